// synth.h -- deterministic synthetic corpus, identical on host and device.
//
// a counter-based generator keyed by (seed, string index, word
// index) so that a 10 GB corpus never has to cross PCIe and the CPU baseline
// can regenerate any sample of it.  Bytes are printable ASCII 0x20..0x7E; every
// `plant_every`-th string carries one planted literal so that matched and
// absorbing states are exercised (the reference's bench corpus, tools/bench/
// test_file, is prose; there is no canonical synthetic corpus upstream).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PIRE_HD __host__ __device__ __forceinline__
#else
#define PIRE_HD inline
#endif

namespace pire_b200 {

constexpr int kMaxPlants = 32;

struct SynthParams {
    uint64_t seed;
    uint64_t first_string;
    uint64_t n_strings;
    uint32_t string_len;       // multiple of 16
    uint32_t plant_every;      // 0 = never
    uint32_t n_plants;
    uint32_t tail;             // last byte of a randomly planted string (0 = leave)
    uint32_t plant_off[kMaxPlants + 1];   // offsets into plant_bytes
    uint8_t plant_mode[kMaxPlants];       // 0 random offset, 1 string start, 2 string end
};

PIRE_HD uint64_t SynthMix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// 8 printable bytes for (global string index, 8-byte word index).
PIRE_HD uint64_t SynthWord(uint64_t seed, uint64_t string_index, uint32_t word_index, uint32_t words_per_string)
{
    uint64_t r = SynthMix(seed ^ SynthMix(string_index * words_per_string + word_index));
    uint64_t out = 0;
    for (int k = 0; k < 8; ++k) {
        uint32_t b = (uint32_t) (r >> (8 * k)) & 0xffu;
        out |= (uint64_t) (0x20u + ((b * 95u) >> 8)) << (8 * k);
    }
    return out;
}

// Where string `string_index` carries its plant: returns plant id or -1.
// plant_mode[id]: 0 = pseudo-random offset, 1 = at the start of the string (for
// '^'-anchored patterns), 2 = at the end (for '$'-anchored patterns).
PIRE_HD int SynthPlant(const SynthParams& p, uint64_t string_index, uint32_t* offset)
{
    if (p.plant_every == 0 || p.n_plants == 0 || string_index % p.plant_every != 0)
        return -1;
    int id = (int) ((string_index / p.plant_every) % p.n_plants);
    uint32_t len = p.plant_off[id + 1] - p.plant_off[id];
    if (len + 2 > p.string_len)
        return -1;
    if (p.plant_mode[id] == 1)
        *offset = 0;
    else if (p.plant_mode[id] == 2)
        *offset = p.string_len - len;
    else
        *offset = (uint32_t) (SynthMix(p.seed ^ (string_index * 0xD1B54A32D192ED03ull)) % (p.string_len - len - 1));
    return id;
}

// Byte `pos` of string `string_index`.
PIRE_HD uint8_t SynthByte(const SynthParams& p, const char* plant_bytes, uint64_t string_index, uint32_t pos)
{
    uint32_t off = 0;
    int id = SynthPlant(p, string_index, &off);
    if (id >= 0) {
        uint32_t len = p.plant_off[id + 1] - p.plant_off[id];
        if (pos >= off && pos < off + len)
            return (uint8_t) plant_bytes[p.plant_off[id] + (pos - off)];
        if (p.tail && p.plant_mode[id] == 0 && pos == p.string_len - 1)
            return (uint8_t) p.tail;
    }
    uint64_t w = SynthWord(p.seed, string_index, pos / 8, p.string_len / 8);
    return (uint8_t) (w >> (8 * (pos % 8)));
}

} // namespace pire_b200
