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

// ---- kind 1: mixed-length UTF-8 corpus (BASELINE config 4) --------------------
// Lengths are log-uniform over octaves in [16, 65536), multiples of 4; every
// 4-byte cell is one of: 4 printable ASCII bytes (11/16), two 2-byte Cyrillic
// code points U+0410..U+044F (4/16), one 3-byte code point U+2000..U+2FFF plus
// one ASCII byte (1/16) -- so any cell-aligned slice is valid UTF-8.  Every
// plant_every-th string of at least 32 bytes ends with a mixed-case hit for the
// case-insensitive headline pattern.

PIRE_HD uint32_t SynthMixedLength(uint64_t seed, uint64_t string_index)
{
    uint64_t h = SynthMix(seed ^ SynthMix(string_index * 0xA24BAED4963EE407ull));
    uint32_t k = (uint32_t) (h % 12u);
    uint32_t frac = (uint32_t) (h >> 16) & 0xffffu;
    uint32_t base = 16u << k;
    uint32_t len = base + (uint32_t) (((uint64_t) base * frac) >> 16);
    return len & ~3u;
}


PIRE_HD uint32_t SynthMixedCell(uint64_t seed, uint64_t string_index, uint32_t cell)
{
    uint64_t h = SynthMix(seed ^ SynthMix((string_index << 20) ^ cell ^ 0x5851F42D4C957F2Dull));
    uint32_t kind = (uint32_t) (h & 15u);
    uint32_t r = (uint32_t) (h >> 8);
    uint32_t b[4];
    if (kind < 11) {
        for (int k = 0; k < 4; ++k)
            b[k] = 0x20u + ((((r >> (8 * k)) & 0xffu) * 95u) >> 8);
    } else if (kind < 15) {
        for (int k = 0; k < 2; ++k) {
            uint32_t cp = 0x410u + ((r >> (8 * k)) & 0x3fu);           // U+0410..U+044F
            b[2 * k] = 0xC0u | (cp >> 6);
            b[2 * k + 1] = 0x80u | (cp & 0x3fu);
        }
    } else {
        uint32_t cp = 0x2000u + (r & 0xfffu);                            // U+2000..U+2FFF
        b[0] = 0xE0u | (cp >> 12);
        b[1] = 0x80u | ((cp >> 6) & 0x3fu);
        b[2] = 0x80u | (cp & 0x3fu);
        b[3] = 0x20u + ((((r >> 16) & 0xffu) * 95u) >> 8);
    }
    return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
}

// Cell `cell` of string `string_index` whose length is `len`, plant applied.
PIRE_HD uint32_t SynthMixedCellPlanted(uint64_t seed, uint32_t plant_every, uint64_t string_index, uint32_t len, uint32_t cell)
{
    if (plant_every && string_index % plant_every == 0 && len >= 32 && cell >= len / 4 - 4) {
        // "HeLLo   WoRRRRLD" as four little-endian cells
        const uint32_t plant[4] = {0x4C4C6548u /* HeLL */, 0x2020206Fu /* o    */, 0x52526F57u /* WoRR */, 0x444C5252u /* RRLD */};
        return plant[cell - (len / 4 - 4)];
    }
    return SynthMixedCell(seed, string_index, cell);
}

} // namespace pire_b200
