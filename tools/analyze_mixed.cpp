// tools/analyze_mixed.cpp -- the bank-conflict model of tools/analyze.cpp for the mixed UTF-8 text of BASELINE
// configs[3] (synth.h SynthMixedCell): shared-memory wavefronts per warp-wide table read for
//   (A) the shipped layout, entry of (id g, byte b) at g * kHotStride + b, and
//   (B) the same with bit 6 of every byte >= 0x80 flipped before the lookup (b' = b ^ ((b & 0x80) >> 1), two
//       ALU instructions per 4-byte word), which moves the UTF-8 continuation bytes 0x80-0xBF from the banks of
//       0x00-0x3F (space, digits, punctuation) to those of 0x40-0x7F and the lead bytes 0xC0-0xFF the other way.
// Host-only experiment; not part of the product.
//   analyze_mixed <scanner.img> <n_strings> <string_len>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "../pire_b200/csrc/dfa_tables.hpp"
#include "../pire_b200/csrc/pire_image.hpp"
#include "../pire_b200/csrc/synth.h"

using namespace pire_b200;

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: analyze_mixed image n len\n");
        return 2;
    }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> img((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    Dfa dfa;
    std::string err = ParsePireImage(img.data(), img.size(), &dfa);
    if (!err.empty()) {
        std::fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    const uint64_t n = std::strtoull(argv[2], nullptr, 10);
    const uint32_t len = (uint32_t) std::atoi(argv[3]) & ~3u;
    std::vector<uint8_t> corpus(n * len);
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t c = 0; c < len / 4; ++c) {
            uint32_t v = SynthMixedCellPlanted(42, 8, i, len, c);
            std::memcpy(&corpus[i * len + 4 * c], &v, 4);
        }
    // visit counts -> tuned hot order, as pire_gpu_scanner_tune does
    std::vector<uint64_t> visits(dfa.states, 0);
    uint64_t high = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = dfa.Next(dfa.initial, kBeginMark);
        for (uint32_t k = 0; k < len; ++k) {
            ++visits[s];
            high += corpus[i * len + k] >= 0x80;
            s = dfa.Next(s, corpus[i * len + k]);
        }
    }
    ScanTables t;
    BuildScanTables(dfa, HotOrderFromCounts(dfa, visits), kMaxHot, &t);
    std::printf("%u states, hot %u, bytes >= 0x80: %.3f\n", dfa.states, t.hot, (double) high / (double) (n * len));
    const uint32_t H = t.hot;
    for (int variant = 0; variant < 2; ++variant) {
        uint64_t steps = 0, wavefronts = 0;
        for (uint64_t base = 0; base + 32 <= n; base += 32) {
            uint32_t g[32];
            for (int l = 0; l < 32; ++l) {
                uint32_t st = t.start[1];
                g[l] = st < H ? st : H;
            }
            for (uint32_t k = 0; k < len; ++k) {
                uint32_t words[32][8];
                int cnt[32] = {0};
                for (int l = 0; l < 32; ++l) {
                    const uint8_t b = corpus[(base + l) * len + k];
                    const uint32_t pos = variant == 0 ? b : (uint32_t) (b ^ ((b & 0x80) >> 1));
                    const uint32_t word = (g[l] * kHotStride + pos) >> 2, bank = word & 31;
                    bool seen = false;
                    for (int q = 0; q < cnt[bank] && q < 8; ++q)
                        seen = seen || words[bank][q] == word;
                    if (!seen) {
                        if (cnt[bank] < 8)
                            words[bank][cnt[bank]] = word;
                        ++cnt[bank];
                    }
                    g[l] = t.hot8[(size_t) g[l] * kHotStride + b];          // cold lanes stay in the sink: good enough here
                }
                int worst = 0;
                for (int bank = 0; bank < 32; ++bank)
                    worst = std::max(worst, cnt[bank]);
                ++steps;
                wavefronts += worst;
            }
        }
        std::printf("%s: %.3f wavefronts per table read\n", variant == 0 ? "(A) shipped layout        " : "(B) bit 6 of high bytes flipped", (double) wavefronts / (double) steps);
    }
    return 0;
}
