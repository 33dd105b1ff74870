// tools/analyze.cpp -- host-side model of the scan kernel's shared-memory behaviour.
//
// Replays the kernel's walk (dfa_tables.hpp layout, one string per lane, 32 lanes
// per warp) over a synthetic corpus on the CPU and counts, per warp-wide step,
// the shared-memory wavefronts the LDS.U8 would need: max over the 32 banks of
// the number of distinct 32-bit words addressed by active lanes.  Used to choose
// layouts and variants without spending GPU time.  Not part of the product.
//
//   analyze <scanner.img> <n_strings> <string_len> <static|tuned> [plant ...]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../pire_b200/csrc/dfa_tables.hpp"
#include "../pire_b200/csrc/pire_image.hpp"
#include "../pire_b200/csrc/synth.h"

using namespace pire_b200;

int main(int argc, char** argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: analyze image n len static|tuned [plants...]\n");
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
    const uint32_t len = (uint32_t) std::atoi(argv[3]);
    const bool tuned = std::strcmp(argv[4], "tuned") == 0;

    SynthParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.seed = 42;
    sp.n_strings = n;
    sp.string_len = len;
    sp.plant_every = 8;
    std::string packed;
    for (int i = 5; i < argc && sp.n_plants < (uint32_t) kMaxPlants; ++i) {
        const char* lit = argv[i];
        sp.plant_off[sp.n_plants] = (uint32_t) packed.size();
        sp.plant_mode[sp.n_plants] = lit[0] == '^' ? 1 : lit[0] == '$' ? 2 : 0;
        packed += lit + (sp.plant_mode[sp.n_plants] ? 1 : 0);
        ++sp.n_plants;
    }
    sp.plant_off[sp.n_plants] = (uint32_t) packed.size();

    std::vector<uint8_t> corpus(n * len);
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t* dst = &corpus[i * len];
        for (uint32_t w = 0; w < len / 8; ++w) {
            uint64_t v = SynthWord(sp.seed, i, w, len / 8);
            std::memcpy(dst + w * 8, &v, 8);
        }
        uint32_t off;
        int id = SynthPlant(sp, i, &off);
        if (id >= 0)
            std::memcpy(dst + off, packed.data() + sp.plant_off[id], sp.plant_off[id + 1] - sp.plant_off[id]);
    }

    const uint32_t start = dfa.Next(dfa.initial, kBeginMark);
    std::vector<uint64_t> visits(dfa.states, 0);
    uint64_t matches = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = start;
        for (uint32_t k = 0; k < len; ++k) {
            ++visits[s];
            s = dfa.Next(s, corpus[i * len + k]);
        }
        matches += dfa.Final(dfa.Next(s, kEndMark));
    }
    std::vector<uint32_t> order = tuned ? HotOrderFromCounts(dfa, visits) : StaticHotOrder(dfa);
    if (std::getenv("SNAKE")) {
        // rows whose ids differ by a multiple of 32 share a bank rotation: reverse every other block of 32 ranks
        // so that a rotation class pairs a busy row of one block with a quiet row of the next
        const size_t hot_n = std::min<size_t>(order.size(), kMaxHot);
        for (size_t b = 32; b + 32 <= hot_n; b += 64)
            std::reverse(order.begin() + b, order.begin() + b + 32);
    }
    if (const char* sh = std::getenv("SHUFFLE")) {
        // which id a hot row gets decides its bank rotation (9 * id mod 32): shuffle the ids of the hot rows
        // (id 0 stays: the exit filter is built on it) to see how much the assignment matters
        uint64_t x = std::strtoull(sh, nullptr, 10) * 0x9E3779B97F4A7C15ull + 1;
        const size_t hot_n = std::min<size_t>(order.size(), kMaxHot);
        for (size_t i = hot_n - 1; i > 1; --i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            std::swap(order[i], order[1 + x % i]);
        }
    }
    ScanTables t;
    BuildScanTables(dfa, order, kMaxHot, &t);

    std::printf("states %u letters %u regexps %u hot %u matches %llu/%llu\n", dfa.states, dfa.letters, dfa.regexps, t.hot,
                (unsigned long long) matches, (unsigned long long) n);
    {
        std::vector<std::pair<uint64_t, uint32_t>> top;
        uint64_t total = 0, visited = 0;
        for (uint32_t s = 0; s < dfa.states; ++s) {
            total += visits[s];
            visited += visits[s] != 0;
            top.push_back({visits[s], s});
        }
        std::sort(top.rbegin(), top.rend());
        double cum = 0;
        std::printf("visited states %llu; coverage:", (unsigned long long) visited);
        for (size_t k = 0; k < top.size() && k < 1024; ++k) {
            cum += (double) top[k].first;
            if (k == 0 || k == 3 || k == 15 || k == 63 || k == 127 || k == 254 || k == 511 || k == 1023)
                std::printf(" top%zu=%.5f", k + 1, cum / (double) total);
        }
        std::printf("\n");
        uint32_t s0 = t.old_of_new[0];
        int exits = 0;
        std::string ex;
        for (uint32_t b = 0; b < 256; ++b)
            if (dfa.Next(s0, b) != s0) {
                ++exits;
                if (b >= 0x20 && b < 0x7f)
                    ex += (char) b;
            }
        std::printf("hot id 0 = state %u: %d exit bytes [%s], bitmap %08x (%d slots)\n", s0, exits, ex.c_str(),
                    t.exit_bitmap0, __builtin_popcount(t.exit_bitmap0));
    }

    // Private-row model: P most visited rows are lane-private (never conflict); a lane that
    // needs another row misses.  Report per-lane-step, per-warp-word(4 B) and per-warp-chunk(16 B) miss rates.
    {
        std::vector<uint32_t> rank(dfa.states, UINT32_MAX);
        std::vector<uint32_t> ord = HotOrderFromCounts(dfa, visits);
        for (uint32_t i = 0; i < ord.size(); ++i)
            rank[ord[i]] = i;
        for (uint32_t P : {16u, 24u, 32u, 40u, 48u, 56u, 64u, 72u, 96u, 128u}) {
            uint64_t lane_steps = 0, lane_miss = 0, words = 0, word_miss = 0, chunks16 = 0, chunk_miss = 0, nonascii = 0;
            for (uint64_t base = 0; base + 32 <= n; base += 32) {
                uint32_t st[32];
                for (int l = 0; l < 32; ++l)
                    st[l] = start;
                bool cm = false;
                for (uint32_t k = 0; k < len; ++k) {
                    bool wm = false;
                    for (int l = 0; l < 32; ++l) {
                        uint8_t b = corpus[(base + l) * len + k];
                        ++lane_steps;
                        bool miss = rank[st[l]] >= P || b >= 128;
                        nonascii += b >= 128;
                        lane_miss += miss;
                        wm = wm || miss;
                        st[l] = dfa.Next(st[l], b);
                    }
                    static bool wacc = false;
                    wacc = wacc || wm;
                    cm = cm || wm;
                    if ((k & 3) == 3) { ++words; word_miss += wacc; wacc = false; }
                    if ((k & 15) == 15) { ++chunks16; chunk_miss += cm; cm = false; }
                }
            }
            std::printf("private P=%3u: lane-step miss %.5f  warp-word(4B) miss %.4f  warp-chunk(16B) miss %.4f\n", P,
                        (double) lane_miss / lane_steps, (double) word_miss / words, (double) chunk_miss / chunks16);
        }
    }

    // Warp model.
    const uint32_t H = t.hot;
    const char* fm = std::getenv("FILTER");
    const int filter_mode = fm ? std::atoi(fm) : 0;      // -1 exact, k = slot (b >> k) & 31
    uint32_t bitmap = 0;
    for (uint32_t b = 0; b < 256; ++b)
        if (t.hot8[b] != 0 && filter_mode >= 0 && filter_mode < 8)
            bitmap |= 1u << ((b >> filter_mode) & 31);
    uint64_t bitmap64 = 0;
    for (uint32_t b = 0; b < 256; ++b)
        if (t.hot8[b] != 0)
            bitmap64 |= 1ull << (b & 63);
    std::printf("filter mode %d bitmap %08x\n", filter_mode, bitmap);
    const char* rotenv = std::getenv("ROT");
    const int rot = rotenv ? std::atoi(rotenv) : -1;
    const int rowrot = std::getenv("ROWROT") ? std::atoi(std::getenv("ROWROT")) : (int) (kHotStride - 256) / 4;      // >= 0: lanes 16..31 use a second table copy rotated by `rot` banks
    uint64_t steps = 0, wf_plain = 0, wf_pred = 0, active_pred = 0, lds_pred = 0;
    uint64_t chunks = 0, replay_lane_chunks = 0, replay_warp_chunks = 0;
    uint64_t hist_plain[8] = {0}, hist_pred[8] = {0};
    for (uint64_t base = 0; base + 32 <= n; base += 32) {
        uint32_t g[32], full[32];
        for (int l = 0; l < 32; ++l) {
            full[l] = t.start[1];
            g[l] = full[l] < H ? full[l] : H;
        }
        for (uint32_t c = 0; c < len; c += 16) {
            uint32_t g_before[32], full_before[32];
            std::memcpy(g_before, g, sizeof(g));
            std::memcpy(full_before, full, sizeof(full));
            for (uint32_t k = 0; k < 16; ++k) {
                uint32_t words_plain[32][4], words_pred[32][4];
                int cnt_plain[32] = {0}, cnt_pred[32] = {0};
                int act = 0;
                for (int l = 0; l < 32; ++l) {
                    uint8_t b = corpus[(base + l) * len + c + k];
                    uint32_t idx = (g[l] << 8) | b;
                    uint32_t word = idx >> 2, bank = (word + g[l] * (uint32_t) rowrot) & 31;    // ROWROT: row stride 256 + 4 * rowrot bytes
                    if (rot >= 0 && l >= 16) {
                        bank = (bank + rot) & 31;
                        word |= 0x40000000u;            // a different copy: never the same word as copy A
                    }
                    auto add = [&](uint32_t (*words)[4], int* cnt) {
                        bool seen = false;
                        for (int q = 0; q < cnt[bank] && q < 4; ++q)
                            seen = seen || words[bank][q] == word;
                        if (!seen) {
                            if (cnt[bank] < 4)
                                words[bank][cnt[bank]] = word;
                            ++cnt[bank];
                        }
                    };
                    add(words_plain, cnt_plain);
                    bool need;
                    if (filter_mode < 0)
                        need = g[l] != 0 || t.hot8[b] != 0;                       // exact
                    else if (filter_mode == 64)
                        need = g[l] != 0 || ((bitmap64 >> (b & 63)) & 1);
                    else
                        need = g[l] != 0 || ((bitmap >> ((b >> filter_mode) & 31)) & 1);
                    if (need) {
                        add(words_pred, cnt_pred);
                        ++act;
                    }
                    g[l] = t.hot8[(size_t) g[l] * kHotStride + b];
                }
                int wp = 0, wq = 0;
                for (int bnk = 0; bnk < 32; ++bnk) {
                    wp = std::max(wp, cnt_plain[bnk]);
                    wq = std::max(wq, cnt_pred[bnk]);
                }
                ++steps;
                wf_plain += wp;
                wf_pred += wq;
                lds_pred += act != 0;
                active_pred += act;
                ++hist_plain[std::min(wp, 7)];
                ++hist_pred[std::min(wq, 7)];
            }
            ++chunks;
            bool any = false;
            for (int l = 0; l < 32; ++l) {
                if (g[l] == H) {
                    any = true;
                    ++replay_lane_chunks;
                    uint32_t s = full_before[l];
                    for (uint32_t k = 0; k < 16; ++k)
                        s = t.wide ? t.full32[(size_t) s * t.letters + t.cls[corpus[(base + l) * len + c + k]]]
                                   : t.full16[(size_t) s * t.letters + t.cls[corpus[(base + l) * len + c + k]]];
                    full[l] = s;
                    g[l] = s < H ? s : H;
                } else {
                    full[l] = g[l];
                }
            }
            (void) g_before;
            replay_warp_chunks += any;
        }
    }
    std::printf("warp-steps %llu\n", (unsigned long long) steps);
    std::printf("plain: %.3f wavefronts/step  hist[1..6+]:", (double) wf_plain / steps);
    for (int i = 1; i < 8; ++i)
        std::printf(" %.3f", (double) hist_plain[i] / steps);
    std::printf("\npred : %.3f wavefronts/step, %.2f active lanes/step, LDS issued on %.3f of steps  hist[0..6+]:",
                (double) wf_pred / steps, (double) active_pred / steps, (double) lds_pred / steps);
    for (int i = 0; i < 8; ++i)
        std::printf(" %.3f", (double) hist_pred[i] / steps);
    std::printf("\nreplays: %.5f of lane-chunks, %.5f of warp-chunks\n", (double) replay_lane_chunks / (chunks * 32.0),
                (double) replay_warp_chunks / (double) chunks);
    return 0;
}
