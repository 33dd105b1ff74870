// tools/model_dflt.cpp -- host model of the "default row" step.
//
// Idea under test: most transitions of the hot states of a scanner lead back to one rest state d (the start
// state on random text).  If a per-lane filter over the input byte says "this byte cannot lead anywhere but
// d", the lane needs no table read at all for that step: its next state is d.  Only lanes whose byte passes
// the filter read the table, so a warp-wide LDS has few active lanes and few bank conflicts.
//
// Hot states are grouped by the most frequent target of their row (nd); a group's filter is the union of the
// (hashed) bytes on which any of its states goes somewhere else than the group's rest state.  Transitions
// between groups sink (replay at the end of the 16-byte chunk), like transitions into cold states.
//
//   model_dflt <scanner.img> <n_strings> <string_len> [plant ...]
// env: STRIDE (292), EVICT (visit share below which a state may be evicted from its group, 0 = never)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <string>
#include <vector>

#include "../pire_b200/csrc/dfa_tables.hpp"
#include "../pire_b200/csrc/pire_image.hpp"
#include "../pire_b200/csrc/synth.h"

using namespace pire_b200;

struct Hash {
    std::string name;
    uint32_t slots;
    std::function<uint32_t(uint32_t)> slot;
};

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: model_dflt image n len [plants...]\n");
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
    const uint32_t stride = std::getenv("STRIDE") ? std::atoi(std::getenv("STRIDE")) : 292;
    const double evict = std::getenv("EVICT") ? std::atof(std::getenv("EVICT")) : 0.0;

    SynthParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.seed = 42;
    sp.n_strings = n;
    sp.string_len = len;
    sp.plant_every = 8;
    std::string packed;
    for (int i = 4; i < argc && sp.n_plants < (uint32_t) kMaxPlants; ++i) {
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
    uint64_t total_visits = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = start;
        for (uint32_t k = 0; k < len; ++k) {
            ++visits[s];
            ++total_visits;
            s = dfa.Next(s, corpus[i * len + k]);
        }
    }
    std::vector<uint32_t> order = HotOrderFromCounts(dfa, visits);
    ScanTables t;
    BuildScanTables(dfa, order, kMaxHot, &t);
    const uint32_t H = t.hot;
    auto T = [&](uint32_t g, uint32_t b) -> uint32_t { return t.hot8[(size_t) g * kHotStride + b]; };

    // most frequent target of every hot row
    std::vector<uint32_t> nd(H + 1, H);
    for (uint32_t g = 0; g < H; ++g) {
        uint32_t cnt[257] = {0};
        for (uint32_t b = 0; b < 256; ++b)
            ++cnt[T(g, b)];
        uint32_t best = 0;
        for (uint32_t v = 0; v <= H; ++v)
            if (cnt[v] > cnt[best])
                best = v;
        nd[g] = best;
    }
    // groups: key = rest state d with nd(d) == d
    const uint32_t kNone = 0xffffffffu;
    std::vector<uint32_t> group(H + 1, kNone);
    for (uint32_t g = 0; g < H; ++g) {
        uint32_t d = nd[g];
        if (d < H && nd[d] == d)
            group[g] = d;
    }
    auto share = [&](uint32_t g) { return (double) visits[t.old_of_new[g]] / (double) total_visits; };
    // deviation bytes
    auto deviates = [&](uint32_t g, uint32_t b) {
        uint32_t to = T(g, b);
        return to != group[g] || (to < H && group[to] != group[g]);
    };
    if (evict > 0) {
        // a rarely visited state that would add many deviation bytes to its group's filter leaves the group
        for (uint32_t g = 0; g < H; ++g) {
            if (group[g] == kNone || group[g] == g)
                continue;
            if (share(g) < evict)
                group[g] = kNone;
        }
    }
    {
        std::vector<uint32_t> keys;
        for (uint32_t g = 0; g < H; ++g)
            if (group[g] == g)
                keys.push_back(g);
        std::printf("H %u; rest states (groups): %zu\n", H, keys.size());
        for (uint32_t d : keys) {
            double sh = 0;
            uint32_t members = 0;
            bool dev[256] = {false};
            for (uint32_t g = 0; g < H; ++g)
                if (group[g] == d) {
                    ++members;
                    sh += share(g);
                    for (uint32_t b = 0; b < 256; ++b)
                        dev[b] = dev[b] || deviates(g, b);
                }
            std::string s;
            int cntp = 0, cnta = 0;
            for (uint32_t b = 0; b < 256; ++b)
                if (dev[b]) {
                    ++cnta;
                    if (b >= 0x20 && b < 0x7f) {
                        s += (char) b;
                        ++cntp;
                    }
                }
            if (sh > 0.0005)
                std::printf("  group d=%u: %u members, visit share %.4f, deviation bytes %d (%d printable) [%s]\n", d, members, sh, cnta,
                            cntp, s.c_str());
        }
        double none = 0;
        for (uint32_t g = 0; g < H; ++g)
            if (group[g] == kNone)
                none += share(g);
        std::printf("  ungrouped hot states: visit share %.5f\n", none);
    }

    // hash families
    std::vector<Hash> hashes;
    hashes.push_back({"exact256", 256, [](uint32_t b) { return b; }});
    for (int s = 0; s < 4; ++s)
        hashes.push_back({"b>>" + std::to_string(s) + "&31", 32, [s](uint32_t b) { return (b >> s) & 31u; }});
    hashes.push_back({"b&63", 64, [](uint32_t b) { return b & 63u; }});
    hashes.push_back({"(b>>1)&63", 64, [](uint32_t b) { return (b >> 1) & 63u; }});
    // multiplicative: slot = hi32((b + c) * K) & 31; search K for the lowest pass rate of group-0's filter over printable bytes
    {
        // deviation set of the biggest group
        uint32_t big = 0;
        double bigsh = -1;
        for (uint32_t d = 0; d < H; ++d)
            if (group[d] == d) {
                double sh = 0;
                for (uint32_t g = 0; g < H; ++g)
                    if (group[g] == d)
                        sh += share(g);
                if (sh > bigsh) {
                    bigsh = sh;
                    big = d;
                }
            }
        bool dev[256] = {false};
        for (uint32_t g = 0; g < H; ++g)
            if (group[g] == big)
                for (uint32_t b = 0; b < 256; ++b)
                    dev[b] = dev[b] || deviates(g, b);
        for (uint32_t slots : {32u, 64u}) {
            uint32_t bestK = 0;
            int bestpass = 1000;
            uint64_t x = 88172645463325252ull;
            for (int it = 0; it < 4000000; ++it) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                uint32_t K = (uint32_t) x;
                uint64_t dirty = 0;
                for (uint32_t b = 0; b < 256; ++b)
                    if (dev[b])
                        dirty |= 1ull << ((uint32_t) (((uint64_t) b * K) >> 32) & (slots - 1));
                int pass = 0;
                for (uint32_t b = 0x20; b < 0x7f; ++b)
                    pass += (dirty >> ((uint32_t) (((uint64_t) b * K) >> 32) & (slots - 1))) & 1;
                if (pass < bestpass) {
                    bestpass = pass;
                    bestK = K;
                }
            }
            std::printf("multiplicative %u slots: best K %08x passes %d of 95 printable bytes\n", slots, bestK, bestpass);
            const uint32_t K = bestK, sl = slots;
            hashes.push_back({"mulhi" + std::to_string(slots), slots, [K, sl](uint32_t b) { return (uint32_t) (((uint64_t) b * K) >> 32) & (sl - 1); }});
        }
    }

    for (const Hash& h : hashes) {
        // filters per group (key d) and, for the current kernel, of hot id 0 alone
        std::vector<std::vector<uint8_t>> filt(H + 1, std::vector<uint8_t>(h.slots, 0));
        for (uint32_t g = 0; g < H; ++g)
            if (group[g] != kNone)
                for (uint32_t b = 0; b < 256; ++b)
                    if (deviates(g, b))
                        filt[group[g]][h.slot(b)] = 1;
        std::vector<uint8_t> filt0(h.slots, 0);
        for (uint32_t b = 0; b < 256; ++b)
            if (T(0, b) != 0)
                filt0[h.slot(b)] = 1;

        uint64_t steps = 0, wf_dflt = 0, wf_pred = 0, act_dflt = 0, act_pred = 0, sink_lane_chunks = 0, sink_warp_chunks = 0, chunks = 0;
        uint64_t hist[8] = {0};
        for (uint64_t base = 0; base + 32 <= n; base += 32) {
            uint32_t full[32], g[32];
            bool sunk[32];
            for (int l = 0; l < 32; ++l) {
                full[l] = t.start[1];
                g[l] = full[l] < H ? full[l] : H;
                sunk[l] = g[l] == H;
            }
            for (uint32_t c = 0; c < len; c += 16) {
                for (int l = 0; l < 32; ++l)
                    sunk[l] = g[l] == H;       // cold lanes idle in the sink row for the whole chunk
                for (uint32_t k = 0; k < 16; ++k) {
                    uint32_t words[2][32][4];
                    int cnt[2][32];
                    std::memset(cnt, 0, sizeof(cnt));
                    int act[2] = {0, 0};
                    for (int l = 0; l < 32; ++l) {
                        const uint8_t b = corpus[(base + l) * len + c + k];
                        const uint32_t row = sunk[l] ? H : g[l];
                        const uint32_t addr = row * stride + b;
                        const uint32_t word = addr >> 2, bank = word & 31;
                        auto add = [&](int which) {
                            bool seen = false;
                            for (int q = 0; q < cnt[which][bank] && q < 4; ++q)
                                seen = seen || words[which][bank][q] == word;
                            if (!seen) {
                                if (cnt[which][bank] < 4)
                                    words[which][bank][cnt[which][bank]] = word;
                                ++cnt[which][bank];
                            }
                            ++act[which];
                        };
                        // current kernel: only lanes resting in id 0 on a byte outside the filter skip the load
                        if (row != 0 || filt0[h.slot(b)])
                            add(1);
                        // default-row step
                        bool need = sunk[l] || group[row] == kNone || filt[group[row]][h.slot(b)];
                        if (need)
                            add(0);
                        if (!sunk[l]) {
                            uint32_t to = T(g[l], b);
                            if (to == H || group[to] != group[g[l]] || (group[to] == kNone && false))
                                sunk[l] = to == H || group[to] != group[g[l]];
                            if (!sunk[l])
                                g[l] = to;
                        }
                        full[l] = t.wide ? t.full32[(size_t) full[l] * t.letters + t.cls[b]] : t.full16[(size_t) full[l] * t.letters + t.cls[b]];
                    }
                    int w0 = 0, w1 = 0;
                    for (int bnk = 0; bnk < 32; ++bnk) {
                        w0 = std::max(w0, cnt[0][bnk]);
                        w1 = std::max(w1, cnt[1][bnk]);
                    }
                    ++steps;
                    wf_dflt += w0;
                    wf_pred += w1;
                    act_dflt += act[0];
                    act_pred += act[1];
                    ++hist[std::min(w0, 7)];
                }
                ++chunks;
                bool any = false;
                for (int l = 0; l < 32; ++l) {
                    const bool was_cold_all_along = g[l] == H;
                    if (sunk[l] && !was_cold_all_along) {
                        ++sink_lane_chunks;
                        any = true;
                    }
                    if (sunk[l] && was_cold_all_along && full[l] < H) {
                        ++sink_lane_chunks;      // a cold lane is replayed every chunk anyway; count its return
                        any = true;
                    }
                    g[l] = full[l] < H ? full[l] : H;
                }
                sink_warp_chunks += any;
            }
        }
        std::printf("%-12s pred(now): %.3f wf/step %.2f active | dflt: %.3f wf/step %.2f active, hist[0..4]: %.3f %.3f %.3f %.3f %.3f | sinks: %.5f lane-chunks %.4f warp-chunks\n",
                    h.name.c_str(), (double) wf_pred / steps, (double) act_pred / steps, (double) wf_dflt / steps, (double) act_dflt / steps,
                    (double) hist[0] / steps, (double) hist[1] / steps, (double) hist[2] / steps, (double) hist[3] / steps, (double) hist[4] / steps,
                    (double) sink_lane_chunks / (chunks * 32.0), (double) sink_warp_chunks / (double) chunks);
    }
    return 0;
}
