// sharded_check.cpp -- the multi-GPU entry of the C ABI driven from plain C++ (no Python, no PyTorch): what a
// C++ Pire user (tools/bench/bench.cpp:241-254, samples/pigrep/pigrep.cpp:38-45) would write.
//
//   sharded_check <scanner.pire> <n_strings> <world>
//
// `world` threads, one GPU each (world = 1 runs on one GPU and still goes through NCCL's one-rank communicator).
// Every rank scans its shard of a synthetic corpus with pire_gpu_run_sharded and gathers the bitmap; the result must
// equal, bit for bit, a single-GPU pire_gpu_run_batch over the whole corpus, and the accept sets gathered from the
// state indices must equal the 32-bit accept masks of the scan.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <thread>
#include <vector>

#include "pire_gpu.hpp"

namespace {

const char kPlants[] = "$ABCDEFGHIJKLMNOPQRSTUVWXYZ\0$XABCDEFGHIJKLMNOPQRSTUVWXYZ\0$ABCDEFGHIJKLMNOPQRSTUVWXYZ\0$(555) 123-4567\0"
                       "$hello \t world\0error\0fatal\0https://\0^GET \0$timeout\0";

pire_gpu_synth Spec(uint64_t first, uint64_t n)
{
    pire_gpu_synth s;
    std::memset(&s, 0, sizeof(s));
    s.seed = 42;
    s.first_string = first;
    s.n_strings = n;
    s.string_len = 1024;
    s.plant_every = 8;
    s.n_plants = 10;
    s.plants = kPlants;
    s.plants_bytes = sizeof(kPlants);
    return s;
}

#define CU(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            std::fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(e__));             \
            std::exit(2);                                                                 \
        }                                                                                 \
    } while (0)

struct RankResult {
    std::vector<uint32_t> bits_all;
    long mismatches = 0;
};

void RunRank(const std::vector<char>& image, uint64_t n_global, int world, int rank, const unsigned char* id, RankResult* out)
{
    using namespace Pire::Gpu;
    CU(cudaSetDevice(rank));
    Scanner sc(image.data(), image.size(), rank);
    Comm comm(id, world, rank, rank);
    const std::pair<uint64_t, uint64_t> b = comm.Bounds(n_global);
    const uint64_t n_local = b.second - b.first;
    uint8_t* d_corpus = nullptr;
    CU(cudaMalloc(&d_corpus, n_local * 1024 + 64));
    pire_gpu_synth spec = Spec(b.first, n_local);
    if (n_local)
        Check(pire_gpu_synth_fill_device(&spec, d_corpus, rank, nullptr), "synth");
    uint32_t *d_bits = nullptr, *d_masks = nullptr, *d_states = nullptr, *d_sets = nullptr;
    const uint64_t words = comm.Words(n_global);
    CU(cudaMalloc(&d_bits, words * 4 + 4));
    CU(cudaMemset(d_bits, 0xff, words * 4 + 4));         // the call must overwrite every word
    CU(cudaMalloc(&d_masks, n_local * 4 + 4));
    CU(cudaMalloc(&d_states, n_local * 4 + 4));
    const uint32_t aw = AcceptWords(sc);
    CU(cudaMalloc(&d_sets, n_local * 4 * aw + 4));
    Batch shard{d_corpus, nullptr, 1024, n_local};
    comm.RunSharded(sc, shard, n_global, PIRE_GPU_RUN_BEGIN | PIRE_GPU_RUN_END, d_bits, d_masks, d_states, nullptr);
    AcceptSets(sc, d_states, n_local, d_sets, nullptr);
    CU(cudaDeviceSynchronize());
    out->bits_all.resize(words);
    CU(cudaMemcpy(out->bits_all.data(), d_bits, words * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> masks(n_local), sets((size_t) n_local * aw);
    CU(cudaMemcpy(masks.data(), d_masks, n_local * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sets.data(), d_sets, n_local * 4 * aw, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < n_local; ++i)
        if (sets[i * aw] != masks[i])
            ++out->mismatches;
    // this rank's slot against its own masks: a string matches iff its accept mask is not empty
    const uint64_t words_per = words / world;
    for (uint64_t i = 0; i < n_local; ++i) {
        const uint32_t bit = (out->bits_all[rank * words_per + i / 32] >> (i % 32)) & 1u;
        if (bit != (masks[i] != 0))
            ++out->mismatches;
    }
    cudaFree(d_corpus);
    cudaFree(d_bits);
    cudaFree(d_masks);
    cudaFree(d_states);
    cudaFree(d_sets);
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: sharded_check scanner.pire n_strings world\n");
        return 2;
    }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> image((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const uint64_t n = std::strtoull(argv[2], nullptr, 10);
    const int world = std::atoi(argv[3]);
    try {
        unsigned char id[PIRE_GPU_COMM_ID_BYTES];
        Pire::Gpu::Comm::MakeId(id);
        std::vector<RankResult> res(world);
        std::vector<std::thread> threads;
        for (int r = 0; r < world; ++r)
            threads.emplace_back(RunRank, std::cref(image), n, world, r, id, &res[r]);
        for (std::thread& t : threads)
            t.join();

        // the whole corpus on GPU 0 through the single-GPU entry
        using namespace Pire::Gpu;
        CU(cudaSetDevice(0));
        Scanner sc(image.data(), image.size(), 0);
        uint8_t* d_corpus = nullptr;
        uint32_t* d_bits = nullptr;
        CU(cudaMalloc(&d_corpus, n * 1024 + 64));
        CU(cudaMalloc(&d_bits, (n + 31) / 32 * 4 + 4));
        pire_gpu_synth spec = Spec(0, n);
        if (n)
            Check(pire_gpu_synth_fill_device(&spec, d_corpus, 0, nullptr), "synth");
        Batch all{d_corpus, nullptr, 1024, n};
        Runner(sc).Begin().Run(all).End().Launch(d_bits, nullptr, nullptr, nullptr);
        CU(cudaDeviceSynchronize());
        std::vector<uint32_t> want((n + 31) / 32);
        CU(cudaMemcpy(want.data(), d_bits, want.size() * 4, cudaMemcpyDeviceToHost));
        long mismatches = 0, matches = 0;
        for (int r = 0; r < world; ++r) {
            mismatches += res[r].mismatches;
            if (res[r].bits_all != res[0].bits_all)
                ++mismatches;                                   // identical on every rank
        }
        for (size_t w = 0; w < res[0].bits_all.size(); ++w) {
            const uint32_t expect = w < want.size() ? want[w] : 0u;     // words past n are zero
            if (res[0].bits_all[w] != expect)
                ++mismatches;
            matches += __builtin_popcount(res[0].bits_all[w]);
        }
        if (matches < (long) (n / 8))
            ++mismatches;                                       // every planted string is reported
        std::printf("world %d strings %llu matches %ld: %ld mismatches\n", world, (unsigned long long) n, matches, mismatches);
        return mismatches == 0 ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
