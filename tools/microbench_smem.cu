// tools/microbench_smem.cu -- which input path leaves the shared-memory data pipe to the table walk?
//
// The scan kernels are bound by the SM's L1/shared-memory data pipe: one dependent LDS.U8 per input byte, plus
// whatever it costs to bring 32 bytes per lane into registers.  This bench keeps the walk fixed (a dependent
// LDS.U8 chain over a 255-row table with 292-byte rows, `act` lanes of 32 active, 3 x 512 threads per SM like
// the product kernels) and swaps the input path next to it:
//   0  none (the walk alone)
//   1  one LDG.256 (L1::no_allocate) per lane per 32 steps             -- the uniform kernels today
//   2  one 1 KiB cp.async.bulk global->shared per warp per 32 steps, never read  -- what do TMA writes cost?
//   3  mode 2 + two LDS.128 per lane per 32 steps                        -- TMA-staged input, read by the LSU
//   4  per four warps one tcgen05.cp.128x256b shared->TMEM per 32 steps + one tcgen05.ld.32x32b.x8 per warp
//      (no global traffic)                                              -- what do TMEM fills and reads cost?
//   5  mode 2 + mode 4                                                  -- TMA -> shared -> TMEM -> registers
//   6  one 2-D tensor-map TMA tile (32 strings x 32 bytes, SWIZZLE_32B) per warp per 32 steps + two LDS.128 per lane
//      -- the faithful TMA-staged input: rows are one string length apart in global memory, like mode 1's lanes
// Loaded values are XOR-folded into a sink so nothing is optimised away; the walk's bytes come from an LCG in
// registers so the chain is the same in every mode.  Not part of the product library.
//
//   microbench_smem [steps=32768] [act=19] [only mode]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

constexpr int kBlock = 512;
constexpr int kWarps = kBlock / 32;
constexpr int kGroups = kWarps / 4;
constexpr int kStride = 292;
constexpr int kRows = 96;            // small enough that every mode runs 3 CTAs per SM with the same allocation
constexpr int kTableBytes = kRows * kStride;             // 74 752
constexpr int kStageBytes = kWarps * 1024 * 2;           // two 1 KiB slots per warp: 32 KB
constexpr int kCpBytes = 8192;                           // source region of the tcgen05.cp (4 KiB used)

extern __shared__ __align__(1024) uint8_t smem[];

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(SmemAddr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void Bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(SmemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void Tile(void* dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     SmemAddr(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(SmemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void Ld32(const uint8_t* p, uint4& a, uint4& b)
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ uint4 Lds16(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t Fold(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// shared-memory matrix descriptor (tcgen05): start address, leading / stride byte offsets in 16-byte units, no swizzle
__device__ __forceinline__ uint64_t MakeDesc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3fffu);
    d |= (uint64_t) ((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t) ((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}

template <int kMode>
__global__ void __launch_bounds__(kBlock, 3) Bench(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ table,
                                                   const uint8_t* __restrict__ corpus, uint64_t corpus_bytes, uint32_t steps, uint32_t act,
                                                   uint32_t* __restrict__ out)
{
    uint8_t* hot = smem;
    uint8_t* stage = smem + kTableBytes;
    uint8_t* cpsrc = stage + kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cpsrc + kCpBytes);
    uint64_t* tma_bar = bars;                   // [kWarps][2]
    uint64_t* cp_bar = bars + kWarps * 2;       // [kGroups][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cp_bar + kGroups * 2);

    for (uint32_t i = threadIdx.x * 16; i < kTableBytes; i += kBlock * 16)
        *reinterpret_cast<uint4*>(hot + i) = *reinterpret_cast<const uint4*>(table + i);
    if (kMode >= 4)
        for (uint32_t i = threadIdx.x * 16; i < kCpBytes; i += kBlock * 16)
            *reinterpret_cast<uint4*>(cpsrc + i) = make_uint4(i, i + 1, i + 2, i + 3);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kWarps * 2; ++i)
            MbarInit(&tma_bar[i], 1);
        for (int i = 0; i < kGroups * 2; ++i)
            MbarInit(&cp_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, group = warp >> 2;
    uint32_t tmem_base = 0;
    if (kMode >= 4) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(SmemAddr(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    // generic-proxy writes above (cpsrc) must be visible to the async proxy (tcgen05.cp reads them)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (kMode >= 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_base = *tmem_slot;
    }

    const uint32_t base = SmemAddr(hot);
    const uint64_t gwarp = (uint64_t) blockIdx.x * kWarps + warp;
    const uint8_t* src = corpus + gwarp * 32 * (uint64_t) steps;        // the warp's 32 strings of `steps` bytes each
    const bool active = lane < act;
    uint32_t g = lane % (kRows - 1), sink = 0, rnd = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    uint32_t phase = 0;
    const bool tma = kMode == 2 || kMode == 3 || kMode == 5 || kMode == 6;
    const bool tm = kMode >= 4;
    // TMEM: 16 columns per group: two buffers of 8 columns (32 bytes per lane)
    const uint32_t my_tmem = tmem_base + ((uint32_t) ((warp & 3) * 32) << 16) + group * 16;
    const uint64_t desc = MakeDesc(SmemAddr(cpsrc), 128, 256);

    if (tma && lane == 0) {
        MbarExpectTx(&tma_bar[warp * 2 + 0], 1024);
        if (kMode == 6)
            Tile(stage + (warp * 2 + 0) * 1024, &tmap, 0, (uint32_t) (gwarp * 32), &tma_bar[warp * 2 + 0]);
        else
            Bulk(stage + (warp * 2 + 0) * 1024, src, 1024, &tma_bar[warp * 2 + 0]);
    }
    if (tm && (warp & 3) == 0 && lane == 0) {
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_base + group * 16), "l"(desc) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemAddr(&cp_bar[group * 2 + 0])) : "memory");
    }

    for (uint32_t blk = 0; blk * 32 < steps; ++blk) {
        const uint32_t slot = blk & 1;
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (kMode == 1)
            Ld32(src + (uint64_t) lane * steps + blk * 32, a, b);
        if (tma) {
            // next block's copy goes out before this block is read (its slot was consumed one block ago, by every lane)
            __syncwarp();
            if (lane == 0) {
                MbarExpectTx(&tma_bar[warp * 2 + (slot ^ 1)], 1024);
                const uint32_t nb = (blk + 1) * 32 < steps ? blk + 1 : 0;     // the last block fetches the first again (drained below)
                if (kMode == 6)
                    Tile(stage + (warp * 2 + (slot ^ 1)) * 1024, &tmap, nb * 32, (uint32_t) (gwarp * 32), &tma_bar[warp * 2 + (slot ^ 1)]);
                else
                    Bulk(stage + (warp * 2 + (slot ^ 1)) * 1024, src + (uint64_t) nb * 1024, 1024, &tma_bar[warp * 2 + (slot ^ 1)]);
            }
            MbarWait(&tma_bar[warp * 2 + slot], phase);
            if (kMode == 3) {
                const uint32_t sa = SmemAddr(stage + (warp * 2 + slot) * 1024) + lane * 16;
                a = Lds16(sa);
                b = Lds16(sa + 512);
            }
            if (kMode == 6) {
                // SWIZZLE_32B: the two 16-byte halves of a 32-byte row trade places in rows 4..7 of every eight
                const uint32_t sa = SmemAddr(stage + (warp * 2 + slot) * 1024) + lane * 32;
                const uint32_t x = (lane & 4u) << 2;
                a = Lds16(sa + x);
                b = Lds16(sa + (x ^ 16u));
            }
        }
        if (tm) {
            // all four warps of the group are done with the other TMEM buffer (they read it one block ago)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
            if ((warp & 3) == 0 && lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_base + group * 16 + (slot ^ 1) * 8), "l"(desc) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 SmemAddr(&cp_bar[group * 2 + (slot ^ 1)]))
                             : "memory");
            }
            MbarWait(&cp_bar[group * 2 + slot], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                         : "r"(my_tmem + slot * 8));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (slot)
            phase ^= 1;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            rnd = rnd * 1664525u + 1013904223u;
            const uint32_t bb = base + 0x20u + ((rnd >> 24) * 95u >> 8);
            const uint32_t addr = g * kStride + bb;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.u32 p, %2, 0;\n"
                "@p ld.shared.u8 %0, [%1];\n"
                "}\n"
                : "+r"(g)
                : "r"(addr), "r"((uint32_t) active));
        }
        sink ^= Fold(a) ^ Fold(b);
    }
    if (tma) {
        // drain the copy that is still in flight
        MbarWait(&tma_bar[warp * 2 + (((steps + 31) / 32) & 1)], phase);
    }
    if (tm) {
        MbarWait(&cp_bar[group * 2 + (((steps + 31) / 32) & 1)], phase);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
    }
    out[(uint64_t) blockIdx.x * kBlock + threadIdx.x] = g ^ sink;
}

template <int kMode>
static float Run(const CUtensorMap& tmap, const uint8_t* table, const uint8_t* corpus, uint64_t corpus_bytes, uint32_t steps, uint32_t act,
                 uint32_t* out, int sms)
{
    const size_t sm = kTableBytes + kStageBytes + kCpBytes + 1024;     // the same in every mode: same occupancy
    CK(cudaFuncSetAttribute(Bench<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, Bench<kMode>, kBlock, sm));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = sms * per_sm;
    Bench<kMode><<<grid, kBlock, sm>>>(tmap, table, corpus, corpus_bytes, steps, act, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        Bench<kMode><<<grid, kBlock, sm>>>(tmap, table, corpus, corpus_bytes, steps, act, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double warp_steps = (double) grid * kWarps * steps;
    std::printf("{\"mode\": %d, \"ctas_per_sm\": %d, \"smem\": %zu, \"ms\": %.4f, \"warp_steps_per_us_per_sm\": %.2f, \"input_GBps\": %.1f}\n", kMode,
                per_sm, sm, best, warp_steps / (best * 1e3) / sms, kMode == 0 ? 0.0 : warp_steps * 32 / (best * 1e6));
    std::fflush(stdout);
    return best;
}

int main(int argc, char** argv)
{
    const uint32_t steps = argc > 1 ? (uint32_t) std::atoi(argv[1]) : 32768;
    const uint32_t act = argc > 2 ? (uint32_t) std::atoi(argv[2]) : 19;
    const int only = argc > 3 ? std::atoi(argv[3]) : -1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::vector<uint8_t> t(kTableBytes);
    uint32_t r = 12345;
    for (auto& x : t) {
        r = r * 1664525u + 1013904223u;
        x = (uint8_t) ((r >> 24) % (uint32_t) (kRows - 1));
    }
    uint8_t *d_table, *d_corpus;
    uint32_t* d_out;
    const uint64_t n_strings = (uint64_t) sms * 3 * kWarps * 32;
    const uint64_t corpus_bytes = n_strings * steps;
    CK(cudaMalloc(&d_table, kTableBytes));
    CK(cudaMalloc(&d_corpus, corpus_bytes));
    CK(cudaMalloc(&d_out, (size_t) sms * 4 * kBlock * 4));
    CK(cudaMemcpy(d_table, t.data(), kTableBytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_corpus, 0x41, corpus_bytes));
    std::printf("# %s, %d SMs, steps %u, active lanes %u, corpus %.2f GB\n", prop.name, sms, steps, act, corpus_bytes / 1e9);
    // tensor map: [n_strings rows][steps bytes], box 32 bytes x 32 rows, 32-byte swizzle
    CUtensorMap tmap;
    {
        typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) {
            std::fprintf(stderr, "no cuTensorMapEncodeTiled\n");
            return 1;
        }
        cuuint64_t dims[2] = {steps, n_strings};
        cuuint64_t strides[1] = {steps};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = ((Encode) fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_corpus, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            std::fprintf(stderr, "cuTensorMapEncodeTiled: %d\n", (int) rc);
            return 1;
        }
    }
    if (only < 0 || only == 0) Run<0>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 1) Run<1>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 2) Run<2>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 3) Run<3>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 4) Run<4>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 5) Run<5>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    if (only < 0 || only == 6) Run<6>(tmap, d_table, d_corpus, corpus_bytes, steps, act, d_out, sms);
    return 0;
}
