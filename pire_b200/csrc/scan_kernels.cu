// scan_kernels.cu -- hand-written sm_100a kernels for Pire's inner scan loop.
//
// Replaces, for a batch of strings, the reference's
//     Runner(sc).Begin().Run(ptr,len).End()   (pire/run.h:365-392)
// whose per-byte body is  state = row(state)[letter_of[byte]]
// (Step run.h:50-57 -> Next/Translate/NextTranslated multi.h:163-192), including
// the ExitMasks early return on NoExit states (multi.h:955-958).
//
// Mapping to the machine (one input string per warp lane):
//   * The table of hot rows (dfa_tables.hpp) is staged once per persistent CTA
//     into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier).
//   * A lane keeps its state as a hot id g in 0..H (H = "not in the table").
//     One step is   bb = IDP.4A(word, 1 << 8k, base)  (input byte k plus the table's
//     256-byte aligned shared address, FMA pipe; PRMT in the prefix kernels; independent
//     of g, so it runs ahead),
//     addr = IMAD(g, 292, bb)  (rows are 292 bytes apart: consecutive rows start nine
//     banks apart, which spreads the conflicts between lanes in different rows),
//     g = LDS.U8 [addr].  No class lookup, no branch.
//   * kPred variant: the LDS is predicated off while the lane sits in hot id 0
//     and the byte cannot leave it (32-slot bitmap probed with a funnel shift),
//     so fewer lanes hit the banks and the load costs fewer wavefronts.
//   * LOOK variant (the glued benchmark scan): the filter looks one byte further -- a resting
//     lane reads the table only if this byte and the next both pass -- in 5.5 instructions per
//     byte (LookProbe / LookStep), two strings per lane (ScanUniformLook2Kernel).
//   * Input bytes: each lane streams its own string: 32-byte read-only loads that
//     bypass L1 allocation, one ahead in a register ping-pong (uniform kernels,
//     LDG.256), or a four-deep cp.async ring of 16-byte chunks in shared memory
//     (CSR kernels).  Long strings of a length-ordered batch are split over a warp
//     (ScanSplitKernel); lines of text are scanned in stream (ScanTextKernel).
//   * Prefix / suffix scans and HalfFinalScanner counting reuse the walk; final hot
//     states carry the highest ids, so a running maximum per chunk says whether any
//     step needs the per-byte work.
//   * A lane whose walk leaves the hot rows reads H from then on (row H is a
//     sink); after the chunk the lane is replayed byte by byte through the
//     complete class-indirect table in global memory / L2.
//   * End(): one 8-byte load of the precomputed per-state report; the match bit
//     of 32 strings is assembled with a warp ballot and written as one word.
// There is no dense contraction anywhere on this path, hence no tensor cores.

#include "scan_kernels.cuh"

#include <atomic>
#include <mutex>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

// One dynamic shared-memory array for every kernel of this file.  The hot rows sit at its start (behind
// the lane-private region in the PRIV kernel, a multiple of 16 KB), 256-byte aligned: FastStep ORs the input
// byte into the table's shared-window address with one PRMT.  The array is declared with 1 KiB alignment so
// that the alignment is a property of the build; StageTables keeps a trap as a backstop.
extern "C" {
extern __shared__ __align__(1024) uint8_t pire_b200_smem[];
}

namespace pire_b200 {

namespace {

constexpr int kBlock = 512;
constexpr int kMinBlocksPerSM = 3;
// The generic kernel keeps 64 + 64 bytes of input per lane in registers and runs at two
// CTAs per SM: ragged batches are bound by the latency of their longest strings (one
// dependent LDS chain per string), which fewer co-resident warps shorten.
constexpr int kGenericBlocksPerSM = 2;
constexpr int kWarpsPerBlock = kBlock / 32;

std::atomic<uint64_t> g_launches{0};

// ---------------------------------------------------------------- PTX helpers

__device__ __forceinline__ uint32_t SmemAddr(const void* p)
{
    return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void FenceBarrierInit()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(SmemAddr(bar)), "r"(parity) : "memory");
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void BulkCopyG2S(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(SmemAddr(bar))
                 : "memory");
}

// Asynchronous 16-byte copy global -> shared (LDGSTS).  Completion is tracked by
// commit/wait groups, not by the register scoreboard, so a pending copy never makes an
// unrelated shared-memory load wait for DRAM (which is what register prefetch did in the
// generic kernel: one exposed DRAM round trip per iteration, r01 experiments).
__device__ __forceinline__ void CopyAsync16(uint32_t dst_shared, const uint8_t* src)
{
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void CopyAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void CopyAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__device__ __forceinline__ uint4 LoadShared16(uint32_t shared_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(shared_addr) : "memory");
    return v;
}

// Streaming read-only load of the corpus (uniform kernels): 32 bytes = one full sector per
// lane per request, never re-used by this SM.
// L2 prefetch of the 128-byte line at p: no destination registers, so nothing tempts the scheduler to delay it.
__device__ __forceinline__ void PrefetchL2(const uint8_t* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// The same load with a 256-byte L2 prefetch: the first touch of a 256-byte region of a string brings all of it
// into L2, so the following seven 32-byte loads of the lane are L2 hits.
__device__ __forceinline__ void LoadStream32P(const uint8_t* p, uint4& a, uint4& b)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ void LoadStream32(const uint8_t* p, uint4& a, uint4& b)
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}

// ---------------------------------------------------------------- shared layout

constexpr int kStageSlots = 4;                                      // 16-byte chunks in flight per lane
constexpr size_t kStageBytes = (size_t) 512 * kStageSlots * 16;      // generic kernel: 32 KB per CTA

struct SharedView {
    uint8_t* priv;       // (priv_rows/4) * 16 KB, PRIV variant only (else empty)
    uint8_t* hot;        // (H+1)*256
    uint16_t* cls;       // 256
    uint8_t* noexit;     // 256 (H+1 used)
    uint64_t* bar;
    uint8_t* stage;      // generic kernel only: cp.async ring, kStageBytes
};

__host__ __device__ inline size_t HotBytes(uint32_t hot) { return (size_t) ((hot + 1 + 3) / 4 * 4) * kHotStride; }   // == HotTableBytes

__host__ __device__ inline size_t PrivBytes(uint32_t priv_rows) { return (size_t) (priv_rows / 4) * 16384; }

__device__ __forceinline__ SharedView CarveShared(uint8_t* smem, uint32_t hot, uint32_t priv_rows = 0)
{
    SharedView v;
    v.priv = smem;
    smem += PrivBytes(priv_rows);
    v.hot = smem;
    v.cls = reinterpret_cast<uint16_t*>(smem + HotBytes(hot));
    v.noexit = smem + HotBytes(hot) + 512;
    v.bar = reinterpret_cast<uint64_t*>(smem + HotBytes(hot) + 512 + 256);
    v.stage = smem + HotBytes(hot) + 512 + 256 + 16;
    return v;
}

// Stage the tables: hot rows by TMA bulk copy, the two small tables by plain loads.
__device__ __forceinline__ void StageTables(const ScanArgs& a, const SharedView& sv, const uint8_t* hot8, uint32_t hot)
{
    const uint32_t total = (uint32_t) HotBytes(hot);
    if ((SmemAddr(sv.hot) & 255u) != 0)
        __trap();                          // FastStep builds base | byte with one PRMT
    if (threadIdx.x == 0) {
        MbarInit(sv.bar, 1);
        FenceBarrierInit();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MbarExpectTx(sv.bar, total);
        constexpr uint32_t kPiece = 16384;
        for (uint32_t off = 0; off < total; off += kPiece) {
            uint32_t n = total - off < kPiece ? total - off : kPiece;
            BulkCopyG2S(sv.hot + off, hot8 + off, n, sv.bar);
        }
    }
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        sv.cls[i] = a.cls[i];
        sv.noexit[i] = i <= hot ? a.noexit[i] : 0;
    }
    MbarWait(sv.bar, 0);
    __syncthreads();
}

// ---------------------------------------------------------------- the walk

struct Tables {
    const uint8_t* hot;
    const uint16_t* cls;
    const void* full;
    uint32_t H;
    uint32_t letters;
    uint32_t wide;
    uint32_t m0;              // 32-slot exit bitmap of hot id 0, slot = byte & 31
    uint32_t base;            // shared-window address of the hot rows; 256-byte aligned (FastStep relies on it)
};

// One byte through the complete table (hot rows first: they are in shared memory).
__device__ __forceinline__ uint32_t SlowStep(const Tables& t, uint32_t s, uint32_t b)
{
    if (s < t.H) {
        uint32_t h = t.hot[s * kHotStride + b];
        if (h != t.H)
            return h;
    }
    size_t at = (size_t) s * t.letters + t.cls[b];
    return t.wide ? __ldg(static_cast<const uint32_t*>(t.full) + at) : (uint32_t) __ldg(static_cast<const uint16_t*>(t.full) + at);
}

// Lane state: g in 0..H; when g == H the real state is `cold`.
struct LaneState {
    uint32_t g;
    uint32_t cold;
};

__device__ __forceinline__ uint32_t FullState(const Tables& t, const LaneState& s) { return s.g == t.H ? s.cold : s.g; }

__device__ __forceinline__ void SetFull(const Tables& t, LaneState& s, uint32_t full)
{
    if (full < t.H) {
        s.g = full;
    } else {
        s.g = t.H;
        s.cold = full;
    }
}

template <bool kPred, bool kIdp = true>
__device__ __forceinline__ void FastStep(const Tables& t, uint32_t& g, uint32_t w, uint32_t sel)
{
    // Entry of (id g, byte b) sits at base + g * kHotStride + b.  `bb` = base | b comes from one PRMT (the
    // base is 256-byte aligned, so its low byte is free) and does not depend on g: the dependent chain of a
    // step is IMAD (FMA pipe) -> LDS, as short as the PRMT -> LDS of an unpadded table.
    // byte `sel` + base: IDP.4A on the FMA pipe (round 2: +0.6 % on the exit-filter scan, +1-2 % on the CSR and counting
    // kernels, which are short of ALU slots), or PRMT on the ALU pipe for the prefix kernels, whose look-ahead pass keeps
    // the FMA side busy (IDP there measured 17 % slower)
    const uint32_t bb = kIdp ? __dp4a(w, 1u << (8 * (sel & 3u)), t.base) : __byte_perm(w, t.base, 0x7650u | (sel & 3u));
    if (kPred) {
        // bit (byte & 31) of the 32-slot exit bitmap: may this byte leave hot id 0?
        // Lanes resting in id 0 on a self-looping byte skip the load (fewer bank
        // conflicts).  Spelled in PTX so that it stays SHF, LOP3 -> predicate,
        // IMAD, @p LDS.  (Sharper filters -- a 64-slot bitmap probed with SHF.R.U64, or a
        // slot of (byte >> 2) & 31 -- pass fewer lanes (2.12 / 2.14 vs 2.23 modelled
        // wavefronts) but measured 8 % slower: a fourth ALU-pipe instruction per byte,
        // issued at half rate, becomes the bound; DESIGN.md results log.)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b32 probe, addr;\n"
            "shf.r.wrap.b32 probe, %2, 0, %1;\n"        // the shift amount is bb & 31 == byte & 31
            "and.b32 probe, probe, 1;\n"
            "or.b32 probe, probe, %0;\n"
            "setp.ne.u32 p, probe, 0;\n"
            "mad.lo.u32 addr, %0, %3, %1;\n"
            "@p ld.shared.u8 %0, [addr];\n"
            "}\n"
            : "+r"(g)
            : "r"(bb), "r"(t.m0), "n"(kHotStride));
    } else {
        const uint32_t addr = g * kHotStride + bb;
        asm("ld.shared.u8 %0, [%1];" : "=r"(g) : "r"(addr));
    }
}

template <bool kPred>
__device__ __forceinline__ void FastWord(const Tables& t, uint32_t& g, uint32_t w)
{
    FastStep<kPred>(t, g, w, 0x5540);
    FastStep<kPred>(t, g, w, 0x5541);
    FastStep<kPred>(t, g, w, 0x5542);
    FastStep<kPred>(t, g, w, 0x5543);
}

// Replay of one 16-byte chunk through the complete table, for a lane that was
// (or fell) outside the hot rows.  Out of line and by value: it is rare, and
// keeping it away from the caller keeps the fast loop free of local memory.
__device__ __noinline__ uint32_t ReplayChunk(const uint8_t* hot, const uint16_t* cls, const void* full, uint32_t H,
                                             uint32_t letters_wide, uint32_t from, uint4 v)
{
    Tables t;
    t.hot = hot;
    t.base = SmemAddr(hot);
    t.cls = cls;
    t.full = full;
    t.H = H;
    t.letters = letters_wide & 0x7fffffffu;
    t.wide = letters_wide >> 31;
    t.m0 = 0;
    uint32_t s = from;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s = SlowStep(t, s, (v.x >> (8 * k)) & 0xffu);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s = SlowStep(t, s, (v.y >> (8 * k)) & 0xffu);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s = SlowStep(t, s, (v.z >> (8 * k)) & 0xffu);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        s = SlowStep(t, s, (v.w >> (8 * k)) & 0xffu);
    return s;
}

// 16 input bytes.
template <bool kPred>
__device__ __forceinline__ void Chunk16(const Tables& t, LaneState& s, uint4 v)
{
    const uint32_t before = s.g;
    uint32_t g = s.g;
    FastWord<kPred>(t, g, v.x);
    FastWord<kPred>(t, g, v.y);
    FastWord<kPred>(t, g, v.z);
    FastWord<kPred>(t, g, v.w);
    s.g = g;
    if (g == t.H) {
        uint32_t from = before == t.H ? s.cold : before;
        uint32_t full = ReplayChunk(t.hot, t.cls, t.full, t.H, t.letters | (t.wide << 31), from, v);
        SetFull(t, s, full);
    }
}

__device__ __forceinline__ void Report(const ScanArgs& a, const Tables& t, const LaneState& s, uint64_t unit, uint64_t i, bool valid)
{
    DeviceFin f = a.fin[FullState(t, s)];
    unsigned matched = __ballot_sync(0xffffffffu, valid && (f.result >> 31));
    if (a.match_bits && (threadIdx.x & 31) == 0)
        a.match_bits[unit] = matched;
    if (valid) {
        if (a.accept_masks)
            a.accept_masks[i] = f.mask;
        if (a.state_idx)
            a.state_idx[i] = f.result & 0x7fffffffu;
    }
}

// Ordered (length-binned) launches: lane -> string is a permutation, so the match
// bit goes to its word with an atomic OR (the caller zeroes the bitmap).
__device__ __forceinline__ void ReportScattered(const ScanArgs& a, const Tables& t, const LaneState& s, uint64_t i, bool valid)
{
    if (!valid)
        return;
    DeviceFin f = a.fin[FullState(t, s)];
    if (a.match_bits && (f.result >> 31))
        atomicOr(&a.match_bits[i >> 5], 1u << (i & 31));
    if (a.accept_masks)
        a.accept_masks[i] = f.mask;
    if (a.state_idx)
        a.state_idx[i] = f.result & 0x7fffffffu;
}

// ---------------------------------------------------------------- kernels

// Uniform batch: fixed length, length % 32 == 0, corpus 32-byte aligned.
template <bool kPred>
__global__ void __launch_bounds__(kBlock, kMinBlocksPerSM) ScanUniformKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0;

    const uint32_t lane = threadIdx.x & 31;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    const uint32_t len = (uint32_t) a.fixed_len;

    for (uint64_t unit = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t i = unit * 32 + lane;
        const bool valid = i < a.n;
        const uint8_t* p = a.corpus + (valid ? i : a.n - 1) * (uint64_t) len;

        LaneState s;
        SetFull(t, s, a.start);

        if (len != 0) {
            // Two register sets in ping-pong: the 32 bytes after the ones being walked are
            // always in flight (software prefetch, one LDG.256 per lane per 32 bytes).
            uint4 a0, a1, b0, b1;
            LoadStream32(p, a0, a1);
            for (uint32_t off = 0;;) {
                off += 32;
                const bool more_b = off < len;
                if (more_b)
                    LoadStream32(p + off, b0, b1);
                Chunk16<kPred>(t, s, a0);
                Chunk16<kPred>(t, s, a1);
                // multi.h:955-958,:979-982: once no byte can leave any lane's state, the
                // rest of the strings cannot change the outcome.
                if (!more_b || __all_sync(0xffffffffu, sv.noexit[s.g] != 0))
                    break;
                off += 32;
                const bool more_a = off < len;
                if (more_a)
                    LoadStream32(p + off, a0, a1);
                Chunk16<kPred>(t, s, b0);
                Chunk16<kPred>(t, s, b1);
                if (!more_a || __all_sync(0xffffffffu, sv.noexit[s.g] != 0))
                    break;
            }
        }
        Report(a, t, s, unit, i, valid);
    }
}

// ---------------------------------------------------------------- LOOK variant
//
// The reference leaves the table walk while a state cannot be left by the bytes ahead (ExitMasks skip loop,
// multi.h:966-989).  The device analogue looks ONE byte further than the round-1 exit filter: a lane resting in
// hot id 0 reads the table only if this byte AND the next one pass the filter F (dfa_tables.hpp: F holds the
// bytes that leave id 0 and the bytes that keep a state entered from id 0 from falling straight back).  If the
// next byte is outside F the lane is back in id 0 after it whatever this byte does, so both reads are skipped
// and the lane never enters the one-step states at all: on random text the lanes outside id 0 drop from 7.4 to
// 5.1 per warp and the active lanes per load from 19 to 12 (host model tools/model_look.cpp: 2.05 -> 1.72
// shared-memory wavefronts per step with the same 32-slot filter; an exact filter would give 1.26).
// A lane that skipped an exit byte is "virtually" in id 0; its true state differs only until the next byte,
// which is outside F and returns it to id 0 on either path, so replays from the register state stay exact.
// The last byte of a string has no successor: it is filtered by F alone.
//
// One step is six instructions, two of them on the dependent chain:
//     bb  = IDP.4A(word, 1 << 8k, base)     byte k + table base, FMA pipe (PRMT in round 1, ALU pipe)
//     pa  = SHF.R.W(F, bb)                  bit (byte & 31) of the filter in bit 0
//     t   = LOP3(pa, pa_next, 1)            both bytes pass
//     p   = LOP3((t | g) != 0)              ... or the lane is outside id 0            [chain]
//     a   = IMAD(g, 292, bb)                                                          [chain]
//     g   = @p LDS.U8 [a]
// k64 = false: 32-slot filter probed with the low five bits of bb (no extra instruction).
// k64 = true : 64-slot filter (slot = byte & 63) probed with SHF.R.U64, which needs the slot alone in a register:
//              one more IDP per byte on the word masked to six bits per byte (FMA pipe, and one LOP3 per word).
//              Printable text folds 3:1 onto 32 slots but only 3:2 onto 64 (model: 1.72 -> 1.42 wavefronts per step).
struct LookFilter {
    uint32_t lo, hi;
    uint32_t zero;      // a kernel argument that is always 0: the addend that keeps the cleaning multiply an IMAD
    uint32_t rev;       // lo with its bits reversed (clean-bit kernels: probes of the odd bytes)
};

// kClean: the probe's bit is moved to bit 31 with the bits below it cleared by one multiply (IMAD, FMA pipe: pa * 2^31
// keeps bit 0 only), so that "both bytes pass, or the lane is outside id 0" is ONE LOP3 with a predicate result over
// (pa, pa_next, g) instead of two: the step keeps six instructions, but two instead of three of them are on the
// half-rate ALU pipe that ncu shows 66 % busy under the look-ahead kernel.
template <bool k64, int kByte, bool kClean = false>
__device__ __forceinline__ void LookProbe(uint32_t w, uint32_t base, const LookFilter& f, uint32_t& bb, uint32_t& pa)
{
    constexpr uint32_t sel = 1u << (8 * kByte);
    bb = __dp4a(w, sel, base);
    if (kClean && (kByte & 1)) {
        // odd bytes: the bit-reversed filter shifted LEFT puts the probe's bit in bit 31 with other filter bits below it.
        // That is good enough: the step ANDs the probes of two neighbouring bytes, one of them is always an even byte,
        // and an even byte's probe is clean -- so an odd byte costs one SHF, an even byte SHF + IMAD: 5.5 instructions
        // per step on average.
        pa = __funnelshift_l(0u, f.rev, bb);
        return;
    }
    if (k64) {
        const uint32_t slot = __dp4a(w & 0x3F3F3F3Fu, sel, 0u);
        pa = (uint32_t) ((((uint64_t) f.hi << 32) | f.lo) >> slot);
    } else {
        pa = __funnelshift_r(f.lo, f.lo, bb);
    }
    if (kClean)
        asm("mad.lo.u32 %0, %1, 0x80000000, %2;" : "=r"(pa) : "r"(pa), "r"(f.zero));
}

template <bool kClean = false>
__device__ __forceinline__ void LookStep(uint32_t& g, uint32_t bb, uint32_t pa, uint32_t pa_next)
{
    if (kClean) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b32 t, addr;\n"
            "lop3.b32 t, %1, %2, %0, 0xEA;\n"          // (pa & pa_next) | g
            "setp.ne.u32 p, t, 0;\n"
            "mad.lo.u32 addr, %0, %4, %3;\n"
            "@p ld.shared.u8 %0, [addr];\n"
            "}\n"
            : "+r"(g)
            : "r"(pa), "r"(pa_next), "r"(bb), "n"(kHotStride));
        return;
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, addr;\n"
        "lop3.b32 t, %1, %2, 1, 0x80;\n"
        "or.b32 t, t, %0;\n"
        "setp.ne.u32 p, t, 0;\n"
        "mad.lo.u32 addr, %0, %4, %3;\n"
        "@p ld.shared.u8 %0, [addr];\n"
        "}\n"
        : "+r"(g)
        : "r"(pa), "r"(pa_next), "r"(bb), "n"(kHotStride));
}

// Four bytes.  (bb0, pa0) belong to byte 0 of `w` and were computed by the previous call; pan is the probe of the
// byte that follows the word.  (Looking ahead from the even bytes only -- half a LOP3 less per byte, 1.91 instead of
// 1.72 wavefronts per step in the model -- measured slower, 2.67 vs 2.63 ms: profiles/r02_experiments_notes.txt.)
template <bool k64, bool kClean = false>
__device__ __forceinline__ void LookWord(uint32_t& g, uint32_t w, uint32_t bb0, uint32_t pa0, uint32_t pan, uint32_t base,
                                         const LookFilter& f)
{
    uint32_t bb1, bb2, bb3, pa1, pa2, pa3;
    LookProbe<k64, 1, kClean>(w, base, f, bb1, pa1);
    LookProbe<k64, 2, kClean>(w, base, f, bb2, pa2);
    LookProbe<k64, 3, kClean>(w, base, f, bb3, pa3);
    LookStep<kClean>(g, bb0, pa0, pa1);
    LookStep<kClean>(g, bb1, pa1, pa2);
    LookStep<kClean>(g, bb2, pa2, pa3);
    LookStep<kClean>(g, bb3, pa3, pan);
}

// Shared-window address of the dynamic shared memory array, as a link-time constant (a cvta of a generic pointer
// costs an S2R + LEA wherever the compiler chooses to rematerialise it).
__device__ __forceinline__ uint32_t SmemWindowBase()
{
    uint32_t v;
    asm("mov.u32 %0, pire_b200_smem;" : "=r"(v));
    return v;
}

// Thirty-two bytes (one LDG.256 per lane).  next0 = the word that follows the block (ignored when !more: the last
// byte of a string is filtered alone).  A lane that left the hot rows reads the sink row from then on; one test
// per block finds it and replays both 16-byte chunks through the complete table.
// Lane state in two registers: g (hot id, H = outside the hot rows) and prev = the complete state the lane had when
// the block began (its cold state while g == H, else g itself) -- exactly what a replay starts from.
// The replay of a whole 32-byte block for the LOOK kernels: out of line, and with every table pointer read from the
// kernel's parameter block in here, so that the walk loop around the (rare) call carries no argument set-up.
__device__ __noinline__ uint32_t ReplayBlock32(const ScanArgs* a, uint32_t from, uint4 v0, uint4 v1)
{
    const SharedView sv = CarveShared(pire_b200_smem, a->hot);
    const uint32_t letters_wide = a->letters | (a->wide << 31);
    const uint32_t mid = ReplayChunk(sv.hot, sv.cls, a->full, a->hot, letters_wide, from, v0);
    return ReplayChunk(sv.hot, sv.cls, a->full, a->hot, letters_wide, mid, v1);
}

template <bool k64, bool kClean = false>
__device__ __forceinline__ void LookBlock32(const Tables& t, uint32_t& g, uint32_t& prev, const uint4& v0, const uint4& v1,
                                            uint32_t next0, bool more, const LookFilter& f, uint32_t opaque_zero, const ScanArgs* args)
{
    prev = g == t.H ? prev : g;
    uint32_t bb, pa, bn, pn;
    LookProbe<k64, 0, kClean>(v0.x, t.base, f, bb, pa);
    LookProbe<k64, 0, kClean>(v0.y, t.base, f, bn, pn);
    LookWord<k64, kClean>(g, v0.x, bb, pa, pn, t.base, f);
    LookProbe<k64, 0, kClean>(v0.z, t.base, f, bb, pa);
    LookWord<k64, kClean>(g, v0.y, bn, pn, pa, t.base, f);
    LookProbe<k64, 0, kClean>(v0.w, t.base, f, bn, pn);
    LookWord<k64, kClean>(g, v0.z, bb, pa, pn, t.base, f);
    LookProbe<k64, 0, kClean>(v1.x, t.base, f, bb, pa);
    LookWord<k64, kClean>(g, v0.w, bn, pn, pa, t.base, f);
    LookProbe<k64, 0, kClean>(v1.y, t.base, f, bn, pn);
    LookWord<k64, kClean>(g, v1.x, bb, pa, pn, t.base, f);
    LookProbe<k64, 0, kClean>(v1.z, t.base, f, bb, pa);
    LookWord<k64, kClean>(g, v1.y, bn, pn, pa, t.base, f);
    LookProbe<k64, 0, kClean>(v1.w, t.base, f, bn, pn);
    LookWord<k64, kClean>(g, v1.z, bb, pa, pn, t.base, f);
    // The word after the block was requested from HBM when this block began: its probe must stay down here (an
    // ordinary intrinsic is hoisted to the top of the block by the compiler, where it waits for the whole DRAM
    // latency -- ncu: 10 % of all stall samples on that one IDP).
    // (a volatile mov is not enough: ptxas schedules across it.  The word is made to depend on the walk itself --
    // plus g times a kernel argument that is always zero -- which costs one IMAD per block.)
    const uint32_t late = next0 + g * opaque_zero;
    LookProbe<k64, 0, kClean>(late, t.base, f, bb, pa);
    LookWord<k64, kClean>(g, v1.w, bn, pn, more ? pa : (kClean ? 0x80000000u : 0xffffffffu), t.base, f);
    if (g == t.H) {
        prev = ReplayBlock32(args, prev, v0, v1);
        g = prev < t.H ? prev : t.H;
    }
}

// Register budget.  The register file is split between the four warp schedulers (16 K registers each), so a
// CTA's warps should be a multiple of four: 512 threads x 3 CTAs leaves 40 registers per thread (12 warps x 1280
// per scheduler), 384 threads x 3 CTAs or 640 threads x 2 CTAs leave 48 (9 / 10 warps x 1536).  The kernel is short
// of independent chains, so the ten-warp shape wins (2.423 ms against 2.480 on the glued scan); both instantiations
// exist (PIRE_B200_LOOK_REGS=40|48, PIRE_B200_LOOK_BLOCK=<threads> for experiments).
constexpr int kLookBlock40 = 512;
constexpr int kLookBlock48 = 640;

template <bool k64, int kRegs, bool kClean = false>
__global__ void __maxnreg__(kRegs) ScanUniformLookKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemWindowBase();          // the hot rows are the first thing in the array (CarveShared, no private region)
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.look_bitmap;
    LookFilter f;
    f.lo = k64 ? (uint32_t) a.look_bitmap64 : a.look_bitmap;
    f.hi = (uint32_t) (a.look_bitmap64 >> 32);
    f.zero = a.opaque_zero;
    f.rev = __brev(f.lo);

    const uint32_t units = (uint32_t) ((a.n + 31) / 32);            // pire_gpu_run_batch keeps n <= 2^40: units below 2^32
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t warps = gridDim.x * warps_per_block;
    const uint32_t len = (uint32_t) a.fixed_len;
    const uint32_t blocks = len >> 5;                               // 32-byte blocks per string (uniform)

    for (uint32_t unit = blockIdx.x * warps_per_block + (threadIdx.x >> 5); unit < units; unit += warps) {
        uint32_t g, prev;
        {
            const uint64_t i = (uint64_t) unit * 32 + (threadIdx.x & 31);
            const uint8_t* p = a.corpus + (i < a.n ? i : a.n - 1) * (uint64_t) len;
            prev = a.start;
            g = a.start < t.H ? a.start : t.H;
            if (blocks != 0) {
                uint4 a0, a1, b0, b1;
                LoadStream32(p, a0, a1);
                for (uint32_t left = blocks;;) {
                    // `left` counts the blocks not yet walked, the one in the a-set included
                    const bool more_b = left > 1;
                    p += 32;
                    if (more_b)
                        LoadStream32(p, b0, b1);
                    __syncwarp();          // see below
                    LookBlock32<k64, kClean>(t, g, prev, a0, a1, b0.x, more_b, f, a.opaque_zero, &a);
                    if (!more_b)
                        break;
                    const bool more_a = left > 2;
                    p += 32;
                    if (more_a)
                        LoadStream32(p, a0, a1);
                    // The warp barrier pins the load HERE.  Left alone, the scheduler sinks the LDG.256 towards its first
                    // use to lend its eight destination registers to the steps in between (SASS: issued 7 steps before
                    // the block's end instead of 32), which exposes most of a DRAM round trip per block (ncu: 14 % of all
                    // stall samples were long-scoreboard waits on the first use of the loaded word).
                    __syncwarp();
                    LookBlock32<k64, kClean>(t, g, prev, b0, b1, a0.x, more_a, f, a.opaque_zero, &a);
                    left -= 2;
                    // multi.h:955-958,:979-982 (NoExit), looked at every 64 bytes here
                    if (!more_a || __all_sync(0xffffffffu, sv.noexit[g] != 0))
                        break;
                }
            }
        }
        const uint64_t i = (uint64_t) unit * 32 + (threadIdx.x & 31);
        LaneState s;
        s.g = t.H;              // Report reads the complete state
        s.cold = g == t.H ? prev : g;
        Report(a, t, s, unit, i, i < a.n);
    }
}

// ---------------------------------------------------------------- LOOK variant, two strings per lane
//
// ncu on the look-ahead kernel (profiles/r02_full_glue10_lookclean.txt): a third of all stall samples sit on the LOP3 that
// waits for the previous step's LDS, and 40 warps per SM run 2.8 % faster than 36 -- the kernel is short of independent
// chains, and registers (48 per thread) cap the warps.  Here every lane walks TWO strings (units 2p and 2p+1 of the
// batch) step by step in turn: the second string's step fills the latency of the first one's table read, and the
// block bookkeeping is shared.  32 data registers (two LDG.256 ping-pong sets), __maxnreg__ chosen by the launch plan.
template <bool kClean>
__device__ __forceinline__ void LookWord2(uint32_t& ga, uint32_t wa, uint32_t bba0, uint32_t paa0, uint32_t pana, uint32_t& gb, uint32_t wb,
                                          uint32_t bbb0, uint32_t pab0, uint32_t panb, uint32_t base, const LookFilter& f)
{
    uint32_t bba1, bba2, bba3, paa1, paa2, paa3, bbb1, bbb2, bbb3, pab1, pab2, pab3;
    LookProbe<false, 1, kClean>(wa, base, f, bba1, paa1);
    LookProbe<false, 1, kClean>(wb, base, f, bbb1, pab1);
    LookStep<kClean>(ga, bba0, paa0, paa1);
    LookStep<kClean>(gb, bbb0, pab0, pab1);
    LookProbe<false, 2, kClean>(wa, base, f, bba2, paa2);
    LookProbe<false, 2, kClean>(wb, base, f, bbb2, pab2);
    LookStep<kClean>(ga, bba1, paa1, paa2);
    LookStep<kClean>(gb, bbb1, pab1, pab2);
    LookProbe<false, 3, kClean>(wa, base, f, bba3, paa3);
    LookProbe<false, 3, kClean>(wb, base, f, bbb3, pab3);
    LookStep<kClean>(ga, bba2, paa2, paa3);
    LookStep<kClean>(gb, bbb2, pab2, pab3);
    LookStep<kClean>(ga, bba3, paa3, pana);
    LookStep<kClean>(gb, bbb3, pab3, panb);
}

template <bool kClean>
__device__ __forceinline__ void LookBlock32x2(const Tables& t, uint32_t& ga, uint32_t& preva, const uint4& a0, const uint4& a1, uint32_t nexta,
                                              uint32_t& gb, uint32_t& prevb, const uint4& b0, const uint4& b1, uint32_t nextb, bool more,
                                              const LookFilter& f, uint32_t opaque_zero, const ScanArgs* args)
{
    preva = ga == t.H ? preva : ga;
    prevb = gb == t.H ? prevb : gb;
    uint32_t bba, paa, bna, pna, bbb, pab, bnb, pnb;
    LookProbe<false, 0, kClean>(a0.x, t.base, f, bba, paa);
    LookProbe<false, 0, kClean>(b0.x, t.base, f, bbb, pab);
    LookProbe<false, 0, kClean>(a0.y, t.base, f, bna, pna);
    LookProbe<false, 0, kClean>(b0.y, t.base, f, bnb, pnb);
    LookWord2<kClean>(ga, a0.x, bba, paa, pna, gb, b0.x, bbb, pab, pnb, t.base, f);
    LookProbe<false, 0, kClean>(a0.z, t.base, f, bba, paa);
    LookProbe<false, 0, kClean>(b0.z, t.base, f, bbb, pab);
    LookWord2<kClean>(ga, a0.y, bna, pna, paa, gb, b0.y, bnb, pnb, pab, t.base, f);
    LookProbe<false, 0, kClean>(a0.w, t.base, f, bna, pna);
    LookProbe<false, 0, kClean>(b0.w, t.base, f, bnb, pnb);
    LookWord2<kClean>(ga, a0.z, bba, paa, pna, gb, b0.z, bbb, pab, pnb, t.base, f);
    LookProbe<false, 0, kClean>(a1.x, t.base, f, bba, paa);
    LookProbe<false, 0, kClean>(b1.x, t.base, f, bbb, pab);
    LookWord2<kClean>(ga, a0.w, bna, pna, paa, gb, b0.w, bnb, pnb, pab, t.base, f);
    LookProbe<false, 0, kClean>(a1.y, t.base, f, bna, pna);
    LookProbe<false, 0, kClean>(b1.y, t.base, f, bnb, pnb);
    LookWord2<kClean>(ga, a1.x, bba, paa, pna, gb, b1.x, bbb, pab, pnb, t.base, f);
    LookProbe<false, 0, kClean>(a1.z, t.base, f, bba, paa);
    LookProbe<false, 0, kClean>(b1.z, t.base, f, bbb, pab);
    LookWord2<kClean>(ga, a1.y, bna, pna, paa, gb, b1.y, bnb, pnb, pab, t.base, f);
    LookProbe<false, 0, kClean>(a1.w, t.base, f, bna, pna);
    LookProbe<false, 0, kClean>(b1.w, t.base, f, bnb, pnb);
    LookWord2<kClean>(ga, a1.z, bba, paa, pna, gb, b1.z, bbb, pab, pnb, t.base, f);
    // the words after the blocks are still on their way from HBM: their probes stay behind the walk (see LookBlock32)
    const uint32_t latea = nexta + ga * opaque_zero;
    const uint32_t lateb = nextb + gb * opaque_zero;
    LookProbe<false, 0, kClean>(latea, t.base, f, bba, paa);
    LookProbe<false, 0, kClean>(lateb, t.base, f, bbb, pab);
    LookWord2<kClean>(ga, a1.w, bna, pna, more ? paa : 0x80000000u, gb, b1.w, bnb, pnb, more ? pab : 0x80000000u, t.base, f);
    if (ga == t.H) {
        preva = ReplayBlock32(args, preva, a0, a1);
        ga = preva < t.H ? preva : t.H;
    }
    if (gb == t.H) {
        prevb = ReplayBlock32(args, prevb, b0, b1);
        gb = prevb < t.H ? prevb : t.H;
    }
}

template <int kRegs>
__global__ void __maxnreg__(kRegs) ScanUniformLook2Kernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemWindowBase();
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.look_bitmap;
    LookFilter f;
    f.lo = a.look_bitmap;
    f.hi = 0;
    f.zero = a.opaque_zero;
    f.rev = __brev(f.lo);

    const uint32_t units = (uint32_t) ((a.n + 31) / 32);
    const uint32_t pairs = (units + 1) / 2;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t warps = gridDim.x * warps_per_block;
    const uint32_t len = (uint32_t) a.fixed_len;
    const uint32_t blocks = len >> 5;

    for (uint32_t pair = blockIdx.x * warps_per_block + (threadIdx.x >> 5); pair < pairs; pair += warps) {
        const bool second = 2 * pair + 1 < units;          // the last pair of an odd batch walks its first unit twice
        uint32_t ga, preva, gb, prevb;
        {
            const uint64_t ia = (uint64_t) pair * 64 + (threadIdx.x & 31);
            const uint64_t ib = ia + (second ? 32 : 0);
            const uint8_t* pa = a.corpus + (ia < a.n ? ia : a.n - 1) * (uint64_t) len;
            const uint8_t* pb = a.corpus + (ib < a.n ? ib : a.n - 1) * (uint64_t) len;
            preva = prevb = a.start;
            ga = gb = a.start < t.H ? a.start : t.H;
            if (blocks != 0) {
                uint4 a0, a1, b0, b1, c0, c1, d0, d1;         // a/c: the first string's ping-pong sets, b/d: the second's
                LoadStream32(pa, a0, a1);
                LoadStream32(pb, b0, b1);
                for (uint32_t left = blocks;;) {
                    const bool more_c = left > 1;
                    pa += 32;
                    pb += 32;
                    if (more_c) {
                        LoadStream32(pa, c0, c1);
                        LoadStream32(pb, d0, d1);
                    }
                    __syncwarp();          // scheduling fence: the loads stay up here (see ScanUniformLookKernel)
                    LookBlock32x2<true>(t, ga, preva, a0, a1, c0.x, gb, prevb, b0, b1, d0.x, more_c, f, a.opaque_zero, &a);
                    if (!more_c)
                        break;
                    const bool more_a = left > 2;
                    pa += 32;
                    pb += 32;
                    if (more_a) {
                        LoadStream32(pa, a0, a1);
                        LoadStream32(pb, b0, b1);
                    }
                    __syncwarp();
                    LookBlock32x2<true>(t, ga, preva, c0, c1, a0.x, gb, prevb, d0, d1, b0.x, more_a, f, a.opaque_zero, &a);
                    left -= 2;
                    if (!more_a || __all_sync(0xffffffffu, (sv.noexit[ga] & sv.noexit[gb]) != 0))
                        break;
                }
            }
        }
        const uint64_t ia = (uint64_t) pair * 64 + (threadIdx.x & 31);
        LaneState s;
        s.g = t.H;
        s.cold = ga == t.H ? preva : ga;
        Report(a, t, s, 2 * (uint64_t) pair, ia, ia < a.n);
        if (second) {
            s.cold = gb == t.H ? prevb : gb;
            Report(a, t, s, 2 * (uint64_t) pair + 1, ia + 32, ia + 32 < a.n);
        }
    }
}

// Edge chunk of a string (its first or last, partially owned 16 bytes): loaded once, then
// handed out byte by byte from registers.
__device__ __forceinline__ uint4 LoadEdge16(const uint8_t* aligned)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(aligned));
    return v;
}

struct EdgeBytes {
    uint64_t lo, hi;
    __device__ __forceinline__ EdgeBytes(uint4 v, uint32_t skip)
    {
        lo = (uint64_t) v.x | ((uint64_t) v.y << 32);
        hi = (uint64_t) v.z | ((uint64_t) v.w << 32);
        if (skip >= 8) {
            lo = hi;
            hi = 0;
            skip -= 8;
        }
        if (skip) {
            lo = (lo >> (8 * skip)) | (hi << (64 - 8 * skip));
            hi >>= 8 * skip;
        }
    }
    __device__ __forceinline__ uint32_t Next()
    {
        uint32_t b = (uint32_t) lo & 0xffu;
        lo = (lo >> 8) | (hi << 56);
        hi >>= 8;
        return b;
    }
    // the remaining bytes as four words, the next byte in bits 0-7 of .x
    __device__ __forceinline__ uint4 Words() const
    {
        return make_uint4((uint32_t) lo, (uint32_t) (lo >> 32), (uint32_t) hi, (uint32_t) (hi >> 32));
    }
};

// The first n (1..16) bytes of `v` through the hot rows: whole words with the fast step, the last one to
// three bytes one by one -- two to three instructions per byte instead of the slow step's compare, branch
// and table choice.  Short strings (lines of text) consist mostly of such edge bytes.  A lane that is, or
// ends up, outside the hot rows replays the bytes through the complete table.
template <bool kPred>
__device__ __forceinline__ void EdgeFast(const Tables& t, LaneState& s, uint4 v, uint32_t n)
{
    const uint32_t before = s.g;
    uint32_t g = before;
    const uint32_t words = n >> 2, rest = n & 3;
    if (words > 0)
        FastWord<kPred>(t, g, v.x);
    if (words > 1)
        FastWord<kPred>(t, g, v.y);
    if (words > 2)
        FastWord<kPred>(t, g, v.z);
    if (words > 3)
        FastWord<kPred>(t, g, v.w);
    const uint32_t last = words == 0 ? v.x : words == 1 ? v.y : words == 2 ? v.z : v.w;
    if (rest > 0)
        FastStep<kPred>(t, g, last, 0x5540);
    if (rest > 1)
        FastStep<kPred>(t, g, last, 0x5541);
    if (rest > 2)
        FastStep<kPred>(t, g, last, 0x5542);
    s.g = g;
    if (g == t.H) {
        uint32_t full = before == t.H ? s.cold : before;
        EdgeBytes eb(v, 0);
        for (uint32_t k = 0; k < n; ++k)
            full = SlowStep(t, full, eb.Next());
        SetFull(t, s, full);
    }
}

// An aligned 16-byte chunk that may stick out of the caller's buffer at either end (the very first and the
// very last chunk of a corpus): bytes outside read as zero.  Out of line, it is rare.
__device__ __noinline__ uint4 LoadEdge16Clipped(const uint8_t* aligned, uintptr_t buf_lo, uintptr_t buf_hi)
{
    uint32_t w[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; ++k) {
        const uintptr_t at = reinterpret_cast<uintptr_t>(aligned) + k;
        if (at >= buf_lo && at < buf_hi)
            w[k >> 2] |= (uint32_t) aligned[k] << (8 * (k & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint4 LoadChunk16(const uint8_t* aligned, uintptr_t buf_lo, uintptr_t buf_hi)
{
    if (reinterpret_cast<uintptr_t>(aligned) >= buf_lo && reinterpret_cast<uintptr_t>(aligned) + 16 <= buf_hi)
        return LoadEdge16(aligned);
    return LoadEdge16Clipped(aligned, buf_lo, buf_hi);
}

// Generic batch: CSR offsets or arbitrary fixed length / alignment.  Head and
// tail bytes (to 16-byte alignment) take the slow step, like run.h:186-226 does
// with its word-aligned body.
// The look-ahead filter of the LOOK variant (see LookStep) inside one 16-byte chunk of a CSR string: byte k reads the
// table only if bytes k and k+1 both pass; the chunk's last byte is filtered alone, so the lane's state is exact at
// every chunk boundary (edges, replays and the NoExit test see true states).
__device__ __forceinline__ void Chunk16Look(const Tables& t, LaneState& s, uint4 v, const LookFilter& f)
{
    const uint32_t before = s.g;
    uint32_t g = s.g;
    uint32_t bb, pa, bn, pn;
    LookProbe<false, 0>(v.x, t.base, f, bb, pa);
    LookProbe<false, 0>(v.y, t.base, f, bn, pn);
    LookWord<false>(g, v.x, bb, pa, pn, t.base, f);
    LookProbe<false, 0>(v.z, t.base, f, bb, pa);
    LookWord<false>(g, v.y, bn, pn, pa, t.base, f);
    LookProbe<false, 0>(v.w, t.base, f, bn, pn);
    LookWord<false>(g, v.z, bb, pa, pn, t.base, f);
    LookWord<false>(g, v.w, bn, pn, 0xffffffffu, t.base, f);
    s.g = g;
    if (g == t.H) {
        uint32_t from = before == t.H ? s.cold : before;
        uint32_t full = ReplayChunk(t.hot, t.cls, t.full, t.H, t.letters | (t.wide << 31), from, v);
        SetFull(t, s, full);
    }
}

// kMode: 0 plain, 1 exit filter (PRED), 2 exit filter with one byte of look-ahead (LOOK) in the 16-byte body chunks
template <int kMode>
__global__ void __launch_bounds__(kBlock, kGenericBlocksPerSM) ScanGenericKernel(const __grid_constant__ ScanArgs a)
{
    constexpr bool kPred = kMode != 0;
    LookFilter look;
    look.lo = a.look_bitmap;
    look.hi = 0;
    look.zero = 0;
    look.rev = 0;
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0;

    const uint32_t lane = threadIdx.x & 31;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    // staging ring: slot j of (warp w, lane l) at ((w * kStageSlots + j) * 32 + l) * 16 -- the 32
    // lanes of a warp read 512 contiguous bytes with one LDS.128 (conflict-free)
    const uint32_t stage = SmemAddr(sv.stage) + (((threadIdx.x >> 5) * kStageSlots) * 32 + lane) * 16;
    // [buf_lo, buf_hi): the caller's corpus buffer; whole aligned chunks may be read inside it only
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + (a.offsets ? a.offsets[a.n] - a.trim : a.n * a.fixed_len);
    const uint64_t split_first = a.split_count ? *a.split_count : 0;

    for (uint64_t unit = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);; unit += warps) {
        if (a.work_counter) {
            // length-binned launch: units are sorted longest first and claimed dynamically
            // (longest-processing-time-first keeps the 64 KiB strings off the tail)
            unsigned int claimed = 0;
            if (lane == 0)
                claimed = atomicAdd(a.work_counter, 1u);
            unit = __shfl_sync(0xffffffffu, claimed, 0);
        }
        if (unit >= units)
            break;
        const uint64_t slot = unit * 32 + lane;
        const bool valid = slot < a.n && slot >= split_first;          // the first split_first entries are the split kernel's
        const uint64_t i = valid && a.order ? a.order[slot] : slot;
        uint64_t b = 0, e = 0;
        if (valid) {
            if (a.offsets) {
                b = a.offsets[i];
                e = a.offsets[i + 1] - a.trim;
                e = e < b ? b : e;       // an empty entry of a trimmed (lines) batch, or caller offsets that step back
            } else {
                b = i * a.fixed_len;
                e = b + a.fixed_len;
            }
        }
        const uint8_t* p = a.corpus + b;
        const uint8_t* end = a.corpus + e;

        LaneState s;
        SetFull(t, s, a.start);

        // head: up to the first 16-byte boundary.  The bytes come from ONE load of the aligned
        // chunk that holds them (the reference's RunChunk does the same with its head word,
        // run.h:129-151) instead of one dependent global load per byte; only a string whose edge
        // chunk would stick out of the corpus buffer reads its edge bytes one by one.
        {
            const uint32_t misalign = (uint32_t) (reinterpret_cast<uintptr_t>(p) & 15);
            if (p < end && misalign != 0) {
                const uint64_t room = (uint64_t) (end - p);
                const uint32_t nhead = room < 16 - misalign ? (uint32_t) room : 16 - misalign;
                const uint8_t* chunk = p - misalign;
                if (reinterpret_cast<uintptr_t>(chunk) >= buf_lo && reinterpret_cast<uintptr_t>(chunk) + 16 <= buf_hi) {
                    EdgeFast<kPred>(t, s, EdgeBytes(LoadEdge16(chunk), misalign).Words(), nhead);
                } else {
                    uint32_t full = FullState(t, s);
                    for (uint32_t k = 0; k < nhead; ++k)
                        full = SlowStep(t, full, p[k]);
                    SetFull(t, s, full);
                }
                p += nhead;
            }
        }
        // body: 16-byte chunks through a four-deep cp.async ring in shared memory (slot
        // c % 4 of this lane holds chunk c); the warp iterates until its longest lane is
        // done, shorter lanes idle (length binning keeps them few).
        const uint32_t chunks = (uint32_t) ((end - p) >> 4);
        bool parked = false;       // lane sits in a NoExit state: its remaining bytes are irrelevant
#pragma unroll
        for (int j = 0; j < kStageSlots; ++j) {
            if ((uint32_t) j < chunks)
                CopyAsync16(stage + j * 512, p + 16 * j);
            CopyAsyncCommit();
        }
        for (uint32_t k = 0; __any_sync(0xffffffffu, k < chunks); k += kStageSlots) {
#pragma unroll
            for (int j = 0; j < kStageSlots; ++j) {
                CopyAsyncWait<kStageSlots - 1>();                 // chunk k + j has landed
                const uint4 v = LoadShared16(stage + j * 512);
                if (k + kStageSlots + j < chunks)
                    CopyAsync16(stage + j * 512, p + 16 * (size_t) (k + kStageSlots + j));
                CopyAsyncCommit();
                if (k + j < chunks) {
                    if (kMode == 2)
                        Chunk16Look(t, s, v, look);
                    else
                        Chunk16<kPred>(t, s, v);
                }
            }
            // multi.h:955-958,:979-982: a NoExit state cannot be left by any byte.
            const bool live = k + kStageSlots < chunks;
            const bool stuck = sv.noexit[s.g] != 0;
            if (__all_sync(0xffffffffu, !live || stuck)) {
                parked = live && stuck;
                break;
            }
        }
        CopyAsyncWait<0>();
        // tail: fewer than 16 bytes, at an aligned address
        if (!parked) {
            p += 16 * (size_t) chunks;
            if (p < end) {
                const uint32_t ntail = (uint32_t) (end - p);
                if (reinterpret_cast<uintptr_t>(p) + 16 <= buf_hi) {
                    EdgeFast<kPred>(t, s, LoadEdge16(p), ntail);
                } else {
                    uint32_t full = FullState(t, s);
                    for (uint32_t k = 0; k < ntail; ++k)
                        full = SlowStep(t, full, p[k]);
                    SetFull(t, s, full);
                }
            }
        }
        if (a.order)
            ReportScattered(a, t, s, i, valid);
        else
            Report(a, t, s, unit, i, valid);
    }
}

// ---------------------------------------------------------------- long strings, split over a warp
//
// One string per lane makes the longest string of a batch the critical path: a 64 KiB string is 65 536 dependent
// table reads, 3.7 ms at the 110 cycles a step takes a warp that shares its SM with 31 others -- as long as the whole
// mixed-length benchmark batch.  Here a WARP owns one long string: lane j walks piece j of 32 (whole 32-byte blocks,
// LDG.256 one block ahead), lane 0 from the string's true state, the others from a guess (hot id 0, the most visited
// state).  Every lane leaves marks -- its hot id after every K blocks, in shared memory.  Then the pieces are stitched:
// lane j's true start is lane j-1's end; a lane whose walk started from another state re-walks its piece from the true
// one until it meets a mark (same hot id at the same position: the rest of the walk is the recorded one), replacing the
// marks it passes.  Two walks over the same bytes fall together within a few bytes for the automata regexps make (a
// byte that continues no match sends every state to the resting state), so a round of re-walks costs about K blocks;
// the loop repeats until every lane's start equals its predecessor's end, which is exact whatever the automaton does:
// lane 0 is exact, and lane j is exact one round after lane j-1 at the latest (31 rounds of a whole piece each in the
// worst case -- the serial walk).  A true start that is a NoExit state needs no walk at all.
constexpr uint32_t kSplitMin = 8192;            // bytes; a bucket boundary of LengthOrder
constexpr uint32_t kSplitMarks = 32;            // marks per lane
constexpr size_t kSplitMarkBytes = (size_t) kBlock * kSplitMarks;

__device__ __forceinline__ uint64_t EndOffset(const ScanArgs& a, uint64_t i)
{
    const uint64_t b = a.offsets[i], e = a.offsets[i + 1] - a.trim;
    return e < b ? b : e;
}

template <bool kPred>
__global__ void __launch_bounds__(kBlock, kMinBlocksPerSM) ScanSplitKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0;

    const uint32_t lane = threadIdx.x & 31;
    // mark m of lane l at marks[m * 32 + l]: the lanes of a warp write one mark with one conflict-free store
    uint8_t* const marks = sv.stage + (threadIdx.x >> 5) * (32 * kSplitMarks) + lane;
    volatile uint64_t* const parked = reinterpret_cast<volatile uint64_t*>(sv.stage + kSplitMarkBytes) + (threadIdx.x >> 5) * 4;
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + (a.offsets[a.n] - a.trim);
    const uint32_t n_long = *a.split_count;

    for (;;) {
        unsigned int slot = 0;
        if (lane == 0)
            slot = atomicAdd(a.split_counter, 1u);         // longest strings first
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n_long)
            break;
        const uint64_t i = a.order[slot];
        const uint8_t* p = a.corpus + a.offsets[i];
        const uint8_t* const end = a.corpus + EndOffset(a, i);

        // head, by every lane alike: to the first 32-byte boundary
        LaneState s;
        SetFull(t, s, a.start);
        {
            const uint32_t mis = (uint32_t) (reinterpret_cast<uintptr_t>(p) & 15);
            if (p < end && mis != 0) {
                const uint64_t room = (uint64_t) (end - p);
                const uint32_t nhead = room < 16 - mis ? (uint32_t) room : 16 - mis;
                EdgeFast<kPred>(t, s, EdgeBytes(LoadChunk16(p - mis, buf_lo, buf_hi), mis).Words(), nhead);
                p += nhead;
            }
            if ((reinterpret_cast<uintptr_t>(p) & 16) && end - p >= 16) {
                Chunk16<kPred>(t, s, LoadEdge16(p));
                p += 16;
            }
        }
        const uint32_t blocks_total = (reinterpret_cast<uintptr_t>(p) & 31) ? 0u : (uint32_t) ((end - p) >> 5);
        const uint32_t base = blocks_total / 32, rem = blocks_total % 32;
        const uint32_t my_blocks = base + (lane < rem ? 1u : 0u);
        const uint32_t my_first = lane * base + (lane < rem ? lane : rem);
        const uint32_t trips = base + (rem ? 1u : 0u);                    // the longest piece
        const uint32_t per_mark = (trips + kSplitMarks - 1) / kSplitMarks;   // K: blocks between two marks (0 only if trips == 0)
        const uint8_t* const piece = p + 32 * (size_t) my_first;

        // the string's bounds wait in shared memory while the pieces are walked: they are the same in every lane and not
        // needed in the loop, which is short of registers (ptxas spilled the loop counter instead)
        if (lane == 0) {
            parked[0] = i;
            parked[1] = reinterpret_cast<uint64_t>(p + 32 * (size_t) blocks_total);
            parked[2] = reinterpret_cast<uint64_t>(end);
        }
        __syncwarp();

        // the pieces, all at once: lane 0 from the string's state, the others from the guess.  Every lane runs the loop
        // `trips` times, so that the scheduling fences behind the loads are reached by the whole warp.  The loads ask for
        // the whole 256-byte line in L2: a piece is contiguous, seven of eight blocks then come from there.
        uint32_t start_full = lane == 0 ? FullState(t, s) : 0u;
        SetFull(t, s, start_full);
        {
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
            if (my_blocks)
                LoadStream32P(piece, a0, a1);
            uint32_t countdown = per_mark, m = 0;
            const uint32_t ahead = a.split_prefetch;
            for (uint32_t k = 0; k < trips; k += 2) {
                // optional L2 prefetch, one per 256-byte line, `ahead` blocks in front of the walk.  ncu puts 31 % of this
                // kernel's stall samples on the first use of a loaded word, but the prefetch measured no gain at 8..64
                // blocks (3.09-3.12 ms with, 3.09 without): the L2::256B hint of the loads already brings the line into
                // L2, what is exposed is the L2 -> SM trip under load.  Off by default.
                if (ahead && (k & 7u) == 0 && k + ahead < my_blocks)
                    PrefetchL2(piece + 32 * (size_t) (k + ahead));
                if (k + 1 < my_blocks)
                    LoadStream32P(piece + 32 * (size_t) (k + 1), b0, b1);
                __syncwarp();              // scheduling fence: the load is issued here, not next to its first use
                if (k < my_blocks) {
                    Chunk16<kPred>(t, s, a0);
                    Chunk16<kPred>(t, s, a1);
                    if (--countdown == 0) {
                        marks[32 * m++] = (uint8_t) s.g;
                        countdown = per_mark;
                    }
                }
                if (k + 2 < my_blocks)
                    LoadStream32P(piece + 32 * (size_t) (k + 2), a0, a1);
                __syncwarp();
                if (k + 1 < my_blocks) {
                    Chunk16<kPred>(t, s, b0);
                    Chunk16<kPred>(t, s, b1);
                    if (--countdown == 0) {
                        marks[32 * m++] = (uint8_t) s.g;
                        countdown = per_mark;
                    }
                }
            }
        }
        uint32_t end_full = FullState(t, s);

        // A lane whose guess was wrong walks the head of its piece again: its first block is asked for now, before the
        // stitch knows who needs it (one 32-byte L2 hit per lane and string), so that the re-walk does not begin with a
        // bare load -- ncu had 31 % of this kernel's stall samples on the first use of a loaded word.
        uint4 head0 = make_uint4(0, 0, 0, 0), head1 = head0;
        bool head_fresh = my_blocks != 0;
        if (head_fresh)
            LoadStream32P(piece, head0, head1);

        // stitch: every lane must have started where its predecessor ended
        for (;;) {
            const uint32_t before = __shfl_up_sync(0xffffffffu, end_full, 1);
            const uint32_t want = lane == 0 ? start_full : before;
            const bool redo = want != start_full;
            if (!__any_sync(0xffffffffu, redo))
                break;
            if (redo) {
                start_full = want;
                const uint32_t n_marks = per_mark ? my_blocks / per_mark : 0u;
                if (want < t.H && sv.noexit[want] != 0) {
                    // multi.h:955-958: no byte leaves this state
                    end_full = want;
                    for (uint32_t m = 0; m < n_marks; ++m)
                        marks[32 * m] = (uint8_t) want;
                } else {
                    // the piece again from the true state, until the walk meets a mark (same hot id at the same place:
                    // from there on it is the recorded walk, and so is its end); marks passed on the way are replaced,
                    // so that they describe this walk afterwards
                    LaneState r;
                    SetFull(t, r, want);
                    uint32_t countdown = per_mark, m = 0;
                    bool met = false;
                    uint4 v0 = head0, v1 = head1;
                    if (!head_fresh && my_blocks)
                        LoadStream32P(piece, v0, v1);
                    head_fresh = false;
                    for (uint32_t k = 0; k < my_blocks; ++k) {
                        uint4 n0 = v0, n1 = v1;
                        if (k + 1 < my_blocks)
                            LoadStream32P(piece + 32 * (size_t) (k + 1), n0, n1);      // one block ahead, like the first walk
                        Chunk16<kPred>(t, r, v0);
                        Chunk16<kPred>(t, r, v1);
                        if (--countdown == 0) {
                            if (r.g != t.H && r.g == marks[32 * m]) {
                                met = true;
                                break;
                            }
                            marks[32 * m++] = (uint8_t) r.g;
                            countdown = per_mark;
                        }
                        v0 = n0;
                        v1 = n1;
                    }
                    if (!met)
                        end_full = FullState(t, r);
                }
            }
        }

        // tail, by every lane alike, from the last piece's end
        SetFull(t, s, __shfl_sync(0xffffffffu, end_full, 31));
        const uint8_t* q = reinterpret_cast<const uint8_t*>(parked[1]);
        const uint8_t* const q_end = reinterpret_cast<const uint8_t*>(parked[2]);
        if (q_end - q >= 16) {
            Chunk16<kPred>(t, s, LoadEdge16(q));
            q += 16;
        }
        if (q < q_end)
            EdgeFast<kPred>(t, s, LoadChunk16(q, buf_lo, buf_hi), (uint32_t) (q_end - q));
        ReportScattered(a, t, s, parked[0], lane == 0);
        __syncwarp();                      // the next string's bounds replace these
    }
}

// n_long = the number of leading entries of `order` whose strings are at least kSplitMin bytes long (LengthOrder puts the
// longest buckets first, so this is all of them; for any other permutation it is just some prefix, and the split kernel
// is exact for strings of any length).  Also resets the split kernel's work counter.
__global__ void SplitCountKernel(const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ order, uint64_t n,
                                 uint32_t split_min, uint32_t* __restrict__ count, unsigned int* __restrict__ counter)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint64_t i = order[mid];
        if (offsets[i + 1] - offsets[i] >= split_min)
            lo = mid + 1;
        else
            hi = mid;
    }
    *count = (uint32_t) lo;
    *counter = 0;
}

// ---------------------------------------------------------------- lines of text
//
// The step before the path for line-oriented input (samples/pigrep/pigrep.cpp:38-45): strings of a few dozen
// bytes of very unequal length (the reference's own benchmark prose: median 21 B, 90 % under 80 B, the longest
// of 32 consecutive lines 115 B on average).  One string per lane in lockstep units wastes three quarters of the
// lanes there -- each unit lasts as long as its longest line.  Here a warp owns 256 consecutive lines and its
// lanes pull them one at a time: every iteration each busy lane walks one aligned 16-byte chunk of its line
// (clipped to the line at both ends, fetched one chunk ahead into registers), and a lane that finishes reports
// its line and takes the next unassigned one (ballot + popcount over a warp-uniform cursor).  No staging ring:
// neighbouring lines share cache lines, so the chunks come from L1/L2.
constexpr uint32_t kLinesPerWarp = 256;
constexpr uint32_t kPiecesPerTurn = 2;      // chunks a busy lane walks before lines are handed out again
constexpr uint32_t kLinesMinIdle = 1;       // lanes that must be waiting before lines are handed out

template <bool kPred>
__global__ void __launch_bounds__(kBlock, kMinBlocksPerSM) ScanLinesKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t below = (1u << lane) - 1u;
    const uint64_t groups = (a.n + kLinesPerWarp - 1) / kLinesPerWarp;
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + (a.offsets[a.n] - a.trim);

    for (uint64_t group = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); group < groups; group += warps) {
        const uint64_t last = (group + 1) * kLinesPerWarp < a.n ? (group + 1) * kLinesPerWarp : a.n;
        uint64_t cursor = group * kLinesPerWarp;             // next unassigned line, the same in every lane
        bool busy = false;
        uint64_t line = 0;
        const uint8_t* chunk = a.corpus;
        uint32_t mis = 0, span = 0, pieces = 0, piece = 0;
        uint4 cur = make_uint4(0, 0, 0, 0);
        LaneState s;
        s.g = 0;
        s.cold = 0;
        for (;;) {
            const unsigned idle = __ballot_sync(0xffffffffu, !busy);
            // lines are handed out when enough lanes wait for one (or nobody is busy): handing out costs the
            // whole warp ~56 instructions however few lanes take part
            if (cursor < last && ((uint32_t) __popc(idle) >= a.lines_min_idle || idle == 0xffffffffu)) {
                const uint64_t mine = cursor + __popc(idle & below);
                cursor += __popc(idle);
                if (!busy && mine < last) {
                    line = mine;
                    const uint64_t b = a.offsets[line];
                    const uint64_t e = a.offsets[line + 1] - a.trim;
                    const uint8_t* p = a.corpus + b;
                    const uint32_t len = e > b ? (uint32_t) (e - b) : 0u;
                    mis = (uint32_t) (reinterpret_cast<uintptr_t>(p) & 15);
                    chunk = p - mis;
                    span = mis + len;
                    pieces = len ? (span + 15) >> 4 : 0;
                    piece = 0;
                    SetFull(t, s, a.start);
                    busy = true;
                    if (pieces)
                        cur = LoadChunk16(chunk, buf_lo, buf_hi);
                }
            }
            if (!__any_sync(0xffffffffu, busy))
                break;
            if (busy) {
                // up to kPiecesPerTurn chunks before the lanes look for new lines again: most lines end within
                // one turn, and the bookkeeping around a turn costs as much as walking two chunks
#pragma unroll 1
                for (uint32_t turn = 0; turn < a.lines_turn && piece < pieces; ++turn) {
                    uint4 next = make_uint4(0, 0, 0, 0);
                    if (piece + 1 < pieces)
                        next = LoadChunk16(chunk + 16 * (size_t) (piece + 1), buf_lo, buf_hi);
                    const uint32_t skip = piece == 0 ? mis : 0;
                    const uint32_t left = span - 16 * piece;
                    const uint32_t upto = left < 16 ? left : 16;
                    EdgeFast<kPred>(t, s, skip ? EdgeBytes(cur, skip).Words() : cur, upto - skip);
                    cur = next;
                    ++piece;
                }
                if (piece >= pieces) {
                    ReportScattered(a, t, s, line, true);
                    busy = false;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- lines of text, in stream
//
// The lines of a newline-delimited text lie back to back, so they can be scanned where they are: the text is cut
// into segments of a.text_segment bytes on 32-byte boundaries of the address space, lane j of a warp walks segment
// 32 * unit + j like the uniform kernel walks a string (LDG.256, one block ahead in registers, every lane busy in
// every step), and owns the lines that START inside its segment -- it runs past the segment's end until the last of
// them is finished.  What makes this cheap:
//   * the copy of the hot rows in shared memory maps '\n' to the start state in every row (the sink row keeps the
//     sink), so the walk restarts by itself behind a line: no branch, no second pass over the rest of a chunk;
//   * the thirty-two states of a block are packed into eight registers as they appear (one IMAD each, FMA pipe), and
//     the newlines of the block are found in its bytes (exact zero-byte test of word ^ 0x0a0a0a0a, compressed to one
//     bit per byte by a multiply): a lane with newlines parks the packed states in shared memory, picks the state in
//     front of each newline with one LDS.U8 and reports it through a copy of the hot states' reports in shared memory
//     -- the steady state reads no offsets and has no dependent global load;
//   * a lane outside the hot rows at the end of a chunk (the sink is absorbing) replays the chunk byte by byte, line
//     ends included; so does a lane over the unaligned start of its first line.
// The offsets must be those of pire_gpu_split_lines for this text: line i + 1 starts right behind the '\n' of line i.
// They are read once per lane and unit, by the binary search for the first line that starts in the segment; from then
// on line numbers just count up.  Bytes outside the text read as '\n': that ends a last line without newline.
constexpr uint32_t kTextSegment = 4096;
constexpr int kTextBlocksPerSM = 2;
constexpr size_t kTextFinBytes = 256 * sizeof(DeviceFin);
constexpr size_t kTextPackBytes = (size_t) kBlock * 32;     // packed states of one 32-byte block per lane

struct TextLane {
    uint32_t line;       // the line being walked
    int32_t left;        // bytes from the chunk being walked to the end of the segment (negative behind it)
    bool active;         // false: no line of this segment is left
};

// which outputs a line writes: bit 0 accept_masks, bit 1 state_idx, bit 2 match_bits (kept in a register: the
// pointers themselves are 64-bit kernel parameters, and testing them costs three instructions each per line)
__device__ __forceinline__ uint32_t TextOutputs(const ScanArgs& a)
{
    uint32_t outs = (a.accept_masks ? 1u : 0u) | (a.state_idx ? 2u : 0u) | (a.match_bits ? 4u : 0u);
    asm volatile("mov.u32 %0, %0;" : "+r"(outs));
    return outs;
}

__device__ __forceinline__ void TextReport(const ScanArgs& a, uint32_t outs, uint32_t line, DeviceFin f)
{
    if ((outs & 4u) && (f.result >> 31))
        atomicOr(&a.match_bits[line >> 5], 1u << (line & 31));
    if (outs & 1u)
        a.accept_masks[line] = f.mask;
    if (outs & 2u)
        a.state_idx[line] = f.result & 0x7fffffffu;
}

// Bytes [from, to) of the block at text position cpos, one at a time from the complete state `full`.
__device__ __forceinline__ void TextSlow(const ScanArgs& a, const Tables& t, const DeviceFin* fin_hot, uint32_t outs, TextLane& c,
                                         LaneState& s, uint32_t full, int64_t cpos, uint32_t from, uint32_t to, uint64_t seg_hi,
                                         uint64_t total)
{
    for (uint32_t j = from; j < to; ++j) {
        const uint64_t at = (uint64_t) (cpos + (int64_t) j);
        const uint32_t b = at < total ? a.corpus[at] : (uint32_t) '\n';
        if (b == '\n') {
            TextReport(a, outs, c.line, full < t.H ? fin_hot[full] : a.fin[full]);
            ++c.line;
            full = a.start;
            if (c.line >= a.n || at + 1 >= seg_hi) {
                c.active = false;
                break;
            }
            continue;
        }
        full = SlowStep(t, full, b);
    }
    SetFull(t, s, full);
}

// Four steps; q collects the state IN FRONT of each byte (byte 3 - i of q for byte i of w).
template <bool kPred>
__device__ __forceinline__ void TextWord(const Tables& t, uint32_t& g, uint32_t w, uint32_t& q)
{
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i) {
        asm("mad.lo.u32 %0, %0, 256, %1;" : "+r"(q) : "r"(g));       // q = q << 8 | g, on the FMA pipe
        FastStep<kPred>(t, g, w, 0x5540u + i);
    }
}

// Bit i of the result: byte i of w is '\n'.  Exact (no borrow between bytes): 0x80 in every byte of x that is zero,
// then the four flags gathered into the top nibble by one multiply (the partial products do not collide).
__device__ __forceinline__ uint32_t NewlineNibble(uint32_t w)
{
    const uint32_t x = w ^ 0x0a0a0a0au;
    const uint32_t flags = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
    return (flags * 0x00204081u) >> 28;
}

// One 32-byte block at text position cpos.  `packs` = this lane's slot in the warp's staging area for packed states:
// word w (0..7) of the block's states at packs + (w / 4) * 512 + (w % 4) * 4, so that the two 16-byte stores of a warp
// are contiguous.
template <bool kPred>
__device__ __forceinline__ void TextBlock(const ScanArgs& a, const Tables& t, const DeviceFin* fin_hot, uint32_t outs, TextLane& c,
                                          LaneState& s, uint4 v0, uint4 v1, uint32_t packs, int64_t cpos, uint64_t seg_hi,
                                          uint64_t total)
{
    const uint32_t before = s.g;
    uint32_t g = before;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    TextWord<kPred>(t, g, v0.x, q0.x);
    TextWord<kPred>(t, g, v0.y, q0.y);
    TextWord<kPred>(t, g, v0.z, q0.z);
    TextWord<kPred>(t, g, v0.w, q0.w);
    TextWord<kPred>(t, g, v1.x, q1.x);
    TextWord<kPred>(t, g, v1.y, q1.y);
    TextWord<kPred>(t, g, v1.z, q1.z);
    TextWord<kPred>(t, g, v1.w, q1.w);
    const int32_t left = c.left;
    c.left = left - 32 > -(1 << 30) ? left - 32 : -(1 << 30);
    if (!c.active)
        return;
    if (g == t.H) {
        // the sink is absorbing (its '\n' entry too): the lane was, or fell, outside the hot rows somewhere in the block
        TextSlow(a, t, fin_hot, outs, c, s, before == t.H ? s.cold : before, cpos, 0, 32, seg_hi, total);
        return;
    }
    s.g = g;
    uint32_t ends = NewlineNibble(v0.x) | (NewlineNibble(v0.y) << 4) | (NewlineNibble(v0.z) << 8) | (NewlineNibble(v0.w) << 12) |
                    (NewlineNibble(v1.x) << 16) | (NewlineNibble(v1.y) << 20) | (NewlineNibble(v1.z) << 24) | (NewlineNibble(v1.w) << 28);
    if (ends == 0)
        return;
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(packs), "r"(q0.x), "r"(q0.y), "r"(q0.z), "r"(q0.w) : "memory");
    asm volatile("st.shared.v4.u32 [%0+512], {%1,%2,%3,%4};" ::"r"(packs), "r"(q1.x), "r"(q1.y), "r"(q1.z), "r"(q1.w) : "memory");
    const uint32_t n_lines = (uint32_t) a.n;
    do {
        const uint32_t k = (uint32_t) __ffs((int) ends) - 1u;
        ends &= ends - 1u;
        // the state in front of byte k: byte 3 - k % 4 of word k / 4 -- a hot state, the lane is not in the sink
        uint32_t st;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(st) : "r"(packs + ((k & 16u) << 5) + ((k & 15u) ^ 3u)) : "memory");
        TextReport(a, outs, c.line, fin_hot[st]);
        ++c.line;
        if (c.line >= n_lines || (int32_t) k + 1 >= left) {         // the next line starts behind the segment
            c.active = false;
            break;
        }
    } while (ends);
}

// A 16-byte chunk that may stick out of the text at either end: bytes outside read as '\n'.  Out of line, it is rare.
__device__ __noinline__ uint4 LoadText16Clipped(const uint8_t* aligned, uintptr_t buf_lo, uintptr_t buf_hi)
{
    uint32_t w[4] = {0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au};
    for (int k = 0; k < 16; ++k) {
        const uintptr_t at = reinterpret_cast<uintptr_t>(aligned) + k;
        if (at >= buf_lo && at < buf_hi)
            w[k >> 2] = (w[k >> 2] & ~(0xffu << (8 * (k & 3)))) | ((uint32_t) aligned[k] << (8 * (k & 3)));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void LoadBlock32(const uint8_t* aligned, uintptr_t buf_lo, uintptr_t buf_hi, uint4& v0, uint4& v1)
{
    if (reinterpret_cast<uintptr_t>(aligned) >= buf_lo && reinterpret_cast<uintptr_t>(aligned) + 32 <= buf_hi) {
        LoadStream32(aligned, v0, v1);
    } else {
        v0 = LoadText16Clipped(aligned, buf_lo, buf_hi);
        v1 = LoadText16Clipped(aligned + 16, buf_lo, buf_hi);
    }
}

template <bool kPred>
__global__ void __launch_bounds__(kBlock, kTextBlocksPerSM) ScanTextKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);
    // behind a line the walk starts over: '\n' leads to the start state from every hot row
    for (uint32_t g = threadIdx.x; g < a.hot; g += blockDim.x)
        sv.hot[g * kHotStride + '\n'] = (uint8_t) a.start;
    // what a line that stops in a hot state reports
    DeviceFin* const fin_hot = reinterpret_cast<DeviceFin*>(sv.stage);
    for (uint32_t g = threadIdx.x; g < a.hot; g += blockDim.x)
        fin_hot[g] = a.fin[g];
    __syncthreads();

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0 | (1u << ('\n' & 31));       // the exit filter must let the separator through

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t outs = TextOutputs(a);
    const uint32_t packs = SmemAddr(sv.stage) + (uint32_t) kTextFinBytes + (threadIdx.x >> 5) * 1024u + lane * 16u;
    const uint64_t total = a.offsets[a.n] - 1;           // position of the last separator (real or the end of the text)
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + total;
    const uint32_t mis0 = (uint32_t) (buf_lo & 31);
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    // Segment size: about kTextSegment bytes, chosen so that the units (32 segments) come out as a whole number of
    // rounds over the grid's warps -- with a couple of units per warp, one unit more or less is a third of the run time.
    uint64_t seg = a.text_segment;
    if (seg == 0) {
        const uint64_t lanes = 32 * warps;
        const uint64_t rounds = (total + 64 + lanes * kTextSegment - 1) / (lanes * kTextSegment);
        seg = ((total + 64 + lanes * rounds - 1) / (lanes * rounds) + 31) / 32 * 32;
        seg = seg < 64 ? 64 : seg;
    }
    const uint64_t segments = (total + mis0) / seg + 1;
    const uint64_t units = (segments + 31) / 32;

    for (uint64_t unit = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t sidx = unit * 32 + lane;
        const uint64_t seg_hi = (sidx + 1) * seg - mis0;
        TextLane c;
        c.line = 0;
        c.left = 0;
        c.active = false;
        LaneState s;
        s.g = a.start;
        s.cold = 0;
        const uint8_t* p = a.corpus;
        int64_t pos = 0;
        if (sidx < segments) {
            const uint64_t seg_lo = sidx * seg > mis0 ? sidx * seg - mis0 : 0;
            uint64_t lo = 0, hi = a.n;              // first line that starts at or behind seg_lo
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (a.offsets[mid] < seg_lo)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < a.n) {
                const uint64_t pos0 = a.offsets[lo];
                if (pos0 < seg_hi) {
                    c.line = (uint32_t) lo;
                    c.active = true;
                    pos = (int64_t) ((pos0 + mis0) & ~31ull) - (int64_t) mis0;      // the 32-byte block the line starts in
                    const uint32_t skip = (uint32_t) ((int64_t) pos0 - pos);
                    if (skip) {
                        // the line starts inside the block: the rest of the block byte by byte
                        TextSlow(a, t, fin_hot, outs, c, s, a.start, pos, skip, 32, seg_hi, total);
                        pos += 32;
                    }
                    p = a.corpus + pos;
                    c.left = (int32_t) ((int64_t) seg_hi - pos);
                }
            }
        }
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
        if (c.active)
            LoadBlock32(p, buf_lo, buf_hi, v0, v1);
        while (__any_sync(0xffffffffu, c.active)) {
            uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
            if (c.active)
                LoadBlock32(p + 32, buf_lo, buf_hi, n0, n1);
            TextBlock<kPred>(a, t, fin_hot, outs, c, s, v0, v1, packs, pos, seg_hi, total);
            v0 = n0;
            v1 = n1;
            p += 32;
            pos += 32;
        }
    }
}

// ---------------------------------------------------------------- PRIV variant
//
// The plain walk is bound by shared-memory wavefronts once lanes sit in different
// rows (glued scanners: ~2.4 wavefronts per load, ncu r01).  Here the hottest rows
// are replicated into all 32 banks: lane l reads only bank l, so every load is one
// wavefront whatever the states and bytes are.  Layout (byte address inside the
// private region):   [19:14] quad q   [13:7] byte b (< 128)   [6:2] lane   [1:0] row-in-quad s
// A lane's state is its private row id e = 4q + s; one step is
//     e = LDS.U8 [priv + (((e * 0x1001) & 0xFC003) | (b << 7) | (lane << 2))]
// (the multiply drops s into bits 1:0 and q into bits 19:14; the other fields are
// disjoint, so one LOP3 assembles the address).  Rows that are not private map to the
// sink row; a lane found in the sink after a 4-byte word (or a word holding a byte
// >= 128) re-walks that word through the shared hot rows / the complete table and
// re-enters a private row when it can.

constexpr int kPrivBlock = 1024;
constexpr uint32_t kPrivMask = 0x000FC003u;

__device__ __forceinline__ uint32_t LoadSharedU8(uint32_t shared_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}

// Four steps.  `e` is the lane's private row id.  One step costs two ALU-pipe
// instructions (PRMT, LOP3), three FMA-pipe ones (IMAD x3) and one conflict-free LDS.
// lane_base = shared-window address of the private region + lane * 4.
__device__ __forceinline__ void PrivWord(uint32_t& e, uint32_t w, uint32_t lane_base)
{
    // The private rows cover bytes 0..127 only.  A byte >= 128 sends the whole chunk to the
    // re-walk anyway (PrivChunk), but its speculative step must not index past the table:
    // clear bit 7 of every byte first (one LOP3 per word).
    w &= 0x7F7F7F7Fu;
    const uint32_t k0 = __byte_perm(w, 0, 0x4440) * 128u + lane_base;
    const uint32_t k1 = __byte_perm(w, 0, 0x4441) * 128u + lane_base;
    const uint32_t k2 = __byte_perm(w, 0, 0x4442) * 128u + lane_base;
    const uint32_t k3 = __umulhi(w, 256u) * 128u + lane_base;           // w >> 24 without the ALU pipe
    e = LoadSharedU8(((e * 0x1001u) & kPrivMask) + k0);
    e = LoadSharedU8(((e * 0x1001u) & kPrivMask) + k1);
    e = LoadSharedU8(((e * 0x1001u) & kPrivMask) + k2);
    e = LoadSharedU8(((e * 0x1001u) & kPrivMask) + k3);
}

// Sixteen input bytes.  The sink row is absorbing, so one test per chunk finds every lane
// that left the private rows (or met a byte >= 128, which the private rows do not cover);
// such a lane re-walks the chunk through the shared hot rows (PRMT + LDS per byte, as in the
// plain kernel) and through the complete table only if that fails too.
__device__ __forceinline__ void PrivChunk(const Tables& t, uint32_t& e, uint32_t& other, uint4 v, uint32_t lane_base,
                                          uint32_t sink_id, uint32_t real_rows)
{
    const uint32_t before = e;
    PrivWord(e, v.x, lane_base);
    PrivWord(e, v.y, lane_base);
    PrivWord(e, v.z, lane_base);
    PrivWord(e, v.w, lane_base);
    if (e == sink_id || ((v.x | v.y | v.z | v.w) & 0x80808080u) != 0) {
        const uint32_t from = before == sink_id ? other : before;
        uint32_t g = from < t.H ? from : t.H;
        FastWord<false>(t, g, v.x);
        FastWord<false>(t, g, v.y);
        FastWord<false>(t, g, v.z);
        FastWord<false>(t, g, v.w);
        uint32_t full = g;
        if (g == t.H)
            full = ReplayChunk(t.hot, t.cls, t.full, t.H, t.letters | (t.wide << 31), from, v);
        if (full < real_rows) {
            e = full;
        } else {
            e = sink_id;
            other = full;
        }
    }
}

__global__ void __launch_bounds__(kPrivBlock, 1) ScanUniformPrivKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot_small, a.priv_rows);
    StageTables(a, sv, a.hot8_small, a.hot_small);
    {
        // replicate every packed word into the 32 banks
        const uint32_t words = (a.priv_rows / 4) * 128 * 32;
        uint32_t* dst = reinterpret_cast<uint32_t*>(sv.priv);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            dst[i] = __ldg(a.priv_packed + (i >> 5));
        __syncthreads();
    }

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot_small;          // the second tier seen by this kernel
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = 0;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lane_base = SmemAddr(sv.priv) + (lane << 2);
    const uint32_t real_rows = a.priv_rows - 1 < a.hot_small ? a.priv_rows - 1 : a.hot_small;
    const uint32_t sink_id = a.priv_rows - 1;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * (kPrivBlock / 32);
    const uint32_t len = (uint32_t) a.fixed_len;

    for (uint64_t unit = (uint64_t) blockIdx.x * (kPrivBlock / 32) + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t i = unit * 32 + lane;
        const bool valid = i < a.n;
        const uint8_t* p = a.corpus + (valid ? i : a.n - 1) * (uint64_t) len;

        uint32_t other = a.start;
        uint32_t e = a.start < real_rows ? a.start : sink_id;

        uint4 c0, c1, d0, d1;
        LoadStream32(p, c0, c1);
        for (uint32_t off = 0;;) {
            off += 32;
            const bool more_d = off < len;
            if (more_d)
                LoadStream32(p + off, d0, d1);
            PrivChunk(t, e, other, c0, lane_base, sink_id, real_rows);
            PrivChunk(t, e, other, c1, lane_base, sink_id, real_rows);
            if (!more_d)
                break;
            off += 32;
            const bool more_c = off < len;
            if (more_c)
                LoadStream32(p + off, c0, c1);
            PrivChunk(t, e, other, d0, lane_base, sink_id, real_rows);
            PrivChunk(t, e, other, d1, lane_base, sink_id, real_rows);
            if (!more_c)
                break;
        }

        LaneState fs;
        fs.g = t.H;
        fs.cold = e == sink_id ? other : e;
        Report(a, t, fs, unit, i, valid);
    }
}

__device__ __forceinline__ uint32_t FullNext(const Tables& t, uint32_t s, uint32_t letter)
{
    size_t at = (size_t) s * t.letters + letter;
    return t.wide ? __ldg(static_cast<const uint32_t*>(t.full) + at) : (uint32_t) __ldg(static_cast<const uint16_t*>(t.full) + at);
}

// ---------------------------------------------------------------- prefix and suffix scans
//
// Pire::LongestPrefix / ShortestPrefix (run.h:277-311, predicates run.h:69-100) and LongestSuffix /
// ShortestSuffix (run.h:316-362) for a batch: the same walk, but Final() and Dead() matter after every
// byte -- the longest scan remembers the last position whose state is final, the shortest stops at the
// first, both stop in a dead state (pire_ut.cpp ScanTermination@475).  The suffix scans walk the string
// from its last byte to its first (the scanner is normally built from Fsm::Reverse()).  One string per
// lane through the generic kernel's machinery (cp.async ring, fused hot rows).  Final hot states carry the
// highest hot ids, so the maximum id over the 16 steps of a chunk says whether the chunk entered a final
// state or left the hot rows; such a chunk is walked again, branch-free, marking where; only a chunk that
// leaves the hot rows is replayed byte by byte with the predicate.  A dead state is never final and only
// leads to dead states, so noticing it late cannot change the answer: it is looked for once per ring round
// (64 bytes) to stop the lane, and a lane that has stopped no longer fetches its string.
struct PrefixLane {
    uint32_t consumed;      // bytes walked so far
    uint32_t pos;           // answer so far, kNoPrefix = none
    bool stop;
};
constexpr uint32_t kNoPrefix = 0xFFFFFFFFu;

// The bytes of an edge chunk from index `top` downwards (suffix scans).
struct EdgeBytesDown {
    uint64_t lo, hi;
    __device__ __forceinline__ EdgeBytesDown(uint4 v, uint32_t top)
    {
        lo = (uint64_t) v.x | ((uint64_t) v.y << 32);
        hi = (uint64_t) v.z | ((uint64_t) v.w << 32);
        uint32_t drop = 15 - top;                       // bytes above `top` are not ours
        if (drop >= 8) {
            hi = lo;
            lo = 0;
            drop -= 8;
        }
        if (drop) {
            hi = (hi << (8 * drop)) | (lo >> (64 - 8 * drop));
            lo <<= 8 * drop;
        }
    }
    __device__ __forceinline__ uint32_t Next()
    {
        uint32_t b = (uint32_t) (hi >> 56);
        hi = (hi << 8) | (lo >> 56);
        lo <<= 8;
        return b;
    }
};

template <bool kShortest>
__device__ __forceinline__ void PrefixCheck(const ScanArgs& a, uint32_t H, const uint8_t* hot_flags, uint32_t state, PrefixLane& l)
{
    const uint32_t fl = state < H ? hot_flags[state] : __ldg(a.flags + state);
    if (fl & 1u) {                                                   // Final: run.h:76-79 / :92-93 / :329-330 / :353
        l.pos = l.consumed;
        l.stop = kShortest;
    }
    if (fl & 2u)                                                     // Dead: run.h:82 / :94 / :328 / :353
        l.stop = true;
}

template <bool kShortest, bool kReverse, bool kPred = false, bool kIdp = false>
__device__ __forceinline__ void PrefixChunk16(const ScanArgs& a, const Tables& t, const uint8_t* hot_flags, LaneState& s, uint4 v,
                                              PrefixLane& l)
{
    const uint32_t before = s.g;
    uint32_t g = before, top = 0, low = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const int at = kReverse ? 3 - w : w;
        const uint32_t word = at == 0 ? v.x : at == 1 ? v.y : at == 2 ? v.z : v.w;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            FastStep<kPred, kIdp>(t, g, word, 0x5540 + (kReverse ? 3 - b : b));
            top = max(top, g);
            if (!kShortest)
                low = min(low, g);
        }
    }
    if (top < a.first_final_hot) {           // sixteen steps through non-final hot states
        s.g = g;
        l.consumed += 16;
        return;
    }
    if (!kShortest && low >= a.first_final_hot && top != t.H) {
        // sixteen steps through final hot states (a lane behind an unanchored match sits in such states for the rest of
        // its string): the longest prefix so far ends with this chunk, no second pass needed
        s.g = g;
        l.consumed += 16;
        l.pos = l.consumed;
        return;
    }
    // ShortestSuffix steps BeginMark from the state it stopped in (run.h:357-359), so that scan needs the
    // state at the first final step, not the one after the chunk: it goes straight to the byte-wise replay
    // (once per string).
    if (before != t.H && !(kShortest && kReverse)) {
        // a final state (or the sink) was entered: the chunk again, branch-free, noting where.  `mark` is the
        // 1-based step of the last (longest) or first (shortest) final state entered; the sink id also
        // compares >= first_final_hot, but a lane that reached the sink takes the slow path below instead.
        uint32_t h = before, mark = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int at = kReverse ? 3 - w : w;
            const uint32_t word = at == 0 ? v.x : at == 1 ? v.y : at == 2 ? v.z : v.w;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                FastStep<kPred, kIdp>(t, h, word, 0x5540 + (kReverse ? 3 - b : b));
                const bool final = h >= a.first_final_hot;
                if (kShortest)
                    mark = final && mark == 0 ? (uint32_t) (4 * w + b + 1) : mark;
                else
                    mark = final ? (uint32_t) (4 * w + b + 1) : mark;
            }
        }
        if (h != t.H) {
            if (mark) {
                l.pos = l.consumed + mark;
                l.stop = kShortest;
            }
            l.consumed += 16;
            s.g = h;
            return;
        }
    }
    uint32_t full = before == t.H ? s.cold : before;
    if (kReverse) {
        EdgeBytesDown eb(v, 15);
        for (int k = 0; k < 16 && !l.stop; ++k) {
            full = SlowStep(t, full, eb.Next());
            ++l.consumed;
            PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
        }
    } else {
        EdgeBytes eb(v, 0);
        for (int k = 0; k < 16 && !l.stop; ++k) {
            full = SlowStep(t, full, eb.Next());
            ++l.consumed;
            PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
        }
    }
    SetFull(t, s, full);
}

template <bool kShortest, bool kReverse>
__global__ void __launch_bounds__(kBlock, kGenericBlocksPerSM) PrefixKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);
    uint8_t* const hot_flags = sv.stage + kStageBytes;              // H + 1 bytes behind the staging ring
    for (uint32_t i = threadIdx.x; i <= a.hot; i += blockDim.x)
        hot_flags[i] = i < a.hot ? a.flags[i] : 0;
    __syncthreads();

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = 0;

    const uint32_t lane = threadIdx.x & 31;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    const uint32_t stage = SmemAddr(sv.stage) + (((threadIdx.x >> 5) * kStageSlots) * 32 + lane) * 16;
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + (a.offsets ? a.offsets[a.n] - a.trim : a.n * a.fixed_len);

    for (uint64_t unit = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t i = unit * 32 + lane;
        const bool valid = i < a.n;
        uint64_t b = 0, e = 0;
        if (valid) {
            if (a.offsets) {
                b = a.offsets[i];
                e = a.offsets[i + 1] - a.trim;
                e = e < b ? b : e;       // an empty entry of a trimmed (lines) batch, or caller offsets that step back
            } else {
                b = i * a.fixed_len;
                e = b + a.fixed_len;
            }
        }
        const uint8_t* const first = a.corpus + b;
        const uint8_t* const end = a.corpus + e;
        const uint32_t len = (uint32_t) (e - b);

        // Initialize() and the first mark: BeginMark for a prefix scan (run.h:282-283), EndMark for a
        // suffix scan (run.h:321-322)
        uint32_t full = a.initial;
        if (a.with_begin)
            full = FullNext(t, full, a.begin_class);
        PrefixLane l;
        l.consumed = 0;
        l.pos = kNoPrefix;
        l.stop = !valid;
        {
            const uint32_t fl = full < t.H ? hot_flags[full] : __ldg(a.flags + full);
            if (fl & 1u) {                                               // run.h:284 / :301-302 / :329-330 / :353
                l.pos = 0;
                l.stop = l.stop || kShortest;
            }
            if (kReverse && (fl & 2u))                                   // run.h:328 / :353: Dead is looked at before the first byte
                l.stop = true;
        }
        // edge bytes on the side the walk starts from, up to a 16-byte boundary
        const uint8_t* p = kReverse ? end : first;                       // forward: next byte; reverse: one past the next byte
        {
            const uint32_t misalign = (uint32_t) (reinterpret_cast<uintptr_t>(p) & 15);
            if (len != 0 && misalign != 0) {
                const uint32_t span = kReverse ? misalign : 16 - misalign;      // bytes between p and the boundary
                const uint32_t nedge = len < span ? len : span;
                const uint8_t* chunk = p - misalign;
                if (!l.stop) {
                    const bool whole = reinterpret_cast<uintptr_t>(chunk) >= buf_lo && reinterpret_cast<uintptr_t>(chunk) + 16 <= buf_hi;
                    if (kReverse) {
                        EdgeBytesDown eb(whole ? LoadEdge16(chunk) : make_uint4(0, 0, 0, 0), misalign - 1);
                        for (uint32_t k = 0; k < nedge && !l.stop; ++k) {
                            const uint32_t byte = whole ? eb.Next() : (uint32_t) p[-1 - (int) k];
                            full = SlowStep(t, full, byte);
                            ++l.consumed;
                            PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
                        }
                    } else {
                        EdgeBytes eb(whole ? LoadEdge16(chunk) : make_uint4(0, 0, 0, 0), misalign);
                        for (uint32_t k = 0; k < nedge && !l.stop; ++k) {
                            const uint32_t byte = whole ? eb.Next() : (uint32_t) p[k];
                            full = SlowStep(t, full, byte);
                            ++l.consumed;
                            PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
                        }
                    }
                }
                p = kReverse ? p - nedge : p + nedge;
            }
        }
        LaneState s;
        SetFull(t, s, full);
        if (!kReverse && a.uniform) {
            // Fixed-length, 32-byte aligned batch (the BASELINE configs' shape): no edge bytes, and the strings stream
            // through registers like in the uniform scan kernel -- one LDG.256 per lane per 32 bytes, prefetched one
            // block ahead -- instead of the cp.async ring, whose shared-memory round trip costs half a wavefront per
            // step on the pipe that bounds the walk.
            // every lane of the warp walks the loop (its control flow holds a warp vote), also the lanes past the end
            // of the batch: they are stopped from the start and read string 0
            const uint32_t ulen = (uint32_t) a.fixed_len;
            const uint8_t* const src = valid ? first : a.corpus;
            if (ulen != 0) {
                uint4 a0, a1, b0, b1;
                LoadStream32(src, a0, a1);
                for (uint32_t off = 0;;) {
                    off += 32;
                    const bool more_b = off < ulen;
                    if (more_b)
                        LoadStream32(src + off, b0, b1);
                    if (!l.stop)
                        PrefixChunk16<kShortest, kReverse>(a, t, hot_flags, s, a0, l);
                    if (!l.stop)
                        PrefixChunk16<kShortest, kReverse>(a, t, hot_flags, s, a1, l);
                    if (!more_b)
                        break;
                    off += 32;
                    const bool more_a = off < ulen;
                    if (more_a)
                        LoadStream32(src + off, a0, a1);
                    if (!l.stop)
                        PrefixChunk16<kShortest, kReverse>(a, t, hot_flags, s, b0, l);
                    if (!l.stop)
                        PrefixChunk16<kShortest, kReverse>(a, t, hot_flags, s, b1, l);
                    // a dead state only leads to dead states: stop the lane once it is noticed (every 64 bytes)
                    if (!l.stop) {
                        const uint32_t at = FullState(t, s);
                        if ((at < t.H ? hot_flags[at] : __ldg(a.flags + at)) & 2u)
                            l.stop = true;
                    }
                    if (!more_a || __all_sync(0xffffffffu, l.stop))
                        break;
                }
            }
            p = end;
        }
        // body: whole 16-byte chunks; chunk c is at p + 16c (forward) or p - 16(c+1) (reverse)
        const uint32_t chunks = (uint32_t) ((kReverse ? p - first : end - p) >> 4);
#pragma unroll
        for (int j = 0; j < kStageSlots; ++j) {
            if ((uint32_t) j < chunks && !l.stop)
                CopyAsync16(stage + j * 512, kReverse ? p - 16 * (j + 1) : p + 16 * j);
            CopyAsyncCommit();
        }
        for (uint32_t k = 0; __any_sync(0xffffffffu, k < chunks && !l.stop); k += kStageSlots) {
#pragma unroll
            for (int j = 0; j < kStageSlots; ++j) {
                CopyAsyncWait<kStageSlots - 1>();
                const uint4 v = LoadShared16(stage + j * 512);
                const bool had = k + j < chunks && !l.stop;              // this slot was filled for this lane
                if (k + kStageSlots + j < chunks && !l.stop) {
                    const size_t c = (size_t) k + kStageSlots + j;
                    CopyAsync16(stage + j * 512, kReverse ? p - 16 * (c + 1) : p + 16 * c);
                }
                CopyAsyncCommit();
                if (had)
                    PrefixChunk16<kShortest, kReverse>(a, t, hot_flags, s, v, l);
            }
            // a dead state only leads to dead states: stop the lane (and its fetches) once it is noticed
            if (!l.stop) {
                const uint32_t at = FullState(t, s);
                if ((at < t.H ? hot_flags[at] : __ldg(a.flags + at)) & 2u)
                    l.stop = true;
            }
        }
        CopyAsyncWait<0>();
        full = FullState(t, s);
        p = kReverse ? p - 16 * (size_t) chunks : p + 16 * (size_t) chunks;
        // edge bytes on the far side: fewer than 16, p is on a 16-byte boundary
        const uint32_t nfar = (uint32_t) (kReverse ? p - first : end - p);
        if (nfar != 0 && !l.stop) {
            const uint8_t* chunk = kReverse ? p - 16 : p;
            const bool whole = reinterpret_cast<uintptr_t>(chunk) >= buf_lo && reinterpret_cast<uintptr_t>(chunk) + 16 <= buf_hi;
            if (kReverse) {
                EdgeBytesDown eb(whole ? LoadEdge16(chunk) : make_uint4(0, 0, 0, 0), 15);
                for (uint32_t k = 0; k < nfar && !l.stop; ++k) {
                    const uint32_t byte = whole ? eb.Next() : (uint32_t) p[-1 - (int) k];
                    full = SlowStep(t, full, byte);
                    ++l.consumed;
                    PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
                }
            } else {
                EdgeBytes eb(whole ? LoadEdge16(chunk) : make_uint4(0, 0, 0, 0), 0);
                for (uint32_t k = 0; k < nfar && !l.stop; ++k) {
                    const uint32_t byte = whole ? eb.Next() : (uint32_t) p[k];
                    full = SlowStep(t, full, byte);
                    ++l.consumed;
                    PrefixCheck<kShortest>(a, t.H, hot_flags, full, l);
                }
            }
        }
        if (valid) {
            if (kReverse && kShortest) {
                // run.h:357-360: the last mark is stepped from wherever the scan stopped, and the answer is
                // the stopping place if that state is final
                if (a.through_end)
                    full = FullNext(t, full, a.end_class);
                l.pos = (__ldg(a.flags + full) & 1u) ? l.consumed : kNoPrefix;
            } else if (a.through_end) {                                  // run.h:286-290 / :305-309 / :336-340
                const uint32_t last = FullNext(t, full, a.end_class);
                if ((__ldg(a.flags + last) & 1u) && (!kShortest || l.pos == kNoPrefix))
                    l.pos = len;
            }
            a.prefix_len[i] = l.pos;
        }
    }
}

// Forward prefix scans of a fixed-length, 32-byte aligned batch (the BASELINE configs' shape) in a kernel of their own:
// no edge bytes and no staging ring, so three CTAs (or two of twenty warps) fit an SM instead of the generic kernel's two
// of sixteen, and the walk can use the exit filter of hot id 0 (kPred): a lane resting there on a byte that cannot leave
// keeps g = 0, which is all the running maximum needs.  (The look-ahead filter does not carry over: it skips the one-step
// states behind an exit byte, and a prefix scan must see them if they are final.)
template <bool kShortest, bool kPred, bool kIdp>
__global__ void __maxnreg__(48) PrefixUniformKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);
    uint8_t* const hot_flags = sv.stage;                            // H + 1 bytes, where the generic kernels keep their ring
    for (uint32_t i = threadIdx.x; i <= a.hot; i += blockDim.x)
        hot_flags[i] = i < a.hot ? a.flags[i] : 0;
    __syncthreads();

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = a.exit_bitmap0;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * warps_per_block;
    const uint32_t ulen = (uint32_t) a.fixed_len;

    for (uint64_t unit = (uint64_t) blockIdx.x * warps_per_block + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t i = unit * 32 + lane;
        const bool valid = i < a.n;
        // Initialize() and BeginMark (run.h:282-283)
        uint32_t full = a.initial;
        if (a.with_begin)
            full = FullNext(t, full, a.begin_class);
        PrefixLane l;
        l.consumed = 0;
        l.pos = kNoPrefix;
        l.stop = !valid;
        if ((full < t.H ? hot_flags[full] : __ldg(a.flags + full)) & 1u) {      // run.h:284 / :301-302
            l.pos = 0;
            l.stop = l.stop || kShortest;
        }
        LaneState s;
        SetFull(t, s, full);
        // every lane of the warp walks the loop (its control flow holds a warp vote), also the lanes past the end of
        // the batch: they are stopped from the start and read string 0
        const uint8_t* const src = a.corpus + (valid ? i : 0) * (uint64_t) ulen;
        if (ulen != 0) {
            uint4 a0, a1, b0, b1;
            LoadStream32(src, a0, a1);
            for (uint32_t off = 0;;) {
                off += 32;
                const bool more_b = off < ulen;
                if (more_b)
                    LoadStream32(src + off, b0, b1);
                if (!l.stop)
                    PrefixChunk16<kShortest, false, kPred, kIdp>(a, t, hot_flags, s, a0, l);
                if (!l.stop)
                    PrefixChunk16<kShortest, false, kPred, kIdp>(a, t, hot_flags, s, a1, l);
                if (!more_b)
                    break;
                off += 32;
                const bool more_a = off < ulen;
                if (more_a)
                    LoadStream32(src + off, a0, a1);
                if (!l.stop)
                    PrefixChunk16<kShortest, false, kPred, kIdp>(a, t, hot_flags, s, b0, l);
                if (!l.stop)
                    PrefixChunk16<kShortest, false, kPred, kIdp>(a, t, hot_flags, s, b1, l);
                // a dead state only leads to dead states: stop the lane once it is noticed (every 64 bytes)
                if (!l.stop) {
                    const uint32_t at = FullState(t, s);
                    if ((at < t.H ? hot_flags[at] : __ldg(a.flags + at)) & 2u)
                        l.stop = true;
                }
                if (!more_a || __all_sync(0xffffffffu, l.stop))
                    break;
            }
        }
        if (valid) {
            if (a.through_end) {                                             // run.h:286-290 / :305-309
                const uint32_t last = FullNext(t, FullState(t, s), a.end_class);
                if ((__ldg(a.flags + last) & 1u) && (!kShortest || l.pos == kNoPrefix))
                    l.pos = ulen;
            }
            a.prefix_len[i] = l.pos;
        }
    }
}

// ---------------------------------------------------------------- counting (HalfFinalScanner)
//
// pire/scanners/half_final.h: the same table walk, but after Initialize() and after every symbol
// TakeAction (:154-163) adds one to the counter of each regexp listed for the state when the state is
// final.  The per-string result is the vector of counters (State::Result, :88-90).
//
// The walk is the generic kernel's (one string per lane, 16-byte chunks through the cp.async ring, fused
// hot rows in shared memory).  Two ways to keep the counters, chosen per automaton on the host:
//  * packed (kWords = 1 or 2, up to 16 regexps): every state has its increments as 8-bit fields of one
//    or two 64-bit words (zero for non-final states; the hot states' words sit in shared memory), so
//    TakeAction is an unconditional 64-bit add per word.  Fields are widened to 16 bits after every
//    chunk and written to the string's row of counters every 120 chunks, before they can wrap;
//  * lists (kWords = 0): the accept list of a final state is walked; counters 0..3 in registers, the
//    rest straight in the string's row in global memory (lane-private, no atomics).
// Final hot states carry the highest hot ids, so one running maximum per chunk tells whether any of its
// 16 steps landed in a final state or left the hot rows; if not, the chunk costs what a plain scan costs.
// When a sample showed final states to be frequent (kAlways) that first pass is skipped and every chunk
// is counted.
template <int kWords>
struct Counter {
    uint64_t a8[kWords];            // 8 x 8 bits: at most 16 steps x 15 per field between Widen() calls
    uint64_t even[kWords], odd[kWords];   // 4 x 16 bits each
    uint32_t groups;
    uint32_t* row;
    const uint64_t* hot_w;          // shared: (H + 1) * kWords words, the sink's are zero
    const uint64_t* all_w;          // global: states * kWords

    __device__ __forceinline__ void Reset(const ScanArgs& a, const uint64_t* hot_weights, uint32_t* r)
    {
#pragma unroll
        for (int j = 0; j < kWords; ++j)
            a8[j] = even[j] = odd[j] = 0;
        groups = 0;
        row = r;
        hot_w = hot_weights;
        all_w = a.weights;
    }
    __device__ __forceinline__ void Step(const ScanArgs&, uint32_t H, uint32_t s)          // TakeAction
    {
#pragma unroll
        for (int j = 0; j < kWords; ++j)
            a8[j] += s < H ? hot_w[s * kWords + j] : __ldg(all_w + (size_t) s * kWords + j);
    }
    __device__ __forceinline__ void Flush(uint32_t regexps)
    {
#pragma unroll
        for (int j = 0; j < kWords; ++j) {
            for (uint32_t f = 0; f < 8 && j * 8 + f < regexps; ++f) {
                const uint64_t src = (f & 1) ? odd[j] : even[j];
                const uint32_t add = (uint32_t) (src >> (16 * (f >> 1))) & 0xffffu;
                if (add)
                    row[j * 8 + f] += add;
            }
            even[j] = odd[j] = 0;
        }
        groups = 0;
    }
    __device__ __forceinline__ void EndGroup(uint32_t regexps)      // after at most 16 steps
    {
#pragma unroll
        for (int j = 0; j < kWords; ++j) {
            even[j] += a8[j] & 0x00FF00FF00FF00FFull;
            odd[j] += (a8[j] >> 8) & 0x00FF00FF00FF00FFull;
            a8[j] = 0;
        }
        if (++groups >= 120)                                        // 120 x 16 x 15 < 65536
            Flush(regexps);
    }
    __device__ __forceinline__ void Discard()
    {
#pragma unroll
        for (int j = 0; j < kWords; ++j)
            a8[j] = 0;
    }
    __device__ __forceinline__ void Finish(uint32_t regexps) { Flush(regexps); }
};

template <>
struct Counter<0> {
    uint32_t c0, c1, c2, c3;
    uint32_t* row;

    __device__ __forceinline__ void Reset(const ScanArgs&, const uint64_t*, uint32_t* r)
    {
        c0 = c1 = c2 = c3 = 0;
        row = r;
    }
    __device__ __forceinline__ void Step(const ScanArgs& a, uint32_t H, uint32_t s)
    {
        const bool final = s < H ? s >= a.first_final_hot : (__ldg(a.flags + s) & 1u) != 0;
        if (!final)
            return;
        uint32_t k = __ldg(a.acc_begin + s);
        const uint32_t e = __ldg(a.acc_begin + s + 1);
        for (; k < e; ++k) {
            const uint32_t id = __ldg(a.acc_ids + k);
            c0 += id == 0;
            c1 += id == 1;
            c2 += id == 2;
            c3 += id == 3;
            if (id >= 4)
                row[id] += 1;
        }
    }
    __device__ __forceinline__ void EndGroup(uint32_t) {}
    __device__ __forceinline__ void Discard() {}
    __device__ __forceinline__ void Finish(uint32_t regexps)
    {
        row[0] += c0;
        if (regexps > 1) row[1] += c1;
        if (regexps > 2) row[2] += c2;
        if (regexps > 3) row[3] += c3;
    }
};

template <int kWords, bool kAlways>
__device__ __forceinline__ void CountChunk16(const ScanArgs& a, const Tables& t, LaneState& s, uint4 v, Counter<kWords>& c)
{
    const uint32_t before = s.g;
    if (!kAlways || kWords == 0) {
        uint32_t g = before, top = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t word = w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w;
            FastStep<false>(t, g, word, 0x5540);
            top = max(top, g);
            FastStep<false>(t, g, word, 0x5541);
            top = max(top, g);
            FastStep<false>(t, g, word, 0x5542);
            top = max(top, g);
            FastStep<false>(t, g, word, 0x5543);
            top = max(top, g);
        }
        if (top < a.first_final_hot) {       // sixteen steps through non-final hot states: nothing to count
            s.g = g;
            return;
        }
    }
    if (kWords > 0 && before != t.H) {
        // the chunk again (or, kAlways, for the first time) with the packed increments of every state entered
        uint32_t g = before;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t word = w == 0 ? v.x : w == 1 ? v.y : w == 2 ? v.z : v.w;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                FastStep<false>(t, g, word, 0x5540 + b);
                c.Step(a, 0xffffffffu, g);              // g <= H: always the shared copy
            }
        }
        if (g != t.H) {
            c.EndGroup(a.regexps);
            s.g = g;
            return;
        }
        // left the hot rows somewhere inside: forget what was added (the sink's increments are zero, but
        // the steps after the miss were not the real ones) and replay through the complete table
        c.Discard();
    }
    uint32_t full = before == t.H ? s.cold : before;
    EdgeBytes eb(v, 0);
    for (int k = 0; k < 16; ++k) {
        full = SlowStep(t, full, eb.Next());
        c.Step(a, t.H, full);
    }
    c.EndGroup(a.regexps);
    SetFull(t, s, full);
}
template <int kWords, bool kAlways>
__global__ void __launch_bounds__(kBlock, kGenericBlocksPerSM) CountKernel(const __grid_constant__ ScanArgs a)
{
    uint8_t* const smem = pire_b200_smem;
    SharedView sv = CarveShared(smem, a.hot);
    StageTables(a, sv, a.hot8, a.hot);
    // packed increments of the hot states (and zeros for the sink) behind the staging ring
    uint64_t* const hot_w = reinterpret_cast<uint64_t*>(sv.stage + kStageBytes);
    if (kWords > 0) {
        for (uint32_t i = threadIdx.x; i < (a.hot + 1) * kWords; i += blockDim.x)
            hot_w[i] = i < a.hot * kWords ? a.weights[i] : 0;
        __syncthreads();
    }

    Tables t;
    t.hot = sv.hot;
    t.base = SmemAddr(sv.hot);
    t.cls = sv.cls;
    t.full = a.full;
    t.H = a.hot;
    t.letters = a.letters;
    t.wide = a.wide;
    t.m0 = 0;

    const uint32_t lane = threadIdx.x & 31;
    const uint64_t units = (a.n + 31) / 32;
    const uint64_t warps = (uint64_t) gridDim.x * kWarpsPerBlock;
    const uint32_t stage = SmemAddr(sv.stage) + (((threadIdx.x >> 5) * kStageSlots) * 32 + lane) * 16;
    const uintptr_t buf_lo = reinterpret_cast<uintptr_t>(a.corpus);
    const uintptr_t buf_hi = buf_lo + (a.offsets ? a.offsets[a.n] - a.trim : a.n * a.fixed_len);

    for (uint64_t unit = (uint64_t) blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); unit < units; unit += warps) {
        const uint64_t i = unit * 32 + lane;
        const bool valid = i < a.n;
        uint64_t b = 0, e = 0;
        if (valid) {
            if (a.offsets) {
                b = a.offsets[i];
                e = a.offsets[i + 1] - a.trim;
                e = e < b ? b : e;       // an empty entry of a trimmed (lines) batch, or caller offsets that step back
            } else {
                b = i * a.fixed_len;
                e = b + a.fixed_len;
            }
        }
        const uint8_t* p = a.corpus + b;
        const uint8_t* end = a.corpus + e;

        Counter<kWords> c;
        c.Reset(a, hot_w, a.counts + (valid ? i : 0) * a.regexps);
        uint32_t full = a.initial;
        if (valid) {
            c.Step(a, t.H, full);                                   // Initialize ends in TakeAction, half_final.h:136-141
            if (a.with_begin) {
                full = FullNext(t, full, a.begin_class);            // Step(BeginMark), run.h:50-57
                c.Step(a, t.H, full);
            }
            c.EndGroup(a.regexps);
        }
        {
            const uint32_t misalign = (uint32_t) (reinterpret_cast<uintptr_t>(p) & 15);
            if (p < end && misalign != 0) {
                const uint64_t room = (uint64_t) (end - p);
                const uint32_t nhead = room < 16 - misalign ? (uint32_t) room : 16 - misalign;
                const uint8_t* chunk = p - misalign;
                if (reinterpret_cast<uintptr_t>(chunk) >= buf_lo && reinterpret_cast<uintptr_t>(chunk) + 16 <= buf_hi) {
                    EdgeBytes eb(LoadEdge16(chunk), misalign);
                    for (uint32_t k = 0; k < nhead; ++k) {
                        full = SlowStep(t, full, eb.Next());
                        c.Step(a, t.H, full);
                    }
                } else {
                    for (uint32_t k = 0; k < nhead; ++k) {
                        full = SlowStep(t, full, p[k]);
                        c.Step(a, t.H, full);
                    }
                }
                c.EndGroup(a.regexps);
                p += nhead;
            }
        }
        LaneState s;
        SetFull(t, s, full);
        if (a.uniform) {
            // fixed-length, 32-byte aligned batch: LDG.256 ping-pong through registers, no staging ring (see PrefixKernel)
            const uint32_t len = (uint32_t) (end - p);
            if (len != 0) {
                uint4 a0, a1, b0, b1;
                LoadStream32(p, a0, a1);
                for (uint32_t off = 0;;) {
                    off += 32;
                    const bool more_b = off < len;
                    if (more_b)
                        LoadStream32(p + off, b0, b1);
                    CountChunk16<kWords, kAlways>(a, t, s, a0, c);
                    CountChunk16<kWords, kAlways>(a, t, s, a1, c);
                    if (!more_b)
                        break;
                    off += 32;
                    const bool more_a = off < len;
                    if (more_a)
                        LoadStream32(p + off, a0, a1);
                    CountChunk16<kWords, kAlways>(a, t, s, b0, c);
                    CountChunk16<kWords, kAlways>(a, t, s, b1, c);
                    if (!more_a)
                        break;
                }
            }
            p = end;
        }
        const uint32_t chunks = (uint32_t) ((end - p) >> 4);
#pragma unroll
        for (int j = 0; j < kStageSlots; ++j) {
            if ((uint32_t) j < chunks)
                CopyAsync16(stage + j * 512, p + 16 * j);
            CopyAsyncCommit();
        }
        for (uint32_t k = 0; __any_sync(0xffffffffu, k < chunks); k += kStageSlots) {
#pragma unroll
            for (int j = 0; j < kStageSlots; ++j) {
                CopyAsyncWait<kStageSlots - 1>();
                const uint4 v = LoadShared16(stage + j * 512);
                if (k + kStageSlots + j < chunks)
                    CopyAsync16(stage + j * 512, p + 16 * (size_t) (k + kStageSlots + j));
                CopyAsyncCommit();
                if (k + j < chunks)
                    CountChunk16<kWords, kAlways>(a, t, s, v, c);
            }
        }
        CopyAsyncWait<0>();
        full = FullState(t, s);
        p += 16 * (size_t) chunks;
        if (p < end) {
            const uint32_t ntail = (uint32_t) (end - p);
            if (reinterpret_cast<uintptr_t>(p) + 16 <= buf_hi) {
                EdgeBytes eb(LoadEdge16(p), 0);
                for (uint32_t k = 0; k < ntail; ++k) {
                    full = SlowStep(t, full, eb.Next());
                    c.Step(a, t.H, full);
                }
            } else {
                for (uint32_t k = 0; k < ntail; ++k) {
                    full = SlowStep(t, full, p[k]);
                    c.Step(a, t.H, full);
                }
            }
        }
        if (valid && a.through_end) {
            full = FullNext(t, full, a.end_class);                   // Step(EndMark)
            c.Step(a, t.H, full);
        }
        c.EndGroup(a.regexps);
        const bool final = valid && (__ldg(a.flags + full) & 1u) != 0;
        const unsigned matched = __ballot_sync(0xffffffffu, final);
        if (a.match_bits && lane == 0)
            a.match_bits[unit] = matched;
        if (valid)
            c.Finish(a.regexps);
    }
}

// Visit counter for pire_gpu_scanner_tune: how many input bytes are consumed in
// each state (new numbering).  Run-length compressed so that a lane resting in
// one state issues one atomic per stay, not one per byte.
__global__ void __launch_bounds__(256) VisitCountKernel(const __grid_constant__ ScanArgs a)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    uint64_t b, e;
    if (a.offsets) {
        b = a.offsets[i];
        e = a.offsets[i + 1] - a.trim;
        e = e < b ? b : e;
    } else {
        b = i * a.fixed_len;
        e = b + a.fixed_len;
    }
    uint32_t s = a.start;
    unsigned long long run = 0;
    for (uint64_t q = b; q < e; ++q) {
        uint32_t c = a.cls[a.corpus[q]];
        size_t at = (size_t) s * a.letters + c;
        uint32_t ns = a.wide ? static_cast<const uint32_t*>(a.full)[at] : (uint32_t) static_cast<const uint16_t*>(a.full)[at];
        ++run;
        if (ns != s) {
            atomicAdd(&a.visits[s], run);
            run = 0;
            s = ns;
        }
    }
    if (run)
        atomicAdd(&a.visits[s], run);
}

__global__ void __launch_bounds__(256) SynthKernel(const __grid_constant__ SynthParams p, const char* __restrict__ plants, uint8_t* __restrict__ out)
{
    const uint32_t pieces = p.string_len / 16;
    const uint64_t total = p.n_strings * pieces;
    const uint32_t words = p.string_len / 8;
    for (uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t local = idx / pieces;
        const uint32_t piece = (uint32_t) (idx % pieces);
        const uint64_t gi = p.first_string + local;
        uint64_t w0 = SynthWord(p.seed, gi, piece * 2, words);
        uint64_t w1 = SynthWord(p.seed, gi, piece * 2 + 1, words);
        uint32_t off = 0;
        int id = SynthPlant(p, gi, &off);
        if (id >= 0) {
            const uint32_t len = p.plant_off[id + 1] - p.plant_off[id];
            const uint32_t lo = piece * 16, hi = lo + 16;
            const bool touches = (off < hi && off + len > lo) || (p.tail && p.plant_mode[id] == 0 && hi == p.string_len);
            if (touches) {
                uint8_t bytes[16];
                for (int k = 0; k < 8; ++k) {
                    bytes[k] = (uint8_t) (w0 >> (8 * k));
                    bytes[8 + k] = (uint8_t) (w1 >> (8 * k));
                }
                for (uint32_t k = 0; k < 16; ++k)
                    bytes[k] = SynthByte(p, plants, gi, lo + k);
                w0 = w1 = 0;
                for (int k = 0; k < 8; ++k) {
                    w0 |= (uint64_t) bytes[k] << (8 * k);
                    w1 |= (uint64_t) bytes[8 + k] << (8 * k);
                }
            }
        }
        uint4 v = make_uint4((uint32_t) w0, (uint32_t) (w0 >> 32), (uint32_t) w1, (uint32_t) (w1 >> 32));
        *reinterpret_cast<uint4*>(out + local * (uint64_t) p.string_len + (uint64_t) piece * 16) = v;
    }
}

__global__ void __launch_bounds__(256) SynthMixedLengthsKernel(uint64_t seed, uint64_t first, uint64_t n, uint64_t* __restrict__ lengths)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        lengths[i] = SynthMixedLength(seed, first + i);
}

// One warp per string: lanes write consecutive 4-byte cells (coalesced).
__global__ void __launch_bounds__(256) SynthMixedFillKernel(uint64_t seed, uint32_t plant_every, uint64_t first, uint64_t n,
                                                            const uint64_t* __restrict__ offsets, uint8_t* __restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t) gridDim.x * (blockDim.x / 32);
    for (uint64_t i = (uint64_t) blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < n; i += warps) {
        const uint64_t b = offsets[i];
        const uint32_t len = (uint32_t) (offsets[i + 1] - b);
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + b);      // offsets are multiples of 4
        for (uint32_t cell = lane; cell < len / 4; cell += 32)
            dst[cell] = SynthMixedCellPlanted(seed, plant_every, first + i, len, cell);
    }
}

template <bool kPred>
const void* UniformKernelPtr() { return reinterpret_cast<const void*>(&ScanUniformKernel<kPred>); }
template <int kMode>
const void* GenericKernelPtr() { return reinterpret_cast<const void*>(&ScanGenericKernel<kMode>); }

int LookRegs()
{
    static const int regs = [] {
        const char* env = getenv("PIRE_B200_LOOK_REGS");
        return env && atoi(env) == 40 ? 40 : 48;
    }();
    return regs;
}

// PIRE_B200_LOOK_CLEAN=0 restores the two-LOP3 step of the look-ahead kernel (see LookProbe) for comparison.
bool LookClean()
{
    static const bool clean = [] {
        const char* env = getenv("PIRE_B200_LOOK_CLEAN");
        return !(env && atoi(env) == 0);
    }();
    return clean;
}

// Two strings per lane (ScanUniformLook2Kernel) is the default shape of the look-ahead variant; PIRE_B200_LOOK_ILP=1
// selects one string per lane (ScanUniformLookKernel).  PIRE_B200_LOOK_ILP_REGS=64|72|80 picks the register budget
// and with it the CTA shape (two CTAs of 512 / 448 / 384 threads per SM); measured 2.739 / 2.374 / 2.384 ms on the
// glued scan (profiles/r02_experiments_notes.txt).
int LookIlp()
{
    static const int ilp = [] {
        const char* env = getenv("PIRE_B200_LOOK_ILP");
        return env && atoi(env) == 1 ? 1 : 2;
    }();
    return ilp;
}
int LookIlpRegs()
{
    static const int regs = [] {
        const char* env = getenv("PIRE_B200_LOOK_ILP_REGS");
        const int v = env ? atoi(env) : 72;
        return v == 64 || v == 80 ? v : 72;
    }();
    return regs;
}

const void* KernelFor(int variant, bool uniform)
{
    if (variant == kVariantPriv && uniform)
        return reinterpret_cast<const void*>(&ScanUniformPrivKernel);
    if (variant == kVariantLook && uniform && LookIlp() == 2)
        return LookIlpRegs() == 64   ? reinterpret_cast<const void*>(&ScanUniformLook2Kernel<64>)
               : LookIlpRegs() == 80 ? reinterpret_cast<const void*>(&ScanUniformLook2Kernel<80>)
                                     : reinterpret_cast<const void*>(&ScanUniformLook2Kernel<72>);
    if ((variant == kVariantLook || variant == kVariantLook1) && uniform && LookClean())
        return LookRegs() == 48 ? reinterpret_cast<const void*>(&ScanUniformLookKernel<false, 48, true>)
                                : reinterpret_cast<const void*>(&ScanUniformLookKernel<false, 40, true>);
    if ((variant == kVariantLook || variant == kVariantLook1) && uniform)
        return LookRegs() == 48 ? reinterpret_cast<const void*>(&ScanUniformLookKernel<false, 48>)
                                : reinterpret_cast<const void*>(&ScanUniformLookKernel<false, 40>);
    if (variant == kVariantLook64 && uniform)
        return LookRegs() == 48 ? reinterpret_cast<const void*>(&ScanUniformLookKernel<true, 48>)
                                : reinterpret_cast<const void*>(&ScanUniformLookKernel<true, 40>);
    if (uniform)
        return variant == kVariantPred ? UniformKernelPtr<true>() : UniformKernelPtr<false>();
    if (variant == kVariantLook || variant == kVariantLook64 || variant == kVariantLook1)      // CSR batches: one look-ahead kernel (32-slot filter)
        return GenericKernelPtr<2>();
    return variant == kVariantPred ? GenericKernelPtr<1>() : GenericKernelPtr<0>();
}

} // namespace

size_t ScanSharedBytes(uint32_t hot, uint32_t priv_rows) { return PrivBytes(priv_rows) + HotBytes(hot) + 512 + 256 + 16; }
size_t GenericSharedBytes(uint32_t hot) { return ScanSharedBytes(hot, 0) + kStageBytes; }

cudaError_t PrepareScanKernels(int device)
{
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return err;
    int optin = 0;
    err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err != cudaSuccess)
        return err;
    for (int variant : {(int) kVariantPlain, (int) kVariantPred, (int) kVariantPriv, (int) kVariantLook, (int) kVariantLook64, (int) kVariantLook1})
        for (bool uniform : {false, true}) {
            err = cudaFuncSetAttribute(KernelFor(variant, uniform), cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
            if (err != cudaSuccess)
                return err;
            // three CTAs of ~75 KB each per SM: ask for the largest shared-memory carve-out (kernels without a
            // blocks-per-SM launch bound would otherwise get a smaller one and run two CTAs)
            err = cudaFuncSetAttribute(KernelFor(variant, uniform), cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
            if (err != cudaSuccess)
                return err;
        }
    return cudaSuccess;
}

cudaError_t PlanScan(int device, uint32_t hot, uint32_t hot_small, uint32_t priv_rows, int variant, bool uniform, LaunchPlan* plan)
{
    const bool priv = variant == kVariantPriv && uniform;
    const bool look = variant == kVariantLook || variant == kVariantLook64 || variant == kVariantLook1;
    plan->block = priv ? kPrivBlock : (look && uniform) ? (LookRegs() == 48 ? kLookBlock48 : kLookBlock40) : kBlock;
    if (look && uniform) {
        static const int look_block = [] {
            const char* env = getenv("PIRE_B200_LOOK_BLOCK");          // experiments: e.g. 640 = two CTAs of 20 warps at 48 registers
            return env && atoi(env) >= 32 && atoi(env) <= 1024 && atoi(env) % 32 == 0 ? atoi(env) : 0;
        }();
        if (variant == kVariantLook && LookIlp() == 2)
            plan->block = LookIlpRegs() == 64 ? 512 : LookIlpRegs() == 80 ? 384 : 448;
        if (look_block)
            plan->block = look_block;
    }
    plan->shared = priv ? ScanSharedBytes(hot_small, priv_rows) : uniform ? ScanSharedBytes(hot, 0) : GenericSharedBytes(hot);
    int sms = 0, per_sm = 0;
    cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess)
        return err;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KernelFor(variant, uniform), plan->block, plan->shared);
    if (err != cudaSuccess)
        return err;
    if (per_sm < 1)
        return cudaErrorLaunchOutOfResources;
    if (!uniform)
        per_sm = per_sm < kGenericBlocksPerSM ? per_sm : kGenericBlocksPerSM;      // a third CTA measured slower twice (r01, r02 notes)
    plan->grid = sms * per_sm;     // persistent: every SM holds its full share of CTAs
    return cudaSuccess;
}

cudaError_t LaunchScan(const ScanArgs& a, int variant, bool uniform, const LaunchPlan& plan, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    uint64_t units = (a.n + 31) / 32;
    if (variant == kVariantLook && uniform && LookIlp() == 2)
        units = (units + 1) / 2;              // a warp of ScanUniformLook2Kernel takes two units at a time
    const uint64_t warps_per_block = (uint64_t) plan.block / 32;
    uint64_t want = (units + warps_per_block - 1) / warps_per_block;
    int grid = (int) (want < (uint64_t) plan.grid ? want : (uint64_t) plan.grid);
    void* args[] = {const_cast<ScanArgs*>(&a)};
    cudaError_t err = cudaLaunchKernel(KernelFor(variant, uniform), dim3(grid), dim3(plan.block), args, plan.shared, stream);
    if (err == cudaSuccess)
        g_launches.fetch_add(1, std::memory_order_relaxed);
    return err;
}

// Length-ordered CSR batch: the leading long strings, one per warp (ScanSplitKernel).  a.split_count / a.split_counter
// are two device words; the generic kernel launched afterwards with the same a.split_count skips those strings.
cudaError_t LaunchSplit(const ScanArgs& a, int variant, int device, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    static const uint32_t split_min = [] {
        const char* env = getenv("PIRE_B200_SPLIT_MIN");          // experiments; a power of two keeps it a bucket boundary
        return env && atoi(env) >= 64 ? (uint32_t) atoi(env) : kSplitMin;
    }();
    SplitCountKernel<<<1, 32, 0, stream>>>(a.offsets, a.order, a.n, split_min, const_cast<uint32_t*>(a.split_count), a.split_counter);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return err;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    static const uint32_t split_prefetch = [] {
        const char* env = getenv("PIRE_B200_SPLIT_PREFETCH");     // blocks of 32 bytes between the walk and its L2 prefetch; 0 = none
        return env ? (uint32_t) atoi(env) : 0u;
    }();
    const void* fn = variant == kVariantPlain ? reinterpret_cast<const void*>(&ScanSplitKernel<false>)
                                              : reinterpret_cast<const void*>(&ScanSplitKernel<true>);
    const size_t shared = ScanSharedBytes(a.hot, 0) + kSplitMarkBytes + kWarpsPerBlock * 32;
    int optin = 0, sms = 0, per_sm = 0;
    err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err == cudaSuccess)
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err == cudaSuccess)
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, shared);
    if (err != cudaSuccess)
        return err;
    if (per_sm < 1)
        return cudaErrorLaunchOutOfResources;
    ScanArgs with_prefetch = a;                 // kernel arguments are copied at launch
    with_prefetch.split_prefetch = split_prefetch;
    void* args[] = {&with_prefetch};
    err = cudaLaunchKernel(fn, dim3(sms * per_sm), dim3(kBlock), args, shared, stream);
    if (err == cudaSuccess)
        g_launches.fetch_add(1, std::memory_order_relaxed);
    return err;
}

cudaError_t LaunchPrefix(const ScanArgs& a, bool shortest, bool reverse, int device, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    const void* fn = reverse ? (shortest ? reinterpret_cast<const void*>(&PrefixKernel<true, true>)
                                         : reinterpret_cast<const void*>(&PrefixKernel<false, true>))
                             : (shortest ? reinterpret_cast<const void*>(&PrefixKernel<true, false>)
                                         : reinterpret_cast<const void*>(&PrefixKernel<false, false>));
    int optin = 0, sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err == cudaSuccess)
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err != cudaSuccess)
        return err;
    if (a.uniform && !reverse) {
        // a.uniform: 1 = plain walk, 2 = exit filter
        static const int block = [] {
            const char* env = getenv("PIRE_B200_PREFIX_BLOCK");       // experiments; 640 = two CTAs of twenty warps at 48 registers
            return env && atoi(env) >= 32 && atoi(env) <= 1024 && atoi(env) % 32 == 0 ? atoi(env) : 640;
        }();
        const bool pred = a.uniform == 2;
        static const bool idp = [] {
            const char* env = getenv("PIRE_B200_PREFIX_IDP");          // experiments: byte extraction on the FMA pipe
            return env && atoi(env) != 0;
        }();
        if (pred)
            fn = shortest ? reinterpret_cast<const void*>(&PrefixUniformKernel<true, true, false>)
                          : reinterpret_cast<const void*>(&PrefixUniformKernel<false, true, false>);
        else if (idp)
            fn = shortest ? reinterpret_cast<const void*>(&PrefixUniformKernel<true, false, true>)
                          : reinterpret_cast<const void*>(&PrefixUniformKernel<false, false, true>);
        else
            fn = shortest ? reinterpret_cast<const void*>(&PrefixUniformKernel<true, false, false>)
                          : reinterpret_cast<const void*>(&PrefixUniformKernel<false, false, false>);
        const size_t shared = ScanSharedBytes(a.hot, 0) + 272;
        int per_sm = 0;
        err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err == cudaSuccess)
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, shared);
        if (err != cudaSuccess)
            return err;
        if (per_sm < 1)
            return cudaErrorLaunchOutOfResources;
        const uint64_t per_block = (uint64_t) block;
        const uint64_t want = (a.n + per_block - 1) / per_block;
        const int grid = (int) (want < (uint64_t) sms * per_sm ? want : (uint64_t) sms * per_sm);
        void* args[] = {const_cast<ScanArgs*>(&a)};
        err = cudaLaunchKernel(fn, dim3(grid), dim3(block), args, shared, stream);
        if (err == cudaSuccess)
            g_launches.fetch_add(1, std::memory_order_relaxed);
        return err;
    }
    const size_t shared = GenericSharedBytes(a.hot) + 272;      // + the hot states' flag bytes
    uint64_t want = (a.n + kBlock - 1) / kBlock;
    int grid = (int) (want < (uint64_t) sms * kGenericBlocksPerSM ? want : (uint64_t) sms * kGenericBlocksPerSM);
    void* args[] = {const_cast<ScanArgs*>(&a)};
    err = cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), args, shared, stream);
    if (err == cudaSuccess)
        g_launches.fetch_add(1, std::memory_order_relaxed);
    return err;
}


// Lines of text (CSR, PIRE_GPU_RUN_LINES, no order).  The caller zeroes the bitmap.  In stream (ScanTextKernel) when the
// start state is a hot row -- it practically always is; lanes pulling lines one by one (ScanLinesKernel) otherwise.
cudaError_t LaunchLines(const ScanArgs& a, int variant, int device, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    static const int forced = [] {
        const char* env = getenv("PIRE_B200_LINES_KERNEL");        // experiments: 1 = pulling lanes, 2 = in stream
        return env ? atoi(env) : 0;
    }();
    const bool in_stream = forced == 2 || (forced != 1 && a.start < a.hot);
    if (in_stream && !(a.start < a.hot))
        return cudaErrorInvalidValue;
    const bool pred = variant == kVariantPred || variant == kVariantLook || variant == kVariantLook64 || variant == kVariantLook1;
    const void* fn = in_stream ? (pred ? reinterpret_cast<const void*>(&ScanTextKernel<true>) : reinterpret_cast<const void*>(&ScanTextKernel<false>))
                               : (pred ? reinterpret_cast<const void*>(&ScanLinesKernel<true>) : reinterpret_cast<const void*>(&ScanLinesKernel<false>));
    const size_t shared = ScanSharedBytes(a.hot, 0) + (in_stream ? kTextFinBytes + kTextPackBytes : 0);
    int optin = 0, sms = 0, per_sm = 0;
    cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err == cudaSuccess)
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err == cudaSuccess)
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, shared);
    if (err != cudaSuccess)
        return err;
    if (per_sm < 1)
        return cudaErrorLaunchOutOfResources;
    ScanArgs tuned = a;
    int grid = sms * per_sm;
    if (in_stream) {
        tuned.text_segment = 0;                                        // chosen on the device from the text's size
        if (const char* env = getenv("PIRE_B200_TEXT_SEGMENT"))       // experiments
            tuned.text_segment = atoi(env) >= 32 && atoi(env) <= (1 << 20) ? (uint32_t) atoi(env) / 32u * 32u : tuned.text_segment;
        // the persistent grid is launched whole: the number of units depends on the text's size, which only the device
        // knows (offsets[n]); warps without a unit leave at once
    } else {
        const uint64_t groups = (a.n + kLinesPerWarp - 1) / kLinesPerWarp;
        const uint64_t want = (groups + kWarpsPerBlock - 1) / kWarpsPerBlock;
        grid = (int) (want < (uint64_t) grid ? want : (uint64_t) grid);
        tuned.lines_turn = kPiecesPerTurn;
        tuned.lines_min_idle = kLinesMinIdle;
        if (const char* env = getenv("PIRE_B200_LINES_TURN"))          // experiments
            tuned.lines_turn = atoi(env) > 0 ? (uint32_t) atoi(env) : tuned.lines_turn;
        if (const char* env = getenv("PIRE_B200_LINES_MIN_IDLE"))
            tuned.lines_min_idle = atoi(env) > 0 ? (uint32_t) atoi(env) : tuned.lines_min_idle;
    }
    void* args[] = {&tuned};
    err = cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), args, shared, stream);
    if (err == cudaSuccess)
        g_launches.fetch_add(1, std::memory_order_relaxed);
    return err;
}

cudaError_t LaunchCount(const ScanArgs& a, int device, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    const void* fn = nullptr;
    switch (a.count_words * 2 + (a.count_always ? 1 : 0)) {
    case 2: fn = reinterpret_cast<const void*>(&CountKernel<1, false>); break;
    case 3: fn = reinterpret_cast<const void*>(&CountKernel<1, true>); break;
    case 4: fn = reinterpret_cast<const void*>(&CountKernel<2, false>); break;
    case 5: fn = reinterpret_cast<const void*>(&CountKernel<2, true>); break;
    default: fn = reinterpret_cast<const void*>(&CountKernel<0, false>); break;
    }
    int optin = 0, sms = 0;
    cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err == cudaSuccess)
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err != cudaSuccess)
        return err;
    const size_t shared = GenericSharedBytes(a.hot) + (size_t) (a.hot + 1) * a.count_words * 8;
    const uint64_t want = ((a.n + 31) / 32 + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int grid = (int) (want < (uint64_t) sms * kGenericBlocksPerSM ? want : (uint64_t) sms * kGenericBlocksPerSM);
    void* args[] = {const_cast<ScanArgs*>(&a)};
    err = cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), args, shared, stream);
    if (err == cudaSuccess)
        g_launches.fetch_add(1, std::memory_order_relaxed);
    return err;
}

cudaError_t LaunchVisitCount(const ScanArgs& a, cudaStream_t stream)
{
    if (a.n == 0)
        return cudaSuccess;
    int grid = (int) ((a.n + 255) / 256);
    VisitCountKernel<<<grid, 256, 0, stream>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(256) LengthKeysKernel(const uint64_t* __restrict__ offsets, uint64_t n, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ ids)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        // Bucket = half an octave of length.  The sort is stable and only on the bucket, so
        // inside a bucket strings keep their corpus order: the 32 lanes of a warp then read
        // from neighbouring addresses (a full sort by length scatters them over the whole
        // corpus, which measured 4x slower than the imbalance it removes -- r01 experiments).
        uint64_t len = offsets[i + 1] - offsets[i];
        uint32_t l32 = (uint32_t) (len > 0xffffffffull ? 0xffffffffull : len);
        uint32_t msb = l32 ? 31u - (uint32_t) __clz(l32) : 0u;
        uint32_t half = msb ? (l32 >> (msb - 1)) & 1u : 0u;
        keys[i] = 255u - (msb * 2u + half);                                     // ascending keys = longest bucket first
        ids[i] = (uint32_t) i;
    }
}
} // namespace

// Stream-ordered scratch (work counters, sort/select temporaries) comes from a private per-device pool
// that keeps up to 64 MiB cached.  The device's default pool returns everything to the driver at each
// synchronisation, after which the next 4-byte counter allocation costs a 10-40 ms mapping stall -- seen as
// one slow step in twenty on the length-binned path.
cudaError_t ScratchAlloc(void** out, size_t bytes, cudaStream_t stream)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess)
        return err;
    if (device < 0 || device >= 64)
        return cudaMallocAsync(out, bytes, stream);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pools[device]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            err = cudaMemPoolCreate(&pools[device], &props);
            if (err != cudaSuccess)
                return err;
            uint64_t keep = 64ull << 20;
            err = cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
            if (err != cudaSuccess)
                return err;
        }
        pool = pools[device];
    }
    return cudaMallocFromPoolAsync(out, bytes, pool, stream);
}

cudaError_t LengthOrder(const uint64_t* d_offsets, uint64_t n, uint32_t* d_order, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    uint32_t *keys = nullptr, *keys_out = nullptr, *ids = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t err = ScratchAlloc((void**) &keys, n * 4, stream);
    if (err == cudaSuccess) err = ScratchAlloc((void**) &keys_out, n * 4, stream);
    if (err == cudaSuccess) err = ScratchAlloc((void**) &ids, n * 4, stream);
    if (err == cudaSuccess) {
        LengthKeysKernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(d_offsets, n, keys, ids);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess)
        err = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, keys_out, ids, d_order, (int) n, 0, 8, stream);
    if (err == cudaSuccess) err = ScratchAlloc(&temp, temp_bytes, stream);
    if (err == cudaSuccess)
        err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, ids, d_order, (int) n, 0, 8, stream);
    if (keys) cudaFreeAsync(keys, stream);
    if (keys_out) cudaFreeAsync(keys_out, stream);
    if (ids) cudaFreeAsync(ids, stream);
    if (temp) cudaFreeAsync(temp, stream);
    return err;
}

namespace {
struct IsNewline {
    const uint8_t* text;
    __host__ __device__ bool operator()(const unsigned long long& i) const { return text[i] == '\n'; }
};

__global__ void __launch_bounds__(256) CountNewlinesKernel(const uint8_t* __restrict__ text, uint64_t n_bytes, unsigned long long* __restrict__ count)
{
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += (uint64_t) gridDim.x * blockDim.x)
        local += text[i] == '\n';
    for (int d = 16; d; d >>= 1)
        local += __shfl_down_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local)
        atomicAdd(count, local);
}

__global__ void FinishLineOffsetsKernel(const uint8_t* __restrict__ text, uint64_t n_bytes, uint64_t* __restrict__ offsets,
                                        const unsigned long long* __restrict__ n_newlines, uint64_t capacity,
                                        unsigned long long* __restrict__ n_lines_out)
{
    // offsets[1..k] hold newline positions; turn them into the starts of the following lines
    const unsigned long long k = *n_newlines;
    const uint64_t idx = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < k && idx + 1 <= capacity)
        offsets[idx + 1] += 1;
    if (idx == 0) {
        offsets[0] = 0;
        unsigned long long lines = k;
        if (n_bytes != 0 && text[n_bytes - 1] != '\n') {     // last line has no newline: a virtual one after the end
            if (k + 1 <= capacity)
                offsets[k + 1] = n_bytes + 1;
            lines = k + 1;
        }
        *n_lines_out = lines;
    }
}
} // namespace

cudaError_t SplitLines(const uint8_t* d_text, uint64_t n_bytes, uint64_t* d_offsets, uint64_t capacity, uint64_t* n_lines,
                       cudaStream_t stream)
{
    *n_lines = 0;
    if (n_bytes == 0)
        return cudaSuccess;
    unsigned long long* d_counts = nullptr;       // [0] newlines, [1] lines
    void* temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t err = ScratchAlloc((void**) &d_counts, 2 * sizeof(unsigned long long), stream);
    cub::CountingInputIterator<unsigned long long> positions(0);
    IsNewline pred{d_text};
    unsigned long long* out = reinterpret_cast<unsigned long long*>(d_offsets + 1);
    // capacity guards: the select writes at most min(newlines, n_bytes) entries; the caller sizes d_offsets for the
    // worst case it accepts (capacity + 1 entries) and we verify after the fact.
    // first pass: how many lines?  (the select below writes every newline position, so the
    // caller's buffer must be known to be large enough before it runs)
    unsigned long long precount = 0;
    if (err == cudaSuccess)
        err = cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), stream);
    if (err == cudaSuccess) {
        uint64_t blocks = (n_bytes + 256 * 64 - 1) / (256 * 64);
        CountNewlinesKernel<<<(unsigned) (blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, stream>>>(d_text, n_bytes, d_counts);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(&precount, d_counts, sizeof(precount), cudaMemcpyDeviceToHost, stream);
    if (err == cudaSuccess)
        err = cudaStreamSynchronize(stream);
    if (err == cudaSuccess && (!d_offsets || precount + 1 > capacity)) {
        cudaFreeAsync(d_counts, stream);
        *n_lines = precount + 1;                       // upper bound of lines; caller retries with this capacity
        return cudaErrorInvalidValue;
    }
    if (err == cudaSuccess)
        err = cub::DeviceSelect::If(nullptr, temp_bytes, positions, out, d_counts, (long long) n_bytes, pred, stream);
    if (err == cudaSuccess)
        err = ScratchAlloc(&temp, temp_bytes, stream);
    if (err == cudaSuccess)
        err = cub::DeviceSelect::If(temp, temp_bytes, positions, out, d_counts, (long long) n_bytes, pred, stream);
    unsigned long long newlines = 0;
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(&newlines, d_counts, sizeof(newlines), cudaMemcpyDeviceToHost, stream);
    if (err == cudaSuccess)
        err = cudaStreamSynchronize(stream);
    if (err == cudaSuccess) {
        unsigned blocks = (unsigned) ((newlines + 255) / 256);
        FinishLineOffsetsKernel<<<blocks ? blocks : 1, 256, 0, stream>>>(d_text, n_bytes, d_offsets, d_counts, capacity, d_counts + 1);
        g_launches.fetch_add(2, std::memory_order_relaxed);
        err = cudaGetLastError();
    }
    unsigned long long lines = 0;
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(&lines, d_counts + 1, sizeof(lines), cudaMemcpyDeviceToHost, stream);
    if (err == cudaSuccess)
        err = cudaStreamSynchronize(stream);
    if (d_counts) cudaFreeAsync(d_counts, stream);
    if (temp) cudaFreeAsync(temp, stream);
    *n_lines = lines;
    return err;
}

cudaError_t LaunchSynth(const SynthParams& p, const char* d_plants, uint8_t* d_out, cudaStream_t stream)
{
    if (p.n_strings == 0)
        return cudaSuccess;
    uint64_t total = p.n_strings * (p.string_len / 16);
    uint64_t blocks = (total + 255) / 256;
    int grid = (int) (blocks < 148ull * 64 ? blocks : 148ull * 64);
    SynthKernel<<<grid, 256, 0, stream>>>(p, d_plants, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t LaunchSynthMixedLengths(uint64_t seed, uint64_t first, uint64_t n, uint64_t* d_lengths, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    SynthMixedLengthsKernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(seed, first, n, d_lengths);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t LaunchSynthMixedFill(uint64_t seed, uint32_t plant_every, uint64_t first, uint64_t n, const uint64_t* d_offsets,
                                 uint8_t* d_out, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    uint64_t blocks = (n + 7) / 8;
    SynthMixedFillKernel<<<(unsigned) (blocks < 148ull * 32 ? blocks : 148ull * 32), 256, 0, stream>>>(seed, plant_every, first, n,
                                                                                                   d_offsets, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

namespace {
// AcceptedRegexps of the state every string stopped in, as a bit set of `words` 32-bit words (multi.h:149-158 for
// scanners with more than 32 regexps; the scan kernels' accept mask holds ids 0..31 only).
__global__ void __launch_bounds__(256) AcceptGatherKernel(const uint32_t* __restrict__ table, uint32_t states, uint32_t words,
                                                          const uint32_t* __restrict__ state_idx, uint64_t n, uint32_t* __restrict__ out)
{
    const uint64_t total = n * words;
    for (uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t i = k / words;
        const uint32_t w = (uint32_t) (k % words);
        const uint32_t st = state_idx[i];
        out[k] = st < states ? __ldg(table + (size_t) st * words + w) : 0u;
    }
}
} // namespace

cudaError_t LaunchAcceptGather(const uint32_t* d_table, uint32_t states, uint32_t words, const uint32_t* d_state_idx, uint64_t n,
                               uint32_t* d_out, cudaStream_t stream)
{
    if (n == 0 || words == 0)
        return cudaSuccess;
    const uint64_t total = n * words;
    const uint64_t blocks = (total + 255) / 256;
    AcceptGatherKernel<<<(unsigned) (blocks < 148ull * 32 ? blocks : 148ull * 32), 256, 0, stream>>>(d_table, states, words, d_state_idx, n, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

uint64_t KernelLaunchCount() { return g_launches.load(std::memory_order_relaxed); }

} // namespace pire_b200
