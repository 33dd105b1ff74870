// scan_kernels.cuh -- launch interface between the C ABI (capi.cu) and the
// sm_100a kernels (scan_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dfa_tables.hpp"
#include "synth.h"

namespace pire_b200 {

struct DeviceFin {
    uint32_t result;    // bit31 Final, low bits StateIndex (reference numbering)
    uint32_t mask;      // accepted regexp ids < 32
};

// Everything a scan launch reads.  Device pointers.
struct ScanArgs {
    const uint8_t* corpus;
    const uint64_t* offsets;     // CSR (n+1) or nullptr
    uint32_t trim;               // CSR only: bytes dropped from the end of every string (1 = the newline of a line)
    const uint32_t* order;       // generic kernel: lane i scans string order[i] (length-binned launch), or nullptr
    unsigned int* work_counter;  // with `order`: units are claimed longest-first from this counter
    const uint32_t* split_count; // with `order`, or null: the first *split_count entries of `order` belong to the split kernel
    unsigned int* split_counter; // split kernel: strings are claimed from this counter
    uint32_t split_prefetch;     // split kernel: 32-byte blocks between a lane's walk and its L2 prefetch (0 = none)
    uint64_t fixed_len;          // used when offsets == nullptr
    uint64_t n;                  // strings
    const uint8_t* hot8;         // HotTableBytes(hot) bytes (rows kHotStride apart), 16-byte aligned
    const uint8_t* noexit;       // hot+1 bytes
    const uint16_t* cls;         // 256
    const void* full;            // states*letters, u16 or u32
    const DeviceFin* fin;        // [states] for the chosen with_end
    uint32_t hot;                // H
    uint32_t letters;
    uint32_t wide;               // full is u32
    uint32_t start;              // state after Initialize()[+Begin()], new numbering
    uint32_t exit_bitmap0;       // 32-slot exit bitmap of hot id 0, slot = byte & 31
    uint32_t look_bitmap;        // LOOK variant: 32-slot look-ahead filter (dfa_tables.hpp), slot = byte & 31
    uint64_t look_bitmap64;      // LOOK64 variant: the same filter with 64 slots, slot = byte & 63
    uint32_t uniform;            // prefix / count kernels: fixed length, a multiple of 32 bytes, corpus 32-byte aligned
    uint32_t opaque_zero;        // always 0; the LOOK kernels multiply by it to pin an instruction behind the walk
    const uint32_t* priv_packed; // (priv_rows/4)*128 words, PRIV variant
    uint32_t priv_rows;
    const uint8_t* hot8_small;   // PRIV variant's second tier: HotTableBytes(hot_small)
    uint32_t hot_small;
    uint32_t* match_bits;        // may be null
    uint32_t* accept_masks;      // may be null
    uint32_t* state_idx;         // may be null
    unsigned long long* visits;  // tune kernel only: per-state visit counters (new numbering)
    const uint8_t* flags;        // prefix kernels: [states] bit0 Final, bit1 Dead (new numbering)
    uint32_t end_class;          // prefix kernels: letter class of EndMark
    uint32_t through_end;        // prefix kernels: step EndMark after the bytes
    uint32_t* prefix_len;        // prefix kernels: n words, 0xFFFFFFFF = no accepted prefix
    // counting kernel (HalfFinalScanner::TakeAction, half_final.h:154-163)
    const uint32_t* acc_begin;   // [states + 1] CSR of accept lists, new numbering
    const uint32_t* acc_ids;
    uint32_t first_final_hot;    // hot ids >= this are final states
    uint32_t begin_class;        // letter class of BeginMark
    uint32_t initial;            // Initialize()'s state, new numbering
    uint32_t with_begin;         // step BeginMark first (with_end == through_end)
    uint32_t regexps;            // counters per string (>= 1)
    uint32_t* counts;            // n * regexps words, zeroed by the caller
    uint32_t lines_turn;         // lines kernel: chunks per lane between two hand-outs of lines
    uint32_t lines_min_idle;     // lines kernel: waiting lanes needed for a hand-out
    uint32_t text_segment;       // in-stream lines kernel: bytes of text per lane and unit, a multiple of 32
    const uint64_t* weights;     // [states * count_words] packed per-state increments, or null
    uint32_t count_words;        // 0 = walk the accept lists, 1..2 = packed increments
    uint32_t count_always;       // final states are frequent: count every chunk, skip the look-ahead pass
};

struct LaunchPlan {
    int block = 0;
    int grid = 0;
    size_t shared = 0;
};

enum ScanVariant { kVariantPlain = 1, kVariantPred = 2, kVariantPriv = 3, kVariantLook = 4, kVariantLook64 = 5, kVariantLook1 = 6 };
constexpr int kVariantSlots = 8;      // size of per-variant arrays (variant ids are 1-based)

size_t ScanSharedBytes(uint32_t hot, uint32_t priv_rows);
cudaError_t PrepareScanKernels(int device);                       // raises the dynamic smem limit
cudaError_t PlanScan(int device, uint32_t hot, uint32_t hot_small, uint32_t priv_rows, int variant, bool uniform, LaunchPlan* plan);
cudaError_t LaunchScan(const ScanArgs& a, int variant, bool uniform, const LaunchPlan& plan, cudaStream_t stream);
// CSR batches of short strings (lines of text): lanes pull strings dynamically; a.match_bits must be zeroed
cudaError_t LaunchLines(const ScanArgs& a, int variant, int device, cudaStream_t stream);
// length-ordered CSR batches: the leading long strings, one per warp; sets *a.split_count, which the generic launch honours
cudaError_t LaunchSplit(const ScanArgs& a, int variant, int device, cudaStream_t stream);
cudaError_t LaunchVisitCount(const ScanArgs& a, cudaStream_t stream);
// prefix (left to right) or suffix (right to left) scan; a.with_begin/begin_class name the mark stepped first,
// a.through_end/end_class the mark stepped last
cudaError_t LaunchPrefix(const ScanArgs& a, bool shortest, bool reverse, int device, cudaStream_t stream);
cudaError_t LaunchCount(const ScanArgs& a, int device, cudaStream_t stream);
// d_order <- string indices, longest half-octave length bucket first, corpus order inside a bucket (stable CUB radix sort).
// stream-ordered scratch from the library's own per-device pool (see scan_kernels.cu)
cudaError_t ScratchAlloc(void** out, size_t bytes, cudaStream_t stream);
cudaError_t LengthOrder(const uint64_t* d_offsets, uint64_t n, uint32_t* d_order, cudaStream_t stream);
// Line starts of a newline-delimited text (std::getline semantics): d_offsets[0..n_lines], line i =
// text[off[i] .. off[i+1] - 1).  *n_lines is written on the host after a stream synchronise.
cudaError_t SplitLines(const uint8_t* d_text, uint64_t n_bytes, uint64_t* d_offsets, uint64_t capacity, uint64_t* n_lines,
                       cudaStream_t stream);
cudaError_t LaunchSynth(const SynthParams& p, const char* d_plants, uint8_t* d_out, cudaStream_t stream);

cudaError_t LaunchSynthMixedLengths(uint64_t seed, uint64_t first, uint64_t n, uint64_t* d_lengths, cudaStream_t stream);
cudaError_t LaunchSynthMixedFill(uint64_t seed, uint32_t plant_every, uint64_t first, uint64_t n, const uint64_t* d_offsets,
                                 uint8_t* d_out, cudaStream_t stream);

// out[i * words + w] = table[state_idx[i] * words + w] (a state index outside the table yields zeros)
cudaError_t LaunchAcceptGather(const uint32_t* d_table, uint32_t states, uint32_t words, const uint32_t* d_state_idx, uint64_t n,
                               uint32_t* d_out, cudaStream_t stream);

uint64_t KernelLaunchCount();

} // namespace pire_b200
