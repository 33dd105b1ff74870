/* pire_b200.h -- C ABI of the B200-native Pire scan path.
 *
 * The reference (yandex/pire) has no plugin/FFI layer: its seam is the
 * compile-time "Scanner concept" consumed by the templates of pire/run.h
 * (SURVEY.md 8(b)).  This header is the boundary a maintainer binds instead:
 * every entry point names the reference interface it replaces.  Signatures use
 * plain pointers and sizes only; device buffers are caller-owned CUDA pointers.
 *
 * Compiled automata cross the boundary as the byte stream written by the
 * reference's own  Pire::Scanner::Save()  (pire/scanners/multi.h:557-573), so
 * the regex front end (Lexer -> Fsm -> Compile / Scanner::Glue) stays on the
 * host, unchanged.  See INTEGRATION.md for the reference-side binding and
 * include/pire_gpu.hpp for the C++ mirror of Scanner / Runner / Matches.
 *
 * All functions return 0 on success or a negative pire_gpu_status; the text of
 * the last error on the calling thread is available from pire_gpu_last_error().
 * The run path of the reference never throws and returns no status
 * (pire/run.h); here bad arguments and CUDA failures are reported instead of
 * being undefined behaviour.
 */
#ifndef PIRE_B200_H
#define PIRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pire_gpu_scanner pire_gpu_scanner;

typedef enum pire_gpu_status {
    PIRE_GPU_OK = 0,
    PIRE_GPU_EINVAL = -1,      /* bad argument */
    PIRE_GPU_EIMAGE = -2,      /* scanner image rejected (what Scanner::Load/Mmap throw, multi.h:252-272) */
    PIRE_GPU_ECUDA = -3,       /* CUDA runtime error (message holds cudaGetErrorString) */
    PIRE_GPU_ENODEVICE = -4,   /* no usable CUDA device: the scan path has NO CPU fallback */
    PIRE_GPU_EUNSUPPORTED = -5
} pire_gpu_status;

/* Run flags: which of RunHelper's mark steps surround the bytes (run.h:375-376).
 * Runner(sc).Begin().Run(p,n).End() == BEGIN|END; the free function
 * Pire::Matches(sc,b,e) (run.h:396-400) steps neither mark == 0. */
enum {
    PIRE_GPU_RUN_BEGIN = 1u,
    PIRE_GPU_RUN_END = 2u,
    /* CSR offsets come from pire_gpu_split_lines: string i ends one byte before offsets[i+1]
     * (its newline).  Accepted by every entry point that takes CSR offsets. */
    PIRE_GPU_RUN_LINES = 4u
};

/* Kernel variants (pire_gpu_scanner_set_variant). */
enum {
    PIRE_GPU_VARIANT_AUTO = 0,
    PIRE_GPU_VARIANT_PLAIN = 1,    /* one shared-memory load per byte, unconditional */
    PIRE_GPU_VARIANT_PRED = 2,     /* load predicated off while the resident state self-loops */
    PIRE_GPU_VARIANT_PRIV = 3,     /* hottest rows replicated per bank: conflict-free loads (fixed-length ASCII-heavy batches) */
    PIRE_GPU_VARIANT_LOOK = 4,     /* PRED with one byte of look-ahead: a resting lane reads the table only when this byte and
                                      the next can both matter (the device analogue of the ExitMasks skip loop,
                                      multi.h:966-989); fixed-length batches, PRED otherwise */
    PIRE_GPU_VARIANT_LOOK64 = 5,   /* LOOK with a 64-slot filter (one more FMA-pipe instruction per byte, fewer false passes) */
    PIRE_GPU_VARIANT_LOOK1 = 6,    /* LOOK walks two strings per lane on fixed-length batches (the second string's step fills
                                      the latency of the first one's table read); LOOK1 is the same filter with one string per
                                      lane -- the shape for batches too small to give every resident warp two units */
    PIRE_GPU_VARIANT_SLOTS = 8     /* length of per-variant arrays indexed by variant id */
};

typedef struct pire_gpu_info {
    uint32_t states;           /* Scanner::Size()          multi.h:134 */
    uint32_t letters;          /* Scanner::LettersCount()  multi.h:140 */
    uint32_t regexps;          /* Scanner::RegexpsCount()  multi.h:139 */
    uint32_t initial;          /* StateIndex(Initialize()) multi.h:161,:281 */
    uint32_t empty;            /* Scanner::Empty()         multi.h:135 */
    uint32_t hot_rows;         /* rows of the shared-memory table */
    uint32_t variant;          /* kernel variant in use */
    uint32_t tuned;            /* 1 after pire_gpu_scanner_tune */
    uint64_t table_bytes;      /* device bytes of the complete (L2-resident) table */
    uint64_t shared_bytes;     /* dynamic shared memory per CTA */
    int32_t  device;           /* CUDA device, or -1 for a host-only handle */
    uint32_t reserved;
} pire_gpu_info;

/* ---- scanner lifetime -----------------------------------------------------
 * Replaces: Scanner::Load / Scanner::Mmap (multi.h:244-279,:575-599) followed
 * by taking the address of the scanner for Runner() (run.h:388-389).
 * `image` is the Scanner::Save() stream of a Pire::Scanner (Relocatable; both
 * ExitMasks<2> and NoShortcuts variants are accepted).  The handle owns device
 * copies of its tables and does not retain `image`.  device >= 0 selects the
 * CUDA device; device == -1 builds a host-only handle (tables + the host
 * accessors below; every run entry point then fails with PIRE_GPU_ENODEVICE).
 * A handle is immutable after create/tune: run_batch is re-entrant across
 * streams, like a const Scanner shared between threads. */
int  pire_gpu_scanner_create(const void* image, size_t size, int device, pire_gpu_scanner** out);
void pire_gpu_scanner_destroy(pire_gpu_scanner* sc);
int  pire_gpu_scanner_info(const pire_gpu_scanner* sc, pire_gpu_info* out);
int  pire_gpu_scanner_set_variant(pire_gpu_scanner* sc, uint32_t variant);
int  pire_gpu_scanner_set_max_hot(pire_gpu_scanner* sc, uint32_t max_hot_rows);

/* ---- the hot path -----------------------------------------------------------
 * Replaces, for a whole batch of strings at once:
 *     Pire::Runner(sc).Begin().Run(ptr, len).End()          run.h:365-392
 *     -> operator bool / Final()                            run.h:380-381, multi.h:143
 *     -> AcceptedRegexps(state)                             multi.h:149-158
 *     -> StateIndex(state)                                  multi.h:281-284
 * String i is d_corpus[d_offsets[i] .. d_offsets[i+1])  (CSR, n+1 offsets), or,
 * when d_offsets == NULL, d_corpus[i*fixed_len .. (i+1)*fixed_len).  Empty
 * strings are legal (pire_ut.cpp NullPointer); n == 0 is a no-op.
 * Outputs (each may be NULL):
 *   d_match_bits    ceil(n/32) words; bit (i%32) of word i/32 = Final();
 *                   bits past n in the last word are 0
 *   d_accept_masks  n words; bit r = regexp id r in AcceptedRegexps (r < 32)
 *   d_state_idx     n words; StateIndex() of the last state, reference numbering
 * All pointers are device pointers on the handle's device; the launch is
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default).
 * Side effect shared by every entry point that takes a handle, a communicator or a
 * `device` argument: the device becomes the calling thread's current CUDA device
 * (cudaSetDevice) and stays so on return; a caller that works with several devices
 * from one thread re-selects its own afterwards. */
int pire_gpu_run_batch(const pire_gpu_scanner* sc,
                       const uint8_t* d_corpus, const uint64_t* d_offsets,
                       uint64_t fixed_len, uint64_t n, uint32_t flags,
                       uint32_t* d_match_bits, uint32_t* d_accept_masks, uint32_t* d_state_idx,
                       void* stream);

/* Length-binned form of pire_gpu_run_batch for batches of very unequal strings
 * (BASELINE config 4: 16 B .. 64 KiB).  One string per lane means a warp runs as
 * long as its longest string; with `d_order` (a permutation of 0..n-1 from
 * pire_gpu_length_order: longest half-octave length bucket first, corpus order
 * inside a bucket so that a warp's lanes read neighbouring addresses) warps get
 * strings of similar length and claim them longest-first.  Results are still
 * indexed by the original string number.
 * CSR batches only; n < 2^32. */
int pire_gpu_length_order(const uint64_t* d_offsets, uint64_t n, uint32_t* d_order, int device, void* stream);
int pire_gpu_run_batch_ordered(const pire_gpu_scanner* sc,
                               const uint8_t* d_corpus, const uint64_t* d_offsets, const uint32_t* d_order,
                               uint64_t n, uint32_t flags,
                               uint32_t* d_match_bits, uint32_t* d_accept_masks, uint32_t* d_state_idx,
                               void* stream);

/* Replaces, per string of a batch,
 *     Pire::LongestPrefix (sc, begin, end, throughBeginMark, throughEndMark)   run.h:277-292
 *     Pire::ShortestPrefix(sc, begin, end, throughBeginMark, throughEndMark)   run.h:294-311
 * d_prefix_len[i] = length of the longest / shortest prefix of string i that the
 * scanner accepts, or PIRE_GPU_NO_PREFIX when there is none (the reference returns a
 * null pointer).  flags: PIRE_GPU_RUN_BEGIN = throughBeginMark, PIRE_GPU_RUN_END =
 * throughEndMark.  The scan stops in a dead state (pire_ut.cpp ScanTermination).
 * Semantics are those of the byte-by-byte predicates (run.h:69-100), i.e. of the
 * NoMask scanner variants: the ExitMasks fast-forward of the reference skips the
 * predicate for the bytes it jumps over, which can leave LongestPrefix short when a
 * string ends exactly on a 16-byte boundary of the host address space. */
#define PIRE_GPU_NO_PREFIX 0xFFFFFFFFu
int pire_gpu_prefix_batch(const pire_gpu_scanner* sc,
                          const uint8_t* d_corpus, const uint64_t* d_offsets,
                          uint64_t fixed_len, uint64_t n, uint32_t flags, int shortest,
                          uint32_t* d_prefix_len, void* stream);

/* Replaces, per string of a batch, the run of a Pire::HalfFinalScanner (pire/scanners/half_final.h):
 *     HalfFinalScanner::State st;  sc.Initialize(st);            half_final.h:136-141
 *     [Pire::Step(sc, st, BeginMark);]  Pire::Run(sc, st, begin, end);  [Pire::Step(sc, st, EndMark);]
 *     st.Result(r) for every regexp r                             half_final.h:88-90
 * (the driver of tests/count_ut.cpp:54-63).  TakeAction (half_final.h:154-163) adds one to the counter of
 * every regexp listed for a state each time the walk enters that state while it is final -- so Result(r)
 * counts the positions where a match of regexp r ends (HalfFinalFsm's counters, half_final_fsm.h:11-20),
 * and AcceptedRegexps(st) is { r : Result(r) != 0 }.
 * The image is the Save() stream of the HalfFinalScanner (it inherits Scanner::Save; the same format).
 * d_counts: n rows of max(1, regexps) u32, row i for string i (overwritten).  d_match_bits: packed
 * Final(st) per string, may be null.  Counters are 32 bits wide (the reference's are size_t). */
/* Replaces, per string of a batch,
 *     Pire::LongestSuffix (sc, rbegin, rend, throughEndMark, throughBeginMark)   run.h:316-342
 *     Pire::ShortestSuffix(sc, rbegin, rend, throughEndMark, throughBeginMark)   run.h:345-362
 * with rbegin = the string's last byte and rend = one before its first: the scanner (normally compiled from
 * Fsm::Reverse()) is walked over the string from right to left.  d_suffix_len[i] = length of the longest /
 * shortest suffix accepted that way (the reference returns the pointer rbegin - length), or
 * PIRE_GPU_NO_PREFIX for its null.  PIRE_GPU_RUN_END = throughEndMark (stepped before the bytes),
 * PIRE_GPU_RUN_BEGIN = throughBeginMark (stepped after them).  ShortestSuffix's quirk is kept: with
 * throughBeginMark the mark is stepped from the state the scan stopped in, and the answer is null unless that
 * state is final (run.h:357-360). */
int pire_gpu_suffix_batch(const pire_gpu_scanner* sc,
                          const uint8_t* d_corpus, const uint64_t* d_offsets,
                          uint64_t fixed_len, uint64_t n, uint32_t flags, int shortest,
                          uint32_t* d_suffix_len, void* stream);

/* How pire_gpu_count_batch keeps the counters (results are identical; for tests and measurements).
 * AUTO: packed per-state increments when the automaton has at most 16 regexps, behind a look-ahead pass
 * that skips chunks without final states -- or on every chunk when pire_gpu_scanner_tune saw more than
 * 2.5 % of the sample's steps end in a final state; the accept lists otherwise. */
#define PIRE_GPU_COUNT_AUTO        0u
#define PIRE_GPU_COUNT_LISTS       1u
#define PIRE_GPU_COUNT_PACKED      2u
#define PIRE_GPU_COUNT_EVERY_CHUNK 3u
int pire_gpu_scanner_set_count_mode(pire_gpu_scanner* sc, uint32_t mode);

int pire_gpu_count_batch(const pire_gpu_scanner* sc,
                         const uint8_t* d_corpus, const uint64_t* d_offsets,
                         uint64_t fixed_len, uint64_t n, uint32_t flags,
                         uint32_t* d_counts, uint32_t* d_match_bits, void* stream);

/* The step before the path for line-oriented input (samples/pigrep/pigrep.cpp:38-45 calls
 * std::getline and then Runner(sc).Begin().Run(line).End() per line).
 * pire_gpu_split_lines finds the lines of a newline-delimited text resident in HBM:
 *   d_line_offsets[0..*n_lines] (capacity + 1 entries available), line i =
 *   d_text[off[i] .. off[i+1] - 1) -- the newline itself is excluded, a last line without
 *   newline is kept, an empty text has no lines (std::getline semantics).
 *   If capacity is too small (or d_line_offsets is NULL) nothing is written, *n_lines
 *   receives a sufficient capacity and PIRE_GPU_EINVAL is returned.  Synchronises `stream`.
 * pire_gpu_run_lines scans those lines; outputs as in pire_gpu_run_batch.  d_line_offsets must be the offsets
 *   pire_gpu_split_lines produced for d_text: without d_order the lines are scanned where they lie -- every lane
 *   walks a few KiB of the text and restarts behind each newline it meets -- so line i + 1 has to start right behind the
 *   newline that ends line i.  With d_order (pire_gpu_length_order; usually slower for lines) they are scanned one
 *   string per lane like any CSR batch. */
int pire_gpu_split_lines(const uint8_t* d_text, uint64_t n_bytes, uint64_t* d_line_offsets, uint64_t capacity,
                         uint64_t* n_lines, int device, void* stream);
int pire_gpu_run_lines(const pire_gpu_scanner* sc, const uint8_t* d_text, const uint64_t* d_line_offsets,
                       const uint32_t* d_order, uint64_t n_lines, uint32_t flags,
                       uint32_t* d_match_bits, uint32_t* d_accept_masks, uint32_t* d_state_idx, void* stream);

/* Same call with HOST buffers -- what a Pire user holds: Run(const char* begin, const char* end) takes pageable
 * memory (run.h:271-275; samples/pigrep/pigrep.cpp:38-45).  The corpus is cut into chunks of whole 32-string
 * units (about 64 MiB; PIRE_B200_HOST_CHUNK_MB) and streamed through a ring of three device slots: the
 * host->device copy of chunk k+1 overlaps the scan of chunk k and the device->host copy of its results.
 * Pageable input is staged through the library's own pinned buffers by a few copy threads
 * (PIRE_B200_HOST_THREADS, default min(8, cores / 4)); pinned or cudaHostRegister-ed input is DMA-ed straight
 * from the caller's buffer.  The device never holds more than the ring, so corpora larger than HBM stream through.
 * CSR batches are length-binned per chunk (pire_gpu_length_order).  Results are those of pire_gpu_run_batch on
 * the resident corpus.  Returns after everything has landed in the caller's arrays.
 * corpus_bytes = size of the corpus buffer; offsets (if any) must ascend and end within it, n * fixed_len
 * must fit in it (PIRE_GPU_EINVAL otherwise).
 * Concurrency: calls on one handle from several threads run concurrently, each with its own workspace
 * (streams, slots, staging); workspaces are kept with the handle and freed by pire_gpu_scanner_destroy. */
int pire_gpu_run_batch_host(const pire_gpu_scanner* sc,
                            const uint8_t* corpus, uint64_t corpus_bytes, const uint64_t* offsets,
                            uint64_t fixed_len, uint64_t n, uint32_t flags,
                            uint32_t* match_bits, uint32_t* accept_masks, uint32_t* state_idx);

/* ---- several GPUs of one box ---------------------------------------------------
 * The path shards by string: rank r of `world` scans the contiguous shard
 * pire_gpu_shard_bounds(n_global, world, r) -- boundaries on multiples of 32 strings, so bitmap words never straddle
 * ranks -- straight into its slot of the full-length match bitmap, and one in-place ncclAllGather of the equal-sized
 * slots completes the bitmap on every rank (1/world of the bytes of an all-reduce of a zeroed bitmap and no zeroing;
 * the shards are disjoint and word aligned, so the result is the same OR).  The reference's callers are C++
 * (tools/bench/bench.cpp:241-254, samples/pigrep/pigrep.cpp:38-45): this is their multi-GPU entry, no Python or
 * PyTorch involved.  NCCL is bound at run time (libnccl.so.2 of the process, else the system's); single-GPU users carry
 * no NCCL dependency, and these calls return PIRE_GPU_EUNSUPPORTED where NCCL is absent.
 *
 * One communicator per rank (process or thread; one GPU each):
 *   rank 0: pire_gpu_comm_get_id(id)  -> ship the PIRE_GPU_COMM_ID_BYTES bytes to every rank by any means
 *   every rank: pire_gpu_comm_create(id, world, rank, device, &comm)      (collective, like ncclCommInitRank)
 *   or pire_gpu_comm_adopt(ncclComm_t, device, &comm) around a communicator the caller already owns.
 * pire_gpu_run_sharded: d_corpus / d_offsets / d_accept_masks / d_state_idx describe THIS RANK's shard only (string
 *   0 of the buffers is string lo of the batch); d_match_bits_all has pire_gpu_sharded_words(n_global, world) words
 *   (>= ceil(n_global / 32); bits past n_global are zero) and is identical on every rank afterwards.  The scan runs
 *   on `stream`, the exchange on the communicator's own stream behind it.  Default: `stream` then waits for the
 *   exchange (stream-ordered, like any other call).  With PIRE_GPU_RUN_ASYNC_EXCHANGE the exchange is left running so
 *   that the caller's next work on `stream` (the next batch's scan into ANOTHER bitmap buffer) overlaps it;
 *   pire_gpu_comm_wait(comm, stream) makes `stream` wait for the last exchange before the bitmap is read.
 *   A later run_sharded on the same communicator orders itself after the pending exchange. */
#define PIRE_GPU_COMM_ID_BYTES 128
#define PIRE_GPU_RUN_ASYNC_EXCHANGE 8u
typedef struct pire_gpu_comm pire_gpu_comm;
void     pire_gpu_shard_bounds(uint64_t n_global, int world, int rank, uint64_t* lo, uint64_t* hi);
uint64_t pire_gpu_sharded_words(uint64_t n_global, int world);
int  pire_gpu_comm_get_id(void* id_out);
int  pire_gpu_comm_create(const void* id, int world, int rank, int device, pire_gpu_comm** out);
int  pire_gpu_comm_adopt(void* nccl_comm, int device, pire_gpu_comm** out);
void pire_gpu_comm_destroy(pire_gpu_comm* comm);
int  pire_gpu_comm_info(const pire_gpu_comm* comm, int* world, int* rank);
int  pire_gpu_comm_wait(pire_gpu_comm* comm, void* stream);
/* the exchange alone, for a slot filled some other way (e.g. uploaded after pire_gpu_run_batch_host); flags: 0 or
 * PIRE_GPU_RUN_ASYNC_EXCHANGE */
int  pire_gpu_comm_gather_bits(pire_gpu_comm* comm, uint64_t n_global, uint32_t* d_match_bits_all, uint32_t flags, void* stream);
int  pire_gpu_run_sharded(const pire_gpu_scanner* sc, pire_gpu_comm* comm,
                          const uint8_t* d_corpus, const uint64_t* d_offsets, uint64_t fixed_len,
                          uint64_t n_global, uint32_t flags,
                          uint32_t* d_match_bits_all, uint32_t* d_accept_masks, uint32_t* d_state_idx, void* stream);

/* Re-selects the shared-memory hot rows from the states a device-resident
 * sample of the workload actually visits (the reference has no counterpart; it
 * relies on the CPU cache to keep hot rows close).  Only speed depends on it.
 * Synchronises the device.  Not thread-safe against concurrent runs. */
int pire_gpu_scanner_tune(pire_gpu_scanner* sc,
                          const uint8_t* d_corpus, const uint64_t* d_offsets,
                          uint64_t fixed_len, uint64_t n_sample, uint32_t flags, void* stream);

/* Times every kernel variant that can serve this batch shape (two launches each,
 * results discarded) and makes the fastest one the handle's AUTO choice for that
 * shape (fixed-length/aligned vs generic).  ms_out[PIRE_GPU_VARIANT_SLOTS] (may be NULL) receives the
 * milliseconds per variant id, 0 for variants that do not apply.  Which variant
 * wins depends on the automaton and the text, so it is
 * measured rather than guessed.  Synchronises the device. */
int pire_gpu_scanner_autoselect(pire_gpu_scanner* sc,
                                const uint8_t* d_corpus, const uint64_t* d_offsets,
                                uint64_t fixed_len, uint64_t n, uint32_t flags, void* stream, float* ms_out);

/* AcceptedRegexps for scanners with MORE than 32 regexps (multi.h:149-158 returns a list of any length; Glue only
 * caps the states, multi.h:1092-1103).  The scan entry points' d_accept_masks holds ids 0..31; the complete answer
 * comes from the state each string stopped in: run with d_state_idx, then
 *     pire_gpu_accept_sets(sc, d_state_idx, n, d_sets, stream)
 * writes n rows of pire_gpu_accept_words(sc) = ceil(max(1, RegexpsCount()) / 32) words, bit (r % 32) of word
 * r / 32 of row i set iff regexp r is in AcceptedRegexps of string i's last state (after End() when the run
 * stepped it).  A state index outside the scanner yields an empty set. */
uint32_t pire_gpu_accept_words(const pire_gpu_scanner* sc);
int pire_gpu_accept_sets(const pire_gpu_scanner* sc, const uint32_t* d_state_idx, uint64_t n,
                         uint32_t* d_accept_sets, void* stream);

/* Number of kernels this library has launched in the calling process. */
uint64_t pire_gpu_launch_count(void);

/* ---- host-side Scanner concept on the flattened tables ------------------------
 * Index-space mirror of the concept the templates in run.h consume
 * (multi.h:137-194,:281-284); used by include/pire_gpu.hpp's HostScanner and by
 * the parity tests.  States are StateIndex values of the reference. */
uint32_t pire_gpu_initial(const pire_gpu_scanner* sc);                                   /* Initialize  multi.h:161 */
uint32_t pire_gpu_next(const pire_gpu_scanner* sc, uint32_t state, uint32_t ch);         /* Next        multi.h:189-192; ch in 0..259 */
int      pire_gpu_final(const pire_gpu_scanner* sc, uint32_t state);                     /* Final       multi.h:143 */
int      pire_gpu_dead(const pire_gpu_scanner* sc, uint32_t state);                      /* Dead        multi.h:147 */
size_t   pire_gpu_accepted_regexps(const pire_gpu_scanner* sc, uint32_t state,
                                   uint32_t* ids, size_t cap);                           /* AcceptedRegexps multi.h:149-158 */

/* ---- deterministic synthetic corpora (bench + tests) ---------------------------
 * Counter-based generator, identical bytes on host and device.
 * kind 0: printable ASCII 0x20..0x7E; every `plant_every`-th string carries one
 *         planted literal, plants[(i / plant_every) % n_plants].  A literal
 *         starting with '^' is placed at the start of the string, one starting
 *         with '$' at its end (for anchored patterns; the marker itself is not
 *         written), any other at a pseudo-random offset, and then the string's
 *         last byte is set to `tail` when tail != 0.
 * `plants` is n_plants NUL-terminated literals back to back (plants_bytes in all). */
typedef struct pire_gpu_synth {
    uint64_t seed;
    uint64_t first_string;     /* global index of string 0 of this buffer (sharding) */
    uint64_t n_strings;
    uint32_t string_len;       /* fixed length, multiple of 16 */
    uint32_t kind;
    uint32_t plant_every;      /* 0 = never */
    uint32_t n_plants;
    const char* plants;
    uint32_t plants_bytes;
    uint32_t tail;
} pire_gpu_synth;

int pire_gpu_synth_fill_device(const pire_gpu_synth* spec, uint8_t* d_corpus, int device, void* stream);
int pire_gpu_synth_fill_host(const pire_gpu_synth* spec, uint8_t* corpus, uint64_t first, uint64_t count);
/* strings indices[0..count) of the corpus (relative to spec->first_string), back to back: stratified parity samples */
int pire_gpu_synth_fill_host_indexed(const pire_gpu_synth* spec, uint8_t* corpus, const uint64_t* indices, uint64_t count);

/* kind 1: mixed-length UTF-8 corpus (BASELINE config 4): lengths log-uniform in
 * [16, 65536), multiples of 4; ASCII / 2-byte Cyrillic / 3-byte code points; every
 * plant_every-th string (>= 32 bytes) ends with a mixed-case hit for the
 * case-insensitive headline pattern.  *_lengths_* writes n lengths (the caller
 * prefix-sums them into CSR offsets); *_fill_* writes the bytes for given offsets. */
int pire_gpu_synth_mixed_lengths_device(uint64_t seed, uint64_t first_string, uint64_t n, uint64_t* d_lengths, int device, void* stream);
int pire_gpu_synth_mixed_lengths_host(uint64_t seed, uint64_t first_string, uint64_t n, uint64_t* lengths);
int pire_gpu_synth_mixed_fill_device(uint64_t seed, uint32_t plant_every, uint64_t first_string, uint64_t n,
                                     const uint64_t* d_offsets, uint8_t* d_corpus, int device, void* stream);
int pire_gpu_synth_mixed_fill_host(uint64_t seed, uint32_t plant_every, uint64_t first_string, uint64_t n,
                                   const uint64_t* offsets, uint8_t* corpus);

const char* pire_gpu_last_error(void);
const char* pire_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
