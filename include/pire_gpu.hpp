// pire_gpu.hpp -- header-only C++ mirror of Pire's Scanner / Runner / Matches
// surface over the C ABI of include/pire_b200.h.
//
// For a code base that already uses Pire:
//
//     Pire::Scanner sc = Pire::Lexer("hello\\s+w.+d$").Parse().Surround().Compile<Pire::Scanner>();
//     Pire::Gpu::Scanner gsc(sc, /*device*/ 0);              // Scanner::Save() -> device tables
//     Pire::Gpu::Batch batch{d_corpus, d_offsets, 0, n};     // strings resident in HBM
//     Pire::Gpu::BatchRunner r = Pire::Gpu::Runner(gsc);
//     r.Begin().Run(batch).End();                            // run.h:365-392, for the whole batch
//     r.Matches(i);  r.AcceptedRegexps(i);
//
// Pire::Gpu::Scanner also satisfies the reference's compile-time "Scanner concept"
// (pire/scanners/multi.h:137-194,:281-284) on the HOST in index space, so the
// reference's own templates -- Pire::Step, Pire::Run, Pire::Runner, LongestPrefix,
// ShortestPrefix (pire/run.h) -- compile against it unchanged; the parity tests use
// that to prove the ingest is lossless.  The host concept is for verification; the
// batch path never falls back to it.
//
// This header does not include any Pire header: the templated constructor only
// needs `sc.Save(std::ostream*)`.
#pragma once

#include <cstddef>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pire_b200.h"

namespace Pire {
namespace Gpu {

// Counterpart of Pire::Error (pire/stub/stl.h:213-217).
class Error : public std::runtime_error {
public:
    Error(int code, const std::string& what) : std::runtime_error(what), Code(code) {}
    int Code;
};

inline void Check(int rc, const char* where)
{
    if (rc != PIRE_GPU_OK)
        throw Error(rc, std::string(where) + ": " + pire_gpu_last_error());
}

// A batch of strings on the device: CSR offsets (n+1) or fixed length.
struct Batch {
    const uint8_t* Corpus;
    const uint64_t* Offsets;     // nullptr => fixed-length strings
    uint64_t FixedLen;
    uint64_t Count;
};

class Scanner {
public:
    // ---- Scanner concept (host, index space) --------------------------------
    typedef uint32_t State;
    typedef uint32_t Action;
    typedef unsigned short Char;     // Pire::Char, pire/defs.h:59

    Scanner() : Handle(nullptr), AcceptBegin(1, 0) {}

    // From a Scanner::Save() stream (multi.h:557-573).  device = -1: host only.
    Scanner(const void* image, size_t size, int device = 0) : Handle(nullptr)
    {
        Check(pire_gpu_scanner_create(image, size, device, &Handle), "pire_gpu_scanner_create");
        BuildAcceptCache();
    }

    // From any Pire scanner type whose Save() writes the multi-Scanner format
    // (Pire::Scanner, Pire::NonrelocScanner and their NoMask variants).
    template <class PireScanner>
    explicit Scanner(const PireScanner& sc, int device = 0) : Handle(nullptr)
    {
        std::ostringstream out;
        sc.Save(&out);
        const std::string image = out.str();
        Check(pire_gpu_scanner_create(image.data(), image.size(), device, &Handle), "pire_gpu_scanner_create");
        BuildAcceptCache();
    }

    Scanner(Scanner&& o) noexcept : Handle(o.Handle), AcceptBegin(std::move(o.AcceptBegin)), AcceptIds(std::move(o.AcceptIds))
    {
        o.Handle = nullptr;
        o.AcceptBegin.assign(1, 0);
    }
    Scanner& operator=(Scanner&& o) noexcept
    {
        if (this != &o) {
            pire_gpu_scanner_destroy(Handle);
            Handle = o.Handle;
            AcceptBegin = std::move(o.AcceptBegin);
            AcceptIds = std::move(o.AcceptIds);
            o.Handle = nullptr;
            o.AcceptBegin.assign(1, 0);
        }
        return *this;
    }
    Scanner(const Scanner&) = delete;
    Scanner& operator=(const Scanner&) = delete;
    ~Scanner() { pire_gpu_scanner_destroy(Handle); }

    size_t Size() const { return Info().states; }                       // multi.h:134
    bool Empty() const { return Info().empty != 0; }                    // multi.h:135
    size_t RegexpsCount() const { return Info().regexps; }              // multi.h:139
    size_t LettersCount() const { return Info().letters; }              // multi.h:140

    void Initialize(State& st) const { st = pire_gpu_initial(Handle); }                       // multi.h:161
    Action Next(State& st, Char c) const { st = pire_gpu_next(Handle, st, c); return 0; }     // multi.h:189-192
    void TakeAction(State&, Action) const {}                                                  // multi.h:194
    bool Final(const State& st) const { return pire_gpu_final(Handle, st) != 0; }             // multi.h:143
    bool Dead(const State& st) const { return pire_gpu_dead(Handle, st) != 0; }               // multi.h:147
    size_t StateIndex(State st) const { return st; }                                          // multi.h:281-284

    // multi.h:149-158.  The returned range stays valid for the scanner's lifetime.  The lists are built once in the
    // constructor, so this is a pure read: like a const Pire::Scanner, one object may be shared between threads.
    std::pair<const size_t*, const size_t*> AcceptedRegexps(const State& st) const
    {
        if (st >= AcceptBegin.size() - 1)
            return std::make_pair(AcceptIds.data(), AcceptIds.data());
        return std::make_pair(AcceptIds.data() + AcceptBegin[st], AcceptIds.data() + AcceptBegin[st + 1] - 1);
    }

    // ---- device ----------------------------------------------------------------
    void Tune(const Batch& sample, unsigned flags = PIRE_GPU_RUN_BEGIN | PIRE_GPU_RUN_END, void* stream = nullptr)
    {
        Check(pire_gpu_scanner_tune(Handle, sample.Corpus, sample.Offsets, sample.FixedLen, sample.Count, flags, stream),
              "pire_gpu_scanner_tune");
    }

    pire_gpu_info Info() const
    {
        pire_gpu_info info;
        Check(pire_gpu_scanner_info(Handle, &info), "pire_gpu_scanner_info");
        return info;
    }

    pire_gpu_scanner* Raw() const { return Handle; }

private:
    // every state's accept list, each followed by a terminator like m_final's (multi.h:96)
    void BuildAcceptCache()
    {
        const size_t states = Info().states;
        AcceptBegin.assign(states + 1, 0);
        AcceptIds.clear();
        std::vector<uint32_t> ids(64);
        for (size_t st = 0; st < states; ++st) {
            AcceptBegin[st] = AcceptIds.size();
            size_t k = pire_gpu_accepted_regexps(Handle, (uint32_t) st, ids.data(), ids.size());
            if (k > ids.size()) {
                ids.resize(k);
                k = pire_gpu_accepted_regexps(Handle, (uint32_t) st, ids.data(), ids.size());
            }
            AcceptIds.insert(AcceptIds.end(), ids.begin(), ids.begin() + k);
            AcceptIds.push_back(static_cast<size_t>(-1));
        }
        AcceptBegin[states] = AcceptIds.size();
    }

    pire_gpu_scanner* Handle;
    std::vector<size_t> AcceptBegin;      // [states + 1] into AcceptIds
    std::vector<size_t> AcceptIds;
};

// Counterpart of Pire::RunHelper (run.h:365-386) for a device batch.  Results are
// written to caller-owned device buffers (any may be null); MatchesHost() etc. are
// conveniences that copy them back through the host entry point.
class BatchRunner {
public:
    explicit BatchRunner(const Scanner& sc) : Sc(&sc), Flags(0), Ran(false) {}

    BatchRunner& Begin() { Flags |= PIRE_GPU_RUN_BEGIN; return *this; }      // run.h:375
    BatchRunner& End() { Flags |= PIRE_GPU_RUN_END; return *this; }          // run.h:376
    BatchRunner& Run(const Batch& b) { Input = b; Ran = true; return *this; } // run.h:372

    // Launches the fused Begin/Run/End pass on `stream` (cudaStream_t as void*).
    void Launch(uint32_t* d_match_bits, uint32_t* d_accept_masks, uint32_t* d_state_idx, void* stream = nullptr) const
    {
        if (!Ran)
            throw Error(PIRE_GPU_EINVAL, "BatchRunner::Run() was not called");
        Check(pire_gpu_run_batch(Sc->Raw(), Input.Corpus, Input.Offsets, Input.FixedLen, Input.Count, Flags, d_match_bits,
                                 d_accept_masks, d_state_idx, stream),
              "pire_gpu_run_batch");
    }

private:
    const Scanner* Sc;
    Batch Input;
    unsigned Flags;
    bool Ran;
};

inline BatchRunner Runner(const Scanner& sc) { return BatchRunner(sc); }     // run.h:388-389

// AcceptedRegexps for scanners with more than 32 regexps: rows of AcceptWords(sc) words, bit r of row i set iff
// regexp r is accepted by the state string i stopped in (d_state_idx from BatchRunner::Launch).
inline uint32_t AcceptWords(const Scanner& sc) { return pire_gpu_accept_words(sc.Raw()); }
inline void AcceptSets(const Scanner& sc, const uint32_t* d_state_idx, uint64_t n, uint32_t* d_sets, void* stream = nullptr)
{
    Check(pire_gpu_accept_sets(sc.Raw(), d_state_idx, n, d_sets, stream), "pire_gpu_accept_sets");
}

// One rank of a multi-GPU run (one GPU per process or thread): the communicator of pire_gpu_run_sharded.
//     rank 0:  unsigned char id[PIRE_GPU_COMM_ID_BYTES]; Comm::MakeId(id);  ... ship id to every rank ...
//     all:     Comm comm(id, world, rank, device);
//              comm.RunSharded(gsc, shard, n_global, flags, d_bits_all, d_masks, nullptr, stream);
class Comm {
public:
    static void MakeId(void* id) { Check(pire_gpu_comm_get_id(id), "pire_gpu_comm_get_id"); }
    Comm(const void* id, int world, int rank, int device) : Handle(nullptr)
    {
        Check(pire_gpu_comm_create(id, world, rank, device, &Handle), "pire_gpu_comm_create");
    }
    // around an ncclComm_t the caller owns (not destroyed here)
    Comm(void* ncclComm, int device) : Handle(nullptr) { Check(pire_gpu_comm_adopt(ncclComm, device, &Handle), "pire_gpu_comm_adopt"); }
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    ~Comm() { pire_gpu_comm_destroy(Handle); }

    int World() const { int w = 1; pire_gpu_comm_info(Handle, &w, nullptr); return w; }
    int Rank() const { int r = 0; pire_gpu_comm_info(Handle, nullptr, &r); return r; }
    // [lo, hi) of this rank's shard and the words of the gathered bitmap
    std::pair<uint64_t, uint64_t> Bounds(uint64_t n_global) const
    {
        uint64_t lo = 0, hi = 0;
        pire_gpu_shard_bounds(n_global, World(), Rank(), &lo, &hi);
        return std::make_pair(lo, hi);
    }
    uint64_t Words(uint64_t n_global) const { return pire_gpu_sharded_words(n_global, World()); }

    // `shard` describes this rank's strings only (Count is ignored: the shard is Bounds(n_global)).
    void RunSharded(const Scanner& sc, const Batch& shard, uint64_t n_global, unsigned flags, uint32_t* d_match_bits_all,
                    uint32_t* d_accept_masks = nullptr, uint32_t* d_state_idx = nullptr, void* stream = nullptr)
    {
        Check(pire_gpu_run_sharded(sc.Raw(), Handle, shard.Corpus, shard.Offsets, shard.FixedLen, n_global, flags, d_match_bits_all,
                                   d_accept_masks, d_state_idx, stream), "pire_gpu_run_sharded");
    }
    void Wait(void* stream = nullptr) { Check(pire_gpu_comm_wait(Handle, stream), "pire_gpu_comm_wait"); }
    pire_gpu_comm* Raw() const { return Handle; }

private:
    pire_gpu_comm* Handle;
};

// Batch counterparts of Pire::LongestPrefix / Pire::ShortestPrefix (run.h:277-311): one prefix
// length per string into d_prefix_len (PIRE_GPU_NO_PREFIX where the reference returns null).
inline void LongestPrefix(const Scanner& sc, const Batch& b, uint32_t* d_prefix_len, bool throughBeginMark = false,
                          bool throughEndMark = false, void* stream = nullptr)
{
    unsigned flags = (throughBeginMark ? PIRE_GPU_RUN_BEGIN : 0u) | (throughEndMark ? PIRE_GPU_RUN_END : 0u);
    Check(pire_gpu_prefix_batch(sc.Raw(), b.Corpus, b.Offsets, b.FixedLen, b.Count, flags, 0, d_prefix_len, stream),
          "pire_gpu_prefix_batch");
}

inline void ShortestPrefix(const Scanner& sc, const Batch& b, uint32_t* d_prefix_len, bool throughBeginMark = false,
                           bool throughEndMark = false, void* stream = nullptr)
{
    unsigned flags = (throughBeginMark ? PIRE_GPU_RUN_BEGIN : 0u) | (throughEndMark ? PIRE_GPU_RUN_END : 0u);
    Check(pire_gpu_prefix_batch(sc.Raw(), b.Corpus, b.Offsets, b.FixedLen, b.Count, flags, 1, d_prefix_len, stream),
          "pire_gpu_prefix_batch");
}

// Batch counterparts of Pire::LongestSuffix / Pire::ShortestSuffix (run.h:316-362): every string is walked from
// its last byte to its first; one suffix length per string (PIRE_GPU_NO_PREFIX for null).
inline void LongestSuffix(const Scanner& sc, const Batch& b, uint32_t* d_suffix_len, bool throughEndMark = false,
                          bool throughBeginMark = false, void* stream = nullptr)
{
    unsigned flags = (throughBeginMark ? PIRE_GPU_RUN_BEGIN : 0u) | (throughEndMark ? PIRE_GPU_RUN_END : 0u);
    Check(pire_gpu_suffix_batch(sc.Raw(), b.Corpus, b.Offsets, b.FixedLen, b.Count, flags, 0, d_suffix_len, stream),
          "pire_gpu_suffix_batch");
}

inline void ShortestSuffix(const Scanner& sc, const Batch& b, uint32_t* d_suffix_len, bool throughEndMark = false,
                           bool throughBeginMark = false, void* stream = nullptr)
{
    unsigned flags = (throughBeginMark ? PIRE_GPU_RUN_BEGIN : 0u) | (throughEndMark ? PIRE_GPU_RUN_END : 0u);
    Check(pire_gpu_suffix_batch(sc.Raw(), b.Corpus, b.Offsets, b.FixedLen, b.Count, flags, 1, d_suffix_len, stream),
          "pire_gpu_suffix_batch");
}

// Batch counterpart of running a Pire::HalfFinalScanner over each string the way tests/count_ut.cpp:54-63
// does -- Initialize, [Step(BeginMark)], Run, [Step(EndMark)] -- and reading State::Result(r)
// (pire/scanners/half_final.h:88-90,:136-163) for every regexp: d_counts holds Count rows of
// max(1, RegexpsCount()) u32.  `sc` is built from the HalfFinalScanner (its Save() stream is Scanner's).
inline void HalfFinalCount(const Scanner& sc, const Batch& b, uint32_t* d_counts, uint32_t* d_match_bits = nullptr,
                           unsigned flags = PIRE_GPU_RUN_BEGIN | PIRE_GPU_RUN_END, void* stream = nullptr)
{
    Check(pire_gpu_count_batch(sc.Raw(), b.Corpus, b.Offsets, b.FixedLen, b.Count, flags, d_counts, d_match_bits, stream),
          "pire_gpu_count_batch");
}

// Host-buffer counterpart of `bool Pire::Runner(sc).Begin().Run(p, n).End()` for many
// strings at once (CSR): fills `matched[i]`.
inline void MatchesHost(const Scanner& sc, const uint8_t* corpus, const uint64_t* offsets, uint64_t n,
                        std::vector<bool>& matched, unsigned flags = PIRE_GPU_RUN_BEGIN | PIRE_GPU_RUN_END)
{
    std::vector<uint32_t> bits((n + 31) / 32);
    Check(pire_gpu_run_batch_host(sc.Raw(), corpus, n ? offsets[n] : 0, offsets, 0, n, flags, bits.data(), nullptr, nullptr),
          "pire_gpu_run_batch_host");
    matched.resize(n);
    for (uint64_t i = 0; i < n; ++i)
        matched[i] = (bits[i / 32] >> (i % 32)) & 1u;
}

} // namespace Gpu
} // namespace Pire
