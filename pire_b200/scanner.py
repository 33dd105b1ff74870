"""Python mirror of the reference's Scanner / Runner / Matches surface for the
batch scan path, over the C ABI (include/pire_b200.h).

Reference interface mirrored (same names and argument meaning):
    Pire::Scanner             Size / Empty / RegexpsCount / LettersCount /
                              Initialize / Next / Final / Dead / AcceptedRegexps /
                              StateIndex / Load          pire/scanners/multi.h:134-194,:244-311
    Pire::Runner(sc)          .Begin().Run(...).End()    pire/run.h:365-392
    Pire::Matches(sc, ...)    no Begin/End marks         pire/run.h:396-400

The one difference is the unit of work: Run() takes a *batch* of strings
resident in HBM (``Batch``) instead of one ``const char*`` range, and the
RunHelper answers per string.  torch is used only to own device memory and to
name the CUDA stream.
"""
import ctypes as C

import numpy as np

from . import _native as N

BeginMark = 258   # pire/defs.h:63
EndMark = 259     # pire/defs.h:64


def _torch():
    import torch
    return torch


class Batch:
    """A batch of strings on one GPU: ``corpus`` (uint8 CUDA tensor) plus either CSR
    ``offsets`` (int64/uint64 CUDA tensor, n+1 entries) or a fixed string length."""

    def __init__(self, corpus, offsets=None, fixed_len=0, n=None):
        torch = _torch()
        if corpus.dtype != torch.uint8 or not corpus.is_cuda or not corpus.is_contiguous():
            raise ValueError("corpus must be a contiguous uint8 CUDA tensor")
        self.corpus = corpus
        self.offsets = offsets
        self.fixed_len = int(fixed_len)
        if offsets is not None:
            if offsets.dtype not in (torch.int64, torch.uint64) or not offsets.is_cuda or not offsets.is_contiguous():
                raise ValueError("offsets must be a contiguous int64 CUDA tensor")
            self.n = offsets.numel() - 1 if n is None else int(n)
        else:
            if n is None:
                n = corpus.numel() // self.fixed_len if self.fixed_len else 0
            self.n = int(n)
            if self.n * self.fixed_len > corpus.numel():
                raise ValueError("corpus shorter than n * fixed_len")
        self.device = corpus.device
        self.order = None          # set by bin_by_length()
        self.trim = 0              # 1 for line batches (from_text): the newline ending a line is not part of it

    def bin_by_length(self):
        """Sort the strings by descending length (on the device) so that the lanes of a warp
        scan strings of similar length; results stay indexed by the original string number."""
        torch = _torch()
        if self.offsets is None or self.n == 0:
            return self
        order = torch.empty(self.n, dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(N.lib.pire_gpu_length_order(self.offsets.data_ptr(), self.n, order.data_ptr(), self.device.index or 0, stream),
                "pire_gpu_length_order")
        self.order = order
        return self

    @classmethod
    def from_text(cls, text):
        """Lines of a newline-delimited text (uint8 CUDA tensor), found on the device with
        std::getline semantics -- what samples/pigrep/pigrep.cpp:38-45 feeds to Runner per line."""
        torch = _torch()
        if text.dtype != torch.uint8 or not text.is_cuda or not text.is_contiguous():
            raise ValueError("text must be a contiguous uint8 CUDA tensor")
        dev = text.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        n_lines = C.c_uint64(0)
        cap = max(1024, text.numel() // 64)
        while True:
            offsets = torch.empty(cap + 1, dtype=torch.int64, device=dev)
            rc = N.lib.pire_gpu_split_lines(text.data_ptr(), text.numel(), offsets.data_ptr(), cap, C.byref(n_lines),
                                            dev.index or 0, stream)
            if rc == 0:
                break
            if rc == -1 and n_lines.value > cap:         # buffer too small: retry with the reported capacity
                cap = int(n_lines.value)
                continue
            N.check(rc, "pire_gpu_split_lines")
        b = cls(text, offsets[: n_lines.value + 1].contiguous(), n=int(n_lines.value))
        b.trim = 1
        return b

    @classmethod
    def from_strings(cls, strings, device="cuda:0"):
        """Host convenience: pack Python byte strings (CSR) and upload."""
        torch = _torch()
        offs = np.zeros(len(strings) + 1, np.int64)
        np.cumsum([len(s) for s in strings], out=offs[1:])
        blob = np.frombuffer(b"".join(strings) + b"\0" * 32, dtype=np.uint8).copy()
        return cls(torch.from_numpy(blob).to(device), torch.from_numpy(offs).to(device), n=len(strings))

    def payload_bytes(self):
        if self.offsets is None:
            return self.n * self.fixed_len
        o = self.offsets
        return int(o[self.n].item() - o[0].item()) - self.trim * self.n


class Scanner:
    """A compiled multi-regexp scanner resident on one B200 (Pire::Scanner's role)."""

    def __init__(self, image, device=0):
        image = bytes(image)
        h = C.c_void_p()
        buf = (C.c_char * len(image)).from_buffer_copy(image)
        N.check(N.lib.pire_gpu_scanner_create(buf, len(image), int(device), C.byref(h)), "pire_gpu_scanner_create")
        self._h = h
        self.device = int(device)

    # Scanner::Load(yistream*) -- multi.h:575-599
    @classmethod
    def Load(cls, image, device=0):
        return cls(image, device)

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(N, "lib", None)          # None while the interpreter is shutting down
        if h and lib is not None:
            lib.pire_gpu_scanner_destroy(h)
            self._h = None

    def info(self):
        out = N.Info()
        N.check(N.lib.pire_gpu_scanner_info(self._h, C.byref(out)), "pire_gpu_scanner_info")
        return out

    def Size(self):
        return self.info().states

    def Empty(self):
        return bool(self.info().empty)

    def set_count_mode(self, mode):
        """0 auto, 1 accept lists, 2 packed increments, 3 packed on every chunk (pire_gpu_scanner_set_count_mode)."""
        N.check(N.lib.pire_gpu_scanner_set_count_mode(self._h, mode), "pire_gpu_scanner_set_count_mode")

    def RegexpsCount(self):
        return self.info().regexps

    def LettersCount(self):
        return self.info().letters

    # --- host-side Scanner concept (index space) ------------------------------
    def Initialize(self):
        return N.lib.pire_gpu_initial(self._h)

    def Next(self, state, ch):
        return N.lib.pire_gpu_next(self._h, state, ch)

    def Final(self, state):
        return bool(N.lib.pire_gpu_final(self._h, state))

    def Dead(self, state):
        return bool(N.lib.pire_gpu_dead(self._h, state))

    def AcceptedRegexps(self, state):
        ids = (C.c_uint32 * 1024)()
        k = N.lib.pire_gpu_accepted_regexps(self._h, state, ids, 1024)
        return [int(ids[i]) for i in range(min(k, 1024))]

    @staticmethod
    def StateIndex(state):
        return state

    # --- device ------------------------------------------------------------------
    def set_variant(self, variant):
        N.check(N.lib.pire_gpu_scanner_set_variant(self._h, variant), "pire_gpu_scanner_set_variant")

    def set_max_hot(self, rows):
        N.check(N.lib.pire_gpu_scanner_set_max_hot(self._h, rows), "pire_gpu_scanner_set_max_hot")

    def Tune(self, batch, n_sample=None, begin=True, end=True):
        """Pick the shared-memory rows from the states a sample of `batch` visits."""
        n_sample = batch.n if n_sample is None else min(int(n_sample), batch.n)
        flags = (N.RUN_BEGIN if begin else 0) | (N.RUN_END if end else 0) | (N.RUN_LINES if batch.trim else 0)
        torch = _torch()
        stream = torch.cuda.current_stream(batch.device).cuda_stream
        N.check(N.lib.pire_gpu_scanner_tune(self._h, batch.corpus.data_ptr(),
                                            batch.offsets.data_ptr() if batch.offsets is not None else None,
                                            batch.fixed_len, n_sample, flags, stream), "pire_gpu_scanner_tune")

    def AutoSelect(self, batch, begin=True, end=True):
        """Time the kernel variants on `batch` and keep the fastest as the AUTO choice.
        Returns {variant name: ms}."""
        flags = (N.RUN_BEGIN if begin else 0) | (N.RUN_END if end else 0) | (N.RUN_LINES if batch.trim else 0)
        torch = _torch()
        stream = torch.cuda.current_stream(batch.device).cuda_stream
        ms = (C.c_float * N.VARIANT_SLOTS)()
        N.check(N.lib.pire_gpu_scanner_autoselect(self._h, batch.corpus.data_ptr(),
                                                  batch.offsets.data_ptr() if batch.offsets is not None else None,
                                                  batch.fixed_len, batch.n, flags, stream, ms),
                "pire_gpu_scanner_autoselect")
        return {name: float(ms[v]) for v, name in N.VARIANT_NAMES.items() if ms[v] > 0}

    def run_batch(self, batch, flags, match_bits=None, accept_masks=None, state_idx=None, stream=None):
        """Thin wrapper of pire_gpu_run_batch: asynchronous on the current stream."""
        torch = _torch()
        if stream is None:
            stream = torch.cuda.current_stream(batch.device).cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()
        if getattr(batch, "trim", 0):
            order = batch.order.data_ptr() if batch.order is not None else None
            N.check(N.lib.pire_gpu_run_lines(self._h, batch.corpus.data_ptr(), ptr(batch.offsets), order, batch.n, flags,
                                             ptr(match_bits), ptr(accept_masks), ptr(state_idx), stream),
                    "pire_gpu_run_lines")
            return
        if getattr(batch, "order", None) is not None:
            N.check(N.lib.pire_gpu_run_batch_ordered(self._h, batch.corpus.data_ptr(), ptr(batch.offsets),
                                                     batch.order.data_ptr(), batch.n, flags, ptr(match_bits),
                                                     ptr(accept_masks), ptr(state_idx), stream),
                    "pire_gpu_run_batch_ordered")
            return
        N.check(N.lib.pire_gpu_run_batch(self._h, batch.corpus.data_ptr(), ptr(batch.offsets), batch.fixed_len,
                                         batch.n, flags, ptr(match_bits), ptr(accept_masks), ptr(state_idx), stream),
                "pire_gpu_run_batch")

    def run_batch_host(self, corpus, offsets=None, fixed_len=0, n=None, flags=N.RUN_BEGIN | N.RUN_END,
                       want_masks=False, want_states=False):
        """Host buffers in, host results out (pire_gpu_run_batch_host): numpy arrays
        (or pinned torch CPU tensors viewed as numpy)."""
        corpus = np.ascontiguousarray(corpus, dtype=np.uint8)
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets).view(np.uint64)
            n = len(offsets) - 1 if n is None else n
        elif n is None:
            n = len(corpus) // fixed_len if fixed_len else 0
        bits = np.zeros((n + 31) // 32, np.uint32)
        masks = np.zeros(n, np.uint32) if want_masks else None
        states = np.zeros(n, np.uint32) if want_states else None
        p = lambda a: None if a is None else a.ctypes.data
        N.check(N.lib.pire_gpu_run_batch_host(self._h, p(corpus), corpus.nbytes, p(offsets), fixed_len, n, flags,
                                              p(bits), p(masks), p(states)), "pire_gpu_run_batch_host")
        return bits, masks, states


class RunHelper:
    """Pire::RunHelper (run.h:365-386) over a batch.  ``Begin()`` / ``End()`` record the
    mark steps, ``Run(batch)`` names the strings; the single fused launch happens when
    a result is first asked for."""

    def __init__(self, sc):
        self.Sc = sc
        self._begin = False
        self._end = False
        self._batch = None
        self._bits = self._masks = self._states = None

    def Begin(self):
        if self._batch is not None:
            raise ValueError("Begin() must precede Run()")
        self._begin = True
        return self

    def Run(self, batch):
        if self._batch is not None:
            raise ValueError("one Run() per RunHelper on the batch path")
        self._batch = batch
        return self

    def End(self):
        self._end = True
        return self

    def _launch(self):
        if self._bits is not None:
            return
        if self._batch is None:
            raise ValueError("Run() was not called")
        torch = _torch()
        b = self._batch
        dev = b.device
        self._bits = torch.empty((b.n + 31) // 32, dtype=torch.int32, device=dev)
        self._masks = torch.empty(b.n, dtype=torch.int32, device=dev)
        self._states = torch.empty(b.n, dtype=torch.int32, device=dev)
        flags = (N.RUN_BEGIN if self._begin else 0) | (N.RUN_END if self._end else 0)
        self.Sc.run_batch(b, flags, self._bits, self._masks, self._states)

    # per-string results -------------------------------------------------------------
    def MatchBits(self):
        """Packed device bitmap: bit i%32 of word i/32 = Final() of string i."""
        self._launch()
        return self._bits

    def Matches(self):
        """numpy bool[n]: RunHelper::operator bool (run.h:380-381) per string."""
        self._launch()
        words = self._bits.cpu().numpy().view(np.uint32)
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")
        return bits[: self._batch.n].astype(bool)

    def AcceptMasks(self):
        self._launch()
        return self._masks.cpu().numpy().view(np.uint32)

    def States(self):
        """StateIndex() of each string's last state (reference numbering)."""
        self._launch()
        return self._states.cpu().numpy().view(np.uint32)

    def AcceptedRegexps(self, i):
        """Scanner::AcceptedRegexps(State()) for string i (multi.h:149-158)."""
        return self.Sc.AcceptedRegexps(int(self.States()[i]))


def Runner(sc):
    """Pire::Runner(sc) (run.h:388-389)."""
    return RunHelper(sc)


def Matches(sc, batch):
    """Pire::Matches(scanner, begin, end) (run.h:396-400): Run without Begin/End marks."""
    return Runner(sc).Run(batch).Matches()


def _prefix(sc, batch, shortest, throughBeginMark, throughEndMark, suffix=False):
    torch = _torch()
    out = torch.empty(batch.n, dtype=torch.int32, device=batch.device)
    flags = (N.RUN_BEGIN if throughBeginMark else 0) | (N.RUN_END if throughEndMark else 0) | (N.RUN_LINES if batch.trim else 0)
    stream = torch.cuda.current_stream(batch.device).cuda_stream
    fn = N.lib.pire_gpu_suffix_batch if suffix else N.lib.pire_gpu_prefix_batch
    N.check(fn(sc._h, batch.corpus.data_ptr(), batch.offsets.data_ptr() if batch.offsets is not None else None,
               batch.fixed_len, batch.n, flags, int(shortest), out.data_ptr(), stream),
            "pire_gpu_suffix_batch" if suffix else "pire_gpu_prefix_batch")
    res = out.cpu().numpy().view(np.uint32).astype(np.int64)
    res[res == 0xFFFFFFFF] = -1
    return res


def LongestPrefix(sc, batch, throughBeginMark=False, throughEndMark=False):
    """Pire::LongestPrefix (run.h:277-292) per string: prefix length, or -1 where the reference returns null."""
    return _prefix(sc, batch, False, throughBeginMark, throughEndMark)


def ShortestPrefix(sc, batch, throughBeginMark=False, throughEndMark=False):
    """Pire::ShortestPrefix (run.h:294-311) per string: prefix length, or -1 where the reference returns null."""
    return _prefix(sc, batch, True, throughBeginMark, throughEndMark)


def LongestSuffix(sc, batch, throughEndMark=False, throughBeginMark=False):
    """Pire::LongestSuffix (run.h:316-342) per string, walked from its last byte: suffix length, or -1 for null."""
    return _prefix(sc, batch, False, throughBeginMark, throughEndMark, suffix=True)


def ShortestSuffix(sc, batch, throughEndMark=False, throughBeginMark=False):
    """Pire::ShortestSuffix (run.h:345-362) per string: suffix length, or -1 for null."""
    return _prefix(sc, batch, True, throughBeginMark, throughEndMark, suffix=True)


class HalfFinalResult:
    """What a batch of HalfFinalScanner states reports (half_final.h:58-120)."""

    def __init__(self, counts, final):
        self.counts, self.final = counts, final

    def Result(self, i, regexp_id):
        """State::Result(regexp_id) of string i (half_final.h:88-90)."""
        return int(self.counts[i, regexp_id])

    def AcceptedRegexps(self, i):
        """HalfFinalScanner::AcceptedRegexps (half_final.h:130-133): regexps with a non-zero counter."""
        return [int(r) for r in np.nonzero(self.counts[i])[0]]

    def Final(self, i):
        return bool(self.final[i])


def HalfFinalCount(sc, batch, begin=True, end=True):
    """Per string: Initialize; [Step(BeginMark)]; Run; [Step(EndMark)] of a Pire::HalfFinalScanner
    (half_final.h:136-163, driven as tests/count_ut.cpp:54-63 does) -> HalfFinalResult with
    counts[n, regexps] (numpy u32) and final[n] (bool).  `sc` is a Scanner loaded from the
    HalfFinalScanner's Save() stream."""
    torch = _torch()
    regs = max(1, sc.RegexpsCount())
    counts = torch.empty((batch.n, regs), dtype=torch.int32, device=batch.device)
    bits = torch.zeros((batch.n + 31) // 32, dtype=torch.int32, device=batch.device)
    flags = (N.RUN_BEGIN if begin else 0) | (N.RUN_END if end else 0) | (N.RUN_LINES if batch.trim else 0)
    stream = torch.cuda.current_stream(batch.device).cuda_stream
    N.check(N.lib.pire_gpu_count_batch(sc._h, batch.corpus.data_ptr(),
                                       batch.offsets.data_ptr() if batch.offsets is not None else None,
                                       batch.fixed_len, batch.n, flags, counts.data_ptr(), bits.data_ptr(), stream),
            "pire_gpu_count_batch")
    words = bits.cpu().numpy().view(np.uint32)
    final = ((words[np.arange(batch.n) // 32] >> (np.arange(batch.n) % 32).astype(np.uint32)) & 1).astype(bool)
    return HalfFinalResult(counts.cpu().numpy().view(np.uint32), final)
