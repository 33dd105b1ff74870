"""The BASELINE.json workloads: pattern sets and synthetic corpora.

The reference names no canonical 10-pattern set; this one starts from the four
RE2-paper patterns of tools/bench/run-bench:62-65,:72-75 and adds the headline
pattern of BASELINE.json plus five log-style patterns (SURVEY.md App. D set B),
glued left to right like tools/bench/bench.cpp:108-132.
"""
import ctypes as C

import numpy as np

from . import _native as N

HEADLINE = (rb"hello\s+w.+d$", "")

GLUE10 = [
    (rb"ABCDEFGHIJKLMNOPQRSTUVWXYZ$", ""),
    (rb"[XYZ]ABCDEFGHIJKLMNOPQRSTUVWXYZ$", ""),
    (rb"[ -~]*ABCDEFGHIJKLMNOPQRSTUVWXYZ$", ""),
    (rb"(\d{3}-|\(\d{3}\)\s+)(\d{3}-\d{4})$", ""),
    HEADLINE,
    (rb"error", ""),
    (rb"fatal", ""),
    (rb"https?://", ""),
    (rb"^GET ", ""),
    (rb"timeout$", ""),
]

# One literal per regexp id; a leading '^' / '$' pins it to the start / end of the
# string (the anchored patterns), otherwise it lands at a pseudo-random offset.
GLUE10_PLANTS = [
    b"$ABCDEFGHIJKLMNOPQRSTUVWXYZ",
    b"$XABCDEFGHIJKLMNOPQRSTUVWXYZ",
    b"$ABCDEFGHIJKLMNOPQRSTUVWXYZ",
    b"$(555) 123-4567",
    b"$hello \t world",
    b"error",
    b"fatal",
    b"https://",
    b"^GET ",
    b"$timeout",
]

HEADLINE_PLANTS = [b"$hello \t world"]


class SynthSpec:
    """Deterministic printable-ASCII corpus (pire_b200/csrc/synth.h)."""

    def __init__(self, n_strings, string_len=1024, seed=42, plant_every=8, plants=(), first_string=0):
        self.n_strings, self.string_len, self.seed = int(n_strings), int(string_len), int(seed)
        self.plant_every, self.plants, self.first_string = int(plant_every), list(plants), int(first_string)

    def _c(self, first_string=None, n=None):
        blob = b"\0".join(self.plants) + b"\0"
        s = N.Synth()
        s.seed = self.seed
        s.first_string = self.first_string if first_string is None else first_string
        s.n_strings = self.n_strings if n is None else n
        s.string_len = self.string_len
        s.kind = 0
        s.plant_every = self.plant_every if self.plants else 0
        s.n_plants = len(self.plants)
        s.plants = blob
        s.plants_bytes = len(blob)
        s.tail = 0
        s._keep = blob
        return s

    def total_bytes(self):
        return self.n_strings * self.string_len

    def fill_device(self, tensor, stream=None):
        """Write the whole corpus into a uint8 CUDA tensor of total_bytes()."""
        import torch
        assert tensor.is_cuda and tensor.dtype == torch.uint8 and tensor.numel() >= self.total_bytes()
        if stream is None:
            stream = torch.cuda.current_stream(tensor.device).cuda_stream
        s = self._c()
        N.check(N.lib.pire_gpu_synth_fill_device(C.byref(s), tensor.data_ptr(), tensor.device.index or 0, stream),
                "pire_gpu_synth_fill_device")
        return tensor

    def host_sample(self, first, count, threads=0):
        """Strings [first, first+count) of the same corpus, generated on the host
        (sliced over a few threads: the generator is pure and the C call drops the GIL)."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        out = np.empty(count * self.string_len, np.uint8)
        s = self._c()
        if threads <= 0:
            threads = min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)
        threads = max(1, min(threads, count // 4096 or 1))

        def fill(k):
            lo = count * k // threads
            hi = count * (k + 1) // threads
            N.check(N.lib.pire_gpu_synth_fill_host(C.byref(s), out.ctypes.data + lo * self.string_len, first + lo, hi - lo),
                    "pire_gpu_synth_fill_host")
        if threads == 1:
            fill(0)
        else:
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(fill, range(threads)))
        return out

    def shard(self, rank, world):
        """Contiguous string-index shard, boundaries aligned to 32 strings so bitmap words never straddle ranks."""
        per = (self.n_strings + world - 1) // world
        per = (per + 31) // 32 * 32
        lo = rank * per                                  # stays on the 32-string grid even when the shard is empty
        hi = max(lo, min(self.n_strings, lo + per))
        return SynthSpec(hi - lo, self.string_len, self.seed, self.plant_every, self.plants, self.first_string + lo), lo


class MixedSpec:
    """BASELINE config 4: mixed-length (16 B .. 64 KiB) UTF-8 strings, CSR offsets."""

    def __init__(self, n_strings, seed=42, plant_every=8, first_string=0):
        self.n_strings, self.seed, self.plant_every, self.first_string = int(n_strings), int(seed), int(plant_every), int(first_string)

    def device_batch(self, device):
        """(corpus uint8, offsets int64) CUDA tensors, generated on the device."""
        import torch
        dev = torch.device(device)
        idx = dev.index or 0
        stream = torch.cuda.current_stream(dev).cuda_stream
        lengths = torch.empty(self.n_strings, dtype=torch.int64, device=dev)
        N.check(N.lib.pire_gpu_synth_mixed_lengths_device(self.seed, self.first_string, self.n_strings, lengths.data_ptr(), idx, stream),
                "pire_gpu_synth_mixed_lengths_device")
        offsets = torch.zeros(self.n_strings + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lengths, 0, out=offsets[1:])
        total = int(offsets[-1].item())
        corpus = torch.empty(total + 32, dtype=torch.uint8, device=dev)
        N.check(N.lib.pire_gpu_synth_mixed_fill_device(self.seed, self.plant_every, self.first_string, self.n_strings,
                                                       offsets.data_ptr(), corpus.data_ptr(), idx, stream),
                "pire_gpu_synth_mixed_fill_device")
        return corpus, offsets

    def host_batch(self, first, count):
        """Strings [first, first+count) on the host: (corpus uint8, offsets uint64[count+1])."""
        lengths = np.zeros(count, np.uint64)
        N.check(N.lib.pire_gpu_synth_mixed_lengths_host(self.seed, self.first_string + first, count, lengths.ctypes.data),
                "pire_gpu_synth_mixed_lengths_host")
        offsets = np.zeros(count + 1, np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        corpus = np.zeros(int(offsets[-1]) + 32, np.uint8)
        N.check(N.lib.pire_gpu_synth_mixed_fill_host(self.seed, self.plant_every, self.first_string + first, count,
                                                     offsets.ctypes.data, corpus.ctypes.data),
                "pire_gpu_synth_mixed_fill_host")
        return corpus, offsets


HEADLINE_IU = (rb"hello\s+w.+d$", "iu")      # README:35-47: CaseInsensitive + UTF-8


def load_image(name):
    """Scanner::Save() image of a BASELINE pattern set ('headline' or 'glue10'),
    precompiled by tools/compile_patterns.py with the reference's front end."""
    import lzma
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", name + ".pire.xz")
    with open(path, "rb") as f:
        return lzma.decompress(f.read())


WORKLOADS = {
    # name: (image, plants, description, BASELINE.json config index)
    "headline": ("headline", HEADLINE_PLANTS, r"single regex hello\s+w.+d$ (NonrelocScanner DFA: 11 states)", 1),
    "glue10": ("glue10", GLUE10_PLANTS, "10 regexes glued into one multi-Scanner (29664 states x 54 letters)", 2),
    "utf8mixed": ("headline_iu", None, r"hello\s+w.+d$ with UTF-8 + CaseInsensitive, mixed-length (16 B-64 KiB) UTF-8 strings", 3),
}
