"""Content vs structure: the mixed UTF-8 bytes as fixed-length strings through both kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pire_b200 as P
from pire_b200 import _native as N
from pire_b200 import workloads as W

dev = torch.device("cuda:0")
sc = P.Scanner(W.load_image("headline_iu"), 0)
flags = N.RUN_BEGIN | N.RUN_END


def timeit(batch, label, payload, variant):
    sc.set_variant(variant)
    bits = torch.zeros((batch.n + 31) // 32, dtype=torch.int32, device=dev)
    masks = torch.empty(batch.n, dtype=torch.int32, device=dev)
    for _ in range(2):
        sc.run_batch(batch, flags, bits, masks, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sc.run_batch(batch, flags, bits, masks, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-58s n=%8d  %8.3f ms  %8.1f GB/s" % (label, batch.n, ms, payload / ms / 1e6), flush=True)


spec = W.MixedSpec(400000)
mc, mo = spec.device_batch(dev)
total = int(mo[-1].item()) // 65536 * 65536
tune_batch = P.Batch(mc, fixed_len=1024, n=16384)
sc.Tune(tune_batch, 16384)
for length in (1024, 8192, 65536):
    n = total // length
    for vname, v in (("plain", N.VARIANT_PLAIN), ("pred", N.VARIANT_PRED)):
        timeit(P.Batch(mc, fixed_len=length, n=n), "UTF-8 mix, uniform kernel %s, len %d" % (vname, length), n * length, v)
    offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * length
    timeit(P.Batch(mc, offs, n=n), "UTF-8 mix, generic kernel plain, len %d" % length, n * length, N.VARIANT_PLAIN)
    perm = torch.randperm(n, device=dev)
    # same strings, lanes pointed at scattered strings (what length binning does)
    b = P.Batch(mc, offs, n=n)
    b.order = perm.to(torch.int32)
    timeit(b, "UTF-8 mix, generic, random order, len %d" % length, n * length, N.VARIANT_PLAIN)
mb = P.Batch(mc, mo, n=400000)
timeit(mb, "mixed lengths, unordered", mb.payload_bytes(), N.VARIANT_PLAIN)
mb.bin_by_length()
timeit(mb, "mixed lengths, binned", mb.payload_bytes(), N.VARIANT_PLAIN)
