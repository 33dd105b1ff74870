"""Where does the generic (CSR) kernel lose time?  Same bytes, different shapes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pire_b200 as P
from pire_b200 import _native as N
from pire_b200 import workloads as W

dev = torch.device("cuda:0")
sc = P.Scanner(W.load_image("headline_iu"), 0)
sc.set_variant(N.VARIANT_PLAIN)
flags = N.RUN_BEGIN | N.RUN_END


def timeit(batch, label, payload):
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
    print("%-46s n=%8d  %8.3f ms  %8.1f GB/s" % (label, batch.n, ms, payload / ms / 1e6), flush=True)


total = 2 << 30
corpus = torch.randint(0x20, 0x7F, (total + 64,), dtype=torch.uint8, device=dev)
for length in (1024, 8192, 65536):
    n = total // length
    offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * length
    timeit(P.Batch(corpus, fixed_len=length, n=n), "uniform kernel, len %d" % length, n * length)
    b = P.Batch(corpus, offs, n=n)
    timeit(b, "generic kernel (CSR), len %d" % length, n * length)
    b2 = P.Batch(corpus, offs, n=n).bin_by_length()
    timeit(b2, "generic ordered, len %d" % length, n * length)
    offs_odd = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * (length - 4)
    timeit(P.Batch(corpus, offs_odd, n=n), "generic kernel, len %d (4-byte aligned)" % (length - 4), n * (length - 4))

# the real mixed batch, by octave
spec = W.MixedSpec(320000)
mc, mo = spec.device_batch(dev)
mb = P.Batch(mc, mo, n=320000)
timeit(mb, "mixed, unordered", mb.payload_bytes())
mb.bin_by_length()
timeit(mb, "mixed, ordered", mb.payload_bytes())
lens = (mo[1:] - mo[:-1])
for lo, hi in ((16, 1024), (1024, 8192), (8192, 32768), (32768, 65536)):
    sel = ((lens >= lo) & (lens < hi)).nonzero().flatten()
    # build a sub-batch by gathering strings into a fresh corpus
    sl = lens[sel]
    so = torch.zeros(sel.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(sl, 0, out=so[1:])
    sub = torch.empty(int(so[-1].item()) + 64, dtype=torch.uint8, device=dev)
    # gather with a python loop on the host is too slow; use index arithmetic
    idx = torch.repeat_interleave(mo[:-1][sel] - so[:-1], sl) + torch.arange(int(so[-1].item()), device=dev)
    sub[: idx.numel()] = mc[idx]
    sb = P.Batch(sub, so, n=sel.numel()).bin_by_length()
    timeit(sb, "mixed octaves [%d,%d) ordered" % (lo, hi), int(so[-1].item()))
