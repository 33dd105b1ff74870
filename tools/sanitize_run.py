"""A small pass through every kernel of the library, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Checks results against the oracle on the way (bit-exact)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import pire_b200 as P
from pire_b200 import _native as N
from pire_b200 import workloads as W
from refpire import Oracle

dev = "cuda:0"


def check(sc, orc, batch, corpus_host, offsets_host, fixed_len, n, tag):
    f, m, s = orc.run(corpus_host, offsets_host, fixed_len=fixed_len, n=n, shortcuts=True)
    for variant in (N.VARIANT_PLAIN, N.VARIANT_PRED, N.VARIANT_PRIV, N.VARIANT_LOOK, N.VARIANT_LOOK64, N.VARIANT_LOOK1):
        sc.set_variant(variant)
        r = P.Runner(sc).Begin().Run(batch).End()
        assert (r.Matches().astype(np.uint8) == f).all() and (r.AcceptMasks() == m).all() and (r.States() == s).all(), (tag, variant)
    print("ok", tag, "n=%d matches=%d" % (n, int(f.sum())), flush=True)


# glued 10-pattern scanner, fixed-length strings: uniform kernels incl. PRIV, tune, tiny hot sets (replay/cold paths)
img = W.load_image("glue10")
orc = Oracle(img)
sc = P.Scanner(img, 0)
n = 2048 + 3
spec = W.SynthSpec(n, 256, plants=W.GLUE10_PLANTS)
d = torch.empty(spec.total_bytes(), dtype=torch.uint8, device=dev)
spec.fill_device(d)
host = spec.host_sample(0, n)
batch = P.Batch(d, fixed_len=256, n=n)
check(sc, orc, batch, host, None, 256, n, "glue10 static")
sc.Tune(batch, 512)
check(sc, orc, batch, host, None, 256, n, "glue10 tuned")
sc.set_max_hot(3)
check(sc, orc, batch, host, None, 256, n, "glue10 hot=3")
print(sc.AutoSelect(batch))

# UTF-8 scanner, ragged CSR batch: generic kernels, ordered launch, host entry point
img = W.load_image("headline_iu")
orc = Oracle(img)
sc = P.Scanner(img, 0)
mspec = W.MixedSpec(1500)
mc, mo = mspec.device_batch(dev)
hc, ho = mspec.host_batch(0, 1500)
b = P.Batch(mc, mo, n=1500)
check(sc, orc, b, hc, ho, 0, 1500, "mixed unordered")
b.bin_by_length()
check(sc, orc, b, hc, ho, 0, 1500, "mixed binned")
bits, masks, states = sc.run_batch_host(hc, offsets=ho, want_masks=True, want_states=True)
f, m, s = orc.run(hc, ho)
assert (np.unpackbits(bits.view(np.uint8), bitorder="little")[:1500] == f).all() and (states == s).all()
print("ok host entry")
# odd shapes: empty strings, unaligned starts
strings = [b"", b"x", b"hello world", b"\xd0\xbf" * 7 + b"HeLLo  World", b"a" * 33, b""] * 20
bb = P.Batch.from_strings(strings)
r = P.Runner(sc).Begin().Run(bb).End()
from refpire import csr
c, o = csr(strings)
f, m, s = orc.run(c, o)
assert (r.Matches().astype(np.uint8) == f).all() and (r.States() == s).all()
print("ok ragged")
# tiny automaton (one private quad) over bytes >= 128, fixed length: the PRIV kernel's speculative steps
from conftest import GOLDEN
case = next(c for c in GOLDEN if c.name == "UTF8@181b")
sc = P.Scanner(case.image, 0)
orc = Oracle(case.image)
rng = np.random.default_rng(5)
fx = rng.choice(np.frombuffer("x\u0424y ab".encode() + bytes(range(120, 256)), np.uint8), size=(1024, 64)).reshape(-1)
fb = P.Batch(torch.from_numpy(fx).to(dev), fixed_len=64, n=1024)
check(sc, orc, fb, fx, None, 64, 1024, "tiny DFA, high bytes, PRIV")
# lines of text: split on the device, lines kernel (plain / pred) and the generic kernel on the same lines;
# the text starts and ends off a 16-byte boundary so that the clipped chunk loads are exercised
rng = np.random.default_rng(9)
words = [b"hello", b"world", b"GET /x", b"error", b"timeout", b"", b"fatal https://a", b"x" * 70, b"hello   wd"]
text_lines = [b" ".join(words[int(j)] for j in rng.integers(0, len(words), size=int(k))) for k in rng.integers(0, 9, size=3000)]
blob = b"\n".join(text_lines) + b"\n"
pad = torch.zeros(len(blob) + 64, dtype=torch.uint8, device=dev)
text_dev = pad[5:5 + len(blob)]
text_dev.copy_(torch.from_numpy(np.frombuffer(blob, np.uint8).copy()))
lb = P.Batch.from_text(text_dev)
assert lb.n == len(text_lines)
c, o = csr(text_lines)
for name in ("headline", "glue10"):
    image = W.load_image(name)
    orc = Oracle(image)
    f, m, s_ = orc.run(c, o)
    for max_hot in (255, 2):
        sc = P.Scanner(image, 0)
        sc.set_max_hot(max_hot)
        for variant in (N.VARIANT_PLAIN, N.VARIANT_PRED, N.VARIANT_LOOK):
            sc.set_variant(variant)
            r = P.Runner(sc).Begin().Run(lb).End()
            assert (r.Matches().astype(np.uint8) == f).all() and (r.AcceptMasks() == m).all() and (r.States() == s_).all(), (name, max_hot, variant)
print("ok lines", flush=True)
# prefix and suffix scans (all four kernels), ragged + fixed, tiny hot set
from refpire import oracle_prefix, oracle_suffix
for name in ("headline", "count_words5"):
    image = W.load_image(name)
    orc = Oracle(image)
    strings = [b"", b"x", b"hello  world", b"ab cd" * 40, b"a" * 33, b"zz hello\tworld", b"q" * 100] * 30
    c, o = csr(strings)
    bb = P.Batch.from_strings(strings)
    for max_hot in (255, 2):
        sc = P.Scanner(image, 0)
        sc.set_max_hot(max_hot)
        for shortest in (False, True):
            for m1 in (False, True):
                for m2 in (False, True):
                    got = (P.ShortestPrefix if shortest else P.LongestPrefix)(sc, bb, throughBeginMark=m1, throughEndMark=m2)
                    assert (got == oracle_prefix(orc, c, o, shortest=shortest, through_begin=m1, through_end=m2)).all()
                    got = (P.ShortestSuffix if shortest else P.LongestSuffix)(sc, bb, throughEndMark=m1, throughBeginMark=m2)
                    assert (got == oracle_suffix(orc, c, o, shortest=shortest, through_end=m1, through_begin=m2)).all()
    # fixed-length, 32-byte aligned batch: the uniform prefix kernel (no ring; chunks walked in final states only)
    rng_u = np.random.default_rng(7)
    nu, lu = 1024 + 5, 96
    hu = rng_u.choice(np.frombuffer(b"helo wrd\tab", np.uint8), size=(nu, lu))
    hu[::3, :12] = np.frombuffer(b"hello  world", np.uint8)
    hu = np.ascontiguousarray(hu).reshape(-1)
    bu = P.Batch(torch.from_numpy(hu).to(dev), fixed_len=lu, n=nu)
    scu = P.Scanner(image, 0)
    for shortest in (False, True):
        for m1 in (False, True):
            got = (P.ShortestPrefix if shortest else P.LongestPrefix)(scu, bu, throughBeginMark=m1, throughEndMark=True)
            assert (got == oracle_prefix(orc, hu, fixed_len=lu, n=nu, shortest=shortest, through_begin=m1, through_end=True)).all()
    print("ok prefix/suffix", name, flush=True)
# counting kernels (HalfFinalScanner): accept lists / packed / packed on every chunk, tiny hot sets, ragged + fixed
from refpire import oracle_count
for name in ("hf_glue10", "count_words5"):
    image = W.load_image(name)
    orc = Oracle(image)
    strings = [b"", b"x", b"GET /a error timeout", b"hello  world fatal https://x", b"ab cd" * 40, b"a" * 33, b"zz"] * 40
    c, o = csr(strings)
    want, wfin = oracle_count(orc, c, o)
    bb = P.Batch.from_strings(strings)
    for max_hot in (255, 2):
        sc = P.Scanner(image, 0)
        sc.set_max_hot(max_hot)
        for mode in (1, 2, 3):
            sc.set_count_mode(mode)
            res = P.HalfFinalCount(sc, bb)
            assert (res.counts == want).all() and (res.final == wfin.astype(bool)).all(), (name, max_hot, mode)
    spec = W.SynthSpec(1024 + 5, 256, plants=W.GLUE10_PLANTS)
    d = torch.empty(spec.total_bytes(), dtype=torch.uint8, device=dev)
    spec.fill_device(d)
    host = spec.host_sample(0, 1024 + 5)
    want, wfin = oracle_count(orc, host, fixed_len=256, n=1024 + 5)
    sc = P.Scanner(image, 0)
    fbatch = P.Batch(d, fixed_len=256, n=1024 + 5)
    sc.Tune(fbatch, 512)
    res = P.HalfFinalCount(sc, fbatch)
    assert (res.counts == want).all() and (res.final == wfin.astype(bool)).all(), name
    print("ok counting", name, flush=True)
# round 2: register-streaming bodies of the prefix / counting kernels (fixed length, multiple of 32, partial last warp),
# accept sets, the streaming host entry point (1 MiB chunks, pageable input), the one-rank sharded call
image = W.load_image("headline")
orc = Oracle(image)
sc = P.Scanner(image, 0)
spec = W.SynthSpec(1024 + 7, 160, plants=W.HEADLINE_PLANTS)
d = torch.empty(spec.total_bytes(), dtype=torch.uint8, device=dev)
spec.fill_device(d)
host = spec.host_sample(0, 1024 + 7)
fbatch = P.Batch(d, fixed_len=160, n=1024 + 7)
for shortest in (False, True):
    got = (P.ShortestPrefix if shortest else P.LongestPrefix)(sc, fbatch, throughBeginMark=True, throughEndMark=True)
    assert (got == oracle_prefix(orc, host, fixed_len=160, n=1024 + 7, shortest=shortest, through_begin=True, through_end=True)).all()
image = W.load_image("hf_glue10")
sc_hf = P.Scanner(image, 0)
want, wfin = oracle_count(Oracle(image), host, fixed_len=160, n=1024 + 7)
for mode in (1, 2, 3):
    sc_hf.set_count_mode(mode)
    res = P.HalfFinalCount(sc_hf, fbatch)
    assert (res.counts == want).all(), mode
print("ok uniform bodies", flush=True)
import ctypes as C
r = P.Runner(sc).Begin().Run(fbatch).End()
states = torch.from_numpy(r.States().astype(np.int32)).to(dev)
sets = torch.zeros(1024 + 7, dtype=torch.int32, device=dev)
N.check(N.lib.pire_gpu_accept_sets(sc._h, states.data_ptr(), 1024 + 7, sets.data_ptr(), None), "accept sets")
assert (sets.cpu().numpy().view(np.uint32) == r.AcceptMasks()).all()
os.environ["PIRE_B200_HOST_CHUNK_MB"] = "1"
big = W.SynthSpec(6000, 1024, plants=W.HEADLINE_PLANTS)
hbig = big.host_sample(0, 6000)
bits, masks, _ = sc.run_batch_host(hbig, fixed_len=1024, n=6000, want_masks=True)
f, m, _ = orc.run(hbig, fixed_len=1024, n=6000)
assert (np.unpackbits(bits.view(np.uint8), bitorder="little")[:6000] == f).all() and (masks == m).all()
from pire_b200.dist import Comm
comm = Comm(0, rank=0, world=1)
ball = torch.full((comm.words(1024 + 7),), -1, dtype=torch.int32, device=dev)
comm.run_sharded(sc, fbatch, 1024 + 7, N.RUN_BEGIN | N.RUN_END, ball)
torch.cuda.synchronize()
assert (np.unpackbits(ball.cpu().numpy().view(np.uint8), bitorder="little")[:1024 + 7] == r.Matches().astype(np.uint8)).all()
comm.close()
print("ok accept sets, host streaming, sharded(1)", flush=True)
torch.cuda.synchronize()
print("sanitize_run done, launches:", N.lib.pire_gpu_launch_count())
