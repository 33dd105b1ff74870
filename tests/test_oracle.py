"""The oracle (oracle/pire_oracle.c) pinned against the reference:
 * the committed golden fixtures (generated from the real reference by
   tests/golden/make_golden.py) -- always;
 * the real reference compiled from /root/reference (oracle/_ref) -- when present.
"""
import numpy as np
import pytest

from conftest import GOLDEN
from pire_ut_vectors import VECTORS
from refpire import Oracle, csr


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: c.name)
def test_oracle_matches_golden(case):
    orc = Oracle(case.image)
    assert (orc.states, orc.letters) == (case.states, case.letters) or orc.empty
    corpus, offs = csr(case.strings)
    for shortcuts in (False, True):
        final, mask, state = orc.run(corpus, offs, begin=case.begin, end=case.end, shortcuts=shortcuts)
        assert final.tolist() == case.final
        assert mask.tolist() == case.mask()
        assert state.tolist() == case.state


def test_oracle_alignment_independent():
    # pire_ut.cpp Aligned@729: results must not depend on where a string starts
    case = next(c for c in GOLDEN if c.name == "Aligned@729b")
    orc = Oracle(case.image)
    for shift in range(0, 17):
        strings = [b"J" * shift] + case.strings
        corpus, offs = csr(strings)
        final, _, _ = orc.run(corpus, offs, shortcuts=True)
        assert final.tolist()[1:] == case.final


def test_oracle_rejects_bad_images():
    case = GOLDEN[0]
    with pytest.raises(ValueError):
        Oracle(b"\0" * 128)
    with pytest.raises(ValueError):
        Oracle(case.image[: len(case.image) // 2])


def test_oracle_vs_reference_live(ref):
    """Every vector group, compiled afresh by the real reference, incl. random text."""
    rng = np.random.default_rng(7)
    for name, pat, opts, acc, den in VECTORS:
        sc = ref.compile(pat, opts)
        orc = Oracle(sc.save())
        noise = [bytes(rng.integers(0x20, 0x7F, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 200, size=64)]
        # splice the known strings into noise so that longer walks hit the accept paths too
        mixed = [noise[i] + s + noise[-i - 1] for i, s in enumerate(acc + den)]
        corpus, offs = csr(acc + den + noise + mixed)
        for begin, end in ((True, True), (False, False), (True, False), (False, True)):
            f_ref, m_ref, s_ref = sc.run(corpus, offs, begin=begin, end=end, variant=0)
            for shortcuts in (False, True):
                f, m, s = orc.run(corpus, offs, begin=begin, end=end, shortcuts=shortcuts)
                assert (f == f_ref).all() and (m == m_ref).all() and (s == s_ref).all(), name


def test_reference_variants_agree(ref):
    # NonrelocScanner (ExitMasks) and NonrelocScannerNoMask are the CPU baselines bench.py times
    from pire_b200 import workloads as W
    sc = ref.compile(*W.HEADLINE)
    spec = W.SynthSpec(512, 1024, plants=W.HEADLINE_PLANTS)
    corpus = spec.host_sample(0, 512)
    f0, _, s0 = sc.run(corpus, fixed_len=1024, variant=0)
    f1, _, _ = sc.run(corpus, fixed_len=1024, variant=1)
    f2, _, _ = sc.run(corpus, fixed_len=1024, variant=2, threads=4)
    assert (f0 == f1).all() and (f0 == f2).all()
    assert int(f0.sum()) == 512 // 8          # exactly the planted strings match
    orc = Oracle(sc.save())
    f, _, s = orc.run(corpus, fixed_len=1024, shortcuts=True)
    assert (f == f0).all() and (s == s0).all()
