#!/usr/bin/env python
"""Compiles the BASELINE.json pattern sets with the reference's own front end
(Lexer -> Fsm -> Compile -> Scanner::Glue, unchanged host code) and stores the
Scanner::Save() images under pire_b200/data/ -- the "precompiled scanner" form the
reference itself ships around (samples/blacklist/blacklist.cpp:65-93).

Needs oracle/_ref/libpire_ref.so (i.e. /root/reference at build time).  bench.py
and the GPU box only ever read the stored images.
"""
import lzma
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from refpire import Ref  # noqa: E402


def main():
    # workloads.py is data only; import it without pulling in the CUDA library
    import importlib.util
    spec = importlib.util.spec_from_file_location("workloads_data", os.path.join(ROOT, "pire_b200", "workloads.py"))
    src = open(spec.origin).read().replace("from . import _native as N", "N = None")
    ns = {}
    exec(compile(src, spec.origin, "exec"), ns)
    ref = Ref()
    out = os.path.join(ROOT, "pire_b200", "data")
    os.makedirs(out, exist_ok=True)
    for name, pats in (("headline", [ns["HEADLINE"]]), ("glue10", ns["GLUE10"]), ("headline_iu", [ns["HEADLINE_IU"]])):
        sc = ref.glue_all(pats)
        img = sc.save()
        path = os.path.join(out, name + ".pire.xz")
        with open(path, "wb") as f:
            f.write(lzma.compress(img, preset=9))
        print("%s: %d patterns -> %d states x %d letters, image %d bytes (%d compressed)"
              % (name, len(pats), sc.size, sc.letters, len(img), os.path.getsize(path)))

    # HalfFinalScanner images (pire/scanners/half_final.h) for the counting entry point:
    #  hf_glue10     the same ten patterns, each as HalfFinalScanner(fsm), glued (half_final.h:196-198)
    #  count_words5  the five HalfFinalFsm counters of tests/count_ut.cpp:503-520 for [a-z]+, glued
    hf = None
    for pat, opts in ns["GLUE10"]:
        one = ref.compile_half_final(pat, opts, 0)
        hf = one if hf is None else ref.glue_half_final(hf, one)
    words = None
    for mode in (1, 2, 3, 4, 5):
        one = ref.compile_half_final(b"[a-z]+", "n", mode)
        words = one if words is None else ref.glue_half_final(words, one)
    for name, sc in (("hf_glue10", hf), ("count_words5", words)):
        assert not sc.empty
        img = sc.save()
        path = os.path.join(out, name + ".pire.xz")
        with open(path, "wb") as f:
            f.write(lzma.compress(img, preset=9))
        print("%s: %d states, %d regexps, image %d bytes (%d compressed)" % (name, sc.size, sc.regexps, len(img), os.path.getsize(path)))


if __name__ == "__main__":
    main()
