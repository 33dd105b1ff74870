"""include/pire_gpu.hpp compiled against the reference headers: the reference's own
run.h templates (Step, Run, Runner, LongestPrefix, ShortestPrefix) instantiated on
Pire::Gpu::Scanner must agree with the reference scanner.  Needs /root/reference
(headers) and oracle/_ref (library), i.e. runs in the build container only."""
import os
import subprocess

import pytest

from conftest import ROOT

REF = os.environ.get("PIRE_REF", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pire")), reason="reference headers not present")
def test_reference_templates_run_on_gpu_scanner(tmp_path):
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libpire_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built")
    exe = str(tmp_path / "mirror_check")
    cmd = ["g++", "-std=c++11", "-O1", "-w", "-DPIRE_NO_CONFIG", "-I", REF, "-I", os.path.join(REF, "pire"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_check.cpp"),
           ref_so, os.path.join(ROOT, "pire_b200", "libpire_b200.so"), "-o", exe,
           "-Wl,-rpath," + os.path.dirname(ref_so), "-Wl,-rpath," + os.path.join(ROOT, "pire_b200")]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout
