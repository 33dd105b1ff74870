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


@pytest.mark.gpu
def test_sharded_entry_from_cpp(tmp_path, cuda_device):
    """The multi-GPU entry of the C ABI from plain C++ (no Python in the data path): tests/cpp/sharded_check.cpp through
    include/pire_gpu.hpp's Comm, one rank per visible GPU up to two; the gathered bitmap must equal a single-GPU run."""
    import shutil
    import torch
    from pire_b200 import workloads as W
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not present")
    exe = str(tmp_path / "sharded_check")
    lib_dir = os.path.join(ROOT, "pire_b200")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "sharded_check.cpp"),
                    os.path.join(lib_dir, "libpire_b200.so"), "-o", exe, "-Xlinker", "-rpath=" + lib_dir], check=True)
    image = tmp_path / "glue10.pire"
    image.write_bytes(W.load_image("glue10"))
    for world in sorted({1, min(2, torch.cuda.device_count())}):
        for n in (100_000, 70):                        # 70 strings: rank 1 of 2 holds 6, the slack words must be zero
            out = subprocess.run([exe, str(image), str(n), str(world)], capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stdout + out.stderr
            assert ": 0 mismatches" in out.stdout
