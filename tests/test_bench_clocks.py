"""bench.py's clock sampler with stand-ins for NVML and nvidia-smi (neither exists in the build container)."""
import importlib
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def fake_nvml(clock=1965, mask=0):
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    m.nvmlClocksEventReasonHwSlowdown = 0x8
    m.nvmlClocksEventReasonHwThermalSlowdown = 0x40
    m.nvmlClocksEventReasonSwThermalSlowdown = 0x20
    m.nvmlClocksEventReasonSwPowerCap = 0x4
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: ("gpu", i)
    m.nvmlDeviceGetMaxClockInfo = lambda h, which: 1965
    m.nvmlDeviceGetClockInfo = lambda h, which: clock
    m.nvmlDeviceGetCurrentClocksEventReasons = lambda h: mask
    return m


def test_nvml_samples_are_preferred(monkeypatch):
    bench = load_bench()
    monkeypatch.setitem(sys.modules, "pynvml", fake_nvml(clock=1950, mask=0x4))
    s = bench.ClockSampler(0)
    time.sleep(0.05)
    t0 = time.perf_counter()
    time.sleep(0.06)
    t1 = time.perf_counter()
    out = s.stop(t0, t1)
    assert out["source"].startswith("nvml") and out["samples"] >= 5
    assert out["sm_mhz"] == 1950 and out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"]


def test_without_any_source_the_line_says_so(monkeypatch):
    bench = load_bench()
    broken = types.ModuleType("pynvml")
    monkeypatch.setitem(sys.modules, "pynvml", broken)          # no nvmlInit: NVML path fails
    monkeypatch.setenv("PATH", "/nonexistent")                    # no nvidia-smi either
    s = bench.ClockSampler(0)
    out = s.stop(0.0, 1.0)
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]


def test_short_region_uses_the_nearest_samples():
    bench = load_bench()
    rows = [(1.00, 1, []), (1.02, 2, []), (1.30, 3, [])]
    assert bench.ClockSampler._window(rows, 1.005, 1.015) == rows[:2]      # none inside: +-50 ms
    assert bench.ClockSampler._window(rows, 1.01, 1.05) == [rows[1]]
    assert bench.ClockSampler._window(rows, 5.0, 5.1) == rows[-3:]
