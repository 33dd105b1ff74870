#!/usr/bin/env python
"""bench.py -- scanned GB/s of the Pire hot path on B200 (BASELINE.json metric).

A "step" is one pass of the scan path over one batch of synthetic strings that is
already resident in HBM: Runner(sc).Begin().Run(str).End() for every string of the
batch (pire/run.h:365-392), producing the packed match bitmap and the per-string
accepted-regexp mask; with N > 1 the batch is sharded by string (weak scaling:
every GPU holds its own 10 GB shard) and the step ends with the one NCCL
all-reduce of the match bitmap.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload glue10|headline]
  python bench.py --impl reference ...     # the reference's own CPU scan (oracle/_ref)

One JSON line on stdout (rank 0).  Keys follow the driver's contract; `roofline`
is for the scan kernel (algorithmic bytes = payload bytes, 1 B read per input
byte), `e2e` goes through the host-buffer C-ABI call with pinned
host memory (H2D + D2H inside the timed region), `cpu_baseline` is the reference
library timed on this box's host cores on a bounded sample of the same corpus.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

STRING_LEN = 1024
STRINGS_PER_GPU = 9_765_632          # x 1 KiB = 10.000007 GB per GPU (multiple of 32 strings)
MIXED_STRINGS_PER_GPU = 1_281_024    # mixed 16 B..64 KiB strings, mean 7.8 KB: ~10 GB per GPU
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pire_b200", choices=["pire_b200", "reference"])
    ap.add_argument("--workload", default="glue10", choices=["glue10", "headline", "utf8mixed"])
    ap.add_argument("--strings", type=int, default=0, help="strings per GPU (default: 10 GB worth)")
    ap.add_argument("--variant", default="auto", choices=["auto", "plain", "pred", "priv", "look", "look64"])
    ap.add_argument("--no-tune", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--collective", default="allreduce", choices=["allreduce", "allgather"])
    ap.add_argument("--cpu-sample", type=int, default=1 << 22, help="strings in the CPU-baseline sample (per step of the reference arm)")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return FALLBACK_HBM_GBS, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region.

    Two sources run side by side from before the warm-up: NVML polled every 4 ms from a thread (the same
    counters nvidia-smi prints, exact timestamps), and the recipe's `nvidia-smi --query-gpu=... -lms 20`
    line as a subprocess (its lines reach us through a pipe, so their arrival times are only approximate
    and a short timed region can end before the first one arrives).  NVML samples are preferred."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc = [], None
        self.nvml_rows, self.nvml, self.nvml_max, self._halt = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.nvml_thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.nvml_thread.start()
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._halt:
            try:
                t = time.perf_counter()
                sm = float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                mask = int(reasons_fn(self._handle))
                self.nvml_rows.append((t, sm, [name for bit, name in bits if mask & bit]))
            except Exception:
                break
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    @staticmethod
    def _window(rows, t0, t1):
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if inside:
            return inside
        near = [r for r in rows if t0 - 0.05 <= r[0] <= t1 + 0.05]     # region shorter than the sampling period
        return near or rows[-3:]

    def stop(self, t0, t1):
        if self.nvml:
            self._halt = True
            self.nvml_thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        rows = self._window(self.nvml_rows, t0, t1) if self.nvml_rows else []
        if rows:
            reasons = sorted({name for _, _, names in rows for name in names})
            return {"sm_mhz": statistics.median(sm for _, sm, _ in rows), "sm_max_mhz": self.nvml_max, "reasons": reasons,
                    "samples": len(rows), "source": "nvml, polled every 4 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        rows = self._window(self.rows, t0, t1)
        for _, r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(self.NAMES, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows), "source": "nvidia-smi -lms 20"}


class PortScanner:
    """Stand-in with RefScanner.run's signature over the oracle port (oracle/pire_oracle.c), used
    only when oracle/_ref (the compiled reference) is not on this box.  Python threads over slices:
    the C call drops the GIL."""

    def __init__(self, image):
        from refpire import Oracle
        self.orc = Oracle(image)

    def run(self, corpus, offsets=None, fixed_len=0, n=None, variant=1, threads=1, want=("final", "mask")):
        import numpy as np
        from concurrent.futures import ThreadPoolExecutor
        threads = max(1, min(threads, n // 1024 or 1))
        final = np.zeros(n, np.uint8)
        mask = np.zeros(n, np.uint32)

        def part(k):
            lo, hi = n * k // threads, n * (k + 1) // threads
            if offsets is not None:
                f, m, _ = self.orc.run(corpus, offsets[lo:hi + 1], n=hi - lo, shortcuts=variant != 2)
            else:
                f, m, _ = self.orc.run(corpus[lo * fixed_len:hi * fixed_len], fixed_len=fixed_len, n=hi - lo, shortcuts=variant != 2)
            final[lo:hi] = f
            mask[lo:hi] = m
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(part, range(threads)))
        return final, mask, None


class PortRef:
    kind = "port"

    def hardware_threads(self):
        return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def scanner(self, workload):
        from pire_b200 import workloads as W
        return PortScanner(W.load_image(W.WORKLOADS[workload][0]))


def get_reference():
    """The compiled reference (oracle/_ref) when it is on this box, else the oracle port."""
    from refpire import Ref, have_ref
    if have_ref():
        ref = Ref()
        ref.kind = "reference"
        return ref
    return PortRef()


def cpu_reference(workload, threads, n_sample, reps, first_string=0):
    """The reference's own scan, Runner(sc).Begin().Run().End() per string with NonrelocScanner
    (its fastest variant, multi.h:1119-1123), statically partitioned over `threads` host threads."""
    from pire_b200 import workloads as W
    ref = get_reference()
    sc, sample, offsets, n_sample, gb = reference_sample(ref, workload, n_sample, first_string)
    kw = dict(offsets=offsets) if offsets is not None else dict(fixed_len=STRING_LEN)
    if threads <= 0:
        threads = ref.hardware_threads()
    best, matches = 1e30, 0
    for _ in range(reps):
        t0 = time.perf_counter()
        final, mask, _ = sc.run(sample, n=n_sample, variant=1, threads=threads, want=("final", "mask"), **kw)
        best = min(best, time.perf_counter() - t0)
        matches = int(final.sum())
    # single-thread figures on a slice, both with and without the ExitMasks fast-forward
    k = min(n_sample, 1 << 16 if offsets is None else 1 << 13)
    kb = (k * STRING_LEN if offsets is None else int(offsets[k])) / 1e9
    t0 = time.perf_counter()
    sc.run(sample, n=k, variant=1, threads=1, want=("final",), **kw)
    t_mask = time.perf_counter() - t0
    t0 = time.perf_counter()
    sc.run(sample, n=k, variant=2, threads=1, want=("final",), **kw)
    t_nomask = time.perf_counter() - t0
    return {
        "value": gb / best, "unit": "GB/s", "cores": threads, "kind": ref.kind,
        "sample": "%d strings of the same synthetic corpus (%.2f GB), best of %d, %s, static partition by string count" % (
            n_sample, gb, reps, "NonrelocScanner" if ref.kind == "reference" else "oracle port (oracle/_ref absent)"),
        "matches": matches,
        "one_thread_GBps": kb / t_mask,
        "one_thread_nomask_GBps": kb / t_nomask,
    }


def reference_sample(ref, workload, n_sample, first_string=0):
    """The reference scanner for a workload plus a host-generated sample of its corpus."""
    from pire_b200 import workloads as W
    port = getattr(ref, "kind", "reference") == "port"
    if workload == "utf8mixed":
        sc = ref.scanner(workload) if port else ref.compile(*W.HEADLINE_IU)
        n_sample = min(n_sample, 1 << 17)          # mean string is 7.8 KB: ~1 GB
        sample, offsets = W.MixedSpec(n_sample, first_string=first_string).host_batch(0, n_sample)
        return sc, sample, offsets, n_sample, int(offsets[-1]) / 1e9
    sc = ref.scanner(workload) if port else ref.glue_all(W.GLUE10 if workload == "glue10" else [W.HEADLINE])
    spec = W.SynthSpec(n_sample, STRING_LEN, plants=W.WORKLOADS[workload][1], first_string=first_string)
    return sc, spec.host_sample(0, n_sample), None, n_sample, n_sample * STRING_LEN / 1e9


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = min(args.cpu_sample, args.strings or args.cpu_sample)
    t_all = time.perf_counter()
    cb = cpu_reference(args.workload, 0, min(per_step, 1 << 20), 1)        # warm-up + the single-thread figures
    times = []
    from pire_b200 import workloads as W
    ref = get_reference()
    sc, sample, offsets, per_step, gb = reference_sample(ref, args.workload, per_step)
    kw = dict(offsets=offsets) if offsets is not None else dict(fixed_len=STRING_LEN)
    threads = ref.hardware_threads()
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        sc.run(sample, n=per_step, variant=1, threads=threads, want=("final", "mask"), **kw)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = gb * args.steps / total
    cb.update(value=value, cores=threads, kind=ref.kind,
              sample="%d strings of the same synthetic corpus (%.2f GB) per step, %s" % (
                  per_step, gb, "NonrelocScanner" if ref.kind == "reference" else "oracle port"))
    line = {
        "impl": "reference", "metric": "scanned GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s: %s; each step = a bounded sample of %d strings (%.2f GB) on the host" % (
            args.workload, W.WORKLOADS[args.workload][2], per_step, gb), "host_threads": threads},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "strings_per_s": per_step * args.steps / total,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import pire_b200 as P
    from pire_b200 import _native as N
    from pire_b200 import workloads as W
    from pire_b200.dist import popcount_bits

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scan path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    image_name, plants, desc, cfg_index = W.WORKLOADS[args.workload]
    mixed = args.workload == "utf8mixed"
    n_local = (args.strings or (MIXED_STRINGS_PER_GPU if mixed else STRINGS_PER_GPU)) // 32 * 32
    n_global = n_local * world
    lo = rank * n_local
    if mixed:
        # BASELINE configs[3]: CSR batch of very unequal strings; binned by length (on the device) once
        spec = W.MixedSpec(n_local, first_string=lo)
        corpus, offsets = spec.device_batch(dev)
        batch = P.Batch(corpus, offsets, n=n_local)
        payload_local = batch.payload_bytes()
        t0 = time.perf_counter()
        batch.bin_by_length()
        torch.cuda.synchronize()
        bin_ms = 1e3 * (time.perf_counter() - t0)
    else:
        spec = W.SynthSpec(n_local, STRING_LEN, plants=plants, first_string=lo)
        corpus = torch.empty(spec.total_bytes(), dtype=torch.uint8, device=dev)
        spec.fill_device(corpus)
        batch = P.Batch(corpus, fixed_len=STRING_LEN, n=n_local)
        payload_local = n_local * STRING_LEN
        bin_ms = None
    payload_global = payload_local
    if world > 1:
        t = torch.tensor([payload_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        payload_global = int(t.item())

    sc = P.Scanner(W.load_image(image_name), local)
    tune_ms = None
    if not args.no_tune:
        t0 = time.perf_counter()
        sc.Tune(batch, min(n_local, 16384))
        torch.cuda.synchronize()
        tune_ms = 1e3 * (time.perf_counter() - t0)

    words_local = n_local // 32
    bits_full = torch.zeros(words_local * world, dtype=torch.int32, device=dev)
    bits_local = bits_full[rank * words_local:(rank + 1) * words_local]
    masks = torch.empty(n_local, dtype=torch.int32, device=dev)
    flags = N.RUN_BEGIN | N.RUN_END

    def scan():
        sc.run_batch(batch, flags, bits_local, masks, None)

    def step():
        if world > 1 and args.collective == "allreduce":
            # shards are disjoint and word aligned: SUM over zero-initialised words == OR
            if rank > 0:
                bits_full[: rank * words_local].zero_()
            if rank < world - 1:
                bits_full[(rank + 1) * words_local:].zero_()
        scan()
        if world > 1:
            if args.collective == "allreduce":
                dist.all_reduce(bits_full, op=dist.ReduceOp.SUM)
            else:
                dist.all_gather_into_tensor(bits_full, bits_local.clone())

    def time_scan(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            scan()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # kernel variant: measured on this batch by the library, not guessed
    names = dict(N.VARIANT_NAMES)
    variant_ms = {}
    if args.variant == "auto":
        variant_ms = sc.AutoSelect(batch)
        chosen = names[sc.info().variant]
    else:
        chosen = args.variant
        sc.set_variant({v: k for k, v in names.items()}[chosen])

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("PIRE_B200_NO_CLOCKS") else None        # started early: nvidia-smi needs ~100 ms to begin
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = N.lib.pire_gpu_launch_count()
    barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        step()
        marks[i].record()
    e1 = marks[-1] if marks else e1
    if not marks:
        e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    per_step = sorted((marks[i - 1] if i else e0).elapsed_time(marks[i]) for i in range(args.steps))
    launches = N.lib.pire_gpu_launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = payload_global / 1e9 / (ms_per_step / 1e3)

    # the scan kernel alone (CUDA events around the launches only), for the roofline
    kernel_ms = time_scan(min(args.steps, 10))
    matches_global = popcount_bits(bits_full if world > 1 else bits_local)
    matches_local_masks = int((masks != 0).sum().item())
    # consistency: the all-reduced bitmap holds exactly the union of the shards' matches,
    # and every planted string (1/8) is reported
    local_pop = torch.tensor([popcount_bits(bits_local)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(local_pop, op=dist.ReduceOp.SUM)
    assert int(local_pop.item()) == matches_global, (int(local_pop.item()), matches_global)
    assert matches_local_masks == popcount_bits(bits_local)
    assert matches_local_masks >= (n_local // 8) * (0.9 if mixed else 1.0)      # strings < 32 B carry no plant

    # end to end: host buffers through the C ABI, H2D and D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty(payload_local, dtype=torch.uint8, pin_memory=True)
            host.copy_(corpus[:payload_local])
            host_offs = None
            if mixed:
                host_offs = torch.empty(n_local + 1, dtype=torch.int64, pin_memory=True)
                host_offs.copy_(batch.offsets)
            torch.cuda.synchronize()
            hb = torch.empty(words_local, dtype=torch.int32, pin_memory=True)
            hm = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
            hv = host.numpy()

            def e2e_step():
                N.check(N.lib.pire_gpu_run_batch_host(sc._h, hv.ctypes.data, payload_local,
                                                      host_offs.data_ptr() if mixed else None,
                                                      0 if mixed else STRING_LEN, n_local,
                                                      flags, hb.data_ptr(), hm.data_ptr(), None), "run_batch_host")
            e2e_step()
            e2e_steps = max(1, min(args.steps, 3))
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            assert torch.equal(hb, bits_local.cpu()) and torch.equal(hm, masks.cpu())
            e2e = {"value": payload_global / 1e9 / (dt / e2e_steps), "unit": "GB/s",
                   "h2d_bytes_per_step": payload_local + (8 * (n_local + 1) if mixed else 0),
                   "d2h_bytes_per_step": words_local * 4 + n_local * 4,
                   "steps": e2e_steps, "api": "pire_gpu_run_batch_host (pinned host corpus in, bitmap + accept masks out)"}
            del host, hv
        except Exception as ex:          # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "GB/s", "error": repr(ex)}

    # the other fixed-length BASELINE configuration on the same resident bytes (a few launches)
    also = None
    if not mixed:
        try:
            other = "headline" if args.workload == "glue10" else "glue10"
            sc2 = P.Scanner(W.load_image(W.WORKLOADS[other][0]), local)
            sc2.Tune(batch, min(n_local, 16384))
            ms2 = sc2.AutoSelect(batch)
            best2 = min(ms2, key=ms2.get)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                sc2.run_batch(batch, flags, bits_local, masks, None)
            e1.record()
            torch.cuda.synchronize()
            gbps2 = payload_local / 1e9 / (e0.elapsed_time(e1) / 5 / 1e3)
            also = {"workload": "BASELINE configs[%d]: %s (same resident corpus, this GPU only)" % (W.WORKLOADS[other][3], W.WORKLOADS[other][2]),
                    "value": gbps2, "unit": "GB/s", "kernel_variant": best2}
            scan()           # restore this workload's outputs in bits_local / masks
            torch.cuda.synchronize()
            del sc2
        except Exception as ex:      # noqa: BLE001
            also = {"error": repr(ex)}

    # the rows next to the path on the same resident bytes, a few launches each: LongestPrefix with this
    # workload's automaton and HalfFinalScanner counting with the ten patterns glued as HalfFinalScanners
    next_rows = None
    if not mixed:
        try:
            lens = torch.empty(n_local, dtype=torch.int32, device=dev)
            hf = P.Scanner(W.load_image("hf_glue10"), local)
            counts = torch.empty((n_local, hf.RegexpsCount()), dtype=torch.int32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream

            def prefix_run():
                N.check(N.lib.pire_gpu_prefix_batch(sc._h, corpus.data_ptr(), None, STRING_LEN, n_local, flags, 0, lens.data_ptr(), stream),
                        "pire_gpu_prefix_batch")

            def count_run():
                N.check(N.lib.pire_gpu_count_batch(hf._h, corpus.data_ptr(), None, STRING_LEN, n_local, flags, counts.data_ptr(), None,
                                                   stream), "pire_gpu_count_batch")
            next_rows = {}
            for key, fn in (("longest_prefix", prefix_run), ("half_final_count_hf_glue10", count_run)):
                fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(3):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                next_rows[key] = {"value": payload_local / 1e9 / (e0.elapsed_time(e1) / 3 / 1e3), "unit": "GB/s"}
            next_rows["note"] = "this GPU only, same resident corpus; parity of these entry points is in tests/test_gpu_parity.py"
            del lens, counts, hf
        except Exception as ex:      # noqa: BLE001
            next_rows = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    if also and "value" in also:
        also["frac_of_hbm_peak"] = also["value"] / peak
    achieved = payload_local / 1e9 / (kernel_ms / 1e3)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass
    info = sc.info()
    line = {
        "metric": "scanned GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "step_ms": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]} if per_step else None,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[%d]: %s; %d synthetic strings per GPU (%.2f GB/GPU), "
                        "1/8 of the strings carry a planted match" % (
                            cfg_index, desc if mixed else desc + " over 1 KiB printable-ASCII strings", n_local, payload_local / 1e9),
            "strings_per_gpu": n_local, "string_len": "16..65535 (CSR, binned by length on device)" if mixed else STRING_LEN,
            "bin_ms": bin_ms, "outputs": "match bitmap + u32 accept mask per string",
            "kernel_variant": chosen, "variant_ms": variant_ms or None, "hot_rows": info.hot_rows, "tuned": bool(info.tuned),
            "tune_ms": tune_ms, "l2": "input (%.1f GB) is far larger than L2; no flush needed" % (payload_local / 1e9),
            "collective": (args.collective + " of the match bitmap (NCCL)") if world > 1 else "none (1 GPU)",
        },
        "strings_per_s": n_global / (ms_per_step / 1e3),
        "matches": matches_global, "matches_local_by_mask": matches_local_masks,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": ("ScanGenericKernel<%s>" % chosen) if mixed else "ScanUniformPrivKernel" if chosen == "priv" else "ScanUniformKernel<%s>" % chosen, "kernel_ms": kernel_ms,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": payload_local},
        "clocks": clocks,
        "e2e": e2e,
        "also": also,
        "next_rows": next_rows,
    }
    if not args.no_cpu and world == 1:
        try:
            line["cpu_baseline"] = cpu_reference(args.workload, 0, min(args.cpu_sample, n_local), 3)
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
