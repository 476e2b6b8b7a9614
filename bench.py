#!/usr/bin/env python
"""bench.py — the contract benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: a complete 4-pass
32-bit LSD radix sort (vrdxCmdSort through the C-ABI).

N = 1   workload = BASELINE.json configs[1]: 32-bit keys-only, 2^28 uniform keys
        (DataGenerator(seed=1), i.e. the reference's bench/data_generator.cc stream).
        value  = GKeys/s with the keys resident in HBM (CUDA events around the sort only);
        e2e    = same metric through the same C-ABI call with HOST buffers: pinned-host -> device
                 copy, sort, device -> pinned-host copy, all inside the timed region.
        extra  = key-value (configs[2]) and 2^25 numbers, for the record.
N > 1   workload = BASELINE.json configs[4]: distributed keys-only sort, 2^29 keys per GPU,
        MSD-bucket partition + all-to-all over NVLink + local sort (weak scaling).

--impl reference times the reference's own CPU implementation of the path
(CpuBenchmark::Sort = std::sort, bench/cpu_benchmark.cc:19-28, compiled unmodified into
oracle/_ref) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GKeys/s (32-bit keys-only, N=2^28 uniform, 1xB200; key-value and 2^25 in extra)"
UNIT = "GKeys/s"
BYTES_PER_KEY_KEYS = 36   # 4 (histogram read) + 4 passes x (4 R + 4 W)      SURVEY.md §8(d)
BYTES_PER_KEY_KV = 68     # 4 + 4 x (8 R + 8 W)
PASS_BYTES_PER_KEY_KEYS = 8
PASS_BYTES_PER_KEY_KV = 16


# ------------------------------------------------------------------------------ helpers

def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(kind: str, n: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, scaled per key."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)[kind]
        return float(d["dram_bytes_per_key"]) * n
    except Exception:
        return None


def measure_dram_traffic_with_ncu(log2n: int, kv: bool, timeout_s: int = 300):
    """DRAM bytes moved by ONE sort, per kernel, measured live on this box: tools/ncu_one.py (one sort of
    resident data through the C-ABI) under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`.
    Traffic is a counter, not a time, so taking it under the profiler is legitimate; nothing timed comes
    from this run.  Returns {"kernels": {name: bytes}, "whole_sort": bytes} or None."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           sys.executable, os.path.join(ROOT, "tools", "ncu_one.py"), str(log2n), "kv" if kv else "keys", "1"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT).stdout
    except Exception:
        return None
    start = out.find('"ID"')
    if start < 0:
        return None
    per_kernel, order = {}, []
    for row in csv.DictReader(io.StringIO(out[start:])):
        name = row.get("Kernel Name", "")
        if not any(t in name for t in ("PassKernel", "Upsweep", "Spine", "Histogram", "StampStart", "CopyBack")):
            continue   # torch's own kernels (the restore copy) are not part of the sort
        try:
            val = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "byte").lower()
        val *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        key = name.split("(")[0]
        if key not in per_kernel:
            order.append(key)
        per_kernel[key] = per_kernel.get(key, 0.0) + val
    if not per_kernel:
        return None
    return {"kernels": {k: per_kernel[k] for k in order}, "whole_sort": sum(per_kernel.values())}


def probe_vulkan():
    """Is there a Vulkan loader / NVIDIA ICD on this box?  (north_star: time the reference's Vulkan path
    'if the driver exposes Vulkan'; SURVEY section 8c found none in the build container.)"""
    import ctypes.util
    import glob
    import shutil
    icds = sorted(glob.glob("/usr/share/vulkan/icd.d/*.json") + glob.glob("/etc/vulkan/icd.d/*.json"))
    loader = ctypes.util.find_library("vulkan")
    loaded = False
    for cand in ([loader] if loader else []) + ["libvulkan.so.1"]:
        try:
            ctypes.CDLL(cand)
            loaded = True
            break
        except OSError:
            pass
    return {"loader": loader, "loader_loads": loaded, "icd_files": icds, "vulkaninfo": shutil.which("vulkaninfo"),
            "nvidia_icd": any("nvidia" in os.path.basename(p).lower() for p in icds),
            "reference_vulkan_path": "n/a: no Vulkan loader/ICD on this box" if not (loaded and icds) else "loader and ICD present"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def gen_keys(n: int, seed: int = 1):
    from vulkan_radix_sort_b200.datagen import DataGenerator
    return DataGenerator(seed).generate(n, 32)[0]


# ------------------------------------------------------------------------------ reference arm

def run_reference(args):
    """The reference's CPU path (std::sort / CpuBenchmark::Sort) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_oracle
    import numpy as np
    steps, warmup = args.steps, args.warmup
    if cpu_oracle.have_ref():
        kind = "reference"
        sort = lambda k: cpu_oracle.ref_sort_keys(k)[1] / 1e9      # its own timed region (sort only)
    else:
        kind = "port"
        def sort(k):
            t0 = time.perf_counter(); cpu_oracle.sort_keys(k); return time.perf_counter() - t0
    # Same config as our arm at N=1: the FULL 2^28-key workload per step (std::sort needs ~20-25 s for it, so
    # the step count is capped internally to keep the whole run within a few minutes; ms_per_step is per
    # full sort).  For N>1 (2^29 keys per GPU x N) the CPU arm sorts the same 2^28-key sample and says so.
    log2s = args.log2n if args.log2n_reference is None else args.log2n_reference
    n = 1 << log2s
    est_s = n * np.log2(n) / (2 ** 25 * 25 / 3.9)     # std::sort here: 2^25 keys in ~3.9 s, n log n scaling
    warmup_done = min(warmup, 1 if est_s > 5 else warmup)
    steps_done = max(1, min(steps, int(150.0 / max(est_s, 1e-3)) - warmup_done)) if est_s > 5 else steps
    keys = gen_keys(n)
    for _ in range(warmup_done):
        sort(keys)
    times = [sort(keys) for _ in range(steps_done)]
    mean_s = sum(times) / len(times)
    value = n / mean_s / 1e9
    same = (args.gpus == 1 and log2s == 28)
    sample = (f"the full 2^{log2s}-key workload per step ({steps_done} timed step(s), {warmup_done} warm-up: capped "
              f"internally from --steps {steps} --warmup {warmup}); std::sort, 1 thread (the reference's CpuBenchmark::Sort "
              f"is single-threaded, bench/cpu_benchmark.cc:19-28)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps_done, "warmup": warmup_done, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"32-bit keys-only, N=2^{log2s} uniform random (DataGenerator seed 1), CpuBenchmark::Sort"
                               + ("" if args.gpus == 1 else f" (CPU sample of the {args.gpus}-GPU workload: one 2^{log2s}-key slice)"),
                   "same_config_as_our_arm": same, "host_cores": os.cpu_count(),
                   "steps_requested": steps, "warmup_requested": warmup},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ our arm, N = 1

def time_resident(sorter, torch, api, pristine, n, kv, steps, warmup, vals_pristine=None):
    """K sorts of resident data; CUDA events on the launching stream around each sort only.
    Returns (per-step ms list, mean per-pass-kernel ms, launches per step)."""
    work = torch.empty_like(pristine)
    vwork = torch.empty_like(pristine) if kv else None
    storage = sorter.storage_for(n, kv)
    res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), api.QUERY_COUNT)
    assert res == api.VK_SUCCESS
    ms, pass_ms = [], []
    launches = 0
    for it in range(warmup + steps):
        work.copy_(pristine)          # restores unsorted input AND flushes L2 (1 GiB >> 126 MB)
        if kv:
            vwork.copy_(vals_pristine)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if kv:
            sorter.sort_key_value(work, vwork, storage=storage, query_pool=pool)
        else:
            sorter.sort(work, storage=storage, query_pool=pool)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ms.append(e0.elapsed_time(e1))
            rc, ts = api.vrdxCudaGetQueryPoolResults(pool)
            pass_ms.extend((ts[4 + 3 * p] - ts[3 + 3 * p]) / 1e6 for p in range(4))
            launches = sorter.last_launch_count
    api.vrdxCudaDestroyQueryPool(pool)
    return ms, (sum(pass_ms) / len(pass_ms)), launches, work, vwork


def time_e2e(sorter, torch, host_keys, n, steps, warmup):
    """Same sort through the C-ABI with HOST buffers: H2D + sort + D2H inside the timed region."""
    pinned_in = torch.from_numpy(host_keys.view("int32")).pin_memory()
    pinned_out = torch.empty(n, dtype=torch.int32).pin_memory()
    dev = torch.empty(n, dtype=torch.int32, device="cuda")
    storage = sorter.storage_for(n, False)
    ms = []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dev.copy_(pinned_in, non_blocking=True)
        sorter.sort(dev, storage=storage)
        pinned_out.copy_(dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ms.append(e0.elapsed_time(e1))
    return ms, pinned_out


def time_e2e_kv(sorter, torch, host_keys, host_vals, n, steps, warmup):
    """Key-value sort with HOST buffers: keys and values H2D + vrdxCmdSortKeyValue + both D2H, all timed."""
    pin_k = torch.from_numpy(host_keys.view("int32")).pin_memory()
    pin_v = torch.from_numpy(host_vals.view("int32")).pin_memory()
    out_k = torch.empty(n, dtype=torch.int32).pin_memory()
    out_v = torch.empty(n, dtype=torch.int32).pin_memory()
    dk = torch.empty(n, dtype=torch.int32, device="cuda")
    dv = torch.empty(n, dtype=torch.int32, device="cuda")
    storage = sorter.storage_for(n, True)
    ms = []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dk.copy_(pin_k, non_blocking=True)
        dv.copy_(pin_v, non_blocking=True)
        sorter.sort_key_value(dk, dv, storage=storage)
        out_k.copy_(dk, non_blocking=True)
        out_v.copy_(dv, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ms.append(e0.elapsed_time(e1))
    return sum(ms) / len(ms)


def bench_config4(sorter, torch, api, cpu_oracle):
    """BASELINE.json configs[3] / SURVEY 8(d) config 4: vrdxCmdSortIndirect and ...KeyValueIndirect with the count
    in device memory, maxElementCount = 2^27, n = 2^27 - 4099 (odd, no multiple of any tile size), keys skewed
    (AND of three draws) / 8-bit / all-equal.  Resident data, CUDA events, median of 5 after 2 warm-ups; every
    result is checked (sorted + multiset; key-value: values = index, so stability is visible) and the
    [n, max) tail of the caller's buffers must be untouched."""
    import numpy as np
    from vulkan_radix_sort_b200.datagen import make_keys
    nmax = 1 << 27
    n = nmax - 4099
    out = {"max_element_count": nmax, "element_count": n}
    count = torch.tensor([n], dtype=torch.int32, device="cuda")
    idx = torch.arange(nmax, dtype=torch.int32, device="cuda")
    for dist_name in ("skewed", "bits8", "all_zero"):
        host = make_keys(dist_name, nmax, 1)
        host[n:] = 0xDEADBEEF                      # tail marker: must survive every sort
        src = torch.from_numpy(host.view(np.int32)).cuda()
        for kv in (False, True):
            keys = torch.empty_like(src)
            vals = torch.empty_like(src) if kv else None
            storage = sorter.storage_for(nmax, kv)
            ms = []
            for it in range(7):
                keys.copy_(src)
                if kv:
                    vals.copy_(idx)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if kv:
                    sorter.sort_key_value_indirect(keys, vals, count, max_count=nmax, storage=storage)
                else:
                    sorter.sort_indirect(keys, count, max_count=nmax, storage=storage)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    ms.append(e0.elapsed_time(e1))
            k = keys.cpu().numpy().view(np.uint32)
            ok = cpu_oracle.is_sorted(k[:n]) and bool((k[n:] == 0xDEADBEEF).all()) and \
                cpu_oracle.multiset_fingerprint(k[:n]) == cpu_oracle.multiset_fingerprint(host[:n])
            if kv:
                v = vals.cpu().numpy().view(np.uint32)
                ok = ok and cpu_oracle.check_stable_permutation(host[:n], k[:n], v[:n]) and \
                    bool((v[n:] == np.arange(n, nmax, dtype=np.uint32)).all())
            med = statistics.median(ms)
            out[f"{dist_name}_{'key_value' if kv else 'keys_only'}"] = {
                "value": n / (med * 1e-3) / 1e9, "unit": UNIT, "ms_per_step_median": med, "verified": bool(ok)}
            del keys, vals
        del src
    return out


def time_e2e_pipelined(sorter, torch, host_keys, n, steps, warmup):
    """Throughput of a stream of host batches: batch i+1 uploads while batch i sorts and batch i-1
    downloads (two device buffers, three streams, PCIe full duplex).  Every batch still pays its own
    H2D + sort + D2H inside the timed region; only the overlap between consecutive batches is new."""
    pinned_in = torch.from_numpy(host_keys.view("int32")).pin_memory()
    pinned_out = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
    dev = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    storage = sorter.storage_for(n, False)
    up, srt, down = (torch.cuda.Stream() for _ in range(3))
    downloaded = [None, None]

    def batch(i):
        b = i & 1
        with torch.cuda.stream(up):
            if downloaded[b] is not None:
                up.wait_event(downloaded[b])          # buffer b is free once its previous result left
            dev[b].copy_(pinned_in, non_blocking=True)
            uploaded = up.record_event()
        srt.wait_event(uploaded)
        sorter.sort(dev[b], storage=storage, stream=srt)
        sorted_ev = srt.record_event()
        with torch.cuda.stream(down):
            down.wait_event(sorted_ev)
            pinned_out[b].copy_(dev[b], non_blocking=True)
            downloaded[b] = down.record_event()

    for i in range(warmup):
        batch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in (up, srt, down):
        s_.wait_event(e0)
    for i in range(steps):
        batch(warmup + i)
    cur = torch.cuda.current_stream()
    for s_ in (up, srt, down):
        cur.wait_stream(s_)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, pinned_out


def run_single(args):
    import numpy as np
    import torch
    from vulkan_radix_sort_b200 import Sorter, api
    from oracle import cpu_oracle

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(0)
    steps, warmup = args.steps, max(3, args.warmup)
    log2n = args.log2n
    n = 1 << log2n
    host_keys = gen_keys(n)
    sorter = Sorter(0)
    pristine = torch.from_numpy(host_keys.view(np.int32)).cuda()
    peak, peak_src = measured_peak_gbs()

    sampler = ClockSampler(0)   # runs across warm-up, the timed steps and the e2e steps
    sampler.start()
    torch.cuda.synchronize()
    ms, pass_ms, launches, work, _ = time_resident(sorter, torch, api, pristine, n, False, steps, warmup)
    torch.cuda.synchronize()
    ms_per_step = sum(ms) / len(ms)
    value = n / (ms_per_step * 1e-3) / 1e9

    # parity property check of what was just timed (checker only; never on the timed path)
    out = work.cpu().numpy().view(np.uint32)
    assert cpu_oracle.is_sorted(out), "bench output is not sorted"
    assert cpu_oracle.multiset_fingerprint(out) == cpu_oracle.multiset_fingerprint(host_keys), "bench output lost keys"
    del work

    # end to end with host buffers
    e2e_ms, pinned_out = time_e2e(sorter, torch, host_keys, n, max(3, min(steps, 10)), 3)
    e2e_ms_per_step = sum(e2e_ms) / len(e2e_ms)
    e2e_value = n / (e2e_ms_per_step * 1e-3) / 1e9
    assert cpu_oracle.is_sorted(pinned_out.numpy().view(np.uint32))
    del pinned_out
    pipe_ms, pipe_out = time_e2e_pipelined(sorter, torch, host_keys, n, max(4, min(steps, 10)), 3)
    clocks = sampler.stop()
    assert all(cpu_oracle.is_sorted(o.numpy().view(np.uint32)) for o in pipe_out)
    del pipe_out

    # roofline of the dominant kernel (the scatter pass of one LSD pass): algorithmic 8 B/key per launch.
    # traffic = DRAM bytes measured live by ncu on this box, for that kernel (per launch) AND for the whole
    # sort (reduce-then-scan re-reads the keys once per pass: ~48 B/key moved vs 36 algorithmic).
    pass_gbs = PASS_BYTES_PER_KEY_KEYS * n / (pass_ms * 1e-3) / 1e9
    traffic = None if args.no_ncu else measure_dram_traffic_with_ncu(log2n, False)
    pass_traffic, traffic_src = ncu_traffic_per_launch("keys_pass", n), "committed capture (profiles/ncu_traffic.json)"
    whole_traffic = None
    if traffic:
        pk = [v for k, v in traffic["kernels"].items() if "PassKernel" in k]
        if pk:
            pass_traffic, traffic_src = sum(pk) / 4.0, "measured in this run (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)"
        whole_traffic = traffic["whole_sort"]
    roofline = {"bound": "hbm", "achieved": pass_gbs, "peak": peak, "unit": "GB/s", "frac": pass_gbs / peak,
                "traffic": pass_traffic, "traffic_source": traffic_src,
                "kernel": "PassKernel<.., MODE 1> (scatter pass of one LSD pass, keys-only; mean of the 4 passes)",
                "algorithmic_bytes_per_launch": PASS_BYTES_PER_KEY_KEYS * n, "kernel_ms": pass_ms,
                "peak_source": peak_src,
                "whole_sort": {"bytes_per_key": BYTES_PER_KEY_KEYS,
                               "achieved": BYTES_PER_KEY_KEYS * n / (ms_per_step * 1e-3) / 1e9,
                               "frac": BYTES_PER_KEY_KEYS * n / (ms_per_step * 1e-3) / 1e9 / peak,
                               "traffic": whole_traffic,
                               "traffic_over_algorithmic": (whole_traffic / (BYTES_PER_KEY_KEYS * n)) if whole_traffic else None,
                               "traffic_by_kernel": traffic["kernels"] if traffic else None}}

    # extras: key-value at the same N, and both kinds at 2^25 (parity-test sizes, reported for the record)
    extra = {"e2e_pipelined_batches": {
        "value": n / (pipe_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": pipe_ms,
        "note": "NOT the headline e2e: throughput of back-to-back host batches, upload of batch i+1 and download "
                "of batch i-1 overlapped with the sort of batch i (2 device buffers, 3 streams); each batch "
                "still moves 4N bytes each way inside the timed region"}}
    try:
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        kms, kpass, _, kwork, vwork = time_resident(sorter, torch, api, pristine, n, True, max(3, steps // 2), 3, vals)
        kv_ms = sum(kms) / len(kms)
        ok = cpu_oracle.check_stable_permutation(host_keys, kwork.cpu().numpy().view(np.uint32),
                                                 vwork.cpu().numpy().view(np.uint32))
        assert ok, "key-value bench output is not the stable sort"
        extra["key_value_2^%d" % log2n] = {
            "value": n / (kv_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": kv_ms,
            "roofline_frac_whole_sort": BYTES_PER_KEY_KV * n / (kv_ms * 1e-3) / 1e9 / peak,
            "pass_kernel_ms": kpass, "pass_kernel_frac": PASS_BYTES_PER_KEY_KV * n / (kpass * 1e-3) / 1e9 / peak}
        del vals, kwork, vwork
        if log2n > 25:
            n25 = 1 << 25
            p25 = pristine[:n25].clone()
            v25 = torch.arange(n25, dtype=torch.int32, device="cuda")
            for kv in (False, True):
                m, _, _, _, _ = time_resident(sorter, torch, api, p25, n25, kv, 10, 3, v25)
                med = statistics.median(m)
                extra[("key_value" if kv else "keys_only") + "_2^25"] = {
                    "value": n25 / (med * 1e-3) / 1e9, "unit": UNIT, "ms_per_step_median": med,
                    "note": "3.4e7-key working set fits in the 126 MB L2 in part; restore copy between steps"}
    except torch.cuda.OutOfMemoryError as e:  # pragma: no cover
        extra["error"] = str(e)

    # BASELINE.json configs[3]: indirect sort (device-resident count), adversarial keys, non-power-of-two N
    try:
        extra["config4_indirect_adversarial"] = bench_config4(sorter, torch, api, cpu_oracle)
    except Exception as e:  # pragma: no cover
        extra["config4_error"] = repr(e)
    # key-value end to end (host buffers, copies inside the timed region)
    try:
        host_vals = gen_keys(n, seed=2)
        kv_e2e_ms = time_e2e_kv(sorter, torch, host_keys, host_vals, n, 3, 2)
        extra["key_value_e2e_2^%d" % log2n] = {"value": n / (kv_e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": kv_e2e_ms,
                                               "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n}
        del host_vals
    except Exception as e:  # pragma: no cover
        extra["key_value_e2e_error"] = repr(e)
    # the reference's own bench protocol with the b200 backend (bench/benchmark_factory.cc:14-25): one point
    try:
        exe = os.path.join(ROOT, "bench_cpp", "bench")
        if os.path.exists(exe):
            torch.cuda.empty_cache()
            out = subprocess.run([exe, "b200", "--sizes", "2^25", "--seed", "1", "--runs", "5", "-o", "/tmp/vrdx_b200.csv"],
                                 capture_output=True, text=True, timeout=240)
            passed = "Correctness check passed" in out.stdout
            for ln in open("/tmp/vrdx_b200.csv"):
                f = ln.strip().split(",")
                if len(f) == 7 and f[0] == "b200":
                    extra["bench_cpp_b200_%s_2^25" % ("keys_only" if f[2] == "keys" else "key_value")] = {
                        "value": float(f[5]), "unit": UNIT, "gpu_ms": float(f[3]), "correctness_check_passed": passed,
                        "note": "bench_cpp/bench b200 (the reference's CLI, protocol and CSV schema; fresh data per run, median of 5)"}
    except Exception as e:  # pragma: no cover
        extra["bench_cpp_error"] = repr(e)
    extra["vulkan_probe"] = probe_vulkan()

    # comparison point named by BASELINE.json: CUB Onesweep on the same box, through the reference's
    # bench protocol (bench_cpp/bench cuda).  Comparison only — CUB is not in libvrdx_b200.so.
    try:
        exe = os.path.join(ROOT, "bench_cpp", "bench")
        if os.path.exists(exe) and log2n >= 25 and not args.no_cub:
            torch.cuda.empty_cache()
            out = subprocess.run([exe, "cuda", "--sizes", f"2^{log2n}", "--seed", "1", "--runs", "3", "--no-verify",
                                  "-o", "/tmp/vrdx_cub.csv"], capture_output=True, text=True, timeout=240)
            for ln in open("/tmp/vrdx_cub.csv"):
                f = ln.strip().split(",")
                if len(f) == 7 and f[0] == "cuda":
                    extra["cub_onesweep_%s_2^%d" % ("keys_only" if f[2] == "keys" else "key_value", log2n)] = {
                        "value": float(f[5]), "unit": UNIT, "gpu_ms": float(f[3]),
                        "note": "cub::DeviceRadixSort via bench_cpp (fresh data per run, median of 3)"}
    except Exception as e:  # pragma: no cover
        extra["cub_error"] = str(e)

    # CPU baseline: the reference's own CpuBenchmark::Sort on a bounded sample (rank 0, N=1 only)
    cpu = None
    try:
        sample_log2 = min(25, log2n)
        sample = host_keys[: 1 << sample_log2]
        if cpu_oracle.have_ref():
            _, ns = cpu_oracle.ref_sort_keys(sample)
            cpu = {"value": sample.size / (ns * 1e-9) / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"first 2^{sample_log2} keys of the workload, 1 run of CpuBenchmark::Sort (std::sort, single thread)"}
        else:
            t0 = time.perf_counter(); cpu_oracle.sort_keys(sample); dt = time.perf_counter() - t0
            cpu = {"value": sample.size / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first 2^{sample_log2} keys of the workload, 1 run of the C LSD restatement"}
    except Exception as e:  # pragma: no cover
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "ms_per_step_median": statistics.median(ms), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"32-bit keys-only, N=2^{log2n} uniform random (DataGenerator seed 1), vrdxCmdSort direct",
                   "l2": "inputs (1 GiB) larger than L2; a restore copy of the unsorted keys runs between timed steps",
                   "timing": "CUDA events on the launching stream around each sort; mean of K steps",
                   "algorithm": "AUTO: reduce-then-scan at this N (per pass: chunked upsweep, 2 spine kernels, "
                                "look-back-free scatter), 8-bit digits x 4 passes; onesweep below the measured crossover"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step,
                "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n,
                "path": "pinned host keys -> cudaMemcpyAsync H2D -> vrdxCmdSort (C-ABI) -> cudaMemcpyAsync D2H"},
        "gpu_launches": launches * steps,
        "gpu_launches_per_step": launches,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    sorter.close()
    return 0


# ------------------------------------------------------------------------------ our arm, N > 1

def run_distributed(args):
    import bench_dist
    return bench_dist.run(args, METRIC, UNIT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="developer override of the N=1 workload size")
    ap.add_argument("--log2n-reference", type=int, default=None,
                    help="developer override of the --impl reference size (default: the N=1 workload, 2^28)")
    ap.add_argument("--log2n-per-gpu", type=int, default=29, help="developer override of the N>1 per-GPU size")
    ap.add_argument("--log2n-adversarial", type=int, default=24,
                    help="N>1: keys per GPU of the untimed adversarial parity sorts (all_zero, skewed, bits4, all_ones)")
    ap.add_argument("--no-cub", action="store_true", help="skip the CUB comparison run")
    ap.add_argument("--no-ncu", action="store_true", help="skip the live ncu DRAM-traffic measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        return run_distributed(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
