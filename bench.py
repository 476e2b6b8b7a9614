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
    # bounded sample: the whole run must end within minutes; std::sort does ~8.6 MKeys/s/core
    budget_keys = 8.0e6 * 150.0 / max(1, steps + warmup)
    log2s = max(18, min(25, int(np.floor(np.log2(budget_keys)))))
    n = 1 << log2s
    keys = gen_keys(n)
    for _ in range(warmup):
        sort(keys)
    times = [sort(keys) for _ in range(steps)]
    mean_s = sum(times) / len(times)
    value = n / mean_s / 1e9
    sample = f"first 2^{log2s} keys of the 2^28-key workload per step; std::sort, 1 thread (the reference's CpuBenchmark is single-threaded)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "32-bit keys-only, uniform (DataGenerator seed 1), CPU sample of 2^%d keys" % log2s,
                   "host_cores": os.cpu_count()},
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

    # roofline of the dominant kernel (one onesweep pass): algorithmic 8 B/key per launch
    pass_gbs = PASS_BYTES_PER_KEY_KEYS * n / (pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": pass_gbs, "peak": peak, "unit": "GB/s", "frac": pass_gbs / peak,
                "traffic": ncu_traffic_per_launch("keys_pass", n), "kernel": "OnesweepKernel<.., MODE 1> (scatter pass of one LSD pass, keys-only)",
                "algorithmic_bytes_per_launch": PASS_BYTES_PER_KEY_KEYS * n, "kernel_ms": pass_ms,
                "peak_source": peak_src,
                "whole_sort": {"bytes_per_key": BYTES_PER_KEY_KEYS,
                               "achieved": BYTES_PER_KEY_KEYS * n / (ms_per_step * 1e-3) / 1e9,
                               "frac": BYTES_PER_KEY_KEYS * n / (ms_per_step * 1e-3) / 1e9 / peak}}

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
                                "look-back-free scatter), 8-bit digits x 4 passes; onesweep below 3*2^23 keys"},
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
    ap.add_argument("--log2n-per-gpu", type=int, default=29, help="developer override of the N>1 per-GPU size")
    ap.add_argument("--no-cub", action="store_true", help="skip the CUB comparison run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        return run_distributed(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
