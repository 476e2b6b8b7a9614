"""CPU test of bench_cpp/bench: the reference's benchmark CLI, stdout format and CSV schema
(bench/bench.cc:116-207) with the cpu backend, which needs no GPU."""
import csv
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bench_cpp", "bench")


@pytest.fixture(scope="module")
def bench_exe():
    if not os.path.exists(EXE):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "bench_cpp")], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("bench_cpp does not build here: " + r.stderr[-200:])
    return EXE


def test_cli_help_lists_the_backends(bench_exe):
    out = subprocess.run([bench_exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for word in ("b200", "cuda", "cpu", "--no-verify", "-o"):
        assert word in out.stdout
    assert subprocess.run([bench_exe], capture_output=True).returncode == 1          # type is mandatory (bench.cc:140-143)
    assert subprocess.run([bench_exe, "vulkan"], capture_output=True).returncode == 1  # needs a Vulkan ICD


def test_cpu_backend_writes_the_reference_csv_schema(bench_exe, tmp_path):
    path = tmp_path / "cpu.csv"
    out = subprocess.run([bench_exe, "cpu", "--sizes", "2^12,5000", "--seed", "1", "--runs", "2", "-o", str(path)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Correctness check passed" in out.stdout                      # bench.cc:62
    assert "[  1/2] N=     4096 [keys]" in out.stdout                   # bench.cc:172-177 line format
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["backend", "n", "sort", "gpu_ms", "cpu_ms", "gpu_gitems_s", "cpu_gitems_s"]   # bench.cc:199
    body = rows[1:]
    assert [(r[0], r[1], r[2]) for r in body] == [("cpu", "4096", "keys"), ("cpu", "4096", "kv"),
                                                   ("cpu", "5000", "keys"), ("cpu", "5000", "kv")]
    assert all(float(r[3]) > 0 and float(r[5]) > 0 for r in body)


@pytest.mark.gpu
@pytest.mark.parametrize("backend", ["b200", "cuda"])
def test_gpu_backends_pass_the_reference_correctness_gate(bench_exe, tmp_path, backend):
    """`bench b200` (and the CUB comparison backend) through the reference's own protocol —
    the single correctness gate of bench/bench.cc:41-64,164-166 ("GPU == CpuBenchmark at N = 2^18"), the
    factory line bench/benchmark_factory.cc:14-25, the stdout chain and the CSV schema, incl. the
    up / sp / dn split VulkanBenchmark derives from the 15 timestamps (bench/vulkan_benchmark.cc:330-337)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    path = tmp_path / f"{backend}.csv"
    out = subprocess.run([bench_exe, backend, "--sizes", "2^18,262143,2^22", "--seed", "1", "--runs", "3", "-o", str(path)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Correctness check passed" in out.stdout
    lines = open(path).read().splitlines()
    assert lines[0].startswith("# version:")                               # bench.cc:193-197 (tools/plot.py skips '#')
    rows = list(csv.reader(ln for ln in lines if not ln.startswith("#")))
    assert rows[0] == ["backend", "n", "sort", "gpu_ms", "cpu_ms", "gpu_gitems_s", "cpu_gitems_s"]
    body = rows[1:]
    assert [(r[0], r[1], r[2]) for r in body] == [(backend, n, k) for n in ("262144", "262143", "4194304") for k in ("keys", "kv")]
    assert all(float(r[3]) > 0 and float(r[4]) > 0 and float(r[5]) > 0 for r in body)
    if backend == "b200":
        # the per-stage split only exists for backends with timestamps; large N runs reduce-then-scan,
        # where upsweep, spine and downsweep are separate kernels and all three columns are non-zero
        big = subprocess.run([bench_exe, "b200", "--sizes", "2^26", "--seed", "1", "--runs", "2", "--no-verify",
                              "-o", str(tmp_path / "big.csv")], capture_output=True, text=True, timeout=300)
        assert big.returncode == 0, big.stdout + big.stderr
        line = [ln for ln in big.stdout.splitlines() if "[keys]" in ln][0]
        import re
        m = re.search(r"\[up=([0-9.]+)ms\(\d+%\) sp=([0-9.]+)ms\(\d+%\) dn=([0-9.]+)ms\(\d+%\)\]", line)   # bench.cc:178-186
        assert m, line
        assert all(float(x) > 0 for x in m.groups()), line
