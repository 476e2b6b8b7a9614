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
