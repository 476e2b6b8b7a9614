// bench_cpp/bench_main.cu — the reference's benchmark protocol with a `b200` backend.
//
// Mirrors bench/bench.cc of jaesung-cs/vulkan_radix_sort (CLI :116-129, sweep :15-20,161-162,
// 1 warm-up + 10 timed runs with fresh data and the MEDIAN :66-112, correctness check against the
// CPU backend :41-64, stdout format :172-186, CSV schema :191-203) so that tools/plot.py of the
// reference reads our CSVs unchanged.  Backends (bench/benchmark_factory.cc:14-25):
//   b200   this repo's libvrdx_b200.so through the vrdx* C API, driven the way VulkanBenchmark
//          drives it (bench/vulkan_benchmark.cc:253-433): Sort -> vrdxCmdSort (direct),
//          SortKeyValue -> vrdxCmdSortKeyValueIndirect with keys, values and the count in ONE
//          allocation at offsets 0, inout, 2*inout; times from the 15-slot query pool.
//   cuda   cub::DeviceRadixSort::SortKeys / SortPairs, bits [0,32) — the comparison point the
//          reference uses (bench/cuda_benchmark.cu:48-63, 94-111).  Comparison only: CUB is linked
//          into THIS binary, never into libvrdx_b200.so.
//   cpu    std::sort / std::stable_sort on indices + gather (bench/cpu_benchmark.cc:19-53).
// Inputs and verifier: where the reference tree is present at build time (this container), the binary
// is built against the reference's OWN bench/data_generator.cc and bench/cpu_benchmark.cc, compiled
// where they lie (north_star: the b200 backend "reuses data_generator and the existing verifier");
// without it (make REFERENCE=) the in-file mirrors below are used.  `bench --help` says which.
// cxxopts (a network fetch in the reference) is replaced by hand parsing.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <cub/version.cuh>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#define VRDX_FORCE_VK_SHIM 1
#include "vk_radix_sort.h"
#include "vrdx_cuda.h"

#ifdef VRDX_BENCH_REFERENCE_SOURCES
#include "benchmark_base.h"   // /root/reference/bench (read where it lies, never copied)
#include "cpu_benchmark.h"
#include "data_generator.h"
using Results = BenchmarkBase::Results;
static const char* kInputsAndVerifier = "reference bench/data_generator.cc + bench/cpu_benchmark.cc (compiled unmodified)";
#else
static const char* kInputsAndVerifier = "in-file mirrors of bench/data_generator.cc and bench/cpu_benchmark.cc";
#endif

namespace {

uint64_t NowNs() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#ifndef VRDX_BENCH_REFERENCE_SOURCES
// ---------------------------------------------------------------------------- inputs
// Same stream as the reference's DataGenerator (bench/data_generator.cc:12-27).
struct SortData {
  std::vector<uint32_t> keys, values;
};
class DataGenerator {
 public:
  DataGenerator() { std::random_device rd; gen_ = std::mt19937(rd()); }
  explicit DataGenerator(int seed) : gen_(seed) {}
  SortData Generate(uint32_t size, uint32_t bits = 32) {
    std::uniform_int_distribution<uint32_t> dist_values, dist_keys;
    if (bits < 32) dist_keys = std::uniform_int_distribution<uint32_t>(0, (1u << bits) - 1);
    SortData d;
    d.keys.resize(size);
    d.values.resize(size);
    for (auto& k : d.keys) k = dist_keys(gen_);
    for (auto& v : d.values) v = dist_values(gen_);
    return d;
  }
 private:
  std::mt19937 gen_;
};

// ---------------------------------------------------------------------------- backend interface
struct Results {
  std::vector<uint32_t> keys, values;
  uint64_t total_time = 0, cpu_time = 0;  // ns
  uint64_t upsweep_ns = 0, spine_ns = 0, downsweep_ns = 0;
};
class BenchmarkBase {
 public:
  virtual ~BenchmarkBase() = default;
  virtual std::string LibraryVersion() const { return ""; }
  virtual Results Sort(const std::vector<uint32_t>& keys) = 0;
  virtual Results SortKeyValue(const std::vector<uint32_t>& keys, const std::vector<uint32_t>& values) = 0;
};

class CpuBenchmark : public BenchmarkBase {
 public:
  Results Sort(const std::vector<uint32_t>& keys) override {
    Results r;
    r.keys = keys;
    const uint64_t t0 = NowNs();
    std::sort(r.keys.begin(), r.keys.end());
    r.total_time = r.cpu_time = NowNs() - t0;
    return r;
  }
  Results SortKeyValue(const std::vector<uint32_t>& keys, const std::vector<uint32_t>& values) override {
    std::vector<uint32_t> idx(keys.size());
    std::iota(idx.begin(), idx.end(), 0u);
    const uint64_t t0 = NowNs();
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    const uint64_t t1 = NowNs();
    Results r;
    r.keys.resize(keys.size());
    r.values.resize(keys.size());
    for (size_t i = 0; i < idx.size(); ++i) {
      r.keys[i] = keys[idx[i]];
      r.values[i] = values[idx[i]];
    }
    r.total_time = r.cpu_time = t1 - t0;
    return r;
  }
};

#endif  // !VRDX_BENCH_REFERENCE_SOURCES

#define CK(x)                                                                               \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      std::exit(2);                                                                         \
    }                                                                                       \
  } while (0)

struct DeviceBuffer {
  void* p = nullptr;
  size_t size = 0;
  void Reserve(size_t bytes) {
    if (bytes <= size) return;
    if (p) CK(cudaFree(p));
    CK(cudaMalloc(&p, bytes));
    size = bytes;
  }
  ~DeviceBuffer() { if (p) cudaFree(p); }
};

class CubBenchmark : public BenchmarkBase {
 public:
  CubBenchmark() { CK(cudaStreamCreate(&stream_)); CK(cudaEventCreate(&e0_)); CK(cudaEventCreate(&e1_)); }
  ~CubBenchmark() override { cudaStreamDestroy(stream_); cudaEventDestroy(e0_); cudaEventDestroy(e1_); }
  std::string LibraryVersion() const override {
    return "v" + std::to_string(CUB_MAJOR_VERSION) + "." + std::to_string(CUB_MINOR_VERSION) + "." +
           std::to_string(CUB_SUBMINOR_VERSION);
  }
  Results Sort(const std::vector<uint32_t>& keys) override { return Run(keys, nullptr); }
  Results SortKeyValue(const std::vector<uint32_t>& keys, const std::vector<uint32_t>& values) override {
    return Run(keys, &values);
  }
 private:
  Results Run(const std::vector<uint32_t>& keys, const std::vector<uint32_t>* values) {
    const size_t n = keys.size(), bytes = n * sizeof(uint32_t);
    kin_.Reserve(bytes); kout_.Reserve(bytes);
    if (values) { vin_.Reserve(bytes); vout_.Reserve(bytes); }
    CK(cudaMemcpy(kin_.p, keys.data(), bytes, cudaMemcpyHostToDevice));
    if (values) CK(cudaMemcpy(vin_.p, values->data(), bytes, cudaMemcpyHostToDevice));
    auto* ki = static_cast<const uint32_t*>(kin_.p);
    auto* ko = static_cast<uint32_t*>(kout_.p);
    auto* vi = static_cast<const uint32_t*>(vin_.p);
    auto* vo = static_cast<uint32_t*>(vout_.p);
    size_t temp_bytes = 0;
    if (values) cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, ki, ko, vi, vo, n);
    else cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, ki, ko, n);
    temp_.Reserve(temp_bytes ? temp_bytes : 16);
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c0 = NowNs();
    CK(cudaEventRecord(e0_, stream_));
    if (values) cub::DeviceRadixSort::SortPairs(temp_.p, temp_bytes, ki, ko, vi, vo, n, 0, 32, stream_);
    else cub::DeviceRadixSort::SortKeys(temp_.p, temp_bytes, ki, ko, n, 0, 32, stream_);
    CK(cudaEventRecord(e1_, stream_));
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c1 = NowNs();
    Results r;
    r.keys.resize(n);
    CK(cudaMemcpy(r.keys.data(), ko, bytes, cudaMemcpyDeviceToHost));
    if (values) { r.values.resize(n); CK(cudaMemcpy(r.values.data(), vo, bytes, cudaMemcpyDeviceToHost)); }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0_, e1_));
    r.total_time = (uint64_t)((double)ms * 1e6);
    r.cpu_time = c1 - c0;
    return r;
  }
  cudaStream_t stream_{};
  cudaEvent_t e0_{}, e1_{};
  DeviceBuffer kin_, kout_, vin_, vout_, temp_;
};

// The b200 backend: the vrdx* C API exactly as VulkanBenchmark uses it.
class B200Benchmark : public BenchmarkBase {
 public:
  B200Benchmark() {
    CK(cudaStreamCreate(&stream_));
    VrdxSorterCreateInfo info = {vrdxCudaPhysicalDevice(0), vrdxCudaDevice(0), VK_NULL_HANDLE};
    if (vrdxCreateSorter(&info, &sorter_) != VK_SUCCESS) { std::fprintf(stderr, "vrdxCreateSorter failed\n"); std::exit(2); }
    if (vrdxCudaCreateQueryPool(vrdxCudaDevice(0), 15, &pool_) != VK_SUCCESS) std::exit(2);
    // VRDX_BENCH_NO_POOL=1: time with two CUDA events around the call (as the cuda backend is
    // timed) instead of the 15-slot query pool, to expose the cost of the intermediate timestamps.
    if (const char* e = std::getenv("VRDX_BENCH_NO_POOL")) no_pool_ = std::atoi(e) != 0;
    CK(cudaEventCreate(&e0_));
    CK(cudaEventCreate(&e1_));
  }
  ~B200Benchmark() override {
    cudaStreamSynchronize(stream_);
    vrdxCudaDestroyQueryPool(pool_);
    vrdxDestroySorter(sorter_);
    cudaStreamDestroy(stream_);
  }
  std::string LibraryVersion() const override {
    return "v" + std::to_string(VRDX_VERSION_MAJOR) + "." + std::to_string(VRDX_VERSION_MINOR) + "." +
           std::to_string(VRDX_VERSION_PATCH) + "-b200";
  }
  Results Sort(const std::vector<uint32_t>& keys) override {
    const uint32_t n = (uint32_t)keys.size();
    const size_t inout = Align16((size_t)n * 4);
    keys_.Reserve(inout);
    VrdxSorterStorageRequirements req{};
    vrdxGetSorterStorageRequirements(sorter_, n, &req);
    storage_.Reserve(req.size);
    CK(cudaMemcpyAsync(keys_.p, keys.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c0 = NowNs();
    if (no_pool_) CK(cudaEventRecord(e0_, stream_));
    vrdxCmdSort(vrdxCudaCommandBuffer(stream_), sorter_, n, vrdxCudaBuffer(keys_.p), 0, vrdxCudaBuffer(storage_.p), 0,
                no_pool_ ? VK_NULL_HANDLE : pool_, 0);
    if (no_pool_) CK(cudaEventRecord(e1_, stream_));
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c1 = NowNs();
    Results r;
    r.keys.resize(n);
    CK(cudaMemcpy(r.keys.data(), keys_.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    Fill(r, c1 - c0);
    return r;
  }
  Results SortKeyValue(const std::vector<uint32_t>& keys, const std::vector<uint32_t>& values) override {
    const uint32_t n = (uint32_t)keys.size();
    const size_t inout = Align16((size_t)n * 4);
    keys_.Reserve(2 * inout + 16);  // [keys | values | count] in one buffer (vulkan_benchmark.cc:356-358)
    VrdxSorterStorageRequirements req{};
    vrdxGetSorterKeyValueStorageRequirements(sorter_, n, &req);
    storage_.Reserve(req.size);
    char* base = static_cast<char*>(keys_.p);
    CK(cudaMemcpyAsync(base, keys.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(base + inout, values.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(base + 2 * inout, &n, 4, cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c0 = NowNs();
    if (no_pool_) CK(cudaEventRecord(e0_, stream_));
    vrdxCmdSortKeyValueIndirect(vrdxCudaCommandBuffer(stream_), sorter_, n, vrdxCudaBuffer(base), 2 * inout,
                                vrdxCudaBuffer(base), 0, vrdxCudaBuffer(base), inout, vrdxCudaBuffer(storage_.p), 0,
                                no_pool_ ? VK_NULL_HANDLE : pool_, 0);
    if (no_pool_) CK(cudaEventRecord(e1_, stream_));
    CK(cudaStreamSynchronize(stream_));
    const uint64_t c1 = NowNs();
    Results r;
    r.keys.resize(n);
    r.values.resize(n);
    CK(cudaMemcpy(r.keys.data(), base, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.values.data(), base + inout, (size_t)n * 4, cudaMemcpyDeviceToHost));
    Fill(r, c1 - c0);
    return r;
  }
 private:
  static size_t Align16(size_t x) { return (x + 15) / 16 * 16; }
  void Fill(Results& r, uint64_t cpu_ns) {
    if (int e = vrdxCudaGetLastError(sorter_)) {
      std::fprintf(stderr, "vrdx error: %s\n", vrdxCudaGetErrorString(e));
      std::exit(2);
    }
    r.cpu_time = cpu_ns;
    if (no_pool_) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0_, e1_));
      r.total_time = (uint64_t)((double)ms * 1e6);
      return;
    }
    uint64_t ts[15] = {};
    vrdxCudaGetQueryPoolResults(pool_, 0, 15, ts);
    r.total_time = ts[14] - ts[0];
    for (int p = 0; p < 4; ++p) {  // same slot arithmetic as vulkan_benchmark.cc:330-337
      r.upsweep_ns += ts[2 + 3 * p] - ts[1 + 3 * p];
      r.spine_ns += ts[3 + 3 * p] - ts[2 + 3 * p];
      r.downsweep_ns += ts[4 + 3 * p] - ts[3 + 3 * p];
    }
  }
  cudaStream_t stream_{};
  VrdxSorter sorter_{};
  VkQueryPool pool_{};
  bool no_pool_ = false;
  cudaEvent_t e0_{}, e1_{};
  DeviceBuffer keys_, storage_;
};

std::unique_ptr<BenchmarkBase> Create(const std::string& type) {
  if (type == "cpu") return std::make_unique<CpuBenchmark>();
  if (type == "b200") return std::make_unique<B200Benchmark>();
  if (type == "cuda") return std::make_unique<CubBenchmark>();
  throw std::runtime_error("Unavailable benchmark type: " + type + " (vulkan / fuchsia need a Vulkan ICD)");
}

// ---------------------------------------------------------------------------- protocol
constexpr int kWarmupRuns = 1;
constexpr int kTimedRuns = 10;

double ToMs(uint64_t ns) { return (double)ns / 1e6; }
double ToGItemsS(uint32_t n, uint64_t ns) { return ns ? ((double)n / 1e9) / ((double)ns / 1e9) : 0.0; }
uint64_t Median(std::vector<uint64_t>& v) {
  auto mid = v.begin() + v.size() / 2;
  std::nth_element(v.begin(), mid, v.end());
  return *mid;
}

struct Row {
  uint32_t n;
  std::string sort;
  double gpu_ms, cpu_ms, gpu_gitems_s, cpu_gitems_s, upsweep_ms, spine_ms, downsweep_ms;
};

bool CheckCorrectness(BenchmarkBase* bench, BenchmarkBase* cpu, uint32_t n, DataGenerator& gen) {
  SortData data = gen.Generate(n);
  Results r0 = bench->Sort(data.keys), r1 = cpu->Sort(data.keys);
  for (uint32_t i = 0; i < n; ++i)
    if (r0.keys[i] != r1.keys[i]) { std::cerr << "Sort correctness failed at index " << i << std::endl; return false; }
  Results r2 = bench->SortKeyValue(data.keys, data.values), r3 = cpu->SortKeyValue(data.keys, data.values);
  for (uint32_t i = 0; i < n; ++i)
    if (r2.keys[i] != r3.keys[i] || r2.values[i] != r3.values[i]) {
      std::cerr << "SortKeyValue correctness failed at index " << i << std::endl;
      return false;
    }
  std::cout << "Correctness check passed (N=" << n << ")" << std::endl;
  return true;
}

Row Measure(BenchmarkBase* bench, uint32_t n, const std::string& sort, DataGenerator& gen, int timed_runs) {
  auto run = [&](SortData& d) { return sort == "keys" ? bench->Sort(d.keys) : bench->SortKeyValue(d.keys, d.values); };
  for (int i = 0; i < kWarmupRuns; ++i) { SortData d = gen.Generate(n); run(d); }
  std::vector<uint64_t> g, c, up, sp, dn;
  for (int i = 0; i < timed_runs; ++i) {
    SortData d = gen.Generate(n);
    Results r = run(d);
    g.push_back(r.total_time); c.push_back(r.cpu_time);
    up.push_back(r.upsweep_ns); sp.push_back(r.spine_ns); dn.push_back(r.downsweep_ns);
  }
  const uint64_t gm = Median(g), cm = Median(c);
  return Row{n, sort, ToMs(gm), ToMs(cm), ToGItemsS(n, gm), ToGItemsS(n, cm), ToMs(Median(up)), ToMs(Median(sp)), ToMs(Median(dn))};
}

void Usage() {
  std::cout << "Usage: bench <type> [-o results.csv] [--no-verify] [--sizes a,b,...] [--seed s] [--runs k]\n\n"
               "Types:\n  b200      B200-native CUDA backend of the vrdx API (this repo)\n"
               "  cuda      CUB Onesweep (CUDA)\n  cpu       std::sort reference\n\n"
               "Without --sizes the reference sweep is run: N = 2^18 .. 2^25 in 128 linear steps.\n"
               "Inputs and verifier: " << kInputsAndVerifier << "\n";
}

}  // namespace

int main(int argc, char** argv) {
  std::string type, csv_path = "results.csv", sizes;
  bool no_verify = false;
  int seed = -1, runs = kTimedRuns;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-h" || a == "--help") { Usage(); return 0; }
    else if ((a == "-o" || a == "--output") && i + 1 < argc) csv_path = argv[++i];
    else if (a == "--no-verify") no_verify = true;
    else if (a == "--validation") {}  // Vulkan validation layers: meaningless here, accepted for CLI parity
    else if (a == "--sizes" && i + 1 < argc) sizes = argv[++i];
    else if (a == "--seed" && i + 1 < argc) seed = std::atoi(argv[++i]);
    else if (a == "--runs" && i + 1 < argc) runs = std::atoi(argv[++i]);
    else if (type.empty() && a[0] != '-') type = a;
    else { std::cerr << "unknown argument " << a << "\n"; Usage(); return 1; }
  }
  if (type.empty()) { Usage(); return 1; }
  std::unique_ptr<BenchmarkBase> bench, cpu;
  try {
    bench = Create(type);
    cpu = Create("cpu");
  } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }

  std::vector<uint32_t> ns;
  if (sizes.empty()) {
    constexpr uint32_t kNMin = 1u << 18, kNMax = 1u << 25;
    constexpr int kNCount = 128;
    constexpr uint32_t kNStep = (kNMax - kNMin) / (kNCount - 1);
    for (int i = 0; i < kNCount; ++i) ns.push_back(kNMin + (uint32_t)i * kNStep);
  } else {
    std::stringstream ss(sizes);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
      if (tok.rfind("2^", 0) == 0) ns.push_back(1u << std::atoi(tok.c_str() + 2));
      else ns.push_back((uint32_t)std::strtoul(tok.c_str(), nullptr, 10));
    }
  }
  DataGenerator gen = seed >= 0 ? DataGenerator(seed) : DataGenerator();
  std::vector<Row> rows;
  for (size_t i = 0; i < ns.size(); ++i) {
    const uint32_t n = ns[i];
    if (i == 0 && !no_verify) {
      const uint32_t vn = std::min<uint32_t>(n, 1u << 18);  // the reference verifies at its first sweep point, 2^18
      if (!CheckCorrectness(bench.get(), cpu.get(), vn, gen)) return 1;
    }
    for (const std::string sort : {"keys", "kv"}) {
      Row row = Measure(bench.get(), n, sort, gen, runs);
      rows.push_back(row);
      std::cout << "[" << std::setw(3) << i + 1 << "/" << ns.size() << "]" << " N=" << std::setw(9) << n << " ["
                << std::setw(4) << sort << "]" << "  gpu: " << std::fixed << std::setprecision(3) << row.gpu_ms << "ms"
                << " (" << std::setprecision(2) << row.gpu_gitems_s << " GItems/s)" << "  cpu: " << std::setprecision(3)
                << row.cpu_ms << "ms" << " (" << std::setprecision(2) << row.cpu_gitems_s << " GItems/s)";
      if (row.upsweep_ms > 0 || row.spine_ms > 0 || row.downsweep_ms > 0) {
        auto pct = [&](double ms) { return row.gpu_ms > 0 ? (int)(ms / row.gpu_ms * 100 + 0.5) : 0; };
        std::cout << std::fixed << std::setprecision(3) << "  [up=" << row.upsweep_ms << "ms(" << pct(row.upsweep_ms) << "%)"
                  << " sp=" << row.spine_ms << "ms(" << pct(row.spine_ms) << "%)" << " dn=" << row.downsweep_ms << "ms("
                  << pct(row.downsweep_ms) << "%)" << "]";
      }
      std::cout << std::endl;
    }
  }
  std::ofstream csv(csv_path);
  if (!csv) { std::cerr << "Failed to open " << csv_path << " for writing" << std::endl; return 1; }
  const std::string ver = bench->LibraryVersion();
  if (!ver.empty()) csv << "# version: " << ver << "\n";
  csv << "backend,n,sort,gpu_ms,cpu_ms,gpu_gitems_s,cpu_gitems_s\n";
  for (const Row& r : rows)
    csv << type << "," << r.n << "," << r.sort << "," << std::fixed << std::setprecision(6) << r.gpu_ms << "," << r.cpu_ms
        << "," << r.gpu_gitems_s << "," << r.cpu_gitems_s << "\n";
  std::cout << "\nResults written to " << csv_path << std::endl;
  return 0;
}
