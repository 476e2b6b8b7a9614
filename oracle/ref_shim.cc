// oracle/ref_shim.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" doorway into the reference's OWN CPU implementation, compiled unmodified
// from where it lies under /root/reference (nothing is copied into this repo):
//   bench/cpu_benchmark.cc   CpuBenchmark::Sort          = std::sort            (:19-28)
//                            CpuBenchmark::SortKeyValue  = std::stable_sort on indices + gather (:30-53)
//   bench/data_generator.cc  DataGenerator(seed).Generate(size, bits)            (:8-27)
// This is the definition of "correct" the reference applies to its GPU path
// (bench/bench.cc:41-64).  Built by oracle/Makefile into oracle/_ref/libvrdx_ref.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpu_benchmark.h"   // -I/root/reference/bench
#include "data_generator.h"  // -I/root/reference/bench

extern "C" {

// keys <- first `size` draws, values <- next `size` draws of DataGenerator(seed).Generate(size, bits)
int vrdx_ref_generate(int seed, uint32_t size, uint32_t bits, uint32_t* keys, uint32_t* values) {
  DataGenerator gen(seed);
  SortData data = gen.Generate(size, bits);
  if (keys) std::memcpy(keys, data.keys.data(), sizeof(uint32_t) * size);
  if (values) std::memcpy(values, data.values.data(), sizeof(uint32_t) * size);
  return 0;
}

// out_keys <- CpuBenchmark::Sort(keys).keys ; *ns <- its timed region (sort only)
int vrdx_ref_sort_keys(const uint32_t* keys, uint32_t n, uint32_t* out_keys, uint64_t* ns) {
  std::vector<uint32_t> in(keys, keys + n);
  CpuBenchmark cpu;
  auto r = cpu.Sort(in);
  if (out_keys) std::memcpy(out_keys, r.keys.data(), sizeof(uint32_t) * n);
  if (ns) *ns = r.total_time;
  return 0;
}

int vrdx_ref_sort_key_value(const uint32_t* keys, const uint32_t* values, uint32_t n,
                            uint32_t* out_keys, uint32_t* out_values, uint64_t* ns) {
  std::vector<uint32_t> k(keys, keys + n), v(values, values + n);
  CpuBenchmark cpu;
  auto r = cpu.SortKeyValue(k, v);
  if (out_keys) std::memcpy(out_keys, r.keys.data(), sizeof(uint32_t) * n);
  if (out_values) std::memcpy(out_values, r.values.data(), sizeof(uint32_t) * n);
  if (ns) *ns = r.total_time;
  return 0;
}

}  // extern "C"
