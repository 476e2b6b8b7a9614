// vrdx_api.cu — the vk_radix_sort C API on CUDA streams (host side of libvrdx_b200.so).
//
// Replaces the reference's host layer, src/vk_radix_sort.h.in:
//   :141-277  vrdxCreateSorter / vrdxDestroySorter  (pipelines -> kernel attributes)
//   :279-308  vrdxGetSorter[KeyValue]StorageRequirements
//   :310-342  the four vrdxCmdSort* wrappers
//   :344-507  gpuSort(): command recording  -> stream launches below (EnqueueSort)
// "Recording" into a VkCommandBuffer becomes enqueueing on a cudaStream_t: asynchronous, no
// host synchronisation, no host read of device memory, legal inside CUDA-graph capture.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#define VRDX_FORCE_VK_SHIM 1
#include "vk_radix_sort.h"
#include "vrdx_cuda.h"
#include "vrdx_dist.h"
#include "vrdx_dist_kernels.cuh"
#include "vrdx_kernels.cuh"
#include "vrdx_layout.h"

using namespace vrdx;

// ---------------------------------------------------------------------------- configuration
// Tile shapes of the pass kernel: threads x keys-per-thread, and the CTAs/SM the register
// allocation is bounded for.  One CTA sorts one tile per pass.  The product library carries the
// shapes VRDX_CUDA_ALGORITHM_AUTO can pick plus the few alternates the tuning sweeps compare
// (VrdxCudaSorterOptions::reserved[0..1], index + 1); every other variant that was measured and
// lost lives in vrdx_experiments.cuh and is compiled only with -DVRDX_EXPERIMENTS.
struct TileShape {
  int threads, items, min_ctas;
  uint32_t tile;
  size_t smem;
  cudaError_t (*prepare)(int* ctas_per_sm);  // per device: opt every instantiation into its shared-memory size
  // mode 0: onesweep pass, 1: reduce-then-scan scatter pass; generic: digit plan / key codec from PassArgs
  cudaError_t (*launch_pass)(cudaStream_t, uint32_t grid, const PassArgs&, int mode, bool generic, bool pdl);
  cudaError_t (*launch_upsweep)(cudaStream_t, uint32_t grid, const PassArgs&, bool pdl);
};

// Launch with (or without) the programmatic-dependent-launch attribute; see GridDepWait().
template <typename... KArgs, typename... Args>
cudaError_t LaunchEx(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t stream,
                     bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <class Cfg>
cudaError_t PrepareShape(int* ctas_per_sm) {
  cudaError_t e = cudaSuccess;
  auto opt_in = [&](void (*k)(const PassArgs)) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
  };
  opt_in(PassKernel<Cfg, 0, false>);
  opt_in(PassKernel<Cfg, 1, false>);
  opt_in(PassKernel<Cfg, 0, true>);
  opt_in(PassKernel<Cfg, 1, true>);
  if constexpr (!Cfg::kKeyValue) {
    opt_in(PassKernel<Cfg, 1, false, VRDX_RANK, true>);
    opt_in(PassKernel<Cfg, 1, true, VRDX_RANK, true>);
  }
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, PassKernel<Cfg, 1, false>, Cfg::kThreads,
                                                       Cfg::kSmemBytes);
}
// two: the instantiations whose two-run tiles are block-free too (keys-only reduce-then-scan; PassArgs::two_runs)
template <class Cfg>
cudaError_t LaunchPass(cudaStream_t stream, uint32_t grid, const PassArgs& args, int mode, bool generic, bool pdl) {
  void (*k)(const PassArgs) = mode == 0 ? (generic ? PassKernel<Cfg, 0, true> : PassKernel<Cfg, 0, false>)
                                        : (generic ? PassKernel<Cfg, 1, true> : PassKernel<Cfg, 1, false>);
  if constexpr (!Cfg::kKeyValue) {
    if (mode == 1 && args.two_runs)
      k = generic ? PassKernel<Cfg, 1, true, VRDX_RANK, true> : PassKernel<Cfg, 1, false, VRDX_RANK, true>;
  }
  return LaunchEx(k, grid, Cfg::kThreads, Cfg::kSmemBytes, stream, pdl, args);
}
// Bits below the digit of this pass when keys that agree in them may swap places (PassArgs::words_only, see
// MakePassDigit in the kernels), else all 32 bits.
inline uint32_t LowMask(const PassArgs& a) { return a.words_only ? ((1u << a.shift) - 1u) : 0xFFFFFFFFu; }
template <class Cfg>
cudaError_t LaunchUpsweep(cudaStream_t stream, uint32_t grid, const PassArgs& args, bool pdl) {
  // the first kernel of a sort (pass 0) is a normal launch: it must wait for the caller's prior work
  return LaunchEx(args.two_runs ? UpsweepKernel<Cfg::kTile, true, true> : UpsweepKernel<Cfg::kTile, true, false>, grid,
                  kUpsweepThreads, 0, stream, pdl && args.pass != 0, args.indirect, args.n_or_max, args.shift, args.mask,
                  args.codec_in, args.keys_in, args.status, args.status_next, args.hdr, args.ts_end, args.ts_start,
                  args.tile_flags, LowMask(args));
}
template <int T, int I, bool KV, int M>
constexpr TileShape MakeShape() {
  using Cfg = PassConfig<T, I, KV, M>;
  return TileShape{T, I, M, (uint32_t)Cfg::kTile, Cfg::kSmemBytes, &PrepareShape<Cfg>, &LaunchPass<Cfg>,
                   &LaunchUpsweep<Cfg>};
}

// Shapes (measured on B200, profiles/r02/n_shape_sweep.txt: 2^28 keys in 3.37 ms with 256 x 20, 3.64 ms with
// 256 x 16; 2^28 pairs in 5.39 ms with 320 x 20, 5.55 ms with 384 x 16).  Index 0 is the default of both
// compositions; the others are the alternates the tuning sweeps compare (VrdxCudaSorterOptions::reserved[0..1]).
static const TileShape kKeysShapes[] = {
    MakeShape<256, 20, false, 4>(), MakeShape<256, 16, false, 4>(), MakeShape<384, 16, false, 3>(),
    MakeShape<512, 16, false, 2>(),
};
static const TileShape kPairShapes[] = {
    MakeShape<320, 20, true, 3>(), MakeShape<256, 16, true, 4>(), MakeShape<384, 16, true, 3>(),
    MakeShape<512, 16, true, 2>(),
};
constexpr int kDefaultKeysRtsShape = 0;
constexpr int kDefaultPairRtsShape = 0;
constexpr int kDefaultOnesweepShape = 0;  // keys and pairs
constexpr int kNumKeysShapes = sizeof(kKeysShapes) / sizeof(kKeysShapes[0]);
constexpr int kNumPairShapes = sizeof(kPairShapes) / sizeof(kPairShapes[0]);
#ifdef VRDX_EXPERIMENTS
// Smallest tile of any compiled shape: sizes the per-tile tables of the VRDX_EXPERIMENTS build, whichever shape runs.
constexpr uint32_t kMinTile = 4096;  // 256 x 16
#endif
// AUTO: reduce-then-scan at and above this count, onesweep (fewer launches) below it.  Measured crossovers
// (profiles/r02/s_sweep_n_auto_thresholds.txt): keys-only 2^25 (block-free tiles and the 128-bit upsweep moved it down
// from 2^25.6), key-value ~2^27.5 (the two compositions are within 1 % of each other from 2^27 up).
constexpr uint32_t kAutoRtsThresholdKeys = 1u << 25;
constexpr uint32_t kAutoRtsThresholdPairs = 1u << 27;
// the conflict-free histogram kernel needs enough keys to fill one 1024-thread CTA per SM
constexpr uint32_t kHistPrivateMinCount = 1u << 21;

#ifdef VRDX_EXPERIMENTS
#include "vrdx_experiments_host.inc"
#endif

// Immutable after creation (threading row): every per-sort decision reads these fields
// or the arguments of the call; there is no process-wide state in the library.
struct VrdxSorter_T {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  VrdxCudaAlgorithm algorithm = VRDX_CUDA_ALGORITHM_AUTO;
  VrdxCudaTileLoad tile_load = VRDX_CUDA_TILE_LOAD_AUTO;
  int keys_shape = kDefaultOnesweepShape, pair_shape = kDefaultOnesweepShape;        // onesweep
  int keys_rts_shape = kDefaultKeysRtsShape, pair_rts_shape = kDefaultPairRtsShape;  // reduce-then-scan
  int keys_rts_ctas = 1, pair_rts_ctas = 1;               // co-resident CTAs per SM of the scatter kernels
  bool pdl = true;                                        // programmatic dependent launch between our own kernels
  bool relaxed_equal_low_bits = true;                     // PassArgs::words_only (developer switch VRDX_RELAXED=0)
  int two_runs = -1;                                      // PassArgs::two_runs: -1 by run length (default), 0 never,
                                                          // 1 in every pass above the first (developer switch VRDX_TWO_RUNS)
  uint32_t hist_private_min_count = kHistPrivateMinCount;  // lane-private histogram bins from this count up
#ifdef VRDX_EXPERIMENTS
  ExperimentSelection exp;
#endif
  // The only mutable words: a sticky error and a launch counter (diagnostics, not sort state).
  std::atomic<int> last_error{0};
  std::atomic<uint32_t> last_launches{0};
};

// A VkQueryPool on this backend: device memory the sort's own kernels stamp with %globaltimer
// (like vkCmdWriteTimestamp, the write happens on the GPU timeline and costs no stream
// serialisation).  Slots that coincide (no kernel between them) alias an earlier slot.
// Storage layout of a sorter: 32-bit look-back rows only for the counts it may sort with onesweep.
static StorageLayout SorterLayout(const VrdxSorter_T* s, uint64_t max_count, bool key_value) {
#ifdef VRDX_EXPERIMENTS
  return ComputeLayout(max_count, kMinTile, ~0ull, true);  // the round-1 kernels keep 32-bit per-tile rows everywhere
#else
  // (a NULL sorter — legal for the pure size query — is answered like the default, AUTO, sorter)
  // One table row per tile of the smaller of the two tile shapes this sorter uses for this kind of sort.
  const TileShape* shapes = key_value ? kPairShapes : kKeysShapes;
  const uint32_t t_one = shapes[s ? (key_value ? s->pair_shape : s->keys_shape) : kDefaultOnesweepShape].tile;
  const uint32_t t_rts = shapes[s ? (key_value ? s->pair_rts_shape : s->keys_rts_shape)
                                  : (key_value ? kDefaultPairRtsShape : kDefaultKeysRtsShape)].tile;
  const uint32_t min_tile = t_one < t_rts ? t_one : t_rts;
  if (s && s->algorithm == VRDX_CUDA_ALGORITHM_ONESWEEP) return ComputeLayout(max_count, min_tile, ~0ull);
  if (s && s->algorithm == VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN) return ComputeLayout(max_count, min_tile, 0);
  // AUTO: a count below the crossover of this kind of sort may run onesweep
  return ComputeLayout(max_count, min_tile, key_value ? kAutoRtsThresholdPairs : kAutoRtsThresholdKeys);
#endif
}

struct VrdxQueryPool_T {
  int device = 0;
  uint32_t count = 0;
  unsigned long long* d_slots = nullptr;
  std::vector<int32_t> alias;     // slot -> slot whose device value it reports (itself if written by a kernel)
  std::vector<uint8_t> recorded;  // a sort has been enqueued that writes this slot
};

struct VrdxCudaImportedMemory_T {
  int device = 0;
  cudaExternalMemory_t ext = nullptr;
  void* base = nullptr;
  uint64_t size = 0;
};

struct VrdxCudaImportedSemaphore_T {
  int device = 0;
  cudaExternalSemaphore_t ext = nullptr;
  bool timeline = false;
};

namespace {

int DeviceFromHandle(const void* h) { return (int)(reinterpret_cast<uintptr_t>(h)) - 1; }

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

void NoteError(VrdxSorter s, cudaError_t e) {
  if (e != cudaSuccess) {
    int expected = 0;
    s->last_error.compare_exchange_strong(expected, (int)e);
  }
}

VrdxQueryPool_T* Pool(VkQueryPool p) { return reinterpret_cast<VrdxQueryPool_T*>(p); }

// Query-pool bookkeeping of one sort: slot i of the sort lives at pool slot query + i.
struct Stamps {
  VrdxQueryPool_T* qp = nullptr;
  uint32_t query = 0;
  int last_written = 0;  // most recent slot a kernel writes
  unsigned long long* Slot(int i) const { return qp ? qp->d_slots + query + i : nullptr; }
  // slot i is written by the next kernel enqueued
  unsigned long long* Written(int i) {
    if (!qp) return nullptr;
    qp->alias[query + i] = (int32_t)(query + i);
    qp->recorded[query + i] = 1;
    last_written = i;
    return Slot(i);
  }
  // slot i coincides with the latest written slot
  void Same(int i) {
    if (!qp) return;
    qp->alias[query + i] = (int32_t)(query + last_written);
    qp->recorded[query + i] = 1;
  }
};

// Spine of one reduce-then-scan pass over `rows` chunk rows; returns the number of launches.  The segments of the
// fused kernel wait for each other, so its grid must be co-resident: 128 tiny CTAs, at most 4 per SM.
uint32_t SpineGrid(const VrdxSorter_T* s, uint32_t rows) {
  uint32_t cap = (uint32_t)kSpineSegments;
  if ((uint32_t)s->sm_count * 4u < cap) cap = (uint32_t)s->sm_count * 4u;
  return rows < cap ? (rows ? rows : 1u) : cap;
}
uint32_t LaunchSpine(VrdxSorter s, cudaStream_t stream, bool pdl, uint32_t grid, const uint32_t* indirect, uint32_t n_or_max,
                     uint32_t tile_size, uint32_t fixed_rows, uint32_t pass, uint32_t* rows, uint32_t* seg, StorageHeader* hdr,
                     unsigned long long* ts_end) {
#if VRDX_SPINE_FUSED
  NoteError(s, LaunchEx(SpineKernel, grid, (uint32_t)kRadix, 0, stream, pdl, indirect, n_or_max, tile_size, fixed_rows, pass,
                        rows, seg, hdr, ts_end));
  return 1;
#else
  NoteError(s, LaunchEx(SpineReduceKernel, grid, (uint32_t)kRadix, 0, stream, pdl, indirect, n_or_max, tile_size, fixed_rows,
                        pass, (const uint32_t*)rows, seg, hdr));
  NoteError(s, LaunchEx(SpineApplyKernel, grid, (uint32_t)kRadix, 0, stream, pdl, indirect, n_or_max, tile_size, fixed_rows,
                        rows, (const uint32_t*)seg, ts_end));
  return 2;
#endif
}

// What a sort compares: the reference's plan (uint32 ascending, four 8-bit digits) or the plan
// of a VrdxCudaSortKeyInfo (key type / order as a codec, bit sub-range as fewer, narrower digits).
struct SortPlan {
  DigitPlan digits;
  bool reference;   // exactly the reference's sort: every kernel flavour implements it
  bool all_bits;    // the digits cover all 32 bits: equal keys are indistinguishable
};

SortPlan ReferencePlan() {
  SortPlan p{};
  p.digits.passes = kPasses;
  for (int i = 0; i < kPasses; ++i) {
    p.digits.shift[i] = (uint32_t)(i * kRadixBits);
    p.digits.mask[i] = kRadix - 1;
  }
  p.reference = true;
  p.all_bits = true;
  return p;
}

bool PlanFromKeyInfo(const VrdxCudaSortKeyInfo* info, SortPlan* out) {
  *out = ReferencePlan();
  if (!info) return true;
  if (info->structSize < sizeof(VrdxCudaSortKeyInfo) || info->beginBit > info->endBit || info->endBit > 32 ||
      (uint32_t)info->keyType > (uint32_t)VRDX_CUDA_KEY_TYPE_FLOAT32 ||
      (uint32_t)info->order > (uint32_t)VRDX_CUDA_SORT_ORDER_DESCENDING)
    return false;
  DigitPlan& d = out->digits;
  const uint32_t bits = info->endBit - info->beginBit;
  d.passes = (bits + kRadixBits - 1) / kRadixBits;
  for (uint32_t i = 0; i < (uint32_t)kPasses; ++i) {
    const uint32_t lo = info->beginBit + i * kRadixBits;
    const uint32_t width = i < d.passes ? (info->endBit - lo < (uint32_t)kRadixBits ? info->endBit - lo : kRadixBits) : 0;
    d.shift[i] = i < d.passes ? lo : 0;
    d.mask[i] = (1u << width) - 1u;
  }
  d.codec.cmask = info->keyType == VRDX_CUDA_KEY_TYPE_UINT32 ? 0u : 0x80000000u;
  d.codec.fmask = info->keyType == VRDX_CUDA_KEY_TYPE_FLOAT32 ? 0x7FFFFFFFu : 0u;
  d.codec.dmask = info->order == VRDX_CUDA_SORT_ORDER_DESCENDING ? 0xFFFFFFFFu : 0u;
  out->all_bits = bits == 32;
  out->reference = out->all_bits && !d.codec.cmask && !d.codec.fmask && !d.codec.dmask;
  return true;
}

// The body of every vrdxCmdSort* (reference: gpuSort, h.in:344-507).
void EnqueueSort(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t n_or_max,
                 VkBuffer indirectBuffer, VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                 VkDeviceSize keysOffset, VkBuffer valuesBuffer, VkDeviceSize valuesOffset,
                 VkBuffer storageBuffer, VkDeviceSize storageOffset, VkQueryPool queryPool,
                 uint32_t query, const SortPlan& plan = ReferencePlan()) {
  if (!sorter) return;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(commandBuffer);
  DeviceGuard guard(sorter->device);
  uint32_t launches = 0;

  auto addr = [](VkBuffer b, VkDeviceSize off) -> char* {
    return b ? reinterpret_cast<char*>(b) + off : nullptr;
  };
  const uint32_t* indirect = reinterpret_cast<const uint32_t*>(addr(indirectBuffer, indirectOffset));
  uint32_t* keys = reinterpret_cast<uint32_t*>(addr(keysBuffer, keysOffset));
  uint32_t* values = reinterpret_cast<uint32_t*>(addr(valuesBuffer, valuesOffset));
  char* storage = addr(storageBuffer, storageOffset);
  const bool kv = values != nullptr;

  Stamps st;
  if (queryPool) {
    VrdxQueryPool_T* qp = Pool(queryPool);
    if ((uint64_t)query + 15 > qp->count) {
      NoteError(sorter, cudaErrorInvalidValue);
    } else {
      st.qp = qp;
      st.query = query;
      st.Written(0);  // slot 0 and the reset of slots 1..14 are written by the first kernel of the sort (StampStart)
    }
  }
  const uint32_t passes = plan.digits.passes;
  if (n_or_max == 0 || !keys || !storage || passes == 0) {
    if (n_or_max != 0 && passes != 0) NoteError(sorter, cudaErrorInvalidValue);
    for (int i = 1; i < 15; ++i) st.Same(i);
    sorter->last_launches.store(0);
    return;
  }
  // 30-bit look-back cells: counts of 2^30 and above take the reduce-then-scan path (32-bit counts).
  const bool use_rts = sorter->algorithm == VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN ||
                       (uint64_t)n_or_max >= kMaxOnesweepCount ||
                       (sorter->algorithm == VRDX_CUDA_ALGORITHM_AUTO &&
                        n_or_max >= (kv ? kAutoRtsThresholdPairs : kAutoRtsThresholdKeys));

  const StorageLayout lay = SorterLayout(sorter, n_or_max, kv);
  StorageHeader* hdr = reinterpret_cast<StorageHeader*>(storage + lay.header_offset);
  uint32_t* status[2] = {reinterpret_cast<uint32_t*>(storage + lay.status_a_offset),
                         reinterpret_cast<uint32_t*>(storage + lay.status_b_offset)};
  uint32_t* keys_alt = reinterpret_cast<uint32_t*>(storage + lay.keys_alt_offset);
  uint32_t* vals_alt = reinterpret_cast<uint32_t*>(storage + lay.values_alt_offset);

  // Alignment: the header and tables are accessed as 16-byte vectors, keys/values as words.
  if ((reinterpret_cast<uintptr_t>(storage) & (kOffsetAlignment - 1)) || (reinterpret_cast<uintptr_t>(keys) & 3u) ||
      (kv && (reinterpret_cast<uintptr_t>(values) & 3u)) || (reinterpret_cast<uintptr_t>(indirect) & 3u)) {
    NoteError(sorter, cudaErrorInvalidValue);
    for (int i = 1; i < 15; ++i) st.Same(i);
    sorter->last_launches.store(0);
    return;
  }
  // Key types, order and bit sub-ranges run the GENERIC instantiation of the same tile kernel.
  const bool generic = !plan.reference;
  const bool pdl = sorter->pdl;
  const TileShape* chosen = use_rts ? (kv ? &kPairShapes[sorter->pair_rts_shape] : &kKeysShapes[sorter->keys_rts_shape])
                                    : (kv ? &kPairShapes[sorter->pair_shape] : &kKeysShapes[sorter->keys_shape]);
  const TileShape& shape = *chosen;
  uint32_t tile_size = shape.tile;
#ifdef VRDX_EXPERIMENTS
  const ExperimentKernel* ek = generic ? nullptr : PickExperiment(sorter->exp, kv, keys, values, storage);
  if (ek) tile_size = ek->tile;
#endif
  const uint32_t tiles = (uint32_t)CeilDiv(n_or_max, tile_size);
  uint32_t pass_grid = tiles;
#ifdef VRDX_EXPERIMENTS
  // experiment: reduce-then-scan over ranges of consecutive tiles (RangePassKernel), tables sized by a constant
  const bool persistent = use_rts && ek && ek->range_tiles != 0;
  uint32_t range_tiles = 1, ranges = tiles;
  if (persistent) {
    range_tiles = ek->range_tiles > 0 ? (uint32_t)ek->range_tiles   // fixed range length, capped at kMaxRanges rows
                                      : (uint32_t)CeilDiv(tiles, (uint64_t)sorter->sm_count * (uint64_t)*ek->ctas_per_sm);
    const uint32_t floor_tiles = (uint32_t)CeilDiv(tiles, (uint64_t)kMaxRanges);
    if (range_tiles < floor_tiles) range_tiles = floor_tiles;
    ranges = (uint32_t)CeilDiv(tiles, (uint64_t)range_tiles);
  }
#endif
#ifdef VRDX_EXPERIMENTS
  if (ek) pass_grid = ExperimentGrid(*ek, tiles, use_rts, sorter->sm_count);
#endif

  if (!use_rts) {
    // Reset per-sort state inside the stream (reference: vkCmdFillBuffer of the global
    // histogram, h.in:382): header (histograms, tickets) + the pass-0 look-back cells.
    const uint64_t reset_bytes = lay.status_a_offset + (uint64_t)tiles * kRadix * sizeof(uint32_t);  // multiple of 16
    {
      // a kernel rather than cudaMemsetAsync: it carries the start stamp, and the histogram kernel launched
      // next (programmatic dependent launch) counts the caller's keys while it runs
      const uint64_t blocks = CeilDiv(reset_bytes / 16, (uint64_t)256 * 4);
      const uint64_t cap = (uint64_t)sorter->sm_count * 4;
      NoteError(sorter, LaunchEx(ResetKernel, (uint32_t)(blocks < cap ? blocks : cap), 256u, 0, stream, false,
                                 reinterpret_cast<uint4*>(storage), reset_bytes / 16, st.Slot(0)));
    }
    ++launches;
    if (n_or_max >= sorter->hist_private_min_count) {
      // lane-private (conflict-free) bins, one 1024-thread CTA per SM
      uint64_t chunks = CeilDiv(n_or_max, (uint64_t)kHistPrivChunk);
      uint32_t grid = (uint32_t)(chunks < (uint64_t)sorter->sm_count ? chunks : (uint64_t)sorter->sm_count);
      NoteError(sorter, LaunchEx(generic ? HistogramKernelPrivate<true> : HistogramKernelPrivate<false>, grid,
                                 kHistPrivThreads, kHistPrivSmemBytes, stream, pdl, (const uint32_t*)keys, indirect,
                                 n_or_max, hdr, st.Written(1), plan.digits));
    } else {
      uint64_t chunks = CeilDiv(n_or_max, (uint64_t)kHistChunk);
      uint64_t cap = (uint64_t)sorter->sm_count * 4;
      uint32_t grid = (uint32_t)(chunks < cap ? (chunks ? chunks : 1) : cap);
      NoteError(sorter, LaunchEx(generic ? HistogramKernel<true> : HistogramKernel<false>, grid, kHistThreads, 0, stream,
                                 pdl, (const uint32_t*)keys, indirect, n_or_max, hdr, st.Written(1), plan.digits));
    }
    ++launches;
  } else {
    st.Same(1);  // reduce-then-scan keeps no state across passes: every table it reads is written first
  }

  for (uint32_t pass = 0; pass < passes; ++pass) {
    PassArgs args{};
    args.indirect = indirect;
    args.n_or_max = n_or_max;
    args.pass = pass;
    args.shift = plan.digits.shift[pass];
    args.mask = plan.digits.mask[pass];
    if (pass == 0) args.codec_in = plan.digits.codec;
    if (pass + 1 == passes) args.codec_out = plan.digits.codec;
    args.order_free = (!kv && pass == 0 && plan.all_bits) ? 1u : 0u;
#ifdef VRDX_EXPERIMENTS
    args.range_tiles = range_tiles;
#endif
    args.static_tiles = (!use_rts && tiles <= 2u * (uint32_t)sorter->sm_count) ? 1u : 0u;  // every shape fits >= 2 CTAs/SM
    args.words_only = (!kv && plan.all_bits && sorter->relaxed_equal_low_bits) ? 1u : 0u;
    args.hdr = hdr;
    args.status = status[pass & 1];
    args.status_next = (pass + 1 < passes) ? status[(pass + 1) & 1] : nullptr;
    // ping-pong: user -> scratch on even passes, scratch -> user on odd ones (h.in:417-427)
    args.keys_in = (pass & 1) ? keys_alt : keys;
    args.keys_out = (pass & 1) ? keys : keys_alt;
    args.vals_in = kv ? ((pass & 1) ? vals_alt : values) : nullptr;
    args.vals_out = kv ? ((pass & 1) ? values : vals_alt) : nullptr;

#ifdef VRDX_EXPERIMENTS
    if (use_rts && persistent) {
      // The reference's three stages with tables sized by the machine: `ranges` co-resident CTAs, CTA r owns a
      // contiguous tile range; table A = one digit-count row per range -> exclusive prefixes (spine), the
      // spine's segment sums live in the rows after them.  Timestamps fall where the reference puts them.
      args.status = status[0];
      args.status_next = nullptr;
      if (pass == 0 && st.qp)  // the range upsweep carries no start stamp: an empty reset launch writes slot 0
        NoteError(sorter, LaunchEx(ResetKernel, 1u, 256u, 0, stream, false, reinterpret_cast<uint4*>(storage), (uint64_t)0,
                                   st.Slot(0)));
      args.ts_end = st.Written(2 + 3 * pass + 0);
      NoteError(sorter, ek->launch_upsweep(stream, ranges, args, pdl));
      uint32_t* seg = status[0] + (size_t)ranges * kRadix;
      const uint32_t seg_grid = SpineGrid(sorter, ranges);
      launches += LaunchSpine(sorter, stream, pdl, seg_grid, indirect, n_or_max, tile_size, ranges, pass, status[0], seg, hdr,
                              st.Written(2 + 3 * pass + 1));
      args.ts_end = st.Written(2 + 3 * pass + 2);
      NoteError(sorter, ek->launch(stream, ranges, args, 1, pdl));
      launches += 2;
      continue;
    }
#endif
    if (use_rts) {
      // the reference's three stages: table A = per-tile digit prefixes inside a chunk of kSpineChunk tiles
      // (16-bit), table B = spine chunk sums -> exclusive prefixes; timestamps fall where the reference puts them
      args.status = status[0];
      args.status_next = status[1];
      const uint32_t chunks = (uint32_t)CeilDiv(tiles, (uint64_t)kSpineChunk);
      // keys-only sorts over all 32 bits: one flag byte per tile after the spine's rows (chunk prefixes, segment sums)
      // (only where keys that agree below the digit may swap places: the flags rely on the input of a pass being
      // sorted by exactly those bits)
      if (args.words_only) {
        args.tile_flags = reinterpret_cast<uint8_t*>(status[1] + ((size_t)chunks + kSpineSegments + 1) * kRadix);
        // Tiles of exactly two runs are worth their own kernel flavour only where they are common: on uniform keys
        // the runs of pass p are count / 2^shift keys long, and between half a tile and three tiles that makes
        // more than a third of the tiles two-run (2^28 keys: pass 2, runs of 4096; 2^29: pass 2, runs of 8192).
        const uint64_t run = (uint64_t)n_or_max >> args.shift;
        args.two_runs = (pass > 0 && (sorter->two_runs < 0 ? (2 * run >= tile_size && run <= 3ull * tile_size)
                                                           : sorter->two_runs != 0)) ? 1u : 0u;
      }
      args.ts_end = st.Written(2 + 3 * pass + 0);
      args.ts_start = pass == 0 ? st.Slot(0) : nullptr;
#ifdef VRDX_EXPERIMENTS
      if (ek) NoteError(sorter, ek->launch_upsweep(stream, chunks, args, pdl));
      else
#endif
      NoteError(sorter, shape.launch_upsweep(stream, chunks, args, pdl));
      // spine scratch: chunk prefixes occupy rows [0, chunks) of status B, segment sums the rows after them
      uint32_t* seg = status[1] + (size_t)chunks * kRadix;
      const uint32_t seg_grid = SpineGrid(sorter, chunks);
      launches += LaunchSpine(sorter, stream, pdl, seg_grid, indirect, n_or_max, tile_size, 0u, pass, status[1], seg, hdr,
                              st.Written(2 + 3 * pass + 1));
      args.ts_end = st.Written(2 + 3 * pass + 2);
#ifdef VRDX_EXPERIMENTS
      if (ek) NoteError(sorter, ek->launch(stream, pass_grid, args, 1, pdl));
      else
#endif
      NoteError(sorter, shape.launch_pass(stream, tiles, args, 1, generic, pdl));
      launches += 2;
      continue;
    }
    // One fused kernel per pass: the reference's upsweep and spine slots collapse onto its start.
    st.Same(2 + 3 * pass + 0);
    st.Same(2 + 3 * pass + 1);
    args.ts_end = st.Written(2 + 3 * pass + 2);
#ifdef VRDX_EXPERIMENTS
    if (ek) NoteError(sorter, ek->launch(stream, pass_grid, args, 0, pdl));
    else
#endif
    NoteError(sorter, shape.launch_pass(stream, pass_grid, args, 0, generic, pdl));
    ++launches;
  }
  for (uint32_t pass = passes; pass < (uint32_t)kPasses; ++pass)  // passes a bit sub-range does not need
    for (int j = 0; j < 3; ++j) st.Same(2 + 3 * (int)pass + j);
  if (passes & 1u) {
    // an odd number of passes ends in the scratch halves: bring [0, count) home
    const uint64_t blocks = CeilDiv((uint64_t)n_or_max, (uint64_t)1024);
    const uint64_t cap = (uint64_t)sorter->sm_count * 8;
    NoteError(sorter, LaunchEx(CopyBackKernel, (uint32_t)(blocks < cap ? blocks : cap), 256u, 0, stream, pdl, indirect,
                               n_or_max, (const uint32_t*)keys_alt, keys, kv ? (const uint32_t*)vals_alt : nullptr,
                               kv ? values : nullptr, st.Written(14)));
    ++launches;
  } else {
    st.Same(14);
  }
  sorter->last_launches.store(launches);
}

}  // namespace

// ============================================================================ C ABI

extern "C" {

VkResult vrdxCudaCreateSorter(const VrdxSorterCreateInfo* pCreateInfo,
                              const VrdxCudaSorterOptions* pOptions, VrdxSorter* pSorter) {
  if (!pCreateInfo || !pSorter) return VK_ERROR_INITIALIZATION_FAILED;
  const int dev = DeviceFromHandle(pCreateInfo->device);
  if (pCreateInfo->physicalDevice && DeviceFromHandle(pCreateInfo->physicalDevice) != dev)
    return VK_ERROR_INITIALIZATION_FAILED;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  if (dev < 0 || dev >= count) return VK_ERROR_INITIALIZATION_FAILED;

  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return VK_ERROR_INITIALIZATION_FAILED;
  if (prop.major != 10) return VK_ERROR_FEATURE_NOT_PRESENT;  // kernels are built for sm_100a only

  DeviceGuard guard(dev);
  // Developer overrides (environment) are read ONCE, here, into the sorter; nothing on the sort path
  // reads the environment or any process-wide variable.
  int keys_shape = kDefaultOnesweepShape, pair_shape = kDefaultOnesweepShape;
  int keys_rts_shape = kDefaultKeysRtsShape, pair_rts_shape = kDefaultPairRtsShape;
  uint32_t experiment = 0;
  VrdxCudaTileLoad tile_load = VRDX_CUDA_TILE_LOAD_AUTO;
  if (const char* e = getenv("VRDX_KEYS_SHAPE")) keys_shape = atoi(e);
  if (const char* e = getenv("VRDX_KV_SHAPE")) pair_shape = atoi(e);
  if (const char* e = getenv("VRDX_KEYS_RTS_SHAPE")) keys_rts_shape = atoi(e);
  if (const char* e = getenv("VRDX_KV_RTS_SHAPE")) pair_rts_shape = atoi(e);
  if (const char* e = getenv("VRDX_EXPERIMENT")) experiment = (uint32_t)atoi(e);
  if (const char* e = getenv("VRDX_TILE_LOAD")) tile_load = (VrdxCudaTileLoad)atoi(e);
  if (pOptions && pOptions->structSize >= sizeof(VrdxCudaSorterOptions)) {
    if (pOptions->tileLoad != VRDX_CUDA_TILE_LOAD_AUTO) tile_load = pOptions->tileLoad;
    if (pOptions->reserved[0]) keys_shape = keys_rts_shape = (int)pOptions->reserved[0] - 1;
    if (pOptions->reserved[1]) pair_shape = pair_rts_shape = (int)pOptions->reserved[1] - 1;
    if (pOptions->reserved[2]) experiment = pOptions->reserved[2];
  }
  if (keys_shape < 0 || keys_shape >= kNumKeysShapes || pair_shape < 0 || pair_shape >= kNumPairShapes ||
      keys_rts_shape < 0 || keys_rts_shape >= kNumKeysShapes || pair_rts_shape < 0 || pair_rts_shape >= kNumPairShapes)
    return VK_ERROR_INITIALIZATION_FAILED;
#ifndef VRDX_EXPERIMENTS
  // the TMA-staged persistent kernels and the other losing variants are not in the product library
  if (experiment != 0 || tile_load == VRDX_CUDA_TILE_LOAD_TMA) return VK_ERROR_FEATURE_NOT_PRESENT;
#endif
  // "Pipeline creation" (h.in:141-262): opt every kernel this sorter can launch into its shared-memory
  // footprint.  cudaFuncSetAttribute applies to the CURRENT device, so this runs for every sorter, under
  // the device guard above: a process that drives several GPUs prepares each of them.
  int unused = 0, keys_rts_ctas = 0, pair_rts_ctas = 0;
  if (kKeysShapes[kDefaultOnesweepShape].prepare(&unused) != cudaSuccess ||  // vrdxCudaCmdSortEx runs the defaults too
      kPairShapes[kDefaultOnesweepShape].prepare(&unused) != cudaSuccess ||
      kKeysShapes[kDefaultKeysRtsShape].prepare(&unused) != cudaSuccess ||
      kPairShapes[kDefaultPairRtsShape].prepare(&unused) != cudaSuccess ||
      kKeysShapes[keys_shape].prepare(&unused) != cudaSuccess || kPairShapes[pair_shape].prepare(&unused) != cudaSuccess ||
      kKeysShapes[keys_rts_shape].prepare(&keys_rts_ctas) != cudaSuccess ||
      kPairShapes[pair_rts_shape].prepare(&pair_rts_ctas) != cudaSuccess || keys_rts_ctas < 1 || pair_rts_ctas < 1 ||
      cudaFuncSetAttribute(HistogramKernelPrivate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kHistPrivSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(HistogramKernelPrivate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kHistPrivSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(DistPrefixHistogramKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kDistHistMaxSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(DistClassCountKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kDistCountSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  VrdxSorter_T* s = new (std::nothrow) VrdxSorter_T();
  if (!s) return VK_ERROR_OUT_OF_HOST_MEMORY;
  s->device = dev;
  s->sm_count = prop.multiProcessorCount;
  s->cc_major = prop.major;
  s->cc_minor = prop.minor;
  s->keys_shape = keys_shape;
  s->pair_shape = pair_shape;
  s->keys_rts_shape = keys_rts_shape;
  s->pair_rts_shape = pair_rts_shape;
  s->tile_load = tile_load;
  s->keys_rts_ctas = keys_rts_ctas;
  s->pair_rts_ctas = pair_rts_ctas;
  if (const char* e = getenv("VRDX_ALGORITHM")) s->algorithm = (VrdxCudaAlgorithm)atoi(e);
  if (const char* e = getenv("VRDX_PDL")) s->pdl = atoi(e) != 0;
  if (const char* e = getenv("VRDX_RELAXED")) s->relaxed_equal_low_bits = atoi(e) != 0;
  if (const char* e = getenv("VRDX_TWO_RUNS")) s->two_runs = atoi(e);
  if (const char* e = getenv("VRDX_HIST_PRIVATE_MIN")) s->hist_private_min_count = (uint32_t)strtoul(e, nullptr, 10);
#ifdef VRDX_EXPERIMENTS
  if (!PrepareExperiments(&s->exp, experiment, tile_load, s->sm_count)) {
    cudaGetLastError();
    delete s;
    return VK_ERROR_INITIALIZATION_FAILED;
  }
#endif
  if (pOptions && pOptions->structSize >= sizeof(VrdxCudaSorterOptions) &&
      pOptions->algorithm != VRDX_CUDA_ALGORITHM_AUTO)
    s->algorithm = pOptions->algorithm;
  *pSorter = s;
  return VK_SUCCESS;
}

VkResult vrdxCreateSorter(const VrdxSorterCreateInfo* pCreateInfo, VrdxSorter* pSorter) {
  return vrdxCudaCreateSorter(pCreateInfo, nullptr, pSorter);
}

void vrdxDestroySorter(VrdxSorter sorter) {
  if (!sorter) return;
  delete sorter;
}

void vrdxGetSorterStorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                      VrdxSorterStorageRequirements* requirements) {
  if (!requirements) return;
  const StorageLayout lay = SorterLayout(sorter, maxElementCount, false);
  requirements->size = lay.total_keys;
  requirements->usage = VK_BUFFER_USAGE_STORAGE_BUFFER_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT;
}

void vrdxGetSorterKeyValueStorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                              VrdxSorterStorageRequirements* requirements) {
  if (!requirements) return;
  const StorageLayout lay = SorterLayout(sorter, maxElementCount, true);
  requirements->size = lay.total_kv;
  requirements->usage = VK_BUFFER_USAGE_STORAGE_BUFFER_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT;
}

void vrdxCmdSort(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                 VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer storageBuffer,
                 VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query) {
  EnqueueSort(commandBuffer, sorter, elementCount, nullptr, 0, keysBuffer, keysOffset, nullptr, 0,
              storageBuffer, storageOffset, queryPool, query);
}

void vrdxCmdSortIndirect(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t maxElementCount,
                         VkBuffer indirectBuffer, VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                         VkDeviceSize keysOffset, VkBuffer storageBuffer,
                         VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query) {
  if (sorter && !indirectBuffer) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  EnqueueSort(commandBuffer, sorter, maxElementCount, indirectBuffer, indirectOffset, keysBuffer,
              keysOffset, nullptr, 0, storageBuffer, storageOffset, queryPool, query);
}

void vrdxCmdSortKeyValue(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                         VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer valuesBuffer,
                         VkDeviceSize valuesOffset, VkBuffer storageBuffer,
                         VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query) {
  if (sorter && !valuesBuffer && elementCount) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  EnqueueSort(commandBuffer, sorter, elementCount, nullptr, 0, keysBuffer, keysOffset, valuesBuffer,
              valuesOffset, storageBuffer, storageOffset, queryPool, query);
}

void vrdxCmdSortKeyValueIndirect(VkCommandBuffer commandBuffer, VrdxSorter sorter,
                                 uint32_t maxElementCount, VkBuffer indirectBuffer,
                                 VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                                 VkDeviceSize keysOffset, VkBuffer valuesBuffer,
                                 VkDeviceSize valuesOffset, VkBuffer storageBuffer,
                                 VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query) {
  if (sorter && (!indirectBuffer || (!valuesBuffer && maxElementCount))) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  EnqueueSort(commandBuffer, sorter, maxElementCount, indirectBuffer, indirectOffset, keysBuffer,
              keysOffset, valuesBuffer, valuesOffset, storageBuffer, storageOffset, queryPool, query);
}

// ---------------------------------------------------------------------------- extensions

void vrdxCudaCmdSortEx(VkCommandBuffer commandBuffer, VrdxSorter sorter, const VrdxCudaSortKeyInfo* pKeyInfo,
                       uint32_t elementCount, VkBuffer indirectBuffer, VkDeviceSize indirectOffset,
                       VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer valuesBuffer,
                       VkDeviceSize valuesOffset, VkBuffer storageBuffer, VkDeviceSize storageOffset,
                       VkQueryPool queryPool, uint32_t query) {
  if (!sorter) return;
  SortPlan plan;
  if (!PlanFromKeyInfo(pKeyInfo, &plan)) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  EnqueueSort(commandBuffer, sorter, elementCount, indirectBuffer, indirectOffset, keysBuffer, keysOffset,
              valuesBuffer, valuesOffset, storageBuffer, storageOffset, queryPool, query, plan);
}

int vrdxCudaGetLastError(VrdxSorter sorter) {
  if (!sorter) return (int)cudaErrorInvalidResourceHandle;
  return sorter->last_error.exchange(0);
}

const char* vrdxCudaGetErrorString(int error) { return cudaGetErrorString((cudaError_t)error); }

uint32_t vrdxCudaGetLastLaunchCount(VrdxSorter sorter) {
  return sorter ? sorter->last_launches.load() : 0u;
}

VkResult vrdxCudaCreateQueryPool(VkDevice device, uint32_t queryCount, VkQueryPool* pQueryPool) {
  if (!pQueryPool || queryCount == 0) return VK_ERROR_INITIALIZATION_FAILED;
  const int dev = DeviceFromHandle(device);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  DeviceGuard guard(dev);
  VrdxQueryPool_T* qp = new (std::nothrow) VrdxQueryPool_T();
  if (!qp) return VK_ERROR_OUT_OF_HOST_MEMORY;
  qp->device = dev;
  qp->count = queryCount;
  qp->alias.assign(queryCount, -1);
  qp->recorded.assign(queryCount, 0);
  if (cudaMalloc(&qp->d_slots, sizeof(unsigned long long) * queryCount) != cudaSuccess ||
      cudaMemset(qp->d_slots, 0, sizeof(unsigned long long) * queryCount) != cudaSuccess) {
    cudaGetLastError();
    if (qp->d_slots) cudaFree(qp->d_slots);
    delete qp;
    return VK_ERROR_OUT_OF_DEVICE_MEMORY;
  }
  *pQueryPool = reinterpret_cast<VkQueryPool>(qp);
  return VK_SUCCESS;
}

void vrdxCudaDestroyQueryPool(VkQueryPool queryPool) {
  if (!queryPool) return;
  VrdxQueryPool_T* qp = Pool(queryPool);
  DeviceGuard guard(qp->device);
  cudaFree(qp->d_slots);
  delete qp;
}

VkResult vrdxCudaGetQueryPoolResults(VkQueryPool queryPool, uint32_t firstQuery,
                                     uint32_t queryCount, uint64_t* pNanoseconds) {
  if (!queryPool || !pNanoseconds) return VK_ERROR_INITIALIZATION_FAILED;
  VrdxQueryPool_T* qp = Pool(queryPool);
  if ((uint64_t)firstQuery + queryCount > qp->count) return VK_ERROR_INITIALIZATION_FAILED;
  for (uint32_t i = 0; i < queryCount; ++i)
    if (!qp->recorded[firstQuery + i]) return VK_NOT_READY;
  DeviceGuard guard(qp->device);
  std::vector<unsigned long long> host(qp->count);
  if (cudaMemcpy(host.data(), qp->d_slots, sizeof(unsigned long long) * qp->count, cudaMemcpyDeviceToHost) !=
      cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_UNKNOWN;
  }
  unsigned long long prev = 0, base = 0;
  for (uint32_t i = 0; i < queryCount; ++i) {
    unsigned long long t = host[qp->alias[firstQuery + i]];
    if (t == 0) return VK_NOT_READY;  // the kernel that writes this slot has not finished
    if (t < prev) t = prev;           // stages of one sort are ordered; guard against timer granularity
    if (i == 0) base = t;
    pNanoseconds[i] = t - base;
    prev = t;
  }
  return VK_SUCCESS;
}

VkResult vrdxCudaImportMemoryFd(VkDevice device, int fd, VkDeviceSize allocationSize, int dedicated,
                                VrdxCudaImportedMemory* pMemory) {
  if (!pMemory || fd < 0 || allocationSize == 0) return VK_ERROR_INITIALIZATION_FAILED;
  const int dev = DeviceFromHandle(device);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  DeviceGuard guard(dev);
  cudaExternalMemoryHandleDesc hd{};
  hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  hd.size = allocationSize;
  hd.flags = dedicated ? cudaExternalMemoryDedicated : 0;
  cudaExternalMemory_t ext = nullptr;
  if (cudaImportExternalMemory(&ext, &hd) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  cudaExternalMemoryBufferDesc bd{};
  bd.offset = 0;
  bd.size = allocationSize;
  void* base = nullptr;
  if (cudaExternalMemoryGetMappedBuffer(&base, ext, &bd) != cudaSuccess) {
    cudaGetLastError();
    cudaDestroyExternalMemory(ext);
    return VK_ERROR_OUT_OF_DEVICE_MEMORY;
  }
  VrdxCudaImportedMemory_T* m = new (std::nothrow) VrdxCudaImportedMemory_T();
  if (!m) {
    cudaFree(base);
    cudaDestroyExternalMemory(ext);
    return VK_ERROR_OUT_OF_HOST_MEMORY;
  }
  m->device = dev;
  m->ext = ext;
  m->base = base;
  m->size = allocationSize;
  *pMemory = m;
  return VK_SUCCESS;
}

VkBuffer vrdxCudaImportedMemoryBuffer(VrdxCudaImportedMemory memory, VkDeviceSize memoryOffset) {
  if (!memory || memoryOffset >= memory->size) return VK_NULL_HANDLE;
  return reinterpret_cast<VkBuffer>(static_cast<char*>(memory->base) + memoryOffset);
}

void vrdxCudaReleaseImportedMemory(VrdxCudaImportedMemory memory) {
  if (!memory) return;
  DeviceGuard guard(memory->device);
  cudaFree(memory->base);
  cudaDestroyExternalMemory(memory->ext);
  delete memory;
}

void vrdxCudaGetSorterKeys64StorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                                VrdxSorterStorageRequirements* requirements) {
  if (!requirements) return;
  const StorageLayout lay = SorterLayout(sorter, maxElementCount, true);
  // lo[] | hi[] | storage of a key-value sort
  requirements->size = 2 * AlignUp((uint64_t)maxElementCount * sizeof(uint32_t), (uint64_t)kOffsetAlignment) + lay.total_kv;
  requirements->usage = VK_BUFFER_USAGE_STORAGE_BUFFER_BIT | VK_BUFFER_USAGE_TRANSFER_DST_BIT;
}

void vrdxCudaCmdSortKeys64(VkCommandBuffer commandBuffer, VrdxSorter sorter, const VrdxCudaSortKeyInfo* pKeyInfo,
                           uint32_t elementCount, VkBuffer indirectBuffer, VkDeviceSize indirectOffset,
                           VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer storageBuffer,
                           VkDeviceSize storageOffset) {
  if (!sorter) return;
  KeyCodec64 codec{0ull, 0ull, 0ull};
  if (pKeyInfo) {
    if (pKeyInfo->structSize < sizeof(VrdxCudaSortKeyInfo) ||
        (uint32_t)pKeyInfo->keyType > (uint32_t)VRDX_CUDA_KEY_TYPE_FLOAT32 ||
        (uint32_t)pKeyInfo->order > (uint32_t)VRDX_CUDA_SORT_ORDER_DESCENDING) {
      NoteError(sorter, cudaErrorInvalidValue);
      return;
    }
    codec.cmask = pKeyInfo->keyType == VRDX_CUDA_KEY_TYPE_UINT32 ? 0ull : 0x8000000000000000ull;
    codec.fmask = pKeyInfo->keyType == VRDX_CUDA_KEY_TYPE_FLOAT32 ? 0x7FFFFFFFFFFFFFFFull : 0ull;
    codec.dmask = pKeyInfo->order == VRDX_CUDA_SORT_ORDER_DESCENDING ? ~0ull : 0ull;
  }
  char* keys = keysBuffer ? reinterpret_cast<char*>(keysBuffer) + keysOffset : nullptr;
  char* storage = storageBuffer ? reinterpret_cast<char*>(storageBuffer) + storageOffset : nullptr;
  if (elementCount == 0) {
    sorter->last_launches.store(0);
    return;
  }
  if (!keys || !storage || (reinterpret_cast<uintptr_t>(keys) & 7u) || (reinterpret_cast<uintptr_t>(storage) & (kOffsetAlignment - 1))) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  const uint32_t* indirect =
      indirectBuffer ? reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(indirectBuffer) + indirectOffset) : nullptr;
  const uint64_t half = AlignUp((uint64_t)elementCount * sizeof(uint32_t), (uint64_t)kOffsetAlignment);
  uint32_t* lo = reinterpret_cast<uint32_t*>(storage);
  uint32_t* hi = reinterpret_cast<uint32_t*>(storage + half);
  char* inner = storage + 2 * half;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(commandBuffer);
  uint32_t launches = 0;
  {
    DeviceGuard guard(sorter->device);
    const uint64_t blocks = CeilDiv((uint64_t)elementCount, (uint64_t)1024);
    const uint64_t cap = (uint64_t)sorter->sm_count * 8;
    const uint32_t grid = (uint32_t)(blocks < cap ? blocks : cap);
    NoteError(sorter, LaunchEx(Split64Kernel, grid, 256u, 0, stream, false, indirect, elementCount,
                               reinterpret_cast<const unsigned long long*>(keys), lo, hi, codec));
    ++launches;
  }
  // stable by the low word (payload: high word), then stable by the high word (payload: low word)
  EnqueueSort(commandBuffer, sorter, elementCount, indirectBuffer, indirectOffset, reinterpret_cast<VkBuffer>(lo), 0,
              reinterpret_cast<VkBuffer>(hi), 0, reinterpret_cast<VkBuffer>(inner), 0, VK_NULL_HANDLE, 0);
  launches += sorter->last_launches.load();
  EnqueueSort(commandBuffer, sorter, elementCount, indirectBuffer, indirectOffset, reinterpret_cast<VkBuffer>(hi), 0,
              reinterpret_cast<VkBuffer>(lo), 0, reinterpret_cast<VkBuffer>(inner), 0, VK_NULL_HANDLE, 0);
  launches += sorter->last_launches.load();
  {
    DeviceGuard guard(sorter->device);
    const uint64_t blocks = CeilDiv((uint64_t)elementCount, (uint64_t)1024);
    const uint64_t cap = (uint64_t)sorter->sm_count * 8;
    const uint32_t grid = (uint32_t)(blocks < cap ? blocks : cap);
    NoteError(sorter, LaunchEx(Merge64Kernel, grid, 256u, 0, stream, false, indirect, elementCount,
                               (const uint32_t*)lo, (const uint32_t*)hi, reinterpret_cast<unsigned long long*>(keys),
                               codec));
    ++launches;
  }
  sorter->last_launches.store(launches);
}

VkResult vrdxCudaImportSemaphoreFd(VkDevice device, int fd, int timeline, VrdxCudaImportedSemaphore* pSemaphore) {
  if (!pSemaphore || fd < 0) return VK_ERROR_INITIALIZATION_FAILED;
  const int dev = DeviceFromHandle(device);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  DeviceGuard guard(dev);
  cudaExternalSemaphoreHandleDesc hd{};
  hd.type = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  cudaExternalSemaphore_t ext = nullptr;
  if (cudaImportExternalSemaphore(&ext, &hd) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  VrdxCudaImportedSemaphore_T* s = new (std::nothrow) VrdxCudaImportedSemaphore_T();
  if (!s) {
    cudaDestroyExternalSemaphore(ext);
    return VK_ERROR_OUT_OF_HOST_MEMORY;
  }
  s->device = dev;
  s->ext = ext;
  s->timeline = timeline != 0;
  *pSemaphore = s;
  return VK_SUCCESS;
}

VkResult vrdxCudaCmdWaitSemaphore(VkCommandBuffer commandBuffer, VrdxCudaImportedSemaphore semaphore, uint64_t value) {
  if (!semaphore) return VK_ERROR_INITIALIZATION_FAILED;
  DeviceGuard guard(semaphore->device);
  cudaExternalSemaphoreWaitParams wp{};
  wp.params.fence.value = semaphore->timeline ? value : 0;
  if (cudaWaitExternalSemaphoresAsync(&semaphore->ext, &wp, 1, reinterpret_cast<cudaStream_t>(commandBuffer)) !=
      cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_UNKNOWN;
  }
  return VK_SUCCESS;
}

VkResult vrdxCudaCmdSignalSemaphore(VkCommandBuffer commandBuffer, VrdxCudaImportedSemaphore semaphore,
                                    uint64_t value) {
  if (!semaphore) return VK_ERROR_INITIALIZATION_FAILED;
  DeviceGuard guard(semaphore->device);
  cudaExternalSemaphoreSignalParams sp{};
  sp.params.fence.value = semaphore->timeline ? value : 0;
  if (cudaSignalExternalSemaphoresAsync(&semaphore->ext, &sp, 1, reinterpret_cast<cudaStream_t>(commandBuffer)) !=
      cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_UNKNOWN;
  }
  return VK_SUCCESS;
}

void vrdxCudaReleaseImportedSemaphore(VrdxCudaImportedSemaphore semaphore) {
  if (!semaphore) return;
  DeviceGuard guard(semaphore->device);
  cudaDestroyExternalSemaphore(semaphore->ext);
  delete semaphore;
}

void vrdxCudaGetSorterProperties(VrdxSorter sorter, VrdxCudaSorterProperties* p) {
  if (!sorter || !p) return;
  p->deviceOrdinal = sorter->device;
  p->smCount = sorter->sm_count;
  p->ccMajor = sorter->cc_major;
  p->ccMinor = sorter->cc_minor;
  p->keysTileSize = kKeysShapes[sorter->keys_shape].tile;
  p->keyValueTileSize = kPairShapes[sorter->pair_shape].tile;
  p->offsetAlignment = kOffsetAlignment;
  p->maxOnesweepCount = (uint32_t)kMaxOnesweepCount;
}

// ---------------------------------------------------------------------------- multi-GPU building blocks

void vrdxDistCmdPrefixHistogram(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                                VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t shift, uint32_t digitBits,
                                uint32_t prefixCount, VkBuffer prefixesBuffer, VkDeviceSize prefixesOffset,
                                VkBuffer histogramBuffer, VkDeviceSize histogramOffset) {
  if (!sorter) return;
  const size_t smem = (size_t)prefixCount * ((size_t)1 << (digitBits & 31)) * sizeof(uint32_t);
  if (!keysBuffer || !histogramBuffer || prefixCount == 0 || prefixCount > (uint32_t)kDistMaxSplitters ||
      digitBits == 0 || digitBits > 12 || shift + digitBits > 32 || smem > kDistHistMaxSmemBytes ||
      (!prefixesBuffer && shift + digitBits < 32)) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  if (elementCount == 0) return;
  DeviceGuard guard(sorter->device);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(commandBuffer);
  const uint32_t* keys = reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(keysBuffer) + keysOffset);
  const uint32_t* prefixes =
      prefixesBuffer ? reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(prefixesBuffer) + prefixesOffset)
                     : reinterpret_cast<const uint32_t*>(keys);  // never read when the prefix is empty
  uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(histogramBuffer) + histogramOffset);
  // (the kernel was opted into kDistHistMaxSmemBytes of dynamic shared memory on this device by vrdxCudaCreateSorter)
  const uint64_t vec_blocks = CeilDiv((uint64_t)elementCount / 4 + 1, (uint64_t)kDistHistThreads);
  const uint64_t cap = (uint64_t)sorter->sm_count * (smem > 112 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4));
  const uint32_t grid = (uint32_t)(vec_blocks < cap ? vec_blocks : cap);
  DistPrefixHistogramKernel<<<grid, kDistHistThreads, smem, stream>>>(keys, elementCount, shift, digitBits,
                                                                      prefixCount, prefixes, hist);
  NoteError(sorter, cudaGetLastError());
}

void vrdxDistCmdClassCount(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                           VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                           VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer countsBuffer,
                           VkDeviceSize countsOffset) {
  if (!sorter) return;
  if (!keysBuffer || !countsBuffer || splitterCount > (uint32_t)kDistMaxSplitters ||
      (splitterCount && !splittersBuffer)) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  if (elementCount == 0) return;
  DeviceGuard guard(sorter->device);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(commandBuffer);
  const uint32_t* keys = reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(keysBuffer) + keysOffset);
  const uint32_t* splitters =
      splittersBuffer ? reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(splittersBuffer) + splittersOffset)
                      : keys;
  uint32_t* counts = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(countsBuffer) + countsOffset);
  const uint64_t blocks = CeilDiv((uint64_t)elementCount / 4 + 1, (uint64_t)kDistCountThreads * 4);
  const uint64_t cap = (uint64_t)sorter->sm_count * 3;  // 64 KB of lane-private counters per CTA
  DistClassCountKernel<<<(uint32_t)(blocks < cap ? (blocks ? blocks : 1) : cap), kDistCountThreads, kDistCountSmemBytes, stream>>>(
      keys, elementCount, splitterCount, splitters, counts);
  NoteError(sorter, cudaGetLastError());
}

namespace {
void EnqueuePartition(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount, VkBuffer keysBuffer,
                      VkDeviceSize keysOffset, uint32_t splitterCount, VkBuffer splittersBuffer,
                      VkDeviceSize splittersOffset, VkBuffer cursorsBuffer, VkDeviceSize cursorsOffset,
                      VkBuffer outBuffer, VkDeviceSize outOffset, uint32_t destCount, VkBuffer destTableBuffer,
                      VkDeviceSize destTableOffset) {
  if (!sorter) return;
  const bool scatter = destTableBuffer != nullptr;
  if (!keysBuffer || (!scatter && !outBuffer) || !cursorsBuffer || splitterCount > (uint32_t)kDistMaxSplitters ||
      (splitterCount && !splittersBuffer) || (scatter && (destCount == 0 || destCount > (uint32_t)kDistMaxDests))) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  if (elementCount == 0) return;
  DeviceGuard guard(sorter->device);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(commandBuffer);
  const uint32_t* keys = reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(keysBuffer) + keysOffset);
  const uint32_t* splitters =
      splittersBuffer ? reinterpret_cast<const uint32_t*>(reinterpret_cast<char*>(splittersBuffer) + splittersOffset)
                      : keys;  // never read when splitterCount == 0
  uint32_t* cursors = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(cursorsBuffer) + cursorsOffset);
  const uint32_t grid = (uint32_t)CeilDiv((uint64_t)elementCount, (uint64_t)kDistPartTile);
  if (scatter) {
    const char* table = reinterpret_cast<const char*>(destTableBuffer) + destTableOffset;
    const unsigned long long* ptrs = reinterpret_cast<const unsigned long long*>(table);
    const uint32_t* first_pos = reinterpret_cast<const uint32_t*>(table + sizeof(unsigned long long) * destCount);
    DistPartitionKernel<true><<<grid, kDistPartThreads, 0, stream>>>(keys, elementCount, splitterCount, splitters,
                                                                     cursors, nullptr, destCount, ptrs, first_pos);
  } else {
    uint32_t* out = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(outBuffer) + outOffset);
    DistPartitionKernel<false><<<grid, kDistPartThreads, 0, stream>>>(keys, elementCount, splitterCount, splitters,
                                                                      cursors, out, 0, nullptr, nullptr);
  }
  NoteError(sorter, cudaGetLastError());
}
}  // namespace

void vrdxDistCmdPartition(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                          VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                          VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer cursorsBuffer,
                          VkDeviceSize cursorsOffset, VkBuffer outBuffer, VkDeviceSize outOffset) {
  EnqueuePartition(commandBuffer, sorter, elementCount, keysBuffer, keysOffset, splitterCount, splittersBuffer,
                   splittersOffset, cursorsBuffer, cursorsOffset, outBuffer, outOffset, 0, nullptr, 0);
}

void vrdxDistCmdPartitionScatter(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                                 VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                                 VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer cursorsBuffer,
                                 VkDeviceSize cursorsOffset, uint32_t destCount, VkBuffer destTableBuffer,
                                 VkDeviceSize destTableOffset) {
  if (sorter && !destTableBuffer) {
    NoteError(sorter, cudaErrorInvalidValue);
    return;
  }
  EnqueuePartition(commandBuffer, sorter, elementCount, keysBuffer, keysOffset, splitterCount, splittersBuffer,
                   splittersOffset, cursorsBuffer, cursorsOffset, nullptr, 0, destCount, destTableBuffer,
                   destTableOffset);
}

VkResult vrdxDistAllocShared(VkDevice device, VkDeviceSize size, VkBuffer* pBuffer,
                             unsigned char handle[VRDX_DIST_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == VRDX_DIST_IPC_HANDLE_BYTES, "IPC handle size");
  if (!pBuffer || !handle || size == 0) return VK_ERROR_INITIALIZATION_FAILED;
  DeviceGuard guard(DeviceFromHandle(device));
  void* p = nullptr;
  if (cudaMalloc(&p, size) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_OUT_OF_DEVICE_MEMORY;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    return VK_ERROR_FEATURE_NOT_PRESENT;
  }
  std::memcpy(handle, &h, sizeof(h));
  *pBuffer = reinterpret_cast<VkBuffer>(p);
  return VK_SUCCESS;
}

void vrdxDistFreeShared(VkDevice device, VkBuffer buffer) {
  if (!buffer) return;
  DeviceGuard guard(DeviceFromHandle(device));
  cudaFree(reinterpret_cast<void*>(buffer));
}

VkResult vrdxDistOpenShared(VkDevice device, const unsigned char handle[VRDX_DIST_IPC_HANDLE_BYTES],
                            VkBuffer* pBuffer) {
  if (!pBuffer || !handle) return VK_ERROR_INITIALIZATION_FAILED;
  DeviceGuard guard(DeviceFromHandle(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return VK_ERROR_INITIALIZATION_FAILED;
  }
  *pBuffer = reinterpret_cast<VkBuffer>(p);
  return VK_SUCCESS;
}

void vrdxDistCloseShared(VkDevice device, VkBuffer buffer) {
  if (!buffer) return;
  DeviceGuard guard(DeviceFromHandle(device));
  cudaIpcCloseMemHandle(reinterpret_cast<void*>(buffer));
}

}  // extern "C"
