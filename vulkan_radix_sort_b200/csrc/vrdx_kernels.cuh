// vrdx_kernels.cuh — sm_100a device code of the 32-bit LSD radix sort (8-bit digits x 4 passes).
//
// What these kernels replace in the reference (jaesung-cs/vulkan_radix_sort v0.4.0):
//   src/shader/upsweep.slang:10-45    per-partition digit histogram, re-read of the keys every pass
//   src/shader/spine.slang:11-84      exclusive scan over partitions + global histogram scan
//   src/shader/downsweep.slang:41-224 stable rank, local reorder, scatter (keys / key-value)
// with a different decomposition:
//   ResetKernel + HistogramKernel[Private]   onesweep: ONE read of the keys builds all four 256-bin digit
//                     histograms; the last CTA to finish exclusive-scans them in place (no spine launch).
//   UpsweepKernel + SpineKernel  reduce-then-scan: per-tile digit prefixes (16-bit, inside a chunk of 8 tiles),
//                     chunk prefixes and the global digit offsets of one pass; keys-only sorts over all 32 bits
//                     also get one flag byte per tile (its keys are one / two runs below the digit).
//   PassKernel        one launch per pass, one tile per CTA: warp-level multi-split ranking (returning
//                     shared-memory atomicAdd on warp-private counters, collisions repaired exactly with a
//                     REDUX.OR bloom filter + MATCH.ANY on the few colliding lanes), tile-local reorder
//                     through shared memory, run-wise coalesced scatter; the tile's global offsets come
//                     from a single-pass decoupled look-back (MODE 0) or from the upsweep tables (MODE 1).
//                     Flagged tiles of a keys-only reduce-then-scan pass skip the warp-private ranking
//                     (TileBlockFree: one returning atomic per key on a block-wide cursor row).
//                     Keys and values are two separate arrays end to end, as in the reference.
// Stability: a warp owns 32*IPT consecutive keys and ranks them item by item, lane by lane, so
// (warp, item, lane) order == index order — the same argument as downsweep.slang:79-80.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vrdx_layout.h"

namespace vrdx {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t LdRelaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void StRelaxed(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t LdStream(const uint32_t* p) { return __ldcs(p); }
// Programmatic dependent launch (sm_90+): every kernel of a sort lets its successor start launching
// at once (GridDepLaunch) and waits for its predecessor's memory only where it first needs it
// (GridDepWait), so launch latency, CTA ramp-up and prologues overlap the previous kernel's tail.
// Both are no-ops when the kernel was launched without the programmatic-serialization attribute.
__device__ __forceinline__ void GridDepLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void GridDepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Device-written timestamps (the CUDA stand-in for vkCmdWriteTimestamp): nanoseconds of the
// global timer; a kernel's "end of stage" stamp is the maximum over its CTAs.
__device__ __forceinline__ unsigned long long GlobalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void StampEnd(unsigned long long* slot) {
  if (slot != nullptr && threadIdx.x == 0) atomicMax(slot, GlobalTimerNs());
}
// Start of a sort when a query pool is attached: CTA 0 of the FIRST kernel of the sort clears slots 1..14 and
// writes slot 0 before it does anything else (no extra launch; the reference records timestamp 0 at the top of
// gpuSort, h.in:364-366).  Stamps only grow (%globaltimer is monotonic and StampEnd is an atomicMax), so a CTA
// of the same kernel that finished before CTA 0 got here could at worst be re-stamped by a later one.
constexpr int kTimestampSlots = 15;
__device__ __forceinline__ void StampStart(unsigned long long* slots) {
  if (slots != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long now = GlobalTimerNs();
    for (int i = 1; i < kTimestampSlots; ++i) slots[i] = 0ull;
    slots[0] = now;
    __threadfence();
  }
}
// Onesweep's per-sort reset (header: histograms, tickets; pass-0 look-back cells) as a kernel, so that it can
// carry the start stamp: with a query pool a sort is one launch shorter than memset + stamp kernel.
__global__ void __launch_bounds__(256) ResetKernel(uint4* __restrict__ p, uint64_t n16, unsigned long long* slots) {
  StampStart(slots);
  GridDepLaunch();
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * 256)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ uint32_t LaneMaskLt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t WarpInclusiveScan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Element count of this sort: immediate value (vrdxCmdSort / vrdxCmdSortKeyValue) or the
// uint32 the caller left in device memory (…Indirect), clamped to maxElementCount.
// Reference: vkCmdUpdateBuffer / vkCmdCopyBuffer into the count slot, h.in:368-379.
__device__ __forceinline__ uint32_t ResolveCount(const uint32_t* indirect, uint32_t n_or_max) {
  if (indirect == nullptr) return n_or_max;
  uint32_t c = __ldg(indirect);
  return c < n_or_max ? c : n_or_max;
}

// Key codec (extension, absent from the reference: include/vrdx_cuda.h VrdxCudaSortKeyInfo).
// The passes always sort plain unsigned 32-bit words.  Other key types and descending order are
// order-preserving bijections onto such words, applied where the FIRST pass reads the caller's
// keys (KeyIn) and undone where the LAST pass stores them (KeyOut), so they cost no extra pass:
//   int32    flip the sign bit                        cmask = 0x80000000
//   float32  negative: flip all bits, else sign bit   cmask = 0x80000000, fmask = 0x7FFFFFFF
//   descending: complement the word                   dmask = 0xFFFFFFFF
// All-zero masks are the identity (the reference's uint32 ascending sort).
struct KeyCodec {
  uint32_t fmask, cmask, dmask;
};
__device__ __forceinline__ uint32_t KeyIn(uint32_t k, const KeyCodec c) {
  return k ^ ((uint32_t)((int32_t)k >> 31) & c.fmask) ^ (c.cmask ^ c.dmask);
}
__device__ __forceinline__ uint32_t KeyOut(uint32_t t, const KeyCodec c) {
  const uint32_t u = t ^ c.dmask;
  return u ^ ((uint32_t)(~(int32_t)u >> 31) & c.fmask) ^ c.cmask;
}
// Digit plan of one sort: pass p ranks by (word >> shift[p]) & mask[p].  The reference is
// passes = 4, shift = 0/8/16/24, mask = 0xFF; a bit sub-range [begin, end) uses fewer passes and a
// narrower last digit.
struct DigitPlan {
  uint32_t passes;
  uint32_t shift[kPasses];
  uint32_t mask[kPasses];
  KeyCodec codec;  // applied by whoever reads the caller's keys
};

// ------------------------------------------------------------------------------------------
// HistogramKernel — all four digit histograms in one pass over the keys + exclusive scan.
// Algorithmic traffic: 4 B/key read.  Grid: a multiple of the SM count (persistent, grid-stride
// over chunks of THREADS*16 keys, four 128-bit loads in flight per thread).
// ------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;
constexpr int kHistVecPerThread = 4;                                   // uint4 loads per thread per chunk
constexpr int kHistChunk = kHistThreads * kHistVecPerThread * 4;       // keys per chunk

// GENERIC = false is the reference's plan with everything folded at compile time (the hot default);
// GENERIC = true reads the plan of a VrdxCudaSortKeyInfo.
template <bool GENERIC>
__device__ __forceinline__ void HistCount(uint32_t (*sh)[kRadix], uint32_t raw, const DigitPlan& plan) {
  if (!GENERIC) {
    atomicAdd(&sh[0][raw & 0xFFu], 1u);
    atomicAdd(&sh[1][(raw >> 8) & 0xFFu], 1u);
    atomicAdd(&sh[2][(raw >> 16) & 0xFFu], 1u);
    atomicAdd(&sh[3][raw >> 24], 1u);
  } else {
    const uint32_t k = KeyIn(raw, plan.codec);
#pragma unroll
    for (int p = 0; p < kPasses; ++p)
      if ((uint32_t)p < plan.passes) atomicAdd(&sh[p][(k >> plan.shift[p]) & plan.mask[p]], 1u);
  }
}

template <bool GENERIC>
__global__ void __launch_bounds__(kHistThreads)
HistogramKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ indirect,
                uint32_t n_or_max, StorageHeader* __restrict__ hdr, unsigned long long* ts_end,
                const DigitPlan plan) {
  __shared__ uint32_t sh[kPasses][kRadix];
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x;
  GridDepLaunch();
  const uint32_t n = ResolveCount(indirect, n_or_max);

  for (int i = tid; i < kPasses * kRadix; i += kHistThreads) (&sh[0][0])[i] = 0;
  __syncthreads();

  // Peel to 16-byte alignment so the body can use 128-bit loads whatever the caller's offset.
  const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(keys) >> 2) & 3u);
  uint32_t head = mis ? 4u - mis : 0u;
  if (head > n) head = n;
  const uint4* __restrict__ body = reinterpret_cast<const uint4*>(keys + head);
  const uint64_t nvec = (uint64_t)(n - head) >> 2;
  const uint32_t tail_start = head + (uint32_t)(nvec << 2);

  constexpr uint64_t kVecPerChunk = (uint64_t)kHistThreads * kHistVecPerThread;
  for (uint64_t base = (uint64_t)blockIdx.x * kVecPerChunk; base < nvec;
       base += (uint64_t)gridDim.x * kVecPerChunk) {
    uint4 v[kHistVecPerThread];
    if (base + kVecPerChunk <= nvec) {
#pragma unroll
      for (int j = 0; j < kHistVecPerThread; ++j) v[j] = __ldcs(body + base + j * kHistThreads + tid);
#pragma unroll
      for (int j = 0; j < kHistVecPerThread; ++j) {
        HistCount<GENERIC>(sh, v[j].x, plan); HistCount<GENERIC>(sh, v[j].y, plan);
        HistCount<GENERIC>(sh, v[j].z, plan); HistCount<GENERIC>(sh, v[j].w, plan);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kHistVecPerThread; ++j) {
        uint64_t idx = base + (uint64_t)j * kHistThreads + tid;
        if (idx < nvec) {
          uint4 q = __ldcs(body + idx);
          HistCount<GENERIC>(sh, q.x, plan); HistCount<GENERIC>(sh, q.y, plan);
          HistCount<GENERIC>(sh, q.z, plan); HistCount<GENERIC>(sh, q.w, plan);
        }
      }
    }
  }
  if (blockIdx.x == 0) {  // unaligned head (< 4 keys) and the n % 4 tail
    if ((uint32_t)tid < head) HistCount<GENERIC>(sh, keys[tid], plan);
    const uint32_t t = tail_start + tid;
    if (tid < 4 && t < n) HistCount<GENERIC>(sh, keys[t], plan);
  }
  __syncthreads();

  // The header is zeroed by ResetKernel, the launch before this one: with programmatic dependent launch this
  // kernel has been counting the caller's keys while that one ran, and only waits for it here.
  GridDepWait();
  if (blockIdx.x == 0 && tid == 0) hdr->element_count[0] = n;
  uint32_t* gh = &hdr->global_hist[0][0];
  for (int i = tid; i < kPasses * kRadix; i += kHistThreads) {
    uint32_t c = (&sh[0][0])[i];
    if (c) atomicAdd(gh + i, c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&hdr->hist_blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  StampEnd(ts_end);  // the last CTA ends the stage (its scan below is a few hundred cycles)

  // Last CTA: exclusive scan of each 256-bin histogram in place (spine.slang:62-83 does this
  // once per pass in workgroup 0; here once per sort).  Warp p scans pass p, 8 bins per lane.
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < kPasses) {
    uint32_t* h = gh + warp * kRadix + lane * 8;
    uint32_t c[8], sum = 0;
    bool all_in_one = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = __ldcg(h + j); sum += c[j]; all_in_one |= (c[j] == n); }
    uint32_t excl = WarpInclusiveScan(sum, lane) - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = excl; excl += c[j]; }
    // one digit holds every key: this pass is the identity permutation (tiles will just copy)
    const uint32_t any = __ballot_sync(0xffffffffu, all_in_one && n != 0);
    if (lane == 0) hdr->pass_identity[warp] = any ? 1u : 0u;
  }
}

// HistogramKernelPrivate — the same contract as HistogramKernel, for larger inputs: every bin has 32
// lane-private copies laid out as bin*32 + lane, so the bank of a shared-memory atomic is the
// lane id and no two lanes of a warp ever conflict, whatever the digits (HistogramKernel spends
// 13 cycles per 32 keys on 4 conflicting atomics, profiles/r01_microbench_rank.txt; here 4).
// 4 x 256 x 32 counters = 128 KB of shared memory: one 1024-thread CTA per SM, grid-stride.
constexpr int kHistPrivThreads = 1024;
constexpr int kHistPrivChunk = kHistPrivThreads * kHistVecPerThread * 4;  // keys per CTA per iteration
constexpr size_t kHistPrivSmemBytes = (size_t)kPasses * kRadix * 32 * sizeof(uint32_t);

template <bool GENERIC>
__device__ __forceinline__ void HistCountPrivate(uint32_t* sh, uint32_t lane, uint32_t raw, const DigitPlan& plan) {
  if (!GENERIC) {
    atomicAdd(&sh[((0 * kRadix + (raw & 0xFFu)) << 5) + lane], 1u);
    atomicAdd(&sh[((1 * kRadix + ((raw >> 8) & 0xFFu)) << 5) + lane], 1u);
    atomicAdd(&sh[((2 * kRadix + ((raw >> 16) & 0xFFu)) << 5) + lane], 1u);
    atomicAdd(&sh[((3 * kRadix + (raw >> 24)) << 5) + lane], 1u);
  } else {
    const uint32_t k = KeyIn(raw, plan.codec);
#pragma unroll
    for (int p = 0; p < kPasses; ++p)
      if ((uint32_t)p < plan.passes)
        atomicAdd(&sh[((p * kRadix + ((k >> plan.shift[p]) & plan.mask[p])) << 5) + lane], 1u);
  }
}

template <bool GENERIC>
__global__ void __launch_bounds__(kHistPrivThreads, 1)
HistogramKernelPrivate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ indirect,
                       uint32_t n_or_max, StorageHeader* __restrict__ hdr, unsigned long long* ts_end,
                       const DigitPlan plan) {
  extern __shared__ __align__(16) uint32_t sh[];  // [4][256][32]
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  GridDepLaunch();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  {
    uint4* z = reinterpret_cast<uint4*>(sh);
#pragma unroll
    for (int j = 0; j < (int)(kPasses * kRadix * 32 / 4 / kHistPrivThreads); ++j)
      z[j * kHistPrivThreads + tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(keys) >> 2) & 3u);
  uint32_t head = mis ? 4u - mis : 0u;
  if (head > n) head = n;
  const uint4* __restrict__ body = reinterpret_cast<const uint4*>(keys + head);
  const uint64_t nvec = (uint64_t)(n - head) >> 2;
  const uint32_t tail_start = head + (uint32_t)(nvec << 2);
  constexpr uint64_t kVecPerChunk = (uint64_t)kHistPrivThreads * kHistVecPerThread;
  for (uint64_t base = (uint64_t)blockIdx.x * kVecPerChunk; base < nvec; base += (uint64_t)gridDim.x * kVecPerChunk) {
    uint4 v[kHistVecPerThread];
#pragma unroll
    for (int j = 0; j < kHistVecPerThread; ++j) {
      const uint64_t idx = base + (uint64_t)j * kHistPrivThreads + tid;
      v[j] = idx < nvec ? __ldcs(body + idx) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < kHistVecPerThread; ++j) {
      if (base + (uint64_t)j * kHistPrivThreads + tid < nvec) {
        HistCountPrivate<GENERIC>(sh, lane, v[j].x, plan); HistCountPrivate<GENERIC>(sh, lane, v[j].y, plan);
        HistCountPrivate<GENERIC>(sh, lane, v[j].z, plan); HistCountPrivate<GENERIC>(sh, lane, v[j].w, plan);
      }
    }
  }
  if (blockIdx.x == 0) {  // unaligned head (< 4 keys) and the n % 4 tail
    if ((uint32_t)tid < head) HistCountPrivate<GENERIC>(sh, lane, keys[tid], plan);
    const uint32_t t = tail_start + tid;
    if (tid < 4 && t < n) HistCountPrivate<GENERIC>(sh, lane, keys[t], plan);
  }
  __syncthreads();

  // one bin per thread: sum its 32 lane copies (rotated start, so the 32 threads of a warp read
  // 32 different banks), then one global atomic per non-empty bin
  GridDepWait();  // the header is zeroed by the launch before this one (see HistogramKernel)
  if (blockIdx.x == 0 && tid == 0) hdr->element_count[0] = n;
  uint32_t* gh = &hdr->global_hist[0][0];
  {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) c += sh[(tid << 5) + ((lane + j) & 31)];
    if (c) atomicAdd(gh + tid, c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&hdr->hist_blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  StampEnd(ts_end);
  const int warp = tid >> 5;
  if (warp < kPasses) {
    uint32_t* h = gh + warp * kRadix + lane * 8;
    uint32_t c[8], sum = 0;
    bool all_in_one = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = __ldcg(h + j); sum += c[j]; all_in_one |= (c[j] == n); }
    uint32_t excl = WarpInclusiveScan(sum, lane) - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = excl; excl += c[j]; }
    const uint32_t any = __ballot_sync(0xffffffffu, all_in_one && n != 0);
    if (lane == 0) hdr->pass_identity[warp] = any ? 1u : 0u;
  }
}

// ------------------------------------------------------------------------------------------
// Arguments of one pass (the kernel ABI; reference: the descriptor bindings b0..b6 + the `pass` push
// constant, h.in:403-440).
// Algorithmic traffic per pass: 4 B/key read + 4 B/key write (+ 4 + 4 for values).
//
// Ranking primitives measured on B200 (tools/microbench_rank.cu; cycles per 32 keys per SM, 32 warps/SM):
//   hardware MATCH.ANY on an 8-bit digit, all lanes                        60.7  (1.7 per distinct value)
//   8-round ballot loop (reference downsweep.slang:92-99; CUB's choice)    28.8
//   shared-memory atomicOr peer mask + counter cell                        16.5
//   shared-memory atomicAdd (returning) on a warp-private counter           4.0
// So the rank of a key among equal digits of its warp is computed OPTIMISTICALLY: every lane does one
// returning atomicAdd on the warp-private counter of its digit.  A lane that is alone with its digit in this
// warp-instruction (89 % of lanes on uniform keys) gets its rank straight from the returned value.  Lanes
// that collided are served by the hardware in an unspecified order, so they are detected (counter
// read-back) and REPAIRED in lane order (RepairRank below).  Ranks are therefore exactly those of a stable
// counting sort whatever order the hardware serialises colliding atomics in.
// ------------------------------------------------------------------------------------------
struct PassArgs {
  const uint32_t* indirect;   // device count or nullptr
  uint32_t n_or_max;          // elementCount (direct) or maxElementCount (indirect)
  uint32_t pass;              // 0..3
  StorageHeader* hdr;
  uint32_t* status;           // look-back cells of this pass: [tile][256]
  uint32_t* status_next;      // onesweep: cells of the next pass, cleared here (nullptr on the last pass);
                              // reduce-then-scan: scanned chunk prefixes [chunk][256]
  const uint32_t* keys_in;
  uint32_t* keys_out;
  const uint32_t* vals_in;
  uint32_t* vals_out;
  unsigned long long* ts_end;  // query-pool slot stamped when this kernel finishes (or nullptr)
  unsigned long long* ts_start;  // first kernel of a sort with a query pool: slot 0 (StampStart), else nullptr
  // digit of this pass: (word >> shift) & mask (the reference: shift = 8 * pass, mask = 0xFF)
  uint32_t shift, mask;
  KeyCodec codec_in;    // non-zero only on the first pass: caller's key type/order -> sortable word
  KeyCodec codec_out;   // non-zero only on the last pass: sortable word -> caller's key type/order
  uint32_t order_free;  // 1: keys-only first pass of a sort over all 32 bits (no order to preserve)
  uint32_t static_tiles; // onesweep: 1 = the whole grid is co-resident, tile id = blockIdx.x (no ticket round trip)
  uint32_t range_tiles; // RangePassKernel / UpsweepRangeKernel (VRDX_EXPERIMENTS): consecutive tiles per CTA
  uint32_t words_only;  // 1: keys-only sort over all 32 bits (any pass): equal words are indistinguishable, so
                        //    keys that agree in every bit below this pass's digit may swap places (PassKernel)
  uint8_t* tile_flags;  // reduce-then-scan, keys-only: tile_flags[tile] = 1 when all keys of a (full) tile agree in the
                        //    bits below this pass's digit — written by UpsweepKernel, read by PassKernel<.., 1>
                        //    (TileBlockFree); nullptr: not kept (key-value sorts)
  uint32_t two_runs;    // 1: this pass runs the kernel flavours that also detect / take tiles of exactly two runs
};

constexpr uint8_t kTileOneRun = 1, kTileTwoRuns = 2;  // PassArgs::tile_flags
constexpr int kSpineChunk = (int)kSpineChunkTiles;  // reduce-then-scan: tiles per upsweep CTA / spine chunk
constexpr int kRepairBallotThreshold = 12;  // colliding lanes above which the 8-round ballot loop is cheaper

template <int THREADS, int IPT, bool KV, int MIN_CTAS, int LOOK_BATCH = 4, bool PAIRED = false>
struct PassConfig {
  static constexpr int kLookBatch = LOOK_BATCH;  // look-back cells fetched per round trip
  static constexpr bool kPaired = KV && PAIRED;  // key-value: stage (key, value) as one 64-bit shared-memory element
  static constexpr int kThreads = THREADS;
  static constexpr int kItems = IPT;
  static constexpr int kMinCtas = MIN_CTAS;
  static constexpr bool kKeyValue = KV;
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kTile = THREADS * IPT;
  static constexpr int kMiscWords = 16;
  // cnt[kWarps][256] | keys[kTile] | vals[kTile] (KV) | gbase[256] | misc
  static constexpr size_t kSmemBytes =
      sizeof(uint32_t) * ((size_t)kWarps * kRadix + (size_t)kTile * (KV ? 2 : 1) + kRadix + kMiscWords);
  static_assert(THREADS % 32 == 0 && THREADS >= kRadix && THREADS <= 1024, "one thread per digit is assumed");
  static_assert(kTile <= (1 << 14), "byte offsets of tile slots are kept below 2^16");
};

// Decoupled look-back for one digit: exclusive prefix over tiles [0, tile).  kLookBatch cells are
// in flight per round trip; `look_s` holds the first batch (tiles tile-1 .. tile-kLookBatch),
// already loaded by the caller.  Cells are consumed strictly nearest-first and the walk stops at
// the first inclusive prefix (tile 0 always publishes one, so the walk never underflows).
#ifndef VRDX_LOOK_WIDE
#define VRDX_LOOK_WIDE 4  // cells per round trip after the first (prefetched) batch; 12 and 24 measured slower
#endif
// One round trip of the walk: W cells in flight, consumed strictly nearest-first.
// (A 32-cell "burst" round for sorts whose tiles are all co-resident — where a tile walks back over every
// predecessor — was measured and changed nothing: small sorts are bound by the launch / drain latency of
// their six dependent kernels, not by the walk; profiles/r02/e_small_n_shapes_burst_*.txt.)
template <int W>
__device__ __forceinline__ void LookBackRound(const uint32_t* status, int digit, uint32_t& look, uint32_t& excl, bool& done,
                                              uint32_t& st_cells, uint32_t& st_notready) {
  uint32_t w[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const uint32_t t = (look >= (uint32_t)j) ? look - j : 0u;
    w[j] = LdRelaxed(status + (size_t)t * kRadix + digit);
  }
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (done) break;
    const uint32_t s = w[j];
    if ((s >> 30) == 0u) {
      ++st_notready;
      break;
    }
    excl += s & kStatusValueMask;
    ++st_cells;
    if (s & kStatusPrefix) { done = true; break; }
    --look;
  }
}

template <int kLookBatch>
__device__ __forceinline__ uint32_t LookBack(const uint32_t* status, uint32_t tile, int digit,
                                             uint32_t (&look_s)[kLookBatch], uint32_t* stats = nullptr) {
  constexpr int kWide = VRDX_LOOK_WIDE > kLookBatch ? VRDX_LOOK_WIDE : kLookBatch;
  uint32_t excl = 0;
  uint32_t look = tile - 1;  // nearest tile not yet consumed
  bool done = false;
  uint32_t st_rounds = 1, st_cells = 0, st_notready = 0;
  // first batch: the cells prefetched by the caller before the reorder
#pragma unroll
  for (int j = 0; j < kLookBatch; ++j) {
    if (done) break;
    const uint32_t s = look_s[j];
    if ((s >> 30) == 0u) break;  // not published yet: re-poll from `look`
    excl += s & kStatusValueMask;
    ++st_cells;
    if (s & kStatusPrefix) { done = true; break; }
    --look;
  }
  // later rounds: the walk is ~20 cells deep at full speed (profiles/r01_lookback_depth_stats.txt)
  while (!done) {
    ++st_rounds;
    LookBackRound<kWide>(status, digit, look, excl, done, st_cells, st_notready);
  }
#ifdef VRDX_STATS
  if (stats != nullptr && digit == 0) {  // one sample per tile (digit 0)
    atomicAdd(stats + 0, st_rounds);
    atomicAdd(stats + 1, st_cells);
    atomicAdd(stats + 2, st_notready);
  }
#endif
  (void)st_rounds;
  return excl;
}

// ------------------------------------------------------------------------------------------
// PassKernel — the tile kernel of round 2: the same algorithm as round 1's OnesweepKernel (one LSD pass over
// one tile per CTA: load, warp-level multi-split ranking, per-digit scan, tile-local reorder
// through shared memory, run-wise coalesced scatter), rewritten around what the ncu source view
// of the round-1 kernel showed (profiles/r02/a_ncu_source_ops_*): 78 warp-instructions and 20.5 shared-memory
// wavefronts per 32 keys, the instruction count being the first limiter.
//   * the digit of the reference plan is one PRMT (byte extract, selector in a register) instead of
//     shift + mask + scale (the old kernel spent 4-5 instructions per digit use, three uses per key).
//   * counters, ranks and slot bases are kept in BYTES (a key adds 4): the shared-memory address
//     of a key's slot is rank + base + constant, no shift.
//   * full tiles take branch-free load and scatter loops (the guard is hoisted out).
//   * collision repair (template RANK): hardware MATCH.ANY over the few lanes a REDUX.OR bloom
//     filter of the colliding digits selects, instead of a shuffle + ballot loop per collision
//     group; its cost is per distinct value among the PARTICIPATING lanes, so it also replaces
//     the 8-round ballot fallback for low-entropy digits.
//   * the counter read-back / repair of item i overlaps the atomic of item i + 1 (software
//     pipelining in source order; the warp-collective operations keep that order).
//   * reduce-then-scan: the upsweep leaves, per tile, the exclusive prefix INSIDE its chunk, so a
//     tile fetches 2 words per digit instead of 8.
// MODE 0: onesweep (tickets + decoupled look-back)   MODE 1: reduce-then-scan scatter pass.
// GENERIC = false: the reference's digit plan (shift = 8 * pass, mask = 0xFF, identity codec);
// GENERIC = true: digit and codec from PassArgs (vrdxCudaCmdSortEx).
// TWO (MODE 1, keys-only): tiles the upsweep flagged kTileTwoRuns take TileBlockFree as well.  A separate
// instantiation, launched only for passes in which such tiles are likely (EnqueueSort): with both block-free
// flavours compiled in, the one-run tiles of the other passes ran 7 % slower (profiles/r02/q_block_free_tiles.txt).
// ------------------------------------------------------------------------------------------
#ifndef VRDX_RANK_PIPELINE
#define VRDX_RANK_PIPELINE 1  // 1: the read-back / repair of item i is issued after the atomic of item i + 1
#endif
#ifndef VRDX_RANK_SYNCWARP
#define VRDX_RANK_SYNCWARP 1  // 1: __syncwarp() between a warp's atomics and the counter read-back; 0: compiler fence only
#endif
#ifndef VRDX_SPINE_FUSED
#define VRDX_SPINE_FUSED 1  // 1: SpineKernel (one launch per pass); 0: SpineReduceKernel + SpineApplyKernel
#endif
#ifndef VRDX_COUNT_FIRST
#define VRDX_COUNT_FIRST 0  // 1: onesweep tiles count their digits and publish the aggregate before ranking (A/B)
#endif
#ifndef VRDX_BLOCK_FREE_LDG128
#define VRDX_BLOCK_FREE_LDG128 1  // 1: block-free tiles load their keys with 128-bit loads
#endif
#ifndef VRDX_BLOCK_FREE
#define VRDX_BLOCK_FREE 1  // 1: reduce-then-scan keys-only tiles the upsweep flagged take TileBlockFree
#endif
#ifndef VRDX_RANK
#define VRDX_RANK 1  // 0: shuffle + ballot loop per collision group; 1: bloom + MATCH.ANY; 2: REDUX.MIN loop
#endif

// Digit of a key.  Reference plan: byte `pass` of the word, one PRMT with the selector 0x4440 | pass
// (bytes 1..3 of the result come from the zero operand).  Generic plan: (k >> shift) & mask.
// Orders a warp's shared-memory atomics of one item before its counter read-back, and that before the
// atomics of the next item.
__device__ __forceinline__ void RankFence() {
#if VRDX_RANK_SYNCWARP
  __syncwarp();
#else
  asm volatile("" ::: "memory");
#endif
}

template <bool GENERIC>
__device__ __forceinline__ uint32_t DigitOf(uint32_t k, uint32_t shift_or_sel, uint32_t mask) {
  if (!GENERIC) return __byte_perm(k, 0u, shift_or_sel);
  return (k >> shift_or_sel) & mask;
}

// Final rank (in bytes) of a key among the keys of its warp with the same digit: `old` is what the
// returning atomicAdd(+4) gave this lane, `fin` the counter after all lanes of this item were
// served.  ge = lanes >= this one.  A lane that was alone keeps old (fin - old == 4).
template <int RANK>
__device__ __forceinline__ uint32_t RepairRank(uint32_t d, uint32_t old, uint32_t fin, uint32_t ge) {
  uint32_t r = old;
  const bool flagged = fin - old > 4u;  // somebody was served after me: I am in a collision group
  if (RANK == 1) {
    // Every member of a group but the one served last is flagged.  A 32-bit bloom filter of the
    // flagged digits (one REDUX.OR) finds that one too, plus a few false positives, and
    // MATCH.ANY among those lanes only gives each its exact peers.
    const uint32_t bit = 1u << (d & 31u);
    const uint32_t bloom = __reduce_or_sync(0xffffffffu, flagged ? bit : 0u);
    const bool cand = (bloom & bit) != 0u;
    const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
    if (cand) {
      const uint32_t peers = __match_any_sync(cmask, d);
      r = fin - 4u * (uint32_t)__popc(peers & ge);  // a false positive has peers == itself: fin - 4 == old
    }
  } else {
    uint32_t suspects = __ballot_sync(0xffffffffu, flagged);
    if (suspects != 0u) {
      if (__popc(suspects) > kRepairBallotThreshold) {
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < kRadixBits; ++b) {
          const bool bt = (d >> b) & 1u;
          const uint32_t m = __ballot_sync(0xffffffffu, bt);
          peers &= bt ? m : ~m;
        }
        r = fin - 4u * (uint32_t)__popc(peers & ge);
      } else if (RANK == 2) {
        uint32_t mine = flagged ? d : 0xffffffffu;
        do {
          const uint32_t dstar = __reduce_min_sync(0xffffffffu, mine);
          const uint32_t peers = __ballot_sync(0xffffffffu, d == dstar);
          if (d == dstar) {
            r = fin - 4u * (uint32_t)__popc(peers & ge);
            mine = 0xffffffffu;
          }
          suspects &= ~peers;
        } while (suspects != 0u);
      } else {
        do {
          const uint32_t dstar = __shfl_sync(0xffffffffu, d, __ffs(suspects) - 1);
          const uint32_t peers = __ballot_sync(0xffffffffu, d == dstar);
          if (d == dstar) r = fin - 4u * (uint32_t)__popc(peers & ge);
          suspects &= ~peers;
        } while (suspects != 0u);
      }
    }
  }
  return r;
}

// ---- the phases of one tile, shared by PassKernel (one tile per CTA) and RangePassKernel (a
// ---- persistent CTA walking a contiguous range of tiles) ---------------------------------------

// Shared-memory carve-up of a tile: cnt[kWarps][256] | keys[kTile] | vals[kTile] (KV) | gbase[256] | misc
template <class Cfg>
struct TileSmem {
  uint32_t *cnt, *keys, *vals, *gbase, *misc;
  __device__ __forceinline__ explicit TileSmem(uint32_t* base) {
    cnt = base;                                              // warp-private digit counters (bytes), later slot bases
    keys = cnt + Cfg::kWarps * kRadix;                       // tile reordered by digit
    vals = keys + Cfg::kTile;                                // (KV only)
    gbase = vals + (Cfg::kKeyValue ? Cfg::kTile : 0);        // global slot of tile-local slot 0, per digit
    misc = gbase + kRadix;                                   // [0..7] warp totals, [8] tile id
  }
};

// Digit plan and codecs of a pass as the tile phases use them.
template <bool GENERIC>
struct PassDigit {
  uint32_t shift, mask, lowmask;  // shift: PRMT selector (reference plan) or bit shift (generic)
  KeyCodec cin, cout;
};

// load: warp-striped, 128 B per warp-instruction.  Tail tile: pads are the largest word, rank after
// every real key and are never stored (the reference pads the same way, downsweep.slang:81,85).
template <class Cfg, bool GENERIC>
__device__ __forceinline__ void TileLoadKeys(uint32_t (&key)[Cfg::kItems], const uint32_t* keys_in, uint64_t tile_start,
                                             uint32_t tile_count, uint32_t woff, const PassDigit<GENERIC>& dg) {
  constexpr int IPT = Cfg::kItems;
  const uint32_t* kin = keys_in + tile_start + woff;
  if (tile_count == (uint32_t)Cfg::kTile) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) key[i] = KeyIn(LdStream(kin + 32 * i), dg.cin);
  } else {
#pragma unroll
    for (int i = 0; i < IPT; ++i)
      key[i] = (woff + 32 * i < tile_count) ? KeyIn(LdStream(kin + 32 * i), dg.cin) : 0xFFFFFFFFu;
  }
}

// warp-level multi-split: rank (bytes) of each key among equal digits inside its warp, two ranks per
// register (they stay below 2^16 even after the slot base is added: kTile <= 2^14).
template <class Cfg, bool GENERIC, int RANK>
__device__ __forceinline__ void TileRank(const uint32_t (&key)[Cfg::kItems], uint32_t (&rank2)[Cfg::kItems / 2],
                                         uint32_t* row, bool full, const PassDigit<GENERIC>& dg) {
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  // Which order must equal digits keep?  The input of pass p is sorted by the bits below digit p, and the
  // output must be sorted by those bits within every digit.  Keys-only, all 32 bits compared: two keys
  // that agree in all the lower bits may swap places without changing ANY later result (equal words are
  // indistinguishable at the end).  So when all 32*IPT keys of this warp agree in the bits below the
  // digit — always in pass 0, and in passes 1 and 2 of a large sort almost always, because a warp's
  // 512 consecutive keys lie inside one run of the previous passes' order — any bijective ranking
  // inside (warp, digit) will do, and the value returned by the atomic is one: no read-back, no
  // repair.  Everything else (key-value sorts, bit sub-ranges, warps that straddle a run boundary,
  // the tail tile whose pads must rank last) is ranked stably below.
  bool relaxed = false;
  if (!KV && full && dg.lowmask != 0xFFFFFFFFu) {
    uint32_t diff = 0;
#pragma unroll
    for (int i = 1; i < IPT; ++i) diff |= key[i] ^ key[0];
    int same = 0;
    __match_all_sync(0xffffffffu, key[0] & dg.lowmask, &same);
    relaxed = __all_sync(0xffffffffu, (diff & dg.lowmask) == 0u) && same != 0;
  }
  if (relaxed) {
#pragma unroll
    for (int i = 0; i < IPT; i += 2) {
      const uint32_t r0 = atomicAdd(row + DigitOf<GENERIC>(key[i], dg.shift, dg.mask), 4u);
      const uint32_t r1 = atomicAdd(row + DigitOf<GENERIC>(key[i + 1], dg.shift, dg.mask), 4u);
      rank2[i / 2] = __byte_perm(r0, r1, 0x5410);
    }
  } else {
    const uint32_t ge = ~LaneMaskLt();
#if VRDX_RANK_PIPELINE
    uint32_t d_prev = 0, old_prev = 0, fin_prev = 0, r_even = 0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = DigitOf<GENERIC>(key[i], dg.shift, dg.mask);
      volatile uint32_t* c = row + d;
      const uint32_t old = atomicAdd(row + d, 4u);  // optimistic: exact if no other lane holds digit d
      RankFence();
      const uint32_t fin = *c;                      // all 32 increments of this item have landed
      RankFence();
      if (i > 0) {  // repair of the previous item overlaps the round trip above
        const uint32_t r = RepairRank<RANK>(d_prev, old_prev, fin_prev, ge);
        if ((i - 1) & 1) rank2[(i - 1) / 2] = __byte_perm(r_even, r, 0x5410); else r_even = r;
      }
      d_prev = d; old_prev = old; fin_prev = fin;
    }
    rank2[IPT / 2 - 1] = __byte_perm(r_even, RepairRank<RANK>(d_prev, old_prev, fin_prev, ge), 0x5410);
#else
    uint32_t r_even = 0;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = DigitOf<GENERIC>(key[i], dg.shift, dg.mask);
      volatile uint32_t* c = row + d;
      const uint32_t old = atomicAdd(row + d, 4u);  // optimistic: exact if no other lane holds digit d
      RankFence();
      const uint32_t fin = *c;                      // all 32 increments of this item have landed
      const uint32_t r = RepairRank<RANK>(d, old, fin, ge);
      RankFence();
      if (i & 1) rank2[i / 2] = __byte_perm(r_even, r, 0x5410); else r_even = r;
    }
#endif
  }
}

// per-digit, first half (threads < 256, between two barriers): counts over warps -> this tile's count of
// digit `tid` (pads removed) and its exclusive prefix inside the digit thread's warp (bytes).
template <class Cfg>
__device__ __forceinline__ void TileDigitSums(const TileSmem<Cfg>& sm, uint32_t (&wcount)[Cfg::kWarps], uint32_t& digit_count,
                                              uint32_t& digit_excl, uint32_t tile_count, uint32_t mask, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t sum = 0;
#pragma unroll
  for (int w = 0; w < Cfg::kWarps; ++w) {
    wcount[w] = sm.cnt[w * kRadix + tid];
    sum += wcount[w];
  }
  // pads were counted as the largest digit of this pass; they are not part of the data
  digit_count = (sum >> 2) - (((uint32_t)tid == mask) ? ((uint32_t)Cfg::kTile - tile_count) : 0u);
  const uint32_t incl = WarpInclusiveScan(sum, lane);
  if (lane == 31) sm.misc[warp] = incl;
  digit_excl = incl - sum;
}
// second half: tile-wide exclusive prefix of the digit (bytes); cnt becomes the tile-local byte offset of
// the first key of every (warp, digit).
template <class Cfg>
__device__ __forceinline__ void TileSlotBases(const TileSmem<Cfg>& sm, const uint32_t (&wcount)[Cfg::kWarps],
                                              uint32_t& digit_excl, int tid) {
  const int warp = tid >> 5;
#pragma unroll
  for (int w = 0; w < kRadix / 32; ++w) digit_excl += (w < warp) ? sm.misc[w] : 0u;
  uint32_t run = digit_excl;
#pragma unroll
  for (int w = 0; w < Cfg::kWarps; ++w) {
    sm.cnt[w * kRadix + tid] = run;
    run += wcount[w];
  }
}

// tile-local reorder through shared memory; values are fetched only now, so they do not occupy
// registers during the ranking
// Values of a tile, warp-striped like the keys.
template <class Cfg>
__device__ __forceinline__ void TileLoadValues(uint32_t (&val)[Cfg::kKeyValue ? Cfg::kItems : 1], const uint32_t* vals_in,
                                               uint64_t tile_start, uint32_t tile_count, uint32_t woff) {
  constexpr int IPT = Cfg::kItems;
  if (!Cfg::kKeyValue) return;
  const uint32_t* vin = vals_in + tile_start + woff;
  if (tile_count == (uint32_t)Cfg::kTile) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) val[Cfg::kKeyValue ? i : 0] = LdStream(vin + 32 * i);
  } else {
#pragma unroll
    for (int i = 0; i < IPT; ++i) val[Cfg::kKeyValue ? i : 0] = (woff + 32 * i < tile_count) ? LdStream(vin + 32 * i) : 0u;
  }
}

// The values are requested only at the reorder: asking for them right after the ranking (to hide their latency
// behind the digit scan) keeps 16 more registers live across two barriers and measured 4-8 % slower at every
// shape (profiles/r02/j_kv_early_values.txt).
// Staging them with 4-byte cp.async straight into their slots (no registers, no STS) was slower still: 6.04 vs
// 5.43 ms at 2^28 pairs (profiles/r02/o_kv_cp_async.txt).
constexpr bool kKvEarlyValues = false;

template <class Cfg, bool GENERIC>
__device__ __forceinline__ void TileReorder(const TileSmem<Cfg>& sm, const uint32_t (&key)[Cfg::kItems],
                                            const uint32_t (&rank2)[Cfg::kItems / 2], const uint32_t* row,
                                            const uint32_t* vals_in, uint64_t tile_start, uint32_t tile_count, uint32_t woff,
                                            const PassDigit<GENERIC>& dg,
                                            uint32_t (&val)[Cfg::kKeyValue ? Cfg::kItems : 1]) {
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  char* const keys_b = reinterpret_cast<char*>(sm.keys);
  uint32_t slot_b[KV ? IPT : 1];  // byte offset of each key's slot (kept for the values only)
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t r = (i & 1) ? (rank2[i / 2] >> 16) : (rank2[i / 2] & 0xFFFFu);
    const uint32_t sb = r + row[DigitOf<GENERIC>(key[i], dg.shift, dg.mask)];
    *reinterpret_cast<uint32_t*>(keys_b + sb) = key[i];
    if (KV) slot_b[i] = sb;
  }
  if (KV) {
    if (!kKvEarlyValues) TileLoadValues<Cfg>(val, vals_in, tile_start, tile_count, woff);
    char* const vals_b = reinterpret_cast<char*>(sm.vals);
#pragma unroll
    for (int i = 0; i < IPT; ++i) *reinterpret_cast<uint32_t*>(vals_b + slot_b[KV ? i : 0]) = val[KV ? i : 0];
  }
}

// scatter: consecutive threads write consecutive slots of a digit run
template <class Cfg, bool GENERIC>
__device__ __forceinline__ void TileScatter(const TileSmem<Cfg>& sm, uint32_t* keys_out, uint32_t* vals_out,
                                            uint32_t tile_count, int tid, const PassDigit<GENERIC>& dg) {
  constexpr int IPT = Cfg::kItems;
  constexpr int THREADS = Cfg::kThreads;
  constexpr bool KV = Cfg::kKeyValue;
  if (tile_count == (uint32_t)Cfg::kTile) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t slot = i * THREADS + tid;
      const uint32_t k = sm.keys[slot];
      const uint32_t g = sm.gbase[DigitOf<GENERIC>(k, dg.shift, dg.mask)] + slot;
      keys_out[g] = KeyOut(k, dg.cout);
      if (KV) vals_out[g] = sm.vals[slot];
    }
  } else {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t slot = i * THREADS + tid;
      if (slot < tile_count) {
        const uint32_t k = sm.keys[slot];
        const uint32_t g = sm.gbase[DigitOf<GENERIC>(k, dg.shift, dg.mask)] + slot;
        keys_out[g] = KeyOut(k, dg.cout);
        if (KV) vals_out[g] = sm.vals[slot];
      }
    }
  }
}

// constant digit: a stable counting sort with one non-empty bucket is a copy of [start, start + count)
template <class Cfg, bool GENERIC>
__device__ __forceinline__ void TileCopy(const PassArgs& a, uint64_t tile_start, uint32_t tile_count, int tid,
                                         const PassDigit<GENERIC>& dg) {
  constexpr int IPT = Cfg::kItems;
  constexpr int THREADS = Cfg::kThreads;
  constexpr bool KV = Cfg::kKeyValue;
  const uint32_t* kin = a.keys_in + tile_start;
  uint32_t* kout = a.keys_out + tile_start;
  // (pairs go in two halves: 2 x IPT words in flight per thread would not fit the register budget of the pass kernel
  // this is inlined into, and spilled)
  constexpr int kStep = KV ? (IPT + 1) / 2 : IPT;
#pragma unroll
  for (int i0 = 0; i0 < IPT; i0 += kStep) {
    uint32_t ck[kStep], cv[KV ? kStep : 1];
#pragma unroll
    for (int j = 0; j < kStep; ++j) {  // all loads first: the stores below may alias them as far as the compiler knows
      const uint32_t idx = (i0 + j) * THREADS + tid;
      const bool in = i0 + j < IPT && idx < tile_count;
      ck[j] = in ? KeyOut(KeyIn(LdStream(kin + idx), dg.cin), dg.cout) : 0u;
      if (KV) cv[j] = in ? LdStream(a.vals_in + tile_start + idx) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kStep; ++j) {
      const uint32_t idx = (i0 + j) * THREADS + tid;
      if (i0 + j < IPT && idx < tile_count) {
        kout[idx] = ck[j];
        if (KV) a.vals_out[tile_start + idx] = cv[j];
      }
    }
  }
}

// Reduce-then-scan, keys-only, a full tile whose keys all agree in the bits below the digit (always in pass 0;
// pass 1 of a large sort: a tile lies inside one run of pass 0's order): ANY bijection of the tile's keys onto the
// slots of their digit runs gives the same final result (TileRank explains why), and the tile's digit counts are
// already in the upsweep's table.  So the slot bases are computed first (while the keys are still in flight) and
// every key takes its final tile slot with ONE returning shared-memory atomicAdd on a block-wide cursor row and
// is stored there: 2 digit-indexed shared-memory accesses per key instead of 3 (atomic, slot base, store), no
// per-warp rows to sum, two barriers fewer.
// TWO (tile flag kTileTwoRuns): the tile is two runs of keys that agree below the digit (it straddles one boundary
// of the previous passes' order; `first` = the tile's first key).  Keys of the first run must precede keys of the
// second inside every digit: the first run takes its slots from the FRONT of the digit's range (ascending cursor),
// the second from the BACK (descending cursor) — the two meet exactly, because the range has as many slots as
// the digit has keys.
template <class Cfg, bool GENERIC, bool TWO>
__device__ __forceinline__ void TileBlockFree(const PassArgs& a, const TileSmem<Cfg>& sm, const uint32_t (&key)[Cfg::kItems],
                                              uint32_t tile, int tid, const PassDigit<GENERIC>& dg, uint32_t first) {
  constexpr int IPT = Cfg::kItems;
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t count = 0, incl = 0, gb = 0;
  if (tid < kRadix) {
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(a.status);
    const uint32_t before = (tile % kSpineChunk) ? t16[(size_t)(tile - 1) * kRadix + tid] : 0u;
    count = (t16[(size_t)tile * kRadix + tid] - before) & 0xFFFFu;
    gb = a.status_next[(size_t)(tile / kSpineChunk) * kRadix + tid] + before + a.hdr->global_hist[a.pass][tid];
    incl = WarpInclusiveScan(count, lane);
    if (lane == 31) sm.misc[warp] = incl;
  }
  __syncthreads();
  if (tid < kRadix) {
    uint32_t excl = incl - count;
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) excl += (w < warp) ? sm.misc[w] : 0u;
    sm.cnt[tid] = excl << 2;      // byte offset of the digit's next free tile slot
    if (TWO) sm.cnt[kRadix + tid] = (excl + count) << 2;  // ... and of the end of its range
    sm.gbase[tid] = gb - excl;    // global slot of tile-local slot 0 for this digit (mod 2^32 arithmetic)
  }
  __syncthreads();
  char* const keys_b = reinterpret_cast<char*>(sm.keys);
  uint32_t slot2[IPT / 2];
  auto claim = [&](uint32_t k) -> uint32_t {
    const uint32_t d = DigitOf<GENERIC>(k, dg.shift, dg.mask);
    if (!TWO) return atomicAdd(sm.cnt + d, 4u);
    const bool second = ((k ^ first) & dg.lowmask) != 0u;
    const uint32_t r = atomicAdd(sm.cnt + (second ? kRadix : 0) + d, second ? 0xFFFFFFFCu : 4u);
    return second ? r - 4u : r;
  };
#pragma unroll
  for (int i = 0; i < IPT; i += 2) {
    const uint32_t s0 = claim(key[i]);
    const uint32_t s1 = claim(key[i + 1]);
    slot2[i / 2] = __byte_perm(s0, s1, 0x5410);
  }
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t sb = (i & 1) ? (slot2[i / 2] >> 16) : (slot2[i / 2] & 0xFFFFu);
    *reinterpret_cast<uint32_t*>(keys_b + sb) = key[i];
  }
  __syncthreads();
}

template <class Cfg, bool GENERIC>
__device__ __forceinline__ PassDigit<GENERIC> MakePassDigit(const PassArgs& a, bool order_free) {
  PassDigit<GENERIC> dg;
  dg.shift = GENERIC ? a.shift : (0x4440u | a.pass);  // reference plan: the PRMT selector of byte `pass`
  dg.mask = GENERIC ? a.mask : (uint32_t)(kRadix - 1);
  dg.cin = GENERIC ? a.codec_in : KeyCodec{0u, 0u, 0u};
  dg.cout = GENERIC ? a.codec_out : KeyCodec{0u, 0u, 0u};
  // bits below this pass's digit (a keys-only sort over all 32 bits has contiguous digits from bit 0)
  dg.lowmask = (!Cfg::kKeyValue && a.words_only != 0u) ? ((1u << (GENERIC ? a.shift : a.pass * kRadixBits)) - 1u)
               : order_free                            ? 0u
                                                       : 0xFFFFFFFFu;
  return dg;
}

template <class Cfg, int MODE, bool GENERIC, int RANK = VRDX_RANK, bool TWO = false>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
PassKernel(const PassArgs a) {
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;
  constexpr int kLookBatch = Cfg::kLookBatch;
  static_assert(MODE == 0 || MODE == 1, "onesweep or reduce-then-scan (per-tile tables)");

  extern __shared__ __align__(128) uint32_t smem[];
  const TileSmem<Cfg> sm(smem);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t pass = a.pass;
  const uint32_t n = ResolveCount(a.indirect, a.n_or_max);
  GridDepLaunch();
  {
    uint4* z = reinterpret_cast<uint4*>(sm.cnt);
#pragma unroll
    for (int j = tid; j < kWarps * kRadix / 4; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
    if (MODE == 0 && VRDX_COUNT_FIRST && tid < kRadix / 4) reinterpret_cast<uint4*>(sm.gbase)[tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  GridDepWait();  // everything below reads what the previous kernel of this sort wrote
  // keys-only first pass of a sort over all 32 bits: no earlier order to preserve (see TileRank), so an
  // onesweep tile may also claim its output ranges with global atomics instead of tickets + look-back
  const bool order_free = !KV && a.order_free != 0u;
  const bool unordered = (MODE == 0) && order_free;
  const PassDigit<GENERIC> dg = MakePassDigit<Cfg, GENERIC>(a, order_free);
  // Tile ids are handed out in arrival order so that every predecessor a tile may wait on in the
  // look-back is already resident (forward progress without relying on blockIdx order).  When the whole
  // grid is co-resident anyway (small sorts: PassArgs::static_tiles) the ticket's round trip is skipped.
  const bool ticketed = MODE == 0 && !unordered && a.static_tiles == 0u;
  if (ticketed && tid == 0) sm.misc[8] = atomicAdd(&a.hdr->tickets[pass], 1u);
  // Onesweep (small and medium sorts, where the chain of dependent round trips inside a pass is what a sort
  // costs): the identity flag and the digit's global offset are requested now / with the keys and consumed
  // later.  Reduce-then-scan (large sorts) keeps them off the register file until they are needed: holding
  // them across the ranking cost 2.7 % at 2^28 keys (profiles/r02/l_early_loads_ab.txt).
  constexpr bool kEarly = MODE == 0;
  const uint32_t identity = kEarly ? a.hdr->pass_identity[pass] : 0u;
  __syncthreads();

  const uint32_t tile = ticketed ? sm.misc[8] : blockIdx.x;
  const uint64_t tile_start = (uint64_t)tile * kTile;
  if (tile_start >= n) return;  // indirect count below max: surplus CTAs retire (upsweep.slang:20-22)
  const uint32_t remaining = (uint32_t)(n - tile_start);
  const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;
  const bool full = tile_count == (uint32_t)kTile;

  if (MODE == 0 && a.status_next != nullptr && tid < kRadix) a.status_next[(size_t)tile * kRadix + tid] = 0;

  // reduce-then-scan, keys-only: did the upsweep find the keys of this tile to be one or two runs below the digit?
  // (TileBlockFree.)  The flag travels with the identity word: one round trip before the keys are requested.
  uint32_t block_free = 0;
  if (MODE == 1 && !KV && VRDX_BLOCK_FREE && a.tile_flags != nullptr) block_free = a.tile_flags[tile];
  if (!kEarly && a.hdr->pass_identity[pass]) {
    TileCopy<Cfg, GENERIC>(a, tile_start, tile_count, tid, dg);
    StampEnd(a.ts_end);
    return;
  }

  uint32_t key[IPT];
  const uint32_t woff = warp * 32 * IPT + lane;
  if (MODE == 1 && !KV && VRDX_BLOCK_FREE && (TWO ? block_free != 0u : block_free == kTileOneRun) && full) {
    // (its own copy of the key loads: the flag is tested BEFORE they are issued, so that the table loads of
    // TileBlockFree go out right behind them instead of waiting for the keys' scoreboard;
    // profiles/r02/q_block_free_tiles.txt)
#if VRDX_BLOCK_FREE_LDG128
    // The order of a block-free tile's keys does not matter, so a thread takes four consecutive keys per 128-bit
    // load (the survey's "128-bit loads + in-register transpose" without the transpose): same wavefronts, 15
    // load instructions fewer per thread; 2^28 keys 3.28 -> 3.22 ms (profiles/r02/r_ldg128_block_free.txt).
    if ((reinterpret_cast<uintptr_t>(a.keys_in) & 15u) == 0u && IPT % 4 == 0) {
      const uint4* k4 = reinterpret_cast<const uint4*>(a.keys_in + tile_start) + tid;
#pragma unroll
      for (int i = 0; i < IPT / 4; ++i) {
        const uint4 q = __ldcs(k4 + i * THREADS);
        key[4 * i + 0] = KeyIn(q.x, dg.cin);
        key[4 * i + 1] = KeyIn(q.y, dg.cin);
        key[4 * i + 2] = KeyIn(q.z, dg.cin);
        key[4 * i + 3] = KeyIn(q.w, dg.cin);
      }
    } else
#endif
    TileLoadKeys<Cfg, GENERIC>(key, a.keys_in, tile_start, tile_count, woff, dg);
    if (!TWO || block_free == kTileOneRun) TileBlockFree<Cfg, GENERIC, false>(a, sm, key, tile, tid, dg, 0u);
    else TileBlockFree<Cfg, GENERIC, true>(a, sm, key, tile, tid, dg, KeyIn(a.keys_in[tile_start], dg.cin));
    TileScatter<Cfg, GENERIC>(sm, a.keys_out, a.vals_out, tile_count, tid, dg);
    StampEnd(a.ts_end);
    return;
  }
  TileLoadKeys<Cfg, GENERIC>(key, a.keys_in, tile_start, tile_count, woff, dg);
  // global digit offset of this pass for digit `tid`: needed at the very end, requested with the keys
  uint32_t digit_base = 0;
  if (kEarly && tid < kRadix) digit_base = a.hdr->global_hist[pass][tid];

  if (kEarly && identity) {
    // constant digit: a stable counting sort with one non-empty bucket is a copy (same warp-striped indices out as in)
    uint32_t cv[KV ? IPT : 1];
    if (KV) TileLoadValues<Cfg>(cv, a.vals_in, tile_start, tile_count, woff);
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if (full || woff + 32 * i < tile_count) {
        a.keys_out[tile_start + woff + 32 * i] = KeyOut(key[i], dg.cout);
        if (KV) a.vals_out[tile_start + woff + 32 * i] = cv[KV ? i : 0];
      }
    }
    StampEnd(a.ts_end);
    return;
  }

#if VRDX_COUNT_FIRST
  // Onesweep, ordered passes: count the tile's digits BEFORE ranking (one non-returning shared-memory atomic per key
  // on a block-wide row) and publish the aggregate now, so that the successors' look-back — and this tile's own,
  // consumed after the reorder — overlaps the whole ranking instead of only the reorder.
  const bool count_first = MODE == 0 && !unordered;
  if (count_first) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) atomicAdd(sm.gbase + DigitOf<GENERIC>(key[i], dg.shift, dg.mask), 1u);
    __syncthreads();
    if (tid < kRadix) {
      const uint32_t c = sm.gbase[tid] - (((uint32_t)tid == dg.mask) ? ((uint32_t)kTile - tile_count) : 0u);
      StRelaxed(a.status + (size_t)tile * kRadix + tid, (tile == 0 ? kStatusPrefix : kStatusAggregate) | c);
    }
  }
#else
  constexpr bool count_first = false;
#endif
  uint32_t rank2[IPT / 2];
  uint32_t* const row = sm.cnt + warp * kRadix;
  TileRank<Cfg, GENERIC, RANK>(key, rank2, row, full, dg);
  uint32_t val[KV ? IPT : 1];
  if (kKvEarlyValues) TileLoadValues<Cfg>(val, a.vals_in, tile_start, tile_count, woff);
  __syncthreads();

  uint32_t digit_count = 0, digit_excl = 0;
  uint32_t wcount[kWarps];
  if (tid < kRadix) {
    TileDigitSums<Cfg>(sm, wcount, digit_count, digit_excl, tile_count, dg.mask, tid);
    if (MODE == 0 && !unordered && !count_first)
      StRelaxed(a.status + (size_t)tile * kRadix + tid,
                (tile == 0 ? kStatusPrefix : kStatusAggregate) | digit_count);
  }
  __syncthreads();
  uint32_t look_s[kLookBatch];
  if (tid < kRadix) {
    TileSlotBases<Cfg>(sm, wcount, digit_excl, tid);
    if (unordered) {
      // claim [excl, excl + digit_count) of this digit's global run; the round trip overlaps the reorder
      look_s[0] = digit_count ? atomicAdd(&a.hdr->claim_cursor[tid], digit_count) : 0u;
    } else if (MODE == 0) {
      // start the first batch of look-back loads now; it is consumed after the reorder below
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        const uint32_t t = (tile > (uint32_t)j) ? tile - 1 - j : 0u;
        look_s[j] = (tile > 0) ? LdRelaxed(a.status + (size_t)t * kRadix + tid) : 0u;
      }
    } else {
      // reduce-then-scan: scanned chunk prefix + this tile's exclusive prefix inside its chunk
      look_s[0] = a.status_next[(size_t)(tile / kSpineChunk) * kRadix + tid] +
                  ((tile % kSpineChunk) ? reinterpret_cast<const uint16_t*>(a.status)[(size_t)(tile - 1) * kRadix + tid] : 0u);
    }
  }
  __syncthreads();

  TileReorder<Cfg, GENERIC>(sm, key, rank2, row, a.vals_in, tile_start, tile_count, woff, dg, val);

  // ---- global offsets of the digit runs ----------------------------------------------------
  if (tid < kRadix) {
    uint32_t excl = 0;
    if (unordered || MODE == 1) {
      excl = look_s[0];
    } else if (tile > 0) {
      excl = LookBack<kLookBatch>(a.status, tile, tid, look_s, a.hdr->reserved);
      StRelaxed(a.status + (size_t)tile * kRadix + tid, kStatusPrefix | (excl + digit_count));
    }
    // global slot of tile-local slot 0 for this digit (mod 2^32 arithmetic)
    if (!kEarly) digit_base = a.hdr->global_hist[pass][tid];
    sm.gbase[tid] = digit_base + excl - (digit_excl >> 2);
  }
  __syncthreads();

  TileScatter<Cfg, GENERIC>(sm, a.keys_out, a.vals_out, tile_count, tid, dg);
  StampEnd(a.ts_end);
}

// ------------------------------------------------------------------------------------------
// Reduce-then-scan (the reference's upsweep / spine / downsweep decomposition, src/shader/upsweep.slang,
// spine.slang, downsweep.slang): what AUTO runs from 2^25 keys / 2^27 pairs up
// (profiles/r02/s_sweep_n_auto_thresholds.txt) and the only path for N >= 2^30 (full 32-bit counts, no flag
// bits).  The scatter pass has no inter-CTA dependency at all.  Per pass, three launches:
//   UpsweepKernel    4 B/key read (128-bit loads); one CTA counts a CHUNK of kSpineChunk consecutive tiles and
//                    writes the 16-bit inclusive prefixes tile_hist[tile][256], the chunk's column sums
//                    chunk_sums[chunk][256] and (keys-only) one flag byte per tile
//   SpineKernel      exclusive scan of chunk_sums over chunks (128 co-resident segments exchange their
//                    sums once) + global digit offsets and the identity flag (last segment)
//   PassKernel<Cfg, 1>  4 B/key read + 4 B/key write; a tile's offset for digit d is
//                    chunk_sums[chunk][d] + the tile_hist row of the tile before it in its chunk
// Algorithmic overhead vs onesweep: the keys are read twice per pass (12 instead of 8 B/key).
// ------------------------------------------------------------------------------------------
constexpr int kUpsweepThreads = 256;

// First CTA of an upsweep (after its GridDepWait: the previous pass has finished, this pass's spine has not begun):
// the spine's per-pass state.
__device__ __forceinline__ void ResetSpineState(StorageHeader* hdr, int tid) {
  if (blockIdx.x != 0) return;
  if (tid == 0) hdr->hist_blocks_done = 0;  // segment ticket of SpineKernel / last-block-done counter of the two-kernel spine
  if (tid < (int)kSpineSegmentRows) hdr->spine_flags[tid] = 0;
}

// INCL: tile_hist is a table of 16-bit words, tile_hist[tile][d] = number of digit-d keys in this tile and the
// EARLIER tiles of the same chunk (<= kSpineChunk * TILE < 2^16).  PassKernel<.., 1> adds the row of the tile
// before it (nothing for the first tile of a chunk) to the chunk prefix; the difference of the two rows is the
// tile's own digit count, known before the tile is ranked (TileBlockFree).  Half the bytes of the reference's
// partHist (h.in:353-362).  !INCL (round-1 kernels, VRDX_EXPERIMENTS): 32-bit words holding the tile's own count.
// tile_flags (keys-only sorts over all 32 bits, else nullptr): tile_flags[tile] = kTileOneRun when the tile is full
// and all its keys agree in the bits of `lowmask` (the bits below this pass's digit), kTileTwoRuns when they take
// two values there (the first key's, then the last key's).
template <int TILE, bool INCL = false, bool TWO = false>
__global__ void __launch_bounds__(kUpsweepThreads)
UpsweepKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, uint32_t shift, uint32_t mask,
              const KeyCodec codec_in, const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ tile_hist,
              uint32_t* __restrict__ chunk_sums, StorageHeader* __restrict__ hdr, unsigned long long* ts_end,
              unsigned long long* ts_start, uint8_t* __restrict__ tile_flags = nullptr, uint32_t lowmask = 0xFFFFFFFFu) {
  constexpr int THREADS = kUpsweepThreads;
  static_assert(THREADS == kRadix, "one thread per digit");
  // (the last row of a chunk may wrap at 2^16: it is only ever used in a difference mod 2^16, never as a prefix)
  static_assert(!INCL || (uint64_t)(kSpineChunk - 1) * TILE < 65536, "in-chunk prefixes are stored as 16-bit words");
  __shared__ uint32_t h[2][kRadix];
  const int tid = threadIdx.x;
  StampStart(ts_start);  // non-null only on the first kernel of a sort
  GridDepLaunch();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  const uint32_t tiles = (uint32_t)CeilDiv(n, (uint64_t)TILE);
  const uint32_t first = blockIdx.x * kSpineChunk;
  h[0][tid] = 0;
  h[1][tid] = 0;
  GridDepWait();
  ResetSpineState(hdr, tid);
  if (first >= tiles) return;
  const uint32_t last = first + kSpineChunk < tiles ? first + kSpineChunk : tiles;
  constexpr int kIters = (TILE + THREADS - 1) / THREADS;
  uint32_t chunk_acc = 0;
  __syncthreads();
  for (uint32_t tile = first; tile < last; ++tile) {
    uint32_t* hh = h[(tile - first) & 1];
    const uint64_t tile_start = (uint64_t)tile * TILE;
    const uint32_t remaining = (uint32_t)(n - tile_start);
    const uint32_t tile_count = remaining < (uint32_t)TILE ? remaining : (uint32_t)TILE;
    const uint32_t* kin = keys_in + tile_start;
    // Tile flag (keys-only sorts over all 32 bits; see the header comment).  The input of this pass is sorted by the
    // bits of lowmask (the invariant of an LSD sort; lowmask == 0 in the first pass), so the tile is ONE run iff
    // its first and last key agree there — two broadcast loads, nothing per key.  TWO runs need every key looked
    // at (does it agree with the first or with the last key?); that loop runs only when two more samples say so.
    int flag = 0;        // uniform over the CTA
    uint32_t stray = 0;  // non-zero: some key differs (below the digit) from both the first and the last key
    if (tile_count == (uint32_t)TILE) {
      // Counting does not care which thread sees which key: 128-bit loads (a thread takes four consecutive keys)
      // when the caller's pointer allows it.  (With 25 scalar loads per thread the compiler interleaved loads and
      // atomics to save registers and the 6400-key instantiation ran at 230 instead of 180 us per 2^28 keys.)
      constexpr int kVec = TILE / (4 * THREADS);          // full rounds of one uint4 per thread
      constexpr int kVecTail = (TILE % (4 * THREADS)) / 4;  // threads that take one more
      static_assert(TILE % 4 == 0, "tiles are whole uint4s");
      uint32_t k[4 * (kVec + 1) > kIters ? 4 * (kVec + 1) : kIters];
      int have = kIters;
      if ((reinterpret_cast<uintptr_t>(keys_in) & 15u) == 0u) {
        const uint4* k4 = reinterpret_cast<const uint4*>(kin);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
          const uint4 q = __ldcs(k4 + i * THREADS + tid);
          k[4 * i + 0] = q.x; k[4 * i + 1] = q.y; k[4 * i + 2] = q.z; k[4 * i + 3] = q.w;
        }
        have = 4 * kVec;
        if (kVecTail != 0 && tid < kVecTail) {
          const uint4 q = __ldcs(k4 + kVec * THREADS + tid);
          k[4 * kVec + 0] = q.x; k[4 * kVec + 1] = q.y; k[4 * kVec + 2] = q.z; k[4 * kVec + 3] = q.w;
          have = 4 * kVec + 4;
        }
      } else {
#pragma unroll
        for (int i = 0; i < kIters; ++i) k[i] = LdStream(kin + i * THREADS + tid);
      }
      uint32_t l0 = 0, ll = 0;
      bool maybe_two = false;
      if (tile_flags != nullptr) {
        l0 = kin[0] & lowmask;  // (flags are kept for passes whose codec_in is the identity, or with lowmask == 0)
        ll = kin[TILE - 1] & lowmask;
        flag = l0 == ll ? kTileOneRun : 0;
        if (TWO && l0 != ll) {
          const uint32_t la = kin[TILE / 3] & lowmask, lb = kin[2 * (TILE / 3)] & lowmask;
          maybe_two = (la == l0 || la == ll) && (lb == l0 || lb == ll) && !(la == ll && lb == l0);
        }
      }
      constexpr int kHeld = (int)(sizeof(k) / sizeof(k[0]));
      if (maybe_two) {
#pragma unroll
        for (int i = 0; i < kHeld; ++i) {
          if (i < have) {
            const uint32_t kk = KeyIn(k[i], codec_in);
            atomicAdd(&hh[(kk >> shift) & mask], 1u);
            stray |= min((kk & lowmask) ^ l0, (kk & lowmask) ^ ll);
          }
        }
        flag = kTileTwoRuns;  // unless a stray key turns up (barrier below)
      } else {
#pragma unroll
        for (int i = 0; i < kHeld; ++i)
          if (i < have) atomicAdd(&hh[(KeyIn(k[i], codec_in) >> shift) & mask], 1u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const uint32_t idx = i * THREADS + tid;
        if (idx < tile_count) atomicAdd(&hh[(KeyIn(LdStream(kin + idx), codec_in) >> shift) & mask], 1u);
      }
    }
    // one barrier per tile: the two histograms alternate
    const int no_stray = __syncthreads_and(stray == 0u);
    const uint32_t c = hh[tid];
    hh[tid] = 0;      // ready for tile + 2 (the next tile uses the other buffer; barrier above orders it)
    chunk_acc += c;
    if (INCL) reinterpret_cast<uint16_t*>(tile_hist)[(size_t)tile * kRadix + tid] = (uint16_t)chunk_acc;
    else tile_hist[(size_t)tile * kRadix + tid] = c;
    if (INCL && tile_flags != nullptr && tid == 0) tile_flags[tile] = (uint8_t)(no_stray ? flag : 0);
  }
  chunk_sums[(size_t)blockIdx.x * kRadix + tid] = chunk_acc;
  StampEnd(ts_end);
}

// Spine: exclusive scan of chunk_sums[chunk][256] over chunks for every digit, and the global
// digit offsets of this pass (spine.slang:32-60 and :62-83).  Thread = digit, so every row access is one
// 1 KB line.  SpineKernel (below, what ships) does it in one launch; the two-kernel form it replaced is kept
// behind VRDX_SPINE_FUSED=0 for A/B:
//   SpineReduceKernel  CTA s sums the rows of segment s -> seg[s][256]; the LAST CTA to finish
//                      scans seg over segments in place and turns the per-digit totals into the
//                      global digit offsets hdr->global_hist[pass]
//   SpineApplyKernel   CTA s rewrites the rows of segment s as exclusive prefixes
constexpr int kSpineSegments = (int)kSpineSegmentRows;

// The grid (number of segments) is sized by the host from maxElementCount; the rows are split
// evenly over however many segments were launched, using the device-resident count.
// `fixed_rows` != 0: the table has exactly that many rows (one per persistent CTA, RangePassKernel).
__device__ __forceinline__ void SpineGeometry(uint32_t n, uint32_t tile_size, uint32_t fixed_rows, uint32_t& chunks,
                                              uint32_t& rows_per) {
  const uint32_t tiles = (uint32_t)CeilDiv(n, tile_size);
  chunks = fixed_rows ? fixed_rows : (uint32_t)CeilDiv(tiles, kSpineChunk);
  rows_per = (chunks + gridDim.x - 1) / gridDim.x;
}

#if !VRDX_SPINE_FUSED
__global__ void __launch_bounds__(kRadix)
SpineReduceKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, uint32_t tile_size, uint32_t fixed_rows,
                  uint32_t pass, const uint32_t* __restrict__ chunk_sums, uint32_t* __restrict__ seg,
                  StorageHeader* __restrict__ hdr) {
  __shared__ uint32_t s_last;
  __shared__ uint32_t s_warp[kRadix / 32];
  GridDepLaunch();
  GridDepWait();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  uint32_t chunks, rows_per;
  SpineGeometry(n, tile_size, fixed_rows, chunks, rows_per);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t r0 = blockIdx.x * rows_per;
  const uint32_t r1 = r0 + rows_per < chunks ? r0 + rows_per : chunks;
  uint32_t sum = 0;
#pragma unroll 16
  for (uint32_t r = r0; r < r1; ++r) sum += chunk_sums[(size_t)r * kRadix + tid];
  seg[(size_t)blockIdx.x * kRadix + tid] = sum;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&hdr->hist_blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  uint32_t run = 0;
  for (uint32_t s0 = 0; s0 < gridDim.x; s0 += 32) {  // 32 rows in flight per round trip
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (s0 + j < gridDim.x) ? __ldcg(seg + (size_t)(s0 + j) * kRadix + tid) : 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (s0 + j < gridDim.x) seg[(size_t)(s0 + j) * kRadix + tid] = run;
      run += v[j];
    }
  }
  // run == number of keys with digit `tid`: exclusive scan over digits -> global digit offsets
  const uint32_t incl = WarpInclusiveScan(run, lane);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t prefix = 0;
#pragma unroll
  for (int w = 0; w < kRadix / 32; ++w) prefix += (w < warp) ? s_warp[w] : 0u;
  hdr->global_hist[pass][tid] = prefix + incl - run;
  if (tid == 0 && pass == 0) hdr->element_count[0] = n;
  // one digit holds every key: this pass is the identity permutation (the scatter tiles just copy)
  const int any = __syncthreads_or(run == n && n != 0);
  if (tid == 0) hdr->pass_identity[pass] = any ? 1u : 0u;
}

__global__ void __launch_bounds__(kRadix)
SpineApplyKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, uint32_t tile_size, uint32_t fixed_rows,
                 uint32_t* __restrict__ chunk_sums, const uint32_t* __restrict__ seg, unsigned long long* ts_end) {
  GridDepLaunch();
  GridDepWait();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  uint32_t chunks, rows_per;
  SpineGeometry(n, tile_size, fixed_rows, chunks, rows_per);
  const int tid = threadIdx.x;
  const uint32_t r0 = blockIdx.x * rows_per;
  const uint32_t r1 = r0 + rows_per < chunks ? r0 + rows_per : chunks;
  uint32_t run = seg[(size_t)blockIdx.x * kRadix + tid];
  for (uint32_t b0 = r0; b0 < r1; b0 += 16) {  // 16 rows in flight; loads never wait behind the aliasing stores
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (b0 + j < r1) ? chunk_sums[(size_t)(b0 + j) * kRadix + tid] : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (b0 + j < r1) chunk_sums[(size_t)(b0 + j) * kRadix + tid] = run;
      run += v[j];
    }
  }
  StampEnd(ts_end);
}
#endif  // !VRDX_SPINE_FUSED

// SpineKernel — the two kernels above as one (what ships; VRDX_SPINE_FUSED=0 keeps the pair for A/B).  The grid is
// at most kSpineSegments (128) small CTAs, all co-resident on a 148-SM part, so a CTA may wait for the others:
//   1. CTA s sums the rows of segment s, publishes seg[s][256] and raises spine_flags[s]
//   2. waits for the flags of segments < s and sums their rows: the segment's exclusive prefix (no chain: every
//      CTA needs only the SUMS of its predecessors, which all appear one round trip after the slowest read)
//   3. rewrites its rows (still in L2) as exclusive prefixes; the last CTA also turns the per-digit totals into
//      the global digit offsets and the identity flag
// One launch and ~one kernel drain less per pass than the pair: 2^28 keys 26 -> ~10 us per pass.
__global__ void __launch_bounds__(kRadix)
SpineKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, uint32_t tile_size, uint32_t fixed_rows, uint32_t pass,
            uint32_t* __restrict__ chunk_sums, uint32_t* __restrict__ seg, StorageHeader* __restrict__ hdr,
            unsigned long long* ts_end) {
  __shared__ uint32_t s_warp[kRadix / 32];
  __shared__ uint32_t s_seg;
  GridDepLaunch();
  GridDepWait();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  uint32_t chunks, rows_per;
  SpineGeometry(n, tile_size, fixed_rows, chunks, rows_per);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // segment ids are handed out in arrival order (like the onesweep tiles): every segment a CTA waits for below
  // belongs to a CTA that is already running, whatever order the hardware starts the grid in
  if (tid == 0) s_seg = atomicAdd(&hdr->hist_blocks_done, 1u);
  __syncthreads();
  const uint32_t seg_id = s_seg;
  const uint32_t r0 = seg_id * rows_per;
  const uint32_t r1 = r0 + rows_per < chunks ? r0 + rows_per : chunks;
  uint32_t sum = 0;
#pragma unroll 16
  for (uint32_t r = r0; r < r1; ++r) sum += __ldcg(chunk_sums + (size_t)r * kRadix + tid);
  StRelaxed(seg + (size_t)seg_id * kRadix + tid, sum);
  __threadfence();
  __syncthreads();
  if (tid == 0) StRelaxed(&hdr->spine_flags[seg_id], 1u);
  // every predecessor is resident (grid <= 128 CTAs of 256 threads and no shared memory to speak of)
  if (warp == 0)
    for (uint32_t s0 = lane; s0 < seg_id; s0 += 32)
      while (LdRelaxed(&hdr->spine_flags[s0]) == 0u) {}
  __threadfence();
  __syncthreads();
  uint32_t run = 0;
  for (uint32_t s0 = 0; s0 < seg_id; s0 += 32) {  // 32 rows in flight per round trip
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (s0 + j < seg_id) ? LdRelaxed(seg + (size_t)(s0 + j) * kRadix + tid) : 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) run += v[j];
  }
  for (uint32_t b0 = r0; b0 < r1; b0 += 16) {  // 16 rows in flight; loads never wait behind the aliasing stores
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (b0 + j < r1) ? __ldcg(chunk_sums + (size_t)(b0 + j) * kRadix + tid) : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (b0 + j < r1) chunk_sums[(size_t)(b0 + j) * kRadix + tid] = run;
      run += v[j];
    }
  }
  if (seg_id == gridDim.x - 1) {
    // run == number of keys with digit `tid`: exclusive scan over digits -> global digit offsets
    const uint32_t incl = WarpInclusiveScan(run, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t prefix = 0;
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) prefix += (w < warp) ? s_warp[w] : 0u;
    hdr->global_hist[pass][tid] = prefix + incl - run;
    if (tid == 0 && pass == 0) hdr->element_count[0] = n;
    // one digit holds every key: this pass is the identity permutation (the scatter tiles just copy)
    const int any = __syncthreads_or(run == n && n != 0);
    if (tid == 0) hdr->pass_identity[pass] = any ? 1u : 0u;
  }
  StampEnd(ts_end);
}

// CopyBackKernel — a sort with an odd number of passes (bit sub-range, extension) ends in the
// scratch halves of the ping-pong; this moves [0, count) back into the caller's buffers and leaves
// [count, max) untouched, like every pass does.  8 B/key (16 for pairs), grid-stride.
__global__ void __launch_bounds__(256)
CopyBackKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, const uint32_t* __restrict__ keys_src,
               uint32_t* __restrict__ keys_dst, const uint32_t* __restrict__ vals_src, uint32_t* __restrict__ vals_dst,
               unsigned long long* ts_end) {
  GridDepLaunch();
  GridDepWait();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  const uint64_t stride = (uint64_t)gridDim.x * 256 * 4;
  for (uint64_t base = (uint64_t)blockIdx.x * 256 * 4 + threadIdx.x; base < n; base += stride) {
    uint32_t k[4], v[4];  // all loads first: four coalesced 1 KB rows in flight per CTA
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      k[j] = (i < n) ? LdStream(keys_src + i) : 0u;
      v[j] = (vals_src != nullptr && i < n) ? LdStream(vals_src + i) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      if (i < n) {
        keys_dst[i] = k[j];
        if (vals_src != nullptr) vals_dst[i] = v[j];
      }
    }
  }
  StampEnd(ts_end);
}

// ------------------------------------------------------------------------------------------
// 64-bit keys by composition (extension; the reference has 32-bit keys only).  An LSD sort of 64-bit
// words is a stable sort by the low word followed by a stable sort by the high word, and each of
// those is exactly the key-value sort above with the other half of the word as the payload:
//   Split64Kernel   keys64 -> lo[], hi[]   (applies the order-preserving codec of the key type)
//   key-value sort (key = lo, value = hi); key-value sort (key = hi, value = lo)
//   Merge64Kernel   lo[], hi[] -> keys64   (undoes the codec)
// 168 B/key moved instead of the 136 B/key of a native 8-pass 64-bit sort, with no new tile kernel.
// ------------------------------------------------------------------------------------------
struct KeyCodec64 {
  unsigned long long fmask, cmask, dmask;  // float64: f = 0x7FF..F, c = 0x800..0; int64: c only; descending: d = ~0
};
__device__ __forceinline__ unsigned long long KeyIn64(unsigned long long k, const KeyCodec64 c) {
  return k ^ ((unsigned long long)((long long)k >> 63) & c.fmask) ^ (c.cmask ^ c.dmask);
}
__device__ __forceinline__ unsigned long long KeyOut64(unsigned long long t, const KeyCodec64 c) {
  const unsigned long long u = t ^ c.dmask;
  return u ^ ((unsigned long long)(~(long long)u >> 63) & c.fmask) ^ c.cmask;
}

__global__ void __launch_bounds__(256)
Split64Kernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, const unsigned long long* __restrict__ keys,
              uint32_t* __restrict__ lo, uint32_t* __restrict__ hi, const KeyCodec64 codec) {
  const uint32_t n = ResolveCount(indirect, n_or_max);
  const uint64_t stride = (uint64_t)gridDim.x * 256 * 4;
  for (uint64_t base = (uint64_t)blockIdx.x * 256 * 4 + threadIdx.x; base < n; base += stride) {
    unsigned long long k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      k[j] = (i < n) ? __ldcs(keys + i) : 0ull;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      if (i < n) {
        const unsigned long long t = KeyIn64(k[j], codec);
        lo[i] = (uint32_t)t;
        hi[i] = (uint32_t)(t >> 32);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
Merge64Kernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, const uint32_t* __restrict__ lo,
              const uint32_t* __restrict__ hi, unsigned long long* __restrict__ keys, const KeyCodec64 codec) {
  const uint32_t n = ResolveCount(indirect, n_or_max);
  const uint64_t stride = (uint64_t)gridDim.x * 256 * 4;
  for (uint64_t base = (uint64_t)blockIdx.x * 256 * 4 + threadIdx.x; base < n; base += stride) {
    uint32_t l[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      l[j] = (i < n) ? LdStream(lo + i) : 0u;
      h[j] = (i < n) ? LdStream(hi + i) : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = base + (uint64_t)j * 256;
      if (i < n) keys[i] = KeyOut64(((unsigned long long)h[j] << 32) | l[j], codec);
    }
  }
}

}  // namespace vrdx
