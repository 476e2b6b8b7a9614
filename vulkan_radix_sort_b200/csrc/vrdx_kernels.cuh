// vrdx_kernels.cuh — sm_100a device code of the 32-bit LSD radix sort (8-bit digits x 4 passes).
//
// What these kernels replace in the reference (jaesung-cs/vulkan_radix_sort v0.4.0):
//   src/shader/upsweep.slang:10-45    per-partition digit histogram, re-read of the keys every pass
//   src/shader/spine.slang:11-84      exclusive scan over partitions + global histogram scan
//   src/shader/downsweep.slang:41-224 stable rank, local reorder, scatter (keys / key-value)
// with a different decomposition:
//   HistogramKernel   ONE read of the keys builds all four 256-bin digit histograms; the last
//                     CTA to finish exclusive-scans them in place (no separate spine launch).
//   OnesweepKernel    one launch per pass: warp-level multi-split ranking (peer masks built with
//                     shared-memory atomicOr in warp-private cells), tile-local reorder through
//                     shared memory, single-pass decoupled look-back across tiles for the digit
//                     offsets, run-wise coalesced scatter.  Keys and values are two separate
//                     arrays end to end, as in the reference.
// Stability: a warp owns 32*IPT consecutive keys and ranks them item by item, lane by lane, so
// (warp, item, lane) order == index order — the same argument as downsweep.slang:79-80.
#pragma once
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

// ------------------------------------------------------------------------------------------
// HistogramKernel — all four digit histograms in one pass over the keys + exclusive scan.
// Algorithmic traffic: 4 B/key read.  Grid: a multiple of the SM count (persistent, grid-stride
// over chunks of THREADS*16 keys, four 128-bit loads in flight per thread).
// ------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;
constexpr int kHistVecPerThread = 4;                                   // uint4 loads per thread per chunk
constexpr int kHistChunk = kHistThreads * kHistVecPerThread * 4;       // keys per chunk

__device__ __forceinline__ void HistCount(uint32_t (*sh)[kRadix], uint32_t k) {
  atomicAdd(&sh[0][k & 0xFFu], 1u);
  atomicAdd(&sh[1][(k >> 8) & 0xFFu], 1u);
  atomicAdd(&sh[2][(k >> 16) & 0xFFu], 1u);
  atomicAdd(&sh[3][k >> 24], 1u);
}

__global__ void __launch_bounds__(kHistThreads)
HistogramKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ indirect,
                uint32_t n_or_max, StorageHeader* __restrict__ hdr) {
  __shared__ uint32_t sh[kPasses][kRadix];
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x;
  const uint32_t n = ResolveCount(indirect, n_or_max);

  for (int i = tid; i < kPasses * kRadix; i += kHistThreads) (&sh[0][0])[i] = 0;
  if (blockIdx.x == 0 && tid == 0) hdr->element_count[0] = n;
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
        HistCount(sh, v[j].x); HistCount(sh, v[j].y); HistCount(sh, v[j].z); HistCount(sh, v[j].w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kHistVecPerThread; ++j) {
        uint64_t idx = base + (uint64_t)j * kHistThreads + tid;
        if (idx < nvec) {
          uint4 q = __ldcs(body + idx);
          HistCount(sh, q.x); HistCount(sh, q.y); HistCount(sh, q.z); HistCount(sh, q.w);
        }
      }
    }
  }
  if (blockIdx.x == 0) {  // unaligned head (< 4 keys) and the n % 4 tail
    if ((uint32_t)tid < head) HistCount(sh, keys[tid]);
    const uint32_t t = tail_start + tid;
    if (tid < 4 && t < n) HistCount(sh, keys[t]);
  }
  __syncthreads();

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

  // Last CTA: exclusive scan of each 256-bin histogram in place (spine.slang:62-83 does this
  // once per pass in workgroup 0; here once per sort).  Warp p scans pass p, 8 bins per lane.
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < kPasses) {
    uint32_t* h = gh + warp * kRadix + lane * 8;
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = __ldcg(h + j); sum += c[j]; }
    uint32_t excl = WarpInclusiveScan(sum, lane) - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = excl; excl += c[j]; }
  }
}

// ------------------------------------------------------------------------------------------
// OnesweepKernel — one LSD pass over one tile per CTA.
// Algorithmic traffic per pass: 4 B/key read + 4 B/key write (+ 4 + 4 for values).
//
// Ranking (measured on B200, tools/microbench_rank.cu, cycles per 32 keys per SM at 32 warps/SM):
//   hardware MATCH.ANY on an 8-bit digit   60.7      (cost grows with the number of distinct values)
//   8-round ballot loop (the reference's downsweep.slang:92-99, also CUB's choice)   28.8
//   shared-memory atomicOr peer mask + counter cell (this kernel)   16.5
// so the peer mask of a key (lanes of its warp holding the same digit) is built with ONE
// shared-memory atomicOr into a warp-private cell and read back; the result is independent of
// the order in which the hardware serialises colliding lanes, so ranks are by lane order and the
// pass is stable.  Mask and running count of a (warp, digit) share one 8-byte cell, so a key
// costs one ATOMS, one LDS.64 and (for the lowest peer lane only) one STS.64.
// ------------------------------------------------------------------------------------------
struct PassArgs {
  const uint32_t* indirect;   // device count or nullptr
  uint32_t n_or_max;          // elementCount (direct) or maxElementCount (indirect)
  uint32_t pass;              // 0..3
  StorageHeader* hdr;
  uint32_t* status;           // look-back cells of this pass: [tile][256]
  uint32_t* status_next;      // cells of the next pass, cleared here (nullptr on the last pass)
  const uint32_t* keys_in;
  uint32_t* keys_out;
  const uint32_t* vals_in;
  uint32_t* vals_out;
};

template <int THREADS, int IPT, bool KV, int MIN_CTAS>
struct PassConfig {
  static constexpr int kThreads = THREADS;
  static constexpr int kItems = IPT;
  static constexpr int kMinCtas = MIN_CTAS;
  static constexpr bool kKeyValue = KV;
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kTile = THREADS * IPT;
  static constexpr int kMiscWords = 16;
  // [kWarps][256] uint2 cells | keys[kTile] | vals[kTile] (KV) | gbase[256] | misc
  static constexpr size_t kSmemBytes =
      sizeof(uint32_t) * ((size_t)kWarps * kRadix * 2 + (size_t)kTile * (KV ? 2 : 1) + kRadix + kMiscWords);
  static_assert(THREADS % 32 == 0 && THREADS >= kRadix && THREADS <= 1024, "one thread per digit is assumed");
  static_assert(kTile <= (1 << 16), "tile ranks are kept below 2^16");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
OnesweepKernel(const PassArgs a) {
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;

  extern __shared__ __align__(16) uint32_t smem[];
  uint2* s_cell = reinterpret_cast<uint2*>(smem);     // [kWarps][256] {peer mask scratch, running count / base}
  uint32_t* s_keys = smem + kWarps * kRadix * 2;      // [kTile] tile reordered by digit
  uint32_t* s_vals = s_keys + kTile;                  // [kTile] (KV only)
  uint32_t* s_gbase = s_vals + (KV ? kTile : 0);      // [256] global slot of tile-local slot 0, per digit
  uint32_t* s_misc = s_gbase + kRadix;                // [0..7] warp totals, [8] tile id

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t shift = a.pass * kRadixBits;
  const uint32_t n = ResolveCount(a.indirect, a.n_or_max);

  // Tile ids are handed out in arrival order so that every predecessor a tile may wait on in
  // the look-back is already resident (forward progress without relying on blockIdx order).
  if (tid == 0) s_misc[8] = atomicAdd(&a.hdr->tickets[a.pass], 1u);
  {
    uint4* z = reinterpret_cast<uint4*>(s_cell);
#pragma unroll
    for (int j = tid; j < kWarps * kRadix / 2; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  const uint32_t tile = s_misc[8];
  const uint64_t tile_start = (uint64_t)tile * kTile;
  if (tile_start >= n) return;  // indirect count below max: surplus CTAs retire (upsweep.slang:20-22)
  const uint32_t remaining = (uint32_t)(n - tile_start);
  const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;
  const bool full = tile_count == (uint32_t)kTile;

  if (a.status_next != nullptr && tid < kRadix) a.status_next[(size_t)tile * kRadix + tid] = 0;

  // ---- load: warp-striped, 128 B per warp-instruction -------------------------------------
  uint32_t key[IPT];
  const uint32_t woff = warp * 32 * IPT + lane;
  {
    const uint32_t* kin = a.keys_in + tile_start + woff;
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = LdStream(kin + 32 * i);
    } else {
      // Tail tile: pad with the largest key so pads rank after every real key (the reference
      // pads the same way, downsweep.slang:81,85); their slots are >= tile_count and never stored.
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = (woff + 32 * i < tile_count) ? LdStream(kin + 32 * i) : 0xFFFFFFFFu;
    }
  }

  // ---- warp-level multi-split: rank of each key among equal digits inside its warp ---------
  uint32_t rank[IPT];
  {
    uint2* cell = s_cell + warp * kRadix;
    const uint32_t lt = LaneMaskLt();
    const uint32_t lanebit = 1u << lane;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = (key[i] >> shift) & 0xFFu;
      atomicOr(&cell[d].x, lanebit);
      __syncwarp();
      const uint2 pc = cell[d];  // {peers of this key in the warp, count of digit d in earlier items}
      const uint32_t below = __popc(pc.x & lt);
      __syncwarp();
      if (below == 0) cell[d] = make_uint2(0u, pc.y + __popc(pc.x));  // lowest peer clears the mask, bumps the count
      __syncwarp();
      rank[i] = pc.y + below;
    }
  }
  __syncthreads();

  // ---- per-digit: counts over warps, publish aggregate, tile-local exclusive scan -----------
  uint32_t digit_count = 0, digit_excl = 0;
  uint32_t wcount[kWarps];
  if (tid < kRadix) {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      wcount[w] = s_cell[w * kRadix + tid].y;
      sum += wcount[w];
    }
    // pads were counted as digit 255; they are not part of the data
    digit_count = sum - ((tid == kRadix - 1) ? ((uint32_t)kTile - tile_count) : 0u);
    StRelaxed(a.status + (size_t)tile * kRadix + tid,
              (tile == 0 ? kStatusPrefix : kStatusAggregate) | digit_count);
    const uint32_t incl = WarpInclusiveScan(sum, lane);
    if (lane == 31) s_misc[warp] = incl;
    digit_excl = incl - sum;  // exclusive within the warp
  }
  __syncthreads();
  uint32_t first_look = 0;
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) digit_excl += (w < warp) ? s_misc[w] : 0u;
    // cell.y becomes the tile-local slot of the first key of (warp, digit)
    uint32_t run = digit_excl;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s_cell[w * kRadix + tid].y = run;
      run += wcount[w];
    }
    // start the first look-back load now; it is consumed after the reorder below
    if (tile > 0) first_look = LdRelaxed(a.status + (size_t)(tile - 1) * kRadix + tid);
  }
  __syncthreads();

  // ---- tile-local reorder through shared memory --------------------------------------------
  {
    const uint2* cell = s_cell + warp * kRadix;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = (key[i] >> shift) & 0xFFu;
      rank[i] += cell[d].y;
      s_keys[rank[i]] = key[i];
    }
    if (KV) {
      // values are fetched only now, so they do not occupy registers during the ranking
      const uint32_t* vin = a.vals_in + tile_start + woff;
      uint32_t val[IPT];
#pragma unroll
      for (int i = 0; i < IPT; ++i) val[i] = (full || woff + 32 * i < tile_count) ? LdStream(vin + 32 * i) : 0u;
#pragma unroll
      for (int i = 0; i < IPT; ++i) s_vals[rank[i]] = val[i];
    }
  }

  // ---- decoupled look-back: exclusive prefix of this digit over all earlier tiles ------------
  if (tid < kRadix) {
    uint32_t excl = 0;
    if (tile > 0) {
      uint32_t look = tile - 1;
      uint32_t s = first_look;
      while (true) {
        if ((s >> 30) != 0u) {
          excl += s & kStatusValueMask;
          if (s & kStatusPrefix) break;
          --look;  // tile 0 always publishes a prefix, so this never underflows
        }
        s = LdRelaxed(a.status + (size_t)look * kRadix + tid);
      }
      StRelaxed(a.status + (size_t)tile * kRadix + tid, kStatusPrefix | (excl + digit_count));
    }
    // global slot of tile-local slot 0 for this digit (mod 2^32 arithmetic)
    s_gbase[tid] = a.hdr->global_hist[a.pass][tid] + excl - digit_excl;
  }
  __syncthreads();

  // ---- scatter: consecutive threads write consecutive slots of a digit run -------------------
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t slot = i * THREADS + tid;
    const uint32_t k = s_keys[slot];
    const uint32_t g = s_gbase[(k >> shift) & 0xFFu] + slot;
    if (full || slot < tile_count) {
      a.keys_out[g] = k;
      if (KV) a.vals_out[g] = s_vals[slot];
    }
  }
}

}  // namespace vrdx
