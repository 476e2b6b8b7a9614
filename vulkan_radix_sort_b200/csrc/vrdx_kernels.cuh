// vrdx_kernels.cuh — sm_100a device code of the 32-bit LSD radix sort (8-bit digits x 4 passes).
//
// What these kernels replace in the reference (jaesung-cs/vulkan_radix_sort v0.4.0):
//   src/shader/upsweep.slang:10-45    per-partition digit histogram, re-read of the keys every pass
//   src/shader/spine.slang:11-84      exclusive scan over partitions + global histogram scan
//   src/shader/downsweep.slang:41-224 stable rank, local reorder, scatter (keys / key-value)
// with a different decomposition:
//   HistogramKernel   ONE read of the keys builds all four 256-bin digit histograms; the last
//                     CTA to finish exclusive-scans them in place (no separate spine launch).
//   OnesweepKernel    one launch per pass: warp-level multi-split ranking (optimistic returning
//                     shared-memory atomicAdd on warp-private counters + lane-ordered repair of
//                     collisions), tile-local reorder through shared memory, single-pass
//                     decoupled look-back across tiles for the digit offsets, run-wise coalesced
//                     scatter.  Keys and values are two separate
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
// Ranking.  Measured on B200 (tools/microbench_rank.cu; cycles per 32 keys per SM, 32 warps/SM):
//   hardware MATCH.ANY on an 8-bit digit                                   60.7
//   8-round ballot loop (reference downsweep.slang:92-99; CUB's choice)    28.8
//   shared-memory atomicOr peer mask + counter cell                        16.5
//   shared-memory atomicAdd (returning) on a warp-private counter           4.0
// and the first full kernel built on atomicOr cells was bound by the shared-memory pipe
// (profiles/: l1tex 77 % busy, 31 wavefronts per 32 keys).  So the rank of a key among equal
// digits of its warp is computed OPTIMISTICALLY: every lane does one returning atomicAdd(+1) on
// the warp-private counter of its digit.  A lane that is alone with its digit in this
// warp-instruction (89 % of lanes on uniform keys) gets its rank straight from the returned
// value.  Lanes that collided are served by the hardware in an unspecified order, so they are
// detected (counter read-back: more than one increment landed after my returned value) and
// REPAIRED in lane order: one shuffle + one ballot per collision group (1.8 groups per 32 uniform
// keys).  If many lanes collide (low-entropy digits) the repair switches to the fixed 8-round
// ballot loop.  Ranks are therefore exactly those of a stable counting sort whatever order the
// hardware serialises colliding atomics in.
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

constexpr int kLookBatch = 4;          // look-back cells fetched per round trip
constexpr int kRepairBallotThreshold = 12;  // colliding lanes above which the 8-round ballot loop is cheaper

template <int THREADS, int IPT, bool KV, int MIN_CTAS>
struct PassConfig {
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
  uint32_t* s_cnt = smem;                             // [kWarps][256] warp-private digit counters, later slot bases
  uint32_t* s_keys = s_cnt + kWarps * kRadix;         // [kTile] tile reordered by digit
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
    uint4* z = reinterpret_cast<uint4*>(s_cnt);
#pragma unroll
    for (int j = tid; j < kWarps * kRadix / 4; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
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
    uint32_t* cnt = s_cnt + warp * kRadix;
    const uint32_t lt = LaneMaskLt();
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = (key[i] >> shift) & 0xFFu;
      const uint32_t old = atomicAdd(&cnt[d], 1u);   // optimistic: exact if no other lane holds digit d
      __syncwarp();
      const uint32_t fin = cnt[d];                   // all 32 increments of this item have landed
      uint32_t r = old;
      uint32_t suspects = __ballot_sync(0xffffffffu, fin - old > 1u);  // someone was served after me
      if (suspects != 0u) {                                            // warp-uniform
        if (__popc(suspects) > kRepairBallotThreshold) {
          // many collisions (low-entropy digit): fixed-cost peer masks for every lane
          uint32_t peers = 0xffffffffu;
#pragma unroll
          for (int b = 0; b < kRadixBits; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t m = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? m : ~m;
          }
          r = fin - __popc(peers) + __popc(peers & lt);
        } else {
          // few collision groups: repair them one by one, in lane order
          do {
            const uint32_t dstar = __shfl_sync(0xffffffffu, d, __ffs(suspects) - 1);
            const uint32_t peers = __ballot_sync(0xffffffffu, d == dstar);
            if (d == dstar) r = fin - __popc(peers) + __popc(peers & lt);
            suspects &= ~peers;
          } while (suspects != 0u);
        }
      }
      rank[i] = r;
      __syncwarp();
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
      wcount[w] = s_cnt[w * kRadix + tid];
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
  uint32_t look_s[kLookBatch];
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) digit_excl += (w < warp) ? s_misc[w] : 0u;
    // s_cnt becomes the tile-local slot of the first key of (warp, digit)
    uint32_t run = digit_excl;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s_cnt[w * kRadix + tid] = run;
      run += wcount[w];
    }
    // start the first batch of look-back loads now; it is consumed after the reorder below
#pragma unroll
    for (int j = 0; j < kLookBatch; ++j) {
      const uint32_t t = (tile > (uint32_t)j) ? tile - 1 - j : 0u;
      look_s[j] = (tile > 0) ? LdRelaxed(a.status + (size_t)t * kRadix + tid) : 0u;
    }
  }
  __syncthreads();

  // ---- tile-local reorder through shared memory --------------------------------------------
  {
    const uint32_t* base = s_cnt + warp * kRadix;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = (key[i] >> shift) & 0xFFu;
      rank[i] += base[d];
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
  // kLookBatch predecessor cells are in flight per round trip; they are consumed strictly in
  // order (nearest tile first) and the walk stops at the first inclusive prefix.
  if (tid < kRadix) {
    uint32_t excl = 0;
    if (tile > 0) {
      uint32_t look = tile - 1;  // nearest tile not yet consumed
      bool done = false;
      while (!done) {
#pragma unroll
        for (int j = 0; j < kLookBatch; ++j) {
          if (done) break;
          const uint32_t s = look_s[j];
          if ((s >> 30) == 0u) break;  // not published yet: re-poll from `look`
          excl += s & kStatusValueMask;
          if (s & kStatusPrefix) { done = true; break; }
          --look;  // tile 0 always publishes a prefix, so this never underflows
        }
        if (!done) {
#pragma unroll
          for (int j = 0; j < kLookBatch; ++j) {
            const uint32_t t = (look >= (uint32_t)j) ? look - j : 0u;
            look_s[j] = LdRelaxed(a.status + (size_t)t * kRadix + tid);
          }
        }
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
