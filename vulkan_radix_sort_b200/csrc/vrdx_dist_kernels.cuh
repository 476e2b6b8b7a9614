// vrdx_dist_kernels.cuh — kernels of the multi-GPU sort: splitter-search histograms and the
// class multi-split that precedes the NVLink exchange.  No reference counterpart (the reference
// is single-device); see include/vrdx_dist.h for the contracts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vrdx_kernels.cuh"

namespace vrdx {

constexpr int kDistMaxSplitters = 15;
constexpr int kDistMaxClasses = 2 * kDistMaxSplitters + 1;
constexpr int kDistClassSlotsForCount = 32;  // classes padded to a power of two (DistClassCountKernel)

struct DistPrefixes {
  uint32_t count;
  uint32_t shift;
};

// ---- splitter search: per-prefix 256-bin histograms -----------------------------------------
// Algorithmic traffic 4 B/key.  Grid-stride, 128-bit loads, shared-memory histograms
// (prefixCount x 256 bins), one global atomic per non-empty bin at the end.
constexpr int kDistHistThreads = 512;
constexpr size_t kDistHistMaxSmemBytes = 160 * 1024;  // dynamic shared memory the kernel is opted into per device (vrdxCudaCreateSorter)

__global__ void __launch_bounds__(kDistHistThreads)
DistPrefixHistogramKernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t digit_bits,
                          uint32_t prefix_count, const uint32_t* __restrict__ prefixes, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t sh[];  // [prefix_count][2^digit_bits]
  __shared__ uint32_t s_prefix[kDistMaxSplitters];
  __shared__ uint32_t s_cand[kRadix];  // per top byte: bit j set if prefix j can match a key with that top byte
  const int tid = threadIdx.x;
  const uint32_t bins = 1u << digit_bits;
  const uint32_t hi_shift = shift + digit_bits;  // bits above the digit
  const bool top = hi_shift >= 32;               // no bits above the digit: every key matches
  for (uint32_t i = tid; i < prefix_count * bins; i += kDistHistThreads) sh[i] = 0;
  if (tid < (int)prefix_count) s_prefix[tid] = prefixes[tid];
  if (tid < kRadix) {
    uint32_t m = 0;
    if (!top) {
      for (uint32_t j = 0; j < prefix_count; ++j) {
        const uint32_t p = prefixes[j];
        // top byte(s) of the keys matching prefix j: the prefix holds (32 - hi_shift) bits
        const uint32_t pbits = 32 - hi_shift;
        const bool hit = pbits >= 8 ? ((p >> (pbits - 8)) == (uint32_t)tid) : (((uint32_t)tid >> (8 - pbits)) == p);
        m |= hit ? (1u << j) : 0u;
      }
    }
    s_cand[tid] = m;
  }
  __syncthreads();
  const uint32_t dmask = bins - 1u;
  auto count_key = [&](uint32_t k) {
    const uint32_t d = (k >> shift) & dmask;
    if (top) {
      atomicAdd(&sh[d], 1u);  // prefix_count == 1 at the first level
      return;
    }
    uint32_t m = s_cand[k >> 24];  // ~(prefix_count / 256) of uniform keys get past this
    const uint32_t hi = k >> hi_shift;
    while (m) {
      const uint32_t j = __ffs(m) - 1;
      m &= m - 1;
      if (hi == s_prefix[j]) atomicAdd(&sh[j * bins + d], 1u);
    }
  };
  const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(keys) >> 2) & 3u);
  uint32_t head = mis ? 4u - mis : 0u;
  if (head > n) head = n;
  const uint4* __restrict__ body = reinterpret_cast<const uint4*>(keys + head);
  const uint64_t nvec = (uint64_t)(n - head) >> 2;
  const uint32_t tail_start = head + (uint32_t)(nvec << 2);
  // four 128-bit loads in flight per thread: the large-histogram configurations run one CTA per SM
  constexpr int kUnroll = 4;
  const uint64_t stride = (uint64_t)gridDim.x * kDistHistThreads;
  for (uint64_t v0 = (uint64_t)blockIdx.x * kDistHistThreads + tid; v0 < nvec; v0 += stride * kUnroll) {
    uint4 q[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t v = v0 + (uint64_t)u * stride;
      q[u] = v < nvec ? __ldcs(body + v) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (v0 + (uint64_t)u * stride < nvec) {
        count_key(q[u].x); count_key(q[u].y); count_key(q[u].z); count_key(q[u].w);
      }
    }
  }
  if (blockIdx.x == 0) {
    if ((uint32_t)tid < head) count_key(keys[tid]);
    const uint32_t t = tail_start + tid;
    if (tid < 4 && t < n) count_key(keys[t]);
  }
  __syncthreads();
  for (uint32_t i = tid; i < prefix_count * bins; i += kDistHistThreads) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(hist + i, c);
  }
}

// Class lookup by the top kDistLutBits bits of a key: if no splitter shares that prefix, every key
// with it has the same class (2 * #splitters below it) and one byte-wide shared-memory load
// classifies the key; otherwise (flag 0x80) the key needs the full comparison.  With 12 bits and 7
// splitters 0.2 % of uniform keys take the slow path (an 8-bit table sent 2.7 % of the keys, i.e.
// 58 % of the warp-instructions, there: 3.8 ms instead of 2.3 ms per 2^29 keys at 15 classes).
constexpr int kDistLutBits = 12;
constexpr int kDistLutSize = 1 << kDistLutBits;
__device__ __forceinline__ void DistBuildClassLut(uint8_t* s_lut, const uint32_t* __restrict__ splitters,
                                                  uint32_t splitter_count, int tid, int threads) {
  // The table is a step function with at most splitter_count steps.  Step 1: every group of 16
  // entries is filled with the value of its first entry (one 128-bit store per group).  Step 2:
  // the <= splitter_count groups that contain a splitter's prefix are recomputed entry by entry,
  // 16 threads per splitter (groups shared by several splitters get the same bytes twice).
  for (uint32_t g = tid; g < (uint32_t)kDistLutSize / 16; g += threads) {
    uint32_t below = 0;
    for (uint32_t j = 0; j < splitter_count; ++j) below += (__ldg(splitters + j) >> (32 - kDistLutBits)) < g * 16;
    const uint32_t v = 2u * below * 0x01010101u;
    reinterpret_cast<uint4*>(s_lut)[g] = make_uint4(v, v, v, v);
  }
  __syncthreads();
  for (uint32_t t = tid; t < 16 * splitter_count; t += threads) {
    const uint32_t b = ((__ldg(splitters + (t >> 4)) >> (32 - kDistLutBits)) & ~15u) + (t & 15u);
    uint32_t lo = 0, inside = 0;
    for (uint32_t j = 0; j < splitter_count; ++j) {
      const uint32_t ub = __ldg(splitters + j) >> (32 - kDistLutBits);
      lo += ub < b;
      inside |= ub == b;
    }
    s_lut[b] = (uint8_t)(inside ? 0x80u : 2u * lo);
  }
}

// Full comparison of a key against the splitters.  Only keys whose top byte equals a splitter's top
// byte get here (a few percent on any input that is not concentrated on the splitters), so the
// loop is kept out of line: inlined at every use it made the kernel 5,700 instructions long and
// instruction-fetch stalls its first limiter (profiles/kernels/k_dist_partition.txt).
__device__ __noinline__ uint32_t DistClassOfSlow(uint32_t k, const uint32_t* s_u, uint32_t splitter_count) {
  uint32_t gt = 0, eq = 0;
  for (uint32_t j = 0; j < splitter_count; ++j) {
    gt += k > s_u[j];
    eq |= k == s_u[j];
  }
  return 2u * gt + eq;
}

// ---- class sizes (count only) ----------------------------------------------------------------------
// Same class function as the multi-split below.  4 B/key read with 128-bit loads; every (class, lane)
// pair has its own shared-memory counter (address = class * 32 + lane, i.e. bank = lane), so the 32
// atomics of a warp-instruction never collide however few classes there are (with 3 classes a plain
// per-class counter would serialise ~11 lanes per instruction).  32 classes x 32 lanes x 16 warps = 64 KB.
constexpr int kDistCountThreads = 512;
constexpr size_t kDistCountSmemBytes = (size_t)(kDistCountThreads / 32) * kDistClassSlotsForCount * 32 * sizeof(uint32_t);

__global__ void __launch_bounds__(kDistCountThreads)
DistClassCountKernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t splitter_count,
                     const uint32_t* __restrict__ splitters, uint32_t* __restrict__ counts) {
  extern __shared__ __align__(16) uint32_t s_lane_cnt[];  // [warp][class][lane]
  __shared__ uint32_t s_u[kDistMaxSplitters];
  __shared__ __align__(16) uint8_t s_top[kDistLutSize];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < (int)splitter_count) s_u[tid] = splitters[tid];
  {
    uint4* z = reinterpret_cast<uint4*>(s_lane_cnt);
    for (int i = tid; i < (int)(kDistCountSmemBytes / 16); i += kDistCountThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  DistBuildClassLut(s_top, splitters, splitter_count, tid, kDistCountThreads);
  __syncthreads();
  uint32_t* mine = s_lane_cnt + warp * (kDistClassSlotsForCount * 32) + lane;
  auto count_key = [&](uint32_t k) {
    uint32_t c = s_top[k >> (32 - kDistLutBits)];
    if (c & 0x80u) c = DistClassOfSlow(k, s_u, splitter_count);
    atomicAdd(mine + c * 32, 1u);
  };
  const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(keys) >> 2) & 3u);
  uint32_t head = mis ? 4u - mis : 0u;
  if (head > n) head = n;
  const uint4* __restrict__ body = reinterpret_cast<const uint4*>(keys + head);
  const uint64_t nvec = (uint64_t)(n - head) >> 2;
  const uint32_t tail_start = head + (uint32_t)(nvec << 2);
  constexpr int kUnroll = 4;
  const uint64_t stride = (uint64_t)gridDim.x * kDistCountThreads;
  for (uint64_t v0 = (uint64_t)blockIdx.x * kDistCountThreads + tid; v0 < nvec; v0 += stride * kUnroll) {
    uint4 q[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t v = v0 + (uint64_t)u * stride;
      q[u] = v < nvec ? __ldcs(body + v) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (v0 + (uint64_t)u * stride < nvec) {
        count_key(q[u].x); count_key(q[u].y); count_key(q[u].z); count_key(q[u].w);
      }
    }
  }
  if (blockIdx.x == 0) {  // unaligned head (< 4 keys) and the n % 4 tail
    if ((uint32_t)tid < head) count_key(keys[tid]);
    const uint32_t t = tail_start + tid;
    if (tid < 4 && t < n) count_key(keys[t]);
  }
  __syncthreads();
  // thread (w, c) sums the 32 lane copies of (warp w, class c); one warp-level add per class, one global atomic
  const uint32_t classes = 2 * splitter_count + 1;
  for (uint32_t c = warp; c < classes; c += kDistCountThreads / 32) {
    uint32_t sum = 0;
    for (int w = 0; w < kDistCountThreads / 32; ++w) sum += s_lane_cnt[w * (kDistClassSlotsForCount * 32) + c * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0 && sum) atomicAdd(&counts[c], sum);
  }
}

// ---- class multi-split ------------------------------------------------------------------------
// One tile per CTA; a key's slot inside its class comes from a returning shared-memory atomicAdd
// on a warp-private class counter (any order inside a class will do) or, with very few classes,
// from a 5-round ballot loop.  Tile-local reorder
// through shared memory makes the global writes run-wise coalesced; each tile reserves its output
// range per class with one global atomic (order between tiles is irrelevant for keys-only).
constexpr int kDistPartThreads = 256;
constexpr int kDistPartItems = 16;
#ifndef VRDX_DIST_PART_MIN_CTAS
#define VRDX_DIST_PART_MIN_CTAS 5
#endif
constexpr int kDistPartMinCtas = VRDX_DIST_PART_MIN_CTAS;
constexpr int kDistPartTile = kDistPartThreads * kDistPartItems;
constexpr int kDistClassSlots = 32;  // classes padded to a power of two

// SCATTER = false: out[p] = key (class-ordered local buffer).
// SCATTER = true : dest_ptr[j][p - first_pos[j]] = key for the destination j that owns position p —
//                  the exchange is fused into the partition; the stores go to peer memory over NVLink.
constexpr int kDistMaxDests = kDistMaxSplitters + 1;
struct DistDestTable {
  unsigned long long ptr[kDistMaxDests];
  uint32_t first_pos[kDistMaxDests + 1];
};

template <bool SCATTER>
__global__ void __launch_bounds__(kDistPartThreads, kDistPartMinCtas)
DistPartitionKernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t splitter_count,
                    const uint32_t* __restrict__ splitters, uint32_t* __restrict__ cursors,
                    uint32_t* __restrict__ out, uint32_t dest_count,
                    const unsigned long long* __restrict__ dest_ptrs, const uint32_t* __restrict__ first_pos) {
  constexpr int kWarps = kDistPartThreads / 32;
  __shared__ uint32_t s_u[kDistMaxSplitters];
  __shared__ uint32_t s_cnt[kWarps][kDistClassSlots];  // per-warp class counts, later slot bases
  __shared__ uint32_t s_base[kDistClassSlots];         // tile-local first slot of each class
  __shared__ uint32_t s_gbase[kDistClassSlots];        // global slot of tile-local slot 0, per class
  __shared__ uint32_t s_keys[kDistPartTile];
  __shared__ __align__(16) uint8_t s_top[kDistLutSize];
  __shared__ unsigned long long s_dptr[kDistMaxDests];
  __shared__ uint32_t s_dpos[kDistMaxDests + 1];
  __shared__ uint32_t s_cdest[kDistClassSlots];        // SCATTER: destination that owns the first key of the class
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile_start = (uint64_t)blockIdx.x * kDistPartTile;
  if (tile_start >= n) return;
  if (SCATTER) {
    if (tid < (int)dest_count) s_dptr[tid] = dest_ptrs[tid];
    if (tid <= (int)dest_count) s_dpos[tid] = first_pos[tid];
  }
  const uint32_t remaining = (uint32_t)(n - tile_start);
  const uint32_t tile_count = remaining < (uint32_t)kDistPartTile ? remaining : (uint32_t)kDistPartTile;
  if (tid < (int)splitter_count) s_u[tid] = splitters[tid];
  if (lane < kDistClassSlots) s_cnt[warp][lane] = 0;
  DistBuildClassLut(s_top, splitters, splitter_count, tid, kDistPartThreads);
  __syncthreads();
  auto class_of = [&](uint32_t k) -> uint32_t {
    const uint32_t t = s_top[k >> (32 - kDistLutBits)];
    if (!(t & 0x80u)) return t;
    return DistClassOfSlow(k, s_u, splitter_count);
  };

  const uint32_t lt = LaneMaskLt();
  const bool few_classes = splitter_count <= 2;  // warp-uniform
  const uint32_t woff = warp * 32 * kDistPartItems + lane;
  const uint32_t* kin = keys + tile_start + woff;
  uint32_t key[kDistPartItems], rank[kDistPartItems];
  uint32_t cls[kDistPartItems / 4] = {};  // four 8-bit class labels per register
#pragma unroll
  for (int i = 0; i < kDistPartItems; ++i)
    key[i] = (woff + 32 * i < tile_count) ? LdStream(kin + 32 * i) : 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < kDistPartItems; ++i) {
    const bool valid = woff + 32 * i < tile_count;
    // invalid (pad) lanes take the unused top slot so that they rank after every real key
    const uint32_t c = valid ? class_of(key[i]) : (uint32_t)(kDistClassSlots - 1);
    cls[i >> 2] |= c << (8 * (i & 3));
    // Order inside a class is irrelevant (keys only; the destination sorts them), so with many
    // classes the slot is simply the value a returning shared-memory atomicAdd on the warp's class
    // counter hands back (the same observation as the order-free first pass of the sort).  With 2-3
    // destinations 16 lanes would hit one counter per instruction, and finding the peers with a
    // 5-round ballot loop is cheaper.  Measured per 2^29 keys: 3 classes 2.01 ms (ballots) vs 2.47
    // (atomics); 15 classes 2.49 vs 2.01.
    if (few_classes) {
      uint32_t peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const bool bit = (c >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
      }
      const uint32_t before = s_cnt[warp][c];
      const uint32_t below = __popc(peers & lt);
      __syncwarp();
      if (below == 0) s_cnt[warp][c] = before + __popc(peers);
      __syncwarp();
      rank[i] = before + below;
    } else {
      rank[i] = atomicAdd(&s_cnt[warp][c], 1u);
    }
  }
  __syncthreads();
  if (tid < kDistClassSlots) {  // one thread per class: totals over warps, tile scan, global reservation
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = s_cnt[w][tid];
      s_cnt[w][tid] = sum;
      sum += c;
    }
    const uint32_t incl = WarpInclusiveScan(sum, lane);
    const uint32_t excl = incl - sum;
    s_base[tid] = excl;
    const bool real = tid < (int)(2 * splitter_count + 1);
    const uint32_t g = (real && sum) ? atomicAdd(&cursors[tid], sum) : 0u;
    s_gbase[tid] = g - excl;
    if (SCATTER) {
      uint32_t j = 0;
      for (uint32_t d = 1; d < dest_count; ++d) j += g >= s_dpos[d];
      s_cdest[tid] = j;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kDistPartItems; ++i) {
    const uint32_t c = (cls[i >> 2] >> (8 * (i & 3))) & 0xFFu;
    rank[i] += s_cnt[warp][c] + s_base[c];
    s_keys[rank[i]] = key[i];
  }
  __syncthreads();
  // Fused exchange: every class run of the tile goes to its destination's receive buffer in LINE-ALIGNED
  // warp stores.  A warp-wide store that straddles a 128-byte line of the destination becomes two NVLink
  // write packets with partial payloads; with slot-linear stores (thread t writes slot t) ncu counted 1.96 GB
  // on the wire for 1.07 GB of keys (nvltx__bytes, profiles/r02/g_nvlink_partition_probe.txt), i.e. the links
  // were saturated at 55 % payload.  Here chunk q of a run covers destination words [32 q - a, 32 q - a + 32),
  // a = misalignment of the run's first word, so every store except the run's ends is one full line.
  // (The class-ordered local output of vrdxDistCmdPartition is written the same way: full 32-byte sectors.)
  const uint32_t classes = 2 * splitter_count + 1;
  for (uint32_t c = 0; c < classes; ++c) {
    const uint32_t b0 = s_base[c];
    const uint32_t cnt = (c + 1 < kDistClassSlots ? s_base[c + 1] : tile_count) - b0;  // pads sit in the last slot class
    if (cnt == 0) continue;
    const uint32_t p0 = s_gbase[c] + b0;                      // class-ordered position of the run's first key
    uint32_t j = SCATTER ? s_cdest[c] : 0u;
    if (SCATTER && j + 1 < dest_count && p0 + cnt > s_dpos[j + 1]) {
      // the run crosses a destination boundary (ties on a splitter value): element-wise
      for (uint32_t r = tid; r < cnt; r += kDistPartThreads) {
        const uint32_t p = p0 + r;
        uint32_t jj = j;
        while (jj + 1 < dest_count && p >= s_dpos[jj + 1]) ++jj;
        reinterpret_cast<uint32_t*>(s_dptr[jj])[p - s_dpos[jj]] = s_keys[b0 + r];
      }
      continue;
    }
    uint32_t* dst = SCATTER ? reinterpret_cast<uint32_t*>(s_dptr[j]) + (p0 - s_dpos[j]) : out + p0;
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 2) & 31u;
    const uint32_t chunks = (cnt + a + 31u) >> 5;
    for (uint32_t q = warp; q < chunks; q += kWarps) {
      const uint32_t rel = q * 32u + lane - a;               // wraps below zero for the lanes before the run
      if (rel < cnt) dst[rel] = s_keys[b0 + rel];
    }
  }
}

}  // namespace vrdx
