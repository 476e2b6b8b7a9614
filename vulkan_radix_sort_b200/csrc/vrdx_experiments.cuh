// vrdx_experiments.cuh — tile-kernel variants that were built, measured on B200 and did NOT win
// (evidence under profiles/r01_*).  They are not part of the product
// library: this file is compiled only with -DVRDX_EXPERIMENTS (VRDX_EXPERIMENTS=1 python -m
// vulkan_radix_sort_b200.build), which also adds their selectors to the sorter options, so the
// A/B measurements stay reproducible.
//   OnesweepKernel         the round-1 tile kernel (run-time shift, ranks in keys, per-group shuffle
//                          repair loop); MODE 2 = tile id from blockIdx.x; PAIRED 64-bit staging
//   OnesweepClusterKernel  one decoupled look-back per thread-block cluster (DSMEM)
//   OnesweepTmaKernel      persistent CTAs, cp.async.bulk (TMA) double-buffered tile staging
//   RangePassKernel        reduce-then-scan whose tables are sized by the machine / a constant instead of by N:
//                          a CTA walks a RANGE of consecutive tiles with running digit offsets (round 2)
#pragma once
#include <cooperative_groups.h>

#include "vrdx_kernels.cuh"

namespace vrdx {

// Rank of one key among the keys of its warp that hold the same digit and come earlier in
// (item, lane) order; `cnt` is the warp-private counter row.  See the block comment above.
__device__ __forceinline__ uint32_t WarpRankDigit(uint32_t* cnt, uint32_t d, uint32_t lt) {
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
  __syncwarp();
  return r;
}


// MODE 0: onesweep (tile ids from the ticket counter, offsets by decoupled look-back).
// MODE 2: onesweep with tile id = blockIdx.x (no ticket): relies on CTAs being dispatched in
//         increasing blockIdx order, as CUB's decoupled-look-back DeviceScan does.
// MODE 1: downsweep of the reduce-then-scan variant — tile id = blockIdx.x and `a.status` already
//         holds the exclusive prefix over tiles of every digit (UpsweepKernel + Spine*Kernel), so
//         the kernel has no inter-CTA communication at all (the reference's downsweep shape).
// GENERIC = false: the reference's plan (shift = 8 * pass, mask = 0xFF, no codec) folded at compile
// time — the code the measurements in DESIGN.md are about.  GENERIC = true: digit and codec from PassArgs.
template <class Cfg, int MODE = 0, bool GENERIC = false>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
OnesweepKernel(const PassArgs a) {
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;
  constexpr int kLookBatch = Cfg::kLookBatch;

  extern __shared__ __align__(128) uint32_t smem[];
  uint32_t* s_cnt = smem;                             // [kWarps][256] warp-private digit counters, later slot bases
  uint32_t* s_keys = s_cnt + kWarps * kRadix;         // [kTile] tile reordered by digit
  uint32_t* s_vals = s_keys + kTile;                  // [kTile] (KV only)
  uint32_t* s_gbase = s_vals + (KV ? kTile : 0);      // [256] global slot of tile-local slot 0, per digit
  uint32_t* s_misc = s_gbase + kRadix;                // [0..7] warp totals, [8] tile id

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t shift = GENERIC ? a.shift : a.pass * kRadixBits;
  const uint32_t mask = GENERIC ? a.mask : (uint32_t)(kRadix - 1);
  const KeyCodec cin = GENERIC ? a.codec_in : KeyCodec{0u, 0u, 0u};    // identity codecs fold away
  const KeyCodec cout = GENERIC ? a.codec_out : KeyCodec{0u, 0u, 0u};
  const uint32_t n = ResolveCount(a.indirect, a.n_or_max);
  GridDepLaunch();
  {
    uint4* z = reinterpret_cast<uint4*>(s_cnt);
#pragma unroll
    for (int j = tid; j < kWarps * kRadix / 4; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  GridDepWait();  // everything below reads what the previous kernel of this sort wrote
  // Tile ids are handed out in arrival order so that every predecessor a tile may wait on in
  // the look-back is already resident (forward progress without relying on blockIdx order).
  // Keys-only onesweep pass 0 is order-free (see the ranking below): tiles claim their output
  // ranges with global atomics, so it needs neither tickets nor the look-back chain.
  const bool order_free = !KV && a.order_free != 0u;
  const bool unordered = (MODE != 1) && order_free;
  if (MODE == 0 && !unordered && tid == 0) s_misc[8] = atomicAdd(&a.hdr->tickets[a.pass], 1u);  // MODE 1/2: blockIdx.x
  __syncthreads();

  const uint32_t tile = (MODE == 0 && !unordered) ? s_misc[8] : blockIdx.x;
  const uint64_t tile_start = (uint64_t)tile * kTile;
  if (tile_start >= n) return;  // indirect count below max: surplus CTAs retire (upsweep.slang:20-22)
  const uint32_t remaining = (uint32_t)(n - tile_start);
  const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;
  const bool full = tile_count == (uint32_t)kTile;

  if (MODE != 1 && a.status_next != nullptr && tid < kRadix) a.status_next[(size_t)tile * kRadix + tid] = 0;

  // ---- constant digit: a stable counting sort with one non-empty bucket is a copy --------------
  if (a.hdr->pass_identity[a.pass]) {
    const uint32_t* kin = a.keys_in + tile_start;
    uint32_t* kout = a.keys_out + tile_start;
    uint32_t ck[IPT], cv[KV ? IPT : 1];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {  // all loads first: the stores below may alias them as far as the compiler knows
      const uint32_t idx = i * THREADS + tid;
      ck[i] = idx < tile_count ? KeyOut(KeyIn(LdStream(kin + idx), cin), cout) : 0u;
      if (KV) cv[i] = idx < tile_count ? LdStream(a.vals_in + tile_start + idx) : 0u;
    }
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t idx = i * THREADS + tid;
      if (idx < tile_count) {
        kout[idx] = ck[i];
        if (KV) a.vals_out[tile_start + idx] = cv[i];
      }
    }
    StampEnd(a.ts_end);
    return;
  }

  // ---- load: warp-striped, 128 B per warp-instruction -------------------------------------
  uint32_t key[IPT];
  const uint32_t woff = warp * 32 * IPT + lane;
  {
    const uint32_t* kin = a.keys_in + tile_start + woff;
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = KeyIn(LdStream(kin + 32 * i), cin);
    } else {
      // Tail tile: pad with the largest word so pads rank after every real key (the reference
      // pads the same way, downsweep.slang:81,85); their slots are >= tile_count and never stored.
#pragma unroll
      for (int i = 0; i < IPT; ++i)
        key[i] = (woff + 32 * i < tile_count) ? KeyIn(LdStream(kin + 32 * i), cin) : 0xFFFFFFFFu;
    }
  }

  // ---- warp-level multi-split: rank of each key among equal digits inside its warp ---------
  uint32_t rank[IPT];
  {
    uint32_t* cnt = s_cnt + warp * kRadix;
    const uint32_t lt = LaneMaskLt();
    if (order_free && full) {
      // Keys-only, first pass of a sort over all 32 bits: there is no earlier order to preserve
      // and equal keys are indistinguishable, so ANY bijective ranking inside a digit gives the
      // same final output.  The value returned by the atomic is such a ranking: no read-back, no
      // collision repair.  (Every later pass, every pass of a key-value or bit-sub-range sort, and
      // the tail tile — whose pads must keep ranking after the real keys — are stable.)
#pragma unroll
      for (int i = 0; i < IPT; ++i) rank[i] = atomicAdd(&cnt[(key[i] >> shift) & mask], 1u);
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        rank[i] = WarpRankDigit(cnt, (key[i] >> shift) & mask, lt);
      }
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
    // pads were counted as the largest digit of this pass; they are not part of the data
    digit_count = sum - (((uint32_t)tid == mask) ? ((uint32_t)kTile - tile_count) : 0u);
    if (MODE != 1 && !unordered)
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
    if (unordered) {
      // claim [excl, excl + digit_count) of this digit's global run; the round trip overlaps the reorder
      look_s[0] = digit_count ? atomicAdd(&a.hdr->claim_cursor[tid], digit_count) : 0u;
    } else if (MODE != 1) {
#pragma unroll
      for (int j = 0; j < kLookBatch; ++j) {
        const uint32_t t = (tile > (uint32_t)j) ? tile - 1 - j : 0u;
        look_s[j] = (tile > 0) ? LdRelaxed(a.status + (size_t)t * kRadix + tid) : 0u;
      }
    } else {
      // reduce-then-scan: scanned chunk prefix + the rows of the earlier tiles of this chunk
      const uint32_t chunk = tile / kSpineChunk;
      const uint32_t in_chunk = tile % kSpineChunk;
      const uint32_t* row = a.status + (size_t)chunk * kSpineChunk * kRadix + tid;
      uint32_t part[kSpineChunk - 1];
#pragma unroll
      for (int j = 0; j < kSpineChunk - 1; ++j) part[j] = ((uint32_t)j < in_chunk) ? row[(size_t)j * kRadix] : 0u;  // all in flight at once
      uint32_t acc = a.status_next[(size_t)chunk * kRadix + tid];
#pragma unroll
      for (int j = 0; j < kSpineChunk - 1; ++j) acc += part[j];
      look_s[0] = acc;
    }
  }
  __syncthreads();

  // ---- tile-local reorder through shared memory --------------------------------------------
  {
    const uint32_t* base = s_cnt + warp * kRadix;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t d = (key[i] >> shift) & mask;
      rank[i] += base[d];
      if (!Cfg::kPaired) s_keys[rank[i]] = key[i];
    }
    if (KV) {
      // values are fetched only now, so they do not occupy registers during the ranking
      const uint32_t* vin = a.vals_in + tile_start + woff;
      uint32_t val[IPT];
#pragma unroll
      for (int i = 0; i < IPT; ++i) val[i] = (full || woff + 32 * i < tile_count) ? LdStream(vin + 32 * i) : 0u;
      if (Cfg::kPaired) {
        // keys were not stored above (see kPaired there): one 64-bit store per pair
        uint2* s_kv = reinterpret_cast<uint2*>(s_keys);
#pragma unroll
        for (int i = 0; i < IPT; ++i) s_kv[rank[i]] = make_uint2(key[i], val[i]);
      } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) s_vals[rank[i]] = val[i];
      }
    }
  }

  // ---- decoupled look-back: exclusive prefix of this digit over all earlier tiles ------------
  // kLookBatch predecessor cells are in flight per round trip; they are consumed strictly in
  // order (nearest tile first) and the walk stops at the first inclusive prefix.
  if (tid < kRadix) {
    uint32_t excl = 0;
    if (unordered || MODE == 1) {
      excl = look_s[0];
    } else if (tile > 0) {
      excl = LookBack<kLookBatch>(a.status, tile, tid, look_s, a.hdr->reserved);
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
    if (Cfg::kPaired) {
      const uint2 kv = reinterpret_cast<const uint2*>(s_keys)[slot];
      const uint32_t g = s_gbase[(kv.x >> shift) & mask] + slot;
      if (full || slot < tile_count) {
        a.keys_out[g] = KeyOut(kv.x, cout);
        a.vals_out[g] = kv.y;
      }
    } else {
      const uint32_t k = s_keys[slot];
      const uint32_t g = s_gbase[(k >> shift) & mask] + slot;
      if (full || slot < tile_count) {
        a.keys_out[g] = KeyOut(k, cout);
        if (KV) a.vals_out[g] = s_vals[slot];
      }
    }
  }
  StampEnd(a.ts_end);
}

// ------------------------------------------------------------------------------------------
// OnesweepClusterKernel — onesweep with ONE look-back per thread-block cluster.
//
// Measured on B200 at N = 2^28: the look-back-free scatter pass (reduce-then-scan downsweep)
// takes 0.92 ms, the same pass with a per-tile decoupled look-back 1.24 ms.  The look-back depth
// is D ~ lambda * (tiles per cycle): with ~440 tiles in flight and ~700 cycles per L2 round trip a
// tile has to sum ~12 predecessor cells, 1 KB each, while its CTA waits.  Clusters shrink the
// chain: the CLUSTER CTAs of a cluster sort CLUSTER consecutive tiles, exchange their per-digit
// counts through distributed shared memory, and the cluster appears on the global chain as ONE
// participant (one status row, one ticket).  The 256 digits are split across the CTAs of the
// cluster, so each CTA runs the look-back for 256/CLUSTER digits only, in its last warp(s), while
// all other warps reorder the tile.  Chain participants, look-back depth, status traffic and
// status memory all drop by CLUSTER x.
// ------------------------------------------------------------------------------------------
template <class Cfg, int CLUSTER>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
OnesweepClusterKernel(const PassArgs a) {
  namespace cg = cooperative_groups;
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;
  constexpr int kLookBatch = Cfg::kLookBatch;
  constexpr int kSlice = kRadix / CLUSTER;  // digits whose look-back this CTA runs
  static_assert(kRadix % CLUSTER == 0 && kSlice % 32 == 0 && kSlice <= THREADS, "digit slice must be whole warps");

  extern __shared__ __align__(128) uint32_t smem[];
  uint32_t* s_cnt = smem;                             // [kWarps][256]
  uint32_t* s_keys = s_cnt + kWarps * kRadix;         // [kTile]
  uint32_t* s_vals = s_keys + kTile;                  // [kTile] (KV only)
  uint32_t* s_gbase = s_vals + (KV ? kTile : 0);      // [256]
  uint32_t* s_misc = s_gbase + kRadix;                // [0..7] warp totals, [8] cluster ticket (rank 0)
  __shared__ uint32_t s_tot[kRadix];                  // this tile's digit counts, read by the slice owners over DSMEM
  __shared__ uint32_t s_ext[kRadix];                  // written by the slice owners over DSMEM: keys of this digit in all earlier tiles

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t crank = cluster.block_rank();
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t shift = a.pass * kRadixBits;
  const uint32_t n = ResolveCount(a.indirect, a.n_or_max);

  GridDepLaunch();
  {
    uint4* z = reinterpret_cast<uint4*>(s_cnt);
#pragma unroll
    for (int j = tid; j < kWarps * kRadix / 4; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  GridDepWait();
  // one ticket per cluster, drawn by rank 0 and read by the other CTAs over DSMEM
  if (crank == 0 && tid == 0) s_misc[8] = atomicAdd(&a.hdr->tickets[a.pass], 1u);
  cluster.sync();
  const uint32_t ctile = *cluster.map_shared_rank(&s_misc[8], 0);
  const uint64_t cluster_start = (uint64_t)ctile * CLUSTER * kTile;
  if (cluster_start >= n) {  // whole cluster past the (indirect) count: retire together
    cluster.sync();          // rank 0's shared memory must outlive the reads above
    return;
  }
  const uint32_t tile = ctile * CLUSTER + crank;
  const uint64_t tile_start = (uint64_t)tile * kTile;
  const bool active = tile_start < n;  // CTAs past the count still serve their digit slice
  const uint32_t remaining = active ? (uint32_t)(n - tile_start) : 0u;
  const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;
  const bool full = tile_count == (uint32_t)kTile;

  // ---- load + warp-level multi-split (identical to OnesweepKernel) ---------------------------
  uint32_t key[IPT];
  uint32_t rank[IPT];
  const uint32_t woff = warp * 32 * IPT + lane;
  if (active) {
    const uint32_t* kin = a.keys_in + tile_start + woff;
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = LdStream(kin + 32 * i);
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = (woff + 32 * i < tile_count) ? LdStream(kin + 32 * i) : 0xFFFFFFFFu;
    }
    uint32_t* cnt = s_cnt + warp * kRadix;
    const uint32_t lt = LaneMaskLt();
#pragma unroll
    for (int i = 0; i < IPT; ++i) rank[i] = WarpRankDigit(cnt, (key[i] >> shift) & 0xFFu, lt);
  }
  __syncthreads();

  // ---- per-digit: counts over warps, tile-local exclusive scan --------------------------------
  uint32_t digit_count = 0, digit_excl = 0;
  uint32_t wcount[kWarps];
  if (tid < kRadix) {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      wcount[w] = s_cnt[w * kRadix + tid];
      sum += wcount[w];
    }
    digit_count = sum - ((active && tid == kRadix - 1) ? ((uint32_t)kTile - tile_count) : 0u);
    const uint32_t incl = WarpInclusiveScan(sum, lane);
    if (lane == 31) s_misc[warp] = incl;
    digit_excl = incl - sum;
  }
  __syncthreads();
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w) digit_excl += (w < warp) ? s_misc[w] : 0u;
    uint32_t run = digit_excl;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s_cnt[w * kRadix + tid] = run;
      run += wcount[w];
    }
    s_tot[tid] = digit_count;
  }
  cluster.sync();  // every tile's digit counts are visible cluster-wide

  // ---- slice owners: cluster aggregate, first look-back loads -----------------------------------
  const bool is_slice = tid >= THREADS - kSlice;
  const int my_digit = (int)crank * kSlice + (tid - (THREADS - kSlice));
  uint32_t tot[CLUSTER];
  uint32_t ctotal = 0;
  uint32_t look_s[kLookBatch];
  if (is_slice) {
#pragma unroll
    for (int r = 0; r < CLUSTER; ++r) {
      tot[r] = *cluster.map_shared_rank(&s_tot[my_digit], r);
      ctotal += tot[r];
    }
    StRelaxed(a.status + (size_t)ctile * kRadix + my_digit,
              (ctile == 0 ? kStatusPrefix : kStatusAggregate) | ctotal);
    if (a.status_next != nullptr) a.status_next[(size_t)ctile * kRadix + my_digit] = 0;
#pragma unroll
    for (int j = 0; j < kLookBatch; ++j) {
      const uint32_t t = (ctile > (uint32_t)j) ? ctile - 1 - j : 0u;
      look_s[j] = (ctile > 0) ? LdRelaxed(a.status + (size_t)t * kRadix + my_digit) : 0u;
    }
  }

  // ---- tile-local reorder through shared memory -------------------------------------------------
  if (active) {
    const uint32_t* base = s_cnt + warp * kRadix;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      rank[i] += base[(key[i] >> shift) & 0xFFu];
      s_keys[rank[i]] = key[i];
    }
    if (KV) {
      const uint32_t* vin = a.vals_in + tile_start + woff;
      uint32_t val[IPT];
#pragma unroll
      for (int i = 0; i < IPT; ++i) val[i] = (full || woff + 32 * i < tile_count) ? LdStream(vin + 32 * i) : 0u;
#pragma unroll
      for (int i = 0; i < IPT; ++i) s_vals[rank[i]] = val[i];
    }
  }

  // ---- slice owners: look-back over earlier CLUSTERS, then hand every tile its offset -----------
  if (is_slice) {
    uint32_t excl = 0;
    if (ctile > 0) {
      excl = LookBack<kLookBatch>(a.status, ctile, my_digit, look_s);
      StRelaxed(a.status + (size_t)ctile * kRadix + my_digit, kStatusPrefix | (excl + ctotal));
    }
#pragma unroll
    for (int r = 0; r < CLUSTER; ++r) {
      *cluster.map_shared_rank(&s_ext[my_digit], r) = excl;  // keys of this digit in all earlier tiles
      excl += tot[r];
    }
  }
  cluster.sync();  // offsets delivered; no distributed shared-memory access after this point

  if (tid < kRadix) s_gbase[tid] = a.hdr->global_hist[a.pass][tid] + s_ext[tid] - digit_excl;
  __syncthreads();

  // ---- scatter ----------------------------------------------------------------------------------
  if (active) {
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
  StampEnd(a.ts_end);
}

// ------------------------------------------------------------------------------------------
// OnesweepTmaKernel — the same pass as OnesweepKernel, as a PERSISTENT kernel with TMA staging.
//
// The grid is one wave of co-resident CTAs (SM count x CTAs/SM).  Each CTA loops over tiles it
// draws from the ticket counter; the raw keys of a tile are brought into shared memory by ONE
// bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP)
// issued a whole tile ahead into the other half of a double buffer, so neither the global-load
// latency nor the ticket's atomic round trip is on the critical path, no warp sleeps on a
// scoreboard for its keys, and no CTA launch/teardown happens between tiles.  The staging buffer
// of the current tile is reused as the reorder buffer once every thread holds its keys in
// registers.  Key-value sorts stage the tile's values the same way (single buffer, issued when
// the tile starts, consumed at the reorder).  Needs 16-byte aligned key/value/storage
// addresses (cp.async.bulk); the host falls back to OnesweepKernel otherwise.  Tail (partial)
// tiles are loaded with guarded scalar loads.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t SmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "VRDX_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra VRDX_WAIT;\n\t}"
      ::"r"(SmemAddr(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA), bytes a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void TmaLoad1D(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(SmemAddr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(SmemAddr(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void FenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int THREADS, int IPT, bool KV, int MIN_CTAS, int LOOK_BATCH = 4>
struct TmaPassConfig {
  static constexpr int kLookBatch = LOOK_BATCH;
  static constexpr int kThreads = THREADS;
  static constexpr int kItems = IPT;
  static constexpr int kMinCtas = MIN_CTAS;
  static constexpr bool kKeyValue = KV;
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kTile = THREADS * IPT;
  static constexpr int kMiscWords = 32;
  // stage[2][kTile] | vstage[kTile] (KV) | cnt[kWarps][256] | gbase[256] | misc (tile ids, mbarriers)
  static constexpr size_t kSmemBytes =
      sizeof(uint32_t) * ((size_t)kTile * (KV ? 3 : 2) + (size_t)kWarps * kRadix + kRadix + kMiscWords);
  static_assert(THREADS % 32 == 0 && THREADS >= kRadix && THREADS <= 1024, "one thread per digit is assumed");
  static_assert(kTile % 4 == 0 && kTile <= (1 << 16), "tile bytes must be a multiple of 16");
};

// MODE 0: onesweep (tickets + decoupled look-back).  MODE 1: scatter pass of reduce-then-scan —
// tiles are strided statically over the persistent CTAs (tile = blockIdx.x + k * gridDim.x) and
// `a.status` holds the exclusive per-tile digit prefixes, so there is no inter-CTA ordering to
// respect and the prefetch distance costs nothing.
template <class Cfg, int MODE = 0>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
OnesweepTmaKernel(const PassArgs a) {
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;
  constexpr uint32_t kTileBytes = kTile * sizeof(uint32_t);
  constexpr int kLookBatch = Cfg::kLookBatch;
  constexpr int kProducer = THREADS - 1;  // lane 31 of the last warp: idle during the per-digit phases when THREADS > 256

  extern __shared__ __align__(128) uint32_t smem[];
  uint32_t* s_stage = smem;                                   // [2][kTile] raw tile, then the tile reordered by digit
  uint32_t* s_vstage = s_stage + 2 * kTile;                   // [kTile] values (KV only)
  uint32_t* s_cnt = s_vstage + (KV ? kTile : 0);              // [kWarps][256]
  uint32_t* s_gbase = s_cnt + kWarps * kRadix;                // [256]
  uint32_t* s_misc = s_gbase + kRadix;                        // [0..7] warp totals, [8..9] tile id per slot
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc + 16); // [0..1] key stages, [2] value stage

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t shift = a.pass * kRadixBits;
  const uint32_t n = ResolveCount(a.indirect, a.n_or_max);
  const uint32_t lt = LaneMaskLt();

  // Producer: draw a tile id and, if it is a full tile, start its bulk copy into `slot`.
  auto fetch_tile = [&](int slot, uint32_t tile) {
    s_misc[8 + slot] = tile;
    const uint64_t start = (uint64_t)tile * kTile;
    if (start + kTile <= (uint64_t)n) {
      MbarExpectTx(&s_bar[slot], kTileBytes);
      TmaLoad1D(s_stage + slot * kTile, a.keys_in + start, kTileBytes, &s_bar[slot]);
    }
  };

  GridDepLaunch();
  GridDepWait();
  if (tid == kProducer) {
    MbarInit(&s_bar[0], 1);
    MbarInit(&s_bar[1], 1);
    MbarInit(&s_bar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    FenceProxyAsync();
    const uint32_t t0 = (MODE == 0) ? atomicAdd(&a.hdr->tickets[a.pass], 1u) : blockIdx.x;
    fetch_tile(0, t0);
    const uint32_t t1 = (MODE == 0) ? atomicAdd(&a.hdr->tickets[a.pass], 1u) : blockIdx.x + gridDim.x;
    fetch_tile(1, t1);
  }
  __syncthreads();

  for (uint32_t iter = 0;; ++iter) {
    const int slot = iter & 1;
    uint32_t* stage = s_stage + slot * kTile;
    const uint32_t tile = s_misc[8 + slot];
    const uint64_t tile_start = (uint64_t)tile * kTile;
    if (tile_start >= n) break;  // tickets are monotonic: every later tile of this CTA is out of range too
    const uint32_t remaining = (uint32_t)(n - tile_start);
    const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;
    const bool full = tile_count == (uint32_t)kTile;

    // values of this tile: one bulk copy, consumed at the reorder (its latency hides behind the ranking)
    if (KV && full && tid == kProducer) {
      MbarExpectTx(&s_bar[2], kTileBytes);
      TmaLoad1D(s_vstage, a.vals_in + tile_start, kTileBytes, &s_bar[2]);
    }
    {
      uint4* z = reinterpret_cast<uint4*>(s_cnt);
#pragma unroll
      for (int j = tid; j < kWarps * kRadix / 4; j += THREADS) z[j] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (MODE == 0 && a.status_next != nullptr && tid < kRadix) a.status_next[(size_t)tile * kRadix + tid] = 0;

    // ---- keys: shared-memory stage -> registers (warp-striped, conflict-free) -------------------
    uint32_t key[IPT];
    const uint32_t woff = warp * 32 * IPT + lane;
    if (full) {
      MbarWait(&s_bar[slot], (iter >> 1) & 1u);
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = stage[woff + 32 * i];
    } else {
      const uint32_t* kin = a.keys_in + tile_start + woff;
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = (woff + 32 * i < tile_count) ? LdStream(kin + 32 * i) : 0xFFFFFFFFu;
    }
    __syncthreads();  // counters zeroed; every thread holds its keys, so `stage` may be overwritten

    // ---- warp-level multi-split ---------------------------------------------------------------
    uint32_t rank[IPT];
    {
      uint32_t* cnt = s_cnt + warp * kRadix;
#pragma unroll
      for (int i = 0; i < IPT; ++i) rank[i] = WarpRankDigit(cnt, (key[i] >> shift) & 0xFFu, lt);
    }
    __syncthreads();

    // ---- per-digit: counts over warps, publish aggregate, tile-local exclusive scan -------------
    uint32_t digit_count = 0, digit_excl = 0;
    uint32_t wcount[kWarps];
    if (tid < kRadix) {
      uint32_t sum = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        wcount[w] = s_cnt[w * kRadix + tid];
        sum += wcount[w];
      }
      digit_count = sum - ((tid == kRadix - 1) ? ((uint32_t)kTile - tile_count) : 0u);
      if (MODE == 0)
        StRelaxed(a.status + (size_t)tile * kRadix + tid,
                  (tile == 0 ? kStatusPrefix : kStatusAggregate) | digit_count);
      const uint32_t incl = WarpInclusiveScan(sum, lane);
      if (lane == 31) s_misc[warp] = incl;
      digit_excl = incl - sum;
    }
    __syncthreads();
    uint32_t look_s[kLookBatch];
    uint32_t next_ticket = 0;
    if (tid < kRadix) {
#pragma unroll
      for (int w = 0; w < kRadix / 32; ++w) digit_excl += (w < warp) ? s_misc[w] : 0u;
      uint32_t run = digit_excl;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        s_cnt[w * kRadix + tid] = run;
        run += wcount[w];
      }
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < kLookBatch; ++j) {
          const uint32_t t = (tile > (uint32_t)j) ? tile - 1 - j : 0u;
          look_s[j] = (tile > 0) ? LdRelaxed(a.status + (size_t)t * kRadix + tid) : 0u;
        }
      } else {
        const uint32_t chunk = tile / kSpineChunk;
        const uint32_t in_chunk = tile % kSpineChunk;
        const uint32_t* row = a.status + (size_t)chunk * kSpineChunk * kRadix + tid;
        uint32_t part[kSpineChunk - 1];
#pragma unroll
        for (int j = 0; j < kSpineChunk - 1; ++j) part[j] = ((uint32_t)j < in_chunk) ? row[(size_t)j * kRadix] : 0u;
        uint32_t acc = a.status_next[(size_t)chunk * kRadix + tid];
#pragma unroll
        for (int j = 0; j < kSpineChunk - 1; ++j) acc += part[j];
        look_s[0] = acc;
      }
    }
    // the id of the tile that will replace this one is drawn here, off the critical path
    if (tid == kProducer)
      next_ticket = (MODE == 0) ? atomicAdd(&a.hdr->tickets[a.pass], 1u) : tile + 2u * gridDim.x;
    __syncthreads();

    // ---- tile-local reorder through shared memory (into the stage this tile arrived in) ---------
    {
      const uint32_t* base = s_cnt + warp * kRadix;
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        rank[i] += base[(key[i] >> shift) & 0xFFu];
        stage[rank[i]] = key[i];
      }
    }
    if (KV) {
      uint32_t val[IPT];
      if (full) {
        MbarWait(&s_bar[2], iter & 1u);
#pragma unroll
        for (int i = 0; i < IPT; ++i) val[i] = s_vstage[woff + 32 * i];
      } else {
        const uint32_t* vin = a.vals_in + tile_start + woff;
#pragma unroll
        for (int i = 0; i < IPT; ++i) val[i] = (woff + 32 * i < tile_count) ? LdStream(vin + 32 * i) : 0u;
      }
      __syncthreads();  // every thread holds its values before the buffer is permuted in place
#pragma unroll
      for (int i = 0; i < IPT; ++i) s_vstage[rank[i]] = val[i];
    }

    // ---- decoupled look-back --------------------------------------------------------------------
    if (tid < kRadix) {
      uint32_t excl = 0;
      if (MODE == 0) {
        if (tile > 0) {
          excl = LookBack<kLookBatch>(a.status, tile, tid, look_s);
          StRelaxed(a.status + (size_t)tile * kRadix + tid, kStatusPrefix | (excl + digit_count));
        }
      } else {
        excl = look_s[0];
      }
      s_gbase[tid] = a.hdr->global_hist[a.pass][tid] + excl - digit_excl;
    }
    __syncthreads();

    // ---- scatter --------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t slot_i = i * THREADS + tid;
      const uint32_t k = stage[slot_i];
      const uint32_t g = s_gbase[(k >> shift) & 0xFFu] + slot_i;
      if (full || slot_i < tile_count) {
        a.keys_out[g] = k;
        if (KV) a.vals_out[g] = s_vstage[slot_i];
      }
    }
    __syncthreads();  // the stage, the value buffer and the per-digit tables are free again

    if (tid == kProducer) {
      FenceProxyAsync();  // order the generic-proxy accesses above before the async-proxy refill
      fetch_tile(slot, next_ticket);
    }
    // s_misc[8+slot] is next read two iterations from now, after several barriers
  }
  StampEnd(a.ts_end);
}

// ------------------------------------------------------------------------------------------
// RangePassKernel — the scatter pass of reduce-then-scan with tables sized by a CONSTANT, not
// by N (north_star: "a temp-storage layout sized for 148 SMs"; the reference keeps one 1 KB
// row per 4096 keys, h.in:353-362).
//
// CTA r owns the `range_tiles` consecutive tiles of range r.  UpsweepRangeKernel counted the
// range's digits into ONE row range_counts[r][256], the spine turned the rows into exclusive
// prefixes, and this kernel walks the range tile by tile with a running per-digit offset in a
// register of the digit thread: a tile's own digit counts fall out of its ranking anyway, so
// no per-tile table exists.  The number of ranges is capped at kMaxRanges rows (the host grows
// range_tiles with N), so the tables take at most 8.4 MB for any N <= 2^32 - 1 where the
// reference's partHist takes N / 16 bytes (64 MB at 2^28).  Two more things the loop buys: the keys
// of the next tile are requested as soon as the current tile sits in shared memory (their
// latency hides behind the scatter), and every warp re-zeroes only its own counter row (no
// block barrier for it).
// Why ranges of 8..32 tiles and not ONE range per co-resident CTA (148 x 5 = 740 rows): measured
// (profiles/r02/b_persistent_static_ranges.txt), 740 static ranges run the keys-only sort at 2^28
// in 5.99 ms instead of 3.76 ms.  The CTAs of a wave then write 740 x 256 run fronts that are
// megabytes apart instead of 256 fronts a wave's tiles share: ncu shows 1.31x the algorithmic DRAM
// traffic (partially written sectors are evicted before the same CTA's next tile extends them)
// and every SM keeps > 1000 distinct 2 MB pages hot.  Short ranges handed out in order by the
// hardware CTA scheduler keep the wave's working set as compact as one-tile CTAs do.
// ------------------------------------------------------------------------------------------
constexpr int kMaxRanges = 8192;  // rows of the range table
constexpr int kMinRangeTiles = 8;

// Tiles [first, first + count) of range r.
__device__ __forceinline__ void RangeOf(uint32_t tiles, uint32_t range_tiles, uint32_t r, uint32_t& first, uint32_t& count) {
  const uint64_t f = (uint64_t)r * range_tiles;
  first = f < tiles ? (uint32_t)f : tiles;
  count = tiles - first < range_tiles ? tiles - first : range_tiles;
}

template <class Cfg, bool GENERIC, int RANK = VRDX_RANK>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinCtas)
RangePassKernel(const PassArgs a) {
  constexpr int THREADS = Cfg::kThreads;
  constexpr int IPT = Cfg::kItems;
  constexpr bool KV = Cfg::kKeyValue;
  constexpr int kWarps = Cfg::kWarps;
  constexpr int kTile = Cfg::kTile;

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
  }
  GridDepWait();
  const bool order_free = !KV && a.order_free != 0u;
  const PassDigit<GENERIC> dg = MakePassDigit<Cfg, GENERIC>(a, order_free);
  const uint32_t tiles = (uint32_t)CeilDiv((uint64_t)n, (uint64_t)kTile);
  uint32_t tile, range_tiles;
  RangeOf(tiles, a.range_tiles, blockIdx.x, tile, range_tiles);
  if (range_tiles == 0) return;  // indirect count below max: surplus CTAs retire
  const uint32_t tile_end = tile + range_tiles;

  if (a.hdr->pass_identity[pass]) {
    for (; tile < tile_end; ++tile) {
      const uint64_t start = (uint64_t)tile * kTile;
      const uint32_t rest = (uint32_t)(n - start);
      TileCopy<Cfg, GENERIC>(a, start, rest < (uint32_t)kTile ? rest : (uint32_t)kTile, tid, dg);
    }
    StampEnd(a.ts_end);
    return;
  }

  // digit thread: global slot of the next key of digit `tid` this range writes
  uint32_t run = 0;
  if (tid < kRadix) run = a.hdr->global_hist[pass][tid] + a.status[(size_t)blockIdx.x * kRadix + tid];

  uint32_t key[IPT];
  const uint32_t woff = warp * 32 * IPT + lane;
  uint32_t* const row = sm.cnt + warp * kRadix;
  {
    const uint64_t start = (uint64_t)tile * kTile;
    const uint32_t rest = (uint32_t)(n - start);
    TileLoadKeys<Cfg, GENERIC>(key, a.keys_in, start, rest < (uint32_t)kTile ? rest : (uint32_t)kTile, woff, dg);
  }
  __syncthreads();  // counters zeroed

  for (;;) {
    const uint64_t tile_start = (uint64_t)tile * kTile;
    const uint32_t remaining = (uint32_t)(n - tile_start);
    const uint32_t tile_count = remaining < (uint32_t)kTile ? remaining : (uint32_t)kTile;

    uint32_t rank2[IPT / 2];
    TileRank<Cfg, GENERIC, RANK>(key, rank2, row, tile_count == (uint32_t)kTile, dg);
    __syncthreads();

    uint32_t digit_count = 0, digit_excl = 0;
    uint32_t wcount[kWarps];
    if (tid < kRadix) TileDigitSums<Cfg>(sm, wcount, digit_count, digit_excl, tile_count, dg.mask, tid);
    __syncthreads();
    if (tid < kRadix) {
      TileSlotBases<Cfg>(sm, wcount, digit_excl, tid);
      sm.gbase[tid] = run - (digit_excl >> 2);  // read only after the next barrier; last read two barriers ago
      run += digit_count;
    }
    __syncthreads();

    uint32_t val[KV ? IPT : 1];
    if (kKvEarlyValues) TileLoadValues<Cfg>(val, a.vals_in, tile_start, tile_count, woff);
    TileReorder<Cfg, GENERIC>(sm, key, rank2, row, a.vals_in, tile_start, tile_count, woff, dg, val);
    // this warp's counter row is free again (only this warp reads its slot bases): zero it for the next tile
    __syncwarp();
    {
      uint4* z = reinterpret_cast<uint4*>(row);
      z[lane] = make_uint4(0u, 0u, 0u, 0u);
      z[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
    }
    // next tile's keys: in flight while this tile is scattered
    const bool more = tile + 1 < tile_end;
#ifndef VRDX_RANGE_PREFETCH
#define VRDX_RANGE_PREFETCH 1
#endif
    if (VRDX_RANGE_PREFETCH && more) {
      const uint64_t next = tile_start + kTile;
      const uint32_t rest = (uint32_t)(n - next);
      TileLoadKeys<Cfg, GENERIC>(key, a.keys_in, next, rest < (uint32_t)kTile ? rest : (uint32_t)kTile, woff, dg);
    }
    __syncthreads();

    TileScatter<Cfg, GENERIC>(sm, a.keys_out, a.vals_out, tile_count, tid, dg);
    if (!more) break;
    ++tile;
    if (!VRDX_RANGE_PREFETCH) {
      const uint64_t next = (uint64_t)tile * kTile;
      const uint32_t rest = (uint32_t)(n - next);
      TileLoadKeys<Cfg, GENERIC>(key, a.keys_in, next, rest < (uint32_t)kTile ? rest : (uint32_t)kTile, woff, dg);
    }
    // no barrier here: the next write to keys / vals / gbase comes after the next tile's three barriers,
    // the counters of a warp are touched by that warp alone until the first of them
  }
  StampEnd(a.ts_end);
}

// UpsweepRangeKernel — the upsweep that goes with RangePassKernel: CTA r counts the digits of ALL tiles
// of range r into one shared histogram and writes one row range_counts[r][256].  4 B/key read, no
// per-tile output.  Launched with the same grid and range_tiles as RangePassKernel.
template <int TILE>
__global__ void __launch_bounds__(kUpsweepThreads)
UpsweepRangeKernel(const uint32_t* __restrict__ indirect, uint32_t n_or_max, uint32_t range_tiles, uint32_t shift,
                   uint32_t mask, const KeyCodec codec_in, const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ range_counts,
                   StorageHeader* __restrict__ hdr, unsigned long long* ts_end) {
  constexpr int THREADS = kUpsweepThreads;
  static_assert(THREADS == kRadix, "one thread per digit");
  __shared__ uint32_t h[kRadix];
  const int tid = threadIdx.x;
  GridDepLaunch();
  const uint32_t n = ResolveCount(indirect, n_or_max);
  const uint32_t tiles = (uint32_t)CeilDiv((uint64_t)n, (uint64_t)TILE);
  uint32_t first, count;
  RangeOf(tiles, range_tiles, blockIdx.x, first, count);
  h[tid] = 0;
  GridDepWait();
  ResetSpineState(hdr, tid);
  __syncthreads();
  constexpr int kIters = (TILE + THREADS - 1) / THREADS;
  for (uint32_t tile = first; tile < first + count; ++tile) {
    const uint64_t tile_start = (uint64_t)tile * TILE;
    const uint32_t remaining = (uint32_t)(n - tile_start);
    const uint32_t tile_count = remaining < (uint32_t)TILE ? remaining : (uint32_t)TILE;
    const uint32_t* kin = keys_in + tile_start;
    if (tile_count == (uint32_t)TILE) {
      uint32_t k[kIters];
#pragma unroll
      for (int i = 0; i < kIters; ++i) k[i] = LdStream(kin + i * THREADS + tid);
#pragma unroll
      for (int i = 0; i < kIters; ++i) atomicAdd(&h[(KeyIn(k[i], codec_in) >> shift) & mask], 1u);
    } else {
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const uint32_t idx = i * THREADS + tid;
        if (idx < tile_count) atomicAdd(&h[(KeyIn(LdStream(kin + idx), codec_in) >> shift) & mask], 1u);
      }
    }
  }
  __syncthreads();
  range_counts[(size_t)blockIdx.x * kRadix + tid] = h[tid];  // ranges without tiles write zeros
  StampEnd(ts_end);
}


}  // namespace vrdx
