/*
 * oracle/lsd_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the hot path of jaesung-cs/vulkan_radix_sort v0.4.0
 * (vrdxCmdSort* -> gpuSort -> upsweep / spine / downsweep).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product library (libvrdx_b200.so) never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py checks both restatements below against
 * oracle/_ref/libvrdx_ref.so, which is the reference's own bench/cpu_benchmark.cc +
 * bench/data_generator.cc compiled unmodified from /root/reference (the same check
 * the reference applies to its GPU path, bench/bench.cc:41-64), and against the
 * committed fixtures in tests/golden/ generated from that library.
 *
 * Two restatements, both bit-exact with each other and with the reference:
 *
 *  1. vrdx_oracle_sort_*            "net effect": 4 stable counting-sort passes on
 *     bytes 0..3, ping-pong user -> scratch -> user -> scratch -> user
 *     (src/vk_radix_sort.h.in:400-427 pass loop and binding swap;
 *      src/shader/upsweep.slang:33 digit = bitfieldExtract(key, 8*pass, 8)).
 *
 *  2. vrdx_oracle_sort_partitioned_* "structural": the reference's three dispatches
 *     per pass with 4096-key partitions:
 *       upsweep   per-partition histogram, pads count as key 0xFFFFFFFF
 *                 (src/shader/upsweep.slang:30-44),
 *       spine     exclusive scan over partitions per digit + exclusive scan of the
 *                 256-bin global histogram (src/shader/spine.slang:32-60, 62-83),
 *       downsweep stable local rank, dst = global[d] + partExcl[p][d] + (rank - localStart[d]),
 *                 store only if dst < count (src/shader/downsweep.slang:77-115, 164-201,
 *                 values 203-224).
 *     It also reproduces the indirect-count contract: the grid is sized from
 *     max_count, only the first `count` elements move, the tail [count, max) of the
 *     user buffers is untouched (src/vk_radix_sort.h.in:368-379, upsweep.slang:20-22,
 *     downsweep.slang:58-59).
 *
 * All arithmetic is uint32; no floating point.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VRDX_RADIX 256u
#define VRDX_PARTITION_SIZE 4096u /* src/vk_radix_sort.h.in:100-103 */

/* ---------------------------------------------------------------- net effect */

static void counting_pass(const uint32_t* src_k, const uint32_t* src_v, uint32_t* dst_k,
                          uint32_t* dst_v, uint64_t n, int pass) {
  uint64_t hist[VRDX_RADIX];
  memset(hist, 0, sizeof(hist));
  const int shift = 8 * pass;
  for (uint64_t i = 0; i < n; ++i) hist[(src_k[i] >> shift) & 0xFFu]++;
  uint64_t sum = 0;
  for (uint32_t d = 0; d < VRDX_RADIX; ++d) {
    uint64_t c = hist[d];
    hist[d] = sum;
    sum += c;
  }
  if (src_v) {
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t dst = hist[(src_k[i] >> shift) & 0xFFu]++;
      dst_k[dst] = src_k[i];
      dst_v[dst] = src_v[i];
    }
  } else {
    for (uint64_t i = 0; i < n; ++i) dst_k[hist[(src_k[i] >> shift) & 0xFFu]++] = src_k[i];
  }
}

/* Sort keys[0:n] in place (result in `keys`, like the reference: pass 3 writes the
 * user buffer).  scratch_k must hold n elements. Returns 0, or -1 on bad arguments. */
int vrdx_oracle_sort_keys(uint32_t* keys, uint64_t n, uint32_t* scratch_k) {
  if (n && (!keys || !scratch_k)) return -1;
  for (int pass = 0; pass < 4; ++pass) {
    if (pass % 2 == 0)
      counting_pass(keys, NULL, scratch_k, NULL, n, pass);
    else
      counting_pass(scratch_k, NULL, keys, NULL, n, pass);
  }
  return 0;
}

int vrdx_oracle_sort_key_value(uint32_t* keys, uint32_t* values, uint64_t n, uint32_t* scratch_k,
                               uint32_t* scratch_v) {
  if (n && (!keys || !values || !scratch_k || !scratch_v)) return -1;
  for (int pass = 0; pass < 4; ++pass) {
    if (pass % 2 == 0)
      counting_pass(keys, values, scratch_k, scratch_v, n, pass);
    else
      counting_pass(scratch_k, scratch_v, keys, values, n, pass);
  }
  return 0;
}

/* ---------------------------------------------------------------- structural */

/* One pass = upsweep + spine + downsweep over partitions of 4096 keys.
 * `count` is the device-side element count, `max_count` sizes the grid. */
static int partitioned_pass(const uint32_t* src_k, const uint32_t* src_v, uint32_t* dst_k,
                            uint32_t* dst_v, uint32_t count, uint32_t max_count, int pass,
                            uint32_t* part_hist /* P*256 */, uint32_t* global_hist /* 256 */) {
  const uint32_t P = (max_count + VRDX_PARTITION_SIZE - 1) / VRDX_PARTITION_SIZE;
  const int shift = 8 * pass;
  memset(global_hist, 0, VRDX_RADIX * sizeof(uint32_t)); /* vkCmdFillBuffer, h.in:382 */

  /* upsweep: workgroups whose partition starts past `count` exit (upsweep.slang:20-22) */
  for (uint32_t p = 0; p < P; ++p) {
    uint64_t start = (uint64_t)p * VRDX_PARTITION_SIZE;
    if (start >= count) continue;
    uint32_t* h = part_hist + (size_t)p * VRDX_RADIX;
    memset(h, 0, VRDX_RADIX * sizeof(uint32_t));
    for (uint32_t i = 0; i < VRDX_PARTITION_SIZE; ++i) {
      uint64_t idx = start + i;
      uint32_t key = idx < count ? src_k[idx] : 0xFFFFFFFFu; /* upsweep.slang:32 */
      h[(key >> shift) & 0xFFu]++;
    }
    for (uint32_t d = 0; d < VRDX_RADIX; ++d) global_hist[d] += h[d]; /* upsweep.slang:43 */
  }

  /* spine: partitionCount recomputed from the device count (spine.slang:25) */
  const uint32_t Pc = (uint32_t)(((uint64_t)count + VRDX_PARTITION_SIZE - 1) / VRDX_PARTITION_SIZE);
  for (uint32_t d = 0; d < VRDX_RADIX; ++d) {
    uint32_t running = 0;
    for (uint32_t p = 0; p < Pc; ++p) {
      uint32_t v = part_hist[(size_t)p * VRDX_RADIX + d];
      part_hist[(size_t)p * VRDX_RADIX + d] = running; /* spine.slang:57 */
      running += v;
    }
  }
  {
    uint32_t running = 0;
    for (uint32_t d = 0; d < VRDX_RADIX; ++d) { /* spine.slang:62-83 */
      uint32_t v = global_hist[d];
      global_hist[d] = running;
      running += v;
    }
  }

  /* downsweep */
  uint32_t local_k[VRDX_PARTITION_SIZE];
  uint32_t local_v[VRDX_PARTITION_SIZE];
  uint32_t sorted_k[VRDX_PARTITION_SIZE];
  uint32_t sorted_v[VRDX_PARTITION_SIZE];
  for (uint32_t p = 0; p < P; ++p) {
    uint64_t start = (uint64_t)p * VRDX_PARTITION_SIZE;
    if (start >= count) continue; /* downsweep.slang:58-59 */
    uint32_t local_hist[VRDX_RADIX];
    uint32_t local_start[VRDX_RADIX];
    memset(local_hist, 0, sizeof(local_hist));
    for (uint32_t i = 0; i < VRDX_PARTITION_SIZE; ++i) {
      uint64_t idx = start + i;
      local_k[i] = idx < count ? src_k[idx] : 0xFFFFFFFFu;        /* downsweep.slang:81 */
      local_v[i] = (src_v && idx < count) ? src_v[idx] : 0u;      /* downsweep.slang:85 */
      local_hist[(local_k[i] >> shift) & 0xFFu]++;
    }
    uint32_t running = 0;
    for (uint32_t d = 0; d < VRDX_RADIX; ++d) {
      local_start[d] = running;
      running += local_hist[d];
    }
    /* stable local rank: (wave, round, lane) order == index order (downsweep.slang:79-80) */
    uint32_t cursor[VRDX_RADIX];
    memcpy(cursor, local_start, sizeof(cursor));
    for (uint32_t i = 0; i < VRDX_PARTITION_SIZE; ++i) {
      uint32_t r = cursor[(local_k[i] >> shift) & 0xFFu]++;
      sorted_k[r] = local_k[i]; /* downsweep.slang:188-191 */
      sorted_v[r] = local_v[i]; /* downsweep.slang:211-215 */
    }
    for (uint32_t i = 0; i < VRDX_PARTITION_SIZE; ++i) {
      uint32_t d = (sorted_k[i] >> shift) & 0xFFu;
      /* localHistogramSum[d] = global[d] + partExcl[p][d] - localStart[d]  (downsweep.slang:179-183) */
      uint32_t dst = global_hist[d] + part_hist[(size_t)p * VRDX_RADIX + d] - local_start[d] + i;
      if (dst < count) { /* downsweep.slang:199, 220 */
        dst_k[dst] = sorted_k[i];
        if (src_v) dst_v[dst] = sorted_v[i];
      }
    }
  }
  return 0;
}

/* Structural restatement with the indirect contract.  keys/values hold max_count
 * elements; only [0, count) is sorted, [count, max_count) is left untouched.
 * scratch_* hold max_count elements.  values/scratch_v may be NULL (keys-only). */
int vrdx_oracle_sort_partitioned(uint32_t* keys, uint32_t* values, uint32_t count,
                                 uint32_t max_count, uint32_t* scratch_k, uint32_t* scratch_v) {
  if (count > max_count) return -1; /* forbidden by README.md:175-176 */
  if (max_count == 0) return 0;
  if (!keys || !scratch_k) return -1;
  if (values && !scratch_v) return -1;
  const uint32_t P = (max_count + VRDX_PARTITION_SIZE - 1) / VRDX_PARTITION_SIZE;
  uint32_t* part_hist = (uint32_t*)malloc((size_t)P * VRDX_RADIX * sizeof(uint32_t));
  uint32_t global_hist[VRDX_RADIX];
  if (!part_hist) return -2;
  for (int pass = 0; pass < 4; ++pass) {
    if (pass % 2 == 0)
      partitioned_pass(keys, values, scratch_k, values ? scratch_v : NULL, count, max_count, pass,
                       part_hist, global_hist);
    else
      partitioned_pass(scratch_k, values ? scratch_v : NULL, keys, values, count, max_count, pass,
                       part_hist, global_hist);
  }
  free(part_hist);
  return 0;
}

/* ---------------------------------------------------------------- properties */

/* Size-independent checks used at BASELINE.json's full sizes, where a full CPU sort
 * is too slow for a unit test but an O(n) scan is not. */

/* 1 if keys[0:n] is non-decreasing as unsigned, else 0. */
int vrdx_oracle_is_sorted(const uint32_t* keys, uint64_t n) {
  for (uint64_t i = 1; i < n; ++i)
    if (keys[i - 1] > keys[i]) return 0;
  return 1;
}

/* Order-independent multiset fingerprint of (key, value) pairs: sum and xor of a
 * 64-bit mix of each pair.  values may be NULL.  out[0]=sum, out[1]=xor. */
void vrdx_oracle_multiset_fingerprint(const uint32_t* keys, const uint32_t* values, uint64_t n,
                                      uint64_t out[2]) {
  uint64_t s = 0, x = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t z = ((uint64_t)keys[i] << 32) | (values ? values[i] : 0u);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    s += z;
    x ^= z;
  }
  out[0] = s;
  out[1] = x;
}

/* For key-value output where the input values were the identity permutation
 * (values[i] == i): checks sortedness, that values form a permutation consistent
 * with the original keys (keys_sorted[j] == keys_in[values_sorted[j]]) and
 * stability (equal keys keep ascending original index).  Returns 1 if all hold. */
int vrdx_oracle_check_stable_permutation(const uint32_t* keys_in, const uint32_t* keys_sorted,
                                         const uint32_t* values_sorted, uint64_t n) {
  for (uint64_t j = 0; j < n; ++j) {
    uint32_t src = values_sorted[j];
    if (src >= n) return 0;
    if (keys_in[src] != keys_sorted[j]) return 0;
    if (j) {
      if (keys_sorted[j - 1] > keys_sorted[j]) return 0;
      if (keys_sorted[j - 1] == keys_sorted[j] && values_sorted[j - 1] >= src) return 0;
    }
  }
  return 1;
}
