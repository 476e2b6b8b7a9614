// vrdx_layout.h — temp-storage carve-up shared by host and device code.
//
// Replaces the reference's storage layout (src/vk_radix_sort.h.in:353-362):
//   reference: [count 16 B][globalHist 4x256][partHist P x 256, P = ceil(N/4096)][keysAlt][valuesAlt]
//   here:      [StorageHeader][table A: R x 256][table B: R x 256][keysAlt][valuesAlt]
// Onesweep (small and medium N): both tables hold one row of 256 decoupled look-back cells per tile (2 flag
// bits + 30-bit count), alternating between passes.  Reduce-then-scan (large N): table A holds one row of
// 256 SIXTEEN-bit in-chunk prefixes per tile (half of the reference's partHist), table B one 32-bit row per
// chunk of 8 tiles + the spine's segment sums + one flag byte per tile.  2^28 keys: 1,120 MB in total
// (reference: 1,141 MB).
// All offsets are multiples of 16 bytes and all size arithmetic is 64-bit.
#pragma once
#include <stdint.h>

namespace vrdx {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kPasses = 4;

// Look-back cell: [31:30] flag, [29:0] value.
constexpr uint32_t kStatusValueMask = (1u << 30) - 1u;
constexpr uint32_t kStatusAggregate = 1u << 30;  // tile-local digit count is available
constexpr uint32_t kStatusPrefix = 2u << 30;     // inclusive prefix over tiles [0, t] is available
constexpr uint64_t kMaxOnesweepCount = 1ull << 30;

constexpr uint32_t kOffsetAlignment = 16;  // the reference's minStorageBufferOffsetAlignment "usually 16"
#ifndef VRDX_TABLE_ALIGN
#define VRDX_TABLE_ALIGN 256
#endif
constexpr uint32_t kTableAlignment = VRDX_TABLE_ALIGN;  // of the tables / scratch halves inside the storage

constexpr uint32_t kSpineChunkTiles = 8;     // reduce-then-scan: tiles per upsweep CTA / spine row
constexpr uint32_t kSpineSegmentRows = 128;  // spine: segment sums kept after the chunk rows (one co-resident CTA each)

struct alignas(16) StorageHeader {
  uint32_t element_count[4];             // slot 0: resolved element count (parity with the reference's count slot)
  uint32_t global_hist[kPasses][kRadix];  // digit histograms; exclusive-scanned in place by the histogram kernel
  uint32_t tickets[kPasses];              // dynamic tile ids, one counter per pass
  uint32_t hist_blocks_done;              // last-block-done counter of the histogram / spine kernels
  uint32_t reserved[3];
  uint32_t pass_identity[kPasses];        // 1: every key holds the same digit in this pass -> the pass is a copy
  uint32_t claim_cursor[kRadix];          // keys-only onesweep pass 0: per-digit output cursor tiles claim space from
  uint32_t spine_flags[kSpineSegmentRows];  // reduce-then-scan: segment s of this pass's spine has published its sums
};
static_assert(sizeof(StorageHeader) % 16 == 0, "header must keep 16-byte alignment");

#if defined(__CUDACC__)
#define VRDX_HD __host__ __device__
#else
#define VRDX_HD
#endif
VRDX_HD inline constexpr uint64_t AlignUp(uint64_t a, uint64_t b) { return (a + b - 1) / b * b; }
VRDX_HD inline constexpr uint64_t CeilDiv(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

struct StorageLayout {
  uint64_t header_offset;
  uint64_t status_a_offset;
  uint64_t status_b_offset;
  uint64_t status_a_bytes, status_b_bytes;
  uint64_t keys_alt_offset;
  uint64_t values_alt_offset;
  uint64_t inout_bytes;   // Align(4N, 16)
  uint64_t total_keys;    // size for keys-only
  uint64_t total_kv;      // size for key-value
};

// `min_tile` = smallest tile (keys per CTA) any kernel of this sorter may use, so the tables are large
// enough whichever kernel configuration runs.
// `onesweep_below` = counts below this may be sorted by onesweep (AUTO: the measured crossover; ~0: every
// count; 0: none).  Onesweep keeps two tables of 32-bit look-back cells, one row (256 cells) per tile.
// Reduce-then-scan keeps table A = one row of 16-BIT in-chunk prefixes per tile and table B = one 32-bit row
// per chunk of 8 tiles + the spine's segment sums + one flag byte per tile.  `wide_rows`: the round-1 kernels
// (VRDX_EXPERIMENTS) keep 32-bit rows in both.  Every term is non-decreasing in max_count: storage sized for maxElementCount also
// serves any smaller count (the reference's formula is monotone too, h.in:279-308).
inline StorageLayout ComputeLayout(uint64_t max_count, uint32_t min_tile, uint64_t onesweep_below = ~0ull,
                                   bool wide_rows = false) {
  StorageLayout l{};
  const uint64_t row = kRadix * sizeof(uint32_t);
  const uint64_t tiles = CeilDiv(max_count, min_tile) + 1;
  const uint64_t one_count = onesweep_below == 0 ? 0 : (max_count < onesweep_below ? max_count : onesweep_below - 1);
  const uint64_t one_bytes = one_count ? (CeilDiv(one_count, min_tile) + 1) * row : 0;
  const uint64_t rts_a = wide_rows ? tiles * row : tiles * (row / 2);
  // (+ one byte per tile after the segment sums: the upsweep's "all keys agree below the digit" tile flags)
  const uint64_t spine_rows = kSpineSegmentRows + 1 + CeilDiv(tiles, row);  // segment sums + the flag bytes
  const uint64_t rts_b = ((wide_rows ? tiles : CeilDiv(tiles, kSpineChunkTiles)) + spine_rows) * row;
  // Tables and scratch halves start on kTableAlignment (256 B) boundaries of the storage: a warp's 128-byte row of
  // the scratch keys / values is then one L1 line when the caller's storage is itself 256-byte aligned.
  l.header_offset = 0;
  l.status_a_offset = AlignUp(sizeof(StorageHeader), kTableAlignment);
  l.status_a_bytes = AlignUp(one_bytes > rts_a ? one_bytes : rts_a, kTableAlignment);
  l.status_b_bytes = AlignUp(one_bytes > rts_b ? one_bytes : rts_b, kTableAlignment);
  l.status_b_offset = l.status_a_offset + l.status_a_bytes;
  l.inout_bytes = AlignUp(max_count * sizeof(uint32_t), kTableAlignment);
  l.keys_alt_offset = l.status_b_offset + l.status_b_bytes;
  l.values_alt_offset = l.keys_alt_offset + l.inout_bytes;
  l.total_keys = l.values_alt_offset;
  l.total_kv = l.values_alt_offset + l.inout_bytes;
  return l;
}

}  // namespace vrdx
