// vrdx_layout.h — temp-storage carve-up shared by host and device code.
//
// Replaces the reference's storage layout (src/vk_radix_sort.h.in:353-362):
//   reference: [count 16 B][globalHist 4x256][partHist P x 256, P = ceil(N/4096)][keysAlt][valuesAlt]
//   here:      [StorageHeader][status A: T x 256][status B: T x 256][keysAlt][valuesAlt]
// where T = ceil(N / tile) tiles of the pass kernel and a status word is the decoupled
// look-back cell (2 flag bits + 30-bit count).  In reduce-then-scan mode status A holds the
// per-tile digit histograms (full 32-bit counts) and status B the per-chunk spine sums.
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

struct alignas(16) StorageHeader {
  uint32_t element_count[4];             // slot 0: resolved element count (parity with the reference's count slot)
  uint32_t global_hist[kPasses][kRadix];  // digit histograms; exclusive-scanned in place by the histogram kernel
  uint32_t tickets[kPasses];              // dynamic tile ids, one counter per pass
  uint32_t hist_blocks_done;              // last-block-done counter of the histogram / spine kernels
  uint32_t reserved[3];
  uint32_t pass_identity[kPasses];        // 1: every key holds the same digit in this pass -> the pass is a copy
  uint32_t claim_cursor[kRadix];          // keys-only onesweep pass 0: per-digit output cursor tiles claim space from
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
  uint64_t status_bytes;  // per buffer
  uint64_t keys_alt_offset;
  uint64_t values_alt_offset;
  uint64_t inout_bytes;   // Align(4N, 16)
  uint64_t total_keys;    // size for keys-only
  uint64_t total_kv;      // size for key-value
};

// `min_tile` = smallest tile (keys per CTA) any kernel of this sorter may use for this count,
// so the status buffers are large enough whichever kernel configuration runs.
inline StorageLayout ComputeLayout(uint64_t max_count, uint32_t min_tile) {
  StorageLayout l{};
  const uint64_t tiles = CeilDiv(max_count, min_tile) + 1;
  l.header_offset = 0;
  l.status_a_offset = AlignUp(sizeof(StorageHeader), kOffsetAlignment);
  l.status_bytes = AlignUp(tiles * kRadix * sizeof(uint32_t), kOffsetAlignment);
  l.status_b_offset = l.status_a_offset + l.status_bytes;
  l.inout_bytes = AlignUp(max_count * sizeof(uint32_t), kOffsetAlignment);
  l.keys_alt_offset = l.status_b_offset + l.status_bytes;
  l.values_alt_offset = l.keys_alt_offset + l.inout_bytes;
  l.total_keys = l.values_alt_offset;
  l.total_kv = l.values_alt_offset + l.inout_bytes;
  return l;
}

}  // namespace vrdx
