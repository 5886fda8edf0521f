// pipeline.cuh — TMA-staged column tiles.
//
// The streaming side of every aggregation kernel (predicate columns, keys, aggregate inputs) is
// read exactly once, so it is moved by the TMA engine instead of by the compute warps:
//   * one producer warp per CTA issues `cp.async.bulk` (1-D bulk copy, global -> shared) for the
//     next tiles of every referenced column, completion counted on an mbarrier (full[stage]);
//   * 8 consumer warps wait on full[stage], evaluate terms / look up keys / update aggregates out
//     of shared memory (conflict-free 8-byte LDS), then arrive on empty[stage].
// With S stages of (ncols x 16 KB) per CTA the SM always has >= 100 KB of column data in flight
// without a single compute warp stalled on a streaming load; the consumers only ever wait for the
// random accesses (join lookup, group-table atomics).  Tiles are copied with an L2 evict_first
// policy so that the L2-resident lookup / group tables survive the scan.
#pragma once
#include "common.cuh"
#include <stdlib.h>

#define B2_PIPE_R 8
#define B2_PIPE_TILE (B2_BLOCK * B2_PIPE_R)   // 2048 rows per tile
#define B2_PIPE_THREADS (B2_BLOCK + 32)       // 8 consumer warps + 1 producer warp
#define B2_PIPE_MAX_STAGES 4
#define B2_PIPE_SMEM_BUDGET (100 * 1024)      // per CTA: keeps >= 2 CTAs per SM for <= 3 columns

struct b2_pipe_t {  // computed on the host per launch
  int32_t col_off[B2_MAX_COLS];
  int32_t col_bytes[B2_MAX_COLS];
  int32_t stage_bytes;
  int32_t stages;
  int32_t enabled;
  int32_t smem_bytes;
};

// host: lay out one stage; disable the pipeline when a buffer is not 16-byte aligned or the
// partition is too small to be worth it.
static inline void b2_make_pipe(const b2_scan_t& s, b2_pipe_t* pp) {
  memset(pp, 0, sizeof(*pp));
  int off = 0;
  // Opt-in (B200SQL_PIPELINE=1): measured on B200 (profiles/r01_ncu_notes.md, capture C) the staged
  // path currently loses to the direct path because shared memory caps it at 2 CTAs/SM while the
  // consumers are still issue-bound; the direct path runs 5 CTAs/SM.
  static const bool want = [] { const char* e = getenv("B200SQL_PIPELINE"); return e && e[0] == '1'; }();
  bool ok = want && s.ncols > 0 && s.n >= 4 * (int64_t)B2_PIPE_TILE;
  for (int c = 0; c < s.ncols; ++c) {
    const int w = s.cols[c].dtype == B2_U8 ? 1 : 8;
    pp->col_off[c] = off;
    pp->col_bytes[c] = w * B2_PIPE_TILE;
    off += (w * B2_PIPE_TILE + 127) & ~127;
    if (reinterpret_cast<uintptr_t>(s.cols[c].data) & 15) ok = false;
  }
  pp->stage_bytes = off;
  int stages = off > 0 ? B2_PIPE_SMEM_BUDGET / off : 0;
  if (stages > B2_PIPE_MAX_STAGES) stages = B2_PIPE_MAX_STAGES;
  if (stages < 2) {  // many columns: take what one SM offers
    stages = off > 0 ? (200 * 1024) / off : 0;
    if (stages > 2) stages = 2;
  }
  if (stages < 2) ok = false;
  pp->stages = stages;
  pp->enabled = ok ? 1 : 0;
  pp->smem_bytes = ok ? stages * off + 2 * stages * 8 + 128 : 0;
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t b2_smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void b2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b2_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void b2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b2_smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void b2_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b2_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void b2_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = b2_smem_addr(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk copy global -> shared through the TMA engine, completion on `bar`
__device__ __forceinline__ void b2_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(b2_smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(b2_smem_addr(bar)), "l"(policy)
      : "memory");
}

// ---- the tile loop ---------------------------------------------------------------------------------
// body(ld) is called by the 8 consumer warps once per 2048-row tile with a loader that reads the
// tile from shared memory; the ragged tail (< 2048 rows) is handled once with the global loader.
template <class Body>
__device__ __forceinline__ void b2_tile_pipeline(const b2_scan_t& s, const b2_pipe_t& pp, Body&& body) {
  extern __shared__ __align__(128) uint8_t b2_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(b2_smem + (size_t)pp.stages * pp.stage_bytes);
  uint64_t* empty = full + pp.stages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nfull = s.n / B2_PIPE_TILE;
  if (tid == 0) {
    for (int st = 0; st < pp.stages; ++st) {
      b2_mbar_init(&full[st], 1);
      b2_mbar_init(&empty[st], B2_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t my_count = nfull > blockIdx.x ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp == B2_WARPS) {
    // ---------------- producer warp: one elected lane feeds the ring
    if (lane == 0) {
      const uint64_t pol = b2_policy_stream();
      uint32_t tx = 0;
      for (int c = 0; c < s.ncols; ++c) tx += (uint32_t)pp.col_bytes[c];
      int st = 0;
      uint32_t round = 0;
      for (int64_t k = 0; k < my_count; ++k) {
        if (round > 0) b2_mbar_wait(&empty[st], (round - 1) & 1);
        const int64_t tile = blockIdx.x + k * gridDim.x;
        uint8_t* dst = b2_smem + (size_t)st * pp.stage_bytes;
        b2_mbar_expect_tx(&full[st], tx);
        for (int c = 0; c < s.ncols; ++c) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(s.cols[c].data) + tile * (int64_t)pp.col_bytes[c];
          b2_bulk_g2s(dst + pp.col_off[c], src, (uint32_t)pp.col_bytes[c], &full[st], pol);
        }
        if (++st == pp.stages) { st = 0; ++round; }
      }
    }
  } else {
    // ---------------- consumer warps
    int st = 0;
    uint32_t round = 0;
    const int tile_off = warp * (32 * B2_PIPE_R) + lane;
    for (int64_t k = 0; k < my_count; ++k) {
      const int64_t tile = blockIdx.x + k * gridDim.x;
      b2_mbar_wait(&full[st], round & 1);
      const b2_sld ld{&s, tile * B2_PIPE_TILE + tile_off, b2_smem + (size_t)st * pp.stage_bytes, pp.col_off,
                      tile_off};
      body(ld);
      __syncwarp();
      if (lane == 0) b2_mbar_arrive(&empty[st]);
      if (++st == pp.stages) { st = 0; ++round; }
    }
    if (nfull * B2_PIPE_TILE < s.n && (int64_t)blockIdx.x == nfull % gridDim.x) {
      const b2_gld ld{&s, nfull * B2_PIPE_TILE + tile_off};
      body(ld);
    }
  }
}

// the same body over a plain grid-stride loop (small or unaligned partitions)
template <int R, class Body>
__device__ __forceinline__ void b2_tile_direct(const b2_scan_t& s, Body&& body) {
  const int tile_off = (threadIdx.x >> 5) * (32 * R) + (threadIdx.x & 31);
  const int64_t tile = (int64_t)B2_BLOCK * R;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < s.n; base += (int64_t)gridDim.x * tile) {
    const b2_gld ld{&s, base + tile_off};
    body(ld);
  }
}

// Tiles handed out IN ORDER by a global ticket instead of a fixed stride per CTA: the rows being
// processed at any moment then form one contiguous window of (resident CTAs x tile) rows no matter how
// unevenly the CTAs progress.  With a fixed stride a CTA that runs 10 % faster is 10 % of the input
// ahead by the end -- fatal when the input is ordered so that a window should touch one L2-sized slice
// of a table (partition.cuh).  *ticket must be 0 at launch.
template <int R, class Body>
__device__ __forceinline__ void b2_tile_ticket(const b2_scan_t& s, unsigned long long* __restrict__ ticket, Body&& body) {
  __shared__ long long sh_tile;
  const int tile_off = (threadIdx.x >> 5) * (32 * R) + (threadIdx.x & 31);
  const int64_t tile = (int64_t)B2_BLOCK * R;
  const int64_t ntiles = (s.n + tile - 1) / tile;
  for (;;) {
    if (threadIdx.x == 0) sh_tile = (long long)atomicAdd(ticket, 1ULL);
    __syncthreads();
    const int64_t t = sh_tile;
    __syncthreads();
    if (t >= ntiles) break;
    const b2_gld ld{&s, t * tile + tile_off};
    body(ld);
  }
}

// host: launch geometry of a pipelined kernel.  Occupancy is limited by shared memory.
template <class K>
static inline int b2_pipe_grid(K kernel, const b2_pipe_t& pp, int64_t n) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pp.smem_bytes);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, B2_PIPE_THREADS, pp.smem_bytes);
  if (occ < 1) occ = 1;
  int64_t g = (int64_t)b2_sm_count() * occ;
  const int64_t tiles = (n + B2_PIPE_TILE - 1) / B2_PIPE_TILE;
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}
