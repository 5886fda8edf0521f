// common.cuh — shared device helpers for libb200sql (sm_100a).
// Everything here is hand-written CUDA; no CUB/Thrust/cuDF.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "b200sql.h"

#define B2_BLOCK 256
#define B2_WARPS (B2_BLOCK / 32)
#define FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------
extern thread_local char g_b2_err[512];
static inline int32_t b2_fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_b2_err, sizeof(g_b2_err), fmt, ap);
  va_end(ap);
  return code;
}
#define B2_CUDA_TRY(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return b2_fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                              \
  } while (0)
#define B2_CHECK_LAUNCH(name)                                                          \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return b2_fail(B2_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)
#define B2_REQUIRE(cond, msg)                                                          \
  do {                                                                                 \
    if (!(cond)) return b2_fail(B2_ERR_ARG, "%s: %s", __func__, msg);                  \
  } while (0)

// persistent grid: one resident wave of CTAs (sm_count x occupancy), rows are grid-strided.
int b2_sm_count();
template <class K>
static inline int b2_wave_grid(K kernel, int block, int64_t work_items_blocks) {
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0);
  if (occ < 1) occ = 1;
  int64_t g = (int64_t)b2_sm_count() * occ;
  if (g > work_items_blocks) g = work_items_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t b2_mix64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ bool b2_bit(const uint8_t* __restrict__ bm, int64_t i) {
  return (bm[i >> 3] >> (i & 7)) & 1;
}

// L2 eviction policies (sm_80+ createpolicy; plain ld only takes .L2::evict_* on 256-bit vectors).
// Non-volatile asm without inputs: the compiler hoists/CSEs it, one instruction per kernel.
__device__ __forceinline__ uint64_t b2_policy_stream() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t b2_policy_keep() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// streaming 8-byte load: read-only path, do not pollute L1, first-out of L2 so that the
// L2-resident lookup / group tables survive the scan.
__device__ __forceinline__ int64_t b2_ld_stream(const int64_t* p) {
  int64_t v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(b2_policy_stream()));
  return v;
}
// table load that should stay in L2
__device__ __forceinline__ int32_t b2_ld_keep_i32(const int32_t* p) {
  int32_t v;
  asm("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(b2_policy_keep()));
  return v;
}
__device__ __forceinline__ int64_t b2_ld_cg_i64(const int64_t* p) {
  return __ldcg(reinterpret_cast<const long long*>(p));
}

__host__ __device__ __forceinline__ int64_t b2_ordered_from_bits(int64_t b) {
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

__device__ __forceinline__ bool b2_cmp_i(int op, int64_t a, int64_t b) {
  switch (op) {
    case B2_EQ: return a == b;
    case B2_NE: return a != b;
    case B2_LT: return a < b;
    case B2_LE: return a <= b;
    case B2_GT: return a > b;
    default: return a >= b;
  }
}
__device__ __forceinline__ bool b2_cmp_f(int op, double a, double b) {
  switch (op) {
    case B2_EQ: return a == b;
    case B2_NE: return a != b;   // IEEE: true when either is NaN (numpy semantics)
    case B2_LT: return a < b;
    case B2_LE: return a <= b;
    case B2_GT: return a > b;
    default: return a >= b;
  }
}

// raw 64-bit load of any column type (U8 widened to 0/1)
__device__ __forceinline__ int64_t b2_load_raw(const b2_col_t& c, int64_t row) {
  if (c.dtype == B2_U8) return (int64_t) reinterpret_cast<const uint8_t*>(c.data)[row];
  return b2_ld_stream(reinterpret_cast<const int64_t*>(c.data) + row);
}
// NULL test in the pandas sense: bitmap bit clear, or NaN in a float column
__device__ __forceinline__ bool b2_is_null(const b2_col_t& c, int64_t row, int64_t raw) {
  if (c.valid && !b2_bit(c.valid, row)) return true;
  if (c.dtype == B2_F64) { double d = __longlong_as_double(raw); return d != d; }
  return false;
}

// Evaluate all predicate terms for R rows of one lane: rows row0 + j*32, j < R.
// Returns a bitmask (bit j = row j passes).  Loads of one term are issued back-to-back
// (R independent 8-byte loads in flight per lane) before they are consumed.
template <int R>
__device__ __forceinline__ uint32_t b2_eval_terms(const b2_scan_t& s, int64_t row0) {
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < R; ++j)
    if (row0 + (int64_t)j * 32 < s.n) bits |= 1u << j;
  for (int t = 0; t < s.nterms; ++t) {
    const b2_term_t& tm = s.terms[t];
    const b2_col_t& c = s.cols[tm.col];
    int64_t raw[R];
    if (c.dtype == B2_U8) {
      const uint8_t* p = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
      for (int j = 0; j < R; ++j) raw[j] = (bits >> j) & 1 ? (int64_t)p[row0 + (int64_t)j * 32] : 0;
    } else {
      const int64_t* p = reinterpret_cast<const int64_t*>(c.data);
#pragma unroll
      for (int j = 0; j < R; ++j) raw[j] = (bits >> j) & 1 ? b2_ld_stream(p + row0 + (int64_t)j * 32) : 0;
    }
    uint32_t ok = 0;
    const int op = tm.op;
    if (op == B2_IS_NULL || op == B2_IS_NOT_NULL) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if ((bits >> j) & 1) {
          bool isn = b2_is_null(c, row0 + (int64_t)j * 32, raw[j]);
          ok |= (uint32_t)(isn == (op == B2_IS_NULL)) << j;
        }
      }
    } else {
      uint32_t vbits = bits;
      if (c.valid) {
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (((bits >> j) & 1) && !b2_bit(c.valid, row0 + (int64_t)j * 32)) vbits &= ~(1u << j);
      }
      if (op == B2_IS_TRUE) {
#pragma unroll
        for (int j = 0; j < R; ++j) ok |= (uint32_t)(raw[j] != 0) << j;
      } else if (c.dtype == B2_F64) {
        const double lit = tm.lit_f;
#pragma unroll
        for (int j = 0; j < R; ++j) ok |= (uint32_t)b2_cmp_f(op, __longlong_as_double(raw[j]), lit) << j;
      } else if (tm.as_f64) {
        const double lit = tm.lit_f;
#pragma unroll
        for (int j = 0; j < R; ++j) ok |= (uint32_t)b2_cmp_f(op, (double)raw[j], lit) << j;
      } else {
        const int64_t lit = tm.lit_i;
#pragma unroll
        for (int j = 0; j < R; ++j) ok |= (uint32_t)b2_cmp_i(op, raw[j], lit) << j;
      }
      ok &= vbits;
    }
    bits &= ok;
  }
  return bits;
}

// ---------------------------------------------------------------------------------------
// aggregate updates
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void b2_atomic_update(int op, int dtype, void* acc, int64_t slot, int64_t raw) {
  if (op == B2_AGG_SUM) {
    if (dtype == B2_F64)
      atomicAdd(reinterpret_cast<double*>(acc) + slot, __longlong_as_double(raw));
    else
      atomicAdd(reinterpret_cast<unsigned long long*>(acc) + slot, (unsigned long long)raw);
  } else if (op == B2_AGG_SUMF) {
    double d = dtype == B2_F64 ? __longlong_as_double(raw) : (double)raw;
    atomicAdd(reinterpret_cast<double*>(acc) + slot, d);
  } else if (op == B2_AGG_MIN) {
    long long v = dtype == B2_F64 ? b2_ordered_from_bits(raw) : raw;
    atomicMin(reinterpret_cast<long long*>(acc) + slot, v);
  } else if (op == B2_AGG_MAX) {
    long long v = dtype == B2_F64 ? b2_ordered_from_bits(raw) : raw;
    atomicMax(reinterpret_cast<long long*>(acc) + slot, v);
  }
}

// For R rows of one lane with resolved slots (slot < 0 = row does not contribute):
// load each aggregate's input column (batched), skip NULLs, apply atomics.
template <int R>
__device__ __forceinline__ void b2_apply_aggs(const b2_scan_t& s, const b2_agg_t* __restrict__ aggs,
                                              int naggs, const b2_aggstate_t& st, int64_t row0,
                                              const int64_t (&slot)[R]) {
  if (st.out_slot) {
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (row0 + (int64_t)j * 32 < s.n) st.out_slot[row0 + (int64_t)j * 32] = (int32_t)slot[j];
  }
  if (st.rows) {
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (slot[j] >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(st.rows) + slot[j], 1ULL);
  }
  if (st.present) {
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (slot[j] >= 0) {
        uint32_t w = (uint32_t)(slot[j] >> 5), b = 1u << (slot[j] & 31);
        // read first: after warm-up almost every group is already marked, so the atomic is rare
        if (!(__ldcg(st.present + w) & b)) atomicOr(st.present + w, b);
      }
  }
  for (int a = 0; a < naggs; ++a) {
    const b2_agg_t ag = aggs[a];
    if (ag.col < 0) continue;  // COUNT(*) is st.rows
    const b2_col_t& c = s.cols[ag.col];
    int64_t raw[R];
#pragma unroll
    for (int j = 0; j < R; ++j) raw[j] = slot[j] >= 0 ? b2_load_raw(c, row0 + (int64_t)j * 32) : 0;
    void* acc = st.acc[a];
    int64_t* cnt = st.cnt[a];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (slot[j] < 0) continue;
      if (b2_is_null(c, row0 + (int64_t)j * 32, raw[j])) continue;
      if (acc) b2_atomic_update(ag.op, c.dtype, acc, slot[j], raw[j]);
      if (cnt) atomicAdd(reinterpret_cast<unsigned long long*>(cnt) + slot[j], 1ULL);
    }
  }
}

struct b2_aggs_arg {  // aggs passed by value in kernel params
  b2_agg_t a[B2_MAX_AGGS];
  int32_t n;
};
