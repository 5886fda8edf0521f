// join.cuh — hash join: chained build (duplicates allowed), direct-address build (unique dense
// keys), two-pass probe that emits (probe row, build row) pairs in probe-row order.
#pragma once
#include "common.cuh"

#define B2_JOIN_R 16
static_assert(B2_BLOCK * B2_JOIN_R == B2_TILE, "tile geometry");

struct b2_keycols_arg {
  b2_col_t c[B2_MAX_KEYS];
  int32_t n;
};
struct b2_probekeys_arg {
  int32_t cols[B2_MAX_KEYS];
};

// normalised key image: -0.0 -> +0.0 so float keys compare like pandas; ints unchanged
__device__ __forceinline__ int64_t b2_key_image(const b2_col_t& c, int64_t raw) {
  if (c.dtype == B2_F64 && raw == (int64_t)0x8000000000000000LL) return 0;
  return raw;
}
__device__ __forceinline__ uint64_t b2_hash_keys(const int64_t* key, int nkeys) {
  uint64_t hv = 0x9e3779b97f4a7c15ULL;
  for (int k = 0; k < nkeys; ++k) hv = b2_mix64(hv ^ (uint64_t)key[k]);
  return hv;
}

__global__ void __launch_bounds__(B2_BLOCK)
b2_join_build_kernel(const __grid_constant__ b2_keycols_arg keys, int64_t n, int32_t* __restrict__ head,
                     int32_t* __restrict__ next, int64_t cap) {
  for (int64_t i = (int64_t)blockIdx.x * B2_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * B2_BLOCK) {
    int64_t key[B2_MAX_KEYS];
    bool isnull = false;
#pragma unroll
    for (int k = 0; k < B2_MAX_KEYS; ++k) {
      key[k] = 0;
      if (k < keys.n) {
        const int64_t raw = b2_load_raw(keys.c[k], i);
        isnull |= b2_is_null(keys.c[k], i, raw);
        key[k] = b2_key_image(keys.c[k], raw);
      }
    }
    if (isnull) { next[i] = -1; continue; }  // NULL keys never join (join.py:202-213)
    const uint64_t h = b2_hash_keys(key, keys.n) & (uint64_t)(cap - 1);
    next[i] = atomicExch(head + h, (int32_t)i);
  }
}

__global__ void __launch_bounds__(B2_BLOCK)
b2_join_build_dense_kernel(const __grid_constant__ b2_col_t key, int64_t n, int64_t kmin, int64_t range,
                           int32_t* __restrict__ lookup, int32_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * B2_BLOCK + threadIdx.x; i < n; i += (int64_t)gridDim.x * B2_BLOCK) {
    const int64_t raw = b2_load_raw(key, i);
    if (b2_is_null(key, i, raw)) continue;
    const uint64_t d = (uint64_t)raw - (uint64_t)kmin;
    if (d >= (uint64_t)range) continue;
    if (atomicExch(lookup + d, (int32_t)i) != -1) flags[0] = 1;
  }
}

// Walk the matches of one probe row.  F(build_row) is called for every match.
template <class F>
__device__ __forceinline__ int b2_for_matches(const b2_jointable_t& jt, const int64_t* pkey, F f) {
  int cnt = 0;
  if (jt.dense) {
    const uint64_t d = (uint64_t)pkey[0] - (uint64_t)jt.kmin;
    if (d < (uint64_t)jt.range) {
      const int32_t r = __ldg(jt.lookup + d);
      if (r >= 0) { f(r); cnt = 1; }
    }
    return cnt;
  }
  const uint64_t h = b2_hash_keys(pkey, jt.nkeys) & (uint64_t)(jt.cap - 1);
  for (int32_t r = __ldg(jt.head + h); r >= 0; r = __ldg(jt.next + r)) {
    bool same = true;
    for (int k = 0; same && k < jt.nkeys; ++k) {
      const int64_t braw = __ldg(reinterpret_cast<const long long*>(jt.keys[k].data) + r);
      same = b2_key_image(jt.keys[k], braw) == pkey[k];
    }
    if (same) { f(r); ++cnt; }
  }
  return cnt;
}

// load + normalise the probe keys of one row; returns false if any key is NULL
__device__ __forceinline__ bool b2_probe_key(const b2_scan_t& s, const b2_probekeys_arg& pk, int nkeys,
                                             int64_t row, int64_t* key) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < B2_MAX_KEYS; ++k) {
    key[k] = 0;
    if (k < nkeys) {
      const b2_col_t& c = s.cols[pk.cols[k]];
      const int64_t raw = b2_load_raw(c, row);
      ok &= !b2_is_null(c, row, raw);
      key[k] = b2_key_image(c, raw);
    }
  }
  return ok;
}

__device__ __forceinline__ int b2_emit_count(int mode, int matches) {
  switch (mode) {
    case B2_JOIN_INNER: return matches;
    case B2_JOIN_LEFT: return matches ? matches : 1;
    case B2_JOIN_SEMI: return matches ? 1 : 0;
    default: return matches ? 0 : 1;  // ANTI
  }
}

__global__ void __launch_bounds__(B2_BLOCK)
b2_join_count_kernel(const __grid_constant__ b2_scan_t s, const __grid_constant__ b2_probekeys_arg pk,
                     const __grid_constant__ b2_jointable_t jt, int mode, int64_t ntiles,
                     int64_t* __restrict__ tile_cnt) {
  __shared__ int64_t sh[B2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * B2_TILE + (int64_t)warp * (32 * B2_JOIN_R) + lane;
    const uint32_t bits = b2_eval_terms<B2_JOIN_R>(s, row0);
    int64_t c = 0;
#pragma unroll 4
    for (int j = 0; j < B2_JOIN_R; ++j) {
      if (!((bits >> j) & 1)) continue;
      int64_t key[B2_MAX_KEYS];
      int m = 0;
      if (b2_probe_key(s, pk, jt.nkeys, row0 + (int64_t)j * 32, key))
        m = b2_for_matches(jt, key, [](int32_t) {});
      c += b2_emit_count(mode, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if (lane == 0) sh[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int64_t t = 0;
      for (int w = 0; w < B2_WARPS; ++w) t += sh[w];
      tile_cnt[tile] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(B2_BLOCK)
b2_join_write_kernel(const __grid_constant__ b2_scan_t s, const __grid_constant__ b2_probekeys_arg pk,
                     const __grid_constant__ b2_jointable_t jt, int mode, int64_t ntiles,
                     const int64_t* __restrict__ tile_off, int32_t* __restrict__ out_probe,
                     int32_t* __restrict__ out_build, uint8_t* __restrict__ build_matched) {
  __shared__ int64_t sh[B2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * B2_TILE + (int64_t)warp * (32 * B2_JOIN_R) + lane;
    const uint32_t bits = b2_eval_terms<B2_JOIN_R>(s, row0);
    // pass 1 (registers): emit counts per row, remember the first match (the common unique case)
    int cnt[B2_JOIN_R];
    int32_t first[B2_JOIN_R];
    int64_t wtotal = 0;
#pragma unroll
    for (int j = 0; j < B2_JOIN_R; ++j) {
      cnt[j] = 0;
      first[j] = -1;
      if ((bits >> j) & 1) {
        int64_t key[B2_MAX_KEYS];
        int m = 0;
        int32_t f0 = -1;
        if (b2_probe_key(s, pk, jt.nkeys, row0 + (int64_t)j * 32, key))
          m = b2_for_matches(jt, key, [&](int32_t r) { if (f0 < 0) f0 = r; });
        first[j] = f0;
        cnt[j] = b2_emit_count(mode, m) | (m > 1 ? 0x40000000 : 0);  // flag: chain must be re-walked
      }
      wtotal += cnt[j] & 0x3fffffff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wtotal += __shfl_xor_sync(FULL_MASK, wtotal, o);
    if (lane == 0) sh[warp] = wtotal;
    __syncthreads();
    int64_t off = tile_off[tile];
    for (int w = 0; w < warp; ++w) off += sh[w];
    __syncthreads();
    // pass 2: in-order output positions through a warp prefix sum per row group
#pragma unroll
    for (int j = 0; j < B2_JOIN_R; ++j) {
      const int c = cnt[j] & 0x3fffffff;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(FULL_MASK, incl, 31);
      int64_t pos = off + incl - c;
      off += total;
      if (c == 0) continue;
      const int32_t prow = (int32_t)(row0 + (int64_t)j * 32);
      if (mode == B2_JOIN_SEMI || mode == B2_JOIN_ANTI) {
        out_probe[pos] = prow;
        if (out_build) out_build[pos] = -1;
        if (build_matched && first[j] >= 0) build_matched[first[j]] = 1;
      } else if (!(cnt[j] & 0x40000000)) {
        out_probe[pos] = prow;
        out_build[pos] = first[j];  // -1 for an unmatched LEFT row
        if (build_matched && first[j] >= 0) build_matched[first[j]] = 1;
      } else {
        int64_t key[B2_MAX_KEYS];
        b2_probe_key(s, pk, jt.nkeys, prow, key);
        b2_for_matches(jt, key, [&](int32_t r) {
          out_probe[pos] = prow;
          out_build[pos] = r;
          if (build_matched) build_matched[r] = 1;
          ++pos;
        });
      }
    }
  }
}

extern "C" {

int32_t b2_join_build(const b2_col_t* keys, int32_t nkeys, int64_t n, int32_t* head, int32_t* next,
                      int64_t cap, void* stream) {
  B2_REQUIRE(keys && head && (next || n == 0), "null argument");
  B2_REQUIRE(nkeys >= 1 && nkeys <= B2_MAX_KEYS, "bad nkeys");
  B2_REQUIRE(b2_pow2(cap), "cap must be a power of two");
  B2_REQUIRE(n < ((int64_t)1 << 31), "build side must hold < 2^31 rows");
  if (n <= 0) return B2_OK;
  b2_keycols_arg ka;
  memset(&ka, 0, sizeof(ka));
  ka.n = nkeys;
  for (int k = 0; k < nkeys; ++k) {
    B2_REQUIRE(keys[k].dtype != B2_U8, "join keys must be 64-bit columns");
    ka.c[k] = keys[k];
  }
  int grid = b2_wave_grid(b2_join_build_kernel, B2_BLOCK, (n + B2_BLOCK - 1) / B2_BLOCK);
  b2_join_build_kernel<<<grid, B2_BLOCK, 0, (cudaStream_t)stream>>>(ka, n, head, next, cap);
  B2_CHECK_LAUNCH("b2_join_build_kernel");
  return B2_OK;
}

int32_t b2_join_build_dense(const b2_col_t* key, int64_t n, int64_t kmin, int64_t range, int32_t* lookup,
                            int32_t* d_flags, void* stream) {
  B2_REQUIRE(key && lookup && d_flags, "null argument");
  B2_REQUIRE(key->dtype == B2_I64, "dense join needs an int64 key");
  B2_REQUIRE(range > 0, "bad range");
  if (n <= 0) return B2_OK;
  int grid = b2_wave_grid(b2_join_build_dense_kernel, B2_BLOCK, (n + B2_BLOCK - 1) / B2_BLOCK);
  b2_join_build_dense_kernel<<<grid, B2_BLOCK, 0, (cudaStream_t)stream>>>(*key, n, kmin, range, lookup, d_flags);
  B2_CHECK_LAUNCH("b2_join_build_dense_kernel");
  return B2_OK;
}

static int32_t b2_check_join(const b2_scan_t* scan, const int32_t* probe_keys, const b2_jointable_t* jt,
                             int32_t mode, b2_probekeys_arg* pk) {
  int32_t rc = b2_check_scan(scan);
  if (rc) return rc;
  B2_REQUIRE(probe_keys && jt, "null argument");
  B2_REQUIRE(jt->nkeys >= 1 && jt->nkeys <= B2_MAX_KEYS, "bad nkeys");
  B2_REQUIRE(mode >= B2_JOIN_INNER && mode <= B2_JOIN_ANTI, "bad join mode");
  memset(pk, 0, sizeof(*pk));
  for (int k = 0; k < jt->nkeys; ++k) {
    B2_REQUIRE(probe_keys[k] >= 0 && probe_keys[k] < scan->ncols, "probe key out of range");
    B2_REQUIRE(scan->cols[probe_keys[k]].dtype == jt->keys[k].dtype, "probe/build key types differ");
    pk->cols[k] = probe_keys[k];
  }
  if (jt->dense) B2_REQUIRE(jt->nkeys == 1 && jt->lookup && jt->range > 0, "bad dense table");
  else B2_REQUIRE(jt->head && b2_pow2(jt->cap), "bad chained table");
  return B2_OK;
}

int32_t b2_join_count(const b2_scan_t* scan, const int32_t* probe_keys, const b2_jointable_t* jt, int32_t mode,
                      int64_t* d_tile_off, void* stream) {
  b2_probekeys_arg pk;
  int32_t rc = b2_check_join(scan, probe_keys, jt, mode, &pk);
  if (rc) return rc;
  B2_REQUIRE(d_tile_off, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntiles = b2_num_tiles(scan->n);
  if (ntiles > 0) {
    int grid = b2_wave_grid(b2_join_count_kernel, B2_BLOCK, ntiles);
    b2_join_count_kernel<<<grid, B2_BLOCK, 0, st>>>(*scan, pk, *jt, mode, ntiles, d_tile_off);
    B2_CHECK_LAUNCH("b2_join_count_kernel");
  }
  b2_exclusive_scan_kernel<<<1, B2_SCAN_THREADS, 0, st>>>(d_tile_off, ntiles);
  B2_CHECK_LAUNCH("b2_exclusive_scan_kernel");
  return B2_OK;
}

int32_t b2_join_write(const b2_scan_t* scan, const int32_t* probe_keys, const b2_jointable_t* jt, int32_t mode,
                      const int64_t* d_tile_off, int32_t* out_probe_idx, int32_t* out_build_idx,
                      uint8_t* build_matched, void* stream) {
  b2_probekeys_arg pk;
  int32_t rc = b2_check_join(scan, probe_keys, jt, mode, &pk);
  if (rc) return rc;
  B2_REQUIRE(d_tile_off && out_probe_idx, "null argument");
  B2_REQUIRE(out_build_idx || mode == B2_JOIN_SEMI || mode == B2_JOIN_ANTI, "out_build_idx required");
  const int64_t ntiles = b2_num_tiles(scan->n);
  if (ntiles == 0) return B2_OK;
  int grid = b2_wave_grid(b2_join_write_kernel, B2_BLOCK, ntiles);
  b2_join_write_kernel<<<grid, B2_BLOCK, 0, (cudaStream_t)stream>>>(*scan, pk, *jt, mode, ntiles, d_tile_off,
                                                                     out_probe_idx, out_build_idx, build_matched);
  B2_CHECK_LAUNCH("b2_join_write_kernel");
  return B2_OK;
}

}  // extern "C"
