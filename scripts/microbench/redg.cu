// redg.cu — B200 micro-benchmarks behind the group-by design decisions:
//   * peak rate of fire-and-forget global atomics (REDG.E.ADD.F64 / .64) vs table size,
//     lanes-per-sector packing and same-address contention -- the ceiling C2 is reported against;
//   * MATCH.ANY.U64 and shared-memory ATOMS.ADD throughput (ranking / warp aggregation cost).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o redg redg.cu ; run on one B200.
// Prints one JSON object per line.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k;
}

// mode 0: one f64 RED per row into acc[slot]
// mode 1: two REDs per row: acc[slot] (f64) and cnt[slot] (u64), separate arrays
// mode 2: two REDs per row, AoS {sum, cnt} adjacent (two instructions, same sector)
// mode 3: paired lanes: ONE f64 RED instruction covers 16 rows x {sum, cnt-as-double} (lanes 2i, 2i+1 share a sector)
// mode 4: one u64 RED per row
template <int MODE>
__global__ void __launch_bounds__(256) red_kernel(double* acc, unsigned long long* cnt, uint64_t nslots, uint64_t nrows,
                                                  uint64_t hot, int rounds) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  for (uint64_t i = tid; i < nrows; i += nthreads) {
    uint64_t h = mix64(i * 0x9e3779b97f4a7c15ULL + 1);
    uint64_t slot = hot ? (h % hot) * (nslots / hot) : h % nslots;
    double v = (double)(h & 1023) * 0.5;
    if (MODE == 0) atomicAdd(acc + slot, v);
    else if (MODE == 1) { atomicAdd(acc + slot, v); atomicAdd(cnt + slot, 1ULL); }
    else if (MODE == 2) { atomicAdd(acc + 2 * slot, v); atomicAdd(reinterpret_cast<unsigned long long*>(acc) + 2 * slot + 1, 1ULL); }
    else if (MODE == 3) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int src = half * 16 + (lane >> 1);
        const uint64_t s2 = __shfl_sync(0xffffffffu, slot, src);
        const double v2 = __shfl_sync(0xffffffffu, v, src);
        atomicAdd(acc + 2 * s2 + (lane & 1), (lane & 1) ? 1.0 : v2);
      }
    } else if (MODE == 4) atomicAdd(cnt + slot, (unsigned long long)(h & 1023));
  }
}

// C2-like: streaming loads of key + val (16 B/row) then one RED per row
__global__ void __launch_bounds__(256) red_stream_kernel(const int64_t* __restrict__ key, const double* __restrict__ val,
                                                         double* acc, uint64_t nrows) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i + 7 * nthreads < nrows; i += 8 * nthreads) {
    int64_t k[8]; double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { k[j] = __ldcs(key + i + j * nthreads); v[j] = __ldcs(val + i + j * nthreads); }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(acc + k[j], v[j]);
  }
}
__global__ void fill_keys(int64_t* key, double* val, uint64_t n, uint64_t nslots) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t h = mix64(i + 12345);
    key[i] = (int64_t)(h % nslots);
    val[i] = (double)(h & 4095) / 4096.0;
  }
}

__global__ void __launch_bounds__(256) match_kernel(const int64_t* __restrict__ in, int* out, int iters, int spread) {
  int64_t v = in[threadIdx.x & 31] + (spread ? threadIdx.x : 0);
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    acc += __match_any_sync(0xffffffffu, v);
    v = v * 3 + (int64_t)acc;          // dependent: measures latency-bound chain per warp, many warps hide it
    if (!spread) v &= 7;
  }
  if (acc == 0x12345) out[0] = acc;
}

__global__ void __launch_bounds__(256) atoms_kernel(int* out, int iters, int nb) {
  __shared__ int hist[8][1024];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 1024; i += 256) (&hist[0][0])[i] = 0;
  __syncthreads();
  uint64_t h = mix64(threadIdx.x + blockIdx.x * 977);
  int acc = 0;
  for (int i = 0; i < iters; ++i) {
    h = h * 6364136223846793005ULL + 1442695040888963407ULL;
    acc += atomicAdd(&hist[warp][(h >> 33) % nb], 1);
  }
  if (acc == -1) out[0] = acc;
}

static float time_it(void (*launch)(void*), void* arg, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(arg);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch(arg);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

struct RedArgs { int mode; double* acc; unsigned long long* cnt; uint64_t nslots, nrows, hot; int grid; };
static void launch_red(void* p) {
  RedArgs* a = (RedArgs*)p;
  switch (a->mode) {
    case 0: red_kernel<0><<<a->grid, 256>>>(a->acc, a->cnt, a->nslots, a->nrows, a->hot, 1); break;
    case 1: red_kernel<1><<<a->grid, 256>>>(a->acc, a->cnt, a->nslots, a->nrows, a->hot, 1); break;
    case 2: red_kernel<2><<<a->grid, 256>>>(a->acc, a->cnt, a->nslots, a->nrows, a->hot, 1); break;
    case 3: red_kernel<3><<<a->grid, 256>>>(a->acc, a->cnt, a->nslots, a->nrows, a->hot, 1); break;
    default: red_kernel<4><<<a->grid, 256>>>(a->acc, a->cnt, a->nslots, a->nrows, a->hot, 1); break;
  }
}
struct StreamArgs { const int64_t* key; const double* val; double* acc; uint64_t nrows; int grid; };
static void launch_stream(void* p) { StreamArgs* a = (StreamArgs*)p; red_stream_kernel<<<a->grid, 256>>>(a->key, a->val, a->acc, a->nrows); }
struct MatchArgs { const int64_t* in; int* out; int iters, spread, grid; };
static void launch_match(void* p) { MatchArgs* a = (MatchArgs*)p; match_kernel<<<a->grid, 256>>>(a->in, a->out, a->iters, a->spread); }
struct AtomsArgs { int* out; int iters, nb, grid; };
static void launch_atoms(void* p) { AtomsArgs* a = (AtomsArgs*)p; atoms_kernel<<<a->grid, 256>>>(a->out, a->iters, a->nb); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d}\n", prop.name, sms, prop.l2CacheSize);
  const uint64_t nrows = 200000000ULL;
  const uint64_t max_slots = 64ULL << 20;   // 512 MB of doubles, x2 for AoS
  double* acc; unsigned long long* cnt;
  CK(cudaMalloc(&acc, max_slots * 16));
  CK(cudaMalloc(&cnt, max_slots * 8));
  CK(cudaMemset(acc, 0, max_slots * 16));
  CK(cudaMemset(cnt, 0, max_slots * 8));
  const char* names[] = {"f64_1red", "f64+u64_2red_soa", "f64+u64_2red_aos", "paired_lanes_aos_1instr_per_16rows", "u64_1red"};
  const uint64_t sizes[] = {1ULL << 20, 4ULL << 20, 16ULL << 20, 64ULL << 20};
  for (int occ : {8}) {
    for (int mode = 0; mode < 5; ++mode) {
      for (uint64_t ns : sizes) {
        RedArgs a{mode, acc, cnt, ns, nrows, 0, sms * occ};
        float ms = time_it(launch_red, &a);
        printf("{\"test\": \"red_spread\", \"mode\": \"%s\", \"table_mb_per_array\": %llu, \"rows\": %llu, \"ms\": %.4f, \"g_rows_per_s\": %.2f}\n",
               names[mode], (unsigned long long)(ns * 8 >> 20), (unsigned long long)nrows, ms, nrows / ms / 1e6);
        fflush(stdout);
      }
    }
  }
  // same-address contention: all rows spread over `hot` addresses
  for (uint64_t hot : {1ULL, 4ULL, 16ULL, 64ULL, 256ULL, 4096ULL}) {
    RedArgs a{0, acc, cnt, 1ULL << 20, nrows / 4, hot, sms * 8};
    float ms = time_it(launch_red, &a);
    printf("{\"test\": \"red_hot\", \"hot_addresses\": %llu, \"rows\": %llu, \"ms\": %.4f, \"g_rows_per_s\": %.3f, \"ns_per_atomic_per_address\": %.3f}\n",
           (unsigned long long)hot, (unsigned long long)(nrows / 4), ms, nrows / 4 / ms / 1e6, ms * 1e6 / (nrows / 4.0 / hot));
    fflush(stdout);
  }
  // C2-like: 16 B/row streamed + one RED per row, 1M and 16M slots
  {
    int64_t* key; double* val;
    CK(cudaMalloc(&key, nrows * 8)); CK(cudaMalloc(&val, nrows * 8));
    for (uint64_t ns : {1ULL << 20, 16ULL << 20}) {
      fill_keys<<<sms * 8, 256>>>(key, val, nrows, ns);
      CK(cudaDeviceSynchronize());
      for (int occ : {4, 8}) {
        StreamArgs a{key, val, acc, nrows, sms * occ};
        float ms = time_it(launch_stream, &a);
        printf("{\"test\": \"red_stream_c2\", \"slots\": %llu, \"ctas_per_sm\": %d, \"rows\": %llu, \"ms\": %.4f, \"g_rows_per_s\": %.2f, \"gbs_16B_per_row\": %.1f}\n",
               (unsigned long long)ns, occ, (unsigned long long)nrows, ms, nrows / ms / 1e6, nrows * 16 / ms / 1e6);
        fflush(stdout);
      }
    }
    cudaFree(key); cudaFree(val);
  }
  // MATCH.ANY.U64 and shared ATOMS.ADD
  {
    int64_t* in; int* out;
    CK(cudaMalloc(&in, 256)); CK(cudaMemset(in, 0, 256)); CK(cudaMalloc(&out, 64));
    for (int spread : {0, 1}) {
      MatchArgs a{in, out, 4096, spread, sms * 8};
      float ms = time_it(launch_match, &a);
      double warp_instr = (double)sms * 8 * 8 * 4096;
      printf("{\"test\": \"match_any_u64\", \"distinct_values\": %s, \"ms\": %.4f, \"warp_instr_per_s_G\": %.2f, \"cycles_per_warp_instr_per_sm_at_1.9GHz\": %.2f}\n",
             spread ? "\"32 per warp\"" : "\"<=8 per warp\"", ms, warp_instr / ms / 1e6, 1.9e9 * (ms * 1e-3) / (8 * 8 * 4096.0));
    }
    for (int nb : {1024, 128, 32, 1}) {
      AtomsArgs a{out, 4096, nb, sms * 8};
      float ms = time_it(launch_atoms, &a);
      double lane_ops = (double)sms * 8 * 256 * 4096;
      printf("{\"test\": \"smem_atoms_add_u32\", \"buckets\": %d, \"ms\": %.4f, \"g_lane_ops_per_s\": %.2f, \"lane_ops_per_clk_per_sm_at_1.9GHz\": %.2f}\n",
             nb, ms, lane_ops / ms / 1e6, lane_ops / (ms * 1e-3) / 1.9e9 / sms);
    }
  }
  return 0;
}
