// expr.cuh — per-row postfix expression interpreter (b2_expr_eval) and column statistics.
// One pass per expression TREE (the reference does one pandas pass per operator NODE,
// physical/rex/core/call.py:1158-1216).  The program is warp-uniform (kernel parameter
// space), so the interpreter loop never diverges; the value stack lives in local memory.
#pragma once
#include "common.cuh"

#define B2_STACK 16

struct b2_cols_arg {
  b2_col_t c[B2_MAX_COLS];
};

__global__ void __launch_bounds__(B2_BLOCK)
b2_expr_kernel(const __grid_constant__ b2_prog_t prog, const __grid_constant__ b2_cols_arg cols,
               int64_t n, void* __restrict__ out_data, uint32_t* __restrict__ out_valid) {
  const int64_t n32 = (n + 31) & ~(int64_t)31;
  for (int64_t row = (int64_t)blockIdx.x * B2_BLOCK + threadIdx.x; row < n32;
       row += (int64_t)gridDim.x * B2_BLOCK) {
    const bool active = row < n;
    int64_t sv[B2_STACK];
    bool sn[B2_STACK];
    int sp = 0;
    if (active) {
      for (int pc = 0; pc < prog.n; ++pc) {
        const b2_instr_t ins = prog.code[pc];
        const int op = ins.op;
        if (op == B2_OP_LOAD) {
          const b2_col_t& c = cols.c[ins.a];
          int64_t raw = c.dtype == B2_U8 ? (int64_t) reinterpret_cast<const uint8_t*>(c.data)[row]
                                         : reinterpret_cast<const int64_t*>(c.data)[row];
          sv[sp] = raw;
          sn[sp] = c.valid ? !b2_bit(c.valid, row) : false;
          ++sp;
        } else if (op == B2_OP_CONST_I) {
          sv[sp] = ins.imm_i; sn[sp] = false; ++sp;
        } else if (op == B2_OP_CONST_F) {
          sv[sp] = __double_as_longlong(ins.imm_f); sn[sp] = false; ++sp;
        } else if (op == B2_OP_CONST_NULL) {
          sv[sp] = 0; sn[sp] = true; ++sp;
        } else if (op == B2_OP_I2F) {
          sv[sp - 1] = __double_as_longlong((double)sv[sp - 1]);
        } else if (op == B2_OP_F2I) {
          double d = __longlong_as_double(sv[sp - 1]);
          if (d != d) { sn[sp - 1] = true; sv[sp - 1] = 0; }
          else sv[sp - 1] = (int64_t)d;
        } else if (op == B2_OP_NEG_I) {
          sv[sp - 1] = (int64_t)(0ULL - (uint64_t)sv[sp - 1]);
        } else if (op == B2_OP_ABS_I) {
          int64_t v = sv[sp - 1]; sv[sp - 1] = v < 0 ? (int64_t)(0ULL - (uint64_t)v) : v;
        } else if (op == B2_OP_NEG_F) {
          sv[sp - 1] = __double_as_longlong(-__longlong_as_double(sv[sp - 1]));
        } else if (op == B2_OP_SQRT_F) {
          sv[sp - 1] = __double_as_longlong(sqrt(__longlong_as_double(sv[sp - 1])));
        } else if (op == B2_OP_ABS_F) {
          sv[sp - 1] = __double_as_longlong(fabs(__longlong_as_double(sv[sp - 1])));
        } else if (op == B2_OP_ORD2F) {
          sv[sp - 1] = b2_ordered_from_bits(sv[sp - 1]);
        } else if (op == B2_OP_NOT) {
          sv[sp - 1] = sv[sp - 1] == 0;
        } else if (op == B2_OP_ISNULL_I) {
          sv[sp - 1] = sn[sp - 1]; sn[sp - 1] = false;
        } else if (op == B2_OP_ISNULL_F) {
          double d = __longlong_as_double(sv[sp - 1]);
          sv[sp - 1] = sn[sp - 1] || d != d; sn[sp - 1] = false;
        } else if (op == B2_OP_CASE) {
          // stack: cond, then, else
          const int64_t ev = sv[sp - 1]; const bool en = sn[sp - 1];
          const int64_t tv = sv[sp - 2]; const bool tn = sn[sp - 2];
          const bool take = !sn[sp - 3] && sv[sp - 3] != 0;
          sp -= 2;
          sv[sp - 1] = take ? tv : ev;
          sn[sp - 1] = take ? tn : en;
        } else if (op == B2_OP_FILLNA) {
          // stack: x, fill
          if (sn[sp - 2]) { sv[sp - 2] = sv[sp - 1]; sn[sp - 2] = sn[sp - 1]; }
          --sp;
        } else {
          // binary operators: pops b then a
          const int64_t b = sv[sp - 1]; const bool bn = sn[sp - 1];
          const int64_t a = sv[sp - 2]; const bool an = sn[sp - 2];
          --sp;
          int64_t r = 0; bool rn = an || bn;
          if (op == B2_OP_AND) {  // Kleene: False wins over NULL
            const bool af = !an && a == 0, bf = !bn && b == 0;
            rn = (an || bn) && !(af || bf);
            r = (!an && a != 0) && (!bn && b != 0);
          } else if (op == B2_OP_OR) {  // Kleene: True wins over NULL
            const bool at = !an && a != 0, bt = !bn && b != 0;
            rn = (an || bn) && !(at || bt);
            r = at || bt;
          } else if (op >= B2_OP_EQ_F && op <= B2_OP_EQ_F + 5) {
            r = b2_cmp_f(op - B2_OP_EQ_F, __longlong_as_double(a), __longlong_as_double(b));
          } else if (op >= B2_OP_EQ_I && op <= B2_OP_EQ_I + 5) {
            r = b2_cmp_i(op - B2_OP_EQ_I, a, b);
          } else if (op == B2_OP_ADD_I) r = (int64_t)((uint64_t)a + (uint64_t)b);
          else if (op == B2_OP_SUB_I) r = (int64_t)((uint64_t)a - (uint64_t)b);
          else if (op == B2_OP_MUL_I) r = (int64_t)((uint64_t)a * (uint64_t)b);
          else if (op == B2_OP_DIV_I) {
            if (b == 0) rn = true;
            else if (b == -1) r = (int64_t)(0ULL - (uint64_t)a);
            else r = a / b;  // C++ '/' truncates toward zero = SQL semantics
          } else if (op == B2_OP_MOD_I) {
            if (b == 0) rn = true;
            else if (b == -1) r = 0;
            else r = a % b;
          } else if (op == B2_OP_ADD_F) r = __double_as_longlong(__longlong_as_double(a) + __longlong_as_double(b));
          else if (op == B2_OP_SUB_F) r = __double_as_longlong(__longlong_as_double(a) - __longlong_as_double(b));
          else if (op == B2_OP_MUL_F) r = __double_as_longlong(__longlong_as_double(a) * __longlong_as_double(b));
          else if (op == B2_OP_DIV_F) r = __double_as_longlong(__longlong_as_double(a) / __longlong_as_double(b));
          sv[sp - 1] = r;
          sn[sp - 1] = rn;
        }
      }
    }
    const bool isnull = active ? sn[0] : true;
    const int64_t v = (active && !isnull) ? sv[0] : 0;
    if (active) {
      if (prog.out_dtype == B2_U8) reinterpret_cast<uint8_t*>(out_data)[row] = (uint8_t)(v != 0);
      else reinterpret_cast<int64_t*>(out_data)[row] = v;
    }
    if (out_valid) {
      const uint32_t w = __ballot_sync(FULL_MASK, active && !isnull);
      if ((threadIdx.x & 31) == 0) out_valid[row >> 5] = w;
    }
  }
}

// ---- column statistics ------------------------------------------------------------------
// out: {min(ordered image), max(ordered image), nulls, nans, repeating rows of the sample, rows sampled}
__global__ void b2_stats_init_kernel(int64_t* out) {
  out[0] = LLONG_MAX; out[1] = LLONG_MIN; out[2] = 0; out[3] = 0; out[4] = 0; out[5] = 0;
}
__global__ void __launch_bounds__(B2_BLOCK)
b2_stats_kernel(const __grid_constant__ b2_col_t col, int64_t n, int64_t* __restrict__ out) {
  long long mn = LLONG_MAX, mx = LLONG_MIN;
  unsigned long long nulls = 0, nans = 0;
  {
    // repeat sample: of the 32 consecutive rows each warp meets first (spread over the whole column by
    // the grid stride), how many share their value with another one?  Decides whether a GROUP BY on
    // this column pre-aggregates per warp (out[4] = such rows, out[5] = rows sampled).
    const int64_t row = (int64_t)blockIdx.x * B2_BLOCK + threadIdx.x;
    bool ok = row < n && !(col.valid && !b2_bit(col.valid, row));
    int64_t raw = ok ? b2_load_raw(col, row) : 0;
    if (ok && col.dtype == B2_F64) { const double d = __longlong_as_double(raw); ok = d == d; }
    const uint32_t act = __ballot_sync(FULL_MASK, ok);
    uint32_t m = 0;
    if (ok) m = __match_any_sync(act, raw);
    const uint32_t rep = __ballot_sync(FULL_MASK, ok && __popc(m) > 1);
    if ((threadIdx.x & 31) == 0 && act) {
      atomicAdd(reinterpret_cast<unsigned long long*>(out + 4), (unsigned long long)__popc(rep));
      atomicAdd(reinterpret_cast<unsigned long long*>(out + 5), (unsigned long long)__popc(act));
    }
  }
  for (int64_t row = (int64_t)blockIdx.x * B2_BLOCK + threadIdx.x; row < n;
       row += (int64_t)gridDim.x * B2_BLOCK) {
    if (col.valid && !b2_bit(col.valid, row)) { ++nulls; continue; }
    int64_t raw = b2_load_raw(col, row);
    if (col.dtype == B2_F64) {
      double d = __longlong_as_double(raw);
      if (d != d) { ++nans; continue; }
      raw = b2_ordered_from_bits(raw);
    }
    mn = raw < mn ? raw : mn;
    mx = raw > mx ? raw : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long omn = __shfl_xor_sync(FULL_MASK, mn, o), omx = __shfl_xor_sync(FULL_MASK, mx, o);
    mn = omn < mn ? omn : mn;
    mx = omx > mx ? omx : mx;
    nulls += __shfl_xor_sync(FULL_MASK, nulls, o);
    nans += __shfl_xor_sync(FULL_MASK, nans, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mn != LLONG_MAX) atomicMin(reinterpret_cast<long long*>(out), mn);
    if (mx != LLONG_MIN) atomicMax(reinterpret_cast<long long*>(out + 1), mx);
    if (nulls) atomicAdd(reinterpret_cast<unsigned long long*>(out + 2), nulls);
    if (nans) atomicAdd(reinterpret_cast<unsigned long long*>(out + 3), nans);
  }
}
__global__ void b2_stats_fini_kernel(int64_t* out, int is_f64) {
  if (is_f64) {
    if (out[0] != LLONG_MAX || out[1] != LLONG_MIN) {
      out[0] = b2_ordered_from_bits(out[0]);
      out[1] = b2_ordered_from_bits(out[1]);
    }
  }
}

extern "C" {

int64_t b2_stats_ws_bytes(void) { return 256; }

int32_t b2_col_stats(const b2_col_t* col, int64_t n, int64_t* d_out, void* ws, void* stream) {
  (void)ws;
  B2_REQUIRE(col && d_out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  b2_stats_init_kernel<<<1, 1, 0, st>>>(d_out);
  if (n > 0) {
    int grid = b2_wave_grid(b2_stats_kernel, B2_BLOCK, (n + B2_BLOCK - 1) / B2_BLOCK);
    b2_stats_kernel<<<grid, B2_BLOCK, 0, st>>>(*col, n, d_out);
  }
  b2_stats_fini_kernel<<<1, 1, 0, st>>>(d_out, col->dtype == B2_F64);
  B2_CHECK_LAUNCH("b2_stats_kernel");
  return B2_OK;
}

int32_t b2_expr_eval(const b2_prog_t* prog, const b2_col_t* cols, int32_t ncols, int64_t n,
                     void* out_data, uint32_t* out_valid, void* stream) {
  B2_REQUIRE(prog && out_data, "null argument");
  B2_REQUIRE(ncols >= 0 && ncols <= B2_MAX_COLS, "too many columns");
  B2_REQUIRE(prog->n > 0 && prog->n <= B2_MAX_PROG, "bad program length");
  // static stack-depth check so the kernel cannot run off its local stack
  int sp = 0;
  for (int i = 0; i < prog->n; ++i) {
    int op = prog->code[i].op;
    if (op == B2_OP_LOAD) {
      B2_REQUIRE(prog->code[i].a >= 0 && prog->code[i].a < ncols, "LOAD of unknown column");
      ++sp;
    } else if (op == B2_OP_CONST_I || op == B2_OP_CONST_F || op == B2_OP_CONST_NULL) ++sp;
    else if (op == B2_OP_I2F || op == B2_OP_F2I || op == B2_OP_NEG_I || op == B2_OP_ABS_I ||
             op == B2_OP_NEG_F || op == B2_OP_ABS_F || op == B2_OP_SQRT_F || op == B2_OP_NOT || op == B2_OP_ISNULL_I ||
             op == B2_OP_ISNULL_F || op == B2_OP_ORD2F) { B2_REQUIRE(sp >= 1, "stack underflow"); }
    else if (op == B2_OP_CASE) { B2_REQUIRE(sp >= 3, "stack underflow"); sp -= 2; }
    else { B2_REQUIRE(sp >= 2, "stack underflow"); sp -= 1; }
    B2_REQUIRE(sp <= B2_STACK, "expression too deep");
  }
  B2_REQUIRE(sp == 1, "program must leave exactly one value");
  if (n <= 0) return B2_OK;
  b2_cols_arg ca;
  memset(&ca, 0, sizeof(ca));
  for (int i = 0; i < ncols; ++i) ca.c[i] = cols[i];
  int grid = b2_wave_grid(b2_expr_kernel, B2_BLOCK, (n + B2_BLOCK - 1) / B2_BLOCK);
  b2_expr_kernel<<<grid, B2_BLOCK, 0, (cudaStream_t)stream>>>(*prog, ca, n, out_data, out_valid);
  B2_CHECK_LAUNCH("b2_expr_kernel");
  return B2_OK;
}

}  // extern "C"
