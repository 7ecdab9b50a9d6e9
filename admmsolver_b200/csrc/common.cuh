// Shared device/host helpers for libadmm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/admm_b200.h"

namespace admm {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define ADMM_REQUIRE(cond, code, ...)        \
  do {                                       \
    if (!(cond)) {                           \
      ::admm::set_error(__VA_ARGS__);        \
      return (code);                         \
    }                                        \
  } while (0)

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the 4 lanes of a quad (lanes differing in bits 0..1): the "t" index of an mma fragment
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// Block-wide sum of NV values per thread; result valid in every thread. scratch: >= NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + wid] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += scratch[i * 32 + w];  // fixed order: deterministic
    v[i] = s;
  }
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 on sm_100a.
// Fragment ownership (lane = 4*g + t):  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16-byte async global->shared copy (LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming (evict-first) 16-byte global load/store for the once-per-iteration state sweep
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream2(double* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};\n" ::"l"(p), "d"(v.x), "d"(v.y));
}

// XOR swizzle of the padded P operand (see DESIGN.md "P layout"): column offset for row r
__host__ __device__ __forceinline__ int p_swz(int r) {
  // sw = [0,2,1,3,2,0,3,1][r & 7]  (in units of 4 doubles)
  const int sw = ((r >> 1) & 1) | (((r & 1) ^ ((r >> 2) & 1)) << 1);
  return 4 * sw;
}

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

}  // namespace admm
