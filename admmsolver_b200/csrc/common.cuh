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

// shared-memory address of a generic pointer / 16-byte load that stays in program order among the volatile MMAs
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_v2_volatile(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 on sm_100a.
// Fragment ownership (lane = 4*g + t):  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  // volatile: keeps the MMAs in program order relative to the (volatile) streaming loads/stores, so
  // a software prefetch issued before a GEMM really is issued before it
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

// ---- mbarrier + TMA bulk copy (global -> shared, SASS UBLKCP) ---------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bulk prefetch of `bytes` (multiple of 16) of global memory into L2 (SASS UBLKPF): no register or
// shared-memory cost, the later LDG hits L2 instead of HBM
__device__ __forceinline__ void l2_prefetch_bulk(const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}
// contiguous `bytes` (multiple of 16, both addresses 16-byte aligned) global -> shared; completion
// is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- distributed shared memory push: remote store that signals the DESTINATION CTA's mbarrier ----
__device__ __forceinline__ unsigned mapa_u32(unsigned smem_addr, unsigned cta_rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
// 8-byte store into a peer CTA's shared memory; completes 8 tx-bytes on the peer's mbarrier (SASS: ST.ASYNC... /
// UBLKCP-free DSMEM path): the receiver only waits on its own barrier -- no cluster barrier, no fence
__device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- programmatic dependent launch (PDL): the next kernel of the stream may be scheduled while this one
// still runs; it blocks in pdl_wait() until this grid has completed and its writes are visible.  Both are
// no-ops for kernels launched without the attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
}
// launch with programmatic stream serialisation (every kernel launched this way starts with pdl_prologue())
// coop: cooperative launch -- the driver guarantees that all CTAs of the grid are resident at the same time (or fails
// the launch up front), which is what the kernels that wait for each other inside one launch rely on.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                               Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

// current device: kernel attributes (opt-in shared memory, cluster sizes) are per device, so every "already
// configured" cache of a launcher is keyed by it
inline int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

}  // namespace admm
