// Structured-matrix primitives of the generic executor (FP64 / complex128).
// These are the device replacements of the NumPy calls in the reference's matrix.py and of the
// vector algebra in optimizer.py; the two fused engines (spm.cu, bp.cu) carry the hot paths.
#include "common.cuh"

#include <cstdlib>

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <map>

namespace cg = cooperative_groups;

namespace admm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return ADMM_OK;
}

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
struct cplx {
  double x, y;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.x, -a.y}; }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm (what NumPy/LAPACK use) to avoid overflow
  if (fabs(b.x) >= fabs(b.y)) {
    double r = b.y / b.x, den = b.x + b.y * r;
    return {(a.x + a.y * r) / den, (a.y - a.x * r) / den};
  } else {
    double r = b.x / b.y, den = b.y + b.x * r;
    return {(a.x * r + a.y) / den, (a.y * r - a.x) / den};
  }
}

template <typename T> struct num;
template <> struct num<double> {
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double one() { return 1.0; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double conj(double a) { return a; }
  static __device__ __forceinline__ double abs2(double a) { return a * a; }
};
template <> struct num<cplx> {
  static __device__ __forceinline__ cplx zero() { return {0.0, 0.0}; }
  static __device__ __forceinline__ cplx one() { return {1.0, 0.0}; }
  static __device__ __forceinline__ cplx mul(cplx a, cplx b) { return cmul(a, b); }
  static __device__ __forceinline__ cplx add(cplx a, cplx b) { return cadd(a, b); }
  static __device__ __forceinline__ cplx sub(cplx a, cplx b) { return csub(a, b); }
  static __device__ __forceinline__ cplx div(cplx a, cplx b) { return cdiv(a, b); }
  static __device__ __forceinline__ cplx conj(cplx a) { return cconj(a); }
  static __device__ __forceinline__ double abs2(cplx a) { return cabs2(a); }
};

// ---------------------------------------------------------------------------------------------
// GEMM: C = op(A) B, 32x32 output tile, 32-deep K slabs in shared memory, 256 threads
// ---------------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(256) gemm_kernel(int m, int n, int k, const T* __restrict__ A, int lda,
                                                   const T* __restrict__ B, int ldb, T* __restrict__ C, int ldc) {
  __shared__ T As[32][33];
  __shared__ T Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in 0..7
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  T acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = num<T>::zero();
  for (int k0 = 0; k0 < k; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      // As[r][tx] = op(A)[row0 + r][k0 + tx]
      T a = num<T>::zero();
      if (OP == ADMM_OP_N) {
        if (row0 + r < m && k0 + tx < k) a = A[(size_t)(row0 + r) * lda + k0 + tx];
        As[r][tx] = a;
      } else {
        // stored k x m: element op(A)[i][kk] = A[kk][i]; read coalesced along i
        if (k0 + r < k && row0 + tx < m) a = A[(size_t)(k0 + r) * lda + row0 + tx];
        if (OP == ADMM_OP_H) a = num<T>::conj(a);
        As[tx][r] = a;
      }
      T b = num<T>::zero();
      if (k0 + r < k && col0 + tx < n) b = B[(size_t)(k0 + r) * ldb + col0 + tx];
      Bs[r][tx] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const T b = Bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = num<T>::add(acc[i], num<T>::mul(As[ty + 8 * i][kk], b));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty + 8 * i, c = col0 + tx;
    if (r < m && c < n) C[(size_t)r * ldc + c] = acc[i];
  }
}

// ---------------------------------------------------------------------------------------------
// real FP64 GEMM on the tensor cores (DMMA.8x8x4):  C[m x n] = op(A)[m x k] * B[k x n]
//   CTA tile 64 x 64, k-step 16, 4 warps in a 2 x 2 arrangement, each 32 x 32 = 4 x 4 mma tiles
//   (32 accumulator doubles per lane).  Global -> registers -> shared memory with the next k-step
//   prefetched into registers while the current one is multiplied; shared pitches 20 / 68 doubles
//   (= 4 mod 16) make both fragment loads conflict free.  Edges are zero-filled.
// This is the x-update GEMM of a generic LeastSquares term whose right-hand sides share one A
// (PartialDiagonalMatrix packing): B = (alpha A^H A + mu)^-1 times an (N x nbatch) block.
// ---------------------------------------------------------------------------------------------
//
// CPLX: a complex128 product on the same tensor-core path through its real form -- no operand is copied or split.
// With interleaved storage a complex row IS the real row (Re, Im, Re, Im, ...), so for C = op(A) B
//     C_view (m x 2n) = A' (m x 2k) * B' (2k x 2n),   A'[i][2l+e] = part_e(op(A)[i][l]),
//     B'[2l][c] = B_view[l][c],   B'[2l+1][c] = (J B_view)[l][c]  with  J: (re, im) -> (-im, re)
// which are exactly the four real products Ar Br - Ai Bi, Ar Bi + Ai Br (8 m n k flops, nothing wasted).  The kernel
// is called with the REAL dimensions (m, 2n, 2k) and leading dimensions in doubles; only the loaders differ.
template <int OP, bool CPLX>
__global__ void __launch_bounds__(128) gemm_dmma_kernel(int m, int n, int k, const double* __restrict__ A, int lda,
                                                        const double* __restrict__ B, int ldb, double* __restrict__ C,
                                                        int ldc) {
  constexpr int BM = 64, BN = 64, BK = 16, PA = BK + 4, PB = BN + 4;
  __shared__ double As[BM * PA];      // As[i][kk] = op(A)[row0 + i][k0 + kk]
  __shared__ double Bs[BK * PB];      // Bs[kk][j] = B[k0 + kk][col0 + j]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[8], rb[8];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + 128 * q;            // 0 .. 1023
      double a = 0.0;
      if (OP == ADMM_OP_N) {
        const int i = idx >> 4, kk = idx & 15;  // coalesced along k
        if (row0 + i < m && k0 + kk < k) a = A[(size_t)(row0 + i) * lda + k0 + kk];
      } else if (!CPLX) {
        const int kk = idx >> 6, i = idx & 63;  // A stored k x m: coalesced along m
        if (row0 + i < m && k0 + kk < k) a = A[(size_t)(k0 + kk) * lda + row0 + i];
      } else {
        const int kk = idx >> 6, i = idx & 63;  // A stored (k/2) x m complex: part e of element (l, i)
        if (row0 + i < m && k0 + kk < k) {
          const int l = (k0 + kk) >> 1, e = (k0 + kk) & 1;
          a = A[(size_t)l * lda + 2 * (row0 + i) + e];
          if (OP == ADMM_OP_H && e) a = -a;
        }
      }
      ra[q] = a;
      const int kb = idx >> 6, j = idx & 63;    // coalesced along n
      if (!CPLX) {
        rb[q] = (k0 + kb < k && col0 + j < n) ? B[(size_t)(k0 + kb) * ldb + col0 + j] : 0.0;
      } else {
        double bv = 0.0;
        if (k0 + kb < k && col0 + j < n) {
          const double* br = B + (size_t)((k0 + kb) >> 1) * ldb;
          const int c = col0 + j;
          if (((k0 + kb) & 1) == 0) bv = br[c];
          else bv = (c & 1) ? br[c - 1] : -br[c + 1];
        }
        rb[q] = bv;
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = tid + 128 * q;
      if (OP == ADMM_OP_N) As[(idx >> 4) * PA + (idx & 15)] = ra[q];
      else As[(idx & 63) * PA + (idx >> 6)] = ra[q];
      Bs[(idx >> 6) * PB + (idx & 63)] = rb[q];
    }
  };
  load_tiles(0);
  for (int k0 = 0; k0 < k; k0 += BK) {
    __syncthreads();                 // previous k-step consumed
    store_tiles();
    __syncthreads();
    if (k0 + BK < k) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wm + 8 * i + g) * PA + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[(kk + t) * PB + wn + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + wm + 8 * i + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + wn + 8 * j + 2 * t;
      if (r < m) {
        if (c < n) C[(size_t)r * ldc + c] = acc[i][j][0];
        if (c + 1 < n) C[(size_t)r * ldc + c + 1] = acc[i][j][1];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// skinny products (n <= 4 right-hand sides: the matrix-vector products of the generic executor; a real
// operator on a complex vector arrives as n = 2 interleaved columns).  Bandwidth-bound: A is read once,
// coalesced, by as many warps as there are rows (op N) or by 8 row-interleaved warps per 32 columns (op T/H).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum_t(T v);
template <>
__device__ __forceinline__ double warp_sum_t<double>(double v) { return warp_sum(v); }
template <>
__device__ __forceinline__ cplx warp_sum_t<cplx>(cplx v) { return {warp_sum(v.x), warp_sum(v.y)}; }

template <typename T, int NC>
__global__ void __launch_bounds__(256) gemv_n_kernel(int m, int n, int k, const T* __restrict__ A, int lda,
                                                     const T* __restrict__ B, int ldb, T* __restrict__ C, int ldc) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  T acc[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) acc[j] = num<T>::zero();
  const T* ar = A + (size_t)row * lda;
  for (int c = lane; c < k; c += 32) {
    const T a = ar[c];
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (j < n) acc[j] = num<T>::add(acc[j], num<T>::mul(a, B[(size_t)c * ldb + j]));
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const T v = warp_sum_t<T>(acc[j]);
    if (lane == 0 && j < n) C[(size_t)row * ldc + j] = v;
  }
}

template <typename T, int OP, int NC>
__global__ void __launch_bounds__(256) gemv_t_kernel(int m, int n, int k, const T* __restrict__ A, int lda,
                                                     const T* __restrict__ B, int ldb, T* __restrict__ C, int ldc) {
  __shared__ T part[8][NC][33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 32 + lane;        // output row i = column of the stored k x m matrix
  T acc[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) acc[j] = num<T>::zero();
  if (col < m) {
#pragma unroll 4
    for (int kk = w; kk < k; kk += 8) {
      T a = A[(size_t)kk * lda + col];
      if (OP == ADMM_OP_H) a = num<T>::conj(a);
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (j < n) acc[j] = num<T>::add(acc[j], num<T>::mul(a, B[(size_t)kk * ldb + j]));
    }
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) part[w][j][lane] = acc[j];
  __syncthreads();
  if (w == 0 && col < m) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      T v = part[0][j][lane];
#pragma unroll
      for (int q = 1; q < 8; ++q) v = num<T>::add(v, part[q][j][lane]);
      if (j < n) C[(size_t)col * ldc + j] = v;
    }
  }
}

template <typename T, int NC>
static int launch_gemv(int op, int m, int n, int k, const T* a, int lda, const T* b, int ldb, T* c, int ldc, cudaStream_t s) {
  if (op == ADMM_OP_N) gemv_n_kernel<T, NC><<<ceil_div(m, 8), 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  else if (op == ADMM_OP_T) gemv_t_kernel<T, ADMM_OP_T, NC><<<ceil_div(m, 32), 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  else gemv_t_kernel<T, ADMM_OP_H, NC><<<ceil_div(m, 32), 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  return check_launch("admm_gemm(skinny)");
}

template <typename T>
static int launch_gemm(int op, int m, int n, int k, const void* A, int lda, const void* B, int ldb, void* C,
                       int ldc, cudaStream_t s) {
  if (n <= 4 && !getenv("ADMM_GEMM_NO_SKINNY")) {
    const T* a = static_cast<const T*>(A);
    const T* b = static_cast<const T*>(B);
    T* c = static_cast<T*>(C);
    return n <= 2 ? launch_gemv<T, 2>(op, m, n, k, a, lda, b, ldb, c, ldc, s) : launch_gemv<T, 4>(op, m, n, k, a, lda, b, ldb, c, ldc, s);
  }
  dim3 grid(ceil_div(n, 32), ceil_div(m, 32));
  const T* a = static_cast<const T*>(A);
  const T* b = static_cast<const T*>(B);
  T* c = static_cast<T*>(C);
  if (op == ADMM_OP_N)
    gemm_kernel<T, ADMM_OP_N><<<grid, 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  else if (op == ADMM_OP_T)
    gemm_kernel<T, ADMM_OP_T><<<grid, 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  else
    gemm_kernel<T, ADMM_OP_H><<<grid, 256, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
  return check_launch("admm_gemm");
}

// ---------------------------------------------------------------------------------------------
// elementwise kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void diag_mul_kernel(int rows_out, int nd, int ncols, const T* __restrict__ d, const T* __restrict__ V,
                                int ldv, T* __restrict__ out, int ldo) {
  const long long total = (long long)rows_out * ncols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = int(idx / ncols), j = int(idx % ncols);
    T v = num<T>::zero();
    if (i < nd) v = num<T>::mul(d[i], V[(size_t)i * ldv + j]);
    out[(size_t)i * ldo + j] = v;
  }
}

__global__ void axpby_kernel(long long n, double a, const double* x, double b,
                             const double* y, double* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = a * x[i];
    if (y) v += b * y[i];
    out[i] = v;
  }
}

template <typename T, int OPC>
__global__ void ewise_unary_kernel(long long n, const T* x, T* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = OPC == 0 ? num<T>::div(num<T>::one(), x[i]) : num<T>::conj(x[i]);
  }
}

__global__ void prox_l1_kernel(long long n, const double* __restrict__ h, int hs, const double* __restrict__ mud,
                               double alpha, double* __restrict__ out, int os) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double mu = mud[i];
    const double y = -(h[i * hs] / mu);
    const double lam = 0.5 * alpha / mu;
    double r = 0.0;
    if (y > lam) r = y - lam;
    if (y < -lam) r = y + lam;
    out[i * os] = r;
    if (os == 2) out[i * 2 + 1] = 0.0;
  }
}

__global__ void prox_nonneg_kernel(long long n, const double* __restrict__ h, int hs, const double* __restrict__ mud,
                                   double* __restrict__ out, int os) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = -(h[i * hs] / mud[i]);
    if (v < 0) v = 0.0;
    out[i * os] = v;
    if (os == 2) out[i * 2 + 1] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// PSD-cone projection of many small real symmetric matrices (SemiPositiveDefinitePenalty.solve,
// objectivefunc.py:294-327): X = -Re(h)/mu elementwise, then per matrix (LOWER triangle, like
// np.linalg.eigh) X <- sum_{lambda_k >= 0} lambda_k u_k u_k^T.  One warp per matrix, cyclic Jacobi
// in shared memory (lane k owns row/column k of the rotations), n <= 32.
// Matrix m, element (p, q) sits at flat index m*sb + p*sr + q*sc of the term's vector.
// ---------------------------------------------------------------------------------------------
constexpr int PSD_WARPS = 2;

__global__ void __launch_bounds__(PSD_WARPS * 32) prox_psd_kernel(int n, long long nbatch, long long sb, long long sr,
                                                                  long long sc, const double* __restrict__ h, int hs,
                                                                  const double* __restrict__ mud, double* __restrict__ out,
                                                                  int os) {
  extern __shared__ double psd_sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pitch = n | 1;
  double* A = psd_sm + (size_t)warp * 2 * n * pitch;
  double* V = A + (size_t)n * pitch;
  for (long long m = (long long)blockIdx.x * PSD_WARPS + warp; m < nbatch; m += (long long)gridDim.x * PSD_WARPS) {
    // ---- load: symmetric matrix from the lower triangle of -Re(h)/mu
    for (int p = 0; p < n; ++p) {
      if (lane < n) {
        const int r = p > lane ? p : lane, c = p > lane ? lane : p;
        const long long off = m * sb + r * sr + c * sc;
        A[p * pitch + lane] = -(h[off * hs] / mud[off]);
        V[p * pitch + lane] = (p == lane) ? 1.0 : 0.0;
      }
    }
    __syncwarp();
    double fro = 0.0;
    if (lane < n)
      for (int q = 0; q < n; ++q) fro += A[lane * pitch + q] * A[lane * pitch + q];
    fro = warp_sum(fro);
    // ---- cyclic Jacobi sweeps
    for (int sweep = 0; sweep < 40; ++sweep) {
      double off2 = 0.0;
      if (lane < n)
        for (int q = 0; q < n; ++q)
          if (q != lane) off2 += A[lane * pitch + q] * A[lane * pitch + q];
      off2 = warp_sum(off2);
      if (!(off2 > 1e-31 * fro)) break;
      for (int p = 0; p < n - 1; ++p) {
        for (int q = p + 1; q < n; ++q) {
          const double apq = A[p * pitch + q];
          if (apq != 0.0) {
            const double app = A[p * pitch + p], aqq = A[q * pitch + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            __syncwarp();
            if (lane < n) {
              const int k = lane;
              const double vkp = V[k * pitch + p], vkq = V[k * pitch + q];
              V[k * pitch + p] = c * vkp - sn * vkq;
              V[k * pitch + q] = sn * vkp + c * vkq;
              if (k != p && k != q) {
                const double akp = A[k * pitch + p], akq = A[k * pitch + q];
                const double np_ = c * akp - sn * akq, nq_ = sn * akp + c * akq;
                A[k * pitch + p] = np_;
                A[p * pitch + k] = np_;
                A[k * pitch + q] = nq_;
                A[q * pitch + k] = nq_;
              }
            }
            __syncwarp();
            if (lane == 0) {
              A[p * pitch + p] = app - t * apq;
              A[q * pitch + q] = aqq + t * apq;
              A[p * pitch + q] = 0.0;
              A[q * pitch + p] = 0.0;
            }
            __syncwarp();
          }
        }
      }
    }
    // ---- X+ = V diag(max(lambda, 0)) V^T : lane i writes row i
    if (lane < n) {
      const int i = lane;
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
          const double lam = A[k * pitch + k];
          if (lam > 0.0) acc += lam * V[i * pitch + k] * V[j * pitch + k];
        }
        const long long off = m * sb + i * sr + j * sc;
        out[off * os] = acc;
        if (os == 2) out[off * 2 + 1] = 0.0;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on the columns of W (m x n, COLUMN-major, leading dimension ld): plane rotations of
// column pairs until all columns are mutually orthogonal, W <- W R, optionally V <- V R (n x n, column-major).
// Then the column norms are the singular values of the input, the normalised columns its left singular vectors and
// (V started as I) V its right singular vectors -- each to high RELATIVE accuracy, which the kernel of the IR basis
// needs (its singular values span 16 decades).  For a symmetric positive definite input the columns end up as
// lambda_k v_k.  Parallel ordering: round r of the round-robin tournament on ne = n rounded up to even positions pairs
// ne/2 disjoint columns; one warp rotates one pair (three dot products, one fused update), all pairs of a round are
// independent.  Used by the PSD projection of matrices larger than a warp (one CTA per matrix, rounds separated by
// __syncthreads) and by the SVD of the IR kernel (one matrix spread over a cooperative grid, rounds separated by
// grid.sync()).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rr_pair(int ne, int r, int i, int& p, int& q) {
  const int md = ne - 1;
  if (i == 0) {
    p = ne - 1;
    q = r;
  } else {
    p = (r + i) % md;
    q = (r - i + md) % md;
  }
}

// returns |cos| of the angle between the two columns before the rotation (0 when one of them vanishes).
// zero2: columns whose squared norm is at most zero2 are rounding noise of a rank-deficient input (the directions of
// vanishing singular values); they are not rotated -- their mutual angles are noise and never settle.
// (Exchanging the columns so that the larger norm comes first -- de Rijk -- was measured to slow the parallel ordering
// down: positions pair up once per sweep, so moving columns breaks the guarantee that a sweep looks at every pair.)
__device__ __forceinline__ double jacobi_pair(double* __restrict__ a, double* __restrict__ b, int m, double* va, double* vb,
                                              int nv, int lane, double tol, double zero2 = 0.0) {
  double al = 0.0, be = 0.0, ga = 0.0;
  for (int i = lane; i < m; i += 32) {
    const double x = a[i], y = b[i];
    al += x * x;
    be += y * y;
    ga += x * y;
  }
  al = warp_sum(al);
  be = warp_sum(be);
  ga = warp_sum(ga);
  double off = 0.0;
  if (al > zero2 && be > zero2) off = fabs(ga) / sqrt(al * be);
  if (off > tol) {
    const double zeta = (be - al) / (2.0 * ga);
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
    for (int i = lane; i < m; i += 32) {
      const double x = a[i], y = b[i];
      a[i] = c * x - sn * y;
      b[i] = sn * x + c * y;
    }
    if (va != nullptr) {
      for (int i = lane; i < nv; i += 32) {
        const double x = va[i], y = vb[i];
        va[i] = c * x - sn * y;
        vb[i] = sn * x + c * y;
      }
    }
  }
  return off;
}

// PSD-cone projection for 32 < n: one CTA per matrix.  X (symmetric, from the lower triangle) is shifted to
// W = X + sigma I with sigma >= |X|_F >= rho(X), which is positive definite, so that the one-sided Jacobi above applies
// (for an indefinite matrix eigenvalue pairs +-lambda would be singular-value ties that mix the two eigenvectors);
// at convergence column k of W is (lambda_k + sigma) v_k, hence  X+ = sum_{|w_k| > sigma} (|w_k| - sigma) w_k w_k^T / |w_k|^2.
// The absolute error eps * sigma in lambda_k is what LAPACK's eigh guarantees as well.  W lives in shared memory
// (n <= 160) or in the caller's workspace (`work`: one n x n slot per CTA, L2 resident).
constexpr int PSDC_THREADS = 512;

__global__ void __launch_bounds__(PSDC_THREADS) prox_psd_cta_kernel(int n, long long nbatch, long long sb, long long sr,
                                                                    long long sc, const double* __restrict__ h, int hs,
                                                                    const double* __restrict__ mud, double* __restrict__ out,
                                                                    int os, double* __restrict__ work, double tol) {
  extern __shared__ double psdc_sm[];
  __shared__ double red[32];
  __shared__ double offmax_sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = PSDC_THREADS / 32;
  double* coef = psdc_sm;                                              // [n]
  double* W = work != nullptr ? work + (size_t)blockIdx.x * n * n : psdc_sm + n;      // column-major, ld = n
  const int ne = (n + 1) & ~1;
  for (long long mtx = blockIdx.x; mtx < nbatch; mtx += gridDim.x) {
    double fro[1] = {0.0};
    for (int idx = tid; idx < n * n; idx += PSDC_THREADS) {
      const int i = idx % n, j = idx / n;
      const int r = i > j ? i : j, c = i > j ? j : i;
      const long long off = mtx * sb + r * sr + c * sc;
      const double v = -(h[off * hs] / mud[off]);
      W[idx] = v;
      fro[0] += v * v;
    }
    block_sum<1>(fro, red);
    const double sigma = sqrt(fro[0]) * (1.0 + 1e-6) + 1e-300;
    __syncthreads();
    for (int i = tid; i < n; i += PSDC_THREADS) W[(size_t)i * n + i] += sigma;
    __syncthreads();
    for (int sweep = 0; sweep < 60; ++sweep) {
      if (tid == 0) offmax_sh = 0.0;
      __syncthreads();
      double mymax = 0.0;
      for (int r = 0; r < ne - 1; ++r) {
        for (int i = warp; i < ne / 2; i += nw) {
          int p, q;
          rr_pair(ne, r, i, p, q);
          if (p < n && q < n) {
            if (p > q) {
              const int tq = p;
              p = q;
              q = tq;
            }
            mymax = fmax(mymax, jacobi_pair(W + (size_t)p * n, W + (size_t)q * n, n, nullptr, nullptr, 0, lane, tol));
          }
        }
        __syncthreads();
      }
      if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(&offmax_sh), (unsigned long long)__double_as_longlong(mymax));
      __syncthreads();
      const bool done = !(offmax_sh > tol);
      __syncthreads();
      if (done) break;
    }
    // coefficients (|w_k| - sigma) / |w_k|^2 of the positive part
    for (int k = warp; k < n; k += nw) {
      double a = 0.0;
      for (int i = lane; i < n; i += 32) a += W[(size_t)k * n + i] * W[(size_t)k * n + i];
      a = warp_sum(a);
      const double nk = sqrt(a);
      if (lane == 0) coef[k] = nk > sigma ? (nk - sigma) / a : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += PSDC_THREADS) {
      const int i = idx % n, j = idx / n;
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += coef[k] * W[(size_t)k * n + i] * W[(size_t)k * n + j];
      const long long off = mtx * sb + i * sr + j * sc;
      out[off * os] = acc;
      if (os == 2) out[off * 2 + 1] = 0.0;
    }
    __syncthreads();
  }
}

// SVD of ONE m x n matrix (m >= 1, any n) spread over a cooperative grid: Wt holds the columns of the input as rows
// (n rows of length m, row stride ldw), Vt the columns of V as rows (n x n, identity on entry; may be NULL).
// On exit row k of Wt is sigma_k u_k, row k of Vt is v_k, sv[k] = sigma_k (unsorted), info[0] = sweeps used (negative: not
// converged within max_sweeps).  offmax: max_sweeps doubles of scratch, zero-initialised.
constexpr int SVDJ_THREADS = 256;

__global__ void __launch_bounds__(SVDJ_THREADS) svd_jacobi_kernel(int m, int n, double* __restrict__ Wt, int ldw,
                                                                  double* __restrict__ Vt, int ldv, double* __restrict__ sv,
                                                                  double* __restrict__ offmax, double tol, int max_sweeps,
                                                                  int* __restrict__ info) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * SVDJ_THREADS + threadIdx.x) >> 5, tw = gridDim.x * (SVDJ_THREADS / 32);
  const int ne = (n + 1) & ~1;
  int used = -max_sweeps;
  // |K|_F^2 -> the noise floor of a column (sv[] doubles as scratch: it is written last)
  if (blockIdx.x == 0 && threadIdx.x == 0) sv[0] = 0.0;
  grid.sync();
  {
    double f = 0.0;
    for (int k = gw; k < n; k += tw)
      for (int i = lane; i < m; i += 32) f += Wt[(size_t)k * ldw + i] * Wt[(size_t)k * ldw + i];
    f = warp_sum(f);
    if (lane == 0 && f != 0.0) atomicAdd(sv, f);
  }
  grid.sync();
  const double zero2 = 256.0 * 4.930380657631324e-32 * (double)m * __ldcg(sv);      // (16 eps sqrt(m) |K|_F)^2
  grid.sync();
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double mymax = 0.0;
    for (int r = 0; r < ne - 1; ++r) {
      for (int i = gw; i < ne / 2; i += tw) {
        int p, q;
        rr_pair(ne, r, i, p, q);
        if (p < n && q < n) {
          if (p > q) {
            const int tq = p;
            p = q;
            q = tq;
          }
          mymax = fmax(mymax, jacobi_pair(Wt + (size_t)p * ldw, Wt + (size_t)q * ldw, m, Vt ? Vt + (size_t)p * ldv : nullptr,
                                          Vt ? Vt + (size_t)q * ldv : nullptr, n, lane, tol, zero2));
        }
      }
      grid.sync();
    }
    if (lane == 0 && mymax > 0.0)
      atomicMax(reinterpret_cast<unsigned long long*>(offmax + sweep), (unsigned long long)__double_as_longlong(mymax));
    grid.sync();
    const double om = __ldcg(offmax + sweep);
    if (!(om > tol)) {
      used = sweep + 1;
      break;
    }
  }
  for (int k = gw; k < n; k += tw) {
    double a = 0.0;
    for (int i = lane; i < m; i += 32) a += Wt[(size_t)k * ldw + i] * Wt[(size_t)k * ldw + i];
    a = warp_sum(a);
    if (lane == 0) sv[k] = sqrt(a);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && info != nullptr) info[0] = used;
}

__global__ void __launch_bounds__(256) sumsq_stage1(long long n, const double* __restrict__ x,
                                                    const double* __restrict__ y, double* __restrict__ part) {
  __shared__ double scratch[32];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long b0 = per * blockIdx.x, b1 = min(n, b0 + per);
  double v[1] = {0.0};
  for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    double t = x[i];
    if (y) t -= y[i];
    v[0] += t * t;
  }
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = v[0];
}

__global__ void __launch_bounds__(256) sumsq_stage2(int nparts, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double scratch[32];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) v[0] += part[i];
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) out[0] = v[0];
}

// the six squared norms of a coupled pair in one pass: |p1-p2|, |p1|, |p2| over np doubles, |d1-d2|, |d1|, |d2| over nd
__global__ void __launch_bounds__(256) pair_norms_stage1(long long np, const double* __restrict__ p1,
                                                         const double* __restrict__ p2, long long nd,
                                                         const double* __restrict__ d1, const double* __restrict__ d2,
                                                         double* __restrict__ part) {
  __shared__ double scratch[6 * 32];
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  {
    const long long per = (np + gridDim.x - 1) / gridDim.x;
    const long long b0 = per * blockIdx.x, b1 = min(np, b0 + per);
    for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
      const double a = p1[i], b = p2[i];
      v[0] += (a - b) * (a - b);
      v[1] += a * a;
      v[2] += b * b;
    }
  }
  {
    const long long per = (nd + gridDim.x - 1) / gridDim.x;
    const long long b0 = per * blockIdx.x, b1 = min(nd, b0 + per);
    for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
      const double a = d1[i], b = d2[i];
      v[3] += (a - b) * (a - b);
      v[4] += a * a;
      v[5] += b * b;
    }
  }
  block_sum<6>(v, scratch);
  if (threadIdx.x < 6) part[blockIdx.x * 6 + threadIdx.x] = v[threadIdx.x];
}

__global__ void __launch_bounds__(256) pair_norms_stage2(int nparts, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double scratch[6 * 32];
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] += part[i * 6 + q];
  }
  block_sum<6>(v, scratch);
  if (threadIdx.x < 6) out[threadIdx.x] = v[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// general inverse: Gauss-Jordan with partial pivoting on the augmented matrix [A | I], one CTA
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) inverse_kernel(int n, const T* __restrict__ A, int lda, T* __restrict__ Ainv,
                                                       int ldi, T* __restrict__ W, int* __restrict__ info) {
  extern __shared__ unsigned char smem_raw[];
  T* colk = reinterpret_cast<T*>(smem_raw);  // n
  __shared__ double s_best[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int w2 = 2 * n;
  for (int idx = tid; idx < n * w2; idx += nt) {
    const int i = idx / w2, j = idx % w2;
    T v = num<T>::zero();
    if (j < n) v = A[(size_t)i * lda + j];
    else if (j - n == i) v = num<T>::one();
    W[idx] = v;
  }
  if (tid == 0) info[0] = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    // pivot search: argmax |W[i][k]|, i >= k (ties -> smallest i, as LAPACK's idamax)
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += nt) {
      const double a = num<T>::abs2(W[(size_t)i * w2 + k]);
      if (a > best) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_idx[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = s_best[0];
      int ix = s_idx[0];
      for (int w = 1; w < (nt + 31) / 32; ++w)
        if (s_best[w] > b || (s_best[w] == b && s_idx[w] < ix)) { b = s_best[w]; ix = s_idx[w]; }
      s_piv = ix;
      if (!(b > 0.0)) info[0] = k + 1;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k) {
      for (int j = tid; j < w2; j += nt) {
        const T a = W[(size_t)k * w2 + j];
        W[(size_t)k * w2 + j] = W[(size_t)p * w2 + j];
        W[(size_t)p * w2 + j] = a;
      }
    }
    __syncthreads();
    const T piv = W[(size_t)k * w2 + k];
    for (int i = tid; i < n; i += nt) colk[i] = W[(size_t)i * w2 + k];
    __syncthreads();
    for (int j = tid; j < w2; j += nt) W[(size_t)k * w2 + j] = num<T>::div(W[(size_t)k * w2 + j], piv);
    __syncthreads();
    // eliminate column k from every other row; columns < k of the left half are already e_j
    for (int idx = tid; idx < n * (w2 - k); idx += nt) {
      const int i = idx / (w2 - k), j = k + idx % (w2 - k);
      if (i == k) continue;
      const size_t o = (size_t)i * w2 + j;
      W[o] = num<T>::sub(W[o], num<T>::mul(colk[i], W[(size_t)k * w2 + j]));
    }
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx % n;
    Ainv[(size_t)i * ldi + j] = W[(size_t)i * w2 + n + j];
  }
}

// ---------------------------------------------------------------------------------------------
// batched in-place inverse of real SPD matrices (Gauss-Jordan without pivoting), one CTA each
// ---------------------------------------------------------------------------------------------
template <bool IN_SMEM>
__global__ void __launch_bounds__(512) spd_inverse_kernel(int n, double* __restrict__ Aall, long long bstride, int lda,
                                                          const int* __restrict__ mask, int* __restrict__ info) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  if (mask && mask[b] == 0) return;
  double* Ag = Aall + (size_t)b * bstride;
  double* rowk = sm;          // n
  double* colk = sm + n;      // n
  double* a = IN_SMEM ? (sm + 2 * n) : Ag;
  const int ld = IN_SMEM ? n : lda;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (IN_SMEM) {
    for (int idx = tid; idx < n * n; idx += nt) a[idx] = Ag[(size_t)(idx / n) * lda + idx % n];
  }
  __syncthreads();
  int bad = 0;
  for (int k = 0; k < n; ++k) {
    const double p = a[(size_t)k * ld + k];
    if (!(p > 0.0)) bad = k + 1;
    const double ip = 1.0 / p;
    for (int j = tid; j < n; j += nt) {
      rowk[j] = (j == k ? 1.0 : a[(size_t)k * ld + j]) * ip;
      colk[j] = (j == k ? 0.0 : a[(size_t)j * ld + k]);
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx - i * n;
      const size_t o = (size_t)i * ld + j;
      if (i == k) {
        a[o] = rowk[j];
      } else {
        const double base = (j == k) ? 0.0 : a[o];
        a[o] = base - colk[i] * rowk[j];
      }
    }
    __syncthreads();
  }
  if (IN_SMEM) {
    // symmetrise (the exact inverse is symmetric; rounding is not) and write back
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx - i * n;
      Ag[(size_t)i * lda + j] = 0.5 * (a[(size_t)i * n + j] + a[(size_t)j * n + i]);
    }
  }
  if (tid == 0 && info) info[b] = bad;
}

// ---------------------------------------------------------------------------------------------
// batched in-place inverse of real SPD matrices of order n <= 128 on the FP64 tensor cores:
// block Gauss-Jordan with 8x8 blocks, the whole matrix REGISTER-resident in the CTA.
//   warp w owns block row w: its 8 x n strip lives in mma C-fragment layout (lane 4g+t holds
//   (row g, columns 2t, 2t+1) of every 8x8 tile), 2*NB doubles per lane.
//   step k:  P = A_kk^-1 (warp k, in-warp shuffles);  T_w = A_wk P;  A_wj -= T_w A_kj (j != k);
//            A_kj = P A_kj;  A_wk = -T_w;  A_kk = P.
//   The C fragment of a tile is reused as the A operand of the next MMA by permuting the k index
//   (k-slot (e, t) <-> column 2t+e), so nothing is transposed: the row panel A_k travels through
//   shared memory once per step in the two operand layouts the MMAs need.
// One CTA (NB warps) per matrix, 2 __syncthreads per block step.  No pivoting (SPD).
// ---------------------------------------------------------------------------------------------
constexpr int SPDI_MAXB = 16;     // up to 16 block rows of 8 -> n <= 128

__global__ void __launch_bounds__(SPDI_MAXB * 32, 1)
    spd_inverse_dmma_kernel(int n, double* __restrict__ Aall, long long bstride, int lda, const int* __restrict__ mask,
                            int* __restrict__ info) {
  __shared__ __align__(16) double panF[SPDI_MAXB * 64];      // row panel, fragment-major: [j][lane][e] = A_k[2t+e][8j+g]
  __shared__ double panC[8 * (8 * SPDI_MAXB + 4)];           // row panel, canonical, row pitch 8*NBmax+4 (conflict-free column reads)
  __shared__ __align__(16) double Psm[64];                    // P = A_kk^-1, canonical 8x8
  __shared__ int bad_sm;
  constexpr int PITCH = 8 * SPDI_MAXB + 4;
  const int b = blockIdx.x;
  if (mask && mask[b] == 0) return;
  double* Ag = Aall + (size_t)b * bstride;
  const int NB = (n + 7) >> 3;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) bad_sm = 0;

  // ---- load the strip (identity on the padding keeps the padded matrix SPD)
  double c[SPDI_MAXB][2];
  const int row = 8 * w + g;
#pragma unroll
  for (int j = 0; j < SPDI_MAXB; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * j + 2 * t + e;
      double v = (row == col) ? 1.0 : 0.0;
      if (j < NB && row < n && col < n) v = Ag[(size_t)row * lda + col];
      c[j][e] = v;
    }
  }
  __syncthreads();

  for (int k = 0; k < NB; ++k) {
    if (w == k) {
      // ---- publish the (old) row panel in both operand layouts
#pragma unroll
      for (int j = 0; j < SPDI_MAXB; ++j) {
        if (j < NB) {
          // canonical: rows g, columns 8j+2t+e
          panC[g * PITCH + 8 * j + 2 * t] = c[j][0];
          panC[g * PITCH + 8 * j + 2 * t + 1] = c[j][1];
        }
      }
      // ---- P = A_kk^-1 by in-place Gauss-Jordan on the C fragment (lane holds (g, 2t), (g, 2t+1))
      double a0, a1;
      {
        double akk[SPDI_MAXB][2];
#pragma unroll
        for (int j = 0; j < SPDI_MAXB; ++j) { akk[j][0] = c[j][0]; akk[j][1] = c[j][1]; }
        a0 = 0.0; a1 = 0.0;
#pragma unroll
        for (int j = 0; j < SPDI_MAXB; ++j) if (j == k) { a0 = akk[j][0]; a1 = akk[j][1]; }
      }
      int bad = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int src_pp = 4 * p + (p >> 1);                  // lane holding (p, p)
        const double piv = __shfl_sync(0xffffffffu, (p & 1) ? a1 : a0, src_pp);
        if (!(piv > 0.0)) bad = 8 * k + p + 1;
        const double ip = 1.0 / piv;
        // pivot row, my two columns; pivot column, my row
        double r0 = __shfl_sync(0xffffffffu, a0, 4 * p + t);
        double r1 = __shfl_sync(0xffffffffu, a1, 4 * p + t);
        double cg = __shfl_sync(0xffffffffu, (p & 1) ? a1 : a0, 4 * g + (p >> 1));
        if (2 * t == p) r0 = 1.0;
        if (2 * t + 1 == p) r1 = 1.0;
        r0 *= ip;
        r1 *= ip;
        if (g == p) {
          a0 = r0;
          a1 = r1;
        } else {
          const double b0 = (2 * t == p) ? 0.0 : a0;
          const double b1 = (2 * t + 1 == p) ? 0.0 : a1;
          a0 = b0 - cg * r0;
          a1 = b1 - cg * r1;
        }
      }
      Psm[g * 8 + 2 * t] = a0;
      Psm[g * 8 + 2 * t + 1] = a1;
      if (bad && lane == 0) bad_sm = bad;
    }
    __syncthreads();
    // fragment-major copy of the panel (cooperative): panF[j][lane'][e] = A_k[2t'+e][8j+g']
    for (int idx = tid; idx < NB * 64; idx += blockDim.x) {
      const int e = idx & 1, ln = (idx >> 1) & 31, j = idx >> 6;
      panF[idx] = panC[(2 * (ln & 3) + e) * PITCH + 8 * j + (ln >> 2)];
    }
    __syncthreads();

    if (w < NB) {
      if (w != k) {
        // ---- T = A_wk P   (A operand: my C fragment of tile (w,k), k-slot (e,t) <-> column 2t+e)
        double ak0 = 0.0, ak1 = 0.0;
#pragma unroll
        for (int j = 0; j < SPDI_MAXB; ++j) if (j == k) { ak0 = c[j][0]; ak1 = c[j][1]; }
        double T0 = 0.0, T1 = 0.0;
        dmma(T0, T1, ak0, Psm[(2 * t) * 8 + g]);
        dmma(T0, T1, ak1, Psm[(2 * t + 1) * 8 + g]);
        const double nT0 = -T0, nT1 = -T1;
        // ---- A_wj -= T A_kj  for j != k;  A_wk = -T
#pragma unroll
        for (int j = 0; j < SPDI_MAXB; ++j) {
          if (j < NB) {
            if (j == k) {
              c[j][0] = nT0;
              c[j][1] = nT1;
            } else {
              const double2 bb = *reinterpret_cast<const double2*>(panF + (j * 32 + lane) * 2);
              dmma(c[j][0], c[j][1], nT0, bb.x);
              dmma(c[j][0], c[j][1], nT1, bb.y);
            }
          }
        }
      } else {
        // ---- row k:  A_kj = P A_kj (j != k),  A_kk = P     (A operand: P[g][4s+t]; B: A_k[4s+t][8j+g])
        const double p0 = Psm[g * 8 + t], p1 = Psm[g * 8 + 4 + t];
#pragma unroll
        for (int j = 0; j < SPDI_MAXB; ++j) {
          if (j < NB) {
            if (j == k) {
              c[j][0] = Psm[g * 8 + 2 * t];
              c[j][1] = Psm[g * 8 + 2 * t + 1];
            } else {
              double d0 = 0.0, d1 = 0.0;
              dmma(d0, d1, p0, panC[t * PITCH + 8 * j + g]);
              dmma(d0, d1, p1, panC[(4 + t) * PITCH + 8 * j + g]);
              c[j][0] = d0;
              c[j][1] = d1;
            }
          }
        }
      }
    }
    __syncthreads();     // the panel buffers are rewritten in the next step
  }

  // ---- write back (the exact inverse is symmetric; rounding differences stay at the 1e-16 level)
#pragma unroll
  for (int j = 0; j < SPDI_MAXB; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * j + 2 * t + e;
      if (j < NB && row < n && col < n) Ag[(size_t)row * lda + col] = c[j][e];
    }
  }
  if (tid == 0 && info) info[b] = bad_sm;
}

// ---------------------------------------------------------------------------------------------
// the same block Gauss-Jordan for 128 < n <= 512: the matrix stays in global memory (L2-resident:
// <= 2 MB), one CTA of 16 warps per matrix, warp w owns block rows w, w+16, ...; every block step
// streams the strips it updates through the mma C fragments (load, 2 DMMAs per tile, store).
// ---------------------------------------------------------------------------------------------
constexpr int SPDG_MAXB = 64;      // up to 64 block rows of 8 -> n <= 512
constexpr int SPDG_WARPS = 16;

__global__ void __launch_bounds__(SPDG_WARPS * 32, 1)
    spd_inverse_dmma_global_kernel(int n, double* __restrict__ Aall, long long bstride, int lda,
                                   const int* __restrict__ mask, int* __restrict__ info) {
  extern __shared__ __align__(16) double spdg_sm[];
  constexpr int PITCH = 8 * SPDG_MAXB + 4;
  double* panF = spdg_sm;                          // [NB][32][2]   row panel, fragment-major
  double* panC = panF + SPDG_MAXB * 64;            // [8][PITCH]    row panel, canonical
  double* Psm = panC + 8 * PITCH;                  // [64]          P = A_kk^-1
  __shared__ int bad_sm;
  const int b = blockIdx.x;
  if (mask && mask[b] == 0) return;
  double* Ag = Aall + (size_t)b * bstride;
  const int NB = (n + 7) >> 3;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  if (tid == 0) bad_sm = 0;
  __syncthreads();
  // element (r, c) of the padded matrix (identity on the padding)
  auto ld = [&](int r, int c) -> double {
    return (r < n && c < n) ? Ag[(size_t)r * lda + c] : (r == c ? 1.0 : 0.0);
  };
  auto st = [&](int r, int c, double v) {
    if (r < n && c < n) Ag[(size_t)r * lda + c] = v;
  };

  for (int k = 0; k < NB; ++k) {
    // ---- (old) row panel -> shared memory, canonical layout (all warps cooperate)
    for (int idx = tid; idx < 8 * 8 * NB; idx += blockDim.x) {
      const int r = idx / (8 * NB), c = idx - r * (8 * NB);
      panC[r * PITCH + c] = ld(8 * k + r, c);
    }
    __syncthreads();
    if (w == 0) {
      // ---- P = A_kk^-1 by in-place Gauss-Jordan on the C fragment (lane holds (g, 2t), (g, 2t+1))
      double a0 = panC[g * PITCH + 8 * k + 2 * t], a1 = panC[g * PITCH + 8 * k + 2 * t + 1];
      int bad = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const double piv = __shfl_sync(0xffffffffu, (p & 1) ? a1 : a0, 4 * p + (p >> 1));
        if (!(piv > 0.0)) bad = 8 * k + p + 1;
        const double ip = 1.0 / piv;
        double r0 = __shfl_sync(0xffffffffu, a0, 4 * p + t);
        double r1 = __shfl_sync(0xffffffffu, a1, 4 * p + t);
        const double cg = __shfl_sync(0xffffffffu, (p & 1) ? a1 : a0, 4 * g + (p >> 1));
        if (2 * t == p) r0 = 1.0;
        if (2 * t + 1 == p) r1 = 1.0;
        r0 *= ip;
        r1 *= ip;
        if (g == p) {
          a0 = r0;
          a1 = r1;
        } else {
          const double b0 = (2 * t == p) ? 0.0 : a0;
          const double b1 = (2 * t + 1 == p) ? 0.0 : a1;
          a0 = b0 - cg * r0;
          a1 = b1 - cg * r1;
        }
      }
      Psm[g * 8 + 2 * t] = a0;
      Psm[g * 8 + 2 * t + 1] = a1;
      if (bad && lane == 0) bad_sm = bad;
    }
    // fragment-major copy of the panel: panF[j][lane'][e] = A_k[2t'+e][8j+g']
    for (int idx = tid; idx < NB * 64; idx += blockDim.x) {
      const int e = idx & 1, ln = (idx >> 1) & 31, j = idx >> 6;
      panF[idx] = panC[(2 * (ln & 3) + e) * PITCH + 8 * j + (ln >> 2)];
    }
    __syncthreads();

    for (int i = w; i < NB; i += SPDG_WARPS) {
      const int row = 8 * i + g;
      if (i != k) {
        // T = A_ik P, then A_ij -= T A_kj (j != k), A_ik = -T
        const double ak0 = ld(row, 8 * k + 2 * t), ak1 = ld(row, 8 * k + 2 * t + 1);
        double T0 = 0.0, T1 = 0.0;
        dmma(T0, T1, ak0, Psm[(2 * t) * 8 + g]);
        dmma(T0, T1, ak1, Psm[(2 * t + 1) * 8 + g]);
        const double nT0 = -T0, nT1 = -T1;
        for (int j = 0; j < NB; ++j) {
          const int col = 8 * j + 2 * t;
          if (j == k) {
            st(row, col, nT0);
            st(row, col + 1, nT1);
          } else {
            double c0 = ld(row, col), c1 = ld(row, col + 1);
            const double2 bb = *reinterpret_cast<const double2*>(panF + (j * 32 + lane) * 2);
            dmma(c0, c1, nT0, bb.x);
            dmma(c0, c1, nT1, bb.y);
            st(row, col, c0);
            st(row, col + 1, c1);
          }
        }
      } else {
        // row k:  A_kj = P A_kj (j != k),  A_kk = P
        const double p0 = Psm[g * 8 + t], p1 = Psm[g * 8 + 4 + t];
        for (int j = 0; j < NB; ++j) {
          const int col = 8 * j + 2 * t;
          if (j == k) {
            st(row, col, Psm[g * 8 + 2 * t]);
            st(row, col + 1, Psm[g * 8 + 2 * t + 1]);
          } else {
            double d0 = 0.0, d1 = 0.0;
            dmma(d0, d1, p0, panC[t * PITCH + 8 * j + g]);
            dmma(d0, d1, p1, panC[(4 + t) * PITCH + 8 * j + g]);
            st(row, col, d0);
            st(row, col + 1, d1);
          }
        }
      }
    }
    __syncthreads();     // the strips are in memory before the next panel is read
  }
  if (tid == 0 && info) info[b] = bad_sm;
}

}  // namespace admm

using namespace admm;

// ---------------------------------------------------------------------------------------------
// Hermitian positive definite complex128 inverse on the tensor cores, batched: H = X + iY (X symmetric, Y
// antisymmetric) is HPD iff its real form [[X, -Y], [Y, X]] (2n x 2n) is SPD, and the inverse of the real form is the
// real form of H^-1.  So the batched SPD block Gauss-Jordan on DMMA above does the work; these two kernels only
// move the numbers (interleaved complex <-> real form in `work`).
// ---------------------------------------------------------------------------------------------
__global__ void hpd_embed_kernel(int n, int nbatch, const cplx* __restrict__ A, long long bstride, int lda,
                                 double* __restrict__ W, const int* __restrict__ mask) {
  const long long tot = (long long)nbatch * n * n;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(q / ((long long)n * n));
    if (mask != nullptr && mask[b] == 0) continue;
    const int r = (int)(q % ((long long)n * n));
    const int i = r / n, j = r % n;
    const cplx v = A[(size_t)b * bstride + (size_t)i * lda + j];
    double* w = W + (size_t)b * 4 * n * n;
    const int n2 = 2 * n;
    w[(size_t)i * n2 + j] = v.x;
    w[(size_t)i * n2 + n + j] = -v.y;
    w[(size_t)(n + i) * n2 + j] = v.y;
    w[(size_t)(n + i) * n2 + n + j] = v.x;
  }
}

__global__ void hpd_extract_kernel(int n, int nbatch, const double* __restrict__ W, cplx* __restrict__ A, long long bstride,
                                   int lda, const int* __restrict__ mask, const int* __restrict__ info) {
  const long long tot = (long long)nbatch * n * n;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(q / ((long long)n * n));
    if (mask != nullptr && mask[b] == 0) continue;
    if (info != nullptr && info[b] != 0) continue;       // not positive definite: the input stays untouched
    const int r = (int)(q % ((long long)n * n));
    const int i = r / n, j = r % n;
    const double* w = W + (size_t)b * 4 * n * n;
    const int n2 = 2 * n;
    A[(size_t)b * bstride + (size_t)i * lda + j] = {w[(size_t)i * n2 + j], w[(size_t)(n + i) * n2 + j]};
  }
}

extern "C" {

int admm_abi_version(void) { return ADMM_ABI_VERSION; }

const char* admm_last_error(void) { return admm::g_err; }

int admm_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return ADMM_OK;
}

int admm_gemm(int is_complex, int op_a, int m, int n, int k, const void* A, int lda, const void* B, int ldb,
              void* C, int ldc, admm_stream_t stream) {
  ADMM_REQUIRE(m >= 0 && n >= 0 && k >= 0 && op_a >= 0 && op_a <= 2, ADMM_EINVAL, "admm_gemm: bad dims/op");
  if (m == 0 || n == 0) return ADMM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!is_complex && m >= 32 && n >= 32 && k >= 16 && !getenv("ADMM_GEMM_SCALAR")) {
    // tensor-core path (real data; a conjugate transpose of a real matrix is its transpose)
    dim3 grid(ceil_div(n, 64), ceil_div(m, 64));
    const double* a = static_cast<const double*>(A);
    const double* b = static_cast<const double*>(B);
    double* c = static_cast<double*>(C);
    if (op_a == ADMM_OP_N) gemm_dmma_kernel<ADMM_OP_N, false><<<grid, 128, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
    else gemm_dmma_kernel<ADMM_OP_T, false><<<grid, 128, 0, s>>>(m, n, k, a, lda, b, ldb, c, ldc);
    return check_launch("admm_gemm(dmma)");
  }
  if (is_complex && m >= 32 && n >= 16 && k >= 8 && !getenv("ADMM_GEMM_SCALAR")) {
    // complex128 on the tensor cores: the real form of the product (see gemm_dmma_kernel), dimensions in doubles
    dim3 grid(ceil_div(2 * n, 64), ceil_div(m, 64));
    const double* a = static_cast<const double*>(A);
    const double* b = static_cast<const double*>(B);
    double* c = static_cast<double*>(C);
    if (op_a == ADMM_OP_N)
      gemm_dmma_kernel<ADMM_OP_N, true><<<grid, 128, 0, s>>>(m, 2 * n, 2 * k, a, 2 * lda, b, 2 * ldb, c, 2 * ldc);
    else if (op_a == ADMM_OP_T)
      gemm_dmma_kernel<ADMM_OP_T, true><<<grid, 128, 0, s>>>(m, 2 * n, 2 * k, a, 2 * lda, b, 2 * ldb, c, 2 * ldc);
    else
      gemm_dmma_kernel<ADMM_OP_H, true><<<grid, 128, 0, s>>>(m, 2 * n, 2 * k, a, 2 * lda, b, 2 * ldb, c, 2 * ldc);
    return check_launch("admm_gemm(zdmma)");
  }
  return is_complex ? launch_gemm<cplx>(op_a, m, n, k, A, lda, B, ldb, C, ldc, s)
                    : launch_gemm<double>(op_a, m, n, k, A, lda, B, ldb, C, ldc, s);
}

static int ew_grid(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 16)); }

int admm_diag_mul(int is_complex, int rows_out, int nd, int ncols, const void* d, const void* V, int ldv, void* out,
                  int ldo, admm_stream_t stream) {
  ADMM_REQUIRE(rows_out >= 0 && nd >= 0 && ncols >= 0, ADMM_EINVAL, "admm_diag_mul: bad dims");
  if (rows_out == 0 || ncols == 0) return ADMM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int g = ew_grid((long long)rows_out * ncols);
  if (is_complex)
    diag_mul_kernel<cplx><<<g, 256, 0, s>>>(rows_out, nd, ncols, (const cplx*)d, (const cplx*)V, ldv, (cplx*)out, ldo);
  else
    diag_mul_kernel<double><<<g, 256, 0, s>>>(rows_out, nd, ncols, (const double*)d, (const double*)V, ldv,
                                               (double*)out, ldo);
  return check_launch("admm_diag_mul");
}

int admm_axpby(long long n, double a, const double* x, double b, const double* y, double* out, admm_stream_t stream) {
  if (n <= 0) return ADMM_OK;
  axpby_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, a, x, b, y, out);
  return check_launch("admm_axpby");
}

int admm_ewise_unary(int op, int is_complex, long long n, const void* x, void* out, admm_stream_t stream) {
  ADMM_REQUIRE(op == 0 || op == 1, ADMM_EINVAL, "admm_ewise_unary: op must be 0 (reciprocal) or 1 (conjugate)");
  if (n <= 0) return ADMM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int g = ew_grid(n);
  if (is_complex) {
    if (op == 0) ewise_unary_kernel<cplx, 0><<<g, 256, 0, s>>>(n, (const cplx*)x, (cplx*)out);
    else ewise_unary_kernel<cplx, 1><<<g, 256, 0, s>>>(n, (const cplx*)x, (cplx*)out);
  } else {
    if (op == 0) ewise_unary_kernel<double, 0><<<g, 256, 0, s>>>(n, (const double*)x, (double*)out);
    else ewise_unary_kernel<double, 1><<<g, 256, 0, s>>>(n, (const double*)x, (double*)out);
  }
  return check_launch("admm_ewise_unary");
}

int admm_prox_l1(long long n, const double* h, int h_stride, const double* mu_diag, double alpha, double* out,
                 int out_stride, admm_stream_t stream) {
  ADMM_REQUIRE(alpha > 0, ADMM_EINVAL, "admm_prox_l1: alpha must be > 0");
  if (n <= 0) return ADMM_OK;
  prox_l1_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, h, h_stride, mu_diag, alpha, out,
                                                                           out_stride);
  return check_launch("admm_prox_l1");
}

int admm_prox_nonneg(long long n, const double* h, int h_stride, const double* mu_diag, double* out, int out_stride,
                     admm_stream_t stream) {
  if (n <= 0) return ADMM_OK;
  prox_nonneg_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, h, h_stride, mu_diag, out,
                                                                               out_stride);
  return check_launch("admm_prox_nonneg");
}

int admm_prox_psd_work_doubles(int n, long long nbatch, long long* grid_out) {
  const long long grid = std::max<long long>(1, std::min<long long>(nbatch, 148LL * 2));
  if (grid_out != nullptr) *grid_out = grid;
  return n > 160 ? 1 : 0;          // (1: a workspace of grid * n * n doubles is needed)
}

int admm_prox_psd(int n, long long nbatch, long long stride_batch, long long stride_row, long long stride_col,
                  const double* h, int h_stride, const double* mu_diag, double* out, int out_stride, double* work,
                  admm_stream_t stream) {
  ADMM_REQUIRE(n >= 1 && n <= 1024, ADMM_EUNSUPPORTED, "admm_prox_psd: matrix order %d not supported (1..1024)", n);
  if (nbatch <= 0) return ADMM_OK;
  if (n > 32) {
    // one CTA per matrix, one-sided Jacobi on the shifted matrix (shared memory up to n = 160, the caller's
    // L2-resident workspace beyond)
    long long grid = 1;
    const int need_work = admm_prox_psd_work_doubles(n, nbatch, &grid);
    ADMM_REQUIRE(!need_work || work != nullptr, ADMM_EINVAL,
                 "admm_prox_psd: n=%d needs a workspace of %lld doubles (admm_prox_psd_work_doubles)", n, grid * n * n);
    const size_t smem = (size_t)(need_work ? n : n + (size_t)n * n) * sizeof(double);
    static std::map<int, size_t> configured;
    size_t& cfgd = configured[cur_dev()];
    if (smem > cfgd) {
      cudaFuncSetAttribute(prox_psd_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cfgd = smem;
    }
    const double tol = 2.0 * sqrt((double)n) * 2.220446049250313e-16;
    prox_psd_cta_kernel<<<(int)grid, PSDC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        n, nbatch, stride_batch, stride_row, stride_col, h, h_stride, mu_diag, out, out_stride, need_work ? work : nullptr, tol);
    return check_launch("admm_prox_psd(cta)");
  }
  const size_t smem = (size_t)PSD_WARPS * 2 * n * (n | 1) * sizeof(double);
  const int grid = (int)std::max<long long>(1, std::min<long long>((nbatch + PSD_WARPS - 1) / PSD_WARPS, 148LL * 16));
  prox_psd_kernel<<<grid, PSD_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(n, nbatch, stride_batch, stride_row,
                                                                                    stride_col, h, h_stride, mu_diag, out,
                                                                                    out_stride);
  return check_launch("admm_prox_psd");
}

int admm_sumsq(long long n, const double* x, const double* y, double* out, double* scratch, admm_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int parts = (int)std::max<long long>(1, std::min<long long>((n + 4095) / 4096, 1024));
  sumsq_stage1<<<parts, 256, 0, s>>>(n, x, y, scratch);
  sumsq_stage2<<<1, 256, 0, s>>>(parts, scratch, out);
  return check_launch("admm_sumsq");
}

int admm_pair_norms(long long np, const double* p1, const double* p2, long long nd, const double* d1, const double* d2,
                    double* out, double* scratch, admm_stream_t stream) {
  ADMM_REQUIRE(np >= 0 && nd >= 0 && out != nullptr && scratch != nullptr, ADMM_EINVAL, "admm_pair_norms: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n = std::max(np, nd);
  const int parts = (int)std::max<long long>(1, std::min<long long>((n + 4095) / 4096, 1024));
  pair_norms_stage1<<<parts, 256, 0, s>>>(np, p1, p2, nd, d1, d2, scratch);
  pair_norms_stage2<<<1, 256, 0, s>>>(parts, scratch, out);
  return check_launch("admm_pair_norms");
}

int admm_inverse(int is_complex, int n, const void* A, int lda, void* Ainv, int ldi, void* work, int* info,
                 admm_stream_t stream) {
  ADMM_REQUIRE(n > 0 && n <= 2048, ADMM_EINVAL, "admm_inverse: n=%d out of range (1..2048)", n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = n <= 32 ? 256 : 1024;
  if (is_complex)
    inverse_kernel<cplx><<<1, threads, n * sizeof(cplx), s>>>(n, (const cplx*)A, lda, (cplx*)Ainv, ldi, (cplx*)work, info);
  else
    inverse_kernel<double><<<1, threads, n * sizeof(double), s>>>(n, (const double*)A, lda, (double*)Ainv, ldi,
                                                                  (double*)work, info);
  return check_launch("admm_inverse");
}

int admm_spd_inverse_batched(int n, int nbatch, double* A, long long batch_stride, int lda, const int* mask, int* info,
                             admm_stream_t stream) {
  ADMM_REQUIRE(n > 0 && n <= 4096 && nbatch >= 0 && lda >= n, ADMM_EINVAL, "admm_spd_inverse_batched: bad dims");
  if (nbatch == 0) return ADMM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n <= 8 * SPDI_MAXB && !getenv("ADMM_SPD_SCALAR")) {
    const int NB = (n + 7) / 8;
    spd_inverse_dmma_kernel<<<nbatch, 32 * NB, 0, s>>>(n, A, batch_stride, lda, mask, info);
  } else if (n <= 8 * SPDG_MAXB && !getenv("ADMM_SPD_SCALAR")) {
    const size_t smem = (size_t)(SPDG_MAXB * 64 + 8 * (8 * SPDG_MAXB + 4) + 64) * sizeof(double);
    static bool attr_g = false;
    if (!attr_g) {
      cudaFuncSetAttribute(spd_inverse_dmma_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_g = true;
    }
    spd_inverse_dmma_global_kernel<<<nbatch, SPDG_WARPS * 32, smem, s>>>(n, A, batch_stride, lda, mask, info);
  } else if (n <= 128) {
    const size_t smem = (size_t)(n * n + 2 * n) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(spd_inverse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
      attr_set = true;
    }
    const int threads = n <= 16 ? 64 : (n <= 48 ? 256 : 512);
    spd_inverse_kernel<true><<<nbatch, threads, smem, s>>>(n, A, batch_stride, lda, mask, info);
  } else {
    spd_inverse_kernel<false><<<nbatch, 512, 2 * n * sizeof(double), s>>>(n, A, batch_stride, lda, mask, info);
  }
  return check_launch("admm_spd_inverse_batched");
}

int admm_svd_jacobi(int m, int n, double* Wt, int ldw, double* Vt, int ldv, double* sv, double* scratch, int max_sweeps,
                    int* info, admm_stream_t stream) {
  ADMM_REQUIRE(m >= 1 && n >= 1 && ldw >= m && (Vt == nullptr || ldv >= n) && sv != nullptr && scratch != nullptr &&
                   max_sweeps >= 1 && max_sweeps <= 64,
               ADMM_EINVAL, "admm_svd_jacobi: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, svd_jacobi_kernel, SVDJ_THREADS, 0) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev()) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    set_error("admm_svd_jacobi: occupancy query failed");
    return ADMM_ECUDA;
  }
  const int pairs = (n + 1) / 2, wpc = SVDJ_THREADS / 32;
  const int grid = std::max(1, std::min(per_sm * sms, ceil_div(pairs, wpc)));
  cudaMemsetAsync(scratch, 0, (size_t)max_sweeps * sizeof(double), s);
  const double tol = 2.0 * sqrt((double)m) * 2.220446049250313e-16;
  cudaError_t e = launch_coop(svd_jacobi_kernel, dim3(grid), dim3(SVDJ_THREADS), 0, s, false, m, n, Wt, ldw, Vt, ldv, sv, scratch,
                              tol, max_sweeps, info);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("admm_svd_jacobi: cooperative launch failed: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return check_launch("admm_svd_jacobi");
}

int admm_hpd_inverse_batched(int n, int nbatch, void* A, long long batch_stride, int lda, double* work, const int* mask,
                             int* info, admm_stream_t stream) {
  ADMM_REQUIRE(n > 0 && n <= 2048 && nbatch >= 0 && lda >= n && work != nullptr, ADMM_EINVAL,
               "admm_hpd_inverse_batched: bad dims");
  if (nbatch == 0) return ADMM_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long tot = (long long)nbatch * n * n;
  hpd_embed_kernel<<<ew_grid(tot), 256, 0, s>>>(n, nbatch, static_cast<const cplx*>(A), batch_stride, lda, work, mask);
  if (int rc = check_launch("admm_hpd_inverse_batched(embed)")) return rc;
  if (int rc = admm_spd_inverse_batched(2 * n, nbatch, work, 4LL * n * n, 2 * n, mask, info, stream)) return rc;
  hpd_extract_kernel<<<ew_grid(tot), 256, 0, s>>>(n, nbatch, work, static_cast<cplx*>(A), batch_stride, lda, mask, info);
  return check_launch("admm_hpd_inverse_batched(extract)");
}

}  // extern "C"
