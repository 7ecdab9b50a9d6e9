// Pattern A engine (basis pursuit / LASSO): every problem has its own real A (M x N).
// One CTA per problem keeps x0, x1, h, A^T y and the work vectors in shared memory and runs all
// iterations of SimpleOptimizer.solve (optimizer.py:302-341) without returning to the host; the
// x-update uses a cached inverse (Woodbury M x M when M < N, direct N x N otherwise).
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <map>

namespace cg = cooperative_groups;

namespace admm {

// ---------------------------------------------------------------------------------------------
// setup: aty = alpha A^T y ; gram = A A^T (woodbury) or A^T A
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bp_aty_kernel(admm_bp_buffers b, const double* __restrict__ y, double* __restrict__ aty) {
  const int prob = blockIdx.y;
  const double* A = b.A + (size_t)prob * b.M * b.N;
  const double* yv = y + (size_t)prob * b.M;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < b.N; n += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int m = 0; m < b.M; ++m) acc += A[(size_t)m * b.N + n] * yv[m];
    aty[(size_t)prob * b.N + n] = b.alpha * acc;
  }
}

// C = A A^T (WOOD) : C[i][j] = sum_n A[i][n] A[j][n]   (nk = M, inner = N)
// C = A^T A        : C[i][j] = sum_m A[m][i] A[m][j]   (nk = N, inner = M)
template <bool WOOD>
__global__ void __launch_bounds__(256) bp_gram_kernel(admm_bp_buffers b, double* __restrict__ gram) {
  __shared__ double As[32][33];
  __shared__ double Bs[32][33];
  const int prob = blockIdx.z;
  const double* A = b.A + (size_t)prob * b.M * b.N;
  const int nk = b.nk, inner = WOOD ? b.N : b.M;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  if (col0 > row0) return;  // lower triangle tiles only; mirrored on store
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < inner; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      double a = 0.0, c = 0.0;
      if (WOOD) {
        if (row0 + r < nk && k0 + tx < inner) a = A[(size_t)(row0 + r) * b.N + k0 + tx];
        if (col0 + r < nk && k0 + tx < inner) c = A[(size_t)(col0 + r) * b.N + k0 + tx];
        As[r][tx] = a;   // As[i][kk]
        Bs[r][tx] = c;   // Bs[j][kk]
      } else {
        if (k0 + r < inner && row0 + tx < nk) a = A[(size_t)(k0 + r) * b.N + row0 + tx];
        if (k0 + r < inner && col0 + tx < nk) c = A[(size_t)(k0 + r) * b.N + col0 + tx];
        As[tx][r] = a;
        Bs[tx][r] = c;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double c = Bs[tx][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += As[ty + 8 * i][kk] * c;
    }
    __syncthreads();
  }
  double* G = gram + (size_t)prob * nk * nk;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty + 8 * i, c = col0 + tx;
    if (r < nk && c < nk) {
      G[(size_t)r * nk + c] = acc[i];
      G[(size_t)c * nk + r] = acc[i];
    }
  }
}

// Kinv <- gram shifted by the current mu, for flagged problems (then inverted in place)
__global__ void __launch_bounds__(256) bp_shift_kernel(admm_bp_buffers b) {
  const int prob = blockIdx.y;
  if (!b.need_factor[prob]) return;
  const int nk = b.nk;
  const double mu = b.mu[prob];
  const double* G = b.gram + (size_t)prob * nk * nk;
  double* K = b.Kinv + (size_t)prob * nk * nk;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nk * nk; idx += gridDim.x * blockDim.x) {
    const int i = idx / nk, j = idx - i * nk;
    double v = G[idx];
    if (b.woodbury) {
      if (i == j) v += mu / b.alpha;
    } else {
      v = b.alpha * v + (i == j ? mu : 0.0);
    }
    K[idx] = v;
  }
}

__global__ void bp_clear_flag_kernel(admm_bp_buffers b) {
  const int prob = blockIdx.x * blockDim.x + threadIdx.x;
  if (prob < b.nb) b.need_factor[prob] = 0;
}

// ---------------------------------------------------------------------------------------------
// persistent per-problem iteration kernel
// ---------------------------------------------------------------------------------------------
constexpr int BP_THREADS = 256;

__global__ void __launch_bounds__(BP_THREADS) bp_iterate_kernel(admm_bp_buffers b, int iter_end) {
  extern __shared__ double sm[];
  const int prob = blockIdx.x;
  if (b.done[prob] || b.need_factor[prob]) return;
  int it = b.iters[prob];
  if (it >= iter_end) return;
  const int M = b.M, N = b.N, nk = b.nk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BP_THREADS / 32;
  double* aty = sm;          // N
  double* x0 = aty + N;      // N
  double* xo = x0 + N;       // N
  double* x1 = xo + N;       // N
  double* h = x1 + N;        // N
  double* r = h + N;         // N
  double* tv = r + N;        // nk
  double* sv = tv + nk;      // nk
  double* scratch = sv + nk; // 5*32
  const double* A = b.A + (size_t)prob * M * N;
  const double* Kinv = b.Kinv + (size_t)prob * nk * nk;
  for (int n = tid; n < N; n += BP_THREADS) {
    aty[n] = b.aty[(size_t)prob * N + n];
    x0[n] = b.x0[(size_t)prob * N + n];
    x1[n] = b.x1[(size_t)prob * N + n];
    h[n] = b.h[(size_t)prob * N + n];
  }
  double mu = b.mu[prob];
  int done = 0, need = 0;
  double primal = 0.0, dual = 0.0;
  __syncthreads();

  while (it < iter_end) {
    for (int n = tid; n < N; n += BP_THREADS) {
      xo[n] = x0[n];
      r[n] = aty[n] + h[n] + mu * x1[n];
    }
    __syncthreads();
    if (b.woodbury) {
      // t = A r : one warp per row, lanes stride the row (coalesced)
      for (int m = warp; m < M; m += NW) {
        const double* row = A + (size_t)m * N;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int n = lane;
        for (; n + 96 < N; n += 128) {
          a0 += row[n] * r[n];
          a1 += row[n + 32] * r[n + 32];
          a2 += row[n + 64] * r[n + 64];
          a3 += row[n + 96] * r[n + 96];
        }
        for (; n < N; n += 32) a0 += row[n] * r[n];
        const double s = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) tv[m] = s;
      }
      __syncthreads();
      // s = Kinv t
      for (int m = warp; m < M; m += NW) {
        const double* row = Kinv + (size_t)m * M;
        double a0 = 0.0;
        for (int j = lane; j < M; j += 32) a0 += row[j] * tv[j];
        a0 = warp_sum(a0);
        if (lane == 0) sv[m] = a0;
      }
      __syncthreads();
      // x0 = (r - A^T s) / mu : one thread per column (coalesced across threads)
      for (int n = tid; n < N; n += BP_THREADS) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int m = 0;
        for (; m + 3 < M; m += 4) {
          a0 += A[(size_t)m * N + n] * sv[m];
          a1 += A[(size_t)(m + 1) * N + n] * sv[m + 1];
          a2 += A[(size_t)(m + 2) * N + n] * sv[m + 2];
          a3 += A[(size_t)(m + 3) * N + n] * sv[m + 3];
        }
        for (; m < M; ++m) a0 += A[(size_t)m * N + n] * sv[m];
        x0[n] = (r[n] - ((a0 + a1) + (a2 + a3))) / mu;
      }
    } else {
      // x0 = Ginv r  (N x N)
      for (int n = warp; n < N; n += NW) {
        const double* row = Kinv + (size_t)n * N;
        double a0 = 0.0;
        for (int j = lane; j < N; j += 32) a0 += row[j] * r[j];
        a0 = warp_sum(a0);
        if (lane == 0) x0[n] = a0;
      }
    }
    __syncthreads();
    // z-update (soft threshold), dual ascent, norms
    const double thr = 0.5 * b.lam / mu;
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int n = tid; n < N; n += BP_THREADS) {
      const double xv = x0[n], hv = h[n], xov = xo[n];
      const double yv = -((hv - mu * xv) / mu);
      double z = 0.0;
      if (yv > thr) z = yv - thr;
      if (yv < -thr) z = yv + thr;
      x1[n] = z;
      h[n] = hv + mu * (z - xv);
      v[0] += (xv - z) * (xv - z);
      v[1] += xv * xv;
      v[2] += z * z;
      v[3] += (xv - xov) * (xv - xov);
      v[4] += xov * xov;
    }
    block_sum<5>(v, scratch);
    const double p = sqrt(v[0]), nx0 = sqrt(v[1]), nx1 = sqrt(v[2]), nd = sqrt(v[3]), nxo = sqrt(v[4]);
    primal = p;
    dual = mu * nd;
    if (b.history && it < b.hist_cap && tid == 0) {
      b.history[((size_t)prob * b.hist_cap + it) * 2] = primal;
      b.history[((size_t)prob * b.hist_cap + it) * 2 + 1] = dual;
    }
    const int this_it = it;
    ++it;
    const bool conv = (p / fmax(nx0, nx1) < b.rtol) && (dual / fmax(mu * nx0, mu * nxo) < b.rtol);
    if (conv) {
      done = 1;
      break;
    }
    if (this_it % b.interval_update_mu == 0) {
      double m2 = mu;
      if (primal > b.th_change * dual) m2 *= b.fact_incr;
      if (dual > b.th_change * primal) m2 /= b.fact_incr;
      m2 = fmin(m2, b.max_mu);
      if (m2 != mu) {
        mu = m2;
        need = 1;
        break;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int n = tid; n < N; n += BP_THREADS) {
    b.x0[(size_t)prob * N + n] = x0[n];
    b.x1[(size_t)prob * N + n] = x1[n];
    b.h[(size_t)prob * N + n] = h[n];
    if (b.x0_old != nullptr) b.x0_old[(size_t)prob * N + n] = xo[n];     // `_x_old[0]` (optimizer.py:324)
  }
  if (tid == 0) {
    b.mu[prob] = mu;
    b.iters[prob] = it;
    b.done[prob] = done;
    b.need_factor[prob] = need;
    b.last_res[2 * prob] = primal;
    b.last_res[2 * prob + 1] = dual;
  }
}

// ---------------------------------------------------------------------------------------------
// fused single-sweep iteration (Woodbury path): A is streamed ONCE per iteration
// ---------------------------------------------------------------------------------------------
// One iteration needs  t = A r  (by rows) and  A^T s  (by columns).  The z-update, the dual ascent
// and the next right-hand side r' = alpha A^T y + h' + mu x1' are elementwise in the column index,
// so while a column tile of A is on chip for A^T s, the same tile immediately contributes
// A[:, tile] r'[tile] to the NEXT iteration's t: one pass over A per iteration instead of two
// (8 M N + 8 M^2 bytes per problem-iteration instead of 16 M N + 8 M^2).
//
// Column tiles (M x 32 doubles) arrive in a 2-stage shared-memory ring by per-row TMA bulk copies
// (one 256-byte copy per row, issued by M threads, completion on an mbarrier); the stream is
// continuous across iterations.  Warp w owns rows [16 w, 16 w + 16): it computes s for exactly those
// rows (K^-1 t, rows of K^-1 from L2/HBM), keeps its 16 x 32 patch of the tile in registers
// (one column per lane), and accumulates both products from it.
// A (row-major M x N per problem) -> At[prob][tile][m][32], zero padded: every 32-column tile is one
// contiguous block (one TMA bulk copy, full DRAM pages) instead of M strided 256-byte row segments.
__global__ void __launch_bounds__(256) bp_tile_A_kernel(admm_bp_buffers b, double* __restrict__ At) {
  const int prob = blockIdx.y;
  const int T = (b.N + 31) / 32;
  const double* A = b.A + (size_t)prob * b.M * b.N;
  double* dst = At + (size_t)prob * T * b.M * 32;
  const long long total = (long long)T * b.M * 32;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = int(idx & 31);
    const long long q = idx >> 5;
    const int m = int(q % b.M), k = int(q / b.M);
    const int n = 32 * k + c;
    dst[idx] = n < b.N ? A[(size_t)m * b.N + n] : 0.0;
  }
}

constexpr int BPF_TN = 32;       // columns per tile (one per lane)
constexpr int BPF_RW = 16;       // rows per warp
constexpr int BPF_STAGES = 2;

template <int NW>
struct BpfSmem {
  static constexpr int ROWS = NW * BPF_RW;
  static constexpr int STAGE_D = ROWS * BPF_TN;
  static size_t bytes(int N) {
    const int Np = (N + 1) & ~1;
    return (size_t)(BPF_STAGES * STAGE_D + 5 * Np + 2 * ROWS + NW * BPF_TN + BPF_TN + 5 * 32 + 2 * (ROWS + 8)) * sizeof(double) +
           BPF_STAGES * sizeof(uint64_t) + 16;
  }
};

// CS > 1: ONE problem is spread over a thread-block cluster of CS CTAs (few problems, many idle SMs):
// CTA c streams the column tiles k = c, c+CS, ... and owns those columns of x0 / x1 / h; per iteration
// the partial t = A r' vectors and the five norm partials are exchanged through distributed shared
// memory (each CTA sums the CS partial buffers of its peers in fixed rank order, so every CTA holds
// bit-identical t and norms and takes the same decisions) with one cluster barrier.
template <int NW, int CS>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) bp_fused_kernel(admm_bp_buffers b, int iter_end) {
  extern __shared__ __align__(128) double sm[];
  constexpr int ROWS = BpfSmem<NW>::ROWS, STAGE_D = BpfSmem<NW>::STAGE_D, NT_ = NW * 32;
  const int prob = CS > 1 ? blockIdx.x / CS : blockIdx.x;
  const int crank = CS > 1 ? (int)(blockIdx.x % CS) : 0;
  if (b.done[prob] || b.need_factor[prob]) return;
  int it = b.iters[prob];
  if (it >= iter_end) return;
  const int M = b.M, N = b.N, Np = (N + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* stage = sm;                              // [STAGES][ROWS][TN]
  double* aty = stage + BPF_STAGES * STAGE_D;      // N
  double* x0 = aty + Np;                           // N
  double* x1 = x0 + Np;                            // N
  double* h = x1 + Np;                             // N
  double* r = h + Np;                              // N   current right-hand side
  double* tv = r + Np;                             // ROWS   t = A r
  double* sv = tv + ROWS;                          // ROWS   s = K^-1 t
  double* part = sv + ROWS;                        // NW x TN  per-warp partial column dots
  double* rn = part + NW * BPF_TN;                 // TN   r' of the current tile
  double* scratch = rn + BPF_TN;                   // 5*32
  double* xch = scratch + 5 * 32;                  // [2][ROWS + 8] partial t and norms for the cluster exchange
  uint64_t* full = reinterpret_cast<uint64_t*>(xch + 2 * (ROWS + 8));
  const double* Kinv = b.Kinv + (size_t)prob * M * M;
  const int T = (N + BPF_TN - 1) / BPF_TN;          // tiles per sweep
  const int Tl = CS > 1 ? (T - crank + CS - 1) / CS : T;      // tiles of this CTA: k = crank + j*CS
  const double* At = b.At + (size_t)prob * T * M * BPF_TN;     // tile-major copy of A (bp_tile_A_kernel)
  const unsigned TILE_BYTES = (unsigned)(M * BPF_TN * sizeof(double));

  // rows >= M and the columns of a ragged last tile are never written by a copy: keep them finite
  for (int i = tid; i < BPF_STAGES * STAGE_D; i += NT_) stage[i] = 0.0;
  for (int i = tid; i < 2 * ROWS; i += NT_) tv[i] = 0.0;      // tv and sv
  for (int n = tid; n < N; n += NT_) {
    aty[n] = b.aty[(size_t)prob * N + n];
    x0[n] = b.x0[(size_t)prob * N + n];
    x1[n] = b.x1[(size_t)prob * N + n];
    h[n] = b.h[(size_t)prob * N + n];
  }
  if (tid == 0) {
    for (int s = 0; s < BPF_STAGES; ++s) mbar_init(full + s, 1);
    fence_barrier_init();
  }
  double mu = b.mu[prob];
  __syncthreads();
  for (int n = tid; n < N; n += NT_) r[n] = aty[n] + h[n] + mu * x1[n];

  // tile `gt` of the endless stream (sweep after sweep) -> stage gt % STAGES
  auto issue = [&](int gt) {
    if (tid == 0 && Tl > 0) {
      const int k = crank + (gt % Tl) * CS, sidx = gt % BPF_STAGES;
      fence_proxy_async();
      mbar_expect_tx(full + sidx, TILE_BYTES);
      tma_bulk_g2s(stage + sidx * STAGE_D, At + (size_t)k * M * BPF_TN, TILE_BYTES, full + sidx);
    }
  };
  __syncthreads();
  int gt = 0;                 // next tile to consume; tiles gt .. gt+STAGES-1 are in flight
  for (int s = 0; s < BPF_STAGES; ++s) issue(s);

  int done = 0, need = 0;
  double primal = 0.0, dual = 0.0;
  bool first = true;          // first sweep of this launch: only t = A r
  int xphase = 0;             // ping-pong index of the cluster exchange buffer
  while (true) {
    // ---- s = K^-1 t for this warp's rows, kept in shared memory (only this warp reads them back).
    // The loads of 8 rows are issued together: two L2 round trips per iteration instead of sixteen.
    if (!first) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double kv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = warp * BPF_RW + half * 8 + i;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = lane + 32 * q;
            kv[i][q] = (m < M && j < M) ? __ldg(Kinv + (size_t)m * M + j) : 0.0;
          }
        }
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) a[i] += kv[i][q] * tv[lane + 32 * q];
        }
        for (int j0 = 128; j0 < M; j0 += 32) {        // M > 128 (16-warp variant): remaining columns
          const int j = j0 + lane;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = warp * BPF_RW + half * 8 + i;
            if (m < M && j < M) a[i] += __ldg(Kinv + (size_t)m * M + j) * tv[j];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double sres = warp_sum(a[i]);
          if (lane == 0) sv[warp * BPF_RW + half * 8 + i] = sres;
        }
      }
      __syncthreads();        // everyone has read tv before it is overwritten below; sv visible
    }
    double acc[BPF_RW];
#pragma unroll
    for (int i = 0; i < BPF_RW; ++i) acc[i] = 0.0;
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const double thr = 0.5 * b.lam / mu, inv_mu = 1.0 / mu;
    // K^-1 is needed again right after this sweep: pull it into L2 meanwhile (16 KB per bulk prefetch)
    {
      const size_t kbytes = (size_t)M * M * sizeof(double);
      const size_t off = (size_t)tid * 16384;
      if (off < kbytes && (reinterpret_cast<uintptr_t>(Kinv) & 15) == 0) {
        const unsigned nbytes = (unsigned)min((size_t)16384, (kbytes - off) & ~(size_t)15);
        if (nbytes) l2_prefetch_bulk(reinterpret_cast<const char*>(Kinv) + off, nbytes);
      }
    }

    for (int kl = 0; kl < Tl; ++kl, ++gt) {
      const int k = crank + kl * CS;
      const int sidx = gt % BPF_STAGES;
      mbar_wait(full + sidx, (unsigned)((gt / BPF_STAGES) & 1));
      const double* tile = stage + sidx * STAGE_D + (warp * BPF_RW) * BPF_TN + lane;
      double patch[BPF_RW];
#pragma unroll
      for (int i = 0; i < BPF_RW; ++i) patch[i] = tile[i * BPF_TN];
      if (!first) {
        const double* sw = sv + warp * BPF_RW;
        double pc0 = 0.0, pc1 = 0.0;
#pragma unroll
        for (int i = 0; i < BPF_RW; i += 2) {
          const double2 s2 = *reinterpret_cast<const double2*>(sw + i);     // broadcast
          pc0 += patch[i] * s2.x;
          pc1 += patch[i + 1] * s2.y;
        }
        part[warp * BPF_TN + lane] = pc0 + pc1;
      }
      __syncthreads();                     // stage consumed by everyone; partial dots visible
      issue(gt + BPF_STAGES);
      double rnv;
      if (!first) {
        if (tid < BPF_TN) {
          const int n = k * BPF_TN + tid;
          double rnew = 0.0;
          if (n < N) {
            double cw[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) cw[w] = part[w * BPF_TN + tid];
#pragma unroll
            for (int st = NW / 2; st > 0; st >>= 1)
#pragma unroll
              for (int w = 0; w < st; ++w) cw[w] += cw[w + st];
            const double c = cw[0];
            // x-update by the Woodbury identity, then z-update / dual ascent as in bp_iterate_kernel
            // (this is the serial section of a tile: multiply by 1/mu instead of dividing twice)
            const double xov = x0[n], hv = h[n];
            if (b.x0_old != nullptr) b.x0_old[(size_t)prob * N + n] = xov;     // `_x_old[0]` (optimizer.py:324)
            const double xv = (r[n] - c) * inv_mu;
            const double yv = xv - hv * inv_mu;
            double z = 0.0;
            if (yv > thr) z = yv - thr;
            if (yv < -thr) z = yv + thr;
            const double hn = hv + mu * (z - xv);
            x0[n] = xv;
            x1[n] = z;
            h[n] = hn;
            rnew = aty[n] + hn + mu * z;
            r[n] = rnew;
            v[0] += (xv - z) * (xv - z);
            v[1] += xv * xv;
            v[2] += z * z;
            v[3] += (xv - xov) * (xv - xov);
            v[4] += xov * xov;
          }
          rn[tid] = rnew;
        }
        __syncthreads();
        rnv = rn[lane];
      } else {
        const int n = k * BPF_TN + lane;
        rnv = n < N ? r[n] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < BPF_RW; ++i) acc[i] += patch[i] * rnv;
    }
    // ---- t' = A r' : rows are owned by warps, columns were spread over lanes and tiles
#pragma unroll
    for (int i = 0; i < BPF_RW; ++i) acc[i] = warp_sum(acc[i]);
    if (CS == 1) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < BPF_RW; ++i) tv[warp * BPF_RW + i] = acc[i];
      }
      if (first) {
        first = false;
        __syncthreads();
        continue;
      }
      block_sum<5>(v, scratch);              // (syncs: tv is visible afterwards)
    } else {
      // cluster exchange: my partial t and norm partials -> xch[xphase]; everyone sums all CS buffers
      double* mine = xch + xphase * (ROWS + 8);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < BPF_RW; ++i) mine[warp * BPF_RW + i] = acc[i];
      }
      block_sum<5>(v, scratch);
      if (tid < 5) mine[ROWS + tid] = v[tid];
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();
      for (int i = tid; i < ROWS + 5; i += NT_) {
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < CS; ++c) a += cluster.map_shared_rank(mine, c)[i];
        if (i < ROWS) tv[i] = a;
        else scratch[i - ROWS] = a;
      }
      xphase ^= 1;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 5; ++i) v[i] = scratch[i];
      if (first) {
        first = false;
        continue;
      }
    }
    const double p = sqrt(v[0]), nx0 = sqrt(v[1]), nx1 = sqrt(v[2]), nd = sqrt(v[3]), nxo = sqrt(v[4]);
    primal = p;
    dual = mu * nd;
    if (b.history && it < b.hist_cap && tid == 0 && crank == 0) {
      b.history[((size_t)prob * b.hist_cap + it) * 2] = primal;
      b.history[((size_t)prob * b.hist_cap + it) * 2 + 1] = dual;
    }
    const int this_it = it;
    ++it;
    const bool conv = (p / fmax(nx0, nx1) < b.rtol) && (dual / fmax(mu * nx0, mu * nxo) < b.rtol);
    if (conv) {
      done = 1;
      break;
    }
    if (this_it % b.interval_update_mu == 0) {
      double m2 = mu;
      if (primal > b.th_change * dual) m2 *= b.fact_incr;
      if (dual > b.th_change * primal) m2 /= b.fact_incr;
      m2 = fmin(m2, b.max_mu);
      if (m2 != mu) {
        mu = m2;
        need = 1;
        break;
      }
    }
    if (it >= iter_end) break;
    __syncthreads();
  }
  // drain the tiles still in flight before the shared memory goes away
  if (Tl > 0)
    for (int g2 = gt; g2 < gt + BPF_STAGES; ++g2) mbar_wait(full + g2 % BPF_STAGES, (unsigned)((g2 / BPF_STAGES) & 1));
  __syncthreads();
  if (CS > 1) cg::this_cluster().sync();     // nobody leaves while a peer may still read its exchange buffer
  for (int n = tid; n < N; n += NT_) {
    if (CS > 1 && ((n / BPF_TN) % CS) != crank) continue;      // every CTA writes the columns it owns
    b.x0[(size_t)prob * N + n] = x0[n];
    b.x1[(size_t)prob * N + n] = x1[n];
    b.h[(size_t)prob * N + n] = h[n];
  }
  if (tid == 0 && crank == 0) {
    b.mu[prob] = mu;
    b.iters[prob] = it;
    b.done[prob] = done;
    b.need_factor[prob] = need;
    b.last_res[2 * prob] = primal;
    b.last_res[2 * prob + 1] = dual;
  }
}

// ---------------------------------------------------------------------------------------------
// solo: ONE problem (or a handful) resident in the shared memory of a thread-block cluster
// ---------------------------------------------------------------------------------------------
// The notebook / test instance of the reference (basis_pursuit.ipynb, test_optimizer.py:52-82: one
// 100..200 x 1000 problem) is pure latency for the batch kernels: A has to be streamed from L2 every
// iteration by a few CTAs.  Here a cluster of CS CTAs keeps the problem on chip for the whole solve:
//   * CTA c owns N/CS columns: its slice of A (column-major; pitch even with pitch/2 odd: 16-byte loads, conflict-free in both
//     products) stays in shared memory, x0 / x1 / h / alpha A^T y / r of a column in the registers of
//     one thread; it also owns M/CS rows of K^-1 (shared memory);
//   * per iteration  c = A^T s  (own columns), x-update + soft threshold + dual ascent (registers),
//     partial t' = A r' (all M rows, own columns), REDUCE-SCATTER of t' (CTA c sums rows c M/CS ...) with the
//     five norm partials riding along, ALL-GATHER of t', the own rows of s = K^-1 t', ALL-GATHER of s;
//   * the exchanges are DSMEM pushes (st.async + complete_tx on the receiver's mbarrier, ping-pong
//     buffers armed one use ahead): no cluster barrier, no fence in the loop.  Every CTA sums the
//     partials in rank order, so all CTAs hold bit-identical t, s and norms and take identical decisions.
constexpr int BPS_THREADS = 512;
constexpr int BPS_PARTS = 8;       // row parts of the A^T s product

struct BpsLayout {      // offsets in doubles
  int Nc, Nce, Ncp, Mr, Mp, Mq, Cq, pitch, XW, nparts3, A, K, s, t, rs, cp, tp, nrm, nrmt, x1, bars, total;
};
__host__ __device__ inline BpsLayout bps_layout(int M, int N, int cs) {
  BpsLayout o;
  o.Nc = (N + cs - 1) / cs;             // columns per CTA
  o.Ncp = (o.Nc + 31) & ~31;
  o.Mr = (M + cs - 1) / cs;             // rows of K^-1 per CTA
  o.Mp = (M + 31) & ~31;
  o.Mq = (((M + BPS_PARTS - 1) / BPS_PARTS) + 1) & ~1;      // rows per part of A^T s, even: 16-byte loads
  // pitch: even (16-byte loads along a column) with pitch/2 odd (the 16-byte slots of 8 neighbouring columns fall
  // into 8 different bank groups) and >= the zero-padded BPS_PARTS * Mq rows
  o.pitch = BPS_PARTS * o.Mq > M ? BPS_PARTS * o.Mq : ((M + 1) & ~1);
  if (o.pitch % 4 == 0) o.pitch += 2;
  o.Nce = (o.Nc + 1) & ~1;                                   // columns of the slice, zero padded to even
  if (o.Mp < BPS_PARTS * o.Mq) o.Mp = (BPS_PARTS * o.Mq + 31) & ~31;
  o.XW = (o.Mr + 5 + 1) & ~1;           // slot pitch of the reduce-scatter: my Mr rows of the partial t' + 5 norm partials
  o.nparts3 = BPS_THREADS / o.Mp > 0 ? BPS_THREADS / o.Mp : 1;
  o.Cq = (((o.Nc + o.nparts3 - 1) / o.nparts3) + 1) & ~1;    // columns per part of A r', even
  int at = 0;
  auto take = [&](int n) { const int r = at; at += (n + 1) & ~1; return r; };
  o.A = take(o.Nce * o.pitch);
  o.K = take(o.Mr * M);
  o.s = take(2 * o.Mp);                 // all-gather receive buffers (ping-pong)
  o.t = take(o.Mp);
  o.rs = take(o.Ncp);
  o.cp = take(BPS_PARTS * o.Ncp);
  o.tp = take(o.nparts3 * o.Mp);
  o.nrm = take((BPS_THREADS / 32) * 5);
  o.nrmt = take(6);
  o.x1 = take(2 * cs * o.XW);           // reduce-scatter receive slots (ping-pong)
  o.bars = take(6);
  o.total = at;
  return o;
}

#ifdef SOLO_TRACE   // tools only: clock64 stamps of warp 0 of CTA 0 for iterations 10..13 into b.last_res + 16
#define BPS_STAMP(idx)                                                                        \
  if (crank == 0 && tid == 0 && it >= 3000 && it < 3004 && !first)                                \
    reinterpret_cast<long long*>(b.history)[4096 + (it - 3000) * 16 + (idx)] = clock64();
#else
#define BPS_STAMP(idx)
#endif

template <int CS>
__global__ void __launch_bounds__(BPS_THREADS, 1) bp_solo_kernel(admm_bp_buffers b, int iter_end) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int prob = blockIdx.x / CS;
  if (b.done[prob] || b.need_factor[prob]) return;          // uniform over the cluster
  int it = b.iters[prob];
  if (it >= iter_end) return;
  constexpr int NW = BPS_THREADS / 32;
  const int M = b.M, N = b.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(128) double sm[];
  const BpsLayout lay = bps_layout(M, N, CS);
  const int Nc = lay.Nc, Ncp = lay.Ncp, Mr = lay.Mr, Mp = lay.Mp, pitch = lay.pitch, XW = lay.XW;
  double* As = sm + lay.A;          // [nl][pitch]: A[m][n0 + nl]
  double* Ks = sm + lay.K;          // [rl][M]: K^-1[m0 + rl][:]
  double* sbuf = sm + lay.s;        // [2][Mp]
  double* tfull = sm + lay.t;       // [Mp]
  double* rs = sm + lay.rs;         // [Ncp]: r of my columns
  double* cp = sm + lay.cp;         // [part][Ncp]
  double* tp = sm + lay.tp;         // [part3][Mp]
  double* nrm = sm + lay.nrm;       // [warp][5]
  double* nrmt = sm + lay.nrmt;     // [5] cluster totals
  double* x1s = sm + lay.x1;        // [2][CS][XW]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + lay.bars);   // 0,1: reduce-scatter; 2,3: all-gather of s; 4: all-gather of t
  const int n0 = crank * Nc, ncol = max(0, min(Nc, N - n0));
  const int m0 = crank * Mr, nrowK = max(0, min(Mr, M - m0));
  const double* A = b.A + (size_t)prob * M * N;
  const double* Kinv = b.Kinv + (size_t)prob * M * M;

  for (int idx = tid; idx < lay.Nce * pitch; idx += BPS_THREADS) As[idx] = 0.0;
  for (int i = tid; i < 2 * Mp; i += BPS_THREADS) sbuf[i] = 0.0;
  for (int i = tid; i < Mp; i += BPS_THREADS) tfull[i] = 0.0;
  for (int i = tid; i < Ncp; i += BPS_THREADS) rs[i] = 0.0;
  for (int i = tid; i < BPS_PARTS * Ncp; i += BPS_THREADS) cp[i] = 0.0;
  for (int i = tid; i < NW * 5; i += BPS_THREADS) nrm[i] = 0.0;
  __syncthreads();
  for (int idx = tid; idx < M * Nc; idx += BPS_THREADS) {
    const int m = idx / Nc, nl = idx - m * Nc;
    if (nl < ncol) As[nl * pitch + m] = A[(size_t)m * N + n0 + nl];
  }
  for (int idx = tid; idx < nrowK * M; idx += BPS_THREADS) Ks[idx] = Kinv[(size_t)m0 * M + idx];
  double mu = b.mu[prob];
  // my column (thread tid < ncol): everything of size N lives in registers
  const bool own = tid < ncol;
  double c_aty = 0.0, c_x0 = 0.0, c_x1 = 0.0, c_h = 0.0, c_r = 0.0;
  double c_xold = 0.0;       // x0 at the start of the last executed iteration (`_x_old[0]`, optimizer.py:324)
  bool ran_any = false;
  if (own) {
    const size_t o = (size_t)prob * N + n0 + tid;
    c_aty = b.aty[o];
    c_x0 = b.x0[o];
    c_x1 = b.x1[o];
    c_h = b.h[o];
    c_r = c_aty + c_h + mu * c_x1;
    rs[tid] = c_r;
  }
  const unsigned X1BYTES = (unsigned)(CS * (nrowK + 5) * sizeof(double)), X2BYTES = (unsigned)(M * sizeof(double));
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
    fence_barrier_init();
    mbar_expect_tx(bars + 0, X1BYTES);
    mbar_expect_tx(bars + 1, X1BYTES);
    mbar_expect_tx(bars + 2, X2BYTES);
    mbar_expect_tx(bars + 3, X2BYTES);
    mbar_expect_tx(bars + 4, X2BYTES);
  }
  cluster.sync();                  // barriers armed everywhere before anybody pushes; shared data visible
  const unsigned x1_u32 = smem_u32(x1s), s_u32 = smem_u32(sbuf), t_u32 = smem_u32(tfull), bar_u32 = smem_u32(bars);
  unsigned par = 0u;               // bit i: parity of the next completion of barrier i
  int ph1 = 0, ph2 = 0;            // buffers of the NEXT all-reduce / all-gather
  const double rtol2 = b.rtol * b.rtol;
  double thr = 0.5 * b.lam / mu, inv_mu = 1.0 / mu;
  int done = 0, need = 0;
  double sq_p = 0.0, sq_d = 0.0, mu_res = mu;
  bool first = true;               // first pass of this launch: only t = A r, s = K^-1 t
  const int Mq = lay.Mq, Cq = lay.Cq;

  while (true) {
    BPS_STAMP(0)
    if (!first) {
      // ---- c = A^T s for my columns: warp = (row part, block of 32 columns)
      const double* sf = sbuf + (ph2 ^ 1) * Mp;             // the s received last
      const int part = warp & (BPS_PARTS - 1);
      const int mlo = part * Mq, mhi = mlo + Mq;            // rows >= M are zero in As and in s
      for (int blk = warp / BPS_PARTS; blk * 32 < Nc; blk += NW / BPS_PARTS) {
        const int nl = blk * 32 + lane;
        const double* ac = As + min(nl, Nc - 1) * pitch;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int m = mlo;
        for (; m + 3 < mhi; m += 4) {
          const double2 p01 = *reinterpret_cast<const double2*>(ac + m), p23 = *reinterpret_cast<const double2*>(ac + m + 2);
          const double2 s01 = *reinterpret_cast<const double2*>(sf + m), s23 = *reinterpret_cast<const double2*>(sf + m + 2);
          a0 += p01.x * s01.x;
          a1 += p01.y * s01.y;
          a2 += p23.x * s23.x;
          a3 += p23.y * s23.y;
        }
        if (m < mhi) {                                      // Mq is even: one 16-byte pair left
          const double2 p01 = *reinterpret_cast<const double2*>(ac + m), s01 = *reinterpret_cast<const double2*>(sf + m);
          a0 += p01.x * s01.x;
          a1 += p01.y * s01.y;
        }
        if (nl < Nc) cp[part * Ncp + nl] = (a0 + a1) + (a2 + a3);
      }
      BPS_STAMP(1)
      __syncthreads();
      BPS_STAMP(2)
      // everybody has read this s buffer: re-arm its barrier for the all-gather after the next
      if (tid == 0) {
        mbar_expect_tx(bars + 2 + (ph2 ^ 1), X2BYTES);
        mbar_expect_tx(bars + 4, X2BYTES);      // t was consumed by the K^-1 rows of the previous iteration
      }
      // ---- x-update by the Woodbury identity, soft threshold, dual ascent (optimizer.py:322-341)
      double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (own) {
        double c = 0.0;
#pragma unroll
        for (int q = 0; q < BPS_PARTS; ++q) c += cp[q * Ncp + tid];
        const double xov = c_x0;
        c_xold = xov;
        ran_any = true;
        const double xv = (c_r - c) * inv_mu;
        const double yv = xv - c_h * inv_mu;
        double z = 0.0;
        if (yv > thr) z = yv - thr;
        if (yv < -thr) z = yv + thr;
        c_h = c_h + mu * (z - xv);
        c_x0 = xv;
        c_x1 = z;
        c_r = c_aty + c_h + mu * z;
        rs[tid] = c_r;
        v[0] = (xv - z) * (xv - z);
        v[1] = xv * xv;
        v[2] = z * z;
        v[3] = (xv - xov) * (xv - xov);
        v[4] = xov * xov;
      }
      if (warp * 32 < Nc) {
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = warp_sum(v[i]);
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 5; ++i) nrm[warp * 5 + i] = v[i];
        }
      }
      BPS_STAMP(3)
      __syncthreads();
      BPS_STAMP(4)
    }
    // ---- partial t' = A r' over my columns: thread = (column part, row)
    {
      const int q = tid / Mp, m = tid - q * Mp;
      if (q < lay.nparts3 && m < M) {
        const int clo = q * Cq, chi = min(Nc, clo + Cq);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double* ar = As + m;
        int nl = clo;                                       // clo is even, rs is zero padded
        for (; nl + 3 < chi; nl += 4) {
          const double2 r01 = *reinterpret_cast<const double2*>(rs + nl), r23 = *reinterpret_cast<const double2*>(rs + nl + 2);
          a0 += ar[nl * pitch] * r01.x;
          a1 += ar[(nl + 1) * pitch] * r01.y;
          a2 += ar[(nl + 2) * pitch] * r23.x;
          a3 += ar[(nl + 3) * pitch] * r23.y;
        }
        for (; nl < chi; ++nl) a0 += ar[nl * pitch] * rs[nl];
        tp[q * Mp + m] = (a0 + a1) + (a2 + a3);
      }
    }
    BPS_STAMP(5)
    __syncthreads();
    BPS_STAMP(6)
    // ---- reduce-scatter (push): row m of the partial t' goes to the CTA that owns it; the five norm partials to all
    if (tid < M + 5) {
      const unsigned bar = bar_u32 + (unsigned)(ph1 * sizeof(uint64_t));
      if (tid < M) {
        double a = 0.0;
        for (int q = 0; q < lay.nparts3; ++q) a += tp[q * Mp + tid];
        const int dst = tid / Mr, e = tid - dst * Mr;
        const unsigned slot = x1_u32 + (unsigned)(((ph1 * CS + crank) * XW + e) * sizeof(double));
        st_async_f64(mapa_u32(slot, dst), a, mapa_u32(bar, dst));
      } else {
        double a = 0.0;
        if (!first)
          for (int w = 0; w * 32 < Nc; ++w) a += nrm[w * 5 + tid - M];
        const unsigned slot = x1_u32 + (unsigned)(((ph1 * CS + crank) * XW + Mr + tid - M) * sizeof(double));
#pragma unroll
        for (int c = 0; c < CS; ++c) st_async_f64(mapa_u32(slot, c), a, mapa_u32(bar, c));
      }
    }
    BPS_STAMP(7)
    mbar_wait_cluster(bars + ph1, (par >> ph1) & 1u);
    par ^= 1u << ph1;
    BPS_STAMP(8)
    // ---- my rows of t' complete -> all-gather (push); norm totals
    if (tid < nrowK || (tid >= Mr && tid < Mr + 5)) {
      const double* sl = x1s + (size_t)ph1 * CS * XW + tid;
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < CS; ++c) a += sl[c * XW];
      if (tid < nrowK) {
        const unsigned slot = t_u32 + (unsigned)((m0 + tid) * sizeof(double));
        const unsigned bar = bar_u32 + (unsigned)(4 * sizeof(uint64_t));
#pragma unroll
        for (int c = 0; c < CS; ++c) st_async_f64(mapa_u32(slot, c), a, mapa_u32(bar, c));
      } else {
        nrmt[tid - Mr] = a;
      }
    }
    BPS_STAMP(9)
    __syncthreads();
    BPS_STAMP(10)
    if (tid == 0) mbar_expect_tx(bars + ph1, X1BYTES);      // consumed by everybody: re-arm for the use after the next
    ph1 ^= 1;

    if (!first) {
      // ---- residual() / check_convergence() / update_mu() (optimizer.py:232-299) on squared norms
      const double v0 = nrmt[0], v1 = nrmt[1], v2 = nrmt[2], v3 = nrmt[3], v4 = nrmt[4];
      sq_p = v0;
      sq_d = v3;
      mu_res = mu;
      if (b.history && it < b.hist_cap && tid == 0 && crank == 0) {
        b.history[((size_t)prob * b.hist_cap + it) * 2] = sqrt(v0);
        b.history[((size_t)prob * b.hist_cap + it) * 2 + 1] = mu * sqrt(v3);
      }
      const int this_it = it;
      ++it;
      const bool conv = (v0 < rtol2 * fmax(v1, v2)) && (v3 < rtol2 * fmax(v1, v4));
      if (conv) {
        done = 1;
        break;
      }
      if (this_it % b.interval_update_mu == 0) {
        const double primal = sqrt(v0), dual = mu * sqrt(v3);
        double m2 = mu;
        if (primal > b.th_change * dual) m2 *= b.fact_incr;
        if (dual > b.th_change * primal) m2 /= b.fact_incr;
        m2 = fmin(m2, b.max_mu);
        if (m2 != mu) {
          mu = m2;
          need = 1;
          break;
        }
      }
      if (it >= iter_end) break;
    }
    BPS_STAMP(11)
    // ---- my rows of s = K^-1 t' (t' complete once every owner's rows have arrived), all-gather (push)
    mbar_wait_cluster(bars + 4, (par >> 4) & 1u);
    par ^= 1u << 4;
    for (int rl = warp; rl < nrowK; rl += NW) {
      const double* kr = Ks + rl * M;
      double a = 0.0;
      for (int j = lane; j < M; j += 32) a += kr[j] * tfull[j];
      a = warp_sum(a);
      if (lane < CS) {
        const unsigned slot = s_u32 + (unsigned)((ph2 * Mp + m0 + rl) * sizeof(double));
        const unsigned bar = bar_u32 + (unsigned)((2 + ph2) * sizeof(uint64_t));
        st_async_f64(mapa_u32(slot, lane), a, mapa_u32(bar, lane));
      }
    }
    BPS_STAMP(12)
    mbar_wait_cluster(bars + 2 + ph2, (par >> (2 + ph2)) & 1u);
    BPS_STAMP(13)
    par ^= 1u << (2 + ph2);
    ph2 ^= 1;
    first = false;
  }

  // ---- state back to global memory (every CTA its columns)
  if (own) {
    const size_t o = (size_t)prob * N + n0 + tid;
    b.x0[o] = c_x0;
    if (b.x0_old != nullptr && ran_any) b.x0_old[o] = c_xold;
    b.x1[o] = c_x1;
    b.h[o] = c_h;
  }
  if (tid == 0 && crank == 0) {
    b.mu[prob] = mu;
    b.iters[prob] = it;
    b.done[prob] = done;
    b.need_factor[prob] = need;
    b.last_res[2 * prob] = sqrt(sq_p);
    b.last_res[2 * prob + 1] = mu_res * sqrt(sq_d);
  }
  cluster.sync();                  // nobody leaves while a peer may still push into its shared memory
}

static int check_bp(const admm_bp_buffers* b, const char* who) {
  ADMM_REQUIRE(b != nullptr, ADMM_EINVAL, "%s: null buffers", who);
  ADMM_REQUIRE(b->nb >= 1 && b->M >= 1 && b->N >= 1, ADMM_EINVAL, "%s: bad dims", who);
  ADMM_REQUIRE(b->nk == (b->woodbury ? b->M : b->N), ADMM_EINVAL, "%s: nk inconsistent with woodbury flag", who);
  ADMM_REQUIRE(b->interval_update_mu >= 1, ADMM_EINVAL, "%s: interval_update_mu must be >= 1", who);
  return ADMM_OK;
}

static size_t bp_smem_bytes(const admm_bp_buffers* b) { return (size_t)(6 * b->N + 2 * b->nk + 5 * 32) * sizeof(double); }

// kernel attributes (opt-in shared memory, cluster sizes) are per device: the "already configured" caches below are
// keyed by (device, kernel)
using DevKern = std::pair<int, const void*>;

}  // namespace admm

using namespace admm;

extern "C" {

int admm_bp_supported(int M, int N) {
  if (M < 1 || N < 1) return 0;
  admm_bp_buffers b = {};
  b.M = M;
  b.N = N;
  b.nk = M < N ? M : N;
  return bp_smem_bytes(&b) <= 220 * 1024 ? 1 : 0;
}

int admm_bp_setup(const admm_bp_buffers* b, const double* y, double* aty, double* gram, admm_stream_t stream) {
  if (int rc = check_bp(b, "admm_bp_setup")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ADMM_REQUIRE(b->nb <= 65535 * 1024, ADMM_EINVAL, "admm_bp_setup: batch too large");
  for (int p0 = 0; p0 < b->nb; p0 += 65535) {
    admm_bp_buffers bb = *b;
    const int cnt = std::min(65535, b->nb - p0);
    bb.A = b->A + (size_t)p0 * b->M * b->N;
    if (y != nullptr && aty != nullptr) {
      dim3 g1(ceil_div(b->N, 256), cnt);
      bp_aty_kernel<<<g1, 256, 0, s>>>(bb, y + (size_t)p0 * b->M, aty + (size_t)p0 * b->N);
    }
    if (gram != nullptr) {
      const int tiles = ceil_div(b->nk, 32);
      dim3 g2(tiles, tiles, cnt);
      double* gp = gram + (size_t)p0 * b->nk * b->nk;
      if (b->woodbury) bp_gram_kernel<true><<<g2, 256, 0, s>>>(bb, gp);
      else bp_gram_kernel<false><<<g2, 256, 0, s>>>(bb, gp);
    }
  }
  return check_launch("admm_bp_setup");
}

int admm_bp_tile_A(const admm_bp_buffers* b, double* At, admm_stream_t stream) {
  if (int rc = check_bp(b, "admm_bp_tile_A")) return rc;
  ADMM_REQUIRE(At != nullptr, ADMM_EINVAL, "admm_bp_tile_A: null destination");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int T = (b->N + 31) / 32;
  const long long per = (long long)T * b->M * 32;
  for (int p0 = 0; p0 < b->nb; p0 += 65535) {
    admm_bp_buffers bb = *b;
    const int cnt = std::min(65535, b->nb - p0);
    bb.A = b->A + (size_t)p0 * b->M * b->N;
    dim3 g((unsigned)std::max<long long>(1, std::min<long long>((per + 255) / 256, 64)), cnt);
    bp_tile_A_kernel<<<g, 256, 0, s>>>(bb, At + (size_t)p0 * per);
  }
  return check_launch("admm_bp_tile_A");
}

int admm_bp_factor(const admm_bp_buffers* b, int* info, admm_stream_t stream) {
  if (int rc = check_bp(b, "admm_bp_factor")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nk = b->nk;
  for (int p0 = 0; p0 < b->nb; p0 += 65535) {
    admm_bp_buffers bb = *b;
    const int cnt = std::min(65535, b->nb - p0);
    bb.need_factor = b->need_factor + p0;
    bb.mu = b->mu + p0;
    bb.gram = b->gram + (size_t)p0 * nk * nk;
    bb.Kinv = b->Kinv + (size_t)p0 * nk * nk;
    dim3 g(std::max(1, std::min(8, ceil_div(nk * nk, 256))), cnt);
    bp_shift_kernel<<<g, 256, 0, s>>>(bb);
  }
  if (int rc = check_launch("admm_bp_factor(shift)")) return rc;
  if (int rc = admm_spd_inverse_batched(nk, b->nb, b->Kinv, (long long)nk * nk, nk, b->need_factor, info, stream)) return rc;
  bp_clear_flag_kernel<<<ceil_div(b->nb, 256), 256, 0, s>>>(*b);
  return check_launch("admm_bp_factor");
}

int admm_bp_iterate(const admm_bp_buffers* b, int iter_end, admm_stream_t stream) {
  if (int rc = check_bp(b, "admm_bp_iterate")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // a handful of problems: cluster-resident solve (A and K^-1 distributed over the shared memory of 16 or 8 CTAs per
  // problem; more problems than clusters fit run in waves)
  // measured on 128x512 problems (tools/bp_solo_nb_sweep.py): 8 problems 6.8 vs 10.7 us per iteration of the batch,
  // 16: 9.5 vs 12.1, 24: 13.4 vs 12.2 (nine 16-CTA clusters per wave)
  static const int solo_max_nb = getenv("ADMM_BP_SOLO_MAX") ? atoi(getenv("ADMM_BP_SOLO_MAX")) : 16;
  if (b->woodbury && b->nb <= solo_max_nb && b->M >= 8 && b->M <= 480 && b->N >= 128 && !getenv("ADMM_BP_NO_SOLO")) {
    auto try_solo = [&](auto kern, int cs) -> int {       // 0: launched, 1: not possible with this cluster size, <0: error
      const BpsLayout lay = bps_layout(b->M, b->N, cs);
      const size_t smem = (size_t)lay.total * sizeof(double);
      if (lay.Nc > BPS_THREADS || smem > 220 * 1024) return 1;
      static std::map<DevKern, int> state;                 // per (device, kernel): 0 unknown, 1 usable, -1 cluster does not fit
      int& stt = state[DevKern(cur_dev(), reinterpret_cast<const void*>(kern))];
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(b->nb * cs);
      cfg.blockDim = dim3(BPS_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (stt <= 0) {
        if (stt < 0) return 1;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int nclusters = 0;
        cudaLaunchConfig_t probe = cfg;
        probe.dynamicSmemBytes = 220 * 1024;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &probe) != cudaSuccess || nclusters < 1) {
          cudaGetLastError();
          stt = -1;
          return 1;
        }
        stt = 1;
      }
      cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *b, iter_end);
      if (e != cudaSuccess) {
        set_error("admm_bp_iterate(solo): %s", cudaGetErrorString(e));
        return -1;
      }
      return check_launch("admm_bp_iterate(solo)") == ADMM_OK ? 0 : -1;
    };
    int r = try_solo(bp_solo_kernel<16>, 16);
    if (r == 1) r = try_solo(bp_solo_kernel<8>, 8);
    if (r == 0) return ADMM_OK;
    if (r < 0) return ADMM_ECUDA;
  }
  // fused single-sweep kernel: Woodbury path with a tile-major copy of A (admm_bp_tile_A), M <= 256
  // (16 rows per warp, up to 16 warps) and the vectors + 2 tile stages fit in shared memory
  if (b->woodbury && b->At != nullptr && b->M <= 256 && !getenv("ADMM_BP_TWO_SWEEP")) {
    const bool small = b->M <= 128;
    const size_t smem = small ? BpfSmem<8>::bytes(b->N) : BpfSmem<16>::bytes(b->N);
    if (smem <= 220 * 1024) {
      // few problems: spread each over a cluster of 8, 4 or 2 CTAs (distributed shared memory exchange)
      // so that the clusters still fit the 148 SMs in one wave
      int cs = 1;
      if (!getenv("ADMM_BP_NO_CLUSTER")) {
        int sm_count = 0;
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cur_dev());
        const int slots = (small ? 2 : 1) * sm_count;   // resident CTAs of this kernel (296 / 148 on B200)
        for (int c = 8; c >= 2; c >>= 1)
          if (b->nb * c <= slots / 2) { cs = c; break; }
      }
      auto launch = [&](auto kern, int threads, int cs) -> int {
        static std::map<DevKern, size_t> configured;         // largest smem size configured per (device, kernel)
        size_t& a = configured[DevKern(cur_dev(), reinterpret_cast<const void*>(kern))];
        if (smem > a) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
          a = smem;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(b->nb * cs);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = cs > 1 ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *b, iter_end);
        if (e != cudaSuccess) {
          set_error("admm_bp_iterate(fused): %s", cudaGetErrorString(e));
          return ADMM_ECUDA;
        }
        return check_launch("admm_bp_iterate(fused)");
      };
      if (small) {
        switch (cs) {
          case 8: return launch(bp_fused_kernel<8, 8>, 256, 8);
          case 4: return launch(bp_fused_kernel<8, 4>, 256, 4);
          case 2: return launch(bp_fused_kernel<8, 2>, 256, 2);
          default: return launch(bp_fused_kernel<8, 1>, 256, 1);
        }
      }
      switch (cs) {
        case 8: return launch(bp_fused_kernel<16, 8>, 512, 8);
        case 4: return launch(bp_fused_kernel<16, 4>, 512, 4);
        case 2: return launch(bp_fused_kernel<16, 2>, 512, 2);
        default: return launch(bp_fused_kernel<16, 1>, 512, 1);
      }
    }
  }
  const size_t smem = bp_smem_bytes(b);
  ADMM_REQUIRE(smem <= 220 * 1024, ADMM_EUNSUPPORTED, "admm_bp_iterate: N=%d too large for the shared-memory resident path", b->N);
  static std::map<int, size_t> attr;      // per device
  size_t& a = attr[cur_dev()];
  if (smem > 48 * 1024 && smem > a) {
    cudaFuncSetAttribute(bp_iterate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    a = smem;
  }
  bp_iterate_kernel<<<b->nb, BP_THREADS, smem, st>>>(*b, iter_end);
  return check_launch("admm_bp_iterate");
}

}  // extern "C"
