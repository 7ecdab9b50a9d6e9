// Pattern B engine (SpM): many problems sharing s, P, C.   See include/admm_b200.h and DESIGN.md.
//
// One ADMM iteration is   step (= x-update + pass, one kernel)  -> [reduce] -> decide
// (or  xupdate -> pass -> [reduce] -> decide  when the batch is too small to fill the GPU and the
// sampling points are split over several CTAs).
//   * pass streams the implicit (Re h20, x2) state once (8 B read + 8 B written per sampling point
//     and problem) and runs both skinny GEMMs (s' = h - mu P x0,  V = P^T |s'|) on the FP64 tensor
//     cores (mma.sync m8n8k4 -> SASS DMMA.8x8x4) chained through registers; P arrives in
//     fragment-major chunks by TMA bulk copies (UBLKCP) behind an mbarrier ring;
//   * xupdate does the L x L work (cached inverse, KKT correction, soft threshold, dual ascent);
//   * decide evaluates residual()/check_convergence()/update_mu() on device.
#include "common.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <map>

namespace cg = cooperative_groups;

namespace admm {

// ---------------------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t frag_index(int ct, int NT, int j, int lane) {
  return ((size_t)(ct * NT + j) * 32 + lane) * 2;
}

// fragment-major L x L operand:  Bf[jk][jn][lane][e] = B[8 jk + 2 t + e][8 jn + g]   (lane = 4 g + t)
__host__ __device__ __forceinline__ size_t bfrag_index(int NT, int jk, int jn, int lane) {
  return ((size_t)(jk * NT + jn) * 32 + lane) * 2;
}
__host__ __device__ __forceinline__ size_t bfrag_of(int NT, int row, int col) {
  return bfrag_index(NT, row >> 3, col >> 3, 4 * (col & 7) + ((row & 7) >> 1)) + (row & 1);
}

// P (Nw x L) -> fragment-major Pf[rt][which][j][lane][e]  (zero padded):
//   which 0 (GEMM1' B operand):  P[8 rt + g      ][8 j + 2 t + e]
//   which 1 (GEMM2' B operand):  P[8 rt + 2 t + e][8 j + g      ]
// so that every operand fetch of the pass kernel is one conflict-free 16-byte load at
// (lane * 16 + immediate).
//
// Folded (d.fold, see admm_spm_dims): per PAIR tile i (sampling points 8i..8i+7 and their mirror images) first the NT
// slices of GEMM1' as above (rt = i), then 2 * NTE slices of GEMM2' whose 8 columns have ONE parity:
//   slice jj < NTE:  P[8 i + 2 t + e][16 jj + 2 g]          slice NTE + jj:  P[8 i + 2 t + e][16 jj + 2 g + 1]
__global__ void prepare_P_fold_kernel(admm_spm_dims d, const double* __restrict__ P, int ldP, double* __restrict__ Pf) {
  const int NT = d.Lp / 8, NTE = (NT + 1) / 2, NS = NT + 2 * NTE;
  const int npair = d.nrt / 2, half = d.Nw / 2;
  const long long total = (long long)npair * NS * 64;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = int(idx & 1), lane = int((idx >> 1) & 31);
    const long long q = idx >> 6;
    const int sl = int(q % NS), i = int(q / NS);
    const int g = lane >> 2, t = lane & 3;
    int r, l;
    if (sl < NT) {
      r = 8 * i + g;
      l = 8 * sl + 2 * t + e;
    } else {
      const int jj = sl - NT;
      r = 8 * i + 2 * t + e;
      l = jj < NTE ? 16 * jj + 2 * g : 16 * (jj - NTE) + 2 * g + 1;
    }
    double v = 0.0;
    if (r < half && l < d.L) v = P[(size_t)r * ldP + l];
    Pf[idx] = v;
  }
}

// P[row][l] read back from Pf (the Q = P x0 slices); folded: the mirror image of a point of the upper half, odd columns negated
__device__ __forceinline__ double pf_elem(const admm_spm_dims& d, const double* __restrict__ Pf, int row, int l) {
  const int NT = d.Lp / 8;
  if (row >= d.Nw) return 0.0;
  if (!d.fold) return Pf[((size_t)(row >> 3) * 2 * NT + (l >> 3)) * 64 + (4 * (row & 7) + ((l & 7) >> 1)) * 2 + (l & 1)];
  const bool mir = row >= d.Nw / 2;
  const int r = mir ? d.Nw - 1 - row : row, NS = NT + 2 * ((NT + 1) / 2);
  const double v = Pf[((size_t)(r >> 3) * NS + (l >> 3)) * 64 + (4 * (r & 7) + ((l & 7) >> 1)) * 2 + (l & 1)];
  return (mir && (l & 1)) ? -v : v;
}

__global__ void prepare_P_kernel(admm_spm_dims d, const double* __restrict__ P, int ldP, double* __restrict__ Pf) {
  const int NT = d.Lp / 8;
  const long long total = (long long)d.nrt * 2 * NT * 64;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = int(idx & 1), lane = int((idx >> 1) & 31);
    long long q = idx >> 6;
    const int j = int(q % NT);
    q /= NT;
    const int which = int(q & 1), rt = int(q >> 1);
    const int g = lane >> 2, t = lane & 3;
    const int r = which == 0 ? 8 * rt + g : 8 * rt + 2 * t + e;
    const int l = which == 0 ? 8 * j + 2 * t + e : 8 * j + g;
    double v = 0.0;
    if (r < d.Nw && l < d.L) v = P[(size_t)r * ldP + l];
    Pf[idx] = v;
  }
}

// canonical (L x nb) -> fragment layout
__global__ void pack_L_kernel(admm_spm_dims d, const double* __restrict__ canon, int src_cplx, double* __restrict__ frag) {
  const int NT = d.Lp / 8;
  const long long total = (long long)d.npt * d.nplanes * NT * 64;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = int(idx & 1), lane = int((idx >> 1) & 31);
    const long long q = idx >> 6;
    const int j = int(q % NT), ct = int(q / NT);
    const int pt = ct / d.nplanes, pl = ct % d.nplanes;
    const int g = lane >> 2, t = lane & 3;
    const int l = 8 * j + 2 * t + e, prob = 8 * pt + g;
    double v = 0.0;
    if (l < d.L && prob < d.nb) {
      const size_t o = (size_t)l * d.nb + prob;
      if (src_cplx) v = canon[2 * o + pl];
      else if (pl == 0) v = canon[o];
    }
    frag[idx] = v;
  }
}

__global__ void unpack_L_kernel(admm_spm_dims d, const double* __restrict__ frag, double* __restrict__ canon, int dst_cplx) {
  const int NT = d.Lp / 8;
  const long long total = (long long)d.L * d.nb;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int l = int(idx / d.nb), prob = int(idx % d.nb);
    const int pt = prob >> 3, g = prob & 7;
    const int j = l >> 3, t = (l & 7) >> 1, e = l & 1;
    const int lane = 4 * g + t;
    const double re = frag[frag_index(pt * d.nplanes, NT, j, lane) + e];
    const double im = d.nplanes == 2 ? frag[frag_index(pt * d.nplanes + 1, NT, j, lane) + e] : 0.0;
    if (dst_cplx) {
      canon[2 * idx] = re;
      canon[2 * idx + 1] = im;
    } else {
      canon[idx] = re;
    }
  }
}

// implicit (h20, x2) state: ONE real plane (the imaginary part of h20 never leaves L-space: see
// xupdate), stored chunk-major per CTA tile group so that the 4-row-tile chunk a pass CTA needs is
// ONE contiguous block (one TMA bulk copy):
//   S[grp][chunk][tig][r4][lane][2],  grp = pt / GT, tig = pt % GT, GT = 4 * mt tiles per CTA,
//   chunk = rt / 4, r4 = rt % 4.
__host__ __device__ __forceinline__ size_t state_index(const admm_spm_dims& d, int pt, int rt, int lane) {
  const int GT = 4 * d.mt;
  const int grp = pt / GT, tig = pt - grp * GT;
  const int chunk = rt >> 2, r4 = rt & 3;
  return ((((size_t)grp * (d.nrt >> 2) + chunk) * GT + tig) * 4 + r4) * 64 + lane * 2;
}

// sampling point r -> (state tile, row inside the tile).  Folded: tile 2i holds the points 8i..8i+7 of the lower half,
// tile 2i+1 their mirror images Nw-1-(8i+w) at the same row w; the padding rows Nw .. 8 nrt - 1 (the cluster-resident
// kernel walks over them like over real ones) are the unpaired slots behind the Nw/2 pairs, alternating between the planes.
__host__ __device__ __forceinline__ void state_row(const admm_spm_dims& d, int r, int& rt, int& w) {
  if (!d.fold) {
    rt = r >> 3;
    w = r & 7;
  } else if (r >= d.Nw) {
    const int k = r - d.Nw, base = d.Nw / 2 + (k >> 1);
    rt = 2 * (base >> 3) + (k & 1);
    w = base & 7;
  } else if (r < d.Nw / 2) {
    rt = 2 * (r >> 3);
    w = r & 7;
  } else {
    const int rr = d.Nw - 1 - r;
    rt = 2 * (rr >> 3) + 1;
    w = rr & 7;
  }
}
// ... and back (-1: padding)
__host__ __device__ __forceinline__ int state_point(const admm_spm_dims& d, int rt, int w) {
  if (!d.fold) return 8 * rt + w < d.Nw ? 8 * rt + w : -1;
  const int base = 8 * (rt >> 1) + w;
  if (base >= d.Nw / 2) return -1;
  return (rt & 1) ? d.Nw - 1 - base : base;
}
// the state element of (sampling point r, problem 8 pt + g)
__host__ __device__ __forceinline__ size_t state_elem(const admm_spm_dims& d, int pt, int g, int r) {
  int rt, w;
  state_row(d, r, rt, w);
  return state_index(d, pt, rt, 4 * g + (w >> 1)) + (w & 1);
}

__global__ void pack_state_kernel(admm_spm_dims d, const double* __restrict__ h20, const double* __restrict__ x2,
                                  int src_cplx, const double* __restrict__ mu20, double* __restrict__ S,
                                  int* __restrict__ flag) {
  const int GT = 4 * d.mt;
  const int npt_pad = (d.npt + GT - 1) / GT * GT;
  const long long total = (long long)npt_pad * d.nrt * 64;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = int(idx & 1), lane = int((idx >> 1) & 31);
    const long long q = idx >> 6;
    const int rt = int(q % d.nrt), pt = int(q / d.nrt);
    const int g = lane >> 2, t = lane & 3;
    const int r = state_point(d, rt, 2 * t + e), prob = 8 * pt + g;
    double v = 0.0;
    if (r >= 0 && prob < d.nb) {
      const size_t o = (size_t)r * d.nb + prob;
      const double hre = src_cplx ? h20[2 * o] : h20[o];
      const double him = src_cplx ? h20[2 * o + 1] : 0.0;
      const double xre = src_cplx ? x2[2 * o] : x2[o];
      const double xim = src_cplx ? x2[2 * o + 1] : 0.0;
      v = hre - mu20[prob] * xre;
      if (xre < 0.0 || xim != 0.0 || hre < 0.0 || (hre != 0.0 && xre != 0.0)) flag[0] = 1;
      if (d.nplanes == 1 && him != 0.0) flag[0] = 1;
    }
    S[state_index(d, pt, rt, lane) + e] = v;
  }
}

__global__ void unpack_state_kernel(admm_spm_dims d, const double* __restrict__ S, const double* __restrict__ mu20_used,
                                    const double* __restrict__ him_all, double* __restrict__ h20, double* __restrict__ x2,
                                    int dst_cplx) {
  const long long total = (long long)d.Nw * d.nb;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = int(idx / d.nb), prob = int(idx % d.nb);
    const int pt = prob >> 3, g = prob & 7;
    const double s = S[state_elem(d, pt, g, r)];
    const double him = him_all ? him_all[idx] : 0.0;
    const double hre = s > 0.0 ? s : 0.0;
    const double xv = s < 0.0 ? (-s) / mu20_used[prob] : 0.0;
    if (dst_cplx) {
      h20[2 * idx] = hre;
      h20[2 * idx + 1] = him;
      x2[2 * idx] = xv;
      x2[2 * idx + 1] = 0.0;
    } else {
      h20[idx] = hre;
      x2[idx] = xv;
    }
  }
}

// canonical Lp x Lp (row-major) -> fragment-major operand of the x-update GEMMs
__global__ void pack_operator_kernel(int Lp, const double* __restrict__ canon, double* __restrict__ Bf) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Lp * Lp; idx += gridDim.x * blockDim.x) {
    const int i = idx / Lp, j = idx - i * Lp;
    Bf[bfrag_of(Lp / 8, i, j)] = canon[idx];
  }
}

// ---------------------------------------------------------------------------------------------
// factor: Ginv = (G0 + mu10 I + mu20 PtP)^-1 by in-place Gauss-Jordan in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spm_factor_kernel(admm_spm_dims d, const int* __restrict__ slots,
                                                         const double* __restrict__ mu10s, const double* __restrict__ mu20s,
                                                         const double* __restrict__ G0, const double* __restrict__ PtP,
                                                         const double* __restrict__ Cvec, double* __restrict__ Ginv_cache,
                                                         double* __restrict__ w_cache, double* __restrict__ sigma_cache,
                                                         int* __restrict__ info) {
  extern __shared__ double sm[];
  const int n = d.L, Lp = d.Lp;
  double* a = sm;               // n x n
  double* rowk = a + n * n;     // n
  double* colk = rowk + n;      // n
  double* wv = colk + n;        // n
  __shared__ double red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int slot = slots[blockIdx.x];
  const double mu10 = mu10s[blockIdx.x], mu20 = mu20s[blockIdx.x];
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx / n, j = idx - i * n;
    a[idx] = G0[(size_t)i * Lp + j] + (i == j ? mu10 : 0.0) + mu20 * PtP[(size_t)i * Lp + j];
  }
  __syncthreads();
  int bad = 0;
  for (int k = 0; k < n; ++k) {
    const double p = a[k * n + k];
    if (!(p > 0.0)) bad = k + 1;
    const double ip = 1.0 / p;
    for (int j = tid; j < n; j += nt) {
      rowk[j] = (j == k ? 1.0 : a[k * n + j]) * ip;
      colk[j] = (j == k ? 0.0 : a[j * n + k]);
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
      const int i = idx / n, j = idx - i * n;
      if (i == k) {
        a[idx] = rowk[j];
      } else {
        const double base = (j == k) ? 0.0 : a[idx];
        a[idx] = base - colk[i] * rowk[j];
      }
    }
    __syncthreads();
  }
  double* Gi = Ginv_cache + (size_t)slot * Lp * Lp;   // fragment-major (bfrag_of), zero padded
  for (int idx = tid; idx < Lp * Lp; idx += nt) {
    const int i = idx / Lp, j = idx - i * Lp;
    Gi[bfrag_of(Lp / 8, i, j)] = (i < n && j < n) ? 0.5 * (a[i * n + j] + a[j * n + i]) : 0.0;
  }
  __syncthreads();
  // w_k = Ginv c_k (symmetrised Ginv) for every constraint row, Sigma = C W; nc = 1: sigma itself is cached, nc > 1: the
  // inverse of the nc x nc matrix Sigma (Gauss-Jordan by one thread; Sigma is symmetric positive definite)
  const int nc = d.nc < 1 ? 1 : d.nc;
  __shared__ double sig[ADMM_SPM_MAX_NC * ADMM_SPM_MAX_NC];
  for (int k = 0; k < nc; ++k) {
    const double* ck = Cvec + (size_t)k * Lp;
    for (int i = tid; i < n; i += nt) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += 0.5 * (a[i * n + j] + a[j * n + i]) * ck[j];
      wv[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < Lp; i += nt) w_cache[((size_t)slot * nc + k) * Lp + i] = i < n ? wv[i] : 0.0;
    for (int m = 0; m < nc; ++m) {
      const double* cm = Cvec + (size_t)m * Lp;
      double v[1] = {0.0};
      for (int i = tid; i < n; i += nt) v[0] += cm[i] * wv[i];
      block_sum<1>(v, red);
      if (tid == 0) sig[m * nc + k] = v[0];
      __syncthreads();
    }
  }
  if (tid == 0) {
    if (nc == 1) {
      sigma_cache[slot] = sig[0];
    } else {
      double inv[ADMM_SPM_MAX_NC * ADMM_SPM_MAX_NC];
      for (int i = 0; i < nc; ++i)
        for (int j = 0; j < nc; ++j) inv[i * nc + j] = i == j ? 1.0 : 0.0;
      for (int k = 0; k < nc; ++k) {
        const double pv = sig[k * nc + k];
        if (!(pv > 0.0) && bad == 0) bad = n + k + 1;
        const double ip = 1.0 / pv;
        for (int j = 0; j < nc; ++j) {
          sig[k * nc + j] *= ip;
          inv[k * nc + j] *= ip;
        }
        for (int i = 0; i < nc; ++i) {
          if (i == k) continue;
          const double f = sig[i * nc + k];
          for (int j = 0; j < nc; ++j) {
            sig[i * nc + j] -= f * sig[k * nc + j];
            inv[i * nc + j] -= f * inv[k * nc + j];
          }
        }
      }
      for (int i = 0; i < nc * nc; ++i) sigma_cache[(size_t)slot * nc * nc + i] = inv[i];
    }
    if (info) info[blockIdx.x] = bad;
  }
}

// ---------------------------------------------------------------------------------------------
// batch-wide criterion: peer mailboxes and the "lazy" decision folded into the iteration kernels
// ---------------------------------------------------------------------------------------------
// Mailbox word = (seq << 32) | 32-bit half of a double: one atomic 8-byte store carries data and validity
// (the receiver polls the word itself -- one-way NVLink latency, no fence, no separate flag).
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const int* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const int* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }

constexpr int PEER_WORDS = 20;        // ten doubles as 32-bit halves
constexpr int PEER_SLOT = 32;         // words per (buffer, source rank) slot of a mailbox
constexpr long long PEER_TIMEOUT = 10000000000LL;   // ~5 s of clock64: a peer died, give up instead of hanging the GPU

// push the ten sums gs[] (shared memory) into every rank's mailbox under sequence number seq; all threads of the CTA
__device__ __forceinline__ void peer_post(const admm_peer_comm& c, const double* gs, unsigned seq) {
  for (int i = threadIdx.x; i < PEER_WORDS * c.world; i += blockDim.x) {
    const int peer = i / PEER_WORDS, k = i - peer * PEER_WORDS;
    const double val = gs[k >> 1];
    const unsigned half = (k & 1) ? (unsigned)__double2hiint(val) : (unsigned)__double2loint(val);
    unsigned long long* box = c.mbox[peer] + ((size_t)(seq & 1u) * ADMM_MAX_PEERS + c.rank) * PEER_SLOT + k;
    st_relaxed_sys_u64(box, ((unsigned long long)seq << 32) | half);
  }
}

// sum number i (< 10) of ALL ranks for sequence seq from the local mailbox, added in rank order; false: timed out
__device__ __forceinline__ bool peer_gather(const admm_peer_comm& c, unsigned seq, int i, double& out) {
  const unsigned long long* box = c.mbox[c.rank] + (size_t)(seq & 1u) * ADMM_MAX_PEERS * PEER_SLOT + 2 * i;
  const long long t0 = clock64();
  while (true) {
    bool ok = true;
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < ADMM_MAX_PEERS; ++r) {        // independent loads: all in flight at once
      if (r < c.world) {
        const unsigned long long lo = ld_relaxed_sys_u64(box + r * PEER_SLOT), hi = ld_relaxed_sys_u64(box + r * PEER_SLOT + 1);
        ok = ok && (unsigned)(lo >> 32) == seq && (unsigned)(hi >> 32) == seq;
        a += __hiloint2double((int)(unsigned)hi, (int)(unsigned)lo);
      }
    }
    if (ok) {
      out = a;
      return true;
    }
    if (clock64() - t0 > PEER_TIMEOUT) return false;
  }
}

// residual() and check_convergence() (optimizer.py:232-274) from the ten squared norms; 0/0 -> NaN -> not converged
__device__ __forceinline__ bool residual_conv(const double (&s)[10], double mu10, double mu20, double rtol, double& primal,
                                              double& dual, double (&parts)[4]) {
  const double p10 = sqrt(s[0]), nx0 = sqrt(s[1]), nx1 = sqrt(s[2]), nd = sqrt(s[3]), nxo = sqrt(s[4]);
  const double nPd = sqrt(s[5]), nPxo = sqrt(s[6]), p20 = sqrt(s[7]), nx2 = sqrt(s[8]), nPx0 = sqrt(s[9]);
  const double d10 = mu10 * nd, d20 = mu20 * nPd;
  primal = p10 + p20;
  dual = d10 + d20;
  parts[0] = p10;
  parts[1] = d10;
  parts[2] = p20;
  parts[3] = d20;
  return (p10 / fmax(nx0, nx1) < rtol) && (d10 / fmax(mu10 * nx0, mu10 * nxo) < rtol) &&
         (p20 / fmax(nPx0, nx2) < rtol) && (d20 / fmax(mu20 * nPx0, mu20 * nPxo) < rtol);
}

// Head of a lazy iteration kernel: the decision of the PREVIOUS iteration.  Every CTA evaluates it redundantly and
// identically; true = the batch has converged, the caller returns without touching the state.  All threads of the CTA.
__device__ __forceinline__ bool lazy_head(const admm_spm_dims& d, const admm_spm_buffers& b, const admm_peer_comm& c, int pending) {
  if (__ldcg(b.lazy + 2) != 0) return true;             // converged in an earlier kernel
  if (!pending) return false;
  __shared__ double lz_gs[10];
  __shared__ int lz_fail;
  const int tid = threadIdx.x;
  if (tid == 0) lz_fail = 0;
  __syncthreads();
  if (tid < 10) {
    if (c.world == 0) {
      lz_gs[tid] = __ldcg(b.gsum + tid);                // left by the last CTA of the previous kernel
    } else {
      double a = 0.0;
      if (!peer_gather(c, c.ctrl[0], tid, a)) lz_fail = 1;
      lz_gs[tid] = a;
    }
  }
  __syncthreads();
  const bool first_cta = blockIdx.x == 0 && blockIdx.y == 0;
  if (lz_fail) {
    if (first_cta && tid == 0) b.flags[2] = -2;
    return true;
  }
  double s[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) s[i] = lz_gs[i];
  // check_convergence (optimizer.py:232-249) on the squared norms: p / max(a, b) < rtol  <=>  p^2 < rtol^2 max(a^2, b^2)
  // (mu > 0 cancels in the dual tests; 0/0 and x/0 stay "not converged") -- FP64 square roots and divisions are long
  // instruction sequences, and this runs at the head of every CTA
  const double rtol2 = b.rtol * b.rtol;
  const bool conv = (s[0] < rtol2 * fmax(s[1], s[2])) && (s[3] < rtol2 * fmax(s[1], s[4])) &&
                    (s[7] < rtol2 * fmax(s[9], s[8])) && (s[5] < rtol2 * fmax(s[9], s[6]));
  if (first_cta && tid == 0) {
    const double primal = sqrt(s[0]) + sqrt(s[7]), dual = b.mu10[0] * sqrt(s[3]) + b.mu20[0] * sqrt(s[5]);
    const int it = b.iters[0];
    if (b.history && it < b.hist_cap) {
      b.history[2 * it] = primal;
      b.history[2 * it + 1] = dual;
    }
    b.iters[0] = it + 1;
    b.iter_counter[0] = it + 1;
    b.last_res[0] = primal;
    b.last_res[1] = dual;
    if (conv) {
      b.flags[1] = d.nb;
      __threadfence();
      b.lazy[2] = 1;
    }
  }
  return conv;
}

// Tail of a lazy iteration kernel.  `mine` (shared memory, [nwarps][10]) holds the partial sums of this CTA's warps;
// they go to cta_partA (x-update stage or fused kernel: all ten) or cta_partB (pass of the unfused path: entries 7, 8).
// `ticketed`: this kernel closes the iteration -- the CTA that arrives last adds all nA + nB CTA partials in a fixed
// order and publishes the batch-wide sums (gsum, or the peers' mailboxes).  `scratch`: >= 256 doubles of shared
// memory nobody else uses any more.  All threads of the CTA (blockDim.x == 128).
__device__ __forceinline__ void lazy_tail(const admm_spm_buffers& b, const admm_peer_comm& c, double* mine, int nwarps,
                                          bool to_A, bool ticketed, int nA, int nB, double* scratch, int stamp = 0) {
  __shared__ int lz_last;
  const int tid = threadIdx.x;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
  __syncthreads();
  if (tid < 10) {
    double a = 0.0;
    for (int w = 0; w < nwarps; ++w) a += mine[w * 10 + tid];
    if (to_A) __stcg(b.cta_partA + (size_t)cta * 10 + tid, a);
    else if (tid == 7 || tid == 8) __stcg(b.cta_partB + (size_t)cta * 2 + (tid - 7), a);
  }
  if (!ticketed) return;
  __syncthreads();
  if (tid == 0) {
    __threadfence();                  // (cumulative: orders the partials the barrier made visible to this thread)
    lz_last = (atomicAdd(reinterpret_cast<unsigned*>(b.lazy), 1u) == (unsigned)(ncta - 1));
  }
  __syncthreads();
  if (!lz_last) return;
  __threadfence();
  // fixed-order final sum with many loads in flight: thread t < 120 owns value index t % 10 of the x-update-stage
  // partials (flat stride 120), every thread owns index 7 + (t & 1) of the pass partials (flat stride 128)
  double a = 0.0, bs = 0.0;
  if (tid < 120) {      // (16 loads in flight per thread: the whole sum of a 444-CTA grid is three round trips)
    const int n = nA * 10;
#pragma unroll 16
    for (int i = tid; i < n; i += 120) a += __ldcg(b.cta_partA + i);
  }
  {
    const int n = nB * 2;
#pragma unroll 16
    for (int i = tid; i < n; i += 128) bs += __ldcg(b.cta_partB + i);
  }
  scratch[tid] = a;
  scratch[128 + tid] = bs;
  __syncthreads();
  if (tid < 10) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) tot += scratch[tid + 10 * k];
    if (tid == 7 || tid == 8) {
      for (int k = 0; k < 64; ++k) tot += scratch[128 + 2 * k + (tid - 7)];
    }
    mine[tid] = tot;
    b.gsum[tid] = tot;
  }
  __syncthreads();
  if (c.world > 0) {
    const unsigned seq = c.ctrl[0] + 1u;
    __syncthreads();
    peer_post(c, mine, seq);
    if (tid == 0) c.ctrl[0] = seq;
  }
  if (tid == 0) {
    b.lazy[0] = 0;
    if (stamp != 0) b.lazy[3] = stamp;         // launch sequence number of the fused balanced kernel
  }
}

#ifdef SPM_TRACE   // tools only: %globaltimer stamps of CTA 0 / thread 0 inside the x-update, int64 slots 1024.. of b.gpart
__device__ __forceinline__ long long gtimer_x() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
#define XSTAMP(idx)                                                                                             \
  if (threadIdx.x == 0) {                                                                                       \
    const long long t_now = gtimer_x();                                                                         \
    if (blockIdx.x == 0) reinterpret_cast<long long*>(b.gpart)[1024 + (idx)] = t_now;                           \
    if (blockIdx.x < 512) reinterpret_cast<long long*>(b.gpart)[4096 + 16 * blockIdx.x + (idx)] = t_now;         \
  }
#else
#define XSTAMP(idx)
#endif

// ---------------------------------------------------------------------------------------------
// x-update of ONE problem tile (8 problems, all planes) by one warp, everything in fragment layout
// ---------------------------------------------------------------------------------------------
// out[p][c][l'] = sum_l a[p][c][l] * B[l][l']  for NP planes that share the B fragments
// (k-slot (jk,e) of lane (g,t) <-> l = 8 jk + 2 t + e); NT*NP independent accumulation chains.
template <int NT, int NP, bool SMEM = false>      // SMEM: the operand was staged in shared memory
__device__ __forceinline__ void frag_gemm(double (&out)[NP][NT][2], const double (&a)[NP][NT][2],
                                          const double* __restrict__ Bf, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int jn = 0; jn < NT; ++jn) out[p][jn][0] = out[p][jn][1] = 0.0;
#pragma unroll
  for (int jk = 0; jk < NT; ++jk) {
    double2 bb[NT];
#pragma unroll
    for (int jn = 0; jn < NT; ++jn)
      bb[jn] = SMEM ? *reinterpret_cast<const double2*>(Bf + bfrag_index(NT, jk, jn, lane))
                    : __ldg(reinterpret_cast<const double2*>(Bf + bfrag_index(NT, jk, jn, lane)));
#pragma unroll
    for (int jn = 0; jn < NT; ++jn)
#pragma unroll
      for (int p = 0; p < NP; ++p) dmma(out[p][jn][0], out[p][jn][1], a[p][jk][0], bb[jn].x);
#pragma unroll
    for (int jn = 0; jn < NT; ++jn)
#pragma unroll
      for (int p = 0; p < NP; ++p) dmma(out[p][jn][0], out[p][jn][1], a[p][jk][1], bb[jn].y);
  }
}

// Term 0 solve (ConstrainedLeastSquares with the cached inverse and KKT correction), L1 z-update,
// dual ascent of pair (1,0), Gram-form norms of pair (2,0) (|P v|^2 = v^T (P^T P) v with
// y = P^T P x0 cached from iteration to iteration) and -- imaginary plane -- the L-space
// recursion of z = P^T Im(h20).  Re(x0) (new) is returned in registers for the pass.
// y0 = P^T P x0_old must be current (spm_refresh_y_kernel after a state change).
// Returns false when every problem of the tile is frozen (nothing was touched, x0 not loaded).
// Handles planes p0 .. p0+NP-1 of the tile (plane 0 = real, plane 1 = imaginary parts).
// wpart != NULL (lazy batch-wide iterations): instead of per-problem norms in normsA, the sums over the tile's live
// problems are added to wpart[0..9] (this warp's row of the CTA's partial sums, indexed like gather_problem's s[]).
// Stand-alone kernel (PRE): the iteration is a chain of dependent steps on a few warps, so every global-memory
// round trip shows.  The L x L operands are staged in shared memory by the whole CTA (sm_PtP; sm_Ginv = the factor of
// cache row sm_slot, -1: none) and every load that does not depend on a computed value is issued up front.
// EARLY: the late operands (old x0, y0, h10) are kept in registers from the start (stand-alone kernel); PRE without EARLY
// (owner CTAs of the fused balanced step, whose registers belong to the pass): they are prefetched into L1 instead.
template <int NT, int NP, bool SPLIT, bool PRE = false, bool EARLY = PRE>   // SPLIT: V arrives as d.nsplit partial sums
__device__ __forceinline__ bool xupdate_tile(const admm_spm_dims& d, const admm_spm_buffers& b, int pt, int lane,
                                             int p0 = 0, double* wpart = nullptr, const double* sm_PtP = nullptr,
                                             const double* sm_Ginv = nullptr, int sm_slot = -1, uint64_t* ops_bar = nullptr) {
  const int g = lane >> 2, t = lane & 3;
  const int prob = 8 * pt + g;
  const int is_done = b.done[prob];
  const double mu10 = b.mu10[prob], mu20 = b.mu20[prob];      // (issued together with `done`: one round trip)
  const int slot = b.slot[prob];
  XSTAMP(0)
  if (__all_sync(0xffffffffu, is_done)) return false;
  XSTAMP(1)
  const int Lp = d.Lp;
  const int ct0 = pt * d.nplanes + p0;
  const size_t vstride = (size_t)d.npt * d.nplanes * NT * 64;

  // PRE: old x0, y0 = P^T P x0_old and the KKT vectors are requested now, long before they are used
  double xo_pre[EARLY ? NP : 1][NT][2], yo_pre[EARLY ? NP : 1][NT][2], hh_pre[EARLY ? NP : 1][NT][2];
  double cw_pre[PRE ? NT : 1][4];      // C[8j+2t], C[8j+2t+1], w[8j+2t], w[8j+2t+1]
  double sig_pre = 0.0, D_pre[PRE ? NP : 1];
  if (PRE) {
    const double* wv = b.w_cache + (size_t)slot * Lp;
    sig_pre = b.sigma_cache[slot];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      D_pre[p] = b.Dre[(size_t)(p0 + p) * 8 * d.npt + prob];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const size_t o = frag_index(ct0 + p, NT, j, lane);
        if (EARLY) {
          const double2 a = *reinterpret_cast<const double2*>(b.x0 + o), y2 = *reinterpret_cast<const double2*>(b.y0 + o);
          xo_pre[p][j][0] = a.x;
          xo_pre[p][j][1] = a.y;
          yo_pre[p][j][0] = y2.x;
          yo_pre[p][j][1] = y2.y;
        } else {
          asm volatile("prefetch.global.L1 [%0];\n" ::"l"(b.x0 + o));
          asm volatile("prefetch.global.L1 [%0];\n" ::"l"(b.y0 + o));
          if (p0 + p == 1) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(b.aim + o));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      cw_pre[j][0] = b.Cvec[8 * j + 2 * t];
      cw_pre[j][1] = b.Cvec[8 * j + 2 * t + 1];
      cw_pre[j][2] = wv[8 * j + 2 * t];
      cw_pre[j][3] = wv[8 * j + 2 * t + 1];
    }
  }

  // ---- rhs = alpha A^H y + h10 + mu10 x1 + P^T(h20 + mu20 x2)
  double x0[NP][NT][2];
  {
    double rhs[NP][NT][2];
    // Every load of this block is issued before the first dependent instruction: all (plane, row tile) pieces of the
    // four base vectors at once, then the partial sums of V four splits at a time (predicated, summed in split order --
    // a loop with a run-time trip count per row tile serialised ~15 L2 round trips here).
    double2 vsum[NP][NT];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const size_t o = frag_index(ct0 + p, NT, j, lane);
        const double2 b0 = *reinterpret_cast<const double2*>(b.b0 + o);
        const double2 hh = *reinterpret_cast<const double2*>(b.h10 + o);
        const double2 x1 = *reinterpret_cast<const double2*>(b.x1 + o);
        vsum[p][j] = *reinterpret_cast<const double2*>(b.V + o);
        rhs[p][j][0] = b0.x + hh.x + mu10 * x1.x;
        rhs[p][j][1] = b0.y + hh.y + mu10 * x1.y;
        if (EARLY) {
          hh_pre[p][j][0] = hh.x;
          hh_pre[p][j][1] = hh.y;
        }
      }
    }
    if (SPLIT && p0 == 0) {      // (imaginary plane: z lives in split 0, owned by this function)
      const int nsp = d.nsplit;
#pragma unroll 1
      for (int sp0 = 1; sp0 < nsp; sp0 += 4) {
        double2 q[4][NT];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            q[k][j] = make_double2(0.0, 0.0);
            if (sp0 + k < nsp)
              q[k][j] = *reinterpret_cast<const double2*>(b.V + (size_t)(sp0 + k) * vstride + frag_index(ct0, NT, j, lane));
          }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (sp0 + k < nsp) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              vsum[0][j].x += q[k][j].x;
              vsum[0][j].y += q[k][j].y;
            }
          }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        rhs[p][j][0] += vsum[p][j].x;
        rhs[p][j][1] += vsum[p][j].y;
      }
    // ---- xi1 = Ginv rhs, one tensor-core GEMM per distinct factor slot in the tile (one slot unless
    // the problems of the tile sit on different (mu10, mu20): per-problem mode only)
#ifdef SPM_TRACE
    if (rhs[0][0][0] == 1.2345e300) return false;      // (forces the loads to have arrived before the stamp)
#endif
    XSTAMP(2)
    if (PRE && ops_bar != nullptr) mbar_wait(ops_bar, 0);      // staged operands (bulk copies issued at kernel start) have landed
    unsigned remaining = __ballot_sync(0xffffffffu, !is_done);
#pragma unroll 1
    do {
      const int cur = __shfl_sync(0xffffffffu, slot, __ffs(remaining) - 1);
      double acc[NP][NT][2];
      if (PRE && cur == sm_slot) frag_gemm<NT, NP, true>(acc, rhs, sm_Ginv, lane);
      else frag_gemm<NT, NP>(acc, rhs, b.Ginv_cache + (size_t)cur * Lp * Lp, lane);
      if (slot == cur) {
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            x0[p][j][0] = acc[p][j][0];
            x0[p][j][1] = acc[p][j][1];
          }
      }
      remaining &= ~__ballot_sync(0xffffffffu, slot == cur);
    } while (remaining);
  }

  XSTAMP(3)
  // ---- KKT correction enforcing C x0 = D   (objectivefunc.py:148-157)
  if (d.nc > 1) {
    // several constraint rows: nu = (C G^-1 C^T)^-1 (D - C xi1), x0 = xi1 + sum_k w_k nu_k   (operands from L1/L2)
    const int nc = d.nc;
    const double* Sinv = b.sigma_cache + (size_t)slot * nc * nc;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      double resid[ADMM_SPM_MAX_NC];
#pragma unroll
      for (int k = 0; k < ADMM_SPM_MAX_NC; ++k) {
        resid[k] = 0.0;
        if (k < nc) {
          const double* ck = b.Cvec + (size_t)k * Lp;
          double cxi = 0.0;
#pragma unroll
          for (int j = 0; j < NT; ++j) cxi += ck[8 * j + 2 * t] * x0[p][j][0] + ck[8 * j + 2 * t + 1] * x0[p][j][1];
          cxi = quad_sum(cxi);
          resid[k] = b.Dre[((size_t)(p0 + p) * nc + k) * 8 * d.npt + prob] - cxi;
        }
      }
#pragma unroll
      for (int k = 0; k < ADMM_SPM_MAX_NC; ++k) {
        if (k < nc) {
          double nu = 0.0;
#pragma unroll
          for (int m = 0; m < ADMM_SPM_MAX_NC; ++m)
            if (m < nc) nu += Sinv[k * nc + m] * resid[m];
          const double* wk = b.w_cache + ((size_t)slot * nc + k) * Lp;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            x0[p][j][0] += wk[8 * j + 2 * t] * nu;
            x0[p][j][1] += wk[8 * j + 2 * t + 1] * nu;
          }
        }
      }
    }
  } else {
    const double* wv = b.w_cache + (size_t)slot * Lp;
    const double isig = 1.0 / (PRE ? sig_pre : b.sigma_cache[slot]);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      double cxi = 0.0;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double c0 = PRE ? cw_pre[j][0] : b.Cvec[8 * j + 2 * t], c1 = PRE ? cw_pre[j][1] : b.Cvec[8 * j + 2 * t + 1];
        cxi += c0 * x0[p][j][0] + c1 * x0[p][j][1];
      }
      cxi = quad_sum(cxi);
      const double nu = ((PRE ? D_pre[p] : b.Dre[(size_t)(p0 + p) * 8 * d.npt + prob]) - cxi) * isig;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        x0[p][j][0] += (PRE ? cw_pre[j][2] : wv[8 * j + 2 * t]) * nu;
        x0[p][j][1] += (PRE ? cw_pre[j][3] : wv[8 * j + 2 * t + 1]) * nu;
      }
    }
  }

  XSTAMP(4)
  // ---- y = P^T P x0 (both planes in one GEMM)
  double y[NP][NT][2];
  if (PRE && sm_PtP != nullptr) frag_gemm<NT, NP, true>(y, x0, sm_PtP, lane);
  else frag_gemm<NT, NP>(y, x0, b.PtPf, lane);

  XSTAMP(5)
  const double thr = 0.5 * b.lam / mu10;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    // Loads first, stores last: the compiler may not move a load above a store that could alias it, so interleaving
    // them row tile by row tile made every row tile a dependent L2 round trip (measured: 4 us per plane).
    // ---- norms of the x0 change, plain and in Gram form
    double n_d = 0.0, n_xo = 0.0, nPd = 0.0, nPxo = 0.0, nPx = 0.0;
    {
      double2 xo2[NT], yo2[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const size_t o = frag_index(ct0 + p, NT, j, lane);
        xo2[j] = EARLY ? make_double2(xo_pre[p][j][0], xo_pre[p][j][1]) : *reinterpret_cast<const double2*>(b.x0 + o);
        yo2[j] = EARLY ? make_double2(yo_pre[p][j][0], yo_pre[p][j][1]) : *reinterpret_cast<const double2*>(b.y0 + o);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double xo[2] = {xo2[j].x, xo2[j].y}, yo[2] = {yo2[j].x, yo2[j].y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double dd = x0[p][j][e] - xo[e];
          n_d += dd * dd;
          n_xo += xo[e] * xo[e];
          nPd += dd * (y[p][j][e] - yo[e]);
          nPxo += xo[e] * yo[e];
          nPx += x0[p][j][e] * y[p][j][e];
        }
      }
      // `_x_old[0]` of the reference (optimizer.py:324): only kept when the caller asks for it
      if (b.x0_old != nullptr && !is_done) {
#pragma unroll
        for (int j = 0; j < NT; ++j) *reinterpret_cast<double2*>(b.x0_old + frag_index(ct0 + p, NT, j, lane)) = xo2[j];
      }
    }
    // ---- second batch of loads: h10 and, imaginary plane, z and a
    double2 hh2[NT], zv2[NT], av2[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const size_t o = frag_index(ct0 + p, NT, j, lane);
      hh2[j] = EARLY ? make_double2(hh_pre[p][j][0], hh_pre[p][j][1]) : *reinterpret_cast<const double2*>(b.h10 + o);
      if (p0 + p == 1) {
        zv2[j] = *reinterpret_cast<const double2*>(b.V + o);
        av2[j] = *reinterpret_cast<const double2*>(b.aim + o);
      }
    }
    // ---- imaginary plane: z <- z - mu20 y,  a += mu20 x0   (Im(h20) never leaves L-space)
    if (p0 + p == 1) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        zv2[j].x -= mu20 * y[p][j][0];
        zv2[j].y -= mu20 * y[p][j][1];
        av2[j].x += mu20 * x0[p][j][0];
        av2[j].y += mu20 * x0[p][j][1];
      }
    }
    // ---- L1 z-update (real plane only; the imaginary part of x1 is identically zero) + dual ascent
    double n_p = 0.0, n_x0 = 0.0, n_x1 = 0.0;
    double2 zz2[NT], hn2[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      double zz[2], hn[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double xv = x0[p][j][e], hv = e == 0 ? hh2[j].x : hh2[j].y;
        double z = 0.0;
        if (p0 + p == 0) {
          const double yv = -((hv - mu10 * xv) / mu10);
          if (yv > thr) z = yv - thr;
          if (yv < -thr) z = yv + thr;
        }
        hn[e] = hv + mu10 * (z - xv);
        zz[e] = z;
        n_p += (xv - z) * (xv - z);
        n_x0 += xv * xv;
        n_x1 += z * z;
      }
      zz2[j] = make_double2(zz[0], zz[1]);
      hn2[j] = make_double2(hn[0], hn[1]);
    }
    // ---- all stores of the plane
    if (!is_done) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const size_t o = frag_index(ct0 + p, NT, j, lane);
        if (p0 + p == 1) {
          *reinterpret_cast<double2*>(b.V + o) = zv2[j];
          *reinterpret_cast<double2*>(b.aim + o) = av2[j];
        }
        *reinterpret_cast<double2*>(b.x0 + o) = make_double2(x0[p][j][0], x0[p][j][1]);
        *reinterpret_cast<double2*>(b.x1 + o) = zz2[j];
        *reinterpret_cast<double2*>(b.h10 + o) = hn2[j];
        *reinterpret_cast<double2*>(b.y0 + o) = make_double2(y[p][j][0], y[p][j][1]);
      }
    }
    n_p = quad_sum(n_p);
    n_x0 = quad_sum(n_x0);
    n_x1 = quad_sum(n_x1);
    n_d = quad_sum(n_d);
    n_xo = quad_sum(n_xo);
    nPd = quad_sum(nPd);
    nPxo = quad_sum(nPxo);
    nPx = quad_sum(nPx);
    if (wpart != nullptr) {
      // sum over the 8 problems of the tile (every lane of a quad holds its problem's value: combine across quads)
      double nv[8] = {n_p, n_x0, n_x1, n_d, n_xo, nPd > 0.0 ? nPd : 0.0, nPxo > 0.0 ? nPxo : 0.0, nPx > 0.0 ? nPx : 0.0};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double v = is_done ? 0.0 : nv[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        nv[i] = v;
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) wpart[i] += nv[i];
        wpart[9] += nv[7];                           // |P x0|^2 (Gram form), both planes
        if (p0 + p == 1) wpart[7] += nv[7];          // |P Im(x0) - 0|^2
      }
    } else if (t == 0 && !is_done) {
      double* nrm = b.normsA + ((size_t)(ct0 + p) * 8 + g) * 8;
      nrm[0] = n_p;
      nrm[1] = n_x0;
      nrm[2] = n_x1;
      nrm[3] = n_d;
      nrm[4] = n_xo;
      nrm[5] = nPd > 0.0 ? nPd : 0.0;
      nrm[6] = nPxo > 0.0 ? nPxo : 0.0;
      nrm[7] = nPx > 0.0 ? nPx : 0.0;     // |P x0|^2 of this plane
    }
    XSTAMP(6 + p)
  }
  return true;
}

// Stand-alone x-update (small batches): one warp per (problem tile, plane) -- the kernel is a chain of
// dependent loads and two small GEMMs, so more, shorter warps finish sooner than fewer, longer ones.
// lazy: 0 classic (per-problem norms, the reduce/decide kernels follow), 1 lazy iteration without, 2 with a pending
// decision of the previous iteration (see lazy_head / lazy_tail)
template <int NT>
__global__ void __launch_bounds__(128) spm_xupdate_kernel(admm_spm_dims d, admm_spm_buffers b, admm_peer_comm c, int lazy) {
  pdl_prologue();
  __shared__ double wsum[4 * 10];
  if (d.batch_wide && b.lazy != nullptr) {
    if (lazy ? lazy_head(d, b, c, lazy == 2) : (__ldcg(b.lazy + 2) != 0)) return;
  }
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  // stage the two L x L operands (fragment-major, NT*NT*64 doubles each) with 16-byte async copies: P^T P at once, the
  // cached inverse as soon as the factor row of the CTA's first problem is known (batch-wide: the row of all of them)
  extern __shared__ __align__(16) double sm_ops[];        // [2][NT*NT*64]
  constexpr int OPV = NT * NT * 32;                        // 16-byte vectors per operand
  for (int i = threadIdx.x; i < OPV; i += 128) cp_async16(sm_ops + 2 * i, b.PtPf + 2 * i);
  const int wfirst = (blockIdx.x * blockDim.x) >> 5;
  const int pfirst = min(8 * (wfirst / d.nplanes), 8 * d.npt - 1);
  const int slot0 = b.slot[pfirst];
  {
    const double* gi = b.Ginv_cache + (size_t)slot0 * d.Lp * d.Lp;
    for (int i = threadIdx.x; i < OPV; i += 128) cp_async16(sm_ops + NT * NT * 64 + 2 * i, gi + 2 * i);
  }
  cp_async_commit();
  if (lazy) {
    if (lane < 10) wsum[wl * 10 + lane] = 0.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (warp < d.npt * d.nplanes)
    xupdate_tile<NT, 1, true, true>(d, b, warp / d.nplanes, lane, warp % d.nplanes, lazy ? wsum + wl * 10 : nullptr, sm_ops,
                              sm_ops + NT * NT * 64, slot0);
  if (lazy) lazy_tail(b, c, wsum, 4, true, false, 0, 0, sm_ops);
}

// y0 = P^T P x0 for every column tile (after the state was loaded from outside)
template <int NT>
__global__ void __launch_bounds__(128) spm_refresh_y_kernel(admm_spm_dims d, admm_spm_buffers b) {
  const int ct = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ct >= d.npt * d.nplanes) return;
  double x[1][NT][2], y[1][NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const double2 v = *reinterpret_cast<const double2*>(b.x0 + frag_index(ct, NT, j, lane));
    x[0][j][0] = v.x;
    x[0][j][1] = v.y;
  }
  frag_gemm<NT, 1>(y, x, b.PtPf, lane);
#pragma unroll
  for (int j = 0; j < NT; ++j)
    *reinterpret_cast<double2*>(b.y0 + frag_index(ct, NT, j, lane)) = make_double2(y[0][j][0], y[0][j][1]);
}

// ---------------------------------------------------------------------------------------------
// pass: the streaming sweep with both skinny GEMMs on the FP64 tensor cores
// ---------------------------------------------------------------------------------------------
constexpr int PASS_WARPS = 4;     // warps per CTA; each warp owns MT tiles of 8 problems
constexpr int PASS_STAGES = 2;    // TMA ring depth (one stage = a P chunk + the CTA's state chunk)
constexpr int PASS_CHUNK_RT = 4;  // 8-row tiles per chunk (32 sampling points)

enum { PASS_STEP = 0, PASS_VINIT = 1 };

// sign-bit helpers on the integer pipe (the FP64 pipe is shared with DMMA: keep it for the MMAs)
__device__ __forceinline__ bool is_neg(double v) { return __double2hiint(v) < 0; }

#ifdef SPM_TRACE   // tools only: %globaltimer stamps (ns) of CTAs 0, 1 and the last one into b.gpart: [launch & 7][cta slot][8]
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
#define PASS_STAMP(idx)                                                                                         \
  if (threadIdx.x == 0 && (blockIdx.x < 2 || blockIdx.x == gridDim.x - 1))                                      \
    reinterpret_cast<long long*>(b.gpart)[((trace_launch & 7) * 3 + (blockIdx.x < 2 ? blockIdx.x : 2)) * 8 + (idx)] = gtimer();
#else
#define PASS_STAMP(idx)
#endif

#ifndef FOLD1_CTAS
#define FOLD1_CTAS 4      // resident CTAs per SM the folded one-tile-per-warp kernel is compiled for
#endif
template <int NT, int MT, bool FOLD = false>
struct PassSmem {
  static constexpr int NTE = (NT + 1) / 2;                            // folded: slices of V per column parity
  static constexpr int TILE_D = FOLD ? (NT + 2 * NTE) * 64 : 2 * NT * 64;      // doubles of Pf per 8-row (pair) tile
  static constexpr int CHUNK_D = (FOLD ? PASS_CHUNK_RT / 2 : PASS_CHUNK_RT) * TILE_D;      // P doubles per chunk
  static constexpr int WARP_STATE_D = MT * PASS_CHUNK_RT * 64;        // state doubles per warp and chunk
  static constexpr int STATE_D = PASS_WARPS * WARP_STATE_D;           // state doubles per CTA and chunk
  static constexpr int STAGE_D = CHUNK_D + STATE_D;
  // (folded: the P chunk is small -- keep room behind stage 0 for both staged L x L operands of the balanced step)
  static constexpr int RING_D = (FOLD && PASS_STAGES * STAGE_D < STAGE_D + 2 * NT * NT * 64) ? STAGE_D + 2 * NT * NT * 64
                                                                                              : PASS_STAGES * STAGE_D;
  static constexpr size_t BYTES = (size_t)RING_D * sizeof(double) + 2 * PASS_STAGES * sizeof(uint64_t) +
                                  PASS_STAGES * sizeof(unsigned) + 16;
};

// Work decomposition.  The unit is a "group-chunk": one chunk (32 sampling points) of one CTA tile
// group (4*MT problem tiles); the state layout makes group-chunks of consecutive linear index
// gc = group * nchunks + chunk consecutive 16 KB blocks.  A CTA processes the linear range
// [g_begin, g_end):
//   * classic grid (d.nbal == 0): blockIdx.x = group, blockIdx.y = split (equal chunk ranges);
//   * balanced (d.nbal > 0, small batches): the T = ngroups * nchunks group-chunks are cut into d.nbal
//     equal pieces, one per CTA, so that one wave of CTAs fills every SM equally; a CTA's range may
//     straddle a group boundary, then it finishes the first group (epilogue) and starts the next.
// Either way the partial V / norm sums of (group, piece) go to split slot `piece`; slots a group does
// not use stay zero, the x-update sums all d.nsplit of them.
// lazy (batch-wide criterion, MODE == PASS_STEP): 0 classic (per-problem norms, the reduce/decide kernels follow);
// 1 / 2: lazy iteration without / with a pending decision of the previous iteration (lazy_head / lazy_tail); nA = CTAs
// of the stand-alone x-update kernel whose partial sums the last CTA of this pass adds (unfused path).
//
// BAL (FNP != 0 with the balanced decomposition): the WHOLE iteration of a small batch in one launch.  First the
// x-update: its units (problem tile, plane) are dealt out to all warps of the grid (V = the sum of the partial slots the
// pieces left in the previous iteration); a warp that finishes a unit adds one to xready[G] of the tile's group
// (release).  Then the pass: before a CTA touches chunks of group G it waits until xready[G] has reached
// (launch sequence number) x (units of G) (acquire).  All CTAs of the single wave are co-resident (cooperative launch;
// the launcher also checks the occupancy) and the x-update waits for nobody, so the waits always end; a watchdog turns a
// broken assumption into flags[2] = -3.
//
// FOLD (d.fold: P has the parity of the IR basis, see admm_spm_dims): the pass works on PAIRS of sampling points
// (r, r' = Nw-1-r), P[r'][l] = (-1)^l P[r][l].  The fragment layout already separates the column parities -- k-step
// e = 0 of GEMM1' carries the even l = 8j+2t, e = 1 the odd ones -- so two accumulation chains started at
// (h + h')/2 and (h - h')/2 end as E, O with s'(r) = E + O, s'(r') = E - O: one row's MMAs serve two.  GEMM2' runs over
// slices of one column parity with u + u' (even) or u - u' (odd) as the A operand; a quad shuffle per segment brings V
// back to the fragment layout of the x-update.  22 instead of 40 DMMAs per 16 sampling points and problem tile.
template <int NT, int MT, int MODE, int FNP, bool BAL = false, bool FOLD = false>   // FNP: 0 = pass only, 1/2 = fused x-update of 1/2 planes
__global__ void __launch_bounds__(PASS_WARPS * 32, (FOLD && MT == 1 && NT == 5 && !BAL) ? FOLD1_CTAS : (MT == 2 || NT > 2) ? 3 : 4)      // (shared memory admits 3 CTAs per SM for NT = 5)
    spm_pass_kernel(admm_spm_dims d, admm_spm_buffers b, admm_peer_comm c, int lazy, int nA) {
#ifdef SPM_TRACE
  const long long t_entry = gtimer();
#endif
  pdl_prologue();
#ifdef SPM_TRACE
  const int trace_launch = b.lazy[3];
  if (threadIdx.x == 0 && (blockIdx.x < 2 || blockIdx.x == gridDim.x - 1))
    reinterpret_cast<long long*>(b.gpart)[((trace_launch & 7) * 3 + (blockIdx.x < 2 ? blockIdx.x : 2)) * 8 + 7] = t_entry;
#endif
  PASS_STAMP(0)
  __shared__ double wsum[PASS_WARPS * 10];        // lazy: partial sums of the ten squared norms, one row per warp
  if (!BAL && MODE == PASS_STEP && d.batch_wide && b.lazy != nullptr) {
    // the fused kernel opens the iteration (decision of the previous one); the stand-alone pass follows the x-update
    // kernel, which has taken it already.  (BAL: further down, with the first copies already in flight)
    if ((lazy && FNP != 0) ? lazy_head(d, b, c, lazy == 2) : (__ldcg(b.lazy + 2) != 0)) return;
  }
  using SM = PassSmem<NT, MT, FOLD>;
  constexpr int TILE_D = SM::TILE_D, CHUNK_D = SM::CHUNK_D, STATE_D = SM::STATE_D, STAGE_D = SM::STAGE_D;
  constexpr unsigned CHUNK_BYTES = CHUNK_D * sizeof(double), STATE_BYTES = STATE_D * sizeof(double);
  constexpr int GT = PASS_WARPS * MT;
  extern __shared__ __align__(128) double ring[];       // [PASS_STAGES][P chunk | state chunk], then barriers
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + SM::RING_D);
  uint64_t* empty_bar = full_bar + PASS_STAGES;
  unsigned* ticket = reinterpret_cast<unsigned*>(empty_bar + PASS_STAGES);
  uint64_t* ops_bar = reinterpret_cast<uint64_t*>(ticket + PASS_STAGES);      // BAL: the staged x-update operands

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nct = d.nrt / PASS_CHUNK_RT;                 // chunks per group
  const int npl = d.nplanes;
  long long g_begin, g_end;
  if (d.nbal > 0) {
    const long long T = (long long)((d.npt + GT - 1) / GT) * nct;
    if (b.bal_bounds != nullptr) {      // pieces of unequal length (the caller's table; see admm_spm_buffers)
      g_begin = b.bal_bounds[blockIdx.x];
      g_end = b.bal_bounds[blockIdx.x + 1];
    } else {
      g_begin = (long long)blockIdx.x * T / d.nbal;
      g_end = (long long)(blockIdx.x + 1) * T / d.nbal;
    }
  } else {
    const int cps = (nct + d.nsplit - 1) / d.nsplit;     // chunks per split
    const int c_begin = min(nct, (int)blockIdx.y * cps), c_end = min(nct, c_begin + cps);
    g_begin = (long long)blockIdx.x * nct + c_begin;
    g_end = (long long)blockIdx.x * nct + c_end;
  }
  const int nchunks = (int)(g_end - g_begin);

  // ---- barrier ring; the first chunks are in flight before anything else happens
  auto fill = [&](int stage, long long gc) {
    mbar_expect_tx(full_bar + stage, CHUNK_BYTES + STATE_BYTES);
    tma_bulk_g2s(ring + stage * STAGE_D, b.Pf + (size_t)(gc % nct) * CHUNK_D, CHUNK_BYTES, full_bar + stage);
    tma_bulk_g2s(ring + stage * STAGE_D + CHUNK_D, b.S + (size_t)gc * STATE_D, STATE_BYTES, full_bar + stage);
  };
  if (tid == 0) {
    for (int s = 0; s < PASS_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, PASS_WARPS);
      ticket[s] = 0;
    }
    if (BAL) mbar_init(ops_bar, 1);
    fence_barrier_init();
  }
  if (MODE == PASS_STEP && lazy && lane < 10) wsum[warp * 10 + lane] = 0.0;      // (own row: ordered by program order per warp)
  // BAL: sequence number of this launch (the CTA that finishes last advances lazy[3]) and this warp's first x-update
  // unit.  A unit is (problem tile, plane); the npt * NPL units are dealt out to ALL warps of the grid, round k to warp
  // (k + cta + cta / nsm) % 4 of CTA (unit % ncta): with the usual breadth-first placement of a 3-CTAs-per-SM grid the
  // CTAs of one SM then work on different SM sub-partitions -- the x-update is bound by the FP64 pipe of its
  // sub-partition, and with whole tile groups on "owner" CTAs (first version) up to three owners shared an SM and took
  // 23 us instead of 12 while two thirds of the SMs idled.
  int bal_stamp = 0, bal_unit0 = -1, bal_ustride = 0;
  if (BAL) {
    constexpr int NPLc = FNP == 0 ? 1 : FNP;
    bal_stamp = __ldcg(b.lazy + 3) + 1;
    const int ncta = gridDim.x, nsm = max(1, ncta / 3);
    const int k0 = (warp - (int)blockIdx.x - (int)blockIdx.x / nsm) & (PASS_WARPS - 1);      // round in which this warp is served first
    bal_unit0 = k0 * ncta + blockIdx.x;
    bal_ustride = PASS_WARPS * ncta;
    if (bal_unit0 >= d.npt * NPLc) bal_unit0 = -1;
  }
  const bool bal_xwork = BAL && __syncthreads_or(bal_unit0 >= 0);      // (also the barrier after the mbarrier init)
  if (!BAL) __syncthreads();
  constexpr int OPD = NT * NT * 64;                                         // doubles per L x L operand
  constexpr bool OPS_BOTH = 2 * OPD <= SM::RING_D - STAGE_D;                // room for P^T P and the cached inverse?
  static_assert(!BAL || OPD <= SM::RING_D - STAGE_D, "ring stage too small for a staged operand");
  int bal_slot0 = -1;
  if (tid == 0) {
    // (a CTA with x-update units stages the L x L operands in ring stage 1 first: that stage is filled afterwards)
    for (int s = 0; s < PASS_STAGES && s < nchunks; ++s)
      if (!(bal_xwork && s == 1)) fill(s, g_begin + s);
  }
  if (bal_xwork) {
    // the factor row of the first problem this CTA touches (batch-wide: the row of all problems) and P^T P: two bulk
    // copies into ring stage 1, completing on their own barrier -- in flight during the head and the loads of the
    // x-update, awaited just before its first GEMM
    constexpr int NPLc = FNP == 0 ? 1 : FNP;
    const int ufirst = min((int)blockIdx.x, d.npt * NPLc - 1);
    bal_slot0 = b.slot[min(8 * (ufirst / NPLc), 8 * d.npt - 1)];
    if (tid == 0) {
      constexpr unsigned OPB = OPD * sizeof(double);
      mbar_expect_tx(ops_bar, OPS_BOTH ? 2 * OPB : OPB);
      tma_bulk_g2s(ring + STAGE_D, b.Ginv_cache + (size_t)bal_slot0 * d.Lp * d.Lp, OPB, ops_bar);
      if (OPS_BOTH) tma_bulk_g2s(ring + STAGE_D + OPD, b.PtPf, OPB, ops_bar);
    }
  }
  if (BAL && MODE == PASS_STEP && d.batch_wide && b.lazy != nullptr) {
    // the decision of the previous iteration (reads only: the copies above are harmless if the batch has converged,
    // but they have to land before the CTA may exit)
    if (lazy ? lazy_head(d, b, c, lazy == 2) : (__ldcg(b.lazy + 2) != 0)) {
      if (tid == 0) {
        for (int s = 0; s < PASS_STAGES && s < nchunks; ++s)
          if (!(bal_xwork && s == 1)) mbar_wait(full_bar + s, 0);
        if (bal_xwork) mbar_wait(ops_bar, 0);
      }
      __syncthreads();
      return;
    }
  }

  PASS_STAMP(1)
  if (bal_xwork) {
    // ---- x-update of this CTA's units, one (tile, plane) per warp and round
    constexpr int NPL = FNP == 0 ? 1 : FNP;
    double* sm_ops = ring + STAGE_D;                                        // ring stage 1
    const int slot0 = bal_slot0;
    constexpr bool BOTH = OPS_BOTH;
    XSTAMP(8)
    XSTAMP(9)
    double* wp = lazy ? wsum + warp * 10 : nullptr;
#pragma unroll 1
    for (int u = bal_unit0; u >= 0 && u < d.npt * NPL; u += bal_ustride) {
      const int ptm = u / NPL;
      xupdate_tile<NT, 1, true, true, false>(d, b, ptm, lane, u - ptm * NPL, wp, BOTH ? sm_ops + OPD : nullptr, sm_ops, slot0,
                                             ops_bar);
      // publish: one more unit of the tile's group is done (release; the waiters acquire)
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(b.xready + ptm / GT, 1);
      }
    }
    XSTAMP(10)
    __syncthreads();
    XSTAMP(12)
#ifdef SPM_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 1000) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;\n" : "=r"(smid));
      reinterpret_cast<long long*>(b.gpart)[2048 + 2 * blockIdx.x] = ((gtimer() - t_entry) << 12) | smid;
    }
#endif
    if (tid == 0 && nchunks > 1) {
      fence_proxy_async();               // the staged operands were written through the generic proxy
      fill(1, g_begin + 1);
    }
  }
  PASS_STAMP(2)
  if (FNP != 0 && !BAL) {
    // x-update of this warp's tiles right here (all planes): x0 reaches the MMA operand registers
    // through L1/L2; the pass of the other CTAs of the SM hides the latency of this L x L work.
    // The L-vectors of the tiles are pulled into L2 up front (one bulk prefetch per vector).
    constexpr int NPL = FNP == 0 ? 1 : FNP;
    const int wpt0 = (blockIdx.x * PASS_WARPS + warp) * MT;
    double* wp = lazy ? wsum + warp * 10 : nullptr;
    {
      const double* vec = lane == 0 ? b.b0 : lane == 1 ? b.h10 : lane == 2 ? b.x1 : lane == 3 ? b.V
                        : lane == 4 ? b.x0 : lane == 5 ? b.y0 : b.aim;
      const int ntl = min(MT, d.npt - wpt0);
      if (lane < 7 && ntl > 0)
        l2_prefetch_bulk(vec + frag_index(wpt0 * NPL, NT, 0, 0), (unsigned)(ntl * NPL * NT * 64 * sizeof(double)));
    }
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      if (wpt0 + m < d.npt) xupdate_tile<NT, NPL, false>(d, b, wpt0 + m, lane, 0, wp);
    }
  }

  // this lane's slot inside a state chunk, in shared memory (read) and in global memory (write back)
  const int st_off = warp * SM::WARP_STATE_D + lane * 2;
  const int nctc = d.npt * npl;
  const size_t vstride = (size_t)nctc * NT * 64;

  int stage = 0;
  unsigned parity = 0;
  long long gc = g_begin;
#pragma unroll 1
  while (gc < g_end) {
    // ================= one segment: chunks [gc, seg_end) of tile group `grp` =================
    const int grp = (int)(gc / nct);
    const long long seg_end = min(g_end, (long long)(grp + 1) * nct);
    int piece;                                           // split slot of this (group, range)
    if (d.nbal > 0) {
      const long long T = (long long)((d.npt + GT - 1) / GT) * nct;
      const long long x = (long long)grp * nct;          // first group-chunk of the group
      piece = b.bal_first != nullptr ? (int)blockIdx.x - b.bal_first[grp] : (int)blockIdx.x - (int)(((x + 1) * d.nbal - 1) / T);
    } else {
      piece = blockIdx.y;
    }

    // ---- this warp's MT problem tiles (real plane only: the imaginary plane lives in L-space)
    int pt[MT];
    bool inr[MT];
    int dn[MT];
    double mu20[MT];
    double xa[MT][NT][2];      // A fragments of GEMM1': -mu20 * Re(x0)   (k-slot (j,e) of lane (g,t) <-> l = 8j+2t+e)
    constexpr int NTE = SM::NTE, NACC = FOLD ? 2 * NTE : NT;
    double acc[MT][NACC][2];   // C fragments of GEMM2': V   (folded: NTE slices of even columns, then NTE of odd ones)
    bool all_done = true;
    // (done / mu20 are not written by this launch: requested before the wait below, one round trip less after it)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      pt[m] = (grp * PASS_WARPS + warp) * MT + m;
      inr[m] = pt[m] < d.npt;
      if (!inr[m]) pt[m] = d.npt - 1;        // clamp: loads stay in range, stores are suppressed
      const int prob = 8 * pt[m] + g;
      dn[m] = inr[m] ? b.done[prob] : 1;
      mu20[m] = b.mu20[prob];
    }
    if (BAL) {
      // x0 of this group comes from the warps that ran its x-update units: wait until all of them have reported
      // (xready counts units, it runs on from launch to launch)
      constexpr int NPLc = FNP == 0 ? 1 : FNP;
      const int bal_expect = bal_stamp * NPLc * (min(d.npt, (grp + 1) * GT) - grp * GT);
      // relaxed polling, ONE acquire fence at the end (an ld.acquire per poll is a load plus an invalidation of the
      // SM's whole L1)
      if (lane == 0) {
        const long long t_start = clock64();
        while ((int)ld_relaxed_u32(b.xready + grp) != bal_expect) {
          if (clock64() - t_start > 4000000000LL) {      // the CTAs are not co-resident after all: give up (~2 s)
            b.flags[2] = -3;
            break;
          }
        }
        fence_acquire_gpu();
      }
      __syncwarp();
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double2 v = BAL ? __ldcg(reinterpret_cast<const double2*>(b.x0 + frag_index(pt[m] * npl, NT, j, lane)))
                              : *reinterpret_cast<const double2*>(b.x0 + frag_index(pt[m] * npl, NT, j, lane));
        xa[m][j][0] = -mu20[m] * v.x;
        xa[m][j][1] = -mu20[m] * v.y;
      }
#pragma unroll
      for (int j = 0; j < NACC; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;
      all_done = all_done && dn[m];
    }
    const bool active = !__all_sync(0xffffffffu, all_done);
    PASS_STAMP(3)
#ifdef SPM_TRACE
    if (threadIdx.x == 0 && blockIdx.x < 1000 && gc == g_begin)
      reinterpret_cast<long long*>(b.gpart)[2048 + 2 * blockIdx.x + 1] = gtimer() - t_entry;
#endif

    double n_dh[MT], n_xm[MT], ratio[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      n_dh[m] = n_xm[m] = 0.0;
      ratio[m] = MODE == PASS_VINIT ? mu20[m] / b.mu20_used[8 * pt[m] + g] : 1.0;
    }
    double* Sg = b.S + (size_t)gc * STATE_D + st_off;

#pragma unroll 1
    for (; gc < seg_end; ++gc) {
      mbar_wait(full_bar + stage, parity);      // idle warps wait too: nobody runs ahead of the ring
      if (active) {
        const double* Pc = ring + stage * STAGE_D + lane * 2;
        const double* Sc = ring + stage * STAGE_D + CHUNK_D + st_off;
        // The 2*NT operand fragments of an 8-row tile, GEMM1' then GEMM2', and the tiles of a chunk follow each other
        // at a fixed stride (64 doubles): fragment k of the chunk sits at Pc + 64 k.  In the step they are fetched ONE
        // FRAGMENT AHEAD with volatile loads, so that the fetch of k + 1 is issued before the MMAs of k (the compiler
        // otherwise sinks every LDS next to its first use: ncu showed each group of four DMMAs waiting ~30 cycles on the
        // short scoreboard).  Measured: the per-clock rate of the sweep does not move (the other warps of the scheduler
        // cover those waits, and the board runs at its power cap) -- kept because it is free; a distance of two
        // fragments (-DPF_DIST=2) and the state of the next tile one tile ahead made no difference either.
        constexpr int NFRAG = CHUNK_D / 64;
        const unsigned pc_addr = smem_u32(Pc);
#ifndef PF_DIST
#define PF_DIST 1
#endif
        double2 pf_q[PF_DIST];        // fragments k .. k + PF_DIST - 1 (a small queue in registers)
#pragma unroll
        for (int i = 0; i < PF_DIST; ++i) pf_q[i] = make_double2(0.0, 0.0);
        if (MODE == PASS_STEP) {
#pragma unroll
          for (int i = 0; i < PF_DIST; ++i) pf_q[i] = lds_v2_volatile(pc_addr + i * 512);
        }
        auto next_frag = [&](int k) -> double2 {      // returns fragment k, requests fragment k + PF_DIST
          const double2 r = pf_q[0];
#pragma unroll
          for (int i = 0; i + 1 < PF_DIST; ++i) pf_q[i] = pf_q[i + 1];
          if (k + PF_DIST < NFRAG) pf_q[PF_DIST - 1] = lds_v2_volatile(pc_addr + (k + PF_DIST) * 512);
          return r;
        };
        if constexpr (FOLD) {
#pragma unroll
          for (int f2 = 0; f2 < PASS_CHUNK_RT / 2; ++f2) {      // pair tiles of the chunk: state tiles 2 f2 (points), 2 f2 + 1 (mirrors)
            const double* P2 = Pc + f2 * TILE_D + NT * 64;
            constexpr int NS = NT + 2 * NTE;
            double2 st[MT], sp[MT];
            const unsigned sc_addr = smem_u32(Sc + 2 * f2 * 64);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              st[m] = *reinterpret_cast<const double2*>(Sc + (m * PASS_CHUNK_RT + 2 * f2) * 64);
              sp[m] = *reinterpret_cast<const double2*>(Sc + (m * PASS_CHUNK_RT + 2 * f2 + 1) * 64);
            }
            double up[MT][2], um[MT][2];      // u + u', u - u'
            if (MODE == PASS_STEP) {
              double qe[MT][2], qo[MT][2], hre[MT][2], hrp[MT][2];
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                hre[m][0] = is_neg(st[m].x) ? 0.0 : st[m].x;
                hre[m][1] = is_neg(st[m].y) ? 0.0 : st[m].y;
                hrp[m][0] = is_neg(sp[m].x) ? 0.0 : sp[m].x;
                hrp[m][1] = is_neg(sp[m].y) ? 0.0 : sp[m].y;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  qe[m][e] = 0.5 * hre[m][e] + 0.5 * hrp[m][e];
                  qo[m][e] = 0.5 * hre[m][e] - 0.5 * hrp[m][e];
                }
              }
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                const double2 bb = next_frag(f2 * NS + j);
#pragma unroll
                for (int m = 0; m < MT; ++m) dmma(qe[m][0], qe[m][1], xa[m][j][0], bb.x);      // even columns
#pragma unroll
                for (int m = 0; m < MT; ++m) dmma(qo[m][0], qo[m][1], xa[m][j][1], bb.y);      // odd columns
              }
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                // (the old state is fetched again rather than kept in registers across the MMA chains)
                st[m] = lds_v2_volatile(sc_addr + (m * PASS_CHUNK_RT) * 512);
                sp[m] = lds_v2_volatile(sc_addr + (m * PASS_CHUNK_RT + 1) * 512);
                hre[m][0] = is_neg(st[m].x) ? 0.0 : st[m].x;
                hre[m][1] = is_neg(st[m].y) ? 0.0 : st[m].y;
                hrp[m][0] = is_neg(sp[m].x) ? 0.0 : sp[m].x;
                hrp[m][1] = is_neg(sp[m].y) ? 0.0 : sp[m].y;
                double sn[2], sq[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const double s_new = qe[m][e] + qo[m][e], s_mir = qe[m][e] - qo[m][e];
                  const bool neg = is_neg(s_new), negp = is_neg(s_mir);
                  const double dh = hre[m][e] - (neg ? 0.0 : s_new), dhp = hrp[m][e] - (negp ? 0.0 : s_mir);
                  const double xm = neg ? s_new : 0.0, xmp = negp ? s_mir : 0.0;
                  n_dh[m] += dh * dh;
                  n_dh[m] += dhp * dhp;
                  n_xm[m] += xm * xm;
                  n_xm[m] += xmp * xmp;
                  const double ua = fabs(s_new), ub = fabs(s_mir);
                  up[m][e] = ua + ub;
                  um[m][e] = ua - ub;
                  sn[e] = dn[m] ? (e == 0 ? st[m].x : st[m].y) : s_new;
                  sq[e] = dn[m] ? (e == 0 ? sp[m].x : sp[m].y) : s_mir;
                }
                st_stream2(Sg + (m * PASS_CHUNK_RT + 2 * f2) * 64, make_double2(sn[0], sn[1]));
                st_stream2(Sg + (m * PASS_CHUNK_RT + 2 * f2 + 1) * 64, make_double2(sq[0], sq[1]));
              }
            } else {
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const double u0 = is_neg(st[m].x) ? -st[m].x * ratio[m] : st[m].x, u1 = is_neg(st[m].y) ? -st[m].y * ratio[m] : st[m].y;
                const double v0 = is_neg(sp[m].x) ? -sp[m].x * ratio[m] : sp[m].x, v1 = is_neg(sp[m].y) ? -sp[m].y * ratio[m] : sp[m].y;
                up[m][0] = u0 + v0;
                up[m][1] = u1 + v1;
                um[m][0] = u0 - v0;
                um[m][1] = u1 - v1;
              }
            }
            // ---- GEMM2' over slices of one column parity
            if constexpr (MT == 1) {      // one tile per warp: two slices at a time, so that no MMA waits for its predecessor
#pragma unroll
              for (int j = 0; j < 2 * NTE; j += 2) {
                double2 b0, b1;
                if (MODE == PASS_STEP) {
                  b0 = next_frag(f2 * NS + NT + j);
                  b1 = next_frag(f2 * NS + NT + j + 1);
                } else {
                  b0 = *reinterpret_cast<const double2*>(P2 + j * 64);
                  b1 = *reinterpret_cast<const double2*>(P2 + (j + 1) * 64);
                }
                dmma(acc[0][j][0], acc[0][j][1], j < NTE ? up[0][0] : um[0][0], b0.x);
                dmma(acc[0][j + 1][0], acc[0][j + 1][1], j + 1 < NTE ? up[0][0] : um[0][0], b1.x);
                dmma(acc[0][j][0], acc[0][j][1], j < NTE ? up[0][1] : um[0][1], b0.y);
                dmma(acc[0][j + 1][0], acc[0][j + 1][1], j + 1 < NTE ? up[0][1] : um[0][1], b1.y);
              }
            } else
#pragma unroll
            for (int j = 0; j < 2 * NTE; ++j) {
              double2 bb;
              if (MODE == PASS_STEP) {
                bb = next_frag(f2 * NS + NT + j);
              } else {
                bb = *reinterpret_cast<const double2*>(P2 + j * 64);
              }
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma(acc[m][j][0], acc[m][j][1], j < NTE ? up[m][0] : um[m][0], bb.x);
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma(acc[m][j][0], acc[m][j][1], j < NTE ? up[m][1] : um[m][1], bb.y);
            }
          }
        } else {
#pragma unroll
          for (int r4 = 0; r4 < PASS_CHUNK_RT; ++r4) {
            const double* P1 = Pc + r4 * TILE_D;        // GEMM1' operand: [j][lane][2]
            const double* P2 = P1 + NT * 64;            // GEMM2' operand: [j][lane][2]
            double2 st[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) st[m] = *reinterpret_cast<const double2*>(Sc + (m * PASS_CHUNK_RT + r4) * 64);

            double u[MT][2];
            if (MODE == PASS_STEP) {
              // ---- GEMM1': s' = max(0,s) - mu20 * (P x0)   (accumulator starts at Re h20 = max(0,s))
              double q[MT][2], q2[MT][2], hre[MT][2];
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                hre[m][0] = is_neg(st[m].x) ? 0.0 : st[m].x;
                hre[m][1] = is_neg(st[m].y) ? 0.0 : st[m].y;
                q[m][0] = hre[m][0];
                q[m][1] = hre[m][1];
                q2[m][0] = q2[m][1] = 0.0;
              }
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                const double2 bb = next_frag(r4 * 2 * NT + j);
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                  if (MT == 1) {      // a single tile per warp: two chains hide the DMMA latency
                    dmma(q[m][0], q[m][1], xa[m][j][0], bb.x);
                    dmma(q2[m][0], q2[m][1], xa[m][j][1], bb.y);
                  } else {
                    dmma(q[m][0], q[m][1], xa[m][j][0], bb.x);
                    dmma(q[m][0], q[m][1], xa[m][j][1], bb.y);
                  }
                }
              }
              // ---- elementwise: s' encodes both the dual ascent and the non-negative projection
              //   Re h20' = max(0, s'),  mu20 x2' = max(0, -s'),  mu20 (P x0 - x2') = Re h20 - Re h20'
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                double sn[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const double s_new = MT == 1 ? q[m][e] + q2[m][e] : q[m][e];
                  const bool neg = is_neg(s_new);
                  const double hnew = neg ? 0.0 : s_new;
                  const double xm = neg ? s_new : 0.0;
                  const double dh = hre[m][e] - hnew;
                  n_dh[m] += dh * dh;
                  n_xm[m] += xm * xm;
                  u[m][e] = fabs(s_new);
                  sn[e] = dn[m] ? (e == 0 ? st[m].x : st[m].y) : s_new;
                }
                st_stream2(Sg + (m * PASS_CHUNK_RT + r4) * 64, make_double2(sn[0], sn[1]));
              }
            } else {
              // V from the current state, no step:  u = Re h20 + mu20 x2  with x2 decoded by mu20_used
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                u[m][0] = is_neg(st[m].x) ? -st[m].x * ratio[m] : st[m].x;
                u[m][1] = is_neg(st[m].y) ? -st[m].y * ratio[m] : st[m].y;
              }
            }

            // ---- GEMM2': V[c][l] += sum_r u[c][r] P[r][l],  k-slot t of step e <-> row 2t+e
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              double2 bb;
              if (MODE == PASS_STEP) {
                bb = next_frag(r4 * 2 * NT + NT + j);
              } else {
                bb = *reinterpret_cast<const double2*>(P2 + j * 64);
              }
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma(acc[m][j][0], acc[m][j][1], u[m][0], bb.x);
#pragma unroll
              for (int m = 0; m < MT; ++m) dmma(acc[m][j][0], acc[m][j][1], u[m][1], bb.y);
            }
          }
        }
      }
      Sg += STATE_D;

      // ---- release the stage; the warp that arrives last refills it with chunk gc + PASS_STAGES
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(empty_bar + stage);
        const unsigned tk = atomicAdd(ticket + stage, 1u);
        if ((tk & (PASS_WARPS - 1)) == PASS_WARPS - 1 && gc + PASS_STAGES < g_end) {
          mbar_wait(empty_bar + stage, parity);
          fill(stage, gc + PASS_STAGES);
        }
      }
      if (++stage == PASS_STAGES) {
        stage = 0;
        parity ^= 1u;
      }
    }

    PASS_STAMP(4)
    // ---- epilogue of the segment: partial V (fragment layout) and per-column norm partials
    if (active) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (!inr[m]) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const size_t o = piece * vstride + frag_index(pt[m] * npl, NT, j, lane);
          if constexpr (FOLD) {
            // fragment slot (j, e) of lane (g, t) is column l = 8j + 2t + e: column 4 (j & 1) + t of parity slice j / 2,
            // which lane (g, 2 (j & 1) + t / 2) holds as element t & 1
            const int src = (lane & ~3) | (2 * (j & 1) + (t >> 1));
            const double e0 = __shfl_sync(0xffffffffu, acc[m][j >> 1][0], src), e1 = __shfl_sync(0xffffffffu, acc[m][j >> 1][1], src);
            const double o0 = __shfl_sync(0xffffffffu, acc[m][NTE + (j >> 1)][0], src),
                         o1 = __shfl_sync(0xffffffffu, acc[m][NTE + (j >> 1)][1], src);
            *reinterpret_cast<double2*>(b.V + o) = make_double2((t & 1) ? e1 : e0, (t & 1) ? o1 : o0);
          } else {
            *reinterpret_cast<double2*>(b.V + o) = make_double2(acc[m][j][0], acc[m][j][1]);
          }
        }
        if (MODE == PASS_STEP) {
          const double inv = 1.0 / mu20[m];
          const double s0 = quad_sum(n_dh[m]) * inv * inv, s1 = quad_sum(n_xm[m]) * inv * inv;
          if (lazy) {
            double a0 = dn[m] ? 0.0 : s0, a1 = dn[m] ? 0.0 : s1;      // sum over the tile's live problems (across quads)
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
              a0 += __shfl_xor_sync(0xffffffffu, a0, o);
              a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            }
            if (lane == 0) {
              wsum[warp * 10 + 7] += a0;      // |P Re(x0) - x2|^2
              wsum[warp * 10 + 8] += a1;      // |x2|^2
            }
            // the decide kernel is not run: record the mu20 this pass encoded x2 with
            if (t == 0 && !dn[m]) b.mu20_used[8 * pt[m] + g] = mu20[m];
          } else if (t == 0 && !dn[m]) {
            double* o = b.normsB + ((size_t)piece * nctc * 8 + (size_t)(pt[m] * npl) * 8 + g) * 2;
            o[0] = s0;      // |P Re(x0) - x2|^2
            o[1] = s1;      // |x2|^2
          }
        }
      }
    }
  }
  PASS_STAMP(5)
#ifdef SPM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 512) reinterpret_cast<long long*>(b.gpart)[4096 + 16 * blockIdx.x + 13] = gtimer() - t_entry;
#endif
  if (MODE == PASS_STEP && lazy) {
    const int ncta = gridDim.x * gridDim.y;
    lazy_tail(b, c, wsum, PASS_WARPS, FNP != 0, true, FNP != 0 ? ncta : nA, FNP != 0 ? 0 : ncta, ring,     // (ring: all chunks consumed)
              BAL ? bal_stamp : 0);
  } else if (BAL) {
    // classic iteration (per-problem norms, the reduce/decide kernels follow): only the launch sequence number moves on
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(reinterpret_cast<unsigned*>(b.lazy), 1u) == gridDim.x * gridDim.y - 1) {
        b.lazy[0] = 0;
        b.lazy[3] = bal_stamp;
      }
    }
  }
  PASS_STAMP(6)
#ifdef SPM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 512) {
    reinterpret_cast<long long*>(b.gpart)[4096 + 16 * blockIdx.x + 14] = gtimer() - t_entry;
    reinterpret_cast<long long*>(b.gpart)[4096 + 16 * blockIdx.x + 15] = t_entry;
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// norms -> residual / convergence / mu
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_problem(const admm_spm_dims& d, const admm_spm_buffers& b, int prob, double (&s)[10]) {
  const int pt = prob >> 3, g = prob & 7;
  const int nct = d.npt * d.nplanes;
#pragma unroll
  for (int i = 0; i < 10; ++i) s[i] = 0.0;
  for (int pl = 0; pl < d.nplanes; ++pl) {
    const size_t col = (size_t)(pt * d.nplanes + pl) * 8 + g;
    const double* a = b.normsA + col * 8;
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] += a[i];
    s[9] += a[7];                                // |P x0|^2 (Gram form), both planes
    if (pl == 0) {
      for (int sp = 0; sp < d.nsplit; ++sp) {
        const double* bb = b.normsB + ((size_t)sp * nct * 8 + col) * 2;
        s[7] += bb[0];
        s[8] += bb[1];
      }
    } else {
      s[7] += a[7];                              // |P Im(x0) - 0|^2
    }
  }
}

__global__ void __launch_bounds__(256) spm_reduce_stage1(admm_spm_dims d, admm_spm_buffers b) {
  pdl_prologue();
  __shared__ double scratch[10 * 32];
  const int per = (d.nb + gridDim.x - 1) / gridDim.x;
  const int p0 = per * blockIdx.x, p1 = min(d.nb, p0 + per);
  double v[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) v[i] = 0.0;
  for (int prob = p0 + threadIdx.x; prob < p1; prob += blockDim.x) {
    double s[10];
    gather_problem(d, b, prob, s);
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] += s[i];
  }
  block_sum<10>(v, scratch);
  if (threadIdx.x < 10) b.gpart[blockIdx.x * 16 + threadIdx.x] = v[threadIdx.x];
}

__global__ void __launch_bounds__(256) spm_reduce_stage2(int nparts, admm_spm_buffers b) {
  pdl_prologue();
  __shared__ double scratch[10 * 32];
  double v[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) v[i] = 0.0;
  for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] += b.gpart[p * 16 + i];
  }
  block_sum<10>(v, scratch);
  if (threadIdx.x < 10) b.gsum[threadIdx.x] = v[threadIdx.x];
}

__device__ __forceinline__ double mu_step(double mu, double primal, double dual, const admm_spm_buffers& b) {
  if (primal > b.th_change * dual) mu *= b.fact_incr;
  if (dual > b.th_change * primal) mu /= b.fact_incr;
  return fmin(mu, b.max_mu);
}

// residual() / check_convergence() / update_mu() of ONE problem from its ten squared norms
__device__ __forceinline__ void decide_one(const admm_spm_dims& d, const admm_spm_buffers& b, int prob, const double (&s)[10],
                                           int do_update_mu) {
  const double mu10 = b.mu10[prob], mu20 = b.mu20[prob];
  const double p10 = sqrt(s[0]), nx0 = sqrt(s[1]), nx1 = sqrt(s[2]), nd = sqrt(s[3]), nxo = sqrt(s[4]);
  const double nPd = sqrt(s[5]), nPxo = sqrt(s[6]), p20 = sqrt(s[7]), nx2 = sqrt(s[8]), nPx0 = sqrt(s[9]);
  const double d10 = mu10 * nd, d20 = mu20 * nPd;
  const double primal = p10 + p20, dual = d10 + d20;
  b.last_res[2 * prob] = primal;
  b.last_res[2 * prob + 1] = dual;
  const int it = b.iters[prob];
  b.iters[prob] = it + 1;
  b.mu20_used[prob] = mu20;
  if (prob == 0) {
    if (b.history && it < b.hist_cap) {
      b.history[2 * it] = primal;
      b.history[2 * it + 1] = dual;
    }
    b.iter_counter[0] = it + 1;
  }
  // check_convergence (optimizer.py:232-249): 0/0 -> NaN -> not converged
  const bool conv = (p10 / fmax(nx0, nx1) < b.rtol) && (d10 / fmax(mu10 * nx0, mu10 * nxo) < b.rtol) &&
                    (p20 / fmax(nPx0, nx2) < b.rtol) && (d20 / fmax(mu20 * nPx0, mu20 * nPxo) < b.rtol);
  if (conv) {
    b.done[prob] = 1;
    atomicAdd(&b.flags[1], 1);
    return;
  }
  if (do_update_mu) {
    const double m10 = mu_step(mu10, p10, d10, b), m20 = mu_step(mu20, p20, d20, b);
    if (m10 != mu10 || m20 != mu20) {
      b.mu10[prob] = m10;
      b.mu20[prob] = m20;
      b.flags[0] = 1;
    }
  }
}

// nparts > 0 (batch-wide, unsharded): gsum has not been formed yet -- every CTA adds the nparts
// stage-1 partials itself, in the same fixed order (saves the stage-2 launch of a short iteration).
__global__ void __launch_bounds__(128) spm_decide_kernel(admm_spm_dims d, admm_spm_buffers b, int do_update_mu, int nparts) {
  pdl_prologue();
  __shared__ double gs[10];
  if (d.batch_wide && b.lazy != nullptr && __ldcg(b.lazy + 2) != 0) return;      // converged in a lazy iteration
  if (nparts > 0) {
    if (threadIdx.x < 10) {
      double a = 0.0;
      for (int p = 0; p < nparts; ++p) a += b.gpart[p * 16 + threadIdx.x];
      gs[threadIdx.x] = a;
      if (blockIdx.x == 0) b.gsum[threadIdx.x] = a;
    }
    __syncthreads();
  }
  const int prob = blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= d.nb) return;
  if (b.done[prob]) return;
  double s[10];
  if (d.batch_wide) {
#pragma unroll
    for (int i = 0; i < 10; ++i) s[i] = nparts > 0 ? gs[i] : b.gsum[i];
  } else {
    gather_problem(d, b, prob, s);
  }
  decide_one(d, b, prob, s, do_update_mu);
}

// ---------------------------------------------------------------------------------------------
// sharded batch, batch-wide criterion: one-shot all-reduce of the ten sums over peer-mapped memory
// ---------------------------------------------------------------------------------------------
// Stage 1 of the batch-wide sum; the CTA that finishes last adds the partials in a fixed order and pushes the
// ten sums to every rank's mailbox (its own included) under sequence number ctrl[0] + 1.
__global__ void __launch_bounds__(256) spm_reduce_post_kernel(admm_spm_dims d, admm_spm_buffers b, admm_peer_comm c) {
  pdl_prologue();
  if (b.lazy != nullptr && __ldcg(b.lazy + 2) != 0) return;      // converged in a lazy iteration (on every rank alike)
  __shared__ double scratch[10 * 32];
  __shared__ double gs[10];
  __shared__ int last;
  const int tid = threadIdx.x;
  const int per = (d.nb + gridDim.x - 1) / gridDim.x;
  const int p0 = per * blockIdx.x, p1 = min(d.nb, p0 + per);
  double v[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) v[i] = 0.0;
  for (int prob = p0 + tid; prob < p1; prob += blockDim.x) {
    double s[10];
    gather_problem(d, b, prob, s);
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] += s[i];
  }
  block_sum<10>(v, scratch);
  if (tid < 10) {
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i)
      if (tid == i) mine = v[i];
    __stcg(b.gpart + blockIdx.x * 16 + tid, mine);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = (atomicAdd(c.ctrl + 1, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  double w[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) w[i] = 0.0;
  for (int p = tid; p < (int)gridDim.x; p += blockDim.x) {
#pragma unroll
    for (int i = 0; i < 10; ++i) w[i] += __ldcg(b.gpart + p * 16 + i);
  }
  block_sum<10>(w, scratch);
  const unsigned seq = c.ctrl[0] + 1u;
  if (tid < 10) {
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i)
      if (tid == i) mine = w[i];
    gs[tid] = mine;
  }
  __syncthreads();
  peer_post(c, gs, seq);
  if (tid == 0) {
    c.ctrl[1] = 0u;
    c.ctrl[0] = seq;
  }
}

// residual() / check_convergence() / update_mu() on the sums of ALL ranks: wait for the `world` posts of sequence
// ctrl[0] in the local mailbox, add them in rank order (identical totals and decisions on every rank), decide.
__global__ void __launch_bounds__(128) spm_decide_peer_kernel(admm_spm_dims d, admm_spm_buffers b, admm_peer_comm c,
                                                              int do_update_mu) {
  pdl_prologue();
  if (b.lazy != nullptr && __ldcg(b.lazy + 2) != 0) return;      // converged in a lazy iteration (on every rank alike)
  __shared__ unsigned w32[ADMM_MAX_PEERS * PEER_WORDS];
  __shared__ double gs[10];
  __shared__ int fail;
  const int tid = threadIdx.x;
  const unsigned seq = c.ctrl[0];
  const unsigned long long* box = c.mbox[c.rank] + (size_t)(seq & 1u) * ADMM_MAX_PEERS * PEER_SLOT;
  if (tid == 0) fail = 0;
  __syncthreads();
  for (int i = tid; i < PEER_WORDS * c.world; i += blockDim.x) {
    const int src = i / PEER_WORDS, k = i - src * PEER_WORDS;
    const unsigned long long* p = box + src * PEER_SLOT + k;
    const long long t0 = clock64();
    unsigned long long v;
    while (true) {
      v = ld_relaxed_sys_u64(p);
      if ((unsigned)(v >> 32) == seq) break;
      if (clock64() - t0 > PEER_TIMEOUT) {       // a peer died: give up after ~5 s instead of hanging the GPU
        fail = 1;
        break;
      }
    }
    w32[i] = (unsigned)v;
  }
  __syncthreads();
  if (fail) {
    if (tid == 0) b.flags[2] = -2;
    return;
  }
  if (tid < 10) {
    double a = 0.0;
    for (int r = 0; r < c.world; ++r)
      a += __hiloint2double((int)w32[r * PEER_WORDS + 2 * tid + 1], (int)w32[r * PEER_WORDS + 2 * tid]);
    gs[tid] = a;
    if (blockIdx.x == 0) b.gsum[tid] = a;
  }
  __syncthreads();
  const int prob = blockIdx.x * blockDim.x + tid;
  if (prob >= d.nb) return;
  if (b.done[prob]) return;
  double s[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) s[i] = gs[i];
  decide_one(d, b, prob, s, do_update_mu);
}

// the pending decision of the last lazy iteration (head logic alone)
__global__ void __launch_bounds__(128) spm_lazy_flush_kernel(admm_spm_dims d, admm_spm_buffers b, admm_peer_comm c) {
  pdl_prologue();
  lazy_head(d, b, c, 1);
}

// ---------------------------------------------------------------------------------------------
// solo: a handful of problems, each on its own thread-block cluster, the WHOLE solve in one launch
// ---------------------------------------------------------------------------------------------
// A single SpM problem (spm.ipynb) has 2 x 156 kflop of work per iteration: launch latency and the
// round trips through L2 between the kernels of an iteration are all that is left to optimise.  Here
// the CS CTAs of a cluster keep one problem resident for the whole solve:
//   * every CTA owns Nw/CS sampling points: its rows of P and of the implicit (h20, x2) state stay in
//     shared memory;
//   * the L-space work (cached inverse, KKT correction, P^T P x0, soft threshold, dual ascent, norms,
//     residual()/check_convergence()/update_mu(), even the re-inversion after a change of mu) is done
//     redundantly by every CTA -- same inputs, same order, bit-identical results, identical decisions;
//   * the only exchange per iteration is the partial V = P^T|s'| (L doubles) and two norm partials,
//     pulled from the peers' shared memory (DSMEM) after ONE cluster barrier.
// Warps 0-3 hold the L-vectors in registers (thread = (plane, l)); warps 4-11 own the sampling points.

constexpr int SOLO_THREADS = 384;
constexpr int SOLO_LWARPS = 4;
constexpr int SOLO_RTHREADS = SOLO_THREADS - 32 * SOLO_LWARPS;

struct SoloLayout {      // offsets in doubles into the dynamic shared memory
  int R, P, Gi, PtP, M2, pw, ybuf, rhs, xn, w, Cv, us, ss, xch, vp, nA, nB, nBt, rowk, colk, total;
};
// LP = padded L (d.Lp: 16, 40 or 64): every L-dimension is zero padded to LP so that the inner loops
// have compile-time trip counts; R = sampling points per CTA, padded to a multiple of 32
// regp: P, the state and u live in the registers of the sampling-point threads (needs R <= SOLO_RTHREADS)
__host__ __device__ inline SoloLayout solo_layout(int LP, int Nwp, int cs, int npl, bool regp) {
  SoloLayout o;
  o.R = (((Nwp + cs - 1) / cs) + 31) & ~31;
  int at = 0;
  auto take = [&](int n) { const int r = at; at += (n + 1) & ~1; return r; };
  o.P = take(regp ? 0 : LP * o.R);
  o.Gi = take(LP * LP);
  o.PtP = take(regp ? LP * LP : 0);     // shared-memory variant: P^T P stays in global memory (budget)
  o.M2 = take(LP * LP);
  o.pw = take(LP);
  o.ybuf = take(npl * LP);
  o.rhs = take(npl * LP);
  o.xn = take(npl * LP);
  o.w = take(LP);
  o.Cv = take(LP);
  o.us = take(regp ? 0 : o.R);
  o.ss = take(regp ? 0 : o.R);
  o.xch = take(2 * cs * (LP + 2) + 2);          // [2][cs][LP + 2] receive slots + two mbarriers
  o.vp = take(regp ? (SOLO_THREADS / 32 - SOLO_LWARPS) * LP : 0);
  o.nA = take(SOLO_LWARPS * 8);
  o.nB = take((SOLO_THREADS / 32 - SOLO_LWARPS) * 2);
  o.nBt = take(2);
  o.rowk = take(LP);
  o.colk = take(LP);
  o.total = at;
  return o;
}

// Gi <- (G0 + mu10 I + mu20 PtP)^-1 (in-place Gauss-Jordan on the leading L x L block of the LP-pitched
// array, symmetrised), w = Gi C^T; returns sigma = C w  (objectivefunc.py:89-96,148-157).  All threads of
// the CTA; ends with the shared data visible.
__device__ __forceinline__ double solo_factor(int L, int LP, double mu10, double mu20, const double* __restrict__ G0,
                                              const double* PtPs, double* Gi, double* wv, const double* Cv, double* rowk,
                                              double* colk, int* bad) {
  const int tid = threadIdx.x;
  for (int idx = tid; idx < L * L; idx += SOLO_THREADS) {
    const int i = idx / L, j = idx - i * L;
    Gi[i * LP + j] = G0[(size_t)i * LP + j] + (i == j ? mu10 : 0.0) + mu20 * PtPs[i * LP + j];
  }
  __syncthreads();
  for (int k = 0; k < L; ++k) {
    const double pv = Gi[k * LP + k];
    if (!(pv > 0.0)) *bad = k + 1;
    const double ip = 1.0 / pv;
    for (int j = tid; j < L; j += SOLO_THREADS) {
      rowk[j] = (j == k ? 1.0 : Gi[k * LP + j]) * ip;
      colk[j] = (j == k ? 0.0 : Gi[j * LP + k]);
    }
    __syncthreads();
    for (int idx = tid; idx < L * L; idx += SOLO_THREADS) {
      const int i = idx / L, j = idx - i * L;
      if (i == k) {
        Gi[i * LP + j] = rowk[j];
      } else {
        const double base = (j == k) ? 0.0 : Gi[i * LP + j];
        Gi[i * LP + j] = base - colk[i] * rowk[j];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < L * L; idx += SOLO_THREADS) {
    const int i = idx / L, j = idx - i * L;
    if (i < j) {
      const double v = 0.5 * (Gi[i * LP + j] + Gi[j * LP + i]);
      Gi[i * LP + j] = v;
      Gi[j * LP + i] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < L; i += SOLO_THREADS) {
    double a = 0.0;
    for (int j = 0; j < L; ++j) a += Gi[i * LP + j] * Cv[j];
    wv[i] = a;
  }
  __syncthreads();
  double sigma = 0.0;
  for (int i = 0; i < L; ++i) sigma += Cv[i] * wv[i];
  return sigma;
}

// out = sum_j M[j * LP + l] * v[j]  (M symmetric, LP-pitched, zero padded; v zero padded): four chains
template <int LP>
__device__ __forceinline__ double solo_matvec(const double* __restrict__ Mcol, const double* __restrict__ v) {
  double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < LP; j += 4) {
    const double2 v01 = *reinterpret_cast<const double2*>(v + j);
    const double2 v23 = *reinterpret_cast<const double2*>(v + j + 2);
    a[0] += Mcol[(j + 0) * LP] * v01.x;
    a[1] += Mcol[(j + 1) * LP] * v01.y;
    a[2] += Mcol[(j + 2) * LP] * v23.x;
    a[3] += Mcol[(j + 3) * LP] * v23.y;
  }
  return (a[0] + a[1]) + (a[2] + a[3]);
}

// Sum each of W (power of two <= 32) per-lane values over the 32 lanes of the warp with W - 1 + (5 - log2 W)
// shuffles instead of 5 W: every halving step trades half of the values with the partner lane.  Afterwards
// v[0] of lane `lane` is the complete sum of value index  lane >> (5 - log2 W).
template <int W>
__device__ __forceinline__ void solo_treduce(double (&v)[W], int lane) {
  int o = 16;
#pragma unroll
  for (int n = W / 2; n >= 1; n /= 2, o /= 2) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const double send = up ? v[k] : v[k + n];
      const double keep = up ? v[k + n] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o >= 1; o /= 2) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

#ifdef SOLO_TRACE   // tools only: clock64 stamps of warps 0 and 4 of CTA 0 for iterations 10..13 into b.gpart
#define SOLO_STAMP(idx)                                                                                   \
  if (crank == 0 && lane == 0 && (warp == 0 || warp == SOLO_LWARPS) && k >= 10 && k < 14)                 \
    reinterpret_cast<long long*>(b.gpart)[((k - 10) * 2 + (warp != 0)) * 16 + (idx)] = clock64();
#else
#define SOLO_STAMP(idx)
#endif

// M2T[j][l] = (P^T P G^-1)[l][j],  pw = P^T P w:  y = P^T P x0 = M2 rhs + nu pw comes out of the same pass over
// rhs as x0 = G^-1 rhs + nu w (one dependent matvec per iteration instead of two).  All threads; ends synchronised.
template <int LP>
__device__ __forceinline__ void solo_m2(const double* PtPs, const double* Gi, const double* wv, double* M2T, double* pw) {
  for (int idx = threadIdx.x; idx < LP * LP; idx += SOLO_THREADS) {
    const int j = idx / LP, l = idx - j * LP;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
    for (int k = 0; k < LP; k += 2) {
      a0 += PtPs[k * LP + l] * Gi[k * LP + j];
      a1 += PtPs[(k + 1) * LP + l] * Gi[(k + 1) * LP + j];
    }
    M2T[idx] = a0 + a1;
  }
  for (int l = threadIdx.x; l < LP; l += SOLO_THREADS) {
    double a = 0.0;
    for (int k = 0; k < LP; ++k) a += PtPs[k * LP + l] * wv[k];
    pw[l] = a;
  }
  __syncthreads();
}

// Same over the 16 lanes that share bit 0 of the lane index (offsets 16, 8, 4, 2): v[0] of lane `lane` is the sum
// of value index  (lane >> (5 - log2 W)) & (W - 1)  over those lanes.
template <int W>
__device__ __forceinline__ void solo_treduce_pairs(double (&v)[W], int lane) {
  int o = 16;
#pragma unroll
  for (int n = W / 2; n >= 1; n /= 2, o /= 2) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const double send = up ? v[k] : v[k + n];
      const double keep = up ? v[k + n] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o >= 2; o /= 2) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

template <int CS, int LP, bool REGP, bool BW>   // BW: batch-wide criterion over several co-resident clusters
__global__ void __launch_bounds__(SOLO_THREADS, 1)
    spm_solo_kernel(admm_spm_dims d, admm_spm_buffers b, const double* __restrict__ G0, const double* __restrict__ PtP,
                    int budget, int interval) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CS > 1 ? (int)cluster.block_rank() : 0;
  const int prob = blockIdx.x / CS;
  if (b.done[prob]) return;                                 // uniform over the cluster
  constexpr int NT = LP / 8, NW = SOLO_THREADS / 32, NRW = NW - SOLO_LWARPS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = d.L, npl = d.nplanes;
  const int pt = prob >> 3, g = prob & 7;
  const int Nwp = d.nrt * 8;
  extern __shared__ __align__(16) double sm[];
  const SoloLayout lay = solo_layout(LP, Nwp, CS, npl, REGP);
  const int R = lay.R;                                      // sampling points per CTA (padded)
  const int row0 = crank * R, nrow = max(0, min(R, Nwp - row0));
  double* Psm = sm + lay.P;        // [l][i]: P[row0 + i][l]
  double* Gi = sm + lay.Gi;        // [LP][LP], symmetric
  // P^T P ([LP][LP], symmetric, canonical): only the set-up, the re-inversion after a mu change and the first y0 read it
  // (P^T P x0 comes from M2T in the loop).  Register variant: a shared-memory copy; shared-memory variant (the budget is
  // tight there): straight from global memory.
  const double* PtPs = REGP ? (sm + lay.PtP) : PtP;
  double* M2T = sm + lay.M2;       // [LP][LP]: (P^T P G^-1) transposed
  double* pw = sm + lay.pw;        // [LP]: P^T P w
  double* ybuf = sm + lay.ybuf;    // [plane][LP]: (P^T P G^-1) rhs, computed by the sampling-point warps during the x-update
  double* rhs = sm + lay.rhs;      // [plane][LP]
  double* xn = sm + lay.xn;        // [plane][LP]: the new x0
  double* wv = sm + lay.w;
  double* Cv = sm + lay.Cv;
  double* us = sm + lay.us;        // [i]: operand of V = P^T u
  double* ss = sm + lay.ss;        // [i]: implicit state  s = Re h20 - mu20 x2
  double* xch = sm + lay.xch;      // [2][CS][LP + 2]: partial V and norm partials pushed by every CTA of the cluster
  uint64_t* xbar = reinterpret_cast<uint64_t*>(xch + 2 * CS * (LP + 2));   // [2]
  double* vp = sm + lay.vp;        // [row warp][LP]: per-warp partial V (REGP)
  double* nA = sm + lay.nA;        // [L-space warp][8]
  double* nB = sm + lay.nB;        // [row warp][2]
  double* nBt = sm + lay.nBt;      // [2] cluster totals
  __shared__ int bad_sh;
  __shared__ double gtot[10];
  int gpar = 0, gtarget = 0;       // batch-wide all-reduce over the clusters: buffer parity, expected arrivals

  // ---- one-time loads (everything zero padded)
  if (tid == 0) bad_sh = 0;
  // sampling-point thread rt owns point row0 + rt (REGP: its row of P, s and u never leave the registers)
  const int rt = tid - 32 * SOLO_LWARPS;
  // Kreg: row of P (sampling-point threads) or column l of the cached inverse (L-space threads): ONE array,
  // the two roles never meet in a thread
  double Kreg[REGP ? LP : 1], s_reg = 0.0, u_reg = 0.0;
  if (REGP) {
#pragma unroll
    for (int j = 0; j < (REGP ? LP : 1); ++j) Kreg[j] = 0.0;
    // two neighbouring threads share two rows: thread (pair i, half h) keeps columns [h LP/2, (h+1) LP/2) of rows
    // 2i and 2i+1 -- Kreg[k] = P[2i][c0 + k], Kreg[LP/2 + k] = P[2i+1][c0 + k] -- so that the partial V of the pair is
    // summed in registers before the warp reduction, which then runs over 16 lanes and LP/2 columns only
    if (rt >= 0 && (rt & ~1) < nrow) {
      const int ra = row0 + (rt & ~1), c0 = (rt & 1) * (LP / 2);
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        const int row = ra + r2;
#pragma unroll
        for (int k2 = 0; k2 < (REGP ? LP / 2 : 0); ++k2) Kreg[REGP ? r2 * (LP / 2) + k2 : 0] = pf_elem(d, b.Pf, row, c0 + k2);
      }
      const int row = row0 + rt;
      s_reg = b.S[state_elem(d, pt, g, row)];
    }
  } else {
    for (int idx = tid; idx < LP * R; idx += SOLO_THREADS) {
      const int l = idx / R, i = idx - l * R;
      const int row = row0 + i;
      double v = 0.0;
      if (i < nrow) v = pf_elem(d, b.Pf, row, l);
      Psm[idx] = v;
    }
  }
  const int slot = b.slot[prob];
  for (int idx = tid; idx < LP * LP; idx += SOLO_THREADS) {
    const int i = idx / LP, j = idx - i * LP;
    const size_t o = bfrag_of(NT, i, j);
    Gi[idx] = b.Ginv_cache[(size_t)slot * LP * LP + o];
    if (REGP) sm[lay.PtP + idx] = PtP[idx];
  }
  for (int i = tid; i < LP; i += SOLO_THREADS) {
    wv[i] = b.w_cache[(size_t)slot * LP + i];
    Cv[i] = b.Cvec[i];
  }
  for (int i = tid; i < npl * LP; i += SOLO_THREADS) rhs[i] = xn[i] = 0.0;
  if (!REGP) {
    for (int i = tid; i < R; i += SOLO_THREADS) {
      const int row = row0 + i;
      ss[i] = i < nrow ? b.S[state_elem(d, pt, g, row)] : 0.0;
      us[i] = 0.0;
    }
  }
  double sigma = b.sigma_cache[slot];
  double mu10 = b.mu10[prob], mu20 = b.mu20[prob];
  double mu20_enc = b.mu20_used[prob];            // the mu20 the negative part of s is scaled with
  int it = b.iters[prob];
  // receive barriers: armed for the first use of each buffer; nobody pushes before every CTA is set up
  constexpr unsigned XBYTES = CS * (LP + 2) * sizeof(double);
  if (tid == 0) {
    mbar_init(xbar, 1);
    mbar_init(xbar + 1, 1);
    fence_barrier_init();
    mbar_expect_tx(xbar, XBYTES);
    mbar_expect_tx(xbar + 1, XBYTES);
  }
  cluster.sync();
  solo_m2<LP>(PtPs, Gi, wv, M2T, pw);
  unsigned xpar = 0u;      // bit p: parity of the next completion of receive barrier p
  const unsigned xch_u32 = smem_u32(xch), xbar_u32 = smem_u32(xbar);
  // push value `v` of entry `e` (column or norm slot) of exchange buffer `ph` to every CTA of the cluster
  auto push_all = [&](int ph, int e, double v) {
    const unsigned slot = xch_u32 + ((ph * CS + crank) * (LP + 2) + e) * (unsigned)sizeof(double);
    const unsigned bar = xbar_u32 + ph * (unsigned)sizeof(uint64_t);
#pragma unroll
    for (int c = 0; c < CS; ++c) st_async_f64(mapa_u32(slot, c), v, mapa_u32(bar, c));
  };

  // L-space threads: (plane, l) and their vector elements, in registers for the whole solve
  const bool lth = warp < SOLO_LWARPS;
  const int pl = tid >> 6, l = tid & 63;
  const bool lact = lth && pl < npl && l < L;
  const size_t fo = lact ? frag_index(pt * npl + pl, NT, l >> 3, 4 * g + ((l & 7) >> 1)) + (l & 1) : 0;
  double r_b0 = 0.0, r_h10 = 0.0, r_x1 = 0.0, r_x0 = 0.0, r_y0 = 0.0, r_V = 0.0, r_aim = 0.0, Dp = 0.0;
  double r_xold = 0.0;         // x0 at the start of the last executed iteration (`_x_old[0]`)
  bool ran_any = false;
  if (lact) {
    r_b0 = b.b0[fo];
    r_h10 = b.h10[fo];
    r_x1 = b.x1[fo];
    r_x0 = b.x0[fo];
    r_aim = b.aim[fo];
    if (pl == 1) r_V = b.V[fo];                   // z = P^T Im(h20): state, not derived
    Dp = b.Dre[(size_t)pl * 8 * d.npt + prob];
    xn[pl * LP + l] = r_x0;
  }
  __syncthreads();
  if (lact) r_y0 = solo_matvec<LP>(PtPs + l, xn + pl * LP);     // y0 = P^T P x0
  auto load_gcol = [&]() {                                      // column l of the cached inverse
    if (REGP && lth) {
#pragma unroll
      for (int j = 0; j < (REGP ? LP : 1); ++j) Kreg[j] = lact ? Gi[j * LP + l] : 0.0;
    }
  };
  load_gcol();

  int phase = 0;
  bool need_v = true;          // V (real plane) has to be rebuilt from the state (start, change of mu20)
  bool mu_changed = false;
  int conv = 0;
  // uniform reciprocals (FP64 division and square root are long instruction sequences on the pipe the
  // whole CTA shares: they are kept out of the per-iteration path)
  double inv_mu10 = 1.0 / mu10, inv_sigma = 1.0 / sigma, inv_mu20sq = 1.0 / (mu20 * mu20), thr = 0.5 * b.lam / mu10;
  const double rtol2 = b.rtol * b.rtol;
  double sq[4] = {0.0, 0.0, 0.0, 0.0};     // |x0-x1|^2, |P x0-x2|^2, |dx0|^2, |P dx0|^2 of the last iteration
  double mu10_res = mu10, mu20_res = mu20; // the mu they are to be scaled with
  bool hist_pending = false;
  int hist_it = 0;
  int until_upd = interval > 0 ? (interval - it % interval) % interval : -1;   // iterations until the next update_mu

  // partial V of my sampling points from us[], exchange, total into r_V of the (0, l) threads; the two
  // norm partials ride along.  Warp w sums the columns l = w, w + NW, ... (at most 4 for L <= 48, else 6).
  constexpr int NQ = (LP + NW - 1) / NW;
  int k = 0;
  auto exchange = [&](bool with_norms) {
    const double* slots = xch + phase * CS * (LP + 2);
    if (REGP) {
      // per-warp partial V: the pair's two rows are combined in registers, then transposed warp reductions over the
      // 16 lanes of equal half (16, 8 or 4 columns at a time)
      if (!lth) {
        constexpr int H = LP / 2;
        double* vw = vp + (warp - SOLO_LWARPS) * LP + (lane & 1) * H;
        const bool hi = (lane & 1) != 0;
        const double uo = __shfl_xor_sync(0xffffffffu, u_reg, 1);
        const double ua = hi ? uo : u_reg, ub = hi ? u_reg : uo;
#pragma unroll
        for (int base = 0; base < (REGP ? H : 0); base += 16) {
          if (H - base >= 16) {
            double v[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) v[k2] = ua * Kreg[REGP ? base + k2 : 0] + ub * Kreg[REGP ? H + base + k2 : 0];
            solo_treduce_pairs<16>(v, lane);
            vw[base + ((lane >> 1) & 15)] = v[0];
          } else if (H - base == 8) {
            double v[8];
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) v[k2] = ua * Kreg[REGP ? base + k2 : 0] + ub * Kreg[REGP ? H + base + k2 : 0];
            solo_treduce_pairs<8>(v, lane);
            if ((lane & 2) == 0) vw[base + ((lane >> 2) & 7)] = v[0];
          } else {
            double v[4];
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) v[k2] = ua * Kreg[REGP ? base + k2 : 0] + ub * Kreg[REGP ? H + base + k2 : 0];
            solo_treduce_pairs<4>(v, lane);
            if ((lane & 6) == 0) vw[base + ((lane >> 3) & 3)] = v[0];
          }
        }
      }
      __syncthreads();
      if (tid < LP) {
        double a = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NRW; ++w2) a += vp[w2 * LP + tid];
        push_all(phase, tid, a);
      }
    } else {
      double a[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) a[q] = 0.0;
      const double* pc = Psm + min(warp, LP - 1) * R + lane;   // (clamped: unused columns are skipped below)
#pragma unroll 4
      for (int i = 0; i < R; i += 32) {
        const double u = us[i + lane];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          if (warp + q * NW < LP) a[q] += u * pc[q * NW * R + i];
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) a[q] = warp_sum(a[q]);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          if (warp + q * NW < LP) push_all(phase, warp + q * NW, a[q]);
      }
    }
    if (tid >= 64 && tid < 66) {
      double a = 0.0;
      if (with_norms) {
#pragma unroll
        for (int w2 = 0; w2 < NRW; ++w2) a += nB[w2 * 2 + tid - 64];
        a *= inv_mu20sq;
      }
      push_all(phase, LP + tid - 64, a);
    }
    SOLO_STAMP(6)
    mbar_wait_cluster(xbar + phase, (xpar >> phase) & 1u);
    xpar ^= 1u << phase;
    SOLO_STAMP(7)
    if (lact && pl == 0) {
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < CS; ++c) a += slots[c * (LP + 2) + l];
      r_V = a;
    }
    if (warp == SOLO_LWARPS && lane < 2) {
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < CS; ++c) a += slots[c * (LP + 2) + LP + lane];
      nBt[lane] = a;
    }
    SOLO_STAMP(8)
    __syncthreads();
    // everybody has read this buffer: re-arm its barrier for the exchange after the next (no peer can push
    // into it before it has received this CTA's NEXT push, which comes later in program order)
    if (tid == 0) mbar_expect_tx(xbar + phase, XBYTES);
    phase ^= 1;
    SOLO_STAMP(9)
  };

  for (k = 0;; ++k) {
    if (need_v) {
      // u = Re h20 + mu20 x2 with x2 decoded by the mu20 it was encoded with
      const double ratio = mu20 / mu20_enc;
      if (REGP) {
        u_reg = is_neg(s_reg) ? -s_reg * ratio : s_reg;
      } else {
        for (int i = tid; i < R; i += SOLO_THREADS) {
          const double s = ss[i];
          us[i] = is_neg(s) ? -s * ratio : s;
        }
        __syncthreads();
      }
      exchange(false);
      need_v = false;
    }
    if (k >= budget) break;
    SOLO_STAMP(0)

    // ---- term 0: rhs, cached inverse, KKT correction (C Gi rhs = w . rhs because Gi is symmetric)
    if (lact) rhs[pl * LP + l] = r_b0 + r_h10 + mu10 * r_x1 + r_V;
    __syncthreads();
    SOLO_STAMP(1)
    if (hist_pending && warp == NW - 1) {
      // residual() of the previous iteration (this warp idles while the L-space warps solve term 0)
      if (crank == 0 && lane == 0 && prob == 0 && b.history && hist_it < b.hist_cap) {
        b.history[2 * hist_it] = sqrt(sq[0]) + sqrt(sq[1]);
        b.history[2 * hist_it + 1] = mu10_res * sqrt(sq[2]) + mu20_res * sqrt(sq[3]);
      }
    }
    hist_pending = false;
    double xv = 0.0, nu_keep = 0.0;
    if (!lth) {
      // the sampling-point warps idle during the x-update: they take the third dot product of every (plane, l)
      if (rt < npl * LP) {
        const int p3 = rt / LP, l3 = rt - p3 * LP;
        ybuf[rt] = solo_matvec<LP>(M2T + l3, rhs + p3 * LP);
      }
    }
    if (lact) {
      const double* rp = rhs + pl * LP;
      double xi;
      if (REGP) {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < LP; j += 4) {
          const double2 r01 = *reinterpret_cast<const double2*>(rp + j), r23 = *reinterpret_cast<const double2*>(rp + j + 2);
          a[0] += Kreg[REGP ? j : 0] * r01.x;
          a[1] += Kreg[REGP ? j + 1 : 0] * r01.y;
          a[2] += Kreg[REGP ? j + 2 : 0] * r23.x;
          a[3] += Kreg[REGP ? j + 3 : 0] * r23.y;
        }
        xi = (a[0] + a[1]) + (a[2] + a[3]);
      } else {
        xi = solo_matvec<LP>(Gi + l, rp);
      }
      double c[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < LP; j += 4) {
        const double2 r01 = *reinterpret_cast<const double2*>(rp + j), r23 = *reinterpret_cast<const double2*>(rp + j + 2);
        const double2 w01 = *reinterpret_cast<const double2*>(wv + j), w23 = *reinterpret_cast<const double2*>(wv + j + 2);
        c[0] += w01.x * r01.x;
        c[1] += w01.y * r01.y;
        c[2] += w23.x * r23.x;
        c[3] += w23.y * r23.y;
      }
      const double nu = (Dp - ((c[0] + c[1]) + (c[2] + c[3]))) * inv_sigma;
      xv = xi + wv[l] * nu;
      xn[pl * LP + l] = xv;
      nu_keep = nu;
    }
    SOLO_STAMP(2)
    __syncthreads();
    SOLO_STAMP(3)

    if (lth) {
      // ---- y = P^T P x0, norms, L1 z-update, dual ascent of pair (1,0), imaginary-plane recursion
      double n[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (lact) {
        const double y = ybuf[pl * LP + l] + pw[l] * nu_keep;      // P^T P x0 = (P^T P G^-1) rhs + nu P^T P w
        const double dd = xv - r_x0;
        n[3] = dd * dd;
        n[4] = r_x0 * r_x0;
        n[5] = dd * (y - r_y0);
        n[6] = r_x0 * r_y0;
        n[7] = xv * y;
        double z = 0.0;
        if (pl == 0) {
          const double yv = -((r_h10 - mu10 * xv) * inv_mu10);
          if (yv > thr) z = yv - thr;
          if (yv < -thr) z = yv + thr;
        } else {
          r_V -= mu20 * y;
          r_aim += mu20 * xv;
        }
        r_h10 += mu10 * (z - xv);
        n[0] = (xv - z) * (xv - z);
        n[1] = xv * xv;
        n[2] = z * z;
        r_xold = r_x0;
        ran_any = true;
        r_x0 = xv;
        r_x1 = z;
        r_y0 = y;
      }
      solo_treduce<8>(n, lane);
      if ((lane & 3) == 0) nA[warp * 8 + (lane >> 2)] = n[0];
    } else {
      // ---- my sampling points: s' = Re h20 - mu20 (P Re x0) encodes dual ascent and projection
      double n_dh = 0.0, n_xm = 0.0;
      if (REGP) {
        // both rows of the pair over my half of the columns, then one exchange with the partner
        constexpr int H = LP / 2;
        const bool hi = (lane & 1) != 0;
        const double* xh = xn + (hi ? H : 0);
        double qa[2] = {0.0, 0.0}, qb[2] = {0.0, 0.0};
#pragma unroll
        for (int k2 = 0; k2 < H; k2 += 2) {
          const double2 x2 = *reinterpret_cast<const double2*>(xh + k2);
          qa[0] += Kreg[REGP ? k2 : 0] * x2.x;
          qa[1] += Kreg[REGP ? k2 + 1 : 0] * x2.y;
          qb[0] += Kreg[REGP ? H + k2 : 0] * x2.x;
          qb[1] += Kreg[REGP ? H + k2 + 1 : 0] * x2.y;
        }
        const double sa = qa[0] + qa[1], sb = qb[0] + qb[1];
        const double qd = (hi ? sb : sa) + __shfl_xor_sync(0xffffffffu, hi ? sa : sb, 1);
        const double hre = is_neg(s_reg) ? 0.0 : s_reg;
        const double s_new = hre - mu20 * qd;
        const bool neg = is_neg(s_new);
        const double hnew = neg ? 0.0 : s_new;
        const double xm = neg ? s_new : 0.0;
        const double dh = hre - hnew;
        n_dh = dh * dh;
        n_xm = xm * xm;
        u_reg = fabs(s_new);
        s_reg = s_new;
      }
      for (int i = REGP ? R : rt; i < R; i += SOLO_RTHREADS) {
        double q[4] = {0.0, 0.0, 0.0, 0.0};
        const double* pr = Psm + i;
#pragma unroll
        for (int j = 0; j < LP; j += 4) {
          const double2 x01 = *reinterpret_cast<const double2*>(xn + j), x23 = *reinterpret_cast<const double2*>(xn + j + 2);
          q[0] += pr[(j + 0) * R] * x01.x;
          q[1] += pr[(j + 1) * R] * x01.y;
          q[2] += pr[(j + 2) * R] * x23.x;
          q[3] += pr[(j + 3) * R] * x23.y;
        }
        const double s = ss[i];
        const double hre = is_neg(s) ? 0.0 : s;
        const double s_new = hre - mu20 * ((q[0] + q[1]) + (q[2] + q[3]));
        const bool neg = is_neg(s_new);
        const double hnew = neg ? 0.0 : s_new;
        const double xm = neg ? s_new : 0.0;
        const double dh = hre - hnew;
        n_dh += dh * dh;
        n_xm += xm * xm;
        us[i] = fabs(s_new);
        ss[i] = s_new;
      }
      n_dh = warp_sum(n_dh);
      n_xm = warp_sum(n_xm);
      if (lane == 0) {
        nB[(warp - SOLO_LWARPS) * 2] = n_dh;
        nB[(warp - SOLO_LWARPS) * 2 + 1] = n_xm;
      }
    }
    mu20_enc = mu20;
    SOLO_STAMP(4)
    __syncthreads();
    SOLO_STAMP(5)
    exchange(true);
    SOLO_STAMP(10)

    // ---- residual() / check_convergence() / update_mu(): every thread, identical numbers
    double s[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) s[i] = 0.0;
    for (int p2 = 0; p2 < npl; ++p2) {
      double a[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 v0 = *reinterpret_cast<const double2*>(nA + (2 * p2) * 8 + i);
        const double2 v1 = *reinterpret_cast<const double2*>(nA + (2 * p2 + 1) * 8 + i);
        a[i] = v0.x + v1.x;
        a[i + 1] = v0.y + v1.y;
      }
#pragma unroll
      for (int i = 5; i < 8; ++i) a[i] = a[i] > 0.0 ? a[i] : 0.0;
#pragma unroll
      for (int i = 0; i < 7; ++i) s[i] += a[i];
      s[9] += a[7];
      if (p2 == 1) s[7] += a[7];
    }
    s[7] += nBt[0];
    s[8] += nBt[1];
    if (BW) {
      // batch-wide criterion over several clusters (a packed PartialDiagonalMatrix batch of a few problems): all-reduce
      // of the ten squared norms through global memory.  Every cluster publishes its sums and arrives on a counter;
      // all CTAs wait for the nb arrivals of this iteration and add the nb entries in problem order -- identical
      // totals, identical decisions everywhere.  The launcher guarantees that all nb clusters are co-resident.
      double* slots = b.gpart + (size_t)gpar * d.nb * 16;
      if (crank == 0) {
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < 10; ++i)
          if (tid == i) mine = s[i];
        if (tid < 10) slots[prob * 16 + tid] = mine;
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(&b.flags[3], 1);
      }
      gtarget += d.nb;
      if (tid == 0) {
        // watchdog: if the clusters are not all co-resident after all (the launcher checks the occupancy), give up
        // after ~2 s instead of hanging the GPU; every CTA of every cluster runs into the same timeout
        const long long t_start = clock64();
        while ((int)ld_acquire_u32(&b.flags[3]) < gtarget) {
          if (clock64() - t_start > 4000000000LL) {
            bad_sh = -1;
            break;
          }
        }
      }
      __syncthreads();
      if (bad_sh < 0) break;
      if (tid < 10) {
        double a = 0.0;
        for (int p2 = 0; p2 < d.nb; ++p2) a += __ldcg(slots + p2 * 16 + tid);
        gtot[tid] = a;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 10; ++i) s[i] = gtot[i];
      gpar ^= 1;
    }
    // check_convergence (optimizer.py:232-249) on the squared norms: p / max(a, b) < rtol  <=>
    // p^2 < rtol^2 max(a^2, b^2)  (mu > 0 cancels in the dual tests; 0/0 and x/0 stay "not converged")
    sq[0] = s[0];
    sq[1] = s[7];
    sq[2] = s[3];
    sq[3] = s[5];
    mu10_res = mu10;
    mu20_res = mu20;
    hist_pending = true;
    hist_it = it;
    ++it;
    const bool cv = (s[0] < rtol2 * fmax(s[1], s[2])) && (s[3] < rtol2 * fmax(s[1], s[4])) &&
                    (s[7] < rtol2 * fmax(s[9], s[8])) && (s[5] < rtol2 * fmax(s[9], s[6]));
    SOLO_STAMP(11)
    if (cv) {
      conv = 1;
      break;
    }
    const bool upd = until_upd == 0;
    until_upd = (upd ? interval : until_upd) - 1;
    if (upd) {
      const double p10 = sqrt(s[0]), p20 = sqrt(s[7]), d10 = mu10 * sqrt(s[3]), d20 = mu20 * sqrt(s[5]);
      const double m10 = mu_step(mu10, p10, d10, b), m20 = mu_step(mu20, p20, d20, b);
      if (m10 != mu10 || m20 != mu20) {
        mu10 = m10;
        mu20 = m20;
        mu_changed = true;
        need_v = true;
        sigma = solo_factor(L, LP, mu10, mu20, G0, PtPs, Gi, wv, Cv, sm + lay.rowk, sm + lay.colk, &bad_sh);
        solo_m2<LP>(PtPs, Gi, wv, M2T, pw);
        load_gcol();
        inv_mu10 = 1.0 / mu10;
        inv_sigma = 1.0 / sigma;
        inv_mu20sq = 1.0 / (mu20 * mu20);
        thr = 0.5 * b.lam / mu10;
      }
    }
  }
  const double primal = sqrt(sq[0]) + sqrt(sq[1]), dual = mu10_res * sqrt(sq[2]) + mu20_res * sqrt(sq[3]);
  if (hist_pending && crank == 0 && tid == 0 && prob == 0 && b.history && hist_it < b.hist_cap) {
    b.history[2 * hist_it] = primal;
    b.history[2 * hist_it + 1] = dual;
  }

  // ---- write the state back in the layouts of the batch kernels
  if (crank == 0 && lact) {
    b.x0[fo] = r_x0;
    if (b.x0_old != nullptr && ran_any) b.x0_old[fo] = r_xold;
    b.x1[fo] = r_x1;
    b.h10[fo] = r_h10;
    b.y0[fo] = r_y0;
    b.aim[fo] = r_aim;
    b.V[fo] = r_V;
    if (pl == 0) {
      const size_t vstride = (size_t)d.npt * npl * NT * 64;
      for (int sp = 1; sp < d.nsplit; ++sp) b.V[sp * vstride + fo] = 0.0;
    }
  }
  if (REGP) {
    if (rt >= 0 && rt < nrow) {
      const int row = row0 + rt;
      b.S[state_elem(d, pt, g, row)] = s_reg;
    }
  } else {
    for (int i = tid; i < nrow; i += SOLO_THREADS) {
      const int row = row0 + i;
      b.S[state_elem(d, pt, g, row)] = ss[i];
    }
  }
  if (crank == 0 && tid == 0) {
    b.mu10[prob] = mu10;
    b.mu20[prob] = mu20;
    b.mu20_used[prob] = mu20_enc;
    if (it != b.iters[prob]) {
      b.last_res[2 * prob] = primal;
      b.last_res[2 * prob + 1] = dual;
    }
    b.iters[prob] = it;
    if (prob == 0) b.iter_counter[0] = it;
    if (conv) {
      b.done[prob] = 1;
      atomicAdd(&b.flags[1], 1);
    }
    if (mu_changed) b.flags[0] = 1;
    if (bad_sh) b.flags[2] = bad_sh;
  }
  if (CS > 1) cluster.sync();      // nobody leaves while a peer may still read its exchange buffer
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static int check_dims(const admm_spm_dims* d, const char* who) {
  ADMM_REQUIRE(d != nullptr, ADMM_EINVAL, "%s: null dims", who);
  ADMM_REQUIRE(d->L >= 1 && (d->Lp == 16 || d->Lp == 40 || d->Lp == 64) && d->Lp >= d->L, ADMM_EUNSUPPORTED,
               "%s: L=%d Lp=%d unsupported (Lp must be 16, 40 or 64 and >= L)", who, d->L, d->Lp);
  ADMM_REQUIRE(d->nrt % PASS_CHUNK_RT == 0 && d->nrt * 8 >= d->Nw && d->Nw >= 1, ADMM_EINVAL,
               "%s: nrt=%d must be a multiple of %d covering Nw=%d", who, d->nrt, PASS_CHUNK_RT, d->Nw);
  ADMM_REQUIRE(d->nb >= 1 && d->npt * 8 >= d->nb, ADMM_EINVAL, "%s: bad nb/npt", who);
  ADMM_REQUIRE(d->nplanes == 1 || d->nplanes == 2, ADMM_EINVAL, "%s: nplanes must be 1 or 2", who);
  ADMM_REQUIRE(d->nc >= 0 && d->nc <= ADMM_SPM_MAX_NC, ADMM_EUNSUPPORTED, "%s: %d constraint rows (at most %d)", who, d->nc,
               ADMM_SPM_MAX_NC);
  ADMM_REQUIRE(d->mt == 1 || (d->mt == 2 && d->Lp <= 40), ADMM_EINVAL, "%s: mt must be 1, or 2 with Lp <= 40", who);
  ADMM_REQUIRE(d->nsplit >= 1 && (d->nbal > 0 || d->nsplit <= d->nrt / PASS_CHUNK_RT), ADMM_EINVAL, "%s: bad nsplit=%d", who,
               d->nsplit);
  if (d->fold) {
    ADMM_REQUIRE(d->fold == 1 && d->Nw % 2 == 0, ADMM_EUNSUPPORTED, "%s: the folded pass needs an even Nw", who);
    ADMM_REQUIRE((d->nrt / 2) * 8 >= d->Nw / 2, ADMM_EINVAL, "%s: nrt=%d does not cover the %d point pairs", who, d->nrt, d->Nw / 2);
  }
  if (d->nbal > 0) {
    const long long T = (long long)ceil_div(d->npt, 4 * d->mt) * (d->nrt / PASS_CHUNK_RT);
    ADMM_REQUIRE(d->nbal <= T, ADMM_EINVAL, "%s: nbal=%d exceeds the %lld group-chunks", who, d->nbal, T);
    const long long kmax = ((long long)(d->nrt / PASS_CHUNK_RT) * d->nbal + T - 1) / T + 1;
    ADMM_REQUIRE(d->nsplit >= kmax, ADMM_EINVAL, "%s: balanced decomposition needs nsplit >= %lld (got %d)", who, kmax,
                 d->nsplit);
  }
  return ADMM_OK;
}

// programmatic dependent launch of the per-iteration kernels (ADMM_NO_PDL=1: plain stream order, for A/B runs)
static bool use_pdl() {
  static const bool on = getenv("ADMM_NO_PDL") == nullptr;
  return on;
}
static bool use_pdl_x() {      // the stand-alone x-update kernel (ADMM_NO_PDL_X=1: A/B runs)
  static const bool on = getenv("ADMM_NO_PDL_X") == nullptr;
  return on && use_pdl();
}

static int ew_grid(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 16)); }

// how kernels whose CTAs wait for each other are launched, per device (see launch_pass_k / launch_solo)
static std::map<int, int>& coop_mode() {
  static std::map<int, int> m;
  return m;
}
static std::map<int, int>& coop_mode_solo() {
  static std::map<int, int> m;
  return m;
}

struct LazyArgs {          // how a pass / step launch takes part in the lazy batch-wide scheme (all zero: classic)
  admm_peer_comm comm;     // world == 0: one rank, the sums go through gsum
  int lazy;                // 0 classic, 1 lazy without, 2 lazy with a pending decision
  int nA;                  // CTAs of the stand-alone x-update kernel (unfused path)
};

template <int NT, int MT, int MODE, int FNP, bool BAL = false, bool FOLD = false>
static int launch_pass_k(const admm_spm_dims* d, const admm_spm_buffers* b, cudaStream_t s, const LazyArgs& lz,
                         int* max_ctas = nullptr) {      // max_ctas != NULL: occupancy query only (co-resident CTAs)
  const dim3 grid = d->nbal > 0 ? dim3(d->nbal, 1) : dim3(ceil_div(d->npt, PASS_WARPS * MT), d->nsplit);
  const size_t smem = PassSmem<NT, MT, FOLD>::BYTES;
  auto k = spm_pass_kernel<NT, MT, MODE, FNP, BAL, FOLD>;
  if (max_ctas != nullptr) {
    static std::map<int, int> cached;      // per instantiation and device
    auto it = cached.find(cur_dev());
    if (it == cached.end()) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      int per_sm = 0, sms = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, PASS_WARPS * 32, smem) != cudaSuccess ||
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev()) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
      }
      it = cached.emplace(cur_dev(), per_sm * sms).first;
    }
    *max_ctas = it->second;
    return ADMM_OK;
  }
  static std::map<int, bool> configured;     // per instantiation and device
  bool& cfgd = configured[cur_dev()];
  if (!cfgd) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cfgd = true;
  }
  if (BAL) {
    // the CTAs of this launch wait for each other (owner CTAs publish x0): cooperative launch, so that the driver
    // guarantees their co-residency whatever else runs on the device.  Probed once per device: cooperative +
    // programmatic serialisation, cooperative alone, or (ADMM_NO_COOP=1 / not supported) the plain launch that relies
    // on the occupancy check of admm_spm_step_supported and the in-kernel watchdog.
    int& mode = coop_mode()[cur_dev()];      // 0 unprobed, 1 coop + pdl, 2 coop, 3 plain
    if (mode == 0) mode = getenv("ADMM_NO_COOP") ? 3 : (use_pdl() ? 1 : 2);
    while (mode < 3) {
      cudaError_t e = launch_coop(k, grid, dim3(PASS_WARPS * 32), smem, s, mode == 1, *d, *b, lz.comm, lz.lazy, lz.nA);
      if (e == cudaSuccess) return check_launch("admm_spm_step");
      cudaGetLastError();
      ++mode;
    }
  }
  launch_pdl(k, grid, dim3(PASS_WARPS * 32), smem, s, use_pdl(), *d, *b, lz.comm, lz.lazy, lz.nA);
  return check_launch(FNP ? "admm_spm_step" : "admm_spm_pass");
}

template <int NT, int MT, bool FOLD = false>
static int launch_pass_mode(const admm_spm_dims* d, const admm_spm_buffers* b, int mode, bool fused, cudaStream_t s,
                            const LazyArgs& lz, int* max_ctas = nullptr) {
  if (fused && d->nbal > 0) {      // the whole iteration of a small batch in one launch (owner CTAs run the x-update)
    return d->nplanes == 2 ? launch_pass_k<NT, MT, PASS_STEP, 2, true, FOLD>(d, b, s, lz, max_ctas)
                           : launch_pass_k<NT, MT, PASS_STEP, 1, true, FOLD>(d, b, s, lz, max_ctas);
  }
  if (fused) {
    return d->nplanes == 2 ? launch_pass_k<NT, MT, PASS_STEP, 2, false, FOLD>(d, b, s, lz)
                           : launch_pass_k<NT, MT, PASS_STEP, 1, false, FOLD>(d, b, s, lz);
  }
  if (mode == PASS_STEP) return launch_pass_k<NT, MT, PASS_STEP, 0, false, FOLD>(d, b, s, lz);
  return launch_pass_k<NT, MT, PASS_VINIT, 0, false, FOLD>(d, b, s, lz);
}

static int launch_pass(const admm_spm_dims* d, const admm_spm_buffers* b, int mode, bool fused, cudaStream_t s,
                       const LazyArgs& lz = LazyArgs(), int* max_ctas = nullptr) {
  if (d->fold) {      // pairs of sampling points share the MMAs (admm_spm_dims.fold)
    switch (d->Lp / 8) {
      case 2:
        return d->mt == 2 ? launch_pass_mode<2, 2, true>(d, b, mode, fused, s, lz, max_ctas)
                          : launch_pass_mode<2, 1, true>(d, b, mode, fused, s, lz, max_ctas);
      case 5:
        return d->mt == 2 ? launch_pass_mode<5, 2, true>(d, b, mode, fused, s, lz, max_ctas)
                          : launch_pass_mode<5, 1, true>(d, b, mode, fused, s, lz, max_ctas);
      default:
        return launch_pass_mode<8, 1, true>(d, b, mode, fused, s, lz, max_ctas);
    }
  }
  switch (d->Lp / 8) {
    case 2:
      return d->mt == 2 ? launch_pass_mode<2, 2>(d, b, mode, fused, s, lz, max_ctas)
                        : launch_pass_mode<2, 1>(d, b, mode, fused, s, lz, max_ctas);
    case 5:
      return d->mt == 2 ? launch_pass_mode<5, 2>(d, b, mode, fused, s, lz, max_ctas)
                        : launch_pass_mode<5, 1>(d, b, mode, fused, s, lz, max_ctas);
    default:
      return launch_pass_mode<8, 1>(d, b, mode, fused, s, lz, max_ctas);
  }
}

// Can admm_spm_step / admm_spm_step_lazy run these dims?  1: whole columns per CTA; 2: balanced decomposition with owner
// CTAs (every tile group must contain the start of a piece, and all pieces must be co-resident); 0: no.
static int step_mode(const admm_spm_dims* d) {
  if (d->nsplit == 1 && d->nbal == 0) return 1;
  if (d->nbal <= 0 || getenv("ADMM_SPM_TWO_KERNELS")) return 0;
  const int ngroups = ceil_div(d->npt, PASS_WARPS * d->mt);
  if (d->nbal < ngroups) return 0;
  int maxc = 0;
  if (launch_pass(d, nullptr, PASS_STEP, true, nullptr, LazyArgs(), &maxc) != ADMM_OK || d->nbal > maxc) return 0;
  return 2;
}

static int xupdate_grid(const admm_spm_dims* d) { return ceil_div(d->npt * d->nplanes, 4); }

static int launch_xupdate(const admm_spm_dims* d, const admm_spm_buffers* b, cudaStream_t s, const LazyArgs& lz = LazyArgs()) {
  const int grid = xupdate_grid(d);
  const int NT = d->Lp / 8;
  const size_t smem = (size_t)2 * NT * NT * 64 * sizeof(double);      // the two staged L x L operands
  switch (NT) {
    case 2: launch_pdl(spm_xupdate_kernel<2>, dim3(grid), dim3(128), smem, s, use_pdl_x(), *d, *b, lz.comm, lz.lazy); break;
    case 5: launch_pdl(spm_xupdate_kernel<5>, dim3(grid), dim3(128), smem, s, use_pdl_x(), *d, *b, lz.comm, lz.lazy); break;
    default: {
      static std::map<int, bool> configured;      // 64 KB of dynamic shared memory: opt-in, per device
      bool& cfgd = configured[cur_dev()];
      if (!cfgd) {
        cudaFuncSetAttribute(spm_xupdate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cfgd = true;
      }
      launch_pdl(spm_xupdate_kernel<8>, dim3(grid), dim3(128), smem, s, use_pdl_x(), *d, *b, lz.comm, lz.lazy);
      break;
    }
  }
  return check_launch("admm_spm_xupdate");
}

static bool solo_regp(const admm_spm_dims* d, int cs) {
  return d->Lp <= 40 && solo_layout(d->Lp, d->nrt * 8, cs, d->nplanes, false).R <= SOLO_RTHREADS && !getenv("ADMM_SOLO_SMEM");
}
static size_t solo_smem_bytes(const admm_spm_dims* d, int cs) {
  return (size_t)solo_layout(d->Lp, d->nrt * 8, cs, d->nplanes, solo_regp(d, cs)).total * sizeof(double);
}

template <int CS, int LP, bool REGP, bool BW>
static int launch_solo(const admm_spm_dims* d, const admm_spm_buffers* b, const double* G0, const double* PtP, int niter,
                       int interval,
                       cudaStream_t st, int* max_clusters = nullptr) {   // max_clusters != NULL: occupancy query only
  const size_t smem = solo_smem_bytes(d, CS);
  auto kern = spm_solo_kernel<CS, LP, REGP, BW>;
  static std::map<int, size_t> configured_by_dev;     // per instantiation and device
  size_t& configured = configured_by_dev[cur_dev()];
  if (smem > configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(d->nb * CS);
  cfg.blockDim = dim3(SOLO_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_clusters != nullptr) {
    static std::map<int, std::pair<size_t, int>> cache_by_dev;
    size_t& cached_smem = cache_by_dev[cur_dev()].first;
    int& cached_n = cache_by_dev[cur_dev()].second;
    if (cached_smem != smem) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
      }
      cached_smem = smem;
      cached_n = n;
    }
    *max_clusters = cached_n;
    return ADMM_OK;
  }
  cudaError_t e = cudaErrorNotSupported;
  if (BW) {
    // the clusters of a batch-wide solve all-reduce through global memory every iteration: cooperative launch (the
    // driver guarantees co-residency); ADMM_NO_COOP=1 or a driver that refuses the combination with clusters: the
    // plain launch, which relies on the occupancy check of admm_spm_solo_supported and the in-kernel watchdog
    int& mode = coop_mode_solo()[cur_dev()];      // 0 unprobed, 1 cooperative, 2 plain
    if (mode == 0) mode = getenv("ADMM_NO_COOP") ? 2 : 1;
    if (mode == 1) {
      cfg.numAttrs = 2;
      e = cudaLaunchKernelEx(&cfg, kern, *d, *b, G0, PtP, niter, interval);
      if (e != cudaSuccess) {
        cudaGetLastError();
        mode = 2;
        cfg.numAttrs = 1;
      }
    }
  }
  if (e != cudaSuccess) e = cudaLaunchKernelEx(&cfg, kern, *d, *b, G0, PtP, niter, interval);
  if (e != cudaSuccess) {
    set_error("admm_spm_solo: %s", cudaGetErrorString(e));
    return ADMM_ECUDA;
  }
  return check_launch("admm_spm_solo");
}

}  // namespace admm

using namespace admm;

extern "C" {

int admm_spm_prepare_P(const admm_spm_dims* d, const double* P, int ldP, double* Pf, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_prepare_P")) return rc;
  if (d->fold)
    prepare_P_fold_kernel<<<ew_grid((long long)d->nrt * 2 * d->Lp * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, P, ldP, Pf);
  else
    prepare_P_kernel<<<ew_grid((long long)d->nrt * 2 * d->Lp * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, P, ldP, Pf);
  return check_launch("admm_spm_prepare_P");
}

int admm_spm_pack_operator(const admm_spm_dims* d, const double* canon, double* Bf, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_pack_operator")) return rc;
  pack_operator_kernel<<<ceil_div(d->Lp * d->Lp, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d->Lp, canon, Bf);
  return check_launch("admm_spm_pack_operator");
}

int admm_spm_pack_L(const admm_spm_dims* d, const void* canon, int src_is_complex, double* frag, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_pack_L")) return rc;
  const long long total = (long long)d->npt * d->nplanes * (d->Lp / 8) * 64;
  pack_L_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, (const double*)canon, src_is_complex, frag);
  return check_launch("admm_spm_pack_L");
}

int admm_spm_unpack_L(const admm_spm_dims* d, const double* frag, void* canon, int dst_is_complex, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_unpack_L")) return rc;
  unpack_L_kernel<<<ew_grid((long long)d->L * d->nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, frag, (double*)canon,
                                                                                                dst_is_complex);
  return check_launch("admm_spm_unpack_L");
}

int admm_spm_pack_state(const admm_spm_dims* d, const void* h20, const void* x2, int src_is_complex, const double* mu20,
                        double* S, int* flag, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_pack_state")) return rc;
  const int GT = 4 * d->mt;
  const long long total = (long long)((d->npt + GT - 1) / GT * GT) * d->nrt * 64;
  pack_state_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, (const double*)h20, (const double*)x2,
                                                                                 src_is_complex, mu20, S, flag);
  return check_launch("admm_spm_pack_state");
}

int admm_spm_unpack_state(const admm_spm_dims* d, const double* S, const double* mu20_used, const double* him,
                          void* h20, void* x2, int dst_is_complex, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_unpack_state")) return rc;
  unpack_state_kernel<<<ew_grid((long long)d->Nw * d->nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *d, S, mu20_used, him, (double*)h20, (double*)x2, dst_is_complex);
  return check_launch("admm_spm_unpack_state");
}

int admm_spm_factor(const admm_spm_dims* d, int nslots, const int* slots, const double* mu10s, const double* mu20s,
                    const double* G0, const double* PtP, const double* Cvec, double* Ginv_cache, double* w_cache,
                    double* sigma_cache, int* info, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_factor")) return rc;
  if (nslots <= 0) return ADMM_OK;
  const size_t smem = (size_t)(d->L * d->L + 3 * d->L) * sizeof(double);
  spm_factor_kernel<<<nslots, 256, smem, static_cast<cudaStream_t>(stream)>>>(*d, slots, mu10s, mu20s, G0, PtP, Cvec,
                                                                             Ginv_cache, w_cache, sigma_cache, info);
  return check_launch("admm_spm_factor");
}

int admm_spm_refresh_y(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_refresh_y")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = ceil_div(d->npt * d->nplanes, 4);
  switch (d->Lp / 8) {
    case 2: spm_refresh_y_kernel<2><<<grid, 128, 0, s>>>(*d, *b); break;
    case 5: spm_refresh_y_kernel<5><<<grid, 128, 0, s>>>(*d, *b); break;
    default: spm_refresh_y_kernel<8><<<grid, 128, 0, s>>>(*d, *b); break;
  }
  return check_launch("admm_spm_refresh_y");
}

int admm_spm_xupdate(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_xupdate")) return rc;
  return launch_xupdate(d, b, static_cast<cudaStream_t>(stream));
}

int admm_spm_pass(const admm_spm_dims* d, const admm_spm_buffers* b, int mode, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_pass")) return rc;
  ADMM_REQUIRE(mode == PASS_STEP || mode == PASS_VINIT, ADMM_EINVAL, "admm_spm_pass: mode must be 0 (step) or 1 (V from state)");
  return launch_pass(d, b, mode, false, static_cast<cudaStream_t>(stream));
}

int admm_spm_step_supported(const admm_spm_dims* d) {
  if (d == nullptr || check_dims(d, "admm_spm_step_supported") != ADMM_OK) return 0;
  return step_mode(d);
}

int admm_spm_step(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_step")) return rc;
  const int mode = step_mode(d);
  ADMM_REQUIRE(mode != 0, ADMM_EUNSUPPORTED,
               "admm_spm_step: needs whole columns per CTA (nsplit == 1, nbal == 0) or a balanced decomposition with "
               "nbal >= tile groups whose CTAs are all co-resident (see admm_spm_step_supported)");
  ADMM_REQUIRE(mode == 1 || (b->lazy != nullptr && b->xready != nullptr), ADMM_EINVAL, "admm_spm_step: lazy / xready buffers missing");
  return launch_pass(d, b, PASS_STEP, true, static_cast<cudaStream_t>(stream));
}

static int reduce_parts(const admm_spm_dims* d) { return std::max(1, std::min(256, ceil_div(d->nb, 256))); }

int admm_spm_reduce(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_reduce")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int parts = reduce_parts(d);
  launch_pdl(spm_reduce_stage1, dim3(parts), dim3(256), 0, s, use_pdl(), *d, *b);
  launch_pdl(spm_reduce_stage2, dim3(1), dim3(256), 0, s, use_pdl(), parts, *b);
  return check_launch("admm_spm_reduce");
}

int admm_spm_decide(const admm_spm_dims* d, const admm_spm_buffers* b, int do_update_mu, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_decide")) return rc;
  launch_pdl(spm_decide_kernel, dim3(ceil_div(d->nb, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), use_pdl(), *d, *b,
             do_update_mu, 0);
  return check_launch("admm_spm_decide");
}

static int check_comm(const admm_peer_comm* c, const char* who) {
  ADMM_REQUIRE(c != nullptr && c->world >= 1 && c->world <= ADMM_MAX_PEERS && c->rank >= 0 && c->rank < c->world &&
                   c->ctrl != nullptr,
               ADMM_EINVAL, "%s: bad peer communicator", who);
  for (int r = 0; r < c->world; ++r) ADMM_REQUIRE(c->mbox[r] != nullptr, ADMM_EINVAL, "%s: mailbox of rank %d not mapped", who, r);
  return ADMM_OK;
}

int admm_spm_reduce_post(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_reduce_post")) return rc;
  if (int rc = check_comm(c, "admm_spm_reduce_post")) return rc;
  ADMM_REQUIRE(d->batch_wide, ADMM_EINVAL, "admm_spm_reduce_post: batch-wide criterion only");
  launch_pdl(spm_reduce_post_kernel, dim3(reduce_parts(d)), dim3(256), 0, static_cast<cudaStream_t>(stream), use_pdl(), *d, *b,
             *c);
  return check_launch("admm_spm_reduce_post");
}

int admm_spm_decide_peer(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int do_update_mu,
                         admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_decide_peer")) return rc;
  if (int rc = check_comm(c, "admm_spm_decide_peer")) return rc;
  ADMM_REQUIRE(d->batch_wide, ADMM_EINVAL, "admm_spm_decide_peer: batch-wide criterion only");
  launch_pdl(spm_decide_peer_kernel, dim3(ceil_div(d->nb, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), use_pdl(), *d,
             *b, *c, do_update_mu);
  return check_launch("admm_spm_decide_peer");
}

static int lazy_args(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int pending, const char* who,
                     LazyArgs* out) {
  ADMM_REQUIRE(d->batch_wide, ADMM_EINVAL, "%s: batch-wide criterion only", who);
  ADMM_REQUIRE(b->lazy != nullptr && b->cta_partA != nullptr && b->cta_partB != nullptr && b->gsum != nullptr, ADMM_EINVAL,
               "%s: lazy / cta_partA / cta_partB / gsum buffers missing", who);
  *out = LazyArgs();
  if (c != nullptr) {
    if (int rc = check_comm(c, who)) return rc;
    out->comm = *c;
  }
  out->lazy = pending ? 2 : 1;
  out->nA = xupdate_grid(d);
  return ADMM_OK;
}

int admm_spm_step_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int pending,
                       admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_step_lazy")) return rc;
  const int mode = step_mode(d);
  ADMM_REQUIRE(mode != 0, ADMM_EUNSUPPORTED, "admm_spm_step_lazy: dims not supported by the fused step (see admm_spm_step_supported)");
  ADMM_REQUIRE(mode == 1 || b->xready != nullptr, ADMM_EINVAL, "admm_spm_step_lazy: xready buffer missing");
  LazyArgs lz;
  if (int rc = lazy_args(d, b, c, pending, "admm_spm_step_lazy", &lz)) return rc;
  return launch_pass(d, b, PASS_STEP, true, static_cast<cudaStream_t>(stream), lz);
}

int admm_spm_xupdate_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int pending,
                          admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_xupdate_lazy")) return rc;
  LazyArgs lz;
  if (int rc = lazy_args(d, b, c, pending, "admm_spm_xupdate_lazy", &lz)) return rc;
  return launch_xupdate(d, b, static_cast<cudaStream_t>(stream), lz);
}

int admm_spm_pass_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_pass_lazy")) return rc;
  LazyArgs lz;
  if (int rc = lazy_args(d, b, c, 0, "admm_spm_pass_lazy", &lz)) return rc;
  return launch_pass(d, b, PASS_STEP, false, static_cast<cudaStream_t>(stream), lz);
}

int admm_spm_flush(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_flush")) return rc;
  LazyArgs lz;
  if (int rc = lazy_args(d, b, c, 1, "admm_spm_flush", &lz)) return rc;
  launch_pdl(spm_lazy_flush_kernel, dim3(1), dim3(128), 0, static_cast<cudaStream_t>(stream), use_pdl(), *d, *b, lz.comm);
  return check_launch("admm_spm_flush");
}

int admm_spm_reduce_decide(const admm_spm_dims* d, const admm_spm_buffers* b, int do_update_mu, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_reduce_decide")) return rc;
  ADMM_REQUIRE(d->batch_wide, ADMM_EINVAL, "admm_spm_reduce_decide: batch-wide criterion only");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int parts = reduce_parts(d);
  launch_pdl(spm_reduce_stage1, dim3(parts), dim3(256), 0, s, use_pdl(), *d, *b);
  launch_pdl(spm_decide_kernel, dim3(ceil_div(d->nb, 128)), dim3(128), 0, s, use_pdl(), *d, *b, do_update_mu, parts);
  return check_launch("admm_spm_reduce_decide");
}

static int dispatch_solo(const admm_spm_dims* d, const admm_spm_buffers* b, const double* G0, const double* PtP, int niter,
                         int interval,
                         cudaStream_t st, int* max_clusters) {
  // sampling points per CTA <= 256: P, the state and the cached inverse stay in registers
  const bool regp = solo_regp(d, 8);
  const bool bw = d->batch_wide && d->nb > 1;
#define SOLO_CASE(LPV, RG)                                                                                        \
  (bw ? launch_solo<8, LPV, RG, true>(d, b, G0, PtP, niter, interval, st, max_clusters)                           \
      : launch_solo<8, LPV, RG, false>(d, b, G0, PtP, niter, interval, st, max_clusters))
  switch (d->Lp) {
    case 16: return regp ? SOLO_CASE(16, true) : SOLO_CASE(16, false);
    case 40: return regp ? SOLO_CASE(40, true) : SOLO_CASE(40, false);
    default: return SOLO_CASE(64, false);             // 64 doubles per row: shared memory
  }
#undef SOLO_CASE
}

int admm_spm_launch_mode(int which) {
  if (which == 0) return coop_mode()[cur_dev()];
  const int m = coop_mode_solo()[cur_dev()];
  return m == 0 ? 0 : (m == 1 ? 2 : 3);
}

int admm_spm_solo_supported(const admm_spm_dims* d) {
  if (d == nullptr || d->L < 1 || (d->Lp != 16 && d->Lp != 40 && d->Lp != 64) || d->nb < 1) return 0;
  if (solo_smem_bytes(d, 8) > 216 * 1024) return 0;
  if (d->nc > 1) return 0;      // several constraint rows: the batch kernels (the cluster-resident solve folds w into its operators)
  if (d->batch_wide && d->nb > 1) {
    // the batch-wide criterion synchronises the clusters every iteration: all of them have to be co-resident
    int n = 0;
    if (dispatch_solo(d, nullptr, nullptr, nullptr, 0, 0, nullptr, &n) != ADMM_OK || d->nb > n) return 0;
  }
  return 8;
}

int admm_spm_solo(const admm_spm_dims* d, const admm_spm_buffers* b, const double* G0, const double* PtP, int niter,
                  int interval_update_mu, admm_stream_t stream) {
  if (int rc = check_dims(d, "admm_spm_solo")) return rc;
  ADMM_REQUIRE(admm_spm_solo_supported(d) != 0, ADMM_EUNSUPPORTED,
               "admm_spm_solo: L=%d, Nw=%d do not fit the shared memory of an 8-CTA cluster, or (batch-wide criterion) the %d "
               "clusters cannot be co-resident", d->L, d->Nw, d->nb);
  ADMM_REQUIRE(G0 != nullptr && PtP != nullptr && niter >= 0 && interval_update_mu >= 0, ADMM_EINVAL, "admm_spm_solo: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->batch_wide && d->nb > 1) cudaMemsetAsync(b->flags + 3, 0, sizeof(int), st);      // arrival counter of the all-reduce
  return dispatch_solo(d, b, G0, PtP, niter, interval_update_mu, st, nullptr);
}

}  // extern "C"
