/*
 * libadmm_b200 -- C ABI of the B200-native ADMM engine (drop-in for the solve loop of
 * SpM-lab/admmsolver).  Plain pointers and sizes only; every pointer is a DEVICE pointer unless
 * its name ends in `_host`.  All work is enqueued on the caller's CUDA stream (`stream` is a
 * `cudaStream_t` passed as void*); nothing here synchronises the device.  The caller owns every
 * buffer; the library is stateless (no handles, no hidden allocations) and therefore re-entrant.
 * Every function returns 0 on success, non-zero on error (`admm_last_error()` has the text).
 *
 * The reference has no FFI: its boundary is the Python API.  Each entry
 * point below names the reference code (file:line under /root/reference/src/admmsolver) whose
 * arithmetic it replaces.  The Python mirror that binds them is admmsolver_b200/_lib.py.
 *
 * Layout conventions
 *   - dense matrices are row-major with an explicit leading dimension (elements, not bytes);
 *   - complex128 is interleaved (re, im) like NumPy/torch; `is_complex` selects the kernels;
 *   - "fragment layout" (SpM engine): an (Lp x ncol) real array stored as
 *       F[ct][j][lane][e],  ct = column tile of 8 columns, j = 8-wide slice of L, lane = 4*g+t,
 *       element = (row l = 8*j + 2*t + e, column c = 8*ct + g)                (see DESIGN.md).
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMM_ABI_VERSION 5

enum admm_status { ADMM_OK = 0, ADMM_EINVAL = 1, ADMM_ECUDA = 2, ADMM_EUNSUPPORTED = 3 };
enum admm_op { ADMM_OP_N = 0, ADMM_OP_T = 1, ADMM_OP_H = 2 };

typedef void* admm_stream_t;

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                     */
/* ------------------------------------------------------------------------------------------ */
int admm_abi_version(void);
const char* admm_last_error(void);
/* SM count, compute capability and opt-in shared memory of the current device. */
int admm_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes);

/* ------------------------------------------------------------------------------------------ */
/* structured-matrix primitives (generic executor; replace the NumPy calls of matrix.py)       */
/* ------------------------------------------------------------------------------------------ */
/* C[m x n] = op(A)[m x k] @ B[k x n].  Replaces `DenseMatrix.__matmul__` (matrix.py:100-118),
 * `_matvec_impl`'s tensordot (matrix.py:392-397) and the matrix-matrix products of
 * `Model.__init__` (optimizer.py:71-76).  op(A)=A (lda>=k), A^T or A^H (A stored k x m, lda>=m). */
int admm_gemm(int is_complex, int op_a, int m, int n, int k, const void* A, int lda,
              const void* B, int ldb, void* C, int ldc, admm_stream_t stream);

/* out[i, :] = d[i] * V[i, :] for i < nd, zero rows for nd <= i < rows_out (rectangular diagonal,
 * zero padded / truncated).  Replaces `DiagonalMatrix.__matmul__` (matrix.py:255-274). */
int admm_diag_mul(int is_complex, int rows_out, int nd, int ncols, const void* d, const void* V,
                  int ldv, void* out, int ldo, admm_stream_t stream);

/* out[i] = a * x[i] + b * y[i] over n doubles (complex vectors: pass 2n); y may be NULL (b
 * ignored).  Replaces the vector algebra of `_hk`/`one_sweep` (optimizer.py:175-207,334-341). */
int admm_axpby(long long n, double a, const double* x, double b, const double* y, double* out,
               admm_stream_t stream);

/* out[i] = 1 / x[i] (op 0) or conj(x[i]) (op 1) over n elements.  Replaces `DiagonalMatrix.inv`
 * and `.conjugate()` (matrix.py:223-226,239-240). */
int admm_ewise_unary(int op, int is_complex, long long n, const void* x, void* out,
                     admm_stream_t stream);

/* L1 prox: out = soft(-Re(h)/mu_diag, 0.5*alpha/mu_diag), strict comparisons.  h has stride
 * h_stride doubles (2 for complex h: the imaginary part is dropped), out has stride out_stride
 * (2: the imaginary slot is zeroed).  Replaces `L1Regularizer.solve` + `_softmax`
 * (objectivefunc.py:174-195,335-355). */
int admm_prox_l1(long long n, const double* h, int h_stride, const double* mu_diag, double alpha,
                 double* out, int out_stride, admm_stream_t stream);

/* Non-negative projection: out = max(0, -Re(h)/mu_diag).  Replaces `NonNegativePenalty.solve` +
 * `_project_plus` (objectivefunc.py:256-271,330-333). */
int admm_prox_nonneg(long long n, const double* h, int h_stride, const double* mu_diag,
                     double* out, int out_stride, admm_stream_t stream);

/* PSD-cone projection of `nbatch` real symmetric n x n matrices (n <= 1024) embedded in one vector:
 * element (p, q) of matrix m is entry m*stride_batch + p*stride_row + q*stride_col.  X = -Re(h)/mu_diag
 * elementwise; per matrix the LOWER triangle defines the symmetric matrix (as np.linalg.eigh does) and
 * the negative eigenvalues are removed.  n <= 32: one warp per matrix, cyclic two-sided Jacobi in shared memory.
 * n > 32: one CTA per matrix, parallel-order one-sided Jacobi on the shifted matrix X + |X|_F I (positive definite, so
 * its orthogonalised columns are (lambda_k + sigma) v_k) -- in shared memory up to n = 160, beyond in `work`
 * (admm_prox_psd_work_doubles returns 1 when a workspace of *grid * n * n doubles is needed; NULL otherwise).
 * Replaces `SemiPositiveDefinitePenalty.solve` incl. its per-slice np.linalg.eigh loop (objectivefunc.py:294-327). */
int admm_prox_psd_work_doubles(int n, long long nbatch, long long* grid);
int admm_prox_psd(int n, long long nbatch, long long stride_batch, long long stride_row,
                  long long stride_col, const double* h, int h_stride, const double* mu_diag,
                  double* out, int out_stride, double* work, admm_stream_t stream);

/* Singular value decomposition of ONE real m x n matrix by one-sided Jacobi on a cooperative grid (high relative
 * accuracy of small singular values: the kernel of the intermediate-representation basis spans 16 decades).
 * Wt: the COLUMNS of the input as rows (n rows of length m, row stride ldw), overwritten with sigma_k u_k;
 * Vt: n x n, the identity on entry (or NULL), row k overwritten with v_k; sv[k] = sigma_k (unsorted);
 * scratch: max_sweeps doubles; info[0] = sweeps used, negative if not converged.  Replaces the `np.linalg.svd` /
 * `sparse_ir` basis construction either side of the solver (spm.ipynb:52-65,155-163). */
int admm_svd_jacobi(int m, int n, double* Wt, int ldw, double* Vt, int ldv, double* sv, double* scratch,
                    int max_sweeps, int* info, admm_stream_t stream);

/* out[0] = sum_i x[i]^2 (y == NULL) or sum_i (x[i]-y[i])^2 over n doubles; deterministic
 * two-stage reduction; `scratch` needs 1024 doubles.  Replaces `np.linalg.norm`
 * (util.py:40, optimizer.py:261,267,284,289). */
int admm_sumsq(long long n, const double* x, const double* y, double* out, double* scratch,
               admm_stream_t stream);

/* The six squared norms residual(), check_convergence() and update_mu() need for ONE coupled pair
 * (optimizer.py:251-299) in a single pass: out[0..2] = |p1-p2|^2, |p1|^2, |p2|^2 over np doubles,
 * out[3..5] = |d1-d2|^2, |d1|^2, |d2|^2 over nd doubles; deterministic two-stage reduction; `scratch`
 * needs 6 * 1024 doubles. */
int admm_pair_norms(long long np, const double* p1, const double* p2, long long nd, const double* d1,
                    const double* d2, double* out, double* scratch, admm_stream_t stream);

/* General inverse by Gauss-Jordan with partial pivoting (one CTA).  `work` is n x 2n elements.
 * info[0] != 0 if a zero pivot was met.  Replaces `np.linalg.inv` (matrix.py:77-78). */
int admm_inverse(int is_complex, int n, const void* A, int lda, void* Ainv, int ldi, void* work,
                 int* info, admm_stream_t stream);

/* Batched inverse of real symmetric positive-definite matrices, in place, one CTA per matrix:
 * block Gauss-Jordan with 8x8 blocks on the FP64 tensor cores -- the matrix register-resident for
 * n <= 128, streamed from L2 for n <= 512 -- and a scalar Gauss-Jordan beyond.  `mask` (may be NULL)
 * selects the matrices to process (mask[b] != 0); info[b] = 1-based index of the first non-positive
 * pivot, 0 if none.  Used for (alpha A^H A + mu)^-1 (objectivefunc.py:89-96): the factor that is
 * cached per mu. */
int admm_spd_inverse_batched(int n, int nbatch, double* A, long long batch_stride, int lda,
                             const int* mask, int* info, admm_stream_t stream);

/* Batched inverse of complex128 Hermitian positive-definite matrices (interleaved storage), in place, on the same
 * FP64 tensor-core block Gauss-Jordan: H = X + iY is HPD iff its real form [[X, -Y], [Y, X]] (2n x 2n) is SPD, and the
 * inverse of the real form is the real form of H^-1.  `work` holds nbatch * (2n)^2 doubles; batch_stride and lda count
 * complex elements; mask / info as above (a matrix whose info is non-zero is left untouched).  This is the cached
 * factor (alpha A^H A + mu)^-1 of a LeastSquares / ConstrainedLeastSquares / L2Regularizer term with a complex A
 * (objectivefunc.py:76-77,89-96). */
int admm_hpd_inverse_batched(int n, int nbatch, void* A, long long batch_stride, int lda, double* work,
                             const int* mask, int* info, admm_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Pattern B engine: SpM  [ConstrainedLeastSquares, L1Regularizer, NonNegativePenalty] with     */
/* conditions (0,1,I,I), (0,2,P,I); many problems share s, P, C (PartialDiagonalMatrix packing).*/
/* One ADMM iteration = admm_spm_step (or admm_spm_xupdate + admm_spm_pass for small batches),     */
/* then admm_spm_reduce_decide (sharded batch: admm_spm_reduce_post + admm_spm_decide_peer).     */
/* ------------------------------------------------------------------------------------------ */
typedef struct admm_spm_dims {
  int L;        /* basis size (size_x of terms 0 and 1)                                        */
  int Lp;       /* L padded to 16, 40 or 64 (the instantiated tensor-core tile counts)         */
  int Nw;       /* number of sampling points (size_x of term 2)                                */
  int nrt;      /* number of 8-row tiles, ceil(Nw / 8) rounded up to a multiple of 4           */
  int nb;       /* number of problems                                                          */
  int npt;      /* number of 8-problem tiles = ceil(nb / 8)                                    */
  int nplanes;  /* 1: real data (imaginary parts identically zero), 2: complex128 state        */
  int nsplit;   /* number of partial-sum slots of V / normsB per problem tile: the row splits of the
                   classic grid, or the bound on pieces per tile group of the balanced one       */
  int mt;       /* problem tiles per warp in the pass kernel (1 or 2)                          */
  int nbal;     /* 0: classic grid (tile groups x nsplit equal row ranges).  > 0: balanced decomposition
                   for small batches -- the ceil(npt/(4 mt)) * (nrt/4) group-chunks are cut into nbal
                   equal contiguous pieces, one CTA each (one wave that loads every SM equally)   */
  int batch_wide; /* 1: mu and the stopping test use norms over the whole batch (packed
                     reference semantics), 0: per-problem mu / stopping                        */
  int nc;       /* rows of the constraint C x0 = D (1 .. ADMM_SPM_MAX_NC; 0 is read as 1).  Cvec is [nc][Lp],
                   w_cache [slot][nc][Lp], Dre [plane][nc][8 npt]; sigma_cache [slot][nc*nc] holds sigma = C w for
                   nc = 1 and the INVERSE of C G^-1 C^T for nc > 1 (objectivefunc.py:148-157)     */
  int fold;     /* 1: P has the parity of the IR basis on a symmetric grid, P[Nw-1-r][l] = (-1)^l P[r][l] bit for
                   bit (the caller has checked it) and Nw is even: the pass
                   works on PAIRS of sampling points (r, Nw-1-r) -- the even and the odd columns of one row of P serve
                   both -- which halves its tensor work.  State tiles then alternate: tile 2i holds points
                   8i..8i+7, tile 2i+1 their mirror images, nrt = 2 * (ceil(Nw/16) rounded up to even); Pf holds,
                   per pair tile, the Lp/8 slices of Q = P x0 and then 2 * ceil(Lp/16) slices of V = P^T u with
                   the even columns first (element (lane, e) of slice j = P[8i+2t+e][16j+2g], then [16j+2g+1]).  */
} admm_spm_dims;
#define ADMM_SPM_MAX_NC 4

/* P (Nw x L, row-major, ld = ldP) -> Pf[nrt][2][Lp/8][32][2], zero padded, fragment-major: per
 * 8-row tile first the B operand of Q = P x0 (element (lane=4g+t, e) of slice j = P[8rt+g][8j+2t+e]),
 * then the B operand of V = P^T u (P[8rt+2t+e][8j+g]); the pass kernel pulls 4-tile chunks of it
 * into shared memory with TMA bulk copies and reads them with conflict-free 16-byte loads.
 * With d->fold the pair-tile layout described at admm_spm_dims.fold is written instead (it needs less than the
 * nrt * 2 * Lp * 8 doubles of the plain layout; P must have the parity stated there -- not checked here). */
int admm_spm_prepare_P(const admm_spm_dims* d, const double* P, int ldP, double* Pf,
                       admm_stream_t stream);

/* canonical Lp x Lp row-major operator (zero padded) -> fragment-major
 * Bf[jk][jn][lane][e] = B[8jk+2t+e][8jn+g]: the layout in which the x-update GEMMs read P^T P and
 * the cached inverses (one coalesced 16-byte load per lane and k-step). */
int admm_spm_pack_operator(const admm_spm_dims* d, const double* canon, double* Bf,
                           admm_stream_t stream);

/* canonical (rows x nb) complex128 or float64 (batch index fastest) <-> fragment layout.
 * `src_is_complex`: canonical array is interleaved complex.  For nplanes == 1 only real parts
 * move.  pack zero-fills padding rows/columns. */
int admm_spm_pack_L(const admm_spm_dims* d, const void* canon, int src_is_complex, double* frag,
                    admm_stream_t stream);
int admm_spm_unpack_L(const admm_spm_dims* d, const double* frag, void* canon, int dst_is_complex,
                      admm_stream_t stream);

/* (h20, x2) canonical (Nw x nb) <-> implicit state
 *   S[grp][chunk][tig][r4][lane][2]   (problem tile pt = grp*GT + tig, GT = 4*mt tiles per pass CTA;
 *                                      8-row tile rt = 4*chunk + r4; lane = 4*g + t holds problem
 *                                      8*pt+g, sampling points 8*rt+2t, +1),
 * chunk-major per CTA tile group so that what one pass CTA needs per chunk is one contiguous block
 * (one TMA bulk copy).  S has ceil(npt/GT)*GT * nrt * 64 doubles.
 *   s = Re(h20) - mu20 * x2   (complementarity: Re(h20) = max(0,s), mu20*x2 = max(0,-s)).
 * Im(h20) is not part of S: it only ever enters the iteration through P^T Im(h20), which the
 * x-update maintains in L-space (z <- z - mu20 P^T P Im(x0)); the caller reconstructs
 * Im(h20) = Im(h20)_initial - P a, a = sum_k mu20_k Im(x0_k) (buffer `aim`), and passes it to unpack
 * as `him` (Nw x nb real, may be NULL = 0).
 * pack sets flag[0] = 1 if the given state is not representable (x2 < 0 or Re(h20)*x2 != 0).
 * With d->fold the 8-row state tiles alternate between sampling points of the lower half and their mirror images
 * (admm_spm_dims.fold); the canonical arrays on the other side of pack / unpack are the same as without. */
int admm_spm_pack_state(const admm_spm_dims* d, const void* h20, const void* x2, int src_is_complex,
                        const double* mu20, double* S, int* flag, admm_stream_t stream);
int admm_spm_unpack_state(const admm_spm_dims* d, const double* S, const double* mu20_used,
                          const double* him, void* h20, void* x2, int dst_is_complex,
                          admm_stream_t stream);

/* Factor cache entry for `nslots` (mu10, mu20) pairs: G = G0 + mu10 I + mu20 PtP; Ginv = G^-1
 * (Lp x Lp, zero padded, stored fragment-major like admm_spm_pack_operator), w = Ginv C^T (Lp),
 * sigma = C w.  PtP is canonical (Lp x Lp row-major).  G0 = alpha A^H A (L x L, ld Lp).
 * Replaces `_get_B` and the per-iteration `B @ Ch`, `inv(C @ xi2)` of
 * `ConstrainedLeastSquares.solve` (objectivefunc.py:89-96,148-153).  slots[i] is the cache row
 * to fill for the pair (mu10s[i], mu20s[i]). */
int admm_spm_factor(const admm_spm_dims* d, int nslots, const int* slots, const double* mu10s,
                    const double* mu20s, const double* G0, const double* PtP, const double* Cvec,
                    double* Ginv_cache, double* w_cache, double* sigma_cache, int* info,
                    admm_stream_t stream);

typedef struct admm_spm_buffers {
  /* shared operators */
  const double* Pf;       /* [nrt][2][Lp/8][32][2] fragment-major P (admm_spm_prepare_P)        */
  const double* PtPf;     /* P^T P, Lp x Lp, fragment-major (admm_spm_pack_operator)            */
  const double* Cvec;     /* Lp                                                                */
  const double* Ginv_cache; /* nslot x (Lp x Lp fragment-major)                                */
  const double* w_cache;    /* nslot x Lp                                                      */
  const double* sigma_cache;/* nslot                                                           */
  /* per problem (length 8*npt) */
  const int* slot;        /* factor-cache row of the problem's current (mu10, mu20)            */
  double* mu10;
  double* mu20;
  double* mu20_used;      /* mu20 of the last executed pass (needed to decode x2 from S)       */
  int* done;              /* 1: converged / padding -> frozen                                  */
  int* iters;             /* iterations executed                                               */
  double* last_res;       /* [8*npt][2] primal, dual residual of the last executed iteration   */
  const double* Dre;      /* [nplanes][8*npt] right-hand side of C x0 = D                      */
  /* fragment-layout arrays, (Lp x 8*npt*nplanes), column tile ct = pt*nplanes + plane */
  const double* b0;       /* alpha A^H y                                                       */
  double* x0;
  double* x1;
  double* h10;
  double* y0;             /* P^T P x0 (Gram-form norms of pair (2,0); maintained by the x-update) */
  double* x0_old;         /* x0 at the start of the last executed iteration (`_x_old[0]`, optimizer.py:324),
                             fragment layout like x0; NULL: not kept (saves the store)             */
  double* V;              /* [nsplit][ncolumn tiles][Lp/8][32][2]  P^T(h20 + mu20 x2) partials;
                             imaginary-plane tiles: z = P^T Im(h20) in split 0 (owned by xupdate)    */
  double* aim;            /* fragment layout; imaginary-plane tiles accumulate sum_k mu20_k Im(x0_k) */
  /* implicit (h20, x2) state */
  double* S;              /* [ceil(npt/GT)][nrt/4][GT][4][32][2], GT = 4*mt (see pack_state)    */
  /* norms */
  double* normsA;         /* [8*npt*nplanes][8] from xupdate; slot 7 = |P x0|^2 of the column   */
  double* normsB;         /* [nsplit][8*npt*nplanes][2] from pass (real-plane columns only):
                             |P Re(x0) - x2|^2, |x2|^2                                         */
  double* gsum;           /* [16] batch-wide sums (reduce), only batch_wide                    */
  double* gpart;          /* [256][16] scratch of the two-stage reduce                         */
  double* cta_partA;      /* [CTAs of the x-update stage][10] per-CTA partial sums of the ten squared norms
                             (lazy batch-wide iterations, see admm_spm_step_lazy)                */
  double* cta_partB;      /* [CTAs of the pass][2] per-CTA partials of |P Re(x0) - x2|^2, |x2|^2 (unfused path) */
  int* lazy;              /* [4], zero-initialised: 0 CTA ticket of the in-kernel reduction, 2 batch converged
                             (every kernel of the engine returns at once when set), 3 launch sequence number
                             of the fused balanced step                                          */
  int* xready;            /* [tile groups], zero-initialised, never reset: x-update units (tile, plane) of the group
                             finished so far, over all launches of the fused balanced step (the pieces of a group
                             wait for launch sequence number x units of the group)                */
  /* control */
  int* iter_counter;      /* device scalar: iterations launched so far in this solve call      */
  int* flags;             /* [4]: 0 any mu changed, 1 number of problems done, 2 first non-positive pivot of an in-kernel re-inversion (admm_spm_solo), 3 arrival counter of its batch-wide all-reduce */
  double* history;        /* [hist_cap][2] primal/dual per iteration (batch_wide or nb==1), or NULL */
  int hist_cap;
  double lam;             /* L1 weight                                                         */
  double rtol;
  double max_mu;
  double fact_incr;
  double th_change;
  /* balanced decomposition with pieces of unequal length (both NULL: nbal equal pieces).  A grid of 3 CTAs per SM does
     not advance evenly -- the warp schedulers favour the CTAs that became resident first -- so the caller may hand
     later CTAs shorter pieces: */
  const long long* bal_bounds;  /* [nbal + 1] ascending group-chunk boundaries, bal_bounds[0] = 0, bal_bounds[nbal] = T;
                                   every tile group must contain the start of at most nsplit - 1 pieces            */
  const int* bal_first;         /* [tile groups] first CTA whose piece overlaps the group                           */
} admm_spm_buffers;

/* x-update (term 0, `ConstrainedLeastSquares.solve`, objectivefunc.py:138-157, with
 * `_hk`/`_mu_k`, optimizer.py:175-230), L1 z-update (objectivefunc.py:174-195) and dual ascent of
 * pair (1,0) (optimizer.py:334-341); norms of pair (1,0) and the Gram-form norms of pair (2,0)
 * (|P v|^2 = v^T (P^T P) v, with y0 = P^T P x0 carried from iteration to iteration: call
 * admm_spm_refresh_y once after x0 was loaded from outside). */
int admm_spm_xupdate(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream);

/* y0 = P^T P x0 for the current x0 (after set_state / reset). */
int admm_spm_refresh_y(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream);

/* One streaming sweep over the implicit (Re h20, x2) state: s' = Re h20 - mu20 P Re(x0) (FP64
 * tensor cores; the accumulator starts at Re h20), which encodes the non-negative z-update
 * (objectivefunc.py:256-271) and the dual ascent of pair (2,0) (optimizer.py:334-341) at once;
 * residual partial sums (optimizer.py:251-274) and V = P^T(h20 + mu20 x2) = P^T |s'| for the next
 * x-update (optimizer.py:194-200), all in one read+write of the state.
 * mode 0: step.  mode 1: no step; V = P^T(Re h20 + mu20 x2) from the current state with x2 decoded
 * by mu20_used -- run after the state or mu20 changed (set_state, update_mu). */
int admm_spm_pass(const admm_spm_dims* d, const admm_spm_buffers* b, int mode,
                  admm_stream_t stream);

/* admm_spm_xupdate + admm_spm_pass(mode 0) in ONE kernel: every warp first does the x-update of
 * its own problem tiles (both planes) and then streams their state, so x0 goes from the
 * x-update to the tensor-core operand registers without a round trip.  Whole columns per CTA
 * (nsplit == 1, nbal == 0), or the balanced decomposition with owner CTAs (admm_spm_step_supported). */
int admm_spm_step(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream);

/* Which form of admm_spm_step / admm_spm_step_lazy these dims get: 1 whole columns per CTA (nsplit == 1,
 * nbal == 0); 2 the balanced decomposition of a small batch with OWNER CTAs -- of the CTAs whose piece starts
 * in a tile group the first runs the x-update of the group's tiles, then releases the group (xready) to the
 * CTAs that stream its chunks, so the whole iteration is one launch (needs nbal >= tile groups and all nbal
 * CTAs co-resident: checked here against the kernel's occupancy); 0: use admm_spm_xupdate + admm_spm_pass. */
int admm_spm_step_supported(const admm_spm_dims* d);

/* Batch-wide norms: deterministic two-stage sum over all problems into gsum[16].  The caller
 * all-reduces gsum across ranks (NCCL) before admm_spm_decide when the batch is sharded. */
int admm_spm_reduce(const admm_spm_dims* d, const admm_spm_buffers* b, admm_stream_t stream);

/* admm_spm_reduce + admm_spm_decide for an unsharded batch with the batch-wide criterion in two
 * launches instead of three: every CTA of the decide kernel adds the stage-1 partials itself (same
 * fixed order).  The sharded case needs gsum in memory for the NCCL all-reduce: use the two calls. */
int admm_spm_reduce_decide(const admm_spm_dims* d, const admm_spm_buffers* b, int do_update_mu,
                           admm_stream_t stream);

/* residual() / check_convergence() / update_mu() (optimizer.py:232-299) per problem or
 * batch-wide; increments iter_counter; appends to history. */
int admm_spm_decide(const admm_spm_dims* d, const admm_spm_buffers* b, int do_update_mu,
                    admm_stream_t stream);

/* ---- batch-wide criterion of a batch sharded over the GPUs of one box --------------
 * The only exchange of the path: every iteration the ranks all-reduce their ten squared-norm sums
 * (`residual()` / `check_convergence()` / `update_mu()` of the packed batch, optimizer.py:232-299).
 * Instead of a library collective between two kernels, the reduction kernel itself PUSHES its sums into
 * every peer's mailbox over NVLink (peer-mapped memory, one-way latency) and the decision kernel
 * polls its own mailbox: no host involvement, graph-capturable, three launches per iteration.
 *
 * Mailbox of a rank: [2][ADMM_MAX_PEERS][32] 64-bit words (ADMM_MAILBOX_BYTES, zero-initialised).
 * Word k of source rank r in buffer (seq & 1) = (seq << 32) | 32-bit half k of r's ten doubles
 * (low half first): an 8-byte store is atomic, so data and validity travel together (no fence,
 * no flag round trip).  seq counts the reductions of the plan; two buffers suffice because a rank
 * can post reduction seq + 2 only after every peer has posted seq + 1, i.e. finished reading seq. */
#define ADMM_MAX_PEERS 16
#define ADMM_MAILBOX_BYTES (2 * ADMM_MAX_PEERS * 32 * 8)
typedef struct admm_peer_comm {
  int rank, world;                        /* world <= ADMM_MAX_PEERS                                  */
  unsigned long long* mbox[ADMM_MAX_PEERS]; /* mbox[r]: mailbox of rank r as mapped into THIS process   */
  unsigned* ctrl;                         /* local device [4], zero-initialised: 0 sequence number of the
                                             last posted reduction, 1 CTA ticket of the reduce kernel */
} admm_peer_comm;

/* Device memory that peers on the same box can map (cudaMalloc + cudaIpcGetMemHandle; zero-filled).
 * `handle_host` receives the 64-byte IPC handle to send to the peers (any host transport: the
 * Python layer uses torch.distributed.all_gather_object).  admm_peer_open maps a peer's handle into
 * this process (peer access is enabled lazily); close / free undo them.  Host-synchronous set-up
 * calls, not part of the iteration. */
int admm_peer_alloc(size_t bytes, void** devptr, unsigned char* handle_host);
int admm_peer_open(const unsigned char* handle_host, void** devptr);
int admm_peer_close(void* devptr);
int admm_peer_free(void* devptr);

/* admm_spm_reduce whose last CTA pushes the ten sums into the mailbox of every rank (its own
 * included) with sequence number ctrl[0] + 1. */
int admm_spm_reduce_post(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c,
                         admm_stream_t stream);

/* admm_spm_decide on the batch-wide sums of ALL ranks: every CTA waits until the `world` posts of
 * sequence ctrl[0] have arrived in the local mailbox and adds them in rank order (identical totals,
 * hence identical decisions, on every rank).  A peer that never posts trips a watchdog (~5 s):
 * flags[2] = -2 and no decision is taken. */
int admm_spm_decide_peer(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c,
                         int do_update_mu, admm_stream_t stream);

/* ---- "lazy" batch-wide iterations: ONE launch per iteration -----------------------------------------
 * Between two update_mu() iterations the only thing residual() / check_convergence() (optimizer.py:232-274)
 * decide is whether to stop, so their work is folded into the kernels on either side of them:
 *   tail of the step (or pass) kernel: every CTA leaves the partial sums of the ten squared norms of ITS
 *     problems in cta_partA/B; the CTA that finishes last adds them in a fixed order into gsum (one rank) or
 *     pushes them into every rank's mailbox (sharded batch, `c` non-NULL: what admm_spm_reduce_post does);
 *   head of the NEXT step (or x-update) kernel (`pending` = 1): every CTA obtains the sums (gsum, or the
 *     `world` posts of its mailbox added in rank order), evaluates the stopping test redundantly and identically,
 *     and returns without touching the state when it fired; CTA 0 appends the residuals to `history`, counts
 *     the iteration (iters[0], iter_counter) and on convergence sets lazy[2] and flags[1] = nb.
 * admm_spm_flush takes the pending decision of the last launched iteration (head logic alone, one CTA): call it
 * before an update_mu() iteration (admm_spm_step + reduce + decide as before) and before the host looks at the
 * flags.  Same arithmetic as the three-kernel iteration; per-problem iters / last_res / done are not maintained
 * (entry 0 is; the batch shares them).  Batch-wide criterion only. */
int admm_spm_step_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int pending,
                       admm_stream_t stream);
int admm_spm_xupdate_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c, int pending,
                          admm_stream_t stream);
int admm_spm_pass_lazy(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c,
                       admm_stream_t stream);
int admm_spm_flush(const admm_spm_dims* d, const admm_spm_buffers* b, const admm_peer_comm* c,
                   admm_stream_t stream);

/* A handful of problems (spm.ipynb: ONE), the whole SimpleOptimizer.solve loop (optimizer.py:302-320)
 * in ONE launch: every problem is kept resident by a thread-block cluster of 8 CTAs (each owns
 * Nw/8 sampling points: its rows of P and of the state live in shared memory, the L-vectors in
 * registers); per iteration the CTAs exchange their partial P^T|s'| through distributed shared
 * memory behind one cluster barrier and take the residual / convergence / mu decisions -- including
 * the re-inversion of alpha A^H A + mu after update_mu -- redundantly and identically.  Runs up to
 * `niter` iterations per problem (stops at convergence), mu update every `interval_update_mu`
 * iterations (0: never).  G0 = alpha A^H A and PtP = P^T P (Lp x Lp, row-major, zero padded, as for admm_spm_factor).  Reads and
 * writes the same buffers as the batch kernels (state in, state out; V, y0, mu20_used consistent),
 * sets flags[0] when a mu changed (the caller re-maps its factor cache), flags[1] += converged
 * problems, flags[2] = first non-positive pivot.  Per-problem criterion: any nb (clusters run in waves).  Batch-wide
 * criterion with nb > 1 (a packed batch of a few problems): the clusters all-reduce their ten squared norms through
 * gpart / flags[3] every iteration, so all nb clusters have to be co-resident (nb <= ~14 on a B200).
 * admm_spm_solo_supported: cluster size used (8) if L, Nw fit the cluster's shared memory (and, batch-wide, the nb
 * clusters fit the GPU at once), else 0. */
int admm_spm_solo_supported(const admm_spm_dims* d);
/* How the two kernels whose CTAs wait for each other inside ONE launch were last launched on the current device
 * (which = 0: fused balanced step of admm_spm_step / admm_spm_step_lazy; which = 1: batch-wide admm_spm_solo):
 * 0 not launched yet, 1 cooperative launch (co-residency guaranteed by the driver) with programmatic stream
 * serialisation, 2 cooperative launch, 3 plain launch (ADMM_NO_COOP=1, or the driver refused the combination:
 * co-residency then rests on the occupancy check and the in-kernel watchdog, flags[2] = -3 / -1). */
int admm_spm_launch_mode(int which);
int admm_spm_solo(const admm_spm_dims* d, const admm_spm_buffers* b, const double* G0, const double* PtP,
                  int niter, int interval_update_mu, admm_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* Pattern A engine: basis pursuit / LASSO  [LeastSquares, L1Regularizer], condition (1,0,I,I);  */
/* every problem has its own real A (M x N).  One CTA per problem runs all iterations.           */
/* ------------------------------------------------------------------------------------------ */
typedef struct admm_bp_buffers {
  int nb, M, N;
  int woodbury;           /* 1: M < N, factor K = A A^T + (mu/alpha) I (M x M); 0: G = alpha A^T A + mu I */
  int nk;                 /* order of the factor (M or N)                                      */
  const double* A;        /* [nb][M][N] row-major                                              */
  const double* At;       /* [nb][ceil(N/32)][M][32] tile-major copy of A (admm_bp_tile_A), zero
                             padded; NULL: admm_bp_iterate falls back to the two-sweep kernel  */
  const double* aty;      /* [nb][N]   alpha A^T y                                             */
  const double* gram;     /* [nb][nk][nk]  A A^T or A^T A                                      */
  double* Kinv;           /* [nb][nk][nk]                                                      */
  double* x0;             /* [nb][N]                                                           */
  double* x1;             /* [nb][N]                                                           */
  double* h;              /* [nb][N]                                                           */
  double* x0_old;         /* [nb][N] x0 at the start of the last executed iteration (`_x_old[0]`,
                             optimizer.py:324), written when a launch of admm_bp_iterate ends; NULL: not kept */
  double* mu;             /* [nb]                                                              */
  int* need_factor;       /* [nb] 1: Kinv stale for the current mu                             */
  int* done;              /* [nb]                                                              */
  int* iters;             /* [nb] iterations executed in this solve call                       */
  double* last_res;       /* [nb][2]                                                           */
  double* history;        /* [nb][hist_cap][2] or NULL                                         */
  int hist_cap;
  double alpha, lam, rtol, max_mu, fact_incr, th_change;
  int interval_update_mu;
} admm_bp_buffers;

/* 1 if an M x N problem fits the shared-memory resident iteration kernels (the N-vectors and two
 * vectors of the factor's order stay in one CTA's shared memory: (6 N + 2 min(M, N) + 160) * 8 bytes
 * <= 220 KB), else 0 -- lets a caller pick another path BEFORE building Gram matrices and factors. */
int admm_bp_supported(int M, int N);

/* aty = alpha A^T y and gram = A A^T (woodbury) or A^T A.  Replaces `LeastSquares.__init__`
 * (objectivefunc.py:76-77) and the per-call `Ac @ y` (objectivefunc.py:108).  Either output may be
 * NULL: new data y for the same operators needs only aty (y, aty non-NULL, gram NULL). */
int admm_bp_setup(const admm_bp_buffers* b, const double* y, double* aty, double* gram,
                  admm_stream_t stream);

/* At = tile-major copy of A: every 32-column tile of a problem is one contiguous M x 32 block, so
 * the fused iteration kernel fetches it with one TMA bulk copy (whole DRAM pages). */
int admm_bp_tile_A(const admm_bp_buffers* b, double* At, admm_stream_t stream);

/* Kinv = (gram + (mu/alpha) I)^-1 or (alpha gram + mu I)^-1 for problems with need_factor set;
 * clears the flag.  Replaces `_get_B` (objectivefunc.py:89-96). */
int admm_bp_factor(const admm_bp_buffers* b, int* info, admm_stream_t stream);

/* Runs iterations [iter_begin, iter_end) of `SimpleOptimizer.solve` (optimizer.py:302-341) for
 * every problem that is neither done nor waiting for a factor: x-update by the cached inverse
 * (Woodbury or direct), soft threshold, dual ascent, residuals, convergence test, and -- when
 * (iter % interval_update_mu == 0) -- update_mu.  With At set (Woodbury path, M <= 256) A is streamed
 * ONCE per iteration: the column tile that yields A^T s also feeds the next iteration's A r'.  A problem whose mu changed sets need_factor
 * and stops; the caller runs admm_bp_factor and calls again with the same iter_end.
 * A handful of problems (nb <= 16, Woodbury, 8 <= M <= 480, N >= 128: the notebook / test instances of the
 * reference) run cluster-resident instead: a thread-block cluster of 16 (or 8) CTAs per problem keeps its
 * column slices of A and row slices of K^-1 in shared memory and the N-vectors in registers for all
 * iterations; the two exchanges per iteration (all-reduce of A r', all-gather of K^-1 t) are DSMEM pushes. */
int admm_bp_iterate(const admm_bp_buffers* b, int iter_end, admm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADMM_B200_H */
