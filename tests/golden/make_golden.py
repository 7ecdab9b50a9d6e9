#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (the reference is not on the GPU box):

    python tests/golden/make_golden.py            # reads /root/reference/src

The reference package is imported from ``/root/reference/src`` under its own name
``admmsolver``; inputs come from ``admmsolver_b200/problems.py`` (loaded by path, so
that the product package -- which needs the CUDA library -- is not imported).
Every ``.npz`` holds the outputs of the reference after a fixed iteration count
(x, h, mu, residual histories, objective) and, where the inputs depend on a
BLAS-computed SVD, the inputs too.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ADMM_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)

from admmsolver.matrix import (DenseMatrix, DiagonalMatrix, PartialDiagonalMatrix,  # noqa: E402
                               ScaledIdentityMatrix, identity)
from admmsolver.objectivefunc import (ConstrainedLeastSquares, L1Regularizer,  # noqa: E402
                                      LeastSquares, NonNegativePenalty, SemiPositiveDefinitePenalty)
from admmsolver.optimizer import Model, SimpleOptimizer  # noqa: E402

spec = importlib.util.spec_from_file_location("problems", os.path.join(ROOT, "admmsolver_b200", "problems.py"))
problems = importlib.util.module_from_spec(spec)
sys.modules["problems"] = problems
spec.loader.exec_module(problems)


def _mu_hist_cb(opt, pairs, store):
    def cb():
        store.append([opt._mu[i, j] for (i, j) in pairs])
    return cb


def run_bp(A, y, lam, niter, alpha=1.0, **kw):
    N = A.shape[1]
    lstsq = LeastSquares(alpha, A, y)
    l1 = L1Regularizer(lam, N)
    opt = SimpleOptimizer(Model([lstsq, l1], [(1, 0, identity(N), identity(N))]))
    hist = []
    opt.solve(niter, callback=_mu_hist_cb(opt, [(1, 0)], hist), **kw)
    return dict(x0=opt.x[0], x1=opt.x[1], h10=opt._h[1, 0], mu10=opt._mu[1, 0],
                primal=np.array(opt._primal_residual), dual=np.array(opt._dual_residual),
                mu_seen=np.array(hist), objective=opt(opt.x))


def run_spm(p, g, D, niter, packed_nb=None, **kw):
    """Single problem (packed_nb None) or the packed PartialDiagonalMatrix formulation."""
    L, Nw = p.s.size, p.P.shape[0]
    if packed_nb is None:
        lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), g, p.C, D)
        l1 = L1Regularizer(p.lam, L)
        nn = NonNegativePenalty(Nw)
        conds = [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]
    else:
        nb = packed_nb
        rest = (nb,)
        A = PartialDiagonalMatrix(-DiagonalMatrix(p.s), rest)
        Cm = PartialDiagonalMatrix(p.C, rest)
        lstsq = ConstrainedLeastSquares(1.0, A, g.ravel(), Cm, D.astype(float))
        l1 = L1Regularizer(p.lam, L * nb)
        nn = NonNegativePenalty(Nw * nb)
        conds = [(0, 1, identity(L * nb), identity(L * nb)),
                 (0, 2, PartialDiagonalMatrix(p.P, rest), identity(Nw * nb))]
    opt = SimpleOptimizer(Model([lstsq, l1, nn], conds), mu=p.mu)
    hist = []
    opt.solve(niter, callback=_mu_hist_cb(opt, [(1, 0), (2, 0)], hist), **kw)
    return dict(x0=opt.x[0], x1=opt.x[1], x2=opt.x[2], h10=opt._h[1, 0], h20=opt._h[2, 0],
                mu10=opt._mu[1, 0], mu20=opt._mu[2, 0],
                primal=np.array(opt._primal_residual), dual=np.array(opt._dual_residual),
                mu_seen=np.array(hist), objective=opt(opt.x))


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    # ---- cfg1a: the notebook / test instance, known-answer vector of basis_pursuit.ipynb:137-138
    A, y, xa = problems.basis_pursuit_instance(100, 1000, 20, 1234)
    r = run_bp(A, y, 0.1, 100)
    print("known answer:", np.abs(xa).max(), np.abs(xa - r["x0"]).max())
    save("bp_notebook", y=y, xanswer=xa, **r)

    # ---- cfg1b: BASELINE config 1 (200x1000, 10-sparse, 1000 iterations)
    A, y, xa = problems.basis_pursuit_instance(200, 1000, 10, 0)
    save("bp_cfg1", y=y, xanswer=xa, **run_bp(A, y, 0.1, 1000))

    # ---- cfg4 samples: 128x512, seeds 0..3, 300 iterations
    for b in range(4):
        A, y, xa = problems.basis_pursuit_instance(128, 512, 10, b)
        save(f"bp_cfg4_seed{b}", y=y, xanswer=xa, **run_bp(A, y, 0.1, 300))

    # ---- tiny LASSO of test_optimizer.py:13-50 (int inputs there; float here)
    Al = np.array([[2.0, 1.0]])
    yl = np.array([2.0])
    save("lasso_1x2", A=Al, y=yl, **run_bp(Al, yl, 0.1, 100))

    # ---- tall basis pursuit (M > N) exercising the direct N x N path
    rs = np.random.RandomState(7)
    At = rs.randn(96, 40)
    xt = np.zeros(40)
    xt[:5] = rs.randn(5)
    yt = At @ xt + 1e-3 * rs.randn(96)
    save("bp_tall", A=At, y=yt, **run_bp(At, yt, 0.05, 250, alpha=0.7))

    # ---- cfg2: SpM single, full size L=39, Nw=2000 (inputs regenerated by the recipe; y & C stored)
    basis = problems.ir_basis()
    p = problems.spm_single(basis, Nw=2000)
    r = run_spm(p, p.g, p.D, 1000)
    # P_probe: every 97th row of P, so a test can tell whether its regenerated P has the same bits
    save("spm_cfg2", s=p.s, C=p.C, g=p.g, P_probe=p.P[::97].copy(), **r)

    # ---- SpM reduced grid with ALL inputs stored (SVD-independent parity)
    p = problems.spm_single(basis, Nw=192)
    r = run_spm(p, p.g, p.D, 700)
    save("spm_small", s=p.s, C=p.C, g=p.g, P=p.P, lam=p.lam, mu=p.mu, **r)

    # ---- packed batch (batch-wide mu / stopping): nb=6 complex spectra, Nw=192
    pb = problems.spm_batch(6, basis, Nw=192, seed=3)
    r = run_spm(pb, pb.g, pb.D, 400, packed_nb=6)
    save("spm_packed", s=pb.s, C=pb.C, g=pb.g, P=pb.P, lam=pb.lam, mu=pb.mu, **r)

    # ---- the same six spectra as independent reference instances (per-problem mode)
    outs = [run_spm(pb, pb.g[:, b].copy(), np.array([1.0]), 400) for b in range(6)]
    save("spm_independent", s=pb.s, C=pb.C, g=pb.g, P=pb.P, lam=pb.lam, mu=pb.mu,
         **{k: np.stack([o[k] for o in outs], axis=-1) for k in ("x0", "x1", "x2", "h10", "h20")},
         mu10=np.array([o["mu10"] for o in outs]), mu20=np.array([o["mu20"] for o in outs]),
         primal=np.stack([o["primal"] for o in outs], axis=-1),
         dual=np.stack([o["dual"] for o in outs], axis=-1),
         objective=np.array([o["objective"] for o in outs]))

    # ---- term-level solves (objectivefunc.py): LeastSquares with dense HPD mu (complex), partial, CLS
    rs = np.random.RandomState(100)
    cr = lambda *sh: rs.randn(*sh) + 1j * rs.randn(*sh)
    N1, N2 = 4, 2
    y = cr(N1)
    A = cr(N1, N2)
    h = cr(N2)
    mu = cr(N2, N2)
    mu = mu @ mu.T.conjugate()
    x = LeastSquares(2.0, A, y).solve(h, DenseMatrix(mu))
    C = cr(1, N2)
    D = cr(1)
    xc = ConstrainedLeastSquares(2.0, A, y, C, D).solve(h, DenseMatrix(mu))
    a2 = cr(2, 1)
    y2 = cr(2 * 20)
    h2 = cr(20)
    mu2 = ScaledIdentityMatrix(20, 1.5)
    xp = LeastSquares(0.3, PartialDiagonalMatrix(a2, (20,)), y2).solve(h2, mu2)
    hl = rs.randn(7)
    xl1 = L1Regularizer(0.3, 7).solve(hl, DiagonalMatrix(np.linspace(0.5, 2.0, 7)))
    xnn = NonNegativePenalty(7).solve(hl + 0.1j, ScaledIdentityMatrix(7, 0.7))
    save("terms", y=y, A=A, h=h, mu=mu, x_ls=x, C=C, D=D, x_cls=xc, a2=a2, y2=y2, h2=h2, x_partial=xp,
         hl=hl, x_l1=xl1, x_nn=xnn)

    # ---- generic 3-term model with dense rectangular couplings (exercises the generic executor)
    rs = np.random.RandomState(5)
    n0, n1, n2 = 6, 5, 4
    Ag = rs.randn(8, n0)
    yg = rs.randn(8)
    E10a = rs.randn(5, n0)
    lst = LeastSquares(1.3, Ag, yg)
    l1 = L1Regularizer(0.2, n1)
    nn = NonNegativePenalty(n2)
    Pg = rs.randn(n2, n0)
    conds = [(0, 1, E10a, identity(n1)), (0, 2, Pg, DiagonalMatrix(np.linspace(1.0, 2.0, n2)))]
    opt = SimpleOptimizer(Model([lst, l1, nn], conds), mu=0.7)
    opt.solve(150, interval_update_mu=20)
    save("generic3", A=Ag, y=yg, E1=E10a, P=Pg, x0=opt.x[0], x1=opt.x[1], x2=opt.x[2],
         h10=opt._h[1, 0], h20=opt._h[2, 0], mu10=opt._mu[1, 0], mu20=opt._mu[2, 0],
         primal=np.array(opt._primal_residual), dual=np.array(opt._dual_residual),
         objective=opt(opt.x))


def main_psd():
    """SemiPositiveDefinitePenalty: term-level solves for every axis / mu type and one
    model through the reference's loop.  `python tests/golden/make_golden.py psd`."""
    rs = np.random.RandomState(100)
    cr = lambda *sh: rs.randn(*sh) + 1j * rs.randn(*sh)
    N, K = 10, 20
    h = cr(N * N * K)
    out = dict(h=h)
    out["x_identity"] = SemiPositiveDefinitePenalty((N, N, K), axis=2).solve(h, identity(N * N * K))
    out["x_partial"] = SemiPositiveDefinitePenalty((N, N, K), axis=2).solve(
        h, PartialDiagonalMatrix(ScaledIdentityMatrix(N * N, 1.7), (K,)))
    dvar = np.linspace(0.5, 2.5, N * N)
    out["dvar"] = dvar
    out["x_partial_diag"] = SemiPositiveDefinitePenalty((N, N, K), axis=2).solve(
        h, PartialDiagonalMatrix(DiagonalMatrix(dvar), (K,)))
    dfull = np.linspace(0.3, 3.0, N * N * K)
    out["dfull"] = dfull
    out["x_diag"] = SemiPositiveDefinitePenalty((N, N, K), axis=2).solve(h, DiagonalMatrix(dfull))
    h0 = cr(6 * 7 * 7)
    out["h_axis0"] = h0
    out["x_axis0"] = SemiPositiveDefinitePenalty((6, 7, 7), axis=0).solve(h0, ScaledIdentityMatrix(6 * 7 * 7, 0.8))
    h1 = cr(5 * 4 * 5)
    out["h_axis1"] = h1
    out["x_axis1"] = SemiPositiveDefinitePenalty((5, 4, 5), axis=1).solve(h1, ScaledIdentityMatrix(5 * 4 * 5, 1.3))
    h32 = rs.randn(32 * 32 * 3)
    out["h_n32"] = h32
    out["x_n32"] = SemiPositiveDefinitePenalty((3, 32, 32), axis=0).solve(h32, ScaledIdentityMatrix(32 * 32 * 3, 1.0))
    # slices larger than a warp (the CTA-wide one-sided Jacobi of admm_prox_psd): own generator, so that the draws above
    # and below keep their values
    rs40 = np.random.RandomState(40)
    h40 = rs40.randn(2 * 40 * 40)
    out["h_n40"] = h40
    out["x_n40"] = SemiPositiveDefinitePenalty((40, 2, 40), axis=1).solve(h40, ScaledIdentityMatrix(2 * 40 * 40, 0.6))
    # matrix-valued least squares with a PSD constraint through SimpleOptimizer.solve
    n, k = 4, 3
    nx = n * n * k
    Am = rs.randn(2 * nx, nx)
    ym = rs.randn(2 * nx)
    opt = SimpleOptimizer(Model([LeastSquares(0.9, Am, ym), SemiPositiveDefinitePenalty((n, n, k), axis=2)],
                                [(0, 1, identity(nx), identity(nx))]), mu=0.5)
    opt.solve(120, interval_update_mu=25)
    out.update(loop_A=Am, loop_y=ym, loop_x0=opt.x[0], loop_x1=opt.x[1], loop_h10=opt._h[1, 0], loop_mu10=opt._mu[1, 0],
               loop_primal=np.array(opt._primal_residual), loop_dual=np.array(opt._dual_residual),
               loop_objective=opt(opt.x))
    # the same through the loop with a slice larger than a warp (36 x 36: the CTA-wide Jacobi of admm_prox_psd);
    # inputs are regenerated from the seed by the test, only the outputs are stored
    rs36 = np.random.RandomState(36)
    n36 = 36
    nx = n36 * n36
    A36 = rs36.randn(nx + 200, nx) / np.sqrt(nx)
    y36 = rs36.randn(nx + 200)
    opt = SimpleOptimizer(Model([LeastSquares(1.0, A36, y36), SemiPositiveDefinitePenalty((n36, n36, 1), axis=2)],
                                [(0, 1, identity(nx), identity(nx))]), mu=0.5)
    opt.solve(60, interval_update_mu=20)
    out.update(loop36_x0=opt.x[0], loop36_x1=opt.x[1], loop36_mu10=opt._mu[1, 0], loop36_primal=np.array(opt._primal_residual),
               loop36_objective=opt(opt.x))
    save("psd", **out)


def _after(opt, loose):
    """What the step-wise public methods return right after solve() (optimizer.py:232-299,324)."""
    out = dict(x_old0=opt._x_old[0].copy(), residual_after=np.array(opt.residual()),
               converged_tight=np.array(opt.check_convergence(1e-12)), converged_loose=np.array(opt.check_convergence(loose)),
               primal=np.array(opt._primal_residual), dual=np.array(opt._dual_residual), x0=opt.x[0].copy())
    opt.update_mu()
    out["mu_after_update"] = np.array([opt._mu[i, j] for (i, j) in ((1, 0), (2, 0)) if i < opt._mu.shape[0]])
    return out


def main_after():
    """`_x_old` hand-over: residual() / check_convergence() / update_mu() called right after solve() on the models the
    fused engines take over.  `python tests/golden/make_golden.py after`."""
    out = {}

    def bp(tag, A, y, lam, niter, loose, **kw):
        N = A.shape[1]
        opt = SimpleOptimizer(Model([LeastSquares(1.0, A, y), L1Regularizer(lam, N)], [(1, 0, identity(N), identity(N))]))
        opt.solve(niter, **kw)
        out.update({f"{tag}_{k}": v for k, v in _after(opt, loose).items()})

    A, y, _ = problems.basis_pursuit_instance(100, 1000, 20, 1234)
    bp("bpnb", A, y, 0.1, 100, 1e-2)
    A, y, _ = problems.basis_pursuit_instance(128, 512, 10, 2)
    bp("bp301", A, y, 0.1, 301, 1e-3)          # last iteration (index 300) is a mu-update iteration
    bp("lasso", np.array([[2.0, 1.0]]), np.array([2.0]), 0.1, 100, 1e-6)     # early exit (iteration 41)

    basis = problems.ir_basis()
    for tag, p, niter, nb in (("spm1", problems.spm_single(basis, Nw=192), 250, None),
                              ("spm6", problems.spm_batch(6, basis, Nw=192, seed=3), 130, 6)):
        L, Nw = p.s.size, p.P.shape[0]
        if nb is None:
            lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), p.g, p.C, p.D)
            terms = [lstsq, L1Regularizer(p.lam, L), NonNegativePenalty(Nw)]
            conds = [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]
        else:
            rest = (nb,)
            lstsq = ConstrainedLeastSquares(1.0, PartialDiagonalMatrix(-DiagonalMatrix(p.s), rest), p.g.ravel(),
                                            PartialDiagonalMatrix(p.C, rest), p.D.astype(float))
            terms = [lstsq, L1Regularizer(p.lam, L * nb), NonNegativePenalty(Nw * nb)]
            conds = [(0, 1, identity(L * nb), identity(L * nb)), (0, 2, PartialDiagonalMatrix(p.P, rest), identity(Nw * nb))]
        opt = SimpleOptimizer(Model(terms, conds), mu=p.mu)
        opt.solve(niter)
        out.update({f"{tag}_{k}": v for k, v in _after(opt, 1e-2).items()})
        out.update({f"{tag}_s": p.s, f"{tag}_C": p.C, f"{tag}_g": p.g, f"{tag}_P": p.P, f"{tag}_lam": p.lam, f"{tag}_mu": p.mu,
                    f"{tag}_D": np.asarray(p.D)})
    save("after_solve", **out)


def main_complex():
    """Complex operators through the reference's loop (objectivefunc.py:76-77,89-96; matrix.py:100-118): a complex LASSO
    and a packed, constrained 3-term model whose A, C (two rows), D and coupling matrix are all complex.
    `python tests/golden/make_golden.py complex`."""
    rs = np.random.RandomState(2024)
    cr = lambda *sh: rs.randn(*sh) + 1j * rs.randn(*sh)
    out = {}
    # (a) complex LASSO, 60 x 100
    M, N = 60, 100
    A = cr(M, N)
    xt = np.zeros(N, dtype=complex)
    xt[rs.permutation(N)[:8]] = rs.randn(8)
    y = A @ xt + 1e-3 * cr(M)
    opt = SimpleOptimizer(Model([LeastSquares(0.8, A, y), L1Regularizer(0.15, N)], [(1, 0, identity(N), identity(N))]))
    opt.solve(200, interval_update_mu=25)
    out.update(a_A=A, a_y=y, a_x0=opt.x[0], a_x1=opt.x[1], a_h10=opt._h[1, 0], a_mu10=opt._mu[1, 0],
               a_primal=np.array(opt._primal_residual), a_dual=np.array(opt._dual_residual), a_objective=opt(opt.x))
    # (b) packed batch of nb = 48 constrained problems sharing complex A (40 x 36), complex C (2 x 36), complex dense
    #     coupling P (50 x 36) to a non-negative block; batch-wide mu / stopping
    nb, Mb, Lb, Nwb = 48, 40, 36, 50
    Ab, Cb, Pb = cr(Mb, Lb), cr(2, Lb), cr(Nwb, Lb)
    yb, Db = cr(Mb * nb), cr(2 * nb)
    rest = (nb,)
    terms = [ConstrainedLeastSquares(1.1, PartialDiagonalMatrix(Ab, rest), yb, PartialDiagonalMatrix(Cb, rest), Db),
             L1Regularizer(0.3, Lb * nb), NonNegativePenalty(Nwb * nb)]
    conds = [(0, 1, identity(Lb * nb), identity(Lb * nb)), (0, 2, PartialDiagonalMatrix(Pb, rest), identity(Nwb * nb))]
    opt = SimpleOptimizer(Model(terms, conds), mu=0.5)
    opt.solve(160, interval_update_mu=20)
    out.update(b_A=Ab, b_C=Cb, b_P=Pb, b_y=yb, b_D=Db, b_x0=opt.x[0], b_x1=opt.x[1], b_x2=opt.x[2], b_h10=opt._h[1, 0],
               b_h20=opt._h[2, 0], b_mu10=opt._mu[1, 0], b_mu20=opt._mu[2, 0], b_primal=np.array(opt._primal_residual),
               b_dual=np.array(opt._dual_residual), b_objective=opt(opt.x))
    save("complex_ops", **out)


def main_multirow():
    """SpM with SEVERAL constraint rows (sum rule + first and second moment of the spectrum; objectivefunc.py:148-157
    with a 3 x L resp. 2 x L matrix C): single problem and packed batch.  `python tests/golden/make_golden.py multirow`."""
    basis = problems.ir_basis()
    out = {}
    # single problem, three rows
    p = problems.spm_single(basis, Nw=192)
    L, Nw = p.s.size, p.P.shape[0]
    wq = basis.womega
    C3 = np.stack([basis.v_omega @ wq, basis.v_omega @ (wq * basis.omega), basis.v_omega @ (wq * basis.omega ** 2)])
    D3 = C3 @ p.rho_l
    lstsq = ConstrainedLeastSquares(1.0, -DiagonalMatrix(p.s), p.g, C3, D3)
    opt = SimpleOptimizer(Model([lstsq, L1Regularizer(p.lam, L), NonNegativePenalty(Nw)],
                                [(0, 1, identity(L), identity(L)), (0, 2, p.P, identity(Nw))]), mu=p.mu)
    opt.solve(300, interval_update_mu=50)
    out.update(a_s=p.s, a_P=p.P, a_g=p.g, a_C=C3, a_D=D3, a_lam=p.lam, a_mu=p.mu, a_x0=opt.x[0], a_x1=opt.x[1], a_x2=opt.x[2],
               a_h20=opt._h[2, 0], a_mu10=opt._mu[1, 0], a_mu20=opt._mu[2, 0], a_primal=np.array(opt._primal_residual),
               a_dual=np.array(opt._dual_residual), a_objective=opt(opt.x))
    # packed batch of 6 complex spectra, two rows, D per problem
    nb = 6
    pb = problems.spm_batch(nb, basis, Nw=192, seed=3)
    C2 = C3[:2]
    D2 = (C2 @ pb.rho_l)                                                 # (2, nb), batch fastest when flattened
    rest = (nb,)
    lstsq = ConstrainedLeastSquares(1.0, PartialDiagonalMatrix(-DiagonalMatrix(pb.s), rest), pb.g.ravel(),
                                    PartialDiagonalMatrix(C2, rest), D2.ravel())
    opt = SimpleOptimizer(Model([lstsq, L1Regularizer(pb.lam, L * nb), NonNegativePenalty(Nw * nb)],
                                [(0, 1, identity(L * nb), identity(L * nb)),
                                 (0, 2, PartialDiagonalMatrix(pb.P, rest), identity(Nw * nb))]), mu=pb.mu)
    opt.solve(250, interval_update_mu=50)
    out.update(b_g=pb.g, b_C=C2, b_D=D2, b_x0=opt.x[0], b_x1=opt.x[1], b_x2=opt.x[2], b_mu10=opt._mu[1, 0], b_mu20=opt._mu[2, 0],
               b_primal=np.array(opt._primal_residual), b_dual=np.array(opt._dual_residual), b_objective=opt(opt.x))
    print("multirow:", out["a_mu10"], out["a_mu20"], out["b_mu10"], out["b_mu20"], len(out["a_primal"]), len(out["b_primal"]))
    save("spm_multirow", **out)


def main_fuzz():
    """Seeded random models of tests/golden/fuzz_models.py through the reference's loop.
    `python tests/golden/make_golden.py fuzz`."""
    import admmsolver.matrix as RM
    import admmsolver.objectivefunc as RF
    import admmsolver.optimizer as RO
    sys.path.insert(0, HERE)
    import fuzz_models
    out = {}
    for seed in range(fuzz_models.NSEEDS):
        opt, nterms = fuzz_models.build((RM, RF, RO), seed)
        opt.solve(fuzz_models.NITER, interval_update_mu=fuzz_models.INTERVAL)
        for k in range(nterms):
            out[f"s{seed}_x{k}"] = opt.x[k]
        out[f"s{seed}_mu"] = np.array([opt._mu[k, 0] for k in range(1, nterms)])
        out[f"s{seed}_primal"] = np.array(opt._primal_residual)
        out[f"s{seed}_dual"] = np.array(opt._dual_residual)
        out[f"s{seed}_objective"] = np.array(opt(opt.x))
        print(seed, nterms, len(opt._primal_residual), out[f"s{seed}_mu"], float(opt(opt.x)))
    save("fuzz_models", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fuzz":
        main_fuzz()
    elif len(sys.argv) > 1 and sys.argv[1] == "multirow":
        main_multirow()
    elif len(sys.argv) > 1 and sys.argv[1] == "psd":
        main_psd()
    elif len(sys.argv) > 1 and sys.argv[1] == "after":
        main_after()
    elif len(sys.argv) > 1 and sys.argv[1] == "complex":
        main_complex()
    else:
        main()
        main_psd()
        main_after()
        main_complex()
        main_fuzz()
        main_multirow()
