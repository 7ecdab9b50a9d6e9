"""Seeded random models for the generic-executor parity fuzz.

``build(mods, seed)`` assembles the same model from the same draws whether ``mods`` are the reference's modules
(``tests/golden/make_golden.py fuzz`` -> ``fuzz_models.npz``) or the drop-in package (``tests/test_api_gpu.py``): 2-4 terms
drawn from LeastSquares / ConstrainedLeastSquares / L2Regularizer / L1Regularizer / NonNegativePenalty, real or complex
data and operators, coupled to term 0 through identity / scaled identity / diagonal / dense rectangular (one or both
sides) / PartialDiagonalMatrix operators -- the term and matrix types of SURVEY 8(a), in combinations neither fused
engine takes.
"""
import numpy as np

NITER = 60
INTERVAL = 15
NSEEDS = 14


def build(mods, seed):
    """mods = (matrix module, objectivefunc module, optimizer module).  Returns (optimizer, number of terms)."""
    M, F, O = mods
    rs = np.random.RandomState(1000 + seed)
    cplx = bool(rs.randint(2))

    def rnd(*shape):
        a = rs.randn(*shape)
        return a + 1j * rs.randn(*shape) if cplx else a

    n0 = int(rs.randint(4, 13))
    m0 = n0 + int(rs.randint(0, 6))
    kind0 = int(rs.randint(3))
    A0 = rnd(m0, n0)
    if kind0 == 0:
        t0 = F.LeastSquares(float(rs.uniform(0.5, 2.0)), A0, rnd(m0))
    elif kind0 == 1:
        nc = int(rs.randint(1, 3))
        t0 = F.ConstrainedLeastSquares(float(rs.uniform(0.5, 2.0)), A0, rnd(m0), rnd(nc, n0), rnd(nc))
    else:
        # packed: A (x) I_rest with a small inner matrix
        inner = int(rs.choice([d for d in (2, 3, 4) if n0 % d == 0] or [1]))
        rest = n0 // inner
        a_in = rnd(inner + 1, inner)
        t0 = F.LeastSquares(float(rs.uniform(0.5, 2.0)), M.PartialDiagonalMatrix(a_in, (rest,)), rnd((inner + 1) * rest))
    terms = [t0]
    conds = []
    nterms = int(rs.randint(2, 5))
    for k in range(1, nterms):
        kind = int(rs.randint(4))
        ek = int(rs.randint(5))
        if kind in (0, 1) and ek == 3:
            ek = 2       # L1 / non-negative terms need a diagonal mu (objectivefunc.py:187,263): no dense operator on their side
        # size of the partner block and the pair of coupling operators E0 x0 = Ek xk
        if ek == 0:
            nk, E0, Ek = n0, M.identity(n0), M.identity(n0)
        elif ek == 1:
            nk, E0, Ek = n0, M.ScaledIdentityMatrix(n0, float(rs.uniform(0.5, 1.5))), M.DiagonalMatrix(rs.uniform(0.5, 2.0, n0))
        elif ek == 2:
            nk = int(rs.randint(3, 10))
            E0, Ek = rs.randn(nk, n0), M.identity(nk)                      # dense rectangular on the x0 side
        elif ek == 3:
            nk = int(rs.randint(3, 10))
            # dense on both sides; the reference's dual residual multiplies E[j,i] @ E[i,j] (optimizer.py:269-270),
            # which only type-checks when the partner's own operator is square
            E0, Ek = rs.randn(nk, n0), rs.randn(nk, nk) + 2.0 * np.eye(nk)
        else:
            inner = int(rs.choice([d for d in (2, 3, 4) if n0 % d == 0] or [1]))
            rest = n0 // inner
            nk = inner * rest
            E0, Ek = M.PartialDiagonalMatrix(rs.randn(inner, inner), (rest,)), M.identity(nk)
        if kind == 0:
            terms.append(F.L1Regularizer(float(rs.uniform(0.05, 0.5)), nk))
        elif kind == 1:
            terms.append(F.NonNegativePenalty(nk))
        elif kind == 2:
            terms.append(F.L2Regularizer(float(rs.uniform(0.1, 1.0)), rnd(nk + 1, nk)))
        else:
            terms.append(F.LeastSquares(float(rs.uniform(0.2, 1.0)), rnd(nk + 2, nk), rnd(nk + 2)))
        conds.append((0, k, E0, Ek))
    mu = float(rs.choice([0.3, 1.0, 3.0]))
    opt = O.SimpleOptimizer(O.Model(terms, conds), mu=mu)
    return opt, nterms
