"""Batched front-ends of the two fused CUDA engines.

``SharedSpM``           -- pattern B: many SpM problems sharing one (s, P, C); the reference reaches
                           this only through ``PartialDiagonalMatrix`` packing (matrix.py:301-401),
                           which makes mu and the stopping test batch-global (``batch_wide=True``).
                           ``batch_wide=False`` is the per-problem mode (every column behaves like
                           its own ``SimpleOptimizer`` instance).
``BatchedBasisPursuit`` -- pattern A: independent problems, each with its own A (M x N).

Host code is orchestration only: it owns torch device buffers, fills the C structs of
``include/admm_b200.h`` and enqueues kernels on the current CUDA stream.  All arithmetic of the
solve loop runs in ``libadmm_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import BpBuffers, SpmBuffers, SpmDims, call, ptr, stream

__all__ = ["SharedSpM", "BatchedBasisPursuit", "SpMHostStream"]

_F64 = torch.float64
_C128 = torch.complex128


def _dev_tensor(a, device, dtype=None) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a.to(device=device)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _pad_L(L: int) -> int:
    for cand in (16, 40, 64):
        if L <= cand:
            return cand
    raise NotImplementedError(f"SpM engine supports basis sizes up to 64, got L={L}")


class SharedSpM:
    """Pattern B on the GPU.

    minimise   alpha ||y - A x0||^2 + lam |x1|_1    s.t.  C x0 = D,  x0 = x1,  P x0 = x2 >= 0

    with real operators (A^H A = ``G0`` is L x L real, P is Nw x L real, C is 1 x L real) and
    real or complex data.  Construct either from the SpM form (``s``, ``g``: A = -diag(s), y = g)
    or from precomputed ``G0 = alpha A^H A`` and ``b0 = alpha A^H y`` (``from_operators``).
    """

    CACHE_SLOTS = 64
    #: relative over-/under-weight of the pieces of the first / last residency tier of a full-wave balanced
    #: decomposition (tools/bal_sweep.py)
    BAL_SKEW = 0.0

    def __init__(self, s, P, C_, D, g, lam: float, mu: float = 0.1, alpha: float = 1.0,
                 batch_wide: bool = False, max_mu: float = 1e3, nsplit: Optional[int] = None,
                 group=None, force_complex: Optional[bool] = None, mt: Optional[int] = None, nbal: Optional[int] = None,
                 collective: Optional[str] = None, keep_x_old: bool = False, bal_skew: Optional[float] = None,
                 fold: Optional[bool] = None):
        dev = _lib.require_cuda()
        s_t = _dev_tensor(s, dev, _F64)
        g_t = _dev_tensor(g, dev)
        if g_t.ndim == 1:
            g_t = g_t[:, None]
        g_t = g_t.contiguous()
        G0 = torch.diag(alpha * s_t * s_t)
        # b0 = alpha A^H y with A = -diag(s):  diag kernel, then scale
        L, nb = g_t.shape
        cplx = g_t.is_complex()
        gv = g_t if cplx else g_t.to(_F64)
        b0 = torch.empty_like(gv)
        sd = (-alpha * s_t).to(gv.dtype)
        call("admm_diag_mul", int(cplx), L, L, nb, ptr(sd), ptr(gv), nb, ptr(b0), nb, stream())
        self._init_common(G0, b0, P, C_, D, lam, mu, mu, batch_wide, max_mu, nsplit, group, force_complex, mt, nbal,
                          collective, keep_x_old, bal_skew, fold)
        self._s = s_t
        self._g = gv
        self._alpha = alpha

    @classmethod
    def from_operators(cls, G0, b0, P, C_, D, lam: float, mu10: float, mu20: float,
                       batch_wide: bool = True, max_mu: float = 1e3, nsplit: Optional[int] = None,
                       group=None, force_complex: Optional[bool] = None, mt: Optional[int] = None,
                       collective: Optional[str] = None, keep_x_old: bool = False,
                       fold: Optional[bool] = None) -> "SharedSpM":
        self = cls.__new__(cls)
        dev = _lib.require_cuda()
        b0_t = _dev_tensor(b0, dev)
        if b0_t.ndim == 1:
            b0_t = b0_t[:, None].contiguous()
        self._init_common(_dev_tensor(G0, dev, _F64), b0_t, P, C_, D, lam, mu10, mu20, batch_wide, max_mu,
                          nsplit, group, force_complex, mt, None, collective, keep_x_old, None, fold)
        self._s = None
        self._g = None
        self._alpha = None
        return self

    # ------------------------------------------------------------------ setup
    def _init_common(self, G0, b0, P, C_, D, lam, mu10, mu20, batch_wide, max_mu, nsplit, group, force_complex,
                     mt=None, nbal_req=None, collective=None, keep_x_old=False, bal_skew=None, fold=None):
        dev = _lib.require_cuda()
        self.device = dev
        self.group = group
        # Sharded batch with the batch-wide criterion: how the ten squared-norm sums are all-reduced every iteration.
        #   "peer" (default): the reduction kernel pushes them into every rank's mailbox over NVLink, the decision
        #                     kernel polls its own -- no library call between kernels, CUDA-graph capturable;
        #   "nccl":           admm_spm_reduce -> torch.distributed.all_reduce -> admm_spm_decide (eager only).
        if collective is None:
            collective = "peer"
        if collective not in ("peer", "nccl"):
            raise ValueError("collective must be 'peer' or 'nccl'")
        self.collective = collective if (group is not None and batch_wide) else None
        self._peer = None
        if self.collective == "peer":
            from .peer import PeerMailbox
            self._peer = PeerMailbox(group)
        P_t = _dev_tensor(P, dev, _F64)
        Nw, L = P_t.shape
        assert b0.shape[0] == L and G0.shape == (L, L)
        nb = b0.shape[1]
        Cv = _dev_tensor(np.asarray(C_) if not isinstance(C_, torch.Tensor) else C_, dev, _F64)
        Cv = Cv.reshape(1, -1) if Cv.ndim == 1 else Cv
        nc = int(Cv.shape[0])
        assert Cv.shape[1] == L, "C must be (rows, L)"
        if nc > 4:
            raise NotImplementedError("the fused SpM engine supports up to 4 constraint rows, got %d" % nc)
        # D: one value per (constraint row, problem): scalar, (nb,) [one row], (nc,) [same for all problems] or (nc, nb)
        D_t = _dev_tensor(np.asarray(D) if not isinstance(D, torch.Tensor) else D, dev)
        if D_t.numel() == 1:
            D_t = D_t.reshape(1, 1).expand(nc, nb)
        elif D_t.numel() == nc * nb:
            D_t = D_t.reshape(nc, nb)
        elif D_t.numel() == nc:
            D_t = D_t.reshape(nc, 1).expand(nc, nb)
        else:
            raise AssertionError("D must have one entry per constraint row (and problem)")
        D_t = D_t.contiguous()
        cplx = bool(b0.is_complex() or D_t.is_complex())
        if force_complex is not None:
            cplx = bool(force_complex) or cplx
        self.L, self.Nw, self.nb = L, Nw, nb
        self.is_complex = cplx
        Lp = _pad_L(L)
        nrt = ((Nw + 7) // 8 + 3) // 4 * 4
        # Folded pass: on a symmetric frequency grid the IR basis functions have the parity of their index,
        # P[Nw-1-r, l] = (-1)^l P[r, l].  When that holds BIT FOR BIT the pass works on pairs of sampling points and
        # needs half the tensor work (admm_spm_dims.fold); the state is then stored pair tile by pair tile.
        if fold is None:
            fold = os.environ.get("ADMM_SPM_NO_FOLD") is None
        fold = bool(fold) and Nw % 2 == 0 and Nw >= 16
        if fold:
            sign = torch.ones(L, dtype=_F64, device=dev)
            sign[1::2] = -1.0
            fold = bool(torch.equal(P_t.flip(0), P_t * sign[None, :]))
        if fold:
            nrt = 2 * ((-(-(Nw // 2) // 8) + 1) // 2 * 2)
        self.fold = fold
        npt = (nb + 7) // 8
        nplanes = 2 if cplx else 1
        nchunks = nrt // 4
        NT = Lp // 8
        # problem tiles per warp in the pass kernel; one CTA = 4 warps = 4*mt tiles of 8 problems
        # ---- launch configuration (measured on B200, tools/cfg3_sweep.py)
        #  * fused x-update + pass, whole columns per CTA, 2 tiles per warp: when the batch gives many
        #    (nearly) full waves of 444 CTAs;
        #  * otherwise the balanced decomposition: the group-chunks are cut into equal contiguous pieces,
        #    one per resident CTA slot, the x-update is a separate kernel that sums the partial V slots
        #    (so the pieces per tile group are capped: every slot costs the x-update a dependent load).
        # resident CTAs of the pass kernel (shared-memory bound): 3 per SM for L <= 40 (444 on B200), 2 for L <= 64
        SLOTS = {16: 4, 40: 3, 64: 2}[Lp] * _lib.device_info()[0]
        nbal = 0
        if nsplit is None and nbal_req is None:
            waves = -(-npt // 8) / SLOTS
            fused_ok = Lp <= 40 and waves >= 1.0 and waves / np.ceil(waves) >= 0.85
            if mt is None:
                mt = 2 if (Lp <= 40 and (fused_ok or npt >= 1024)) else 1
                if fused_ok and os.environ.get("ADMM_SPM_MT"):      # (A/B runs of the large-batch step kernel)
                    mt = int(os.environ["ADMM_SPM_MT"])
            ngroups = -(-npt // (4 * mt))
            if fused_ok and (mt == 2 or os.environ.get("ADMM_SPM_MT")):
                nsplit = 1
            else:
                total = ngroups * nchunks
                # pieces per tile group: up to 16 (measured with the fused balanced step, tools/bal_sweep.py: one wave
                # of pieces as soon as there are 28 groups; with the separate x-update kernel of round 1, where every
                # partial-sum slot cost a dependent load, the optimum was 4 to 8)
                nbal = max(1, min(SLOTS, total, 16 * ngroups))
                nsplit = -(-nchunks * nbal // total) + 1
        else:
            if mt is None:
                mt = 1
            ngroups = -(-npt // (4 * mt))
            if nbal_req is not None:
                total = ngroups * nchunks
                nbal = max(1, min(int(nbal_req), total))
                nsplit = -(-nchunks * nbal // total) + 1
            else:
                nsplit = max(1, min(nsplit, nchunks))
        assert mt in (1, 2) and (mt == 1 or Lp <= 40)
        # Balanced decomposition over a full wave: the CTAs of one SM do not advance evenly (the warp schedulers favour
        # the CTAs that became resident first: measured on cfg3, the third CTA of an SM finished its equal share 13 us
        # after the first), so later CTAs get shorter pieces -- weights 1 + skew, ..., 1 - skew over the residency tiers.
        self._bal_tables = None
        explicit_skew = bal_skew is not None          # (tests: tables for any number of pieces)
        if bal_skew is None:
            bal_skew = float(os.environ.get("ADMM_BAL_SKEW", self.BAL_SKEW))
        per_sm = SLOTS // _lib.device_info()[0]
        if (nbal == SLOTS or (explicit_skew and nbal >= per_sm)) and per_sm > 1 and bal_skew != 0.0:
            total = ngroups * nchunks
            tier = np.minimum(np.arange(nbal) // max(1, nbal // per_sm), per_sm - 1)
            w = np.round(1024 * (1.0 + bal_skew * (1.0 - 2.0 * tier / (per_sm - 1)))).astype(np.int64)
            cum = np.concatenate([[0], np.cumsum(w)])
            bounds = (total * cum) // cum[-1]
            if np.all(np.diff(bounds) >= 1):
                gstart = np.arange(ngroups, dtype=np.int64) * nchunks
                first = np.searchsorted(bounds, gstart, side="right") - 1
                last = np.searchsorted(bounds, gstart + nchunks - 1, side="right") - 1
                nsplit = max(nsplit, int((last - first + 1).max()))
                self._bal_tables = (torch.from_numpy(bounds.astype(np.int64)).to(dev),
                                    torch.from_numpy(first.astype(np.int32)).to(dev))
        self.dims = SpmDims(L, Lp, Nw, nrt, nb, npt, nplanes, nsplit, mt, nbal, int(batch_wide), nc, int(fold))
        self.nc = nc
        self.batch_wide = bool(batch_wide)
        self.lam, self.max_mu = float(lam), float(max_mu)
        nprob = 8 * npt
        nct = npt * nplanes
        z = lambda *shape, dtype=_F64: torch.zeros(*shape, dtype=dtype, device=dev)
        dref = C.byref(self.dims)

        # shared operators
        self.Pf = z(nrt * 2 * NT * 64)
        call("admm_spm_prepare_P", dref, ptr(P_t), L, ptr(self.Pf), stream())
        self.P = P_t
        self.PtP = z(Lp, Lp)
        ptp = z(L, L)
        call("admm_gemm", 0, _lib.OP_T, L, L, Nw, ptr(P_t), L, ptr(P_t), L, ptr(ptp), L, stream())
        self.PtP[:L, :L] = ptp
        self.PtPf = z(Lp * Lp)
        call("admm_spm_pack_operator", dref, ptr(self.PtP), ptr(self.PtPf), stream())
        self.G0 = z(Lp, Lp)
        self.G0[:L, :L] = G0
        self.Cvec = z(nc, Lp)
        self.Cvec[:, :L] = Cv
        self._nslots = self.CACHE_SLOTS
        self.Ginv_cache = z(self._nslots, Lp, Lp)
        self.w_cache = z(self._nslots, nc, Lp)
        self.sigma_cache = z(self._nslots, nc * nc)
        self._slot_of = {}
        # per problem
        self.slot = torch.zeros(nprob, dtype=torch.int32, device=dev)
        self.mu10 = torch.full((nprob,), float(mu10), dtype=_F64, device=dev)
        self.mu20 = torch.full((nprob,), float(mu20), dtype=_F64, device=dev)
        self.mu20_used = self.mu20.clone()
        self.done = torch.zeros(nprob, dtype=torch.int32, device=dev)
        self.done[nb:] = 1
        self.iters = torch.zeros(nprob, dtype=torch.int32, device=dev)
        self.last_res = z(nprob, 2)
        self.Dre = z(nplanes, nc, nprob)
        self.Dre[0, :, :nb] = D_t.real if D_t.is_complex() else D_t.to(_F64)
        if cplx and D_t.is_complex():
            self.Dre[1, :, :nb] = D_t.imag
        # fragment-layout arrays
        fl = nct * NT * 64
        self.b0 = z(fl)
        call("admm_spm_pack_L", dref, ptr(b0.contiguous()), int(b0.is_complex()), ptr(self.b0), stream())
        self.x0f, self.x1f, self.h10f, self.y0f = z(fl), z(fl), z(fl), z(fl)
        # x0 at the start of the last executed iteration (`_x_old[0]` of the reference, optimizer.py:324): kept
        # only on request (the drop-in SimpleOptimizer asks for it) -- one more L-vector store per iteration
        self.x0_oldf = z(fl) if keep_x_old else None
        self.V = z(nsplit * fl)
        self.aim = z(fl)                  # sum_k mu20_k Im(x0_k)  (imaginary-plane tiles only)
        self._him_base = None             # Im(h20) at the time of set_state (None == 0)
        gt = 4 * mt                       # problem tiles per pass CTA; S is padded to whole CTAs
        self.S = z(-(-npt // gt) * gt * nrt * 64)
        self.normsA = z(nct * 8 * 8)
        self.normsB = z(nsplit * nct * 8 * 2)
        self.gsum = z(16)
        self.gpart = z(256 * 16 + (16 * 512 + 64 if os.environ.get('ADMM_B200_LIB') else 0))      # (+ trace slots of tools/pass_trace.py)
        # lazy batch-wide iterations (one launch per iteration: reduction in the tail of the step / pass kernel, decision
        # in the head of the next one): per-CTA partial sums and the control words
        ngrp = -(-npt // gt)
        n_pass = nbal if nbal > 0 else ngrp * nsplit
        self.cta_partA = z(max(ngrp, -(-nct // 4), n_pass) * 10)      # (fused balanced step: one row per pass CTA)
        self.cta_partB = z(n_pass * 2)
        self.lazy = torch.zeros(4, dtype=torch.int32, device=dev)
        self.xready = torch.zeros(max(1, ngrp), dtype=torch.int32, device=dev)
        # 1: fused step with whole columns per CTA; 2: fused step on the balanced decomposition (owner CTAs run the
        # x-update, the whole iteration of a small batch is one launch); 0: x-update kernel + pass kernel
        self._step_mode = int(_lib.lib.admm_spm_step_supported(C.byref(self.dims)))
        self.use_lazy = True          # tests switch it off to compare with the three-kernel iteration
        self._no_solo_bw = False      # set after a co-residency failure of the batch-wide cluster-resident solve
        self._inject_coresidency_failure = False
        self._lazy_pending = False    # a launched lazy iteration still awaits its decision (next kernel's head or flush)
        self.iter_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.flags = torch.zeros(4, dtype=torch.int32, device=dev)
        self.history = None
        self.pass_events = None      # bench.py: list collecting (start, stop) CUDA events around the pass kernel
        self._v_valid = False        # V = P^T(h20 + mu20 x2) matches the current state and mu20
        self._fresh = True           # y0 = P^T P x0 must be recomputed before the next x-update
        self._graphs = {}            # CUDA graphs of runs of plain iterations, keyed by (length, baked args)
        self.primal_residual = []
        self.dual_residual = []
        self.bufs = SpmBuffers()
        self._fill_bufs(1e-12)
        self._refresh_slots(initial=True)

    def _fill_bufs(self, rtol, fact_incr=2.0, th_change=10.0):
        b = self.bufs
        for name, t in (("Pf", self.Pf), ("PtPf", self.PtPf), ("Cvec", self.Cvec), ("Ginv_cache", self.Ginv_cache),
                        ("w_cache", self.w_cache), ("sigma_cache", self.sigma_cache), ("slot", self.slot),
                        ("mu10", self.mu10), ("mu20", self.mu20), ("mu20_used", self.mu20_used), ("done", self.done),
                        ("iters", self.iters), ("last_res", self.last_res), ("Dre", self.Dre), ("b0", self.b0),
                        ("x0", self.x0f), ("x1", self.x1f), ("h10", self.h10f), ("y0", self.y0f), ("V", self.V),
                        ("aim", self.aim), ("S", self.S), ("normsA", self.normsA), ("normsB", self.normsB), ("gsum", self.gsum),
                        ("gpart", self.gpart), ("cta_partA", self.cta_partA), ("cta_partB", self.cta_partB),
                        ("lazy", self.lazy), ("xready", self.xready), ("iter_counter", self.iter_counter),
                        ("flags", self.flags)):
            setattr(b, name, t.data_ptr())
        b.x0_old = self.x0_oldf.data_ptr() if self.x0_oldf is not None else None
        b.bal_bounds, b.bal_first = ((self._bal_tables[0].data_ptr(), self._bal_tables[1].data_ptr())
                                     if self._bal_tables is not None else (None, None))
        b.history = self.history.data_ptr() if self.history is not None else None
        b.hist_cap = int(self.history.shape[0]) if self.history is not None else 0
        b.lam, b.rtol, b.max_mu = self.lam, float(rtol), self.max_mu
        b.fact_incr, b.th_change = float(fact_incr), float(th_change)

    # ------------------------------------------------------------------ factor cache
    def _refresh_slots(self, initial: bool = False) -> None:
        """Map every problem's (mu10, mu20) to a factor-cache row; factor the new pairs."""
        nb = self.nb
        if self.batch_wide or nb == 1:
            pairs = torch.stack([self.mu10[:1], self.mu20[:1]], dim=1).cpu().numpy()
            inverse = None
        else:
            # distinct (mu10, mu20) pairs through three 1-D sorts (torch.unique(dim=0) on a (nb, 2) array is
            # orders of magnitude slower and this runs once per interval on up to 2^20 problems)
            u10, i10 = torch.unique(self.mu10[:nb], return_inverse=True)
            u20, i20 = torch.unique(self.mu20[:nb], return_inverse=True)
            n20 = u20.numel()
            uk, inverse = torch.unique(i10 * n20 + i20, return_inverse=True)
            pairs = torch.stack([u10[uk // n20], u20[uk % n20]], dim=1).cpu().numpy()
        new = [(float(a), float(b)) for a, b in pairs if (float(a), float(b)) not in self._slot_of]
        if len(self._slot_of) + len(new) > self._nslots:
            # evict everything not currently needed; if the pairs in use alone exceed the cache, grow it (per-problem
            # mode: mu10 and mu20 each walk a ladder 0.1 * 2^k up to max_mu, so a large batch can need a few hundred)
            needed = {(float(a), float(b)) for a, b in pairs}
            self._slot_of = {k: v for k, v in self._slot_of.items() if k in needed}
            new = [k for k in needed if k not in self._slot_of]
            if len(self._slot_of) + len(new) > self._nslots:
                self._grow_cache(len(self._slot_of) + len(new))
        if new:
            free = [i for i in range(self._nslots) if i not in set(self._slot_of.values())]
            slots = free[:len(new)]
            for k, sl in zip(new, slots):
                self._slot_of[k] = sl
            dev = self.device
            sl_t = torch.tensor(slots, dtype=torch.int32, device=dev)
            m10 = torch.tensor([k[0] for k in new], dtype=_F64, device=dev)
            m20 = torch.tensor([k[1] for k in new], dtype=_F64, device=dev)
            info = torch.zeros(len(new), dtype=torch.int32, device=dev)
            call("admm_spm_factor", C.byref(self.dims), len(new), ptr(sl_t), ptr(m10), ptr(m20), ptr(self.G0),
                 ptr(self.PtP), ptr(self.Cvec), ptr(self.Ginv_cache), ptr(self.w_cache), ptr(self.sigma_cache),
                 ptr(info), stream())
            if int(info.max().item()) != 0:
                raise _lib.AdmmError("alpha A^H A + mu is not positive definite")
        if inverse is None:
            self.slot.fill_(self._slot_of[(float(pairs[0][0]), float(pairs[0][1]))])
        else:
            lut = torch.tensor([self._slot_of[(float(a), float(b))] for a, b in pairs], dtype=torch.int32,
                               device=self.device)
            self.slot[:nb] = lut[inverse]

    def _grow_cache(self, need: int) -> None:
        """Reallocate the factor cache with room for ``need`` pairs (entries in use keep their rows)."""
        n = self._nslots
        while n < need:
            n *= 2
        Lp = self.dims.Lp
        for name, shape in (("Ginv_cache", (n, Lp, Lp)), ("w_cache", (n, self.nc, Lp)), ("sigma_cache", (n, self.nc * self.nc))):
            old = getattr(self, name)
            t = torch.zeros(*shape, dtype=_F64, device=self.device)
            t[:old.shape[0]] = old
            setattr(self, name, t)
        self._nslots = n
        self._graphs.clear()            # captured launches bake the old cache addresses
        b = self.bufs
        b.Ginv_cache, b.w_cache, b.sigma_cache = (self.Ginv_cache.data_ptr(), self.w_cache.data_ptr(),
                                                  self.sigma_cache.data_ptr())

    # ------------------------------------------------------------------ data reload (same operators)
    def reset(self, g=None, mu: Optional[float] = None) -> None:
        """Zero the ADMM state (x, h), optionally load new data ``g`` (L x nb, device or host) and
        reset the penalties -- lets one plan serve many batches without re-allocating."""
        for t in (self.x0f, self.x1f, self.h10f, self.S, self.V, self.aim):
            t.zero_()
        self._him_base = None
        if mu is not None:
            self.mu10.fill_(float(mu))
            self.mu20.fill_(float(mu))
            self.mu20_used.fill_(float(mu))
            self._refresh_slots()
        if g is not None:
            if self._s is None:
                raise NotImplementedError("reset(g=...) needs the SpM form (s, g)")
            g_t = _dev_tensor(g, self.device)
            if g_t.ndim == 1:
                g_t = g_t[:, None]
            g_t = g_t.contiguous()
            cplx = g_t.is_complex()
            assert g_t.shape == (self.L, self.nb) and (self.is_complex or not cplx)
            gv = g_t if cplx else g_t.to(_F64)
            b0 = torch.empty_like(gv)
            sd = (-self._alpha * self._s).to(gv.dtype)
            call("admm_diag_mul", int(cplx), self.L, self.L, self.nb, ptr(sd), ptr(gv), self.nb, ptr(b0), self.nb, stream())
            call("admm_spm_pack_L", C.byref(self.dims), ptr(b0), int(cplx), ptr(self.b0), stream())
            self._g = gv
        self._v_valid = True          # all-zero state: V = 0
        self._fresh = True
        self.primal_residual, self.dual_residual = [], []

    # ------------------------------------------------------------------ state import / export
    def set_state(self, x0=None, x1=None, x2=None, h10=None, h20=None) -> None:
        """Load canonical (rows x nb) state (complex or real).  Raises ``NotImplementedError`` if
        (h20, x2) is not representable by the implicit complementarity form."""
        dref = C.byref(self.dims)
        for src, dst in ((x0, self.x0f), (x1, self.x1f), (h10, self.h10f)):
            if src is not None:
                t = _dev_tensor(src, self.device).reshape(self.L, self.nb).contiguous()
                if not self.is_complex and t.is_complex():
                    if float(t.imag.abs().max().item()) != 0.0:
                        raise NotImplementedError("complex state on a real-data SpM plan")
                    t = t.real.contiguous()
                call("admm_spm_pack_L", dref, ptr(t), int(t.is_complex()), ptr(dst), stream())
        if x2 is not None or h20 is not None:
            zer = lambda: torch.zeros(self.Nw, self.nb, dtype=_F64, device=self.device)
            h = _dev_tensor(h20, self.device).reshape(self.Nw, self.nb) if h20 is not None else zer()
            x = _dev_tensor(x2, self.device).reshape(self.Nw, self.nb) if x2 is not None else zer()
            cp = h.is_complex() or x.is_complex()
            if cp:
                h, x = h.to(_C128).contiguous(), x.to(_C128).contiguous()
            flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            call("admm_spm_pack_state", dref, ptr(h), ptr(x), int(cp), ptr(self.mu20), ptr(self.S), ptr(flag), stream())
            if int(flag.item()) != 0:
                raise NotImplementedError("(h20, x2) state is not complementary; use the generic executor")
            self.mu20_used.copy_(self.mu20)
            # imaginary part of h20: kept as a base plane + the L-space accumulators (z, a)
            self.aim.zero_()
            him = h.imag.contiguous() if cp else None
            if him is not None and float(him.abs().max().item()) == 0.0:
                him = None
            self._him_base = him
            if self.is_complex:
                zc = torch.zeros(self.L, self.nb, dtype=_C128, device=self.device)
                if him is not None:
                    zr = torch.empty(self.L, self.nb, dtype=_F64, device=self.device)
                    call("admm_gemm", 0, _lib.OP_T, self.L, self.nb, self.Nw, ptr(self.P), self.L, ptr(him), self.nb,
                         ptr(zr), self.nb, stream())
                    zc = torch.complex(torch.zeros_like(zr), zr)
                self._set_imag_tiles(self.V, zc)
            self._v_valid = False
        self._fresh = True

    def _set_imag_tiles(self, frag_arr: torch.Tensor, canon_c: torch.Tensor) -> None:
        """Write Im(canon_c) (L x nb) into the imaginary-plane tiles of a fragment array (split 0)."""
        NT = self.dims.Lp // 8
        fl = self.dims.npt * self.dims.nplanes * NT * 64
        tmp = torch.empty(fl, dtype=_F64, device=self.device)
        call("admm_spm_pack_L", C.byref(self.dims), ptr(canon_c.contiguous()), 1, ptr(tmp), stream())
        frag_arr[:fl].view(self.dims.npt, 2, NT * 64)[:, 1, :] = tmp.view(self.dims.npt, 2, NT * 64)[:, 1, :]

    def _unpack_L(self, frag) -> np.ndarray:
        out = torch.empty(self.L, self.nb, dtype=_C128, device=self.device)
        call("admm_spm_unpack_L", C.byref(self.dims), ptr(frag), ptr(out), 1, stream())
        return out.cpu().numpy()

    def x0(self) -> np.ndarray:
        return self._unpack_L(self.x0f)

    def x1(self) -> np.ndarray:
        return self._unpack_L(self.x1f)

    def h10(self) -> np.ndarray:
        return self._unpack_L(self.h10f)

    def x0_old(self) -> np.ndarray:
        """x0 at the start of the last executed iteration (needs ``keep_x_old=True``)."""
        if self.x0_oldf is None:
            raise _lib.AdmmError("x0_old() needs SharedSpM(..., keep_x_old=True)")
        return self._unpack_L(self.x0_oldf)

    def _unpack_state(self):
        h = torch.empty(self.Nw, self.nb, dtype=_C128, device=self.device)
        x = torch.empty(self.Nw, self.nb, dtype=_C128, device=self.device)
        him = None
        if self.is_complex:
            # Im(h20) = Im(h20)_base - P a,   a = sum_k mu20_k Im(x0_k)
            a_c = torch.empty(self.L, self.nb, dtype=_C128, device=self.device)
            call("admm_spm_unpack_L", C.byref(self.dims), ptr(self.aim), ptr(a_c), 1, stream())
            a_im = a_c.imag.contiguous()
            Pa = torch.empty(self.Nw, self.nb, dtype=_F64, device=self.device)
            call("admm_gemm", 0, _lib.OP_N, self.Nw, self.nb, self.L, ptr(self.P), self.L, ptr(a_im), self.nb,
                 ptr(Pa), self.nb, stream())
            him = torch.empty_like(Pa)
            base = self._him_base
            n = Pa.numel()
            if base is None:
                call("admm_axpby", n, -1.0, ptr(Pa), 0.0, None, ptr(him), stream())
            else:
                call("admm_axpby", n, -1.0, ptr(Pa), 1.0, ptr(base), ptr(him), stream())
        call("admm_spm_unpack_state", C.byref(self.dims), ptr(self.S), ptr(self.mu20_used), ptr(him), ptr(h), ptr(x), 1,
             stream())
        return h, x

    def x2(self) -> np.ndarray:
        return self._unpack_state()[1].cpu().numpy()

    def h20(self) -> np.ndarray:
        return self._unpack_state()[0].cpu().numpy()

    def x0_device(self) -> torch.Tensor:
        """(L, nb) complex128 result on the device (no host copy)."""
        out = torch.empty(self.L, self.nb, dtype=_C128, device=self.device)
        call("admm_spm_unpack_L", C.byref(self.dims), ptr(self.x0f), ptr(out), 1, stream())
        return out

    def set_mu(self, mu10, mu20) -> None:
        """Overwrite the penalties (scalars or per-problem arrays); re-encodes the implicit state."""
        h, x = self._unpack_state()
        self.mu10[:self.nb] = torch.as_tensor(mu10, dtype=_F64, device=self.device)
        self.mu20[:self.nb] = torch.as_tensor(mu20, dtype=_F64, device=self.device)
        self.set_state(h20=h, x2=x)
        self._refresh_slots()

    # ------------------------------------------------------------------ iteration
    def _lazy_ok(self) -> bool:
        """Batch-wide criterion without NCCL between the kernels: plain iterations are ONE launch (fused path) or two."""
        return self.batch_wide and self.use_lazy and (self.group is None or self._peer is not None)

    def _comm_ref(self):
        return C.byref(self._peer.comm) if self._peer is not None else None

    def _flush(self) -> None:
        """Take the pending decision of the last lazy iteration (history, iteration count, stopping test)."""
        if self._lazy_pending:
            call("admm_spm_flush", C.byref(self.dims), C.byref(self.bufs), self._comm_ref(), stream())
            self._lazy_pending = False

    def _iteration(self, do_update_mu: bool, allow_lazy: bool = True) -> None:
        dref, bref, st = C.byref(self.dims), C.byref(self.bufs), stream()
        lazy = allow_lazy and not do_update_mu and self._lazy_ok()
        if not lazy:
            self._flush()
        if not self._v_valid:
            # V = P^T(h20 + mu20 x2) from the current state (after set_state or a change of mu20)
            call("admm_spm_pass", dref, bref, 1, st)
            self._v_valid = True
        if self._fresh:
            call("admm_spm_refresh_y", dref, bref, st)      # y0 = P^T P x0 for the loaded x0
            self._fresh = False
        timed = self.pass_events is not None and not do_update_mu
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fused = self._step_mode != 0
        if lazy:
            # the whole iteration: decision of the previous one in the head, reduction of this one in the tail
            cref, pend = self._comm_ref(), int(self._lazy_pending)
            if fused:
                if timed:
                    e0.record()
                call("admm_spm_step_lazy", dref, bref, cref, pend, st)
            else:
                call("admm_spm_xupdate_lazy", dref, bref, cref, pend, st)
                if timed:
                    e0.record()
                call("admm_spm_pass_lazy", dref, bref, cref, st)
            if timed:
                e1.record()
                self.pass_events.append((e0, e1))
            self._lazy_pending = True
            return
        if fused:
            if timed:
                e0.record()
            call("admm_spm_step", dref, bref, st)
        else:
            call("admm_spm_xupdate", dref, bref, st)
            if timed:
                e0.record()
            call("admm_spm_pass", dref, bref, 0, st)
        if timed:
            e1.record()
            self.pass_events.append((e0, e1))
        if self.batch_wide and self.group is None:
            call("admm_spm_reduce_decide", dref, bref, int(do_update_mu), st)
            return
        if self.batch_wide and self._peer is not None:
            cref = C.byref(self._peer.comm)
            call("admm_spm_reduce_post", dref, bref, cref, st)
            call("admm_spm_decide_peer", dref, bref, cref, int(do_update_mu), st)
            return
        if self.batch_wide:
            call("admm_spm_reduce", dref, bref, st)
            torch.distributed.all_reduce(self.gsum, group=self.group)
        call("admm_spm_decide", dref, bref, int(do_update_mu), st)

    def _after_update_iteration(self) -> bool:
        """Host look at the device flags after an iteration that may have changed mu or finished
        problems.  Returns True when every problem is done."""
        self._flush()
        fl = self.flags.cpu()
        if int(fl[2]) == -3:
            raise _lib.CoResidencyError("fused balanced step: the CTAs of the launch were not co-resident (another kernel "
                                        "holding SMs?)")
        if int(fl[2]) == -2:
            raise _lib.AdmmError("sharded batch-wide criterion: a peer rank never posted its residual sums "
                                 "(watchdog of admm_spm_decide_peer); the state is undefined")
        if int(fl[1]) >= self.nb:
            return True
        if int(fl[0]) != 0:
            self.flags[0] = 0
            self._refresh_slots()
            self._v_valid = False      # V was built with the old mu20
        return False

    #: iterations per captured CUDA graph (see _replay_plain)
    GRAPH_CHUNK = 32

    #: largest batch the cluster-resident single-launch solve is used for (one 8-CTA cluster per problem, 16 clusters
    #: per wave): measured break-even against the batch kernels at ~200 problems (tools/solo_nb_sweep.py:
    #: 32 problems 3.7x, 64: 2.5x, 128: 1.5x faster, 256: 0.9x)
    SOLO_MAX_NB = 128

    # ------------------------------------------------------------------ co-residency safety net
    _SNAP_TENSORS = ("x0f", "x1f", "h10f", "y0f", "V", "aim", "S", "mu10", "mu20", "mu20_used", "slot", "done", "iters",
                     "last_res", "x0_oldf")

    def _snapshot(self):
        snap = {n: getattr(self, n).clone() for n in self._SNAP_TENSORS if getattr(self, n) is not None}
        snap["_host"] = (self._v_valid, self._fresh, len(self.primal_residual), len(self.dual_residual), dict(self._slot_of))
        return snap

    def _restore(self, snap) -> None:
        for n in self._SNAP_TENSORS:
            if n in snap:
                getattr(self, n).copy_(snap[n])
        self._v_valid, self._fresh, npr, ndu, slot_of = snap["_host"]
        del self.primal_residual[npr:], self.dual_residual[ndu:]
        # factor rows computed meanwhile stay valid (they are keyed by the penalties); keep the larger map
        for k, v in slot_of.items():
            self._slot_of.setdefault(k, v)
        self.flags.zero_()
        self.lazy[:3].zero_()
        self._lazy_pending = False

    def _risky_launch_mode(self, use_solo: Optional[bool], callback) -> bool:
        """Will this solve use a kernel whose CTAs wait for each other, launched WITHOUT the driver's co-residency
        guarantee (plain launch, or not yet probed on this device)?"""
        bal = self._step_mode == 2 and int(_lib.lib.admm_spm_launch_mode(0)) not in (1, 2)
        solo_bw = (self.batch_wide and 1 < self.nb <= self.SOLO_MAX_NB and callback is None and self.group is None
                   and use_solo is not False and not self._no_solo_bw
                   and int(_lib.lib.admm_spm_launch_mode(1)) != 2)
        return bal or solo_bw

    def solve(self, niter: int = 10000, interval_update_mu: int = 100, rtol: float = 1e-12,
              callback=None, keep_history: Optional[bool] = None, use_graph: Optional[bool] = None,
              use_solo: Optional[bool] = None) -> int:
        """Run up to ``niter`` iterations with the ordering of ``SimpleOptimizer.solve``
        (optimizer.py:302-320).  Returns the number of iterations launched.

        Two launch shapes let CTAs wait for each other inside one launch (the fused balanced step of a small batch and
        the batch-wide cluster-resident solve).  They are launched cooperatively, so the driver guarantees that their
        CTAs are resident together.  Where that is not available (``admm_spm_launch_mode`` 3, or not probed yet) the
        state is saved first; should the in-kernel watchdog ever give up (``CoResidencyError``), the state is restored
        and the solve repeated with the kernels that need no co-residency -- the caller never sees an undefined state."""
        snap = self._snapshot() if self._risky_launch_mode(use_solo, callback) else None
        try:
            n = self._solve_impl(niter, interval_update_mu, rtol, callback, keep_history, use_graph, use_solo)
            if self._inject_coresidency_failure:      # tests: exercise the recovery path
                self._inject_coresidency_failure = False
                raise _lib.CoResidencyError("injected")
            return n
        except _lib.CoResidencyError:
            if snap is None:
                raise
            import warnings
            warnings.warn("admmsolver_b200: CTAs of a single-launch iteration were not co-resident (another kernel holding "
                          "SMs?); repeating the solve with the multi-kernel iteration", RuntimeWarning)
            torch.cuda.synchronize()
            self._restore(snap)
            if self._step_mode == 2:
                self._step_mode = 0
            self._no_solo_bw = True
            self._graphs.clear()
            return self._solve_impl(niter, interval_update_mu, rtol, callback, keep_history, use_graph,
                                    False if (self.batch_wide and self.nb > 1) else use_solo)

    def _solve_impl(self, niter, interval_update_mu, rtol, callback, keep_history, use_graph, use_solo) -> int:
        """The loop of ``solve`` (see there).

        A single problem (or a handful with the per-problem criterion) and no callback: the whole
        loop is ONE launch of the cluster-resident kernel (``admm_spm_solo``; ``use_solo``).

        The iterations between two mu updates are data-independent launch sequences: they are
        captured once in a CUDA graph and replayed (``use_graph``; default: on when no callback
        is given and no bench timing events are being collected)."""
        nb = self.nb
        self.iters.zero_()
        self.done[:nb] = 0
        self.flags.zero_()
        self.iter_counter.zero_()
        self.lazy[:3].zero_()           # (entry 3, the launch sequence number of the fused balanced step, runs on)
        self._lazy_pending = False
        track = (self.batch_wide or nb == 1) if keep_history is None else keep_history
        if track:
            if self.history is None or self.history.shape[0] < niter:
                self.history = torch.zeros(max(niter, 1), 2, dtype=_F64, device=self.device)
        elif self.history is not None:
            self.history = None
        self._fill_bufs(rtol)
        key = (float(rtol), self.bufs.history, self.bufs.hist_cap)     # everything a captured launch bakes in
        if use_graph is None:
            use_graph = (callback is None and self.pass_events is None
                         and (self.group is None or not self.batch_wide or self._peer is not None))
        solo_ok = (callback is None and self.pass_events is None and self.group is None and nb <= self.SOLO_MAX_NB
                   and not (self._no_solo_bw and self.batch_wide and nb > 1)
                   and _lib.lib.admm_spm_solo_supported(C.byref(self.dims)) != 0)     # batch-wide: co-resident clusters only
        if use_solo is None:
            use_solo = solo_ok and use_graph
        elif use_solo and not solo_ok:
            raise NotImplementedError("the cluster-resident solve needs nb <= %d (batch-wide criterion: as many clusters "
                                      "as fit the GPU at once), no callback and an operator that fits the cluster"
                                      % self.SOLO_MAX_NB)
        launched = 0
        it = 0
        if use_solo:
            call("admm_spm_solo", C.byref(self.dims), C.byref(self.bufs), ptr(self.G0), ptr(self.PtP), int(niter),
                 int(interval_update_mu), stream())
            self._v_valid, self._fresh = True, False
            fl = torch.cat([self.flags, self.iters[:nb]]).cpu()        # one read-back: flags and iteration counts
            if int(fl[2]) < 0:
                raise _lib.CoResidencyError("cluster-resident solve: the clusters of the batch did not become co-resident "
                                            "(another kernel holding SMs?)")
            if int(fl[2]) != 0:
                raise _lib.AdmmError("alpha A^H A + mu is not positive definite")
            if int(fl[0]) != 0:
                self.flags[0] = 0
                self._refresh_slots()
            launched = int(fl[4:].max())
            solo_iters0 = int(fl[4])
            it = niter
        while it < niter:
            upd = (it % interval_update_mu == 0)
            run = 1 if upd else min(niter, (it // interval_update_mu + 1) * interval_update_mu) - it
            if upd or callback is not None or not use_graph or run < 4:
                self._iteration(upd, allow_lazy=callback is None)
                run = 1
            else:
                self._replay_plain(run, key)
            launched += run
            it += run
            if callback is not None:
                callback()
            if upd or callback is not None or it >= niter:
                if self._after_update_iteration():
                    break
        self._flush()
        if self._lazy_ok() and not use_solo:
            # lazy iterations maintain entry 0 of the per-problem bookkeeping only: the batch shares it
            self.iters[:nb] = self.iters[0].clone()
            self.last_res[:nb] = self.last_res[0].clone()
            if int(self.lazy[2].item()) != 0:
                self.done[:nb] = 1
        if track:
            ndone = solo_iters0 if use_solo else int(self.iters[0].item())
            hist = self.history[:ndone].cpu().numpy()
            self.primal_residual.extend(hist[:, 0].tolist())
            self.dual_residual.extend(hist[:, 1].tolist())
        return launched

    def _replay_plain(self, run: int, key) -> None:
        """``run`` iterations without mu update as one CUDA-graph launch (captured on first use)."""
        if self._fresh:
            self._iteration(False)
            run -= 1
        elif not self._v_valid:
            # one-off launch, kept outside the graph: V from the state with the new mu20
            call("admm_spm_pass", C.byref(self.dims), C.byref(self.bufs), 1, stream())
            self._v_valid = True
        # graphs of GRAPH_CHUNK iterations, replayed as often as they fit (capturing and instantiating one graph of
        # all ~interval iterations costs 50-800 ms on the first solve of a plan -- more than a short solve itself);
        # the few iterations left over are launched eagerly
        lazy = self._lazy_ok()
        while run > 0:
            n = min(run, self.GRAPH_CHUNK)
            if n < 8:
                for _ in range(n):
                    self._iteration(False)
                run -= n
                continue
            self._flush()                           # a captured chunk starts without and ends without a pending decision
            graph = self._graphs.get((n, key, lazy))
            if graph is None:
                if len(self._graphs) >= 8:
                    self._graphs.clear()
                graph = torch.cuda.CUDAGraph()
                before = _lib.launch_count
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    for _ in range(n):
                        self._iteration(False)
                    self._flush()
                _lib.launch_count = before          # capture enqueues nothing
                self._graphs[(n, key, lazy)] = graph
            graph.replay()
            _lib.launch_count += n * self._launches_per_iteration() + (1 if lazy else 0)
            run -= n

    def _launches_per_iteration(self) -> int:
        data = 1 if self._step_mode != 0 else 2
        if self._lazy_ok():
            return data
        if self.batch_wide and (self.group is None or self._peer is not None):
            return data + 2
        return data + (2 if self.batch_wide else 0) + 1

    # ------------------------------------------------------------------ objective
    def objective(self) -> float:
        """alpha ||g + s x0||^2 + lam |x1|_1 summed over the batch (SpM form only)."""
        if self._s is None:
            raise NotImplementedError("objective() needs the SpM form (s, g)")
        x0 = self.x0_device()
        cplx = self._g.is_complex()
        xs = x0 if cplx else x0.real.contiguous()
        sx = torch.empty_like(xs)
        sd = self._s.to(xs.dtype)
        call("admm_diag_mul", int(cplx), self.L, self.L, self.nb, ptr(sd), ptr(xs), self.nb, ptr(sx), self.nb, stream())
        n = sx.numel() * (2 if cplx else 1)
        msx = torch.empty_like(sx)
        call("admm_axpby", n, -1.0, ptr(sx), 0.0, None, ptr(msx), stream())
        out = torch.zeros(1, dtype=_F64, device=self.device)
        scratch = torch.zeros(1024, dtype=_F64, device=self.device)
        call("admm_sumsq", n, ptr(self._g), ptr(msx), ptr(out), ptr(scratch), stream())
        l1 = float(np.abs(self.x1()).sum())
        return float(self._alpha * out.item() + self.lam * l1)


class SpMHostStream:
    """Batches from pinned host memory through one ``SharedSpM`` plan and back, with the copies OVERLAPPED with the
    solves: data ``g`` of batch k + 1 is uploaded (own stream, second staging buffer) while batch k iterates, the result
    ``x0`` of batch k is downloaded (own stream) while batch k + 1 iterates.  The pattern of a production sweep over many
    batches that share one basis; ``bench.py`` measures its end-to-end leg through it.

        pipe = SpMHostStream(eng)
        for k, g in enumerate(batches):                       # pinned (L, nb) arrays
            pipe.submit(g, out[k], niter, next_g_host=batches[k + 1] if k + 1 < len(batches) else None)
        pipe.join()                                           # the current stream now waits for all copies
    """

    def __init__(self, eng: "SharedSpM"):
        if eng._s is None:
            raise NotImplementedError("SpMHostStream needs a plan built from the SpM form (s, g)")
        self.eng = eng
        dt = _C128 if eng.is_complex else _F64
        self._g = [torch.empty(eng.L, eng.nb, dtype=dt, device=eng.device) for _ in range(2)]
        self._up, self._down = torch.cuda.Stream(), torch.cuda.Stream()
        self._ev_up = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self._k = 0
        self._prefetched = False

    def _upload(self, i: int, g_host: torch.Tensor) -> None:
        self._up.wait_event(self._ev_free[i])                 # the plan has taken the previous content of this buffer
        with torch.cuda.stream(self._up):
            self._g[i].copy_(g_host, non_blocking=True)
            self._ev_up[i].record(self._up)

    def submit(self, g_host: torch.Tensor, out_host: torch.Tensor, niter: int, mu: float = 0.1,
               next_g_host: Optional[torch.Tensor] = None, **solve_kw) -> int:
        """Solve one batch (``reset`` to the zero state with data ``g_host``, ``niter`` iterations), write x0 (L, nb,
        complex128) to ``out_host``.  Returns what ``SharedSpM.solve`` returns."""
        eng, cur = self.eng, torch.cuda.current_stream()
        i = self._k & 1
        if not self._prefetched:
            self._ev_free[i].record(cur)
            self._upload(i, g_host)
        cur.wait_event(self._ev_up[i])
        eng.reset(g=self._g[i], mu=mu)                        # packs g into the plan's fragment layout
        self._ev_free[i].record(cur)
        self._prefetched = next_g_host is not None
        if self._prefetched:
            if self._k == 0:
                self._ev_free[1 - i].record(cur)
            self._upload(1 - i, next_g_host)
        ran = eng.solve(niter, **solve_kw)
        x = eng.x0_device()
        done = torch.cuda.Event()
        done.record(cur)
        self._down.wait_event(done)
        with torch.cuda.stream(self._down):
            out_host.copy_(x, non_blocking=True)
        x.record_stream(self._down)
        self._k += 1
        return ran

    def join(self) -> None:
        """Make the current stream wait for every copy issued so far."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self._up)
        cur.wait_stream(self._down)


class BatchedBasisPursuit:
    """Pattern A on the GPU: ``nb`` independent problems
    ``alpha ||y_b - A_b x0||^2 + lam |x1|_1  s.t. x0 = x1``, each with its own penalty and
    stopping test (one ``SimpleOptimizer`` instance per problem in the reference)."""

    def __init__(self, A, y, alpha: float = 1.0, lam: float = 0.1, mu: float = 1.0, max_mu: float = 1e3,
                 keep_history: bool = False, tiled: bool = True, keep_x_old: bool = False):
        dev = _lib.require_cuda()
        self.device = dev
        A_t = _dev_tensor(A, dev, _F64)
        y_t = _dev_tensor(y, dev, _F64)
        if A_t.ndim == 2:
            A_t, y_t = A_t[None], y_t[None]
        self.A, self.y = A_t.contiguous(), y_t.contiguous()
        nb, M, N = self.A.shape
        self.nb, self.M, self.N = nb, M, N
        self.woodbury = M < N
        nk = M if self.woodbury else N
        self.nk = nk
        self.alpha, self.lam, self.max_mu = float(alpha), float(lam), float(max_mu)
        z = lambda *shape, dtype=_F64: torch.zeros(*shape, dtype=dtype, device=dev)
        self.aty, self.gram, self.Kinv = z(nb, N), z(nb, nk, nk), z(nb, nk, nk)
        self._x0, self._x1, self._h = z(nb, N), z(nb, N), z(nb, N)
        # x0 at the start of the last executed iteration (`_x_old[0]`, optimizer.py:324), on request
        self._x0_old = z(nb, N) if keep_x_old else None
        self.mu = torch.full((nb,), float(mu), dtype=_F64, device=dev)
        self.need_factor = torch.ones(nb, dtype=torch.int32, device=dev)
        self.done = torch.zeros(nb, dtype=torch.int32, device=dev)
        self.iters = torch.zeros(nb, dtype=torch.int32, device=dev)
        self.last_res = z(nb, 2)
        self.info = torch.zeros(nb, dtype=torch.int32, device=dev)
        self.keep_history = keep_history
        self.history = None
        self._kcache = {}            # nb == 1: mu -> (Kinv, ready event, info, keep-alive)
        self._side = None            # side stream of the speculative factorisations
        self.primal_residual = [[] for _ in range(nb)] if keep_history else None
        self.dual_residual = [[] for _ in range(nb)] if keep_history else None
        # tile-major copy of A for the single-sweep kernel (Woodbury path, M <= 256)
        self.At = z(nb, -(-N // 32), M, 32) if (self.woodbury and M <= 256 and tiled) else None
        self.bufs = BpBuffers()
        self._fill(1e-12, 100)
        call("admm_bp_setup", C.byref(self.bufs), ptr(self.y), ptr(self.aty), ptr(self.gram), stream())
        if self.At is not None:
            call("admm_bp_tile_A", C.byref(self.bufs), ptr(self.At), stream())

    def _fill(self, rtol, interval, fact_incr=2.0, th_change=10.0):
        b = self.bufs
        b.nb, b.M, b.N, b.woodbury, b.nk = self.nb, self.M, self.N, int(self.woodbury), self.nk
        b.At = self.At.data_ptr() if self.At is not None else None
        b.x0_old = self._x0_old.data_ptr() if self._x0_old is not None else None
        for name, t in (("A", self.A), ("aty", self.aty), ("gram", self.gram), ("Kinv", self.Kinv), ("x0", self._x0),
                        ("x1", self._x1), ("h", self._h), ("mu", self.mu), ("need_factor", self.need_factor),
                        ("done", self.done), ("iters", self.iters), ("last_res", self.last_res)):
            setattr(b, name, t.data_ptr())
        b.history = self.history.data_ptr() if self.history is not None else None
        b.hist_cap = int(self.history.shape[1]) if self.history is not None else 0
        b.alpha, b.lam, b.rtol, b.max_mu = self.alpha, self.lam, float(rtol), self.max_mu
        b.fact_incr, b.th_change, b.interval_update_mu = float(fact_incr), float(th_change), int(interval)

    def set_data(self, y) -> None:
        """New right-hand sides y (nb x M, device or host) for the same operators A: only alpha A^T y is
        recomputed (the Gram matrices and cached inverses depend on A and mu alone)."""
        y_t = _dev_tensor(y, self.device, _F64).reshape(self.nb, self.M).contiguous()
        self.y = y_t
        call("admm_bp_setup", C.byref(self.bufs), ptr(self.y), ptr(self.aty), None, stream())

    def set_state(self, x0=None, x1=None, h=None, mu=None) -> None:
        for src, dst in ((x0, self._x0), (x1, self._x1), (h, self._h)):
            if src is not None:
                dst.copy_(_dev_tensor(src, self.device, _F64).reshape(self.nb, self.N))
        if mu is not None:
            self.mu[:] = torch.as_tensor(mu, dtype=_F64, device=self.device)
            self.need_factor.fill_(1)

    # ------------------------------------------------------------------ single problem: factor cache by mu
    def _kinv_entry(self, mu: float, speculative: bool):
        """(alpha A A^T + mu)^-1-type factor of the ONE problem for penalty ``mu`` from the per-mu cache (the
        reference's ``_B_cache``, objectivefunc.py:89-96).  A missing entry is computed on the current stream,
        or -- ``speculative`` -- on a side stream while the iterations run: ``update_mu`` can only move mu to
        ``mu * fact_incr`` or ``mu / fact_incr``, and a single problem leaves most of the GPU idle."""
        main = torch.cuda.current_stream()
        ent = self._kcache.get(mu)
        if ent is None:
            if len(self._kcache) >= 12:
                torch.cuda.synchronize()
                self._kcache.clear()
            if speculative:
                if self._side is None:
                    self._side = torch.cuda.Stream()
                self._side.wait_stream(main)
            s_ = self._side if speculative else main
            with torch.cuda.stream(s_):
                K = torch.empty(1, self.nk, self.nk, dtype=_F64, device=self.device)
                mu_t = torch.full((1,), float(mu), dtype=_F64, device=self.device)
                nf = torch.ones(1, dtype=torch.int32, device=self.device)
                info = torch.zeros(1, dtype=torch.int32, device=self.device)
                sh = BpBuffers.from_buffer_copy(self.bufs)
                sh.mu, sh.need_factor, sh.Kinv = mu_t.data_ptr(), nf.data_ptr(), K.data_ptr()
                call("admm_bp_factor", C.byref(sh), ptr(info), C.c_void_p(s_.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(s_)
            ent = (K, ev, info, (mu_t, nf))
            self._kcache[mu] = ent
        if not speculative:
            main.wait_event(ent[1])
        return ent

    def _solve_single(self, niter: int) -> None:
        """nb == 1: one read-back per interval; the factor for the current mu comes from the cache, the two
        factors ``update_mu`` can ask for next are computed meanwhile on idle SMs."""
        bref, st = C.byref(self.bufs), stream()
        mu = float(self.mu[0].item())
        while True:
            ent = self._kinv_entry(mu, speculative=False)
            self.bufs.Kinv = ent[0].data_ptr()
            self.need_factor.zero_()
            fi = float(self.bufs.fact_incr)
            for m2 in (min(mu * fi, self.max_mu), mu / fi):
                if m2 != mu:
                    self._kinv_entry(m2, speculative=True)
            call("admm_bp_iterate", bref, int(niter), st)
            fl = torch.stack([self.iters.to(_F64), self.done.to(_F64), self.mu]).cpu()
            if int(ent[2].item()) != 0:
                raise _lib.AdmmError("alpha A^H A + mu is not positive definite")
            if int(fl[0, 0]) >= niter or int(fl[1, 0]) != 0:
                break
            mu = float(fl[2, 0])

    def solve(self, niter: int = 10000, interval_update_mu: int = 100, rtol: float = 1e-12) -> None:
        """All iterations of every problem on the device; no host synchronisation inside (a single problem:
        one small read-back per mu interval, see ``_solve_single``)."""
        self.iters.zero_()
        self.done.zero_()
        self.history = (torch.zeros(self.nb, max(niter, 1), 2, dtype=_F64, device=self.device)
                        if self.keep_history else None)
        self._fill(rtol, interval_update_mu)
        if self.nb == 1 and niter > 0:
            self._solve_single(niter)
            self._collect_history()
            return
        bref, st = C.byref(self.bufs), stream()
        rounds = -(-niter // interval_update_mu) + 1
        for _ in range(rounds):
            call("admm_bp_factor", bref, ptr(self.info), st)
            call("admm_bp_iterate", bref, int(niter), st)
            if self.nb <= 16:
                # a handful of problems (latency-bound): one small read-back per round instead of enqueueing
                # all ceil(niter / interval) + 1 rounds blindly -- mu changes only a few times per solve
                fl = torch.stack([self.iters, self.done]).cpu()
                if bool(((fl[0] >= niter) | (fl[1] != 0)).all()):
                    break
        # a factor may still be pending for the next solve() call (mu changed on the last iteration)
        self._collect_history()

    def _collect_history(self) -> None:
        if self.keep_history:
            it = self.iters.cpu().numpy()
            hist = self.history.cpu().numpy()
            for b in range(self.nb):
                self.primal_residual[b].extend(hist[b, :it[b], 0].tolist())
                self.dual_residual[b].extend(hist[b, :it[b], 1].tolist())

    def x0(self) -> np.ndarray:
        return self._x0.cpu().numpy()

    def x1(self) -> np.ndarray:
        return self._x1.cpu().numpy()

    def h(self) -> np.ndarray:
        return self._h.cpu().numpy()

    def x0_old(self) -> np.ndarray:
        """x0 at the start of the last executed iteration (needs ``keep_x_old=True``)."""
        if self._x0_old is None:
            raise _lib.AdmmError("x0_old() needs BatchedBasisPursuit(..., keep_x_old=True)")
        return self._x0_old.cpu().numpy()

    def objective(self) -> np.ndarray:
        """alpha ||y - A x0||^2 + lam |x1|_1 per problem."""
        out = np.empty(self.nb)
        r = torch.empty(self.M, 1, dtype=_F64, device=self.device)
        acc = torch.zeros(1, dtype=_F64, device=self.device)
        scratch = torch.zeros(1024, dtype=_F64, device=self.device)
        l1 = self._x1.abs().sum(dim=1).cpu().numpy()
        for b in range(self.nb):
            call("admm_gemm", 0, _lib.OP_N, self.M, 1, self.N, ptr(self.A[b]), self.N, ptr(self._x0[b]), 1, ptr(r), 1, stream())
            call("admm_sumsq", self.M, ptr(self.y[b]), ptr(r), ptr(acc), ptr(scratch), stream())
            out[b] = self.alpha * acc.item() + self.lam * l1[b]
        return out
