"""ctypes binding of ``libadmm_b200.so`` (the C ABI declared in ``include/admm_b200.h``).

The library is built in-tree by ``__graft_entry__.build()``.  There is NO fallback: if the
shared object is missing, importing this module raises, and every compute entry point needs a
CUDA device (PyTorch is used only for device memory, streams and ``torch.distributed``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADMM_B200_LIB") or os.path.join(_HERE, "libadmm_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python __graft_entry__.py).  admmsolver_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

ABI_VERSION = 5
OP_N, OP_T, OP_H = 0, 1, 2


class AdmmError(RuntimeError):
    pass


class CoResidencyError(AdmmError):
    """A kernel whose CTAs wait for each other inside one launch gave up (in-kernel watchdog): its CTAs were not all
    resident at the same time.  Only possible with plain launches (``admm_spm_launch_mode`` 3); ``SharedSpM.solve``
    restores the state it saved before and repeats the solve with the kernels that need no co-residency."""


class SpmDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("L", "Lp", "Nw", "nrt", "nb", "npt", "nplanes", "nsplit", "mt", "nbal", "batch_wide", "nc", "fold")]


_P = C.c_void_p


class SpmBuffers(C.Structure):
    _fields_ = [
        ("Pf", _P), ("PtPf", _P), ("Cvec", _P), ("Ginv_cache", _P), ("w_cache", _P), ("sigma_cache", _P),
        ("slot", _P), ("mu10", _P), ("mu20", _P), ("mu20_used", _P), ("done", _P), ("iters", _P),
        ("last_res", _P), ("Dre", _P),
        ("b0", _P), ("x0", _P), ("x1", _P), ("h10", _P), ("y0", _P), ("x0_old", _P), ("V", _P), ("aim", _P),
        ("S", _P),
        ("normsA", _P), ("normsB", _P), ("gsum", _P), ("gpart", _P), ("cta_partA", _P), ("cta_partB", _P), ("lazy", _P), ("xready", _P),
        ("iter_counter", _P), ("flags", _P), ("history", _P), ("hist_cap", C.c_int),
        ("lam", C.c_double), ("rtol", C.c_double), ("max_mu", C.c_double),
        ("fact_incr", C.c_double), ("th_change", C.c_double),
        ("bal_bounds", _P), ("bal_first", _P),
    ]


class BpBuffers(C.Structure):
    _fields_ = [
        ("nb", C.c_int), ("M", C.c_int), ("N", C.c_int), ("woodbury", C.c_int), ("nk", C.c_int),
        ("A", _P), ("At", _P), ("aty", _P), ("gram", _P), ("Kinv", _P), ("x0", _P), ("x1", _P), ("h", _P), ("x0_old", _P), ("mu", _P),
        ("need_factor", _P), ("done", _P), ("iters", _P), ("last_res", _P), ("history", _P),
        ("hist_cap", C.c_int),
        ("alpha", C.c_double), ("lam", C.c_double), ("rtol", C.c_double), ("max_mu", C.c_double),
        ("fact_incr", C.c_double), ("th_change", C.c_double), ("interval_update_mu", C.c_int),
    ]


MAX_PEERS = 16
MAILBOX_BYTES = 2 * MAX_PEERS * 32 * 8


class PeerComm(C.Structure):
    """``admm_peer_comm``: rank, world, the mailbox of every rank as mapped into this process, local control words."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("mbox", _P * MAX_PEERS), ("ctrl", _P)]


_LL = C.c_longlong
_I = C.c_int
_D = C.c_double
_SIGS = {
    "admm_abi_version": ([], _I),
    "admm_last_error": ([], C.c_char_p),
    "admm_device_info": ([C.POINTER(_I)] * 4, _I),
    "admm_gemm": ([_I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _P], _I),
    "admm_diag_mul": ([_I, _I, _I, _I, _P, _P, _I, _P, _I, _P], _I),
    "admm_axpby": ([_LL, _D, _P, _D, _P, _P, _P], _I),
    "admm_ewise_unary": ([_I, _I, _LL, _P, _P, _P], _I),
    "admm_prox_l1": ([_LL, _P, _I, _P, _D, _P, _I, _P], _I),
    "admm_prox_nonneg": ([_LL, _P, _I, _P, _P, _I, _P], _I),
    "admm_prox_psd_work_doubles": ([_I, _LL, C.POINTER(_LL)], _I),
    "admm_prox_psd": ([_I, _LL, _LL, _LL, _LL, _P, _I, _P, _P, _I, _P, _P], _I),
    "admm_svd_jacobi": ([_I, _I, _P, _I, _P, _I, _P, _P, _I, _P, _P], _I),
    "admm_sumsq": ([_LL, _P, _P, _P, _P, _P], _I),
    "admm_pair_norms": ([_LL, _P, _P, _LL, _P, _P, _P, _P, _P], _I),
    "admm_inverse": ([_I, _I, _P, _I, _P, _I, _P, _P, _P], _I),
    "admm_spd_inverse_batched": ([_I, _I, _P, _LL, _I, _P, _P, _P], _I),
    "admm_hpd_inverse_batched": ([_I, _I, _P, _LL, _I, _P, _P, _P, _P], _I),
    "admm_spm_prepare_P": ([C.POINTER(SpmDims), _P, _I, _P, _P], _I),
    "admm_spm_pack_operator": ([C.POINTER(SpmDims), _P, _P, _P], _I),
    "admm_spm_pack_L": ([C.POINTER(SpmDims), _P, _I, _P, _P], _I),
    "admm_spm_unpack_L": ([C.POINTER(SpmDims), _P, _P, _I, _P], _I),
    "admm_spm_pack_state": ([C.POINTER(SpmDims), _P, _P, _I, _P, _P, _P, _P], _I),
    "admm_spm_unpack_state": ([C.POINTER(SpmDims), _P, _P, _P, _P, _P, _I, _P], _I),
    "admm_spm_factor": ([C.POINTER(SpmDims), _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "admm_spm_xupdate": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _P], _I),
    "admm_spm_refresh_y": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _P], _I),
    "admm_spm_pass": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _I, _P], _I),
    "admm_spm_step": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _P], _I),
    "admm_spm_step_supported": ([C.POINTER(SpmDims)], _I),
    "admm_spm_reduce": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _P], _I),
    "admm_spm_reduce_decide": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _I, _P], _I),
    "admm_spm_decide": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _I, _P], _I),
    "admm_peer_alloc": ([C.c_size_t, C.POINTER(_P), C.POINTER(C.c_ubyte)], _I),
    "admm_peer_open": ([C.POINTER(C.c_ubyte), C.POINTER(_P)], _I),
    "admm_peer_close": ([_P], _I),
    "admm_peer_free": ([_P], _I),
    "admm_spm_reduce_post": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _P], _I),
    "admm_spm_decide_peer": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _I, _P], _I),
    "admm_spm_step_lazy": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _I, _P], _I),
    "admm_spm_xupdate_lazy": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _I, _P], _I),
    "admm_spm_pass_lazy": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _P], _I),
    "admm_spm_flush": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), C.POINTER(PeerComm), _P], _I),
    "admm_spm_solo_supported": ([C.POINTER(SpmDims)], _I),
    "admm_spm_launch_mode": ([_I], _I),
    "admm_spm_solo": ([C.POINTER(SpmDims), C.POINTER(SpmBuffers), _P, _P, _I, _I, _P], _I),
    "admm_bp_supported": ([_I, _I], _I),
    "admm_bp_setup": ([C.POINTER(BpBuffers), _P, _P, _P, _P], _I),
    "admm_bp_tile_A": ([C.POINTER(BpBuffers), _P, _P], _I),
    "admm_bp_factor": ([C.POINTER(BpBuffers), _P, _P], _I),
    "admm_bp_iterate": ([C.POINTER(BpBuffers), _I, _P], _I),
}
for _name, (_args, _res) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = _res

if lib.admm_abi_version() != ABI_VERSION:
    raise ImportError("libadmm_b200.so ABI version mismatch: rebuild with python __graft_entry__.py")

#: number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0
_LAUNCHES_PER_CALL = {"admm_peer_alloc": 0, "admm_peer_open": 0, "admm_peer_close": 0, "admm_peer_free": 0, "admm_sumsq": 2, "admm_pair_norms": 2, "admm_spm_reduce": 2, "admm_spm_reduce_decide": 2, "admm_bp_factor": 3, "admm_bp_setup": 2, "admm_hpd_inverse_batched": 3}


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.admm_last_error().decode("utf-8", "replace")
        if rc == 3:
            raise NotImplementedError(msg)
        if rc == 1:
            raise ValueError(msg)
        raise AdmmError(msg)


def call(name: str, *args) -> None:
    """Invoke an ABI entry point, raising on a non-zero status."""
    global launch_count
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)
    check(getattr(lib, name)(*args))


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise AdmmError("admmsolver_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_info():
    v = [C.c_int(0) for _ in range(4)]
    check(lib.admm_device_info(*[C.byref(x) for x in v]))
    return tuple(x.value for x in v)
