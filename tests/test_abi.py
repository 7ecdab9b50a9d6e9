"""CPU: the C-ABI library loads and exports every symbol include/admm_b200.h declares; the ctypes
mirror covers them all; the product package has no import path into the oracle."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(admm_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(build_lib):
    lib = ctypes.CDLL(os.path.join(ROOT, "admmsolver_b200", "libadmm_b200.so"))
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/admm_b200.h but not exported"
    lib.admm_abi_version.restype = ctypes.c_int
    assert lib.admm_abi_version() == 5


def test_ctypes_mirror_complete(build_lib):
    from admmsolver_b200 import _lib
    assert sorted(_lib._SIGS) == _declared()


def test_struct_layout_matches_header(build_lib):
    """Field order of the ctypes structs == field order of the C structs."""
    from admmsolver_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    for cname, cls in (("admm_spm_dims", _lib.SpmDims), ("admm_spm_buffers", _lib.SpmBuffers),
                       ("admm_bp_buffers", _lib.BpBuffers), ("admm_peer_comm", _lib.PeerComm)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                part = re.sub(r"\[[^\]]*\]", "", part)          # array members: drop the extent
                fields.append(re.findall(r"[A-Za-z_0-9]+", part)[-1])
        assert fields == [f[0] for f in cls._fields_], cname


def test_no_cpu_fallback_and_no_oracle_import(build_lib):
    """The product never imports oracle/ and refuses to compute without a CUDA device."""
    import torch
    pkg = os.path.join(ROOT, "admmsolver_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("oracle problems", "") or fn == "problems.py", fn
            assert "import oracle" not in src and "from oracle" not in src, fn
    if not torch.cuda.is_available():
        import numpy as np
        import pytest
        from admmsolver_b200 import _lib, batch
        with pytest.raises(_lib.AdmmError):
            batch.BatchedBasisPursuit(np.eye(2), np.ones(2))


def test_solo_support_query_without_gpu(build_lib):
    """admm_spm_solo_supported is pure host logic for the per-problem criterion (shared-memory budget of the cluster)
    and degrades to 'unsupported' -- not to a crash -- when the co-residency query of the batch-wide criterion has no device."""
    import ctypes as C
    import torch
    from admmsolver_b200 import _lib

    def dims(L, Lp, Nw, nb, batch_wide):
        nrt = ((Nw + 7) // 8 + 3) // 4 * 4
        return _lib.SpmDims(L, Lp, Nw, nrt, nb, (nb + 7) // 8, 2, 1, 1, 0, int(batch_wide))

    q = lambda d: int(_lib.lib.admm_spm_solo_supported(C.byref(d)))
    assert q(dims(39, 40, 2000, 1, False)) == 8          # cfg2: registers + 49 KB of shared memory per CTA
    assert q(dims(39, 40, 2000, 100, False)) == 8        # per-problem: any batch size (clusters run in waves)
    assert q(dims(52, 64, 2000, 1, True)) == 8           # L > 40: shared-memory variant, one problem
    assert q(dims(52, 64, 40000, 1, False)) == 0         # 5000 sampling points per CTA x 64 columns do not fit
    assert q(dims(39, 48, 2000, 1, False)) == 0          # padded L must be 16, 40 or 64
    if not torch.cuda.is_available():
        assert q(dims(39, 40, 2000, 6, True)) == 0       # batch-wide over several clusters needs the occupancy query


def test_reference_arm_never_maps_the_product_library(build_lib):
    """`bench.py --impl reference` must time the stock reference with none of this repo's native code in
    the process (the input generators are loaded by path)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, ADMM_BENCH_NO_MP="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "spm_cfg2",
                        "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["native_so_mapped"] is False
    assert line["config"]["problems_total"] == 1 and line["cpu_baseline"]["kind"] in ("reference", "port")
