"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/sqpb200.h declares; without a GPU every compute entry point refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import restartsqp_b200 as r
from restartsqp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sqpb200.h")).read()
    declared = set(re.findall(r"\b(sqpb200_[A-Za-z0-9_]+)\s*\(", hdr))
    declared.discard("sqpb200_handle_s")
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libsqpb200.so does not export %s" % name
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "sqpb200.h")).read()
    for ref in ("src/qpOASESInterface.cpp:137-284", "src/SpHbMat.cpp:196-268", "include/sqphot/Types.hpp:84-89",
                "src/QPhandler.cpp:167-261"):
        assert ref in hdr


def test_no_cpu_fallback_without_gpu():
    L = capi.lib()
    if L.sqpb200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.sqpb200_create(4, 8, 2, capi.QP, 0, None, C.byref(h))
    assert rc == -2  # SQPB200_ERR_CUDA
    with pytest.raises(capi.SqpB200Error):
        r.CudaQPInterface(nV=8, nC=2, batch=4)


def test_product_does_not_import_oracle():
    """The product package must never import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "restartsqp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "import oracle" not in txt, f
                assert not re.search(r"#include\s*[\"<][^\">]*oracle", txt), f


def test_types_mirror_reference_values():
    assert int(r.ActiveType.ACTIVE_BOTH_SIDE) == -99 and int(r.ActiveType.ACTIVE_ABOVE) == 1
    assert int(r.Exitflag.QP_OPTIMAL) == 20 and int(r.Exitflag.QPERROR_INFEASIBLE) == 22
    assert int(r.QPType.LP) == 1 and int(r.QPType.QP) == 2
    o = r.Options()
    assert (o.qp_maxiter, o.lp_maxiter, o.penalty_iter_max, o.rho_max, o.delta_max) == (1000, 100, 200, 1.0e6, 1.0e8)
