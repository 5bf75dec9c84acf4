"""QP dump formats (restartsqp_b200/qp_dump.py): the reference's own dump files (read from /root/reference when it
is present) parse to exactly the committed fixtures, and write -> read round trips of every fixture in both formats."""
import os

import numpy as np
import pytest

from restartsqp_b200 import qp_dump
import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = H.load_qp_fixtures()
KEYS = ("nV", "nC", "lb", "ub", "lbA", "ubA", "g", "A_colptr", "A_rowidx", "A_val", "H_colptr", "H_rowidx", "H_val")


def same(a, b):
    return all(np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)) for k in KEYS)


REF = "/root/reference/test"
REF_DUMPS = [(os.path.join(REF, "unsolved_QP_data", f), f.replace("qpdata.log", "")) for f in sorted(os.listdir(os.path.join(REF, "unsolved_QP_data")))
             if f.endswith(".log")] + \
            [(os.path.join(REF, "unsolved_QPs", f), f.replace(".hpp", "") + "_hpp") for f in sorted(os.listdir(os.path.join(REF, "unsolved_QPs")))
             if f.endswith(".hpp")] if os.path.isdir(REF) else []


@pytest.mark.skipif(not REF_DUMPS, reason="the reference checkout is only present in the build container")
@pytest.mark.parametrize("path,fixture", REF_DUMPS, ids=[f for _, f in REF_DUMPS])
def test_reference_dump_files_read_to_the_fixtures(path, fixture):
    """Every dump file of the reference (read where it lies, never copied) parses to exactly the committed fixture."""
    q = qp_dump.read_dump(path)
    ref = [f for f in FIX if f["name"] == fixture]
    assert ref, fixture
    assert q["name"] == fixture and same(q, ref[0])


@pytest.mark.parametrize("q", FIX, ids=[q["name"] for q in FIX])
def test_write_read_round_trip(tmp_path, q):
    p1, p2 = str(tmp_path / "q.log"), str(tmp_path / "q.hpp")
    qp_dump.write_qore_log(p1, q)
    qp_dump.write_qpoases_hpp(p2, q)
    assert same(qp_dump.read_qpoases_hpp(p2), q)
    # the .log reader goes through the dense matrix like the replay driver (entries <= 1e-16 dropped): compare densely
    import scipy.sparse as sp
    r1 = qp_dump.read_qore_log(p1)
    for k in ("lb", "ub", "lbA", "ubA", "g"):
        assert np.array_equal(np.asarray(r1[k]), np.asarray(q[k], dtype=np.float64))
    for M, shape in (("A", (q["nC"], q["nV"])), ("H", (q["nV"], q["nV"]))):
        d = lambda t: sp.csc_matrix((t[M + "_val"], t[M + "_rowidx"], t[M + "_colptr"]), shape=shape).toarray()
        assert np.array_equal(d(r1), d(q))
