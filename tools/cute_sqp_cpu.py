#!/usr/bin/env python
"""The CPU oracle of Algorithm::Optimize (oracle/oracle_sqp.c) over the small non-HS models of the reference's test/CUTE_examples
directory: 16 perturbed starts per model (SURVEY.md 8d config 3 recipe), exit-flag statistics.  Dev container only.

    python tools/cute_coverage.py --json /tmp/cute_cov.json > /dev/null
    python tools/cute_sqp_cpu.py /tmp/cute_cov.json >> profiles/r2_cute_nl_coverage.md
"""
import json
import os
import signal
import sys
import time
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/test/CUTE_examples"


class _Timeout(Exception):
    pass


def _alarm(sig, frm):
    raise _Timeout()


def main():
    import numpy as np  # noqa: F401
    import restartsqp_b200 as r
    from restartsqp_b200.nl_reader import AmplNLP
    from oracle import oracle_py as orc
    from test_hs_suite import perturbed_starts
    cov = json.load(open(sys.argv[1]))
    is_hs = lambda n: n.startswith("hs") and n[2:5].isdigit()
    cands = sorted((v["n"] + 2 * v["m"], n) for n, v in cov.items()
                   if v["status"] == "ok" and v["finite"] and not is_hs(n) and v["n"] + 2 * v["m"] <= 60 and v["nodes"] < 3000)
    signal.signal(signal.SIGALRM, _alarm)
    out = {}
    for _, name in cands:
        try:
            signal.alarm(40)
            a = AmplNLP(os.path.join(REF, name + ".nl"))
            res = orc.SqpOracle(a, r.Options(iter_max=200)).solve_batch(perturbed_starts(a, 16, 3))
            out[name] = dict(n=a.n, m=a.m, optimal=int((res["exitflag"] == 0).sum()), flags=res["exitflag"].tolist(), iters=float(res["iters"].mean()))
        except _Timeout:
            out[name] = dict(error="timeout")
        except Exception as e:  # noqa: BLE001
            out[name] = dict(error=repr(e)[:80])
        finally:
            signal.alarm(0)
    ok = {k: v for k, v in out.items() if "error" not in v}
    flags = Counter(f for v in ok.values() for f in v["flags"])
    names = {0: "OPTIMAL", 2: "EXCEED_MAX_ITER", 4: "TRUST_REGION_TOO_SMALL", 7: "QP_UNCHANGED", 21: "QPERROR_INTERNAL_ERROR", 28: "QPERROR_PERFORMINGHOMOTOPY"}
    print("\n## Full SQP on the small non-HS models (CPU oracle `oracle/oracle_sqp.c`, 16 perturbed starts each, iter_max 200)\n")
    print("`python tools/cute_sqp_cpu.py`: the %d models with nV = n + 2m <= 60 (and a DAG below 3000 nodes): %d ran, %d ended OPTIMAL from all 16 starts,"
          % (len(cands), len(ok), sum(1 for v in ok.values() if v["optimal"] == 16)))
    print("%d of %d solves OPTIMAL.  Exit flags over all solves: %s.\n" % (
        sum(v["optimal"] for v in ok.values()), 16 * len(ok), ", ".join("%s %d" % (names.get(k, str(k)), c) for k, c in flags.most_common())))
    worst = sorted((v["optimal"], k) for k, v in ok.items() if v["optimal"] < 8)
    print("Models with fewer than 8 of 16 OPTIMAL: %s.\n" % ", ".join("%s (%d)" % (k, c) for c, k in worst))
    print("25 of these models (constrained, tabulated optima; 24 polynomial, one with `if`) are test fixtures: `tests/golden/cute_nl`, `tests/test_cute_suite.py`.")


if __name__ == "__main__":
    main()
