#!/usr/bin/env python
"""Coverage of the `.nl` reader (restartsqp_b200/nl_reader.py) over the whole test/CUTE_examples directory of the reference (735 files;
the reference's own scripts run the 124 hs* files, test/runhs.sh).  Dev container only: reads /root/reference.

    python tools/cute_coverage.py [--jobs 8] [--timeout 30] [--max-bytes 400000] > profiles/r2_cute_nl_coverage.md

Per file: parse, symbolic first and second derivatives, one evaluation of f, c, gradient, Jacobian and Lagrangian Hessian at the
starting point, a finite-difference check of the gradient on the small ones."""
import argparse
import glob
import json
import multiprocessing as mp
import os
import signal
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/test/CUTE_examples"


class _Timeout(Exception):
    pass


def _alarm(sig, frm):
    raise _Timeout()


def one(arg):
    path, timeout, max_bytes = arg
    import numpy as np
    from restartsqp_b200.nl_reader import AmplNLP
    name = os.path.basename(path)[:-3]
    size = os.path.getsize(path)
    if size > max_bytes:
        return name, dict(status="skipped", why="file of %d bytes" % size)
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(timeout)
    t0 = time.time()
    try:
        h = AmplNLP(path)
        x, lam = h.Get_starting_point()
        xl, xu, _, _ = h.Get_bounds_info()
        X = np.clip(np.atleast_2d(np.asarray(x, dtype=np.float64)), xl, xu)
        L = np.ones((1, h.m))
        f, c, g = h.Eval_f(X), h.Eval_constraints(X), h.Eval_gradient(X)
        J, Hh = h.Eval_Jacobian(X), h.Eval_Hessian(X, L)
        out = dict(status="ok", n=h.n, m=h.m, zJ=len(h.J_row1), zH=len(h.H_row1), nodes=len(h.model.G.nodes), seconds=round(time.time() - t0, 2),
                   finite=bool(np.isfinite(f).all() and np.isfinite(c).all() and np.isfinite(g).all() and np.isfinite(J).all() and np.isfinite(Hh).all()))
        if h.n <= 200 and out["finite"]:
            e = 1e-6
            fd = np.zeros(h.n)
            for i in range(h.n):
                d = np.zeros((1, h.n)); d[0, i] = e
                fd[i] = (h.Eval_f(X + d)[0] - h.Eval_f(X - d)[0]) / (2 * e)
            out["grad_fd_err"] = float(np.abs(fd - g[0]).max() / max(1.0, np.abs(g[0]).max()))
        return name, out
    except _Timeout:
        return name, dict(status="timeout", why="more than %d s" % timeout)
    except Exception as ex:  # noqa: BLE001
        return name, dict(status="error", why="%s: %s" % (type(ex).__name__, str(ex)[:90]))
    finally:
        signal.alarm(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--timeout", type=int, default=30)
    ap.add_argument("--max-bytes", type=int, default=400000)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    files = sorted(glob.glob(os.path.join(REF, "*.nl")))
    with mp.Pool(a.jobs, maxtasksperchild=20) as pool:
        res = {}
        for k, (name, r) in enumerate(pool.imap_unordered(one, [(p, a.timeout, a.max_bytes) for p in files])):
            res[name] = r
            if k % 50 == 0:
                print("... %d / %d" % (k, len(files)), file=sys.stderr, flush=True)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=0, sort_keys=True)
    is_hs = lambda n: n.startswith("hs") and n[2:5].isdigit()
    count = lambda pred: sum(1 for n, r in res.items() if pred(n, r))
    print("# r2 — `.nl` reader over the reference's whole `test/CUTE_examples` directory (%d files)\n" % len(files))
    print("`python tools/cute_coverage.py --jobs %d --timeout %d --max-bytes %d` in the dev container (pure Python reader: parse, symbolic\n"
          "first and second derivatives on the hash-consed DAG, one evaluation of f, c, gradient, Jacobian, Lagrangian Hessian at the\n"
          "starting point; central-difference check of the gradient for n <= 200).\n" % (a.jobs, a.timeout, a.max_bytes))
    print("| set | files | read and evaluated | finite at the start | timeout | skipped (large file) | error |")
    print("|---|---|---|---|---|---|---|")
    for label, pred in (("hs* (the reference's runhs.sh set)", is_hs), ("other CUTE files", lambda n: not is_hs(n))):
        c = lambda st: count(lambda n, r: pred(n) and r["status"] == st)
        print("| %s | %d | %d | %d | %d | %d | %d |" % (label, count(lambda n, r: pred(n)), c("ok"),
                                                       count(lambda n, r: pred(n) and r["status"] == "ok" and r["finite"]), c("timeout"), c("skipped"), c("error")))
    bad = sorted((n, r) for n, r in res.items() if r["status"] == "ok" and r.get("grad_fd_err", 0.0) > 1e-4)
    print("\nGradient against central differences (n <= 200, finite start): %d files checked, %d above 1e-4 relative%s.\n" % (
        count(lambda n, r: "grad_fd_err" in r), len(bad), (": " + ", ".join("%s %.1e" % (n, r["grad_fd_err"]) for n, r in bad[:12])) if bad else ""))
    errs = {}
    for n, r in res.items():
        if r["status"] == "error":
            errs.setdefault(r["why"].split(":")[0] + ":" + r["why"].split(":", 1)[1][:50], []).append(n)
    if errs:
        print("Errors by kind:\n")
        for k, v in sorted(errs.items(), key=lambda kv: -len(kv[1])):
            print("* %d x `%s` (%s%s)" % (len(v), k, ", ".join(sorted(v)[:8]), " ..." if len(v) > 8 else ""))
    small = sorted((r["n"] + 2 * r["m"], n) for n, r in res.items() if r["status"] == "ok" and r["finite"] and not is_hs(n) and r["n"] + 2 * r["m"] <= 100)
    print("\nNon-HS files whose QP subproblem fits the warp kernel (nV = n + 2m <= 100): %d, e.g. %s.\n" % (
        len(small), ", ".join("%s (nV %d)" % (n, v) for v, n in small[:25])))
    sizes = sorted((r["n"], n) for n, r in res.items() if r["status"] == "ok")
    print("Largest files read: %s." % ", ".join("%s (n %d, m %d, %d DAG nodes, %.1f s)" % (n, res[n]["n"], res[n]["m"], res[n]["nodes"], res[n]["seconds"])
                                                 for _, n in sizes[-5:]))


if __name__ == "__main__":
    main()
