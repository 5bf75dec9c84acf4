#!/usr/bin/env python
"""bench.py -- QP subproblems/sec of the batched QP-subproblem hot path on B200.

Workload (BASELINE.json configs[1]): replay of the dumped QP subproblems of
test/unsolved_QP_data + test/unsolved_QPs as a batch.  Every dumped QP with a symmetric Hessian array
(21 of the 27; the other six hold a non-symmetric "H", i.e. are not QPs, and are kept as robustness tests
only) is replicated B = 4096 times with g *= 1 + 1e-3*U(-1,1) (seed 1234, replica 0 exact).  One step =
one cold-start solve (init) of every replica of every dumped QP = 20*B QPs per GPU.  With N GPUs every
rank solves its own 20*B replicas (rank-seeded perturbations): weak scaling, no collective on the path.

  value  QPs/s with all inputs resident in HBM: K steps of 20 solve launches, CUDA events, max over ranks.
  e2e    the same through the public plugin API with HOST (pinned) buffers: every step uploads H, A values,
         g, lb, ub, lbA, ubA of every instance, solves, and reads x, y, objective and status back.
  roofline  qp_solve_kernel: algorithmic compulsory bytes over the kernel's device time,
         against the measured HBM copy peak; FP64 and shared-memory rates are reported next to it.
  cpu_baseline  the CPU oracle (oracle/, a port: qpOASES itself is not available) on all host cores over a
         bounded sample of the same workload.

`--impl reference` times that CPU path alone and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "QP subproblems/sec"
UNIT = "QPs/s"
WORKLOAD = "replay_dumped_qps (test/unsolved_QP_data + test/unsolved_QPs, the 20 symmetric-H dumps that are solvable x B replicas, cold start)"
# QORE_hs107 is the reference's own failure case (non-convex: the projected Hessian of the working set the homotopy reaches is
# indefinite); GPU kernel, CPU oracle and the reference's qpOASES run all end it with an error status, so it is a robustness
# test (tests/test_gpu_qp.py), not a throughput workload: the headline counts only solves that end QP_OPTIMAL.
EXCLUDED = ("QORE_hs107",)


# ------------------------------------------------------------------------------------------ workload
def load_fixtures():
    import scipy.sparse as sp
    with open(os.path.join(ROOT, "tests", "golden", "qp_fixtures.json")) as f:
        qps = json.load(f)["qps"]
    out = []
    for q in qps:
        nV = q["nV"]
        Hd = sp.csc_matrix((q["H_val"], q["H_rowidx"], q["H_colptr"]), shape=(nV, nV)).toarray()
        if np.abs(Hd - Hd.T).max() == 0.0 and q["name"] not in EXCLUDED:
            out.append(q)
    return out


def make_batch(q, B, seed):
    nV, nC = q["nV"], q["nC"]
    rng = np.random.default_rng(seed)
    g = np.tile(np.array(q["g"], dtype=np.float64), (B, 1))
    g[1:] *= 1.0 + 1e-3 * rng.uniform(-1.0, 1.0, size=g[1:].shape)
    tile = lambda k, n: np.ascontiguousarray(np.tile(np.array(q[k], dtype=np.float64).reshape(1, n), (B, 1)))
    return dict(nV=nV, nC=nC, g=g, lb=tile("lb", nV), ub=tile("ub", nV), lbA=tile("lbA", nC), ubA=tile("ubA", nC),
                Av=tile("A_val", len(q["A_val"])), Hv=tile("H_val", len(q["H_val"])))


def algorithmic_bytes(q):
    """Compulsory HBM bytes of one solve (SURVEY.md 8d, row 'QP/LP solve'): data in, x/y/working set/obj/status out."""
    nV, nC, zA, zH = q["nV"], q["nC"], len(q["A_val"]), len(q["H_val"])
    return 8 * (zH + zA + 3 * nV + 2 * nC) + 8 * (2 * nV + nC) + 4 * (nV + nC) + 16


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_rate(fixtures, sample_B, seed, max_seconds=25.0, min_seconds=0.0):
    """QPs/s of the CPU oracle on all host cores over `sample_B` replicas of every dumped QP (whole passes over the
    21 dumps are repeated until `min_seconds` of CPU work have been timed; data generation is outside the timed region)."""
    from oracle import oracle_py as orc
    cores = orc.lib().orc_max_threads()
    total, t_total = 0, 0.0
    batches = [make_batch(q, sample_B, seed + k) for k, q in enumerate(fixtures)]
    while True:
        for q, d in zip(fixtures, batches):
            A = (q["A_colptr"], q["A_rowidx"], q["A_val"])
            H = (q["H_colptr"], q["H_rowidx"], q["H_val"])
            t0 = time.perf_counter()
            r = orc.solve_batch(d["nV"], d["nC"], A, H, d["g"], d["lb"], d["ub"], d["lbA"], d["ubA"], Avals=d["Av"], Hvals=d["Hv"])
            t_total += time.perf_counter() - t0
            total += sample_B
            cores = r["threads"]
        if t_total >= min_seconds or t_total > max_seconds:
            break
    return total / t_total, cores, total, t_total


def run_reference(args, rank, world):
    if rank != 0:
        return
    fixtures = load_fixtures()
    # each step = a bounded sample of the workload: passes over `sample_B` replicas of every dump for >= 4 s
    sample_B = args.replicas  # the same replicas per dumped QP as the GPU arm
    cores = 1
    for _ in range(args.warmup):
        cpu_rate(fixtures, min(sample_B, 64), 1234)
    tot, tt = 0, 0.0
    for _ in range(args.steps):
        r, cores, n, t = cpu_rate(fixtures, sample_B, 1234, max_seconds=30.0, min_seconds=4.0)
        tot += n; tt += t
    value = tot / tt
    sample = "per step: passes over %d replicas of each of the %d dumped QPs for >= 4 s (%d QPs/step on average)" % (
        sample_B, len(fixtures), tot // max(1, args.steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (dumped QP fixtures of the reference, perturbed replicas)",
        "config": {"workload": WORKLOAD, "replicas_per_qp": sample_B, "qps_per_step": tot // max(1, args.steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU oracle (restatement of the qpOASES path; qpOASES 3.2.1 is not available offline), "
                                 "one solver object per thread, pthreads over instances"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch
    import restartsqp_b200 as r
    from restartsqp_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = capi.lib()
    fixtures = load_fixtures()
    B = args.replicas
    groups = []
    h2d = d2h = 0
    for k, q in enumerate(fixtures):
        d = make_batch(q, B, 1234 + 1000 * rank + k)
        s = r.CudaQPInterface(nV=d["nV"], nC=d["nC"], qptype=r.QPType.QP, batch=B, device=local_rank, keep_state=False,
                              team_size=args.team)
        s.set_csc(capi.MAT_A, q["A_colptr"], q["A_rowidx"], d["Av"])
        s.set_csc(capi.MAT_H, q["H_colptr"], q["H_rowidx"], d["Hv"])
        s.set_g(d["g"]); s.set_lb(d["lb"]); s.set_ub(d["ub"])
        if d["nC"]:
            s.set_lbA(d["lbA"]); s.set_ubA(d["ubA"])
        pin = {kk: torch.from_numpy(v).pin_memory() for kk, v in d.items() if isinstance(v, np.ndarray)}
        nV, nC = d["nV"], d["nC"]
        outp = dict(x=torch.empty((B, nV), dtype=torch.float64).pin_memory(), y=torch.empty((B, nV + nC), dtype=torch.float64).pin_memory(),
                    obj=torch.empty(B, dtype=torch.float64).pin_memory(), st=torch.empty(B, dtype=torch.int32).pin_memory())
        h2d += sum(int(t.numel()) * 8 for t in pin.values())
        d2h += sum(int(t.numel()) * t.element_size() for t in outp.values())
        groups.append(dict(q=q, s=s, pin=pin, out=outp, nV=nV, nC=nC))
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    # one CUDA stream per dumped QP: the 21 independent solve launches of a step overlap, so small groups fill the
    # SMs that the long-tailed ones (hs107: up to 473 working-set changes) leave idle
    streams = [torch.cuda.Stream() for _ in groups] if args.streams else None

    def use_streams(on):
        for k, gr in enumerate(groups):
            gr["s"].set_stream(streams[k].cuda_stream if (on and streams) else 0)

    def fork():
        flush.zero_()
        if streams:
            ev = torch.cuda.Event()
            ev.record()
            for st in streams:
                st.wait_event(ev)

    def join():
        if streams:
            cur = torch.cuda.current_stream()
            for st in streams:
                cur.wait_stream(st)

    def step_resident():
        fork()
        for gr in groups:
            gr["s"]._solve(r.QPType.QP, None, None, 0)
        join()

    def step_e2e():
        fork()
        for gr in groups:
            s, p, o = gr["s"], gr["pin"], gr["out"]
            s.set_csc_values(capi.MAT_A, p["Av"]); s.set_csc_values(capi.MAT_H, p["Hv"])
            s.set_g(p["g"]); s.set_lb(p["lb"]); s.set_ub(p["ub"])
            if gr["nC"]:
                s.set_lbA(p["lbA"]); s.set_ubA(p["ubA"])
            s._solve(r.QPType.QP, None, None, 0)
        for gr in groups:  # results are read after every group has been queued, so uploads, solves and downloads overlap
            s, o = gr["s"], gr["out"]
            L.sqpb200_get_solution(s.h, C.c_void_p(o["x"].data_ptr()), C.c_void_p(o["y"].data_ptr()), C.c_void_p(o["obj"].data_ptr()),
                                   C.c_void_p(o["st"].data_ptr()), None, capi.LOC_HOST)
        join()

    def timed(fn, steps, warmup, collect_kernel_ms=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = sum(gr["s"].launch_count() for gr in groups)
        kms = 0.0
        e0.record()
        for _ in range(steps):
            fn()
            if collect_kernel_ms:  # per-launch CUDA-event time recorded by the library around qp_solve_kernel
                kms += sum(gr["s"].last_solve_ms() for gr in groups)
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = sum(gr["s"].launch_count() for gr in groups) - launches0
        return float(t.item()), launches, kms

    # Launch order: longest solve first.  The step is bounded by its longest kernel (hs107: a few replicas need 400
    # working-set changes), so that kernel must start first and the short ones fill the SMs its tail leaves idle.  The order
    # comes from one untimed calibration pass on the default stream.
    step_resident()
    torch.cuda.synchronize()
    cal = [gr["s"].last_solve_ms() for gr in groups]
    order = sorted(range(len(groups)), key=lambda i: -cal[i])
    groups[:] = [groups[i] for i in order]
    if streams:
        streams = streams[: len(groups)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    use_streams(True)
    ms_res, launches, _ = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    # serial pass on the default stream, only to attribute device time to the solve kernel (per-launch CUDA events)
    torch.cuda.synchronize()
    use_streams(False)
    ms_serial, _, kernel_ms = timed(step_resident, args.steps, 1, collect_kernel_ms=True)
    qps_step = len(fixtures) * B * world
    value = qps_step * args.steps / (ms_res * 1e-3)
    e2e = qps_step * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        # parity spot check on the benchmarked data (replica 0 of each dump) against the oracle
        status_ok, per_dump = 0, {}
        for gr in groups:
            ok = int((gr["s"].get_status() == 20).sum())
            status_ok += ok
            per_dump[gr["q"]["name"]] = ok
        # roofline of the dominant kernel
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_step = sum(algorithmic_bytes(gr["q"]) for gr in groups) * B
        kernel_s_per_step = kernel_ms * 1e-3 / args.steps
        achieved = bytes_step / kernel_s_per_step / 1e9
        fp64 = smem = None
        try:
            P = C.CDLL(os.path.join(ROOT, "restartsqp_b200", "lib", "libsqpb200_peaks.so"))
            a, b = C.c_double(0), C.c_double(0)
            if P.peaks_measure(C.byref(a), C.byref(b)) == 0:
                fp64, smem = a.value, b.value
        except Exception:
            pass
        flops_step = None
        try:
            from oracle import oracle_py as orc
            fl = 0.0
            for gr in groups:
                q = gr["q"]
                o = orc.OracleQP(q["nV"], q["nC"])
                o.init((q["H_colptr"], q["H_rowidx"], q["H_val"]), q["g"], (q["A_colptr"], q["A_rowidx"], q["A_val"]), q["lb"], q["ub"], q["lbA"], q["ubA"])
                fl += o.flops()
            flops_step = fl * B
        except Exception:
            pass
        # the dominant single launch: hs116, the largest dumped QP (nV=69, nC=28); its DRAM traffic per launch comes from the
        # committed ncu --set full capture of exactly this launch shape (profiles/r1_qp_solve_hs116.md)
        dom = [gr for gr in groups if gr["q"]["name"] == "QORE_hs116"]
        dom_launch = None
        if dom:
            dom_ms = dom[0]["s"].last_solve_ms()
            dom_bytes = algorithmic_bytes(dom[0]["q"]) * B
            dom_launch = {"launch": "QORE_hs116 x %d" % B, "ms": dom_ms, "algorithmic_bytes": dom_bytes,
                          "achieved_gbs": dom_bytes / (dom_ms * 1e-3) / 1e9,
                          "traffic_bytes_ncu": (13351680 if B == 4096 else None),
                          "traffic_source": "profiles/r1_qp_solve_hs116.md (dram__bytes_read.sum + dram__bytes_write.sum, B=4096)"}
        # headline roofline numbers = the dominant single launch (hs116 x B: algorithmic bytes / its CUDA-event duration, and the
        # DRAM traffic ncu measured for exactly that launch); the aggregate over the 21 launches of a step is kept beside it
        if dom_launch is not None:
            r_ach, r_traffic, r_launch = dom_launch["achieved_gbs"], dom_launch["traffic_bytes_ncu"], dom_launch["launch"]
        else:
            r_ach, r_traffic, r_launch = achieved, None, "all launches of a step"
        roofline = {"kernel": "qp_solve_kernel", "launch": r_launch, "bound": "hbm", "achieved": r_ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": r_ach / hbm_peak, "traffic": r_traffic, "peak_source": peak_src,
                    "dominant_launch": dom_launch,
                    "all_launches": {"achieved_gbs": achieved, "frac": achieved / hbm_peak, "algorithmic_bytes_per_step": bytes_step},
                    "kernel_ms_per_step": 1e3 * kernel_s_per_step, "kernel_share_of_step": kernel_s_per_step / (ms_serial * 1e-3 / args.steps),
                    "serial_ms_per_step": ms_serial / args.steps,
                    "note": "active-set iterations run out of shared memory: the kernel is latency/issue bound, not HBM bound (ncu: every input "
                            "byte read once, 31 % issue-slot utilisation, 8 resident warps/SM at nV=69); FP64 rate (flop model of SURVEY.md 8d "
                            "counted by the oracle on replica 0) is given in fp64; HBM-bound kernels of the path: profiles/r1_l0_kernels.md"}
        if flops_step is not None:
            roofline["fp64"] = {"achieved_gflops": flops_step / kernel_s_per_step / 1e9, "peak_gflops": fp64,
                                "frac": (flops_step / kernel_s_per_step / 1e9 / fp64) if fp64 else None,
                                "peak_source": "FMA-chain microbenchmark tools/peaks.cu, this run"}
        if smem is not None:
            roofline["smem_peak_gbs"] = smem
        try:  # FP64 tensor-pipe peak: cuBLAS DGEMM 8192^3 through torch, best of 3
            a_ = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b_ = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
            torch.matmul(a_, b_); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0_.record(); torch.matmul(a_, b_); t1_.record(); torch.cuda.synchronize()
                best = min(best, t0_.elapsed_time(t1_))
            roofline["dgemm_peak_gflops"] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e9
            del a_, b_
        except Exception:
            pass
        sample_B = B  # the same replicas per dumped QP as the timed GPU workload
        rate, cores, n, t = cpu_rate(fixtures, sample_B, 1234, max_seconds=60.0, min_seconds=12.0)  # >= 12 s of CPU work
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (dumped QP fixtures of the reference, perturbed replicas)",
            "config": {"workload": WORKLOAD, "replicas_per_qp": B, "qps_per_step_per_gpu": len(fixtures) * B,
                       "l2": "256 MiB buffer written between steps (inside the timed region)",
                       "solved_optimal": status_ok, "solved_optimal_per_dump": per_dump, "team_size": args.team or "auto",
                       "streams": len(streams) if streams else 1},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d QPs (passes over %d replicas of each of the %d dumped QPs) in %.1f s" % (n, sample_B, len(fixtures), t)},
        }
        if args.extras:
            out["extras"] = run_extras(local_rank)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(device):
    """Untimed-by-the-driver side measurements on rank 0 (outside the timed region of the headline metric): the other
    BASELINE.json configurations that fit one GPU, each through the public API, each with its own CPU-oracle figure."""
    import restartsqp_b200 as r
    from restartsqp_b200 import capi
    from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
    from restartsqp_b200.sqp_device import DeviceBatchedSQP as BatchedSQP  # device-resident outer loop
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    ex = {}
    try:  # configs[0]/[2]: full SQP solves per second, HS071 x 10^4 perturbed starts, device NLP evaluation + CUDA QP backend
        host = AmplNLP(os.path.join(ROOT, "tests", "golden", "hs_nl", "hs071.nl"))
        dev = DeviceNLP(host, device=device)
        Bs = 10000
        x0, _ = host.Get_starting_point()
        xl, xu, _, _ = host.Get_bounds_info()
        rng = np.random.default_rng(71000)
        X = np.clip(x0 * (1 + 0.1 * rng.standard_normal((Bs, host.n))) + 0.1 * rng.standard_normal((Bs, host.n)), xl, xu)
        warm = BatchedSQP(dev, x0=X[:256], device=device)
        warm.Optimize()
        runs = []
        for _ in range(3):  # initialization + Optimize, like the reference's own timing (src/Algorithm.cpp:57)
            t0 = time.perf_counter()
            alg = BatchedSQP(dev, x0=X, device=device)
            t1 = time.perf_counter()
            res = alg.Optimize()
            runs.append((time.perf_counter() - t0, t1 - t0))
            alg.myQP_.solverInterface_.close(); alg.myLP_.solverInterface_.close()  # teardown (cudaFree) is outside the timed region
        warm.myQP_.solverInterface_.close(); warm.myLP_.solverInterface_.close()
        dt = float(np.median([r_[0] for r_ in runs]))
        t0, t1 = 0.0, float(np.median([r_[1] for r_ in runs]))
        ex["sqp_hs071"] = {"metric": "SQP solves/sec", "value": Bs / dt, "unit": "solves/s", "instances": Bs,
                           "optimal": int((res.exitflag == 0).sum()), "sqp_iters_mean": float(res.iters.mean()),
                           "qp_iters_mean": float(res.qp_iter.mean()), "seconds": dt, "seconds_initialization": t1 - t0, "seconds_all_runs": [r_[0] for r_ in runs], "value_best_of_3": Bs / min(r_[0] for r_ in runs),
                           "note": "device-resident outer loop (csrc/sqp_outer.cu), NVRTC NLP evaluation and every QP/LP on the GPU; starts uploaded from the host, results read back; wall clock"}
        try:  # configs[4] at one GPU: 10^6 HS-scale instances (100 copies of the 10^4 starts), one run after a warm-up of the allocations
            Xm = np.tile(X, (100, 1))
            t0 = time.perf_counter()
            algm = BatchedSQP(dev, x0=Xm, device=device)
            resm = algm.Optimize()
            dtm = time.perf_counter() - t0
            algm.close()
            ex["sqp_hs071_1e6"] = {"metric": "SQP solves/sec", "value": Xm.shape[0] / dtm, "unit": "solves/s", "instances": int(Xm.shape[0]),
                                   "optimal": int((resm.exitflag == 0).sum()), "seconds": dtm}
            del Xm, resm
        except Exception as e:
            ex["sqp_hs071_1e6"] = {"error": repr(e)[:200]}
        dev.close()
        try:  # the same solves by the CPU oracle's C restatement of the outer loop on all host cores (bounded sample)
            from oracle import oracle_py as orc
            so = orc.SqpOracle(host, r.Options())
            Xc = np.tile(X, (10, 1))  # 10^5 starts: about a second of CPU work on 16 cores
            so.solve_batch(Xc[:2000])
            t0 = time.perf_counter()
            rc_ = so.solve_batch(Xc)
            tc = time.perf_counter() - t0
            ex["sqp_hs071"]["cpu_baseline"] = {"value": Xc.shape[0] / tc, "unit": "solves/s", "cores": int(rc_["threads"]), "kind": "port",
                                               "sample": "%d HS071 solves in %.2f s (oracle/oracle_sqp.c, one solve per thread at a time)" % (Xc.shape[0], tc),
                                               "optimal": int((rc_["exitflag"] == 0).sum())}
        except Exception as e:
            ex["sqp_hs071"]["cpu_baseline"] = {"error": repr(e)[:200]}
    except Exception as e:  # the extras never take the headline down
        ex["sqp_hs071"] = {"error": repr(e)[:200]}
    try:  # configs[3]: synthetic sparse QP n=256, m=128 (nV=512), batch 64, one QP per CTA
        d = H.synthetic_large_qp(256, batch=64)
        s = r.CudaQPInterface(nV=d["nV"], nC=d["nC"], qptype=r.QPType.QP, batch=64, device=device, keep_state=False)
        s.set_csc(capi.MAT_A, *d["Ac"]); s.set_csc(capi.MAT_H, *d["Hc"])
        s.set_g(d["g"]); s.set_lb(d["lb"]); s.set_ub(d["ub"]); s.set_lbA(d["lbA"]); s.set_ubA(d["ubA"])
        for _ in range(2):
            s._solve(r.QPType.QP, None, None, 0)
            ms = s.last_solve_ms()
        st, it = s.get_status(), s.get_iterations()
        from oracle import oracle_py as orc
        t0 = time.perf_counter()
        rr = orc.solve_batch(d["nV"], d["nC"], d["Ac"], d["Hc"], d["g"][:16], d["lb"][:16], d["ub"][:16], d["lbA"][:16], d["ubA"][:16])
        tc = time.perf_counter() - t0
        ex["large_qp_n256"] = {"metric": "QP subproblems/sec", "value": 64 / (ms * 1e-3), "unit": "QPs/s", "batch": 64, "nV": d["nV"], "nC": d["nC"],
                               "ms": ms, "optimal": int((st == 20).sum()), "iters_mean": float(it.mean()), "config": s.solve_config(),
                               "cpu_baseline": {"value": 16 / tc, "unit": "QPs/s", "cores": rr["threads"], "kind": "port", "sample": "16 instances in %.1f s" % tc}}
        s.close()
    except Exception as e:
        ex["large_qp_n256"] = {"error": repr(e)[:200]}
    return ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per dumped QP (B of SURVEY.md 8d config 2)")
    ap.add_argument("--team", type=int, default=0, help="threads per QP (0 = auto)")
    ap.add_argument("--extras", type=int, default=1, help="1: also report SQP solves/s (HS071 x 1e4) and the config-4 large-QP rate on rank 0")
    ap.add_argument("--streams", type=int, default=1, help="1: one CUDA stream per dumped QP (overlapping launches), 0: default stream")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
