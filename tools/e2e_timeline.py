"""Timeline of one e2e bench step: per dumped QP, when (ms after the step starts) its uploads, its solve kernel and its
downloads finish on its stream, plus the host time spent enqueueing."""
import sys, os, time, ctypes as C, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import bench, restartsqp_b200 as r
from restartsqp_b200 import capi
L = capi.lib(); B = 4096
fixtures = bench.load_fixtures(); groups = []
for k, q in enumerate(fixtures):
    d = bench.make_batch(q, B, 1234 + k)
    s = r.CudaQPInterface(nV=d["nV"], nC=d["nC"], qptype=r.QPType.QP, batch=B, keep_state=False)
    s.set_csc(capi.MAT_A, q["A_colptr"], q["A_rowidx"], d["Av"]); s.set_csc(capi.MAT_H, q["H_colptr"], q["H_rowidx"], d["Hv"])
    s.set_g(d["g"]); s.set_lb(d["lb"]); s.set_ub(d["ub"])
    if d["nC"]: s.set_lbA(d["lbA"]); s.set_ubA(d["ubA"])
    pin = {kk: torch.from_numpy(v).pin_memory() for kk, v in d.items() if isinstance(v, np.ndarray)}
    nV, nC = d["nV"], d["nC"]
    o = dict(x=torch.empty((B, nV), dtype=torch.float64).pin_memory(), y=torch.empty((B, nV + nC), dtype=torch.float64).pin_memory(),
             obj=torch.empty(B, dtype=torch.float64).pin_memory(), st=torch.empty(B, dtype=torch.int32).pin_memory())
    groups.append(dict(q=q, s=s, pin=pin, out=o, nC=nC, st=torch.cuda.Stream()))
for gr in groups: gr["s"]._solve(r.QPType.QP, None, None, 0)
torch.cuda.synchronize()
cal = [gr["s"].last_solve_ms() for gr in groups]
groups = [groups[i] for i in sorted(range(len(groups)), key=lambda i: -cal[i])]
for gr in groups: gr["s"].set_stream(gr["st"].cuda_stream)
def step(log):
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record(); t0 = time.perf_counter()
    for gr in groups: gr["st"].wait_event(ev0)
    for gr in groups:
        s, p = gr["s"], gr["pin"]
        s.set_csc_values(capi.MAT_A, p["Av"]); s.set_csc_values(capi.MAT_H, p["Hv"]); s.set_g(p["g"]); s.set_lb(p["lb"]); s.set_ub(p["ub"])
        if gr["nC"]: s.set_lbA(p["lbA"]); s.set_ubA(p["ubA"])
        gr["e_up"] = torch.cuda.Event(enable_timing=True); gr["e_up"].record(gr["st"])
        s._solve(r.QPType.QP, None, None, 0)
        gr["e_k"] = torch.cuda.Event(enable_timing=True); gr["e_k"].record(gr["st"])
        gr["t_enq"] = time.perf_counter() - t0
    for gr in groups:
        s, o = gr["s"], gr["out"]
        L.sqpb200_get_solution(s.h, C.c_void_p(o["x"].data_ptr()), C.c_void_p(o["y"].data_ptr()), C.c_void_p(o["obj"].data_ptr()), C.c_void_p(o["st"].data_ptr()), None, capi.LOC_HOST)
        gr["t_got"] = time.perf_counter() - t0
    torch.cuda.synchronize()
    if log:
        for gr in groups:
            print(f"{gr['q']['name']:12s} upload done {ev0.elapsed_time(gr['e_up']):7.2f}  kernel done {ev0.elapsed_time(gr['e_k']):7.2f}  host enqueued {1e3*gr['t_enq']:7.2f}  host got results {1e3*gr['t_got']:7.2f}")
        print("step wall ms", 1e3 * (time.perf_counter() - t0))
for i in range(3): step(i == 2)
