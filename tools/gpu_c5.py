"""C5 (random LQR n=32 m=8 N=100) on the GPU: both large-state kernels, agreement and time per batch.
usage: python tools/gpu_c5.py [B]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec = P.random_lqr_problem()
X0 = P.normal_initial_states(spec, B)
res = {}
for eng in ("fused", "phased"):
    pkg.set_default_engine(eng)
    s = pkg.BatchSolver(spec, B)
    s.set_inputs(X0)
    s.solve_al(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        s.set_inputs(X0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.solve_al(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    r = s.results(); X, U = s.trajectory(); K, d = s.gains()
    res[eng] = dict(status=r["status"], iters=r["iters"], cost=r["cost"], X=X, U=U, K=K, d=d)
    print(f"{eng}: {min(ts)*1e3:.2f} ms per batch of {B} -> {B/min(ts):.0f} solves/s; status {np.unique(r['status'])} iters {np.unique(r['iters'][:,0])}", flush=True)
a, b = res["fused"], res["phased"]
print("same status/iters:", np.array_equal(a["status"], b["status"]), np.array_equal(a["iters"], b["iters"]))
for k in ("X", "U", "cost", "K", "d"):
    print(k, np.abs(a[k] - b[k]).max() / max(1.0, np.abs(a[k]).max()))
