"""Run-to-run differences of ONE solver while unrelated kernels co-run on another stream (HBM-saturating copies +
matmuls from a background thread).  Whole solves and the phases one by one.  usage: python tools/gpu_flaky2.py [reps] [B]"""
import sys, threading
import numpy as np, torch
sys.path.insert(0, ".")
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
bits = lambda a: np.ascontiguousarray(a).view(np.int64) if a.dtype == np.float64 else a
dev = torch.device("cuda", 0)

stop = False
def load():
    s2 = torch.cuda.Stream(device=dev)
    a = torch.empty(1 << 27, dtype=torch.float64, device=dev); b = torch.empty_like(a)
    m1 = torch.randn(4096, 4096, device=dev, dtype=torch.float32)
    with torch.cuda.stream(s2):
        while not stop:
            for _ in range(4):
                b.copy_(a); m1 = (m1 @ m1).clamp_(-1, 1)
            s2.synchronize()

def ndiff(a, b):
    return {k: int((bits(a[k]) != bits(b[k])).reshape(len(a[k]), -1).any(axis=1).sum()) for k in a if (bits(a[k]) != bits(b[k])).any()}

s = pkg.BatchSolver(spec, B)
ref = s.solve_al_host(X0)
# phases, reference values without load
def phases(s):
    s.set_inputs(X0); s.solve_setup(); s.rollout(); s.update_expansions()
    e = s.expansion(37)
    s.backward_pass()
    K, d = s.gains()
    s.forward_pass()
    X, U = s.trajectory()
    return dict(A=e["A"], lxx=e["lxx"], lx=e["lx"], K=K, d=d, X=X, U=U)
import os
pref = None
if os.environ.get("FLAKY_PHASES") == "1":
    sp = pkg.BatchSolver(spec, B)
    pref = phases(sp)
    again = phases(sp)
    print("phases reproducible without load:", not ndiff(pref, again))
t = threading.Thread(target=load, daemon=True); t.start()
bad = 0
for r in range(reps):
    out = s.solve_al_host(X0)
    d = ndiff({k: ref[k] for k in ("status", "iters", "cost", "X", "U")}, {k: out[k] for k in ("status", "iters", "cost", "X", "U")})
    if d:
        bad += 1
        if os.environ.get("FLAKY_QUIET") != "1": print(f"  solve rep {r}: differing instances per field {d}", flush=True)
print(f"whole solves under load: {bad} of {reps} differ", flush=True)
if pref is not None:
    badp = {}
    for r in range(reps):
        cur = phases(sp)
        d = ndiff(pref, cur)
        for k in d: badp[k] = badp.get(k, 0) + 1
        if d: print(f"  phases rep {r}: {d}", flush=True)
    print("phases under load, runs differing per output:", badp)
stop = True; t.join(timeout=10)
