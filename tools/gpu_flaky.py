"""Stress for run-to-run differences: sharded (3 solvers concurrently on one GPU) vs single, and fresh sequential
solvers, many repetitions; prints which instances / fields differ.  usage: python tools/gpu_flaky.py [reps] [B]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
bits = lambda a: np.ascontiguousarray(a).view(np.int64) if a.dtype == np.float64 else a
KEYS = ("status", "iters", "cost", "viol", "X", "U")

def diff(a, b, tag):
    bad = {}
    for k in KEYS:
        d = (bits(a[k]) != bits(b[k])).reshape(B, -1).any(axis=1)
        if d.any(): bad[k] = np.where(d)[0].tolist()
    if bad:
        inst = sorted(set(sum(bad.values(), [])))
        print(f"  {tag}: DIFF fields {list(bad)} instances {inst[:12]} (n={len(inst)}) status {a['status'][inst[:6]].tolist()} "
              f"iters {a['iters'][inst[:3]].tolist()} vs {b['iters'][inst[:3]].tolist()} dcost {(a['cost'][inst[:3]] - b['cost'][inst[:3]]).tolist()}", flush=True)
    return bool(bad)

ref = pkg.BatchSolver(spec, B).solve_al_host(X0)
nbad = {"single_fresh": 0, "single_reuse": 0, "sharded_fresh": 0, "sharded_reuse": 0}
reuse_one = pkg.BatchSolver(spec, B)
reuse_sh = pkg.MultiBatchSolver(spec, B, devices=[0, 0, 0])
import os
quiet = os.environ.get("FLAKY_QUIET") == "1"
if quiet:
    _diff = diff
    def diff(a, b, tag):
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            return _diff(a, b, tag)
for r in range(reps):
    if not quiet:
        nbad["single_fresh"] += diff(ref, pkg.BatchSolver(spec, B).solve_al_host(X0), f"rep {r} single fresh")
        nbad["single_reuse"] += diff(ref, reuse_one.solve_al_host(X0), f"rep {r} single reuse")
    nbad["sharded_fresh"] += diff(ref, pkg.MultiBatchSolver(spec, B, devices=[0, 0, 0]).solve_al_host(X0), f"rep {r} sharded fresh")
    nbad["sharded_reuse"] += diff(ref, reuse_sh.solve_al_host(X0), f"rep {r} sharded reuse")
print("differing runs out of", reps, ":", nbad)
