"""A/B timing of build variants and tile widths (run under gpurun). Each config in its own process."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, json
sys.path.insert(0, %r)
import numpy as np, torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
B = int(os.environ.get("BENCH_B", "16384"))
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
s = pkg.BatchSolver(spec, B)
def run():
    s.set_inputs(X0); s.solve_al()
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 2
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
r = s.results()
import hashlib
h = hashlib.sha1(r["cost"].tobytes() + r["iters"].tobytes()).hexdigest()[:12]
print(json.dumps(dict(ms=ms, solves_per_s=B / ms * 1e3, solved=float((r["status"] == 0).mean()),
                      mean_iters=float(r["iters"][:, 2].mean()), hash=h)))
''' % ROOT
configs = [("", "", "16", "100", "0"), ("", "", "16", "100", "4"), ("", "", "32", "100", "0"), ("", "", "8", "100", "0"),
           ("variants/lib_p4.so", "16", "16", "100", "0"), ("variants/lib_p4.so", "16", "16", "100", "4")]
for lib, tile, budget, repack, tile2 in configs:
    env = dict(os.environ)
    if lib:
        env["ALTRO_B200_LIB"] = os.path.join(ROOT, "altro_cpp_b200", lib)
    if tile:
        env["ALTRO_B200_TILE"] = tile
    env["ALTRO_B200_BUDGET"] = budget
    env["ALTRO_B200_REPACK_PCT"] = repack
    env["ALTRO_B200_TILE2"] = tile2
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:]
    except subprocess.TimeoutExpired:
        line = "TIMEOUT"
    print(f"lib={lib or 'default':20s} tile2={tile2:>2s} budget={budget:>3s} repack={repack:>3s}  {line}", flush=True)
