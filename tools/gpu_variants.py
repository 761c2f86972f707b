"""A/B timing of build variants (run under gpurun): python tools/gpu_variants.py lib1.so lib2.so ...
Each library in its own process; 'default' = the in-tree build."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, hashlib
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
reps = 3
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
r = s.results()
h = hashlib.sha1(r["cost"].tobytes() + r["iters"].tobytes()).hexdigest()[:12]
print(json.dumps(dict(engine=s.engine, ms=round(ms, 2), solves_per_s=round(B / ms * 1e3), hash=h)))
''' % ROOT
libs = sys.argv[1:] or ["default"]
for lib in libs:
    env = dict(os.environ)
    if lib != "default":
        env["ALTRO_B200_LIB"] = os.path.join(ROOT, "altro_cpp_b200", "variants", lib)
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-600:]
    except subprocess.TimeoutExpired:
        line = "TIMEOUT"
    print(f"{lib:24s} {line}", flush=True)
