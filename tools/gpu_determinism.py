"""Re-solve determinism of the phased engine under different scheduling knobs (run under gpurun)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
B = int(os.environ.get("BENCH_B", "16384"))
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
s = pkg.BatchSolver(spec, B)
hs = []
prev = None
for rep in range(3):
    s.set_inputs(X0); s.solve_al()
    r = s.results(); X, U = s.trajectory()
    hs.append(hashlib.sha1(r["cost"].tobytes() + r["iters"].tobytes() + X.tobytes() + U.tobytes()).hexdigest()[:10])
    if prev is not None:
        bad = np.where(np.any(prev[0] != r["iters"], axis=1) | (prev[1] != r["cost"]))[0]
        if len(bad): print("   differing instances:", len(bad), bad[:8].tolist(), prev[0][bad[:4]].tolist(), r["iters"][bad[:4]].tolist())
    prev = (r["iters"].copy(), r["cost"].copy())
print(s.engine, hs)
''' % ROOT
configs = [{"ALTRO_B200_SPLIT_MAX": "0", "ALTRO_B200_OVERLAP": "0"},
           {"ALTRO_B200_SPLIT_MAX": "0", "ALTRO_B200_OVERLAP": "0", "ALTRO_B200_REPACK_PCT": "0"},
           {"ALTRO_B200_SPLIT_MAX": "1000000"}, {"ALTRO_B200_SPLIT_MAX": "1000000", "ALTRO_B200_REPACK_PCT": "0"},
           {"ALTRO_B200_SPLIT_MAX": "0", "ALTRO_B200_POLL": "1"}, {}]
if len(sys.argv) > 1:
    configs = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
for cfg in configs:
    env = dict(os.environ); env.update(cfg)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=200)
    print(cfg, (out.stdout.strip() or out.stderr[-400:]), flush=True)
