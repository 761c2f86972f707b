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
for rep in range(int(os.environ.get('REPS', '3'))):
    s.set_inputs(X0); s.solve_al()
    r = s.results(); X, U = s.trajectory()
    hs.append(hashlib.sha1(r["cost"].tobytes() + r["iters"].tobytes() + X.tobytes() + U.tobytes()).hexdigest()[:10])
    if prev is not None:
        dX = np.abs(prev[2] - X).reshape(B, -1).max(axis=1); dU = np.abs(prev[3] - U).reshape(B, -1).max(axis=1)
        bad = np.where(np.any(prev[0] != r["iters"], axis=1) | (prev[1] != r["cost"]) | (dX > 0) | (dU > 0) | (prev[4] != r["viol"]))[0]
        if len(bad):
            print("   differing instances:", len(bad), bad[:16].tolist())
            print("     iters", prev[0][bad[:8]].tolist(), r["iters"][bad[:8]].tolist())
            print("     dcost", (prev[1][bad[:8]] - r["cost"][bad[:8]]).tolist())
            print("     dX", dX[bad[:8]].tolist(), "dU", dU[bad[:8]].tolist(), "dviol", (prev[4][bad[:8]] - r["viol"][bad[:8]]).tolist())
            b0 = bad[0]; kx = np.where(np.abs(prev[2][b0] - X[b0]).max(axis=1) > 0)[0]; print("     knots with dX != 0 for", b0, ":", kx[:10].tolist(), "...", len(kx))
    prev = (r["iters"].copy(), r["cost"].copy(), X.copy(), U.copy(), r["viol"].copy())
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
