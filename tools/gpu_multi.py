"""The batch sharded over the GPUs of a node behind ONE C-ABI handle (altro_b200_multi_*): per-device
timings, throughput through host buffers, and bit identity with a one-GPU solve.
usage: python tools/gpu_multi.py [instances per GPU]        (run under gpurun --gpus N)"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

per = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ndev = torch.cuda.device_count()
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
B = per * ndev
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
multi = pkg.MultiBatchSolver(spec, B, devices=list(range(ndev)))
multi.solve_al_host(X0)  # warm-up (allocations, module loads)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); out = multi.solve_al_host(X0); ts.append(time.perf_counter() - t0)
tm = multi.timings()
one = pkg.BatchSolver(spec, per, device=0)
ref = one.solve_al_host(X0[:per])
same = all(np.array_equal(np.ascontiguousarray(out[k][:per]).view(np.int64) if out[k].dtype == np.float64 else out[k][:per],
                          np.ascontiguousarray(ref[k]).view(np.int64) if ref[k].dtype == np.float64 else ref[k])
           for k in ("status", "iters", "cost", "viol", "X", "U"))
last = pkg.BatchSolver(spec, per, device=ndev - 1)
ref2 = last.solve_al_host(X0[B - per:])
same2 = all(np.array_equal(np.ascontiguousarray(out[k][B - per:]).view(np.int64) if out[k].dtype == np.float64 else out[k][B - per:],
                           np.ascontiguousarray(ref2[k]).view(np.int64) if ref2[k].dtype == np.float64 else ref2[k])
            for k in ("status", "iters", "cost", "viol", "X", "U"))
print(json.dumps({"what": "altro_b200_multi_solve_al_host, C2, host buffers in and out (wall clock of the one call)",
                  "devices": ndev, "instances_per_device": per, "global_batch": B,
                  "ms_per_call": min(ts) * 1e3, "solves_per_s": B / min(ts),
                  "per_device_solve_ms": [float(x) for x in tm["solve_ms"]],
                  "first_shard_bit_identical_to_one_gpu_solver": bool(same),
                  "last_shard_bit_identical_to_a_solver_on_its_device": bool(same2)}))
