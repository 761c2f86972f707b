"""Warm-start / MPC streaming mode (docs/Overview.dox:49-54 of the reference):
a batch of controllers re-solves every control tick from the previous solution shifted by one
knot, with reset_duals = false and initial_penalty = 0 (solver_options.hpp:47-48,
al_solver.hpp:292-297), everything device-resident.  Reports latency per tick.

    python tools/mpc_latency.py [batch] [ticks] [cap]  (run under gpurun)

cap > 0 bounds the work of a tick the way a real-time controller does, with the reference's own options
(max_iterations_total = max_iterations_inner = cap, solver_options.hpp:20-22): the tick then returns the best
iterate found in `cap` iLQR iterations instead of iterating every controller to convergence.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import altro_cpp_b200 as pkg  # noqa: E402
from altro_cpp_b200 import problems as P  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    n, m, N = spec.n, spec.m, spec.N
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    s = pkg.BatchSolver(spec, B)
    Xd = torch.empty((B, N + 1, n), dtype=torch.float64, device=dev)
    Ud = torch.empty((B, N, m), dtype=torch.float64, device=dev)
    x0d = torch.from_numpy(X0).to(dev)
    lat = []
    with torch.cuda.stream(stream):
        # cold solve (tick 0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.set_inputs_dev(x0d.data_ptr(), 0, spec.u0, stream=stream)
        s.solve_al(stream=stream)
        e1.record(stream); torch.cuda.synchronize(dev)
        cold = e0.elapsed_time(e1)
        cold_res = s.results(stream=stream)
        o = pkg.default_options()
        o.reset_duals = 0
        o.initial_penalty = 0.0
        if cap > 0:
            o.max_iterations_total = cap
            o.max_iterations_inner = cap
        s.set_options(o)
        for t in range(ticks):
            s.trajectory_dev(Xd.data_ptr(), Ud.data_ptr(), stream=stream)
            e0.record(stream)
            # the plant moved one step along the plan: new x0 = x_1, plan shifted by one knot
            x0d.copy_(Xd[:, 1, :])
            Ushift = torch.cat([Ud[:, 1:, :], Ud[:, -1:, :]], dim=1).contiguous()
            s.set_inputs_dev(x0d.data_ptr(), Ushift.data_ptr(), None, stream=stream)
            s.solve_al(stream=stream)
            e1.record(stream); torch.cuda.synchronize(dev)
            lat.append(e0.elapsed_time(e1))
    r = s.results()
    print(json.dumps({
        "workload": "C2 unicycle 3 obstacles, warm-started re-solve per tick (reset_duals=0, initial_penalty=0)",
        "batch": B, "engine": s.engine, "ticks": ticks, "iteration_cap_per_tick": cap or None, "cold_solve_ms": cold,
        "cold_mean_iterations": float(cold_res["iters"][:, 2].mean()),
        "tick_ms_median": float(np.median(lat)), "tick_ms_p90": float(np.percentile(lat, 90)),
        "tick_ms_first": lat[0], "ticks_per_s_per_controller_batch": 1e3 / float(np.median(lat)),
        "controller_solves_per_s": B * 1e3 / float(np.median(lat)),
        "last_tick_mean_iterations": float(r["iters"][:, 2].mean()),
        "last_tick_max_iterations": int(r["iters"][:, 2].max()),
        "last_tick_solved_fraction": float((r["status"] == 0).mean()),
        "last_tick_max_violation_median": float(np.median(r["viol"]))}))


if __name__ == "__main__":
    main()
