"""Experiment: G solvers of B/G instances each, driven concurrently from G host threads on G streams,
against one solver of B instances (run under gpurun).  The phases of a slot are latency-bound at
different moments (backward pass: one serial chain per instance, 5 % of the warps a B200 holds);
independent groups on separate streams let the GPU overlap them.
usage: python tools/gpu_two_streams.py [B] [groups...]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
groups = [int(g) for g in sys.argv[2:]] or [1, 2, 4]
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
dev = torch.device("cuda", 0)
x0_dev = torch.from_numpy(X0).to(dev)
ref = None
for G in groups:
    per = B // G
    solvers = [pkg.BatchSolver(spec, per) for _ in range(G)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
    xs = [x0_dev[g * per:(g + 1) * per].contiguous() for g in range(G)]

    def work(g):
        solvers[g].set_inputs_dev(xs[g].data_ptr(), 0, spec.u0, stream=streams[g])
        solvers[g].solve_al(stream=streams[g])

    def step():
        th = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize(dev)

    for _ in range(2): step()
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps): step()
    dt = (time.perf_counter() - t0) / reps
    cost = np.concatenate([s.results()["cost"] for s in solvers])
    if ref is None: ref = cost
    print(f"groups={G} per={per}: {dt*1e3:.1f} ms/step  {B/dt:.0f} solves/s  identical_to_first={np.array_equal(cost, ref)}", flush=True)
    del solvers
