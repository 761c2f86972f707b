"""One-shot GPU diagnostics (run under gpurun): parity diffs per quantity and sweep timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
from oracle import binding as ob

def diffs():
    spec = P.unicycle_problem(P.K_TURN90)
    B = 32
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    s = pkg.BatchSolver(spec, B, use_constraints=False)
    s.set_inputs(X0)
    r = ob.OracleSolver(spec, use_constraints=False); r.set_initial_state(X0[1])
    s.solve_setup(); s.rollout(); r.rollout()
    Xg, Ug = s.trajectory(); Xo, Uo = r.trajectory()
    print("rollout max|dX| per knot (first 5, last 5):", np.abs(Xg[1]-Xo).max(axis=1)[[0,1,2,3,4,96,97,98,99,100]])
    print("cost", s.cost()[1], r.cost())
    s.update_expansions(); r.update_expansions()
    for k in (0, 1, 50, 99, 100):
        e = s.expansion(k); eo = r.expansion(k)
        print("k", k, {nm: float(np.abs(e[nm][1]-eo[nm]).max()) for nm in ("A","B","lxx","lxu","luu","lx","lu")})
    s.backward_pass(); r.backward_pass()
    K, d = s.gains(); Ko, do = r.gains()
    print("K diff per knot (0,50,99):", [float(np.abs(K[1][k]-Ko[k]).max()) for k in (0,50,99)])
    print("d diff per knot (0,50,99):", [float(np.abs(d[1][k]-do[k]).max()) for k in (0,50,99)], d[1][0], do[0])
    P0,p0 = s.ctg(0); Po,po = r.ctg(0)
    print("P0 diff", np.abs(P0[1]-Po).max(), "p0 diff", np.abs(p0[1]-po).max())
    sc = s.scalars(); so = r.scalars()
    print("dV", sc["dV0"][1], so["deltaV"][0], sc["dV1"][1], so["deltaV"][1], "reg", sc["reg"][1], so["rho"])
    s.forward_pass(); r.forward_pass()
    print("alpha", s.scalars()["alpha"][1], r.stat("alpha"), "cost", s.results()["cost"][1], r.stat("cost"))
    # whole solves, small batch, which instances differ and how
    for al, scen in ((False, P.K_TURN90), (True, P.K_TURN90), (True, P.K_THREE_OBSTACLES)):
        spec = P.unicycle_problem(scen)
        X0 = P.perturbed_initial_states(spec, 64, P.UNICYCLE_X0_SCALE)
        s = pkg.BatchSolver(spec, 64, use_constraints=al); s.set_inputs(X0)
        (s.solve_al if al else s.solve_ilqr)()
        res = s.results()
        ref = ob.solve_batch(spec, X0, use_al=al, nthreads=16, want_gains=False)
        same = np.all(res["iters"] == ref["iters"], axis=1) & (res["status"] == ref["status"])
        print("solve al=%s scen=%d same=%.3f" % (al, scen, same.mean()))
        print(" gpu iters[:8]", res["iters"][:8].tolist(), "status", res["status"][:8].tolist())
        print(" ref iters[:8]", ref["iters"][:8].tolist(), "status", ref["status"][:8].tolist())
        print(" cost gpu/ref [:4]", res["cost"][:4], ref["cost"][:4])

def timings():
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    B = 16384
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    s = pkg.BatchSolver(spec, B); s.set_inputs(X0)
    s.solve_setup(); s.rollout(); s.cost()
    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    print("rollout+cost sweep (open)  ms:", timeit(s.rollout))
    print("cost sweep                 ms:", timeit(lambda: s.cost.__func__(s) if False else s._call("cost", 0)))
    print("update_expansions          ms:", timeit(s.update_expansions))
    print("backward_mat (ctg)         ms:", timeit(s.backward_pass))
    print("backward_mat stream only   ms:", timeit(s.backward_pass_stream_only))
    print("backward fused             ms:", timeit(s.backward_pass_fused))
    print("forward_pass (line search) ms:", timeit(s.forward_pass, reps=1))
    o = pkg.default_options(); o.max_iterations_inner = 1; o.max_iterations_outer = 1
    s.set_options(o)
    def solve1():
        s.set_inputs(X0); s.solve_al()
    print("solve_al 1 inner x 1 outer ms:", timeit(solve1, reps=2))
    print("   ls alphas:", np.unique(s.scalars()["alpha"], return_counts=True))
    o.max_iterations_inner = 10; s.set_options(o)
    print("solve_al 10 inner x 1 outer ms:", timeit(solve1, reps=2))

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "diffs"): diffs()
    if what in ("all", "timings"): timings()
    if what == "sanitize":
        # small ragged batches through every kernel (run under compute-sanitizer)
        for scen, al, B in ((P.K_THREE_OBSTACLES, True, 21), (P.K_TURN90, False, 9)):
            spec = P.unicycle_problem(scen, N=30)
            X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
            o = pkg.default_options(); o.max_iterations_inner = 6; o.max_iterations_outer = 3
            s = pkg.BatchSolver(spec, B, use_constraints=al, options=o); s.set_inputs(X0)
            (s.solve_al if al else s.solve_ilqr)()
            s.solve_setup(); s.rollout(); s.cost(); s.update_expansions(); s.backward_pass(); s.forward_pass()
            s.update_convergence_statistics(); s.backward_pass_fused()
            if al: s.update_duals(); s.update_penalties()
            print("sanitize run ok", scen, al, B, s.results()["iters"][:3].tolist())
    if what == "ncu_solve":
        # one k_solve launch of the C2 batch: AL init + rollout + 3 inner iterations of every instance
        spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
        B = 16384
        X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
        o = pkg.default_options(); o.max_iterations_inner = 3; o.max_iterations_outer = 1
        s = pkg.BatchSolver(spec, B, options=o); s.set_inputs(X0); s.solve_al(); torch.cuda.synchronize()
    if what == "ncu_bp_insolve":
        spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
        B = 16384
        X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
        s = pkg.BatchSolver(spec, B); s.set_inputs(X0)
        s.solve_setup(); s.rollout(); s.update_expansions()
        for _ in range(6): s.backward_pass_insolve()
        torch.cuda.synchronize()
    if what == "ncu_bp":
        spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
        B = 16384
        X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
        s = pkg.BatchSolver(spec, B); s.set_inputs(X0)
        s.solve_setup(); s.rollout(); s.update_expansions()
        for _ in range(3): s.backward_pass_stream_only()
        torch.cuda.synchronize()
