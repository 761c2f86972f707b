"""Phased vs fused engine on the GPU: parity against the oracle + timings (run under gpurun)."""
import os, sys, time, json, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
from oracle import binding as ob


def digest(s):
    r = s.results(); X, U = s.trajectory()
    return hashlib.sha1(r["cost"].tobytes() + r["iters"].tobytes() + r["status"].tobytes() + X.tobytes() + U.tobytes()).hexdigest()[:12]


def parity():
    cases = [("unicycle3obs", P.unicycle_problem(P.K_THREE_OBSTACLES), P.UNICYCLE_X0_SCALE, True, 200),
             ("turn90-ilqr", P.unicycle_problem(P.K_TURN90), P.UNICYCLE_X0_SCALE, False, 70),
             ("triple", P.triple_integrator_problem(dof=2, N=50, add_constraints=True), P.TRIPLE_INTEGRATOR_X0_SCALE, True, 70),
             ("cartpole", P.cartpole_problem(N=200), P.CARTPOLE_X0_SCALE, True, 33)]
    for name, spec, scale, al, B in cases:
        X0 = P.perturbed_initial_states(spec, B, scale)
        ref = ob.solve_batch(spec, X0, use_al=al, nthreads=16, want_gains=False)
        out = {}
        for eng in ("fused", "phased"):
            pkg.set_default_engine(eng)
            s = pkg.BatchSolver(spec, B, use_constraints=al); s.set_inputs(X0)
            t0 = time.time(); (s.solve_al if al else s.solve_ilqr)(); torch.cuda.synchronize(); dt = time.time() - t0
            r = s.results(); X, U = s.trajectory()
            same = np.all(r["iters"] == ref["iters"], axis=1) & (r["status"] == ref["status"])
            err = max(np.max(np.abs(X[i] - ref["X"][i])) for i in np.where(same)[0]) if same.any() else float("nan")
            out[eng] = digest(s)
            print(f"{name:14s} {eng:7s} engine={s.engine} same_path={same.mean():.3f} max|dX|={err:.2e} "
                  f"cost_err={np.max(np.abs(r['cost'][same]-ref['cost'][same])):.2e} t={dt*1e3:.1f} ms launches={s.kernel_launches()} digest={out[eng]}", flush=True)
            if not same.all():
                bad = np.where(~same)[0][:5]
                print("   mismatching:", bad.tolist(), r["iters"][bad].tolist(), ref["iters"][bad].tolist(), r["status"][bad].tolist(), ref["status"][bad].tolist())
        print(f"{name:14s} engines bit-identical: {out['fused'] == out['phased']}", flush=True)


def timing(B=16384, reps=3, skip=0):
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    for eng in ("phased", "fused"):
        pkg.set_default_engine(eng)
        o = pkg.default_options(); o.skip_repeated_iterations = skip
        s = pkg.BatchSolver(spec, B, options=o)
        def run():
            s.set_inputs(X0); s.solve_al()
        run(); torch.cuda.synchronize()
        l0 = s.kernel_launches()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        r = s.results()
        print(json.dumps(dict(engine=eng, B=B, skip=skip, ms=ms, solves_per_s=B / ms * 1e3, solved=float((r["status"] == 0).mean()),
                              mean_iters=float(r["iters"][:, 2].mean()), launches=(s.kernel_launches() - l0) // reps, digest=digest(s))), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"): parity()
    if what in ("all", "timing"):
        timing()
        timing(skip=1)
        timing(B=1024)
