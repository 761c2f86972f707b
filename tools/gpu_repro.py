"""Reproducibility probe (run under gpurun): two FRESH solvers on identical inputs must give
bit-identical outputs; a third one with skip_repeated_iterations must too.

For every config: solver A solves `reps` times, solver B (created afterwards, other addresses,
no history) solves once, solver C has skip_repeated_iterations = 1.  Every output (status, iteration
counters, cost, violation, X, U, K, d) is compared as raw bits (NaN-safe).  With ALTRO_B200_POISON=1 in
the environment every device allocation starts as 0xFF bytes, so a read of never-written memory
shows up as NaN/-1 instead of depending on the heap.

usage: python tools/gpu_repro.py [c2[:B]] [c3[:B]] [c4[:B]] ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import altro_cpp_b200 as pkg  # noqa: E402
from bench import workload  # noqa: E402


def outputs(s):
    r = s.results()
    X, U = s.trajectory()
    K, d = s.gains()
    return dict(status=r["status"], iters=r["iters"], cost=r["cost"], viol=r["viol"], X=X, U=U, K=K, d=d)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def diff(a, b, B, tag):
    bad = np.zeros(B, bool)
    per = {}
    for k in a:
        ne = (bits(a[k]) != bits(b[k])).reshape(B, -1).any(axis=1)
        per[k] = int(ne.sum())
        bad |= ne
    nbad = int(bad.sum())
    nan = {k: int(np.isnan(v).sum()) for k, v in b.items() if v.dtype == np.float64 and np.isnan(v).any()}
    print(f"  {tag}: differing instances {nbad}/{B} per field {per}" + (f" NaNs {nan}" if nan else ""), flush=True)
    if nbad:
        idx = np.where(bad)[0]
        print("    first:", idx[:12].tolist(), " tiles(W=8):", sorted(set((idx // 8).tolist()))[:12])
        for i in idx[:4]:
            dx = np.abs(a["X"][i] - b["X"][i]).max()
            du = np.abs(a["U"][i] - b["U"][i]).max()
            dk = np.abs(a["K"][i] - b["K"][i]).max()
            print(f"    inst {i}: iters {a['iters'][i].tolist()} vs {b['iters'][i].tolist()} status {a['status'][i]} vs "
                  f"{b['status'][i]} dcost {a['cost'][i] - b['cost'][i]:.3e} dX {dx:.3e} dU {du:.3e} dK {dk:.3e}")
            kx = np.where(np.abs(a["X"][i] - b["X"][i]).max(axis=1) > 0)[0]
            kk = np.where(np.abs(a["K"][i] - b["K"][i]).reshape(a["K"].shape[1], -1).max(axis=1) > 0)[0]
            print(f"      knots with dX != 0: {kx[:6].tolist()}..({len(kx)})  knots with dK != 0: {kk[:6].tolist()}..({len(kk)})")
    return nbad


def main():
    cfgs = sys.argv[1:] or ["c3", "c2", "c4"]
    reps = int(os.environ.get("REPS", "3"))
    total_bad = 0
    for c in cfgs:
        name, _, b = c.partition(":")
        spec, gen, B, desc = workload(name)
        B = int(b) if b else B
        X0 = gen(spec, B)
        print(f"{name} B={B} poison={os.environ.get('ALTRO_B200_POISON', '0')}  ({desc})", flush=True)
        A = pkg.BatchSolver(spec, B)
        first = None
        for rep in range(reps):
            A.set_inputs(X0)
            A.solve_al()
            o = outputs(A)
            if first is None:
                first = o
            else:
                total_bad += diff(first, o, B, f"A re-solve {rep}")
        Bs = pkg.BatchSolver(spec, B)
        Bs.set_inputs(X0)
        Bs.solve_al()
        total_bad += diff(first, outputs(Bs), B, "fresh solver B vs A")
        o2 = pkg.default_options()
        o2.skip_repeated_iterations = 1
        C = pkg.BatchSolver(spec, B, options=o2)
        C.set_inputs(X0)
        C.solve_al()
        total_bad += diff(first, outputs(C), B, "fresh solver C (skip_repeated_iterations) vs A")
        st, cnt = np.unique(first["status"], return_counts=True)
        print(f"  status histogram {dict(zip(st.tolist(), cnt.tolist()))} mean iters {first['iters'][:, 2].mean():.2f}")
        del A, Bs, C
    print("TOTAL differing:", total_bad)
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
