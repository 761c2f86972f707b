"""Distribution of the GPU-vs-oracle differences after whole solves (run under gpurun).

For every config: solve the batch on the GPU, the first `sample` instances on the CPU oracle, and
report (a) the fraction of instances on the same discrete path (status + iteration counters),
(b) percentiles of the per-instance relative error of X, U, cost, K, d among those, (c) what the
worst instances look like (iterations, status) — the data behind the tolerances in tests/.

usage: python tools/gpu_parity_stats.py [c2[:B[:sample]]] ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import altro_cpp_b200 as pkg  # noqa: E402
from bench import workload  # noqa: E402
from oracle import binding as ob  # noqa: E402


def rel(a, b):
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    return np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))


def main():
    for c in sys.argv[1:] or ["c2:16384:2048", "c3:8192:2048", "c4:4096:256"]:
        name, B, sample = (c.split(":") + ["", ""])[:3]
        spec, gen, Bd, desc = workload(name)
        B = int(B) if B else Bd
        sample = min(int(sample) if sample else 1024, B)
        X0 = gen(spec, B)
        s = pkg.BatchSolver(spec, B)
        s.set_inputs(X0)
        s.solve_al()
        r = s.results()
        X, U = s.trajectory()
        K, d = s.gains()
        ref = ob.solve_batch(spec, X0[:sample], nthreads=os.cpu_count() or 1)
        same = np.all(r["iters"][:sample] == ref["iters"], axis=1) & (r["status"][:sample] == ref["status"])
        print(f"{name} B={B} sample={sample}: same discrete path {same.mean():.4%} ({(~same).sum()} differ)", flush=True)
        for i in np.where(~same)[0][:8]:
            print(f"   inst {i}: gpu iters {r['iters'][i].tolist()} st {r['status'][i]} | cpu iters {ref['iters'][i].tolist()} "
                  f"st {ref['status'][i]} | dcost {r['cost'][i] - ref['cost'][i]:.2e}")
        idx = np.where(same)[0]
        errs = dict(X=rel(X[idx], ref["X"][idx]), U=rel(U[idx], ref["U"][idx]),
                    cost=np.abs(r["cost"][idx] - ref["cost"][idx]) / np.maximum(1.0, np.abs(ref["cost"][idx])),
                    viol=np.abs(r["viol"][idx] - ref["viol"][idx]),
                    K=rel(K[idx], ref["K"][idx]), d=rel(d[idx], ref["d"][idx]))
        for k, e in errs.items():
            q = np.percentile(e, [50, 90, 99, 99.9, 100])
            print(f"   {k:5s} p50 {q[0]:.2e} p90 {q[1]:.2e} p99 {q[2]:.2e} p99.9 {q[3]:.2e} max {q[4]:.2e}")
        worst = idx[np.argsort(-errs["X"])[:6]]
        for i in worst:
            j = np.where(idx == i)[0][0]
            print(f"   worst X: inst {i} err {errs['X'][j]:.2e} iters {r['iters'][i].tolist()} status {r['status'][i]} "
                  f"cost {r['cost'][i]:.6g} errU {errs['U'][j]:.2e} errK {errs['K'][j]:.2e}")
        for st in np.unique(r["status"][idx]):
            m = r["status"][idx] == st
            print(f"   status {st}: {m.sum()} instances, max errX {errs['X'][m].max():.2e} max errU {errs['U'][m].max():.2e} "
                  f"max err cost {errs['cost'][m].max():.2e}")


if __name__ == "__main__":
    main()
