#!/usr/bin/env python
"""bench.py — AL-iLQR solves/sec (batched) on N B200s, with the backward-pass HBM roofline and
the CPU baseline timed beside it.  Contract: see the task statement ("Measurement").

A "step" is one pass of the hot path over one batch: B independent AL-iLQR solves of
BASELINE.json config C2 (unicycle n=3, m=2, N=100, 3 obstacles + control bounds + goal;
instance 0 nominal, the rest perturbed in x0 — SURVEY.md 8d), from the initial guess to the
reference's termination, default SolverOptions.

  value  = (B * n_gpus) / (max-over-ranks device time of one step), inputs resident in HBM
  e2e    = same metric through the reference-facing host-buffer call altro_b200_solve_al_host
           (pinned host x0 in, X/U/cost/viol/status/iters out, copies inside the timed region)
  roofline = the materialised backward-pass kernel (k_backward_mat) timed live with CUDA events:
           algorithmic bytes (SURVEY.md 8d contract, 37,696 B per instance per pass) / duration
  cpu_baseline = the CPU oracle (a port of the reference's algorithm; the reference itself cannot
           be built here — no Eigen) on all host cores over a bounded sample of the same batch

`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from altro_cpp_b200 import problems as P  # noqa: E402

METRIC = "AL-iLQR solves/sec (batched)"
UNIT = "solves/s"


def workload(name: str):
    """-> (spec, x0 generator (spec, count) -> [count, n], default batch per GPU, description)."""
    pert = lambda scale: (lambda spec, count: P.perturbed_initial_states(spec, count, scale))
    if name == "c2":
        return P.unicycle_problem(P.K_THREE_OBSTACLES), pert(P.UNICYCLE_X0_SCALE), 16384, \
            "C2: unicycle n=3 m=2 N=100, 3 obstacles + control bounds + goal, AL-iLQR, default options"
    if name == "c3":
        return P.triple_integrator_problem(dof=2, N=50, add_constraints=True), pert(P.TRIPLE_INTEGRATOR_X0_SCALE), \
            8192, "C3: triple integrator n=6 m=2 N=50, goal + control bounds, AL-iLQR (8192 per GPU)"
    if name == "c4":
        return P.cartpole_problem(N=200), pert(P.CARTPOLE_X0_SCALE), 32768, \
            "C4: cartpole n=4 m=1 N=200, control bound, AL-iLQR"
    if name in ("c5", "c5-literal"):
        return P.random_lqr_problem(literal=name.endswith("literal")), P.normal_initial_states, 4096, \
            "C5: random LQR n=32 m=8 N=100, unconstrained (" + \
            ("SURVEY.md numbers verbatim, ill-conditioned" if name.endswith("literal") else "well-conditioned variant") + ")"
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                clocks.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(clocks)}


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_backward_mat from the committed ncu capture."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k_backward_mat_summary.txt"))):
        txt = open(path).read()
        rd = re.search(r"dram__bytes_read\.sum \[Mbyte\] = ([0-9.]+)", txt)
        wr = re.search(r"dram__bytes_write\.sum \[Mbyte\] = ([0-9.]+)", txt)
        if rd and wr:
            best = ((float(rd.group(1)) + float(wr.group(1))) * 1e6, os.path.basename(path))
    return best


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_leg(spec, X0, steps, warmup, sample, nthreads):
    """Times the CPU oracle (port of the reference algorithm) on a bounded sample."""
    from oracle import binding as ob
    ob.build()
    Xs = X0[:sample]
    times = []
    out = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = ob.solve_batch(spec, Xs, nthreads=nthreads, want_traj=False, want_gains=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sample / float(np.mean(times)), float(np.mean(times)), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--bp-iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--engine", default=None, choices=["phased", "fused"],
                    help="execution engine (default: the library's, phased); results are identical")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec, gen_x0, default_B, wl_name = workload(args.workload)
    B = args.batch or default_B
    ncores = os.cpu_count() or 1

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        X0 = gen_x0(spec, B)
        sample = min(args.cpu_sample, B)
        W = max(0, args.warmup)
        val, dt, out = cpu_leg(spec, X0, args.steps, W, sample, ncores)
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "batch_per_step": sample,
                       "note": "CPU path of the reference algorithm (oracle port; the reference cannot be "
                               "built here: Eigen absent), one independent solve per host thread"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"first {sample} instances of the {B}-instance batch per step"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist
    import altro_cpp_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    # Instances are generated once on rank 0 and scattered over NCCL (the only collective on the
    # path: it shards trivially over the batch axis); results are gathered back.
    from altro_cpp_b200.sharding import gather_rows, scatter_rows
    n, m, N = spec.n, spec.m, spec.N
    total = B * world
    X0_all_dev = None
    if rank == 0:
        X0_all_dev = torch.from_numpy(gen_x0(spec, total)).to(dev)
    if distributed:
        x0_dev = scatter_rows(X0_all_dev, total, (n,), torch.float64, dev)
    else:
        x0_dev = X0_all_dev
    X0_host = x0_dev.cpu().numpy()

    stream = torch.cuda.Stream(device=dev)
    if args.engine:
        pkg.set_default_engine(args.engine)
    solver = pkg.BatchSolver(spec, B, device=local_rank)
    unom = spec.u0

    def step():
        solver.set_inputs_dev(x0_dev.data_ptr(), 0, unom, stream=stream)
        solver.solve_al(stream=stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step()
        barrier()
        l0 = solver.kernel_launches()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        launches = solver.kernel_launches() - l0
        ms = e0.elapsed_time(e1) / args.steps
    res = solver.results()
    X_sol, U_sol = solver.trajectory()  # kept for the bit-identity check of the secondary measurement

    # ---- e2e: host buffers through altro_b200_solve_al_host (pinned memory)
    pin_x0 = torch.from_numpy(X0_host).pin_memory()
    outs = {k: torch.from_numpy(v).pin_memory() for k, v in solver.alloc_outputs(True).items()}
    outs_np = {k: v.numpy() for k, v in outs.items()}
    with torch.cuda.stream(stream):
        for _ in range(2):
            solver.solve_al_host(pin_x0.numpy(), None, True, stream=stream, out=outs_np)
        barrier()
        e2 = torch.cuda.Event(enable_timing=True)
        e3 = torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            solver.solve_al_host(pin_x0.numpy(), None, True, stream=stream, out=outs_np)
        e3.record(stream)
        barrier()
        ms_e2e = e2.elapsed_time(e3) / args.steps
    h2d = B * n * 8
    d2h = B * ((N + 1) * n + N * m) * 8 + B * (8 + 8 + 4 + 12)

    # ---- roofline: the materialised backward-pass kernel, timed live
    ms_bp = float('nan')
    bp_bytes = solver.backward_pass_bytes()
    with torch.cuda.stream(stream):
      if not args.workload.startswith('c5'):
        solver.set_inputs_dev(x0_dev.data_ptr(), 0, unom, stream=stream)
        solver.solve_setup(stream=stream)
        solver.rollout(stream=stream)
        solver.update_expansions(stream=stream)
        for _ in range(3):
            solver.backward_pass_stream_only(stream=stream)
        barrier()
        e4 = torch.cuda.Event(enable_timing=True)
        e5 = torch.cuda.Event(enable_timing=True)
        e4.record(stream)
        for _ in range(args.bp_iters):
            solver.backward_pass_stream_only(stream=stream)
        e5.record(stream)
        barrier()
        ms_bp = e4.elapsed_time(e5) / args.bp_iters
    bp_bytes = solver.backward_pass_bytes()
    peak, peak_src = measured_peak()
    have_bp = ms_bp == ms_bp
    achieved = bp_bytes / (ms_bp * 1e-3) / 1e9 if have_bp else None

    # ---- secondary measurement (not the headline): the same step with skip_repeated_iterations,
    # which accounts for provably identical repeated inner iterations without executing them;
    # results must be (and are checked to be) bit-identical to the faithful run above
    extras = None
    if not args.no_extras and not args.workload.startswith('c5'):
        o2 = pkg.default_options()
        o2.skip_repeated_iterations = 1
        solver2 = pkg.BatchSolver(spec, B, device=local_rank, options=o2)

        def step2():
            solver2.set_inputs_dev(x0_dev.data_ptr(), 0, unom, stream=stream)
            solver2.solve_al(stream=stream)
        with torch.cuda.stream(stream):
            for _ in range(2):
                step2()
            barrier()
            e6 = torch.cuda.Event(enable_timing=True)
            e7 = torch.cuda.Event(enable_timing=True)
            e6.record(stream)
            for _ in range(args.steps):
                step2()
            e7.record(stream)
            barrier()
            ms_skip = e6.elapsed_time(e7) / args.steps
        res2 = solver2.results()
        X2, U2 = solver2.trajectory()
        X1, U1 = X_sol, U_sol
        identical = bool(np.array_equal(res2["cost"], res["cost"]) and np.array_equal(res2["iters"], res["iters"])
                         and np.array_equal(res2["status"], res["status"]) and np.array_equal(X1, X2)
                         and np.array_equal(U1, U2))
        _bits = lambda a: np.ascontiguousarray(a).view(np.int64) if a.dtype == np.float64 else a
        _f = {"cost": (res["cost"], res2["cost"]), "iters": (res["iters"], res2["iters"]), "status": (res["status"], res2["status"]), "X": (X1, X2), "U": (U1, U2)}
        sys.stderr.write("R1DIAG mismatch per field (bitwise): " + str({k: int((_bits(a) != _bits(b)).reshape(B, -1).any(axis=1).sum()) for k, (a, b) in _f.items()}) +
                         " | array_equal per field: " + str({k: bool(np.array_equal(a, b)) for k, (a, b) in _f.items()}) +
                         " | nan X: " + str(int(np.isnan(X1).sum())) + " nan cost " + str(int(np.isnan(res["cost"]).sum())) + "\n")
        extras = {"skip_repeated_iterations": {
            "ms_per_step_rank0": ms_skip, "value_rank0": B / (ms_skip * 1e-3), "unit": UNIT,
            "bit_identical_to_faithful_run": identical,
            "note": "opt-in option, off in the headline: inner iterations that provably repeat the previous "
                    "one (same Z, duals, penalty and regularisation after a fully failed line search) are "
                    "counted, not executed"}}
        del solver2

    # ---- reduce over ranks: max time, summed work
    t = torch.tensor([ms, ms_e2e, ms_bp if have_bp else 0.0], dtype=torch.float64, device=dev)
    stats = torch.tensor([float((res["status"] == 0).sum()), float(res["iters"][:, 2].sum()),
                          float(res["iters"][:, 2].max())], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        smax = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(smax, op=dist.ReduceOp.MAX)
        stats[2] = smax[2]
        # gather per-instance results on rank 0 — the output side of the scatter
        cost_all = gather_rows(torch.from_numpy(res["cost"]).to(dev).reshape(-1, 1), total)
        assert rank != 0 or cost_all.shape[0] == total
    ms, ms_e2e, ms_bp_max = [float(v) for v in t.tolist()]

    cpu = None
    if rank == 0 and not args.no_cpu:
        sample = min(args.cpu_sample, B)
        val, dt, out = cpu_leg(spec, X0_host, 1, 0, sample, ncores)
        same = float(np.mean(np.all(out["iters"] == res["iters"][:sample], axis=1)
                             & (out["status"] == res["status"][:sample])))
        n1 = min(48, sample)
        val1, dt1, _ = cpu_leg(spec, X0_host, 1, 0, n1, 1)
        cpu = {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"first {sample} instances of rank 0's batch, one pass ({dt:.1f} s)",
               "single_thread_value": val1, "single_thread_sample": f"first {n1} instances ({dt1:.1f} s)",
               "same_status_and_iterations_as_gpu": same}

    if rank == 0:
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl_name, "batch_per_gpu": B, "global_batch": total,
                       "parallelism": f"batch-sharded x{world} (no data-path collective)",
                       "engine": solver.engine,
                       "l2": f"working set {solver.device_bytes() / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                       "solved_fraction": float(stats[0].item() / total),
                       "mean_ilqr_iterations": float(stats[1].item() / total),
                       "max_ilqr_iterations": float(stats[2].item())},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": None if not have_bp else {"kernel": "k_backward_mat (materialised backward pass, TMA-streamed)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": (profiled_traffic() or (None, None))[0] if args.workload == "c2" and B == 16384 else None,
                         "traffic_source": (profiled_traffic() or (None, None))[1],
                         "bytes_per_launch": bp_bytes, "ms_per_launch": ms_bp},
            "solve_engine": {"engine": solver.engine,
                             "kernels": ("k_solve(outer/start) || k_update_expansions -> k_backward_mat(TMA) -> "
                                         "k_ls_wide/k_ls_deep (k_roll/k_cost/k_acc when few instances remain)")
                             if solver.engine == "phased" else "k_solve (fused persistent AL-iLQR)",
                             "backward_passes_per_step": float(stats[1].item()),
                             "contract_GBps": float(stats[1].item()) * (bp_bytes / B) / (ms * 1e-3) / 1e9},
            "extras": extras,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
