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
  cpu_baseline = the CPU oracle (a port of the reference's algorithm, bit-identical to the reference's own sources
           compiled on this repo's Eigen stand-in — oracle/_ref, tests/test_oracle_vs_reference_build.py — and 8 x
           faster than that build, so it is the stronger CPU arm) on all host cores over a bounded sample of the batch

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

ROOT = os.path.dirname(os.path.abspath(__file__))
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
    # preferred: the capture of the very launch `roofline` times (altro_b200_backward_pass_insolve, full batch);
    # else the same kernel captured inside a solve (its 25th launch: still the full batch)
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k_backward_mat_phased_summary.txt"))) + \
        sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k_backward_mat_insolve_fullbatch_summary.txt")))
    for path in paths:
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
    """Times the CPU oracle (port of the reference algorithm) on a bounded sample (this process: the
    parity build, -O3 -ffp-contract=off)."""
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


CPU_BUILD = "g++ -O3 -march=native (oracle/Makefile liboracle_native.so, built on this host)"


def cpu_leg_native(workload_name, batch, steps, warmup, sample, nthreads):
    """The timed CPU arm: the same port compiled as BASELINE.md section 3 says (-O3 -march=native), run in
    a child process so that the parity build stays the checker of this one.
    -> (solves/s, seconds per step, {"status", "iters"} of the sample)."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out_npz = os.path.join(tmp, "cpu.npz")
        env = dict(os.environ, ALTRO_ORACLE_VARIANT="native")
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-child", out_npz, "--workload", workload_name,
               "--batch", str(batch), "--cpu-sample", str(sample), "--steps", str(steps), "--warmup", str(warmup),
               "--cpu-threads", str(nthreads)]
        subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL)
        d = np.load(out_npz)
        dt = float(d["dt"])
        return sample / dt, dt, {"status": d["status"], "iters": d["iters"]}


def reference_build_leg(workload_name, spec, X0, per_thread=3):
    """solves/s of oracle/_ref/libaltro_ref.so (the reference's own solver sources compiled on the Eigen stand-in,
    oracle/build_ref.py), one independent solve per host thread like the timed CPU arm, for the workloads its entry
    point covers; None when the library did not travel or anything goes wrong — this is context for cpu_baseline
    (the oracle port is bit-identical to this build and faster, so the port is the baseline), never the headline."""
    try:
        import ctypes
        from concurrent.futures import ThreadPoolExecutor
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libaltro_ref.so")
        if workload_name != "c2" or not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        n, m, N = spec.n, spec.m, spec.N
        threads = os.cpu_count() or 1
        count = min(int(np.asarray(X0).shape[0]), threads * per_thread)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)

        def solve(b):
            X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
            x0 = np.ascontiguousarray(X0[b], dtype=np.float64)
            lib.altro_ref_unicycle(ctypes.c_int(1), ctypes.c_int(1), ptr(x0), None, ptr(X), ptr(U), ptr(sc), ptr(it))
            return int(it[3])

        t0 = time.perf_counter()
        solve(0)
        single = time.perf_counter() - t0
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as pool:  # the C call releases the GIL
            iters = list(pool.map(solve, range(count)))
        dt = time.perf_counter() - t0
        return {"value": count / dt, "unit": UNIT, "cores": threads, "kind": "reference",
                "single_thread_value": 1.0 / single,
                "build": "the reference's altro/**/*.cpp + examples compiled where they lie on this repo's Eigen stand-in "
                         "(oracle/build_ref.py; Eigen itself is absent from the image), g++ -O2",
                "sample": f"first {count} instances, {threads} threads ({dt:.1f} s); mean iLQR iterations "
                          f"{sum(iters) / max(1, len(iters)):.1f}"}
    except Exception:  # noqa: BLE001 - context only
        return None


def cpu_child(args):
    spec, gen_x0, default_B, _ = workload(args.workload)
    X0 = gen_x0(spec, args.batch or default_B)
    _, dt, out = cpu_leg(spec, X0, args.steps, args.warmup, args.cpu_sample, args.cpu_threads)
    np.savez(args.cpu_child, dt=dt, status=out["status"], iters=out["iters"])


def strong_scaling_c3(pkg, torch, dist, world, rank, dev, local_rank, stream, steps):
    """BASELINE config C3 at its stated size — 65 536 randomised triple-integrator instances — with the
    TOTAL fixed and cut into `world` contiguous slices (strong scaling): NCCL scatter of x0 from rank 0,
    per-rank solve, NCCL gather of the results, each timed on the device (max over ranks); rank 0 then
    solves all 65 536 instances on its own GPU and checks that the gathered results are the same bits
    (reference analogue: the nthreads-equivalence tests, test/ilqr/ilqr_class_test.cpp:130-160)."""
    from altro_cpp_b200.sharding import gather_rows, scatter_rows
    spec, gen_x0, _, _ = workload("c3")
    total = 65536
    per = total // world
    n, m, N = spec.n, spec.m, spec.N
    distributed = world > 1

    def timed(fn):
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    X0_all = torch.from_numpy(gen_x0(spec, total)).to(dev) if rank == 0 else None
    if distributed:
        scatter_rows(X0_all, total, (n,), torch.float64, dev)  # warm-up of the NCCL path
        x0_dev, scatter_ms = timed(lambda: scatter_rows(X0_all, total, (n,), torch.float64, dev))
    else:
        x0_dev, scatter_ms = X0_all, 0.0
    solver = pkg.BatchSolver(spec, per, device=local_rank)

    def step():
        solver.set_inputs_dev(x0_dev.data_ptr(), 0, spec.u0, stream=stream)
        solver.solve_al(stream=stream)

    with torch.cuda.stream(stream):
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    Xl = torch.empty((per, N + 1, n), dtype=torch.float64, device=dev)
    Ul = torch.empty((per, N, m), dtype=torch.float64, device=dev)
    solver.trajectory_dev(Xl.data_ptr(), Ul.data_ptr())
    res = solver.results()
    scal = torch.from_numpy(np.stack([res["cost"], res["viol"], res["status"].astype(np.float64),
                                      res["iters"][:, 2].astype(np.float64)], axis=1)).to(dev)
    if distributed:
        gather_rows(scal, total)  # warm-up
        (Xg, Ug, Sg), gather_ms = timed(lambda: (gather_rows(Xl, total), gather_rows(Ul, total), gather_rows(scal, total)))
    else:
        (Xg, Ug, Sg), gather_ms = (Xl, Ul, scal), 0.0
    out = None
    if rank == 0:
        one = solver if world == 1 else pkg.BatchSolver(spec, total, device=local_rank)
        if world > 1:
            one.set_inputs_dev(X0_all.data_ptr(), 0, spec.u0)
            one.solve_al()
        X1 = torch.empty((total, N + 1, n), dtype=torch.float64, device=dev)
        U1 = torch.empty((total, N, m), dtype=torch.float64, device=dev)
        one.trajectory_dev(X1.data_ptr(), U1.data_ptr())
        r1 = one.results()
        S1 = np.stack([r1["cost"], r1["viol"], r1["status"].astype(np.float64), r1["iters"][:, 2].astype(np.float64)], axis=1)
        same = bool(torch.equal(Xg.view(torch.int64), X1.view(torch.int64)) and torch.equal(Ug.view(torch.int64), U1.view(torch.int64))
                    and np.array_equal(Sg.cpu().numpy().view(np.int64), S1.view(np.int64)))
        out = {"workload": "C3: triple integrator n=6 m=2 N=50, goal + control bounds, 65536 instances in total",
               "scaling": "strong", "global_batch": total, "batch_per_gpu": per, "ms_per_step": ms,
               "value": total / (ms * 1e-3), "unit": UNIT,
               "scatter_ms": scatter_ms, "gather_ms": gather_ms,
               "scatter_bytes": total * n * 8, "gather_bytes": total * ((N + 1) * n + N * m + 4) * 8,
               "gathered_bit_identical_to_1gpu_solve": same,
               "solved_fraction": float((S1[:, 2] == 0).mean())}
        assert same, "sharded results differ from the one-GPU solve of the same instances"
    return out


FP64_TENSOR_PEAK_TFLOPS = 37.1  # profiles/r02_dmma_rate.txt: mma.sync.m8n8k4.f64, 148 SMs x 8 warps x 4 chains


def c5_roofline(n, m, N, backward_passes, ms_step, engine):
    """Large-state path: the whole solve is ONE kernel (large_mma.cuh), bound by the fp64 tensor pipe.
    Algorithmic flops of one Riccati step (CalcActionValueExpansion + CalcCostToGo as dense products,
    DESIGN.md section 7): A'P, (A'P)A: 2 n^3 each; B'P, (B'P)A: 2 m n^2 each; (B'P)B, Quu K: 2 m^2 n each;
    K'(Quu K + Qux), Qxu K: 2 m n^2 each.  achieved = those flops x knot points x backward passes of the
    step / the step's duration (the step is that one kernel plus ~10 us of host-side launches)."""
    step_flops = 2.0 * (2 * n ** 3 + 4 * m * n * n + 2 * m * m * n)
    flops = step_flops * N * backward_passes
    achieved = flops / (ms_step * 1e-3) / 1e12
    return {"kernel": "k_solve_large_mma — whole AL-iLQR solves, one instance per warp, dense products on "
                      "mma.sync.m8n8k4.f64" if engine == "phased" else "k_solve_large — exact-order kernel, one instance per CTA",
            "bound": "tensor", "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
            "frac": achieved / FP64_TENSOR_PEAK_TFLOPS,
            "peak_source": "measured fp64 mma.sync rate on this pool's B200 (tools/micro/dmma_rate.cu, "
                           "profiles/r02_dmma_rate.txt); MEASURED_PEAKS.json has no fp64 entry",
            "traffic": None, "flops_per_launch": flops, "ms_per_launch": ms_step,
            "flops_per_riccati_step": step_flops}


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
    ap.add_argument("--cpu-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-threads", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_child:
        return cpu_child(args)

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
        val, dt, out = cpu_leg_native(args.workload, B, args.steps, W, sample, ncores)
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "batch_per_step": sample,
                       "note": "CPU path of the reference algorithm: the oracle port, bit-identical to the reference's own "
                               "sources built on the Eigen stand-in (oracle/_ref) and faster than that build; one "
                               "independent solve per host thread",
                       "build": CPU_BUILD},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "cpu_model": cpu_model(),
                             "build": CPU_BUILD,
                             "sample": f"first {sample} instances of the {B}-instance batch per step",
                             "reference_build": reference_build_leg(args.workload, spec, X0)},
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
    scatter_ms_weak = gather_ms_weak = 0.0
    if distributed:
        scatter_rows(X0_all_dev, total, (n,), torch.float64, dev)  # warm-up of the NCCL path
        torch.cuda.synchronize(dev)
        dist.barrier()
        ts0 = torch.cuda.Event(enable_timing=True)
        ts1 = torch.cuda.Event(enable_timing=True)
        ts0.record()
        x0_dev = scatter_rows(X0_all_dev, total, (n,), torch.float64, dev)
        ts1.record()
        torch.cuda.synchronize(dev)
        scatter_ms_weak = ts0.elapsed_time(ts1)
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

    # ---- roofline: the backward-pass kernel, timed live.  `roofline` is the variant a solve launches
    # (k_backward_mat<..., phased>: per-instance phase mask + regularisation hand-off) at the full batch;
    # the bare streaming variant (writes K, d only, no solve state) is reported beside it.
    ms_bp = float('nan')
    ms_bp_stream = float('nan')
    bp_bytes = solver.backward_pass_bytes()

    def time_bp(fn):
        for _ in range(3):
            fn(stream=stream)
        barrier()
        ea = torch.cuda.Event(enable_timing=True)
        eb = torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for _ in range(args.bp_iters):
            fn(stream=stream)
        eb.record(stream)
        barrier()
        return ea.elapsed_time(eb) / args.bp_iters

    with torch.cuda.stream(stream):
      if not args.workload.startswith('c5'):
        solver.set_inputs_dev(x0_dev.data_ptr(), 0, unom, stream=stream)
        solver.solve_setup(stream=stream)
        solver.rollout(stream=stream)
        solver.update_expansions(stream=stream)
        ms_bp_stream = time_bp(solver.backward_pass_stream_only)
        if solver.engine == "phased":
            ms_bp = time_bp(solver.backward_pass_insolve)
        else:
            ms_bp = ms_bp_stream
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
        # raw-bit comparison per field (NaN-safe: an instance that diverged to NaN on both sides is equal)
        bits = lambda a: np.ascontiguousarray(a).view(np.int64) if a.dtype == np.float64 else a
        fields = {"cost": (res["cost"], res2["cost"]), "viol": (res["viol"], res2["viol"]),
                  "iters": (res["iters"], res2["iters"]), "status": (res["status"], res2["status"]),
                  "X": (X1, X2), "U": (U1, U2)}
        mismatch = {k: int((bits(a) != bits(b)).reshape(B, -1).any(axis=1).sum()) for k, (a, b) in fields.items()}
        identical = not any(mismatch.values())
        extras = {"skip_repeated_iterations": {
            "ms_per_step_rank0": ms_skip, "value_rank0": B / (ms_skip * 1e-3), "unit": UNIT,
            "bit_identical_to_faithful_run": identical,
            "instances_differing_per_field": mismatch,
            "nan_instances": int(np.isnan(X1).reshape(B, -1).any(axis=1).sum()),
            "note": "opt-in option, off in the headline: inner iterations that provably repeat the previous "
                    "one (same Z, duals, penalty and regularisation after a fully failed line search) are "
                    "counted, not executed"}}
        del solver2

    # ---- strong scaling at C3's stated size (all ranks take part; reported by rank 0)
    strong = None
    if not args.no_extras and args.workload == "c2":
        strong = strong_scaling_c3(pkg, torch, dist, world, rank, dev, local_rank, stream, max(3, args.steps))

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
        cost_dev = torch.from_numpy(res["cost"]).to(dev).reshape(-1, 1)
        gather_rows(cost_dev, total)
        torch.cuda.synchronize(dev)
        dist.barrier()
        tg0 = torch.cuda.Event(enable_timing=True)
        tg1 = torch.cuda.Event(enable_timing=True)
        tg0.record()
        cost_all = gather_rows(cost_dev, total)
        tg1.record()
        torch.cuda.synchronize(dev)
        gather_ms_weak = tg0.elapsed_time(tg1)
        assert rank != 0 or cost_all.shape[0] == total
    ms, ms_e2e, ms_bp_max = [float(v) for v in t.tolist()]

    cpu = None
    if rank == 0 and not args.no_cpu:
        sample = min(args.cpu_sample, B)
        # timed arm: the -march=native build in a child process; checker: the parity build here
        val, dt, out_native = cpu_leg_native(args.workload, B, 1, 0, sample, ncores)
        _, _, out = cpu_leg(spec, X0_host, 1, 0, sample, ncores)
        agree = lambda o: float(np.mean(np.all(o["iters"] == res["iters"][:sample], axis=1)
                                        & (o["status"] == res["status"][:sample])))
        n1 = min(48, sample)
        val1, dt1, _ = cpu_leg_native(args.workload, B, 1, 0, n1, 1)
        cpu = {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "cpu_model": cpu_model(),
               "build": CPU_BUILD,
               "sample": f"first {sample} instances of rank 0's batch, one pass ({dt:.1f} s)",
               "single_thread_value": val1, "single_thread_sample": f"first {n1} instances ({dt1:.1f} s)",
               "reference_build": reference_build_leg(args.workload, spec, X0_host),
               "same_status_and_iterations_as_gpu": agree(out),
               "same_status_and_iterations_as_gpu_native_build": agree(out_native)}

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
            "roofline": c5_roofline(n, m, N, float(stats[1].item()) / world, ms, solver.engine) if args.workload.startswith("c5")
            else None if not have_bp else {
                "kernel": "k_backward_mat<phased> — the backward pass as a solve launches it (TMA-streamed "
                          "materialised expansions), full batch",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": (profiled_traffic() or (None, None))[0] if args.workload == "c2" and B == 16384 else None,
                "traffic_source": (profiled_traffic() or (None, None))[1],
                "bytes_per_launch": bp_bytes, "ms_per_launch": ms_bp,
                "stream_only_variant": {"ms_per_launch": ms_bp_stream,
                                        "achieved": bp_bytes / (ms_bp_stream * 1e-3) / 1e9,
                                        "frac": bp_bytes / (ms_bp_stream * 1e-3) / 1e9 / peak}},
            "solve_engine": {"engine": solver.engine,
                             "kernels": ("k_outer_* (dense list, every 2nd slot) -> k_update_expansions -> "
                                         "k_backward_mat(TMA) -> k_ls_wide -> k_ls_deep (k_roll/k_cost/k_acc when few "
                                         "instances remain)")
                             if solver.engine == "phased" and not args.workload.startswith("c5")
                             else ("k_solve_large_mma (whole solves, one instance per warp)" if solver.engine == "phased"
                                   else "k_solve_large (whole solves, one instance per CTA, exact order)")
                             if args.workload.startswith("c5") else "k_solve (fused persistent AL-iLQR)",
                             "backward_passes_per_step": float(stats[1].item()),
                             "contract_GBps": float(stats[1].item()) * (bp_bytes / B) / (ms * 1e-3) / 1e9},
            "extras": extras,
            "strong_scaling": strong,
            "io": {"scatter_ms": scatter_ms_weak, "gather_ms": gather_ms_weak,
                   "note": "NCCL scatter of x0 from rank 0 / gather of the per-instance costs to rank 0, outside the "
                           "timed region (0 on one GPU: the inputs are generated in place)"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
