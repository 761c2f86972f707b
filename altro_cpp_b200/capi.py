"""ctypes binding of include/altro_b200.h.

``BatchSolver`` mirrors the reference's ``AugmentedLagrangianiLQR<n,m>`` / ``iLQR<n,m>``
surface (altro/augmented_lagrangian/al_solver.hpp:28, altro/ilqr/ilqr.hpp:47) for a batch
of independent instances of one ``ProblemSpec``: same method names in snake_case, same
meaning of every option, status codes of ``SolverStatus`` (solver_stats.hpp:20-31).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import build as _build
from .problems import ProblemSpec

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p


class SolverError(RuntimeError):
    pass


class Options(ctypes.Structure):
    """altro_b200_options == SolverOptions numeric fields (solver_options.hpp:19-65)."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "max_iterations_total", "max_iterations_outer", "max_iterations_inner",
        "bp_reg_fail_threshold", "check_forwardpass_bounds", "line_search_max_iterations",
        "reset_duals", "skip_repeated_iterations")] + [(n, ctypes.c_double) for n in (
        "cost_tolerance", "gradient_tolerance", "bp_reg_increase_factor", "bp_reg_initial",
        "bp_reg_max", "bp_reg_min", "state_max", "control_max", "line_search_lower_bound",
        "line_search_upper_bound", "line_search_decrease_factor", "constraint_tolerance",
        "maximum_penalty", "initial_penalty", "penalty_scaling")]


_lib = None


def lib():
    """Loads (building if stale) the CUDA library.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        path = os.environ.get("ALTRO_B200_LIB") or _build.LIB  # override: A/B-testing build variants
        if path == _build.LIB and _build.is_stale():
            path = _build.build()
        if not os.path.exists(path):
            raise SolverError("libaltro_b200.so is missing: run altro_cpp_b200/build.py (no CPU fallback)")
        L = ctypes.CDLL(path)
        L.altro_b200_last_error.restype = ctypes.c_char_p
        L.altro_b200_version.restype = ctypes.c_char_p
        L.altro_b200_backward_pass_bytes.restype = ctypes.c_size_t
        L.altro_b200_backward_pass_bytes.argtypes = [_vp]
        L.altro_b200_device_bytes.restype = ctypes.c_size_t
        L.altro_b200_device_bytes.argtypes = [_vp]
        L.altro_b200_kernel_launches.restype = ctypes.c_int64
        L.altro_b200_kernel_launches.argtypes = [_vp]
        L.altro_b200_solver_create.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(_vp)]
        L.altro_b200_solver_destroy.argtypes = [_vp]
        L.altro_b200_problem_destroy.argtypes = [_vp]
        L.altro_b200_set_default_engine.argtypes = [ctypes.c_int]
        L.altro_b200_set_default_engine.restype = None
        L.altro_b200_solver_engine.argtypes = [_vp]
        if hasattr(L, "altro_b200_multi_create"):  # absent from older builds loaded through ALTRO_B200_LIB
            L.altro_b200_multi_create.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                  ctypes.c_int, ctypes.POINTER(_vp)]
            L.altro_b200_multi_destroy.argtypes = [_vp]
            L.altro_b200_multi_num_devices.argtypes = [_vp]
        _lib = L
    return _lib


def default_options() -> Options:
    o = Options()
    lib().altro_b200_default_options(ctypes.byref(o))
    return o


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().altro_b200_last_error().decode()
        raise SolverError(f"{what} failed ({rc}): {msg}")


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else ctypes.cast(None, _dp)


def _stream_ptr(stream) -> _vp:
    if stream is None:
        return _vp(0)
    if hasattr(stream, "cuda_stream"):
        return _vp(stream.cuda_stream)
    return _vp(int(stream))


ENGINES = {"fused": 0, "phased": 1}


def set_default_engine(name: Optional[str]):
    """Engine of solvers created from now on: "fused", "phased" or None (environment variable
    ALTRO_B200_ENGINE, else the built-in default).  Both engines return the same results."""
    lib().altro_b200_set_default_engine(-1 if name is None else ENGINES[name])


class BatchSolver:
    """B independent AL-iLQR (or plain iLQR) solves of one problem on one GPU."""

    def __init__(self, spec: ProblemSpec, batch: int, use_constraints: bool = True, device: int = 0,
                 options: Optional[Options] = None):
        self.spec, self.B = spec, int(batch)
        self.n, self.m, self.N = spec.n, spec.m, spec.N
        self.use_constraints = bool(use_constraints)
        L = lib()
        self._prob = spec.build(L, "altro_b200_")
        h = _vp()
        rc = L.altro_b200_solver_create(self._prob, self.B, int(use_constraints), int(device), ctypes.byref(h))
        if rc != 0:
            msg = L.altro_b200_last_error().decode()
            L.altro_b200_problem_destroy(self._prob)
            self._prob = None
            raise SolverError(f"solver_create failed ({rc}): {msg}")
        self._h = h
        self.options = options or default_options()
        self.set_options(self.options)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().altro_b200_solver_destroy(self._h)
                self._h = None
            if getattr(self, "_prob", None):
                lib().altro_b200_problem_destroy(self._prob)
                self._prob = None
        except Exception:
            pass

    def _call(self, name, *args):
        fn = getattr(lib(), "altro_b200_" + name)
        _check(fn(self._h, *args), name)

    @property
    def engine(self) -> str:
        e = int(lib().altro_b200_solver_engine(self._h))
        return {v: k for k, v in ENGINES.items()}.get(e, "?")

    # ---- options / inputs -----------------------------------------------------------
    def set_options(self, o: Options):
        self.options = o
        self._call("solver_set_options", ctypes.byref(o))

    def set_inputs(self, X0, U0=None, stream=None):
        """X0 [B,n]; U0 None (spec.u0 at every knot), [N,m] (shared) or [B,N,m]."""
        X0 = _f64(X0, (self.B, self.n))
        unom = None
        U = None
        if U0 is None:
            unom = _f64(self.spec.u0 if self.spec.u0 is not None else np.zeros(self.m))
        else:
            U0 = _f64(U0)
            U = np.ascontiguousarray(np.broadcast_to(U0, (self.B, self.N, self.m)))
        self._keep = (X0, U, unom)
        self._call("solver_set_inputs_host", _p(X0), _p(U), _p(unom), _stream_ptr(stream))

    def set_inputs_dev(self, x0_ptr: int, U0_ptr: int = 0, u_nominal=None, stream=None):
        """Device-resident inputs (raw device pointers, e.g. torch.Tensor.data_ptr())."""
        unom = _f64(u_nominal) if u_nominal is not None else None
        self._keep = (unom,)
        self._call("solver_set_inputs_dev", ctypes.cast(x0_ptr, _dp), ctypes.cast(U0_ptr, _dp), _p(unom),
                   _stream_ptr(stream))

    def trajectory_dev(self, X_ptr: int, U_ptr: int, stream=None):
        """Current trajectories into device buffers X [B][N+1][n], U [B][N][m] (raw device pointers)."""
        self._call("get_trajectory_dev", ctypes.cast(X_ptr, _dp), ctypes.cast(U_ptr, _dp), _stream_ptr(stream))

    def set_states(self, X, stream=None):
        X = _f64(np.broadcast_to(_f64(X), (self.B, self.N + 1, self.n)))
        self._call("solver_set_states_host", _p(X), _stream_ptr(stream))

    def set_penalty(self, rho: float, stream=None):
        self._call("solver_set_penalty", ctypes.c_double(rho), _stream_ptr(stream))

    def set_initial_cost(self, cost: float, stream=None):
        """stats.initial_cost of every instance (step-wise iterations; a whole solve sets it itself)."""
        self._call("solver_set_initial_cost", ctypes.c_double(cost), _stream_ptr(stream))

    def set_duals(self, k: int, lam, stream=None):
        lam = _f64(lam)
        self._call("solver_set_duals_host", ctypes.c_int(k), _p(lam), ctypes.c_int(lam.size), _stream_ptr(stream))

    # ---- solves ------------------------------------------------------------------------
    def solve_al(self, stream=None): self._call("solve_al", _stream_ptr(stream))
    def solve_ilqr(self, stream=None): self._call("solve_ilqr", _stream_ptr(stream))

    def solve_al_host(self, X0, U0=None, want_traj=True, stream=None, out=None):
        """Reference-facing call with host buffers (H2D + Solve() + D2H inside)."""
        X0 = _f64(X0, (self.B, self.n))
        unom, U = None, None
        if U0 is None:
            unom = _f64(self.spec.u0 if self.spec.u0 is not None else np.zeros(self.m))
        else:
            U = np.ascontiguousarray(np.broadcast_to(_f64(U0), (self.B, self.N, self.m)))
        out = out or self.alloc_outputs(want_traj)
        nul = ctypes.cast(None, _dp)
        self._call("solve_al_host", _p(X0), _p(U), _p(unom),
                   _p(out["X"]) if want_traj else nul, _p(out["U"]) if want_traj else nul,
                   _p(out["cost"]), _p(out["viol"]), out["status"].ctypes.data_as(_ip),
                   out["iters"].ctypes.data_as(_ip), _stream_ptr(stream))
        return out

    def alloc_outputs(self, want_traj=True):
        out = dict(cost=np.zeros(self.B), viol=np.zeros(self.B), status=np.zeros(self.B, np.int32),
                   iters=np.zeros((self.B, 3), np.int32))
        if want_traj:
            out["X"] = np.zeros((self.B, self.N + 1, self.n))
            out["U"] = np.zeros((self.B, self.N, self.m))
        return out

    # ---- step-wise phases (public methods of iLQR<n,m>) --------------------------------
    def solve_setup(self, stream=None): self._call("solve_setup", _stream_ptr(stream))
    def rollout(self, stream=None): self._call("rollout", _stream_ptr(stream))
    def cost(self, stream=None) -> np.ndarray:
        self._call("cost", _stream_ptr(stream))
        return self.results()["cost"]
    def update_expansions(self, stream=None): self._call("update_expansions", _stream_ptr(stream))
    def backward_pass(self, stream=None): self._call("backward_pass", _stream_ptr(stream))
    def backward_pass_stream_only(self, stream=None): self._call("backward_pass_stream_only", _stream_ptr(stream))
    def backward_pass_insolve(self, stream=None): self._call("backward_pass_insolve", _stream_ptr(stream))
    def backward_pass_fused(self, stream=None): self._call("backward_pass_fused", _stream_ptr(stream))
    def forward_pass(self, stream=None): self._call("forward_pass", _stream_ptr(stream))
    def update_convergence_statistics(self, stream=None):
        self._call("update_convergence_statistics", _stream_ptr(stream))
    def update_duals(self, stream=None): self._call("update_duals", _stream_ptr(stream))
    def update_penalties(self, stream=None): self._call("update_penalties", _stream_ptr(stream))

    # ---- outputs ----------------------------------------------------------------------
    def trajectory(self, stream=None):
        X = np.zeros((self.B, self.N + 1, self.n)); U = np.zeros((self.B, self.N, self.m))
        self._call("get_trajectory_host", _p(X), _p(U), _stream_ptr(stream))
        return X, U

    def gains(self, stream=None):
        """K [B,N,m,n], d [B,N,m]."""
        K = np.zeros((self.B, self.N, self.n, self.m)); d = np.zeros((self.B, self.N, self.m))
        self._call("get_gains_host", _p(K), _p(d), _stream_ptr(stream))
        return np.ascontiguousarray(K.transpose(0, 1, 3, 2)), d

    def ctg(self, k, stream=None):
        P = np.zeros((self.B, self.n, self.n)); p = np.zeros((self.B, self.n))
        self._call("get_ctg_host", ctypes.c_int(k), _p(P), _p(p), _stream_ptr(stream))
        return np.ascontiguousarray(P.transpose(0, 2, 1)), p

    def costs(self, stream=None):
        """per-knot costs [B,N+1] of the last cost() / update_expansions() (the reference's GetCosts())."""
        c = np.zeros((self.B, self.N + 1))
        self._call("get_costs_host", _p(c), _stream_ptr(stream))
        return c

    def expansion(self, k, stream=None):
        n, m, B = self.n, self.m, self.B
        A = np.zeros((B, n, n)); Bm = np.zeros((B, m, n)); lxx = np.zeros((B, n, n)); lxu = np.zeros((B, m, n))
        luu = np.zeros((B, m, m)); lx = np.zeros((B, n)); lu = np.zeros((B, m))
        self._call("get_expansion_host", ctypes.c_int(k), _p(A), _p(Bm), _p(lxx), _p(lxu), _p(luu), _p(lx),
                   _p(lu), _stream_ptr(stream))
        t = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
        return dict(A=t(A), B=t(Bm), lxx=t(lxx), lxu=t(lxu), luu=t(luu), lx=lx, lu=lu)

    def duals(self, k, stream=None):
        p = ctypes.c_int(0)
        self._call("get_duals_host", ctypes.c_int(k), ctypes.cast(None, _dp), ctypes.byref(p), _stream_ptr(stream))
        lam = np.zeros((self.B, max(p.value, 1)))
        if p.value:
            self._call("get_duals_host", ctypes.c_int(k), _p(lam), ctypes.byref(p), _stream_ptr(stream))
        return lam[:, :p.value]

    def constraint_values(self, k, stream=None):
        """c(x_k, u_k) of the current trajectory, ALCost row order -> [B][p_k]."""
        pmax, pk = ctypes.c_int(0), ctypes.c_int(0)
        self._call("get_duals_host", ctypes.c_int(k), ctypes.cast(None, _dp), ctypes.byref(pmax), _stream_ptr(stream))
        c = np.zeros((self.B, max(pmax.value, 1)))
        self._call("get_constraint_values_host", ctypes.c_int(k), _p(c) if pmax.value else ctypes.cast(None, _dp),
                   ctypes.byref(pk), _stream_ptr(stream))
        return c[:, :pk.value]

    def results(self, stream=None):
        out = self.alloc_outputs(want_traj=False)
        self._call("get_results_host", _p(out["cost"]), _p(out["viol"]), out["status"].ctypes.data_as(_ip),
                   out["iters"].ctypes.data_as(_ip), _stream_ptr(stream))
        return out

    def ilqr_status(self, stream=None):
        st = np.zeros(self.B, np.int32)
        self._call("get_ilqr_status_host", st.ctypes.data_as(_ip), _stream_ptr(stream))
        return st

    def scalars(self, stream=None):
        names = ("reg", "dV0", "dV1", "alpha", "z", "dJ", "grad", "penalty", "initial_cost")
        arrs = {k: np.zeros(self.B) for k in names}
        self._call("get_scalars_host", *[_p(arrs[k]) for k in names], _stream_ptr(stream))
        return arrs

    # ---- per-iteration statistics (SolverStats vectors, altro/common/solver_stats.hpp:54-61) ----
    HISTORY_COLS = ("cost", "alpha", "z", "gradient", "cost_decrease", "regularization", "violations", "max_penalty")

    def enable_history(self, instances: int = 1, rows: int = 1000):
        """Record one row per inner iteration for the first `instances` instances (call before solving;
        a recording solver runs on the fused engine)."""
        self._call("solver_enable_history", ctypes.c_int(int(instances)), ctypes.c_int(int(rows)))
        self._hist_rows = int(rows)

    def history(self, instance: int = 0, stream=None):
        """-> dict name -> array[iterations]: the value each SolverStats vector holds in the row that
        iteration wrote (carry-forward included)."""
        rows = np.zeros((self._hist_rows, len(self.HISTORY_COLS)))
        n = ctypes.c_int(0)
        self._call("get_history_host", ctypes.c_int(int(instance)), _p(rows), ctypes.c_int(self._hist_rows),
                   ctypes.byref(n), _stream_ptr(stream))
        return {k: rows[:n.value, i].copy() for i, k in enumerate(self.HISTORY_COLS)}

    # ---- measurement ------------------------------------------------------------------
    def backward_pass_bytes(self) -> int: return int(lib().altro_b200_backward_pass_bytes(self._h))
    def kernel_launches(self) -> int: return int(lib().altro_b200_kernel_launches(self._h))
    def device_bytes(self) -> int: return int(lib().altro_b200_device_bytes(self._h))


class MultiBatchSolver:
    """The same batch sharded over several GPUs of one node behind ONE handle (altro_b200_multi_*):
    contiguous slices, one solver and one host thread per device, no collective.
    `devices` may name a device more than once (two slices on one GPU: how the 1-GPU tests exercise it)."""

    def __init__(self, spec: ProblemSpec, batch: int, devices, use_constraints: bool = True,
                 options: Optional[Options] = None):
        self.spec, self.B = spec, int(batch)
        self.n, self.m, self.N = spec.n, spec.m, spec.N
        self.devices = [int(d) for d in devices]
        L = lib()
        self._prob = spec.build(L, "altro_b200_")
        h = _vp()
        devs = (ctypes.c_int * len(self.devices))(*self.devices)
        rc = L.altro_b200_multi_create(self._prob, self.B, int(use_constraints), devs, len(self.devices), ctypes.byref(h))
        if rc != 0:
            msg = L.altro_b200_last_error().decode()
            L.altro_b200_problem_destroy(self._prob)
            self._prob = None
            raise SolverError(f"multi_create failed ({rc}): {msg}")
        self._h = h
        if options is not None:
            _check(L.altro_b200_multi_set_options(self._h, ctypes.byref(options)), "multi_set_options")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().altro_b200_multi_destroy(self._h)
                self._h = None
            if getattr(self, "_prob", None):
                lib().altro_b200_problem_destroy(self._prob)
                self._prob = None
        except Exception:
            pass

    def solve_al_host(self, X0, U0=None, want_traj=True, out=None):
        X0 = _f64(X0, (self.B, self.n))
        unom, U = None, None
        if U0 is None:
            unom = _f64(self.spec.u0 if self.spec.u0 is not None else np.zeros(self.m))
        else:
            U = np.ascontiguousarray(np.broadcast_to(_f64(U0), (self.B, self.N, self.m)))
        if out is None:
            out = dict(cost=np.zeros(self.B), viol=np.zeros(self.B), status=np.zeros(self.B, np.int32),
                       iters=np.zeros((self.B, 3), np.int32))
            if want_traj:
                out["X"] = np.zeros((self.B, self.N + 1, self.n))
                out["U"] = np.zeros((self.B, self.N, self.m))
        nul = ctypes.cast(None, _dp)
        rc = lib().altro_b200_multi_solve_al_host(
            self._h, _p(X0), _p(U), _p(unom), _p(out["X"]) if want_traj else nul, _p(out["U"]) if want_traj else nul,
            _p(out["cost"]), _p(out["viol"]), out["status"].ctypes.data_as(_ip), out["iters"].ctypes.data_as(_ip))
        _check(rc, "multi_solve_al_host")
        return out

    def timings(self):
        """Device ms of the last solve per device: inputs H2D + packing, solve, results D2H."""
        G = len(self.devices)
        a, b, c = np.zeros(G), np.zeros(G), np.zeros(G)
        _check(lib().altro_b200_multi_last_timings(self._h, _p(a), _p(b), _p(c)), "multi_last_timings")
        return dict(scatter_ms=a, solve_ms=b, gather_ms=c)


def register_model(name: str, cuda_source: str, n: int, m: int, nparams: int = 0) -> int:
    """Registers a plug-in dynamics model (the CUDA source of a functor struct, concept documented in
    csrc/device.cuh; example plugins/cartpole.cuh) -> model id for ProblemSpec(model=...)."""
    mid = ctypes.c_int(-1)
    rc = lib().altro_b200_register_model(name.encode(), cuda_source.encode(), int(n), int(m), int(nparams), ctypes.byref(mid))
    _check(rc, "register_model")
    return mid.value


def precompile_model(model_id: int, n: int = 0, m: int = 0) -> str:
    """Compiles the model's kernels into the on-disk module cache now (NVRTC; needs no GPU) -> cache file.
    n, m: for a built-in model id whose instantiation is produced at run time (triple integrator, dof != 2)."""
    buf = ctypes.create_string_buffer(1024)
    if n or m:
        _check(lib().altro_b200_precompile_builtin_model(int(model_id), int(n), int(m), buf, 1024), "precompile_model")
    else:
        _check(lib().altro_b200_precompile_model(int(model_id), buf, 1024), "precompile_model")
    return buf.value.decode()
