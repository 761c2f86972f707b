"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by altro_cpp_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ALTRO_ORACLE_VARIANT=native selects the -O3 -march=native build (bench.py's timed CPU arm runs in a
# subprocess with it set); the default is the parity build every check uses.
_NATIVE = os.environ.get("ALTRO_ORACLE_VARIANT", "") == "native"
_LIB_NAME = "liboracle_native.so" if _NATIVE else "liboracle.so"
_LIB_PATH = os.path.join(_HERE, _LIB_NAME)
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("altro_oracle_capi.cpp", "altro_oracle.hpp", "altro_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale or (_NATIVE and not _native_built_here()):
        subprocess.check_call(["make", "-C", _HERE, "-B", _LIB_NAME], stdout=subprocess.DEVNULL)
        if _NATIVE:
            with open(_LIB_PATH + ".host", "w") as f:
                f.write(_host_id())
    return _LIB_PATH


def _host_id() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                import hashlib
                return hashlib.sha1(line.encode()).hexdigest()
    except Exception:
        pass
    return "unknown"


def _native_built_here() -> bool:
    """-march=native objects are only valid on the CPU they were built on."""
    try:
        return open(_LIB_PATH + ".host").read() == _host_id()
    except Exception:
        return False


class Options(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "max_iterations_total", "max_iterations_outer", "max_iterations_inner",
        "bp_reg_fail_threshold", "check_forwardpass_bounds", "line_search_max_iterations",
        "reset_duals", "_pad")] + [(n, ctypes.c_double) for n in (
        "cost_tolerance", "gradient_tolerance", "bp_reg_increase_factor", "bp_reg_initial",
        "bp_reg_max", "bp_reg_min", "state_max", "control_max", "line_search_lower_bound",
        "line_search_upper_bound", "line_search_decrease_factor", "constraint_tolerance",
        "maximum_penalty", "initial_penalty", "penalty_scaling")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp = ctypes.c_void_p
        L.altro_oracle_solver_create.restype = vp
        L.altro_oracle_solver_create.argtypes = [vp, ctypes.c_int]
        L.altro_oracle_solver_cost.restype = ctypes.c_double
        L.altro_oracle_solver_max_violation_stored.restype = ctypes.c_double
        L.altro_oracle_solver_max_penalty.restype = ctypes.c_double
        for name in ("destroy", "rollout", "update_expansions", "backward_pass", "forward_pass",
                     "update_convergence_statistics", "solve_ilqr", "solve_al", "update_duals",
                     "update_penalties", "cost", "max_violation_stored", "max_penalty"):
            getattr(L, "altro_oracle_solver_" + name).argtypes = [vp]
        L.altro_oracle_problem_destroy.argtypes = [vp]
        _lib = L
    return _lib


def default_options() -> Options:
    o = Options()
    lib().altro_oracle_default_options(ctypes.byref(o))
    return o


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class OracleProblem:
    def __init__(self, spec):
        self.spec = spec
        self.handle = spec.build(lib(), "altro_oracle_")

    def __del__(self):
        try:
            lib().altro_oracle_problem_destroy(self.handle)
        except Exception:
            pass


class OracleSolver:
    """Single-instance, step-wise solver: mirrors iLQR<n,m> / AugmentedLagrangianiLQR<n,m>."""

    STAT = dict(cost=0, alpha=1, z=2, gradient=3, cost_decrease=4, regularization=5,
                violations=6, max_penalty=7)

    def __init__(self, spec, use_constraints: bool = True, options: Optional[Options] = None):
        self.spec = spec
        self.prob = OracleProblem(spec)
        self.n, self.m, self.N = spec.n, spec.m, spec.N
        L = lib()
        self.h = ctypes.c_void_p(L.altro_oracle_solver_create(self.prob.handle, int(use_constraints)))
        if not self.h:
            raise RuntimeError(f"oracle has no instantiation for n={spec.n}, m={spec.m}")
        self.options = options or default_options()
        self.set_options(self.options)
        self.set_controls(spec.initial_controls())

    def __del__(self):
        try:
            lib().altro_oracle_solver_destroy(self.h)
        except Exception:
            pass

    def _call(self, name, *args):
        return getattr(lib(), "altro_oracle_solver_" + name)(self.h, *args)

    def set_options(self, o: Options):
        self.options = o
        self._call("set_options", ctypes.byref(o))

    def set_controls(self, U):
        U = _arr(U).reshape(self.N, self.m)
        self._call("set_controls", _p(U))

    def set_states(self, X):
        X = _arr(X).reshape(self.N + 1, self.n)
        self._call("set_states", _p(X))

    def set_initial_state(self, x0):
        x0 = _arr(x0)
        self._call("set_initial_state", _p(x0))

    def set_penalty(self, rho):
        self._call("set_penalty", ctypes.c_double(rho))

    def set_duals(self, k, lam):
        lam = _arr(lam)
        self._call("set_duals", ctypes.c_int(k), _p(lam))

    def rollout(self): self._call("rollout")
    def cost(self) -> float: return float(self._call("cost"))
    def update_expansions(self): self._call("update_expansions")
    def backward_pass(self): self._call("backward_pass")
    def forward_pass(self): self._call("forward_pass")
    def update_convergence_statistics(self): self._call("update_convergence_statistics")
    def solve_ilqr(self): self._call("solve_ilqr")
    def solve_al(self): self._call("solve_al")
    def update_duals(self): self._call("update_duals")
    def update_penalties(self): self._call("update_penalties")
    def max_violation_stored(self) -> float: return float(self._call("max_violation_stored"))
    def max_penalty(self) -> float: return float(self._call("max_penalty"))

    def trajectory(self):
        X = np.zeros((self.N + 1, self.n))
        U = np.zeros((self.N, self.m))
        self._call("get_trajectory", _p(X), _p(U))
        return X, U

    def gains(self):
        """K[k] is m x n (returned as [N, m, n]), d[k] is m."""
        K = np.zeros((self.N, self.n, self.m))  # column-major m x n == C-order [n][m]
        d = np.zeros((self.N, self.m))
        self._call("get_gains", _p(K), _p(d))
        return np.ascontiguousarray(K.transpose(0, 2, 1)), d

    def ctg(self, k):
        P = np.zeros((self.n, self.n))
        p = np.zeros(self.n)
        self._call("get_ctg", ctypes.c_int(k), _p(P), _p(p))
        return P.T.copy(), p

    def expansion(self, k):
        n, m = self.n, self.m
        lxx = np.zeros((n, n)); lxu = np.zeros((m, n)); luu = np.zeros((m, m))
        lx = np.zeros(n); lu = np.zeros(m); jac = np.zeros((n + m, n))
        self._call("get_expansion", ctypes.c_int(k), _p(lxx), _p(lxu), _p(luu), _p(lx), _p(lu), _p(jac))
        return dict(lxx=lxx.T.copy(), lxu=lxu.T.copy(), luu=luu.T.copy(), lx=lx, lu=lu,
                    A=jac.T[:, :n].copy(), B=jac.T[:, n:].copy())

    def action_value(self, k):
        n, m = self.n, self.m
        Qxx = np.zeros((n, n)); Qxu = np.zeros((m, n)); Quu = np.zeros((m, m))
        Qx = np.zeros(n); Qu = np.zeros(m)
        self._call("get_action_value", ctypes.c_int(k), _p(Qxx), _p(Qxu), _p(Quu), _p(Qx), _p(Qu))
        return dict(Qxx=Qxx.T.copy(), Qxu=Qxu.T.copy(), Quu=Quu.T.copy(), Qx=Qx, Qu=Qu)

    def duals(self, k):
        p = int(self._call("num_duals", ctypes.c_int(k)))
        lam = np.zeros(p)
        if p:
            self._call("get_duals", ctypes.c_int(k), _p(lam))
        return lam

    def status(self):
        out = (ctypes.c_int * 4)()
        self._call("get_status", out)
        return dict(status=out[0], iterations_inner=out[1], iterations_outer=out[2],
                    iterations_total=out[3])

    def scalars(self):
        out = (ctypes.c_double * 5)()
        self._call("get_scalars", out)
        return dict(rho=out[0], drho=out[1], deltaV=(out[2], out[3]), initial_cost=out[4])

    def stat(self, name):
        buf = np.zeros(1024)
        ln = int(self._call("get_stat", ctypes.c_int(self.STAT[name]), _p(buf), ctypes.c_int(1024)))
        return buf[:ln].copy()

    def counters(self):
        out = (ctypes.c_long * 4)()
        self._call("get_counters", out)
        return dict(backward=out[0], rollout_cl=out[1], cost=out[2], expansions=out[3])


def solve_batch(spec, X0, U0=None, options: Optional[Options] = None, use_al: bool = True,
                nthreads: int = 1, want_traj: bool = True, want_gains: bool = True):
    """One independent solve per instance, spread over host threads.

    X0: [B, n].  U0: None (spec.initial_controls() for every instance), [N, m] or [B, N, m].
    """
    prob = OracleProblem(spec)
    n, m, N = spec.n, spec.m, spec.N
    X0 = _arr(X0).reshape(-1, n)
    B = X0.shape[0]
    o = options or default_options()
    U0s = None
    U0n = _arr(spec.initial_controls())
    if U0 is not None:
        U0 = _arr(U0)
        if U0.ndim == 3:
            U0s = U0
        else:
            U0n = U0
    X = np.zeros((B, N + 1, n)) if want_traj else None
    U = np.zeros((B, N, m)) if want_traj else None
    K = np.zeros((B, N, n, m)) if want_gains else None
    d = np.zeros((B, N, m)) if want_gains else None
    cost = np.zeros(B); viol = np.zeros(B)
    status = np.zeros(B, dtype=np.int32); iters = np.zeros((B, 3), dtype=np.int32)
    f = lib().altro_oracle_solve_batch
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.POINTER(Options), ctypes.c_int, ctypes.c_int, _dp, _dp, _dp,
                  ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    nul = ctypes.cast(None, _dp)
    rc = f(prob.handle, ctypes.byref(o), int(use_al), B, _p(X0), _p(U0s) if U0s is not None else nul,
           _p(U0n), int(nthreads), _p(X) if want_traj else nul, _p(U) if want_traj else nul,
           _p(K) if want_gains else nul, _p(d) if want_gains else nul, _p(cost), _p(viol),
           status.ctypes.data_as(_ip), iters.ctypes.data_as(_ip))
    if rc != 0:
        raise RuntimeError(f"altro_oracle_solve_batch failed: {rc}")
    out = dict(cost=cost, viol=viol, status=status, iters=iters)
    if want_traj:
        out.update(X=X, U=U)
    if want_gains:
        out.update(K=np.ascontiguousarray(K.transpose(0, 1, 3, 2)), d=d)
    return out
