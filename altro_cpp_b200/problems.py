"""Problem descriptions for the batched AL-iLQR path.

Host-side mirror of the reference's problem factories:
  * ``UnicycleProblem``          -> examples/problems/unicycle.{hpp,cpp}
  * ``TripleIntegratorProblem``  -> examples/problems/triple_integrator.hpp
  * ``LQRCost``                  -> examples/quadratic_cost.hpp:29-39
  * ``ProblemSpec``              -> the data a ``problem::Problem`` holds
                                    (altro/problem/problem.hpp:65) for the closed
                                    set of device-capable functors.

A ``ProblemSpec`` is a recorded list of builder calls.  ``ProblemSpec.build(lib,
prefix)`` replays them on any library exporting ``<prefix>problem_*`` with the
signatures of include/altro_b200.h — the product library (prefix ``altro_b200_``)
or, in tests only, the CPU oracle (prefix ``altro_oracle_``).

Quirks reproduced on purpose: Q1 ``h`` is a float32 and is
promoted to double in every product; Q11 ``uref = 0``; Q13 obstacle constraints are
added regardless of ``add_constraints``; Q14 terminal cost uses ``R*0``.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

MODEL_UNICYCLE = 0
MODEL_TRIPLE_INTEGRATOR = 1
MODEL_CARTPOLE = 2
MODEL_LINEAR = 3

_dp = ctypes.POINTER(ctypes.c_double)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _colmajor(a) -> np.ndarray:
    """Flatten a 2-D matrix to the column-major double array Eigen uses."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)


def lqr_cost(Q, R, xref, uref):
    """examples/quadratic_cost.hpp:29-39 (QuadraticCost::LQRCost)."""
    Q = np.asarray(Q, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    xref = np.asarray(xref, dtype=np.float64)
    uref = np.asarray(uref, dtype=np.float64)
    n, m = Q.shape[0], R.shape[0]
    H = np.zeros((n, m))
    q = -(Q @ xref)
    r = -(R @ uref)
    c = 0.5 * float(xref @ (Q @ xref)) + 0.5 * float(uref @ (R @ uref))
    return Q, R, H, q, r, c


@dataclass
class ProblemSpec:
    n: int
    m: int
    N: int
    calls: List[Tuple] = field(default_factory=list)
    x0: Optional[np.ndarray] = None
    u0: Optional[np.ndarray] = None  # nominal initial control (held for every knot)
    xf: Optional[np.ndarray] = None  # goal state (informational)
    h: float = 0.0
    name: str = ""

    # ---- builder calls (same argument meaning as include/altro_b200.h) ----
    def set_model(self, kind: int, params=()):
        self.calls.append(("set_model", int(kind), _f64(params)))

    def set_uniform_step(self, h):
        self.h = float(np.float32(h))
        self.calls.append(("set_uniform_step", np.float32(h)))

    def set_steps(self, t, h):
        """Per-knot times and steps, float32 [N+1] each (Trajectory::SetTime / SetStep); h[N] is normally 0."""
        t = np.ascontiguousarray(t, dtype=np.float32)
        h = np.ascontiguousarray(h, dtype=np.float32)
        assert t.shape == (self.N + 1,) and h.shape == (self.N + 1,)
        self.h = None
        self.calls.append(("set_steps", t, h))

    def set_cost(self, k0: int, k1: int, Q, R, H, q, r, c):
        self.calls.append(("set_cost", int(k0), int(k1), _colmajor(Q), _colmajor(R), _colmajor(H),
                           _f64(q), _f64(r), float(c)))

    def add_goal(self, k: int, xf):
        self.calls.append(("add_goal", int(k), _f64(xf)))

    def add_control_bound(self, k: int, lb, ub):
        self.calls.append(("add_control_bound", int(k), _f64(lb), _f64(ub)))

    def add_circles(self, k: int, cx, cy, cr, xi=0, yi=1):
        self.calls.append(("add_circles", int(k), _f64(cx), _f64(cy), _f64(cr), int(xi), int(yi)))

    def set_initial_state(self, x0):
        self.x0 = _f64(x0)
        self.calls.append(("set_initial_state", self.x0))

    # ---- replay on a C library ----
    def build(self, lib, prefix: str):
        f = lambda name: getattr(lib, prefix + "problem_" + name)
        create = f("create")
        create.restype = ctypes.c_int
        create.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_void_p)]
        handle = ctypes.c_void_p()
        if create(self.n, self.m, self.N, ctypes.byref(handle)) != 0 or not handle:
            raise RuntimeError("problem_create failed")
        for call in self.calls:
            name, args = call[0], call[1:]
            fn = f(name)
            fn.restype = ctypes.c_int
            if name == "set_model":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, ctypes.c_int]
                rc = fn(handle, args[0], _p(args[1]), len(args[1]))
            elif name == "set_uniform_step":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_float]
                rc = fn(handle, ctypes.c_float(float(args[0])))
            elif name == "set_steps":
                fp = ctypes.POINTER(ctypes.c_float)
                fn.argtypes = [ctypes.c_void_p, fp, fp]
                rc = fn(handle, args[0].ctypes.data_as(fp), args[1].ctypes.data_as(fp))
            elif name == "set_cost":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [_dp] * 5 + [ctypes.c_double]
                rc = fn(handle, args[0], args[1], *[_p(a) for a in args[2:7]], args[7])
            elif name == "add_goal":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
                rc = fn(handle, args[0], _p(args[1]))
            elif name == "add_control_bound":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
                rc = fn(handle, args[0], _p(args[1]), _p(args[2]))
            elif name == "add_circles":
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp,
                               ctypes.c_int, ctypes.c_int]
                rc = fn(handle, args[0], len(args[1]), _p(args[1]), _p(args[2]), _p(args[3]),
                        args[4], args[5])
            elif name == "set_initial_state":
                fn.argtypes = [ctypes.c_void_p, _dp]
                rc = fn(handle, _p(args[0]))
            else:  # pragma: no cover
                raise ValueError(name)
            if rc != 0:
                raise RuntimeError(f"{prefix}problem_{name} failed with {rc}")
        return handle

    def initial_controls(self) -> np.ndarray:
        """InitialTrajectory(): every control = u0 (examples/problems/unicycle.hpp:84-93)."""
        U = np.zeros((self.N, self.m))
        if self.u0 is not None:
            U[:] = self.u0
        return U


# ---------------------------------------------------------------------------
# examples/problems/unicycle.{hpp,cpp}
# ---------------------------------------------------------------------------
K_TURN90 = 0
K_THREE_OBSTACLES = 1


def unicycle_problem(scenario: int = K_TURN90, N: int = 100, add_constraints: bool = True) -> ProblemSpec:
    n, m = 3, 2
    spec = ProblemSpec(n, m, N, name=f"unicycle-{'turn90' if scenario == K_TURN90 else '3obs'}-N{N}")
    spec.set_model(MODEL_UNICYCLE)
    xf = np.array([1.5, 1.5, math.pi / 2])
    x0 = np.zeros(3)
    u0 = np.full(2, 0.1)
    uref = np.zeros(2)
    v_bnd = w_bnd = 1.5
    circles = None
    if scenario == K_TURN90:  # unicycle.cpp:17-26
        tf = np.float32(3.0)
        h = np.float32(tf / np.float32(N))  # float GetTimeStep() const { return tf / N; }
        lb, ub = [-v_bnd, -w_bnd], [+v_bnd, +w_bnd]
        Q = np.eye(3) * (1e-2 * float(h))
        R = np.eye(2) * (1e-2 * float(h))
        Qf = np.eye(3) * 100.0
    else:  # unicycle.cpp:27-60
        tf = np.float32(5.0)
        h = np.float32(tf / np.float32(N))
        Q = np.eye(3) * (1.0 * float(h))
        R = np.eye(2) * (0.5 * float(h))
        Qf = np.eye(3) * 10.0
        x0 = np.zeros(3)
        xf = np.array([3.0, 3.0, 0.0])
        u0 = np.full(2, 0.01)
        scaling = 3.0
        cx = np.array([0.25, 0.5, 0.75]) * scaling
        cy = np.array([0.25, 0.5, 0.75]) * scaling
        cr = np.full(3, 0.425)
        circles = (cx, cy, cr)
        lb, ub = [0.0, -3.0], [3.0, 3.0]
    spec.set_uniform_step(h)
    if circles is not None:  # Q13: added regardless of add_constraints, before the bounds
        for k in range(1, N):
            spec.add_circles(k, *circles)
    spec.set_cost(0, N, *lqr_cost(Q, R, xf, uref))
    spec.set_cost(N, N + 1, *lqr_cost(Qf, R * 0, xf, uref))  # Q14
    if add_constraints:  # unicycle.cpp:76-83
        for k in range(N):
            spec.add_control_bound(k, lb, ub)
        spec.add_goal(N, xf)
    spec.set_initial_state(x0)
    spec.u0 = u0
    spec.xf = xf
    return spec


# ---------------------------------------------------------------------------
# examples/problems/triple_integrator.hpp (and the fixture of test/ilqr/ilqr_test.cpp:21-98)
# ---------------------------------------------------------------------------
def triple_integrator_problem(dof: int = 2, N: int = 10, add_constraints: bool = False,
                              goal_only: bool = False, h=0.1) -> ProblemSpec:
    n, m = 3 * dof, dof
    spec = ProblemSpec(n, m, N, name=f"triple-integrator-dof{dof}-N{N}")
    spec.set_model(MODEL_TRIPLE_INTEGRATOR)
    spec.set_uniform_step(np.float32(h))
    Q = np.eye(n) * 1.0
    R = np.eye(m) * 0.001
    Qf = np.eye(n) * 1e5
    xf = np.zeros(n)
    x0 = np.zeros(n)
    ubnd = np.zeros(dof)
    for i in range(dof):
        xf[i] = i + 1
        x0[i] = -(i + 1)
        ubnd[i] = 100 * (i + 1)
    uref = np.zeros(m)  # Q11: defined as zero here
    spec.set_cost(0, N, *lqr_cost(Q, R, xf, uref))
    spec.set_cost(N, N + 1, *lqr_cost(Qf, R * 0, xf, uref))
    if add_constraints:  # triple_integrator.hpp:74-91
        for k in range(N):
            spec.add_control_bound(k, -ubnd, ubnd)
        spec.add_goal(N, xf)
    elif goal_only:  # test/ilqr/ilqr_test.cpp:78-80
        spec.add_goal(N, xf)
    spec.set_initial_state(x0)
    spec.u0 = np.zeros(m)
    spec.xf = xf
    return spec


# ---------------------------------------------------------------------------
# Models that are not in the reference (BASELINE.json configs C4, C5)
# ---------------------------------------------------------------------------
def cartpole_problem(N: int = 200) -> ProblemSpec:
    n, m = 4, 1
    spec = ProblemSpec(n, m, N, name=f"cartpole-N{N}")
    spec.set_model(MODEL_CARTPOLE, [1.0, 0.2, 0.5, 9.81])  # mc, mp, l, g
    tf = np.float32(5.0)
    h = np.float32(tf / np.float32(N))
    spec.set_uniform_step(h)
    Q = np.eye(n) * (1e-2 * float(h))
    R = np.eye(m) * (1e-1 * float(h))
    Qf = np.eye(n) * 100.0
    xf = np.array([0.0, math.pi, 0.0, 0.0])
    uref = np.zeros(m)
    spec.set_cost(0, N, *lqr_cost(Q, R, xf, uref))
    spec.set_cost(N, N + 1, *lqr_cost(Qf, R * 0, xf, uref))
    for k in range(N):
        spec.add_control_bound(k, [-10.0], [10.0])
    spec.set_initial_state(np.zeros(n))
    spec.u0 = np.zeros(m)
    spec.xf = xf
    return spec


# ---------------------------------------------------------------------------
# Batch generation: counter-based splitmix64, generated once
# on the host and fed to both the GPU path and the oracle.
# ---------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_batch(B: int, dim: int, seed: int = 20240925, first: int = 0) -> np.ndarray:
    """U(-1,1) numbers, shape [B, dim]; stream = instance index (first + row)."""
    with np.errstate(over="ignore"):
        idx = (np.arange(first, first + B, dtype=np.uint64)[:, None] * np.uint64(1024)
               + np.arange(dim, dtype=np.uint64)[None, :])
        bits = _splitmix64(idx ^ _splitmix64(np.full((1, 1), seed, dtype=np.uint64)))
    u01 = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return 2.0 * u01 - 1.0


def perturbed_initial_states(spec: ProblemSpec, B: int, scale, seed: int = 20240925,
                             first: int = 0) -> np.ndarray:
    """Instance 0 is the nominal x0; instance i >= 1 is x0 + scale * U(-1,1)."""
    scale = np.asarray(scale, dtype=np.float64)
    X0 = spec.x0[None, :] + uniform_batch(B, spec.n, seed, first) * scale[None, :]
    if first == 0 and B > 0:
        X0[0] = spec.x0
    return np.ascontiguousarray(X0)


UNICYCLE_X0_SCALE = (0.3, 0.3, math.pi / 6)
TRIPLE_INTEGRATOR_X0_SCALE = (0.5,) * 6
CARTPOLE_X0_SCALE = (0.05,) * 4


def random_lqr_problem(n: int = 32, m: int = 8, N: int = 100, seed: int = 5,
                       literal: bool = False) -> ProblemSpec:
    """BASELINE config C5: discrete LTI x+ = A x + B u, A = I + g G/sqrt(n),
    B = 0.1 H, G, H i.i.d. N(0,1) from `seed`; Q = I h, Qf = 10 I, h = 0.05; unconstrained.
    Not in the reference: parity is GPU vs oracle, and the oracle vs the reference's own solver run on a linear
    functor (tests/test_oracle_vs_reference_build.py).

    literal=True takes SURVEY.md's numbers verbatim (g = 0.05, R = 0.1 h I).  That problem is
    ill-conditioned (spectral radius 1.05 over 100 steps, cond(Quu) ~ 1e8): the first backward pass
    fails its LLT at zero regularisation and the solve needs 25-50 iterations instead of the 2 the
    survey expects.  The default (g = 0.02, R = 0.1 I) is the well-conditioned problem with the
    intended behaviour: no LLT failure, alpha = 1 accepted, 2 iterations."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    Hm = rng.standard_normal((n, m))
    g = 0.05 if literal else 0.02
    A = np.eye(n) + g * G / math.sqrt(n)
    Bm = 0.1 * Hm
    spec = ProblemSpec(n, m, N, name=f"random-lqr-n{n}-m{m}-N{N}" + ("-literal" if literal else ""))
    spec.set_model(MODEL_LINEAR, np.concatenate([_colmajor(A), _colmajor(Bm)]))
    h = np.float32(0.05)
    spec.set_uniform_step(h)
    Q = np.eye(n) * float(h)
    R = np.eye(m) * ((0.1 * float(h)) if literal else 0.1)
    Qf = np.eye(n) * 10.0
    xf = np.zeros(n)
    uref = np.zeros(m)
    spec.set_cost(0, N, *lqr_cost(Q, R, xf, uref))
    spec.set_cost(N, N + 1, *lqr_cost(Qf, R * 0, xf, uref))
    spec.set_initial_state(np.zeros(n))
    spec.u0 = np.zeros(m)
    spec.xf = xf
    return spec


def normal_initial_states(spec: ProblemSpec, B: int, seed: int = 20240925, first: int = 0) -> np.ndarray:
    """x0 ~ N(0, I) per instance from the counter-based generator (Box-Muller on two streams)."""
    u1 = 0.5 * (uniform_batch(B, spec.n, seed, first) + 1.0)
    u2 = 0.5 * (uniform_batch(B, spec.n, seed + 1, first) + 1.0)
    u1 = np.clip(u1, 2.0 ** -53, 1.0)
    return np.ascontiguousarray(np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2))
