"""Host-side problem definition for the batched NMPC path (numpy only, no solver arithmetic).

Everything here mirrors what the reference's closed-loop drivers do *around* ``controller.solve``:

* :func:`load_reference_trajectory` / :func:`do_interpolation` -- simulation.py:201-218,446-449
  (``trajectory.json`` in the reference is a 0-byte file; the tracked trajectory is
  ``data/state_traj.txt`` + ``data/input_traj.txt``, shipped here under ``data/``).
* :func:`window` / :func:`windows_batch` -- the three window regimes of simulation.py:485-499.
* :func:`time_indices` -- the float-accumulated ``k = floor(t/dt)`` of simulation.py:484-485,560.
* :func:`pack_z` / :func:`unpack_z` -- decision-vector layout of trajectory_planning.py:38-84.
* :func:`make_scenarios` -- the synthetic batch of SURVEY.md section 8(d) config 2 (builder-defined:
  the reference has no batch driver; ``test_cases.json`` poses enter as SE(2) scenario families).

Arrays at this level use the *stage-major* layout of the C ABI: states ``[..., N+1, 6]``,
inputs ``[..., N, 2]``.  The reference's ``solve`` takes the transposes (``[6, N+1]``, ``[2, N]``);
the shim classes in ``mpc_control.py`` / ``mpc_control_nmpc.py`` convert.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

from .config import NU, NX, Config

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ----------------------------------------------------------------------------- trajectory
def do_interpolation(state_traj: np.ndarray, input_traj: np.ndarray, dt_1: float, dt_2: float):
    """Linear up-sampling of a ``[6, T+1]`` / ``[2, T]`` trajectory from dt_1 to dt_2 (dt_1 > dt_2).

    Same result as simulation.py:201-218: states interpolated linearly, inputs held, last state copied.
    """
    T = input_traj.shape[1]
    n = math.floor(dt_1 / dt_2)
    frac = (np.arange(n, dtype=np.float64) / n)[None, None, :]
    s0 = state_traj[:, :-1, None]
    s1 = state_traj[:, 1:, None]
    states = ((1.0 - frac) * s0 + frac * s1).reshape(state_traj.shape[0], T * n)
    states = np.concatenate([states, state_traj[:, -1:]], axis=1)
    inputs = np.repeat(input_traj, n, axis=1)
    return states, inputs


def load_reference_trajectory(dt_to: float = 0.1, dt: float = 0.05, data_dir: str | None = None):
    """Returns the tracked trajectory stage-major: ``S [T+1, 6]``, ``U [T, 2]`` (T = 400 at dt 0.05)."""
    d = data_dir or DATA_DIR
    S = np.loadtxt(os.path.join(d, "state_traj.txt"))
    U = np.loadtxt(os.path.join(d, "input_traj.txt"))
    S, U = do_interpolation(S, U, dt_to, dt)
    return np.ascontiguousarray(S.T), np.ascontiguousarray(U.T)


def window(S: np.ndarray, U: np.ndarray, k: int, N: int):
    """Reference window starting at trajectory index k (simulation.py:485-499).

    ``k + N <= T``: plain slices.  ``k < T < k + N``: states padded with the last state, inputs with
    the LAST input.  ``k >= T``: all states = last state, inputs = ZERO.
    """
    T = U.shape[0]
    xs = np.empty((N + 1, NX))
    us = np.empty((N, NU))
    if k >= T:
        xs[:] = S[T]
        us[:] = 0.0
        return xs, us
    idx = np.minimum(k + np.arange(N + 1), T)
    xs[:] = S[idx]
    idu = np.minimum(k + np.arange(N), T - 1)
    us[:] = U[idu]
    return xs, us


def windows_batch(S: np.ndarray, U: np.ndarray, k: np.ndarray, N: int):
    """Vectorised :func:`window` for an int array ``k[B]`` -> ``[B, N+1, 6]``, ``[B, N, 2]``."""
    T = U.shape[0]
    k = np.asarray(k, dtype=np.int64)
    idx = np.minimum(k[:, None] + np.arange(N + 1)[None, :], T)
    xs = S[idx]
    idu = np.minimum(k[:, None] + np.arange(N)[None, :], T - 1)
    us = U[idu].copy()
    past = k >= T
    if past.any():
        xs[past] = S[T]
        us[past] = 0.0
    return np.ascontiguousarray(xs), np.ascontiguousarray(us)


def time_indices(T_sim: float, dt: float) -> np.ndarray:
    """The sequence of ``k = floor(t/dt)`` the reference loop visits (simulation.py:484-485,560).

    ``t`` is accumulated in floating point (``t += dt``) and the loop runs ``while t <= T_sim``; for
    dt = 0.05 the index repeats/skips.  Replicated exactly.
    """
    ks = []
    t = 0.0
    while t <= T_sim:
        ks.append(math.floor(t / dt))
        t += dt
    return np.asarray(ks, dtype=np.int32)


# ----------------------------------------------------------------------------- layouts
def pack_z(states: np.ndarray, inputs: np.ndarray) -> np.ndarray:
    """``[..., N+1, 6]``, ``[..., N, 2]`` -> ``[..., 8N+6]`` in the order of trajectory_planning.py:38-60."""
    N = inputs.shape[-2]
    lead = states.shape[:-2]
    z = np.empty(lead + (8 * N + 6,), dtype=np.float64)
    body = z[..., : 8 * N].reshape(lead + (N, 8))
    body[..., :6] = states[..., :N, :]
    body[..., 6:] = inputs
    z[..., 8 * N :] = states[..., N, :]
    return z


def unpack_z(z: np.ndarray, N: int):
    """Inverse of :func:`pack_z` (trajectory_planning.py:62-84) -> ``[..., N+1, 6]``, ``[..., N, 2]``."""
    lead = z.shape[:-1]
    body = z[..., : 8 * N].reshape(lead + (N, 8))
    states = np.empty(lead + (N + 1, 6), dtype=np.float64)
    states[..., :N, :] = body[..., :6]
    states[..., N, :] = z[..., 8 * N :]
    inputs = np.ascontiguousarray(body[..., 6:])
    return states, inputs


def shift_warm_start(z: np.ndarray, N: int, reference_bug: bool = False) -> np.ndarray:
    """Host twin of ``TruckTrailerNMPC._shift_solution`` (mpc_control_nmpc.py:69-88).

    ``reference_bug=True`` reproduces the reference's slicing of the tail: the vector
    ends ``[x_{N-1}; u_{N-1}; x_N]`` so ``z[-8:-2]`` is ``(u_{N-1}, x_N[0:4])`` and ``z[-2:]`` is
    ``x_N[4:6]``.
    """
    z = np.asarray(z, dtype=np.float64)
    out = np.empty_like(z)
    out[..., : 8 * (N - 1)] = z[..., 8 : 8 * N]
    if reference_bug:
        out[..., 8 * (N - 1) : 8 * (N - 1) + 6] = z[..., -8:-2]
        out[..., 8 * (N - 1) + 6 : 8 * N] = z[..., -2:]
        out[..., 8 * N :] = z[..., -8:-2]
    else:
        out[..., 8 * (N - 1) : 8 * (N - 1) + 6] = z[..., 8 * N :]
        out[..., 8 * (N - 1) + 6 : 8 * N] = z[..., 8 * (N - 1) + 6 : 8 * N]
        out[..., 8 * N :] = z[..., 8 * N :]
    return out


# ----------------------------------------------------------------------------- objective (checker-free)
def objective(cfg: Config, states: np.ndarray, inputs: np.ndarray, ref_states: np.ndarray, ref_inputs: np.ndarray):
    """J of mpc_control.py:17-25 for ``[..., N+1, 6]`` / ``[..., N, 2]`` arrays (no 1/2, terminal weight Q)."""
    Q, R = cfg.Qm(), cfg.Rm()
    dx = states - ref_states
    du = inputs - ref_inputs
    return np.einsum("...ki,ij,...kj->...", dx, Q, dx) + np.einsum("...ki,ij,...kj->...", du, R, du)


def model_f(cfg: Config, q: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Continuous kinematics of truck_trailer_model.py:8-24, vectorised over leading axes."""
    th, psi, phi, v = q[..., 2], q[..., 3], q[..., 4], q[..., 5]
    t = np.tan(phi)
    f = np.empty_like(q)
    f[..., 0] = v * np.cos(th)
    f[..., 1] = v * np.sin(th)
    f[..., 2] = v * t / cfg.L1
    f[..., 3] = -v * t / cfg.L1 * (1.0 + cfg.M / cfg.L2 * np.cos(psi)) - v * np.sin(psi) / cfg.L2
    f[..., 4] = u[..., 1]
    f[..., 5] = u[..., 0]
    return f


def dynamics_defect(cfg: Config, states: np.ndarray, inputs: np.ndarray) -> np.ndarray:
    """``x_{k+1} - x_k - dt f(x_k,u_k)`` (trajectory_planning.py:31-32), shape ``[..., N, 6]``."""
    return states[..., 1:, :] - states[..., :-1, :] - cfg.dt * model_f(cfg, states[..., :-1, :], inputs)


# ----------------------------------------------------------------------------- scenarios
def load_test_cases(path: str | None = None):
    """The 7 named start/goal poses of the reference's ``test_cases.json`` (:8-119)."""
    with open(path or os.path.join(DATA_DIR, "test_cases.json")) as f:
        return json.load(f)["cases"]


def se2_transform(S: np.ndarray, pose_xy_theta) -> np.ndarray:
    """Rigidly move a state trajectory ``[T+1, 6]`` so that its first pose becomes ``(x, y, theta)``.

    The kinematics (truck_trailer_model.py:17-22) only see theta through cos/sin of the heading and
    are invariant under SE(2), so the moved trajectory stays dynamically exact with unchanged inputs.
    """
    x0, y0, th0 = S[0, 0], S[0, 1], S[0, 2]
    x1, y1, th1 = pose_xy_theta
    d = th1 - th0
    c, s = math.cos(d), math.sin(d)
    out = S.copy()
    dx, dy = S[:, 0] - x0, S[:, 1] - y0
    out[:, 0] = x1 + c * dx - s * dy
    out[:, 1] = y1 + s * dx + c * dy
    out[:, 2] = S[:, 2] + d
    return out


@dataclass
class ScenarioBatch:
    x_init: np.ndarray  # [B, 6]
    ref_states: np.ndarray  # [B, N+1, 6]
    ref_inputs: np.ndarray  # [B, N, 2]
    k_index: np.ndarray  # [B] int32 window start
    family: np.ndarray  # [B] int32 test-case family (-1 = untransformed trajectory)
    # the same windows in the compact (shared-trajectory) form: problem i tracks traj_states[traj_index[i]] from k_index[i]
    traj_index: np.ndarray | None = None  # [B] int32
    traj_states: np.ndarray | None = None  # [F, T+1, 6]
    traj_inputs: np.ndarray | None = None  # [F, T, 2]


SIGMA_NARROW = np.full(6, 0.02)  # the reference's own noise level (simulation.py:29)
SIGMA_WIDE = np.array([0.5, 0.5, 0.1, 0.1, 0.05, 0.5])


def make_scenarios(
    cfg: Config,
    B: int,
    seed: int = 20251018,
    sigma: np.ndarray | float = SIGMA_NARROW,
    families: bool = True,
    S: np.ndarray | None = None,
    U: np.ndarray | None = None,
) -> ScenarioBatch:
    """Synthetic batch of SURVEY.md section 8(d) config 2.

    Scenario i: family ``i mod 7`` of test_cases.json (the trajectory is moved by SE(2) so that its
    first pose is the case's start pose with the reference's ``+pi/2`` heading convention,
    get_initial_goal_states.py:13; a family whose rotated heading leaves [-pi, pi] keeps the
    untransformed trajectory), window index ``k_i ~ U{0..T}``, ``x_init = S[k_i] + N(0, sigma^2)`` with
    psi, phi, v, theta clipped into ``[lb + 1e-3, ub - 1e-3]`` so that every NLP is feasible at stage 0.
    """
    N = cfg.horizon
    if S is None or U is None:
        S, U = load_reference_trajectory(dt=cfg.dt)
    T = U.shape[0]
    rng = np.random.default_rng(seed)
    k = rng.integers(0, T + 1, size=B).astype(np.int32)
    noise = rng.normal(0.0, 1.0, size=(B, 6)) * np.asarray(sigma, dtype=np.float64)
    fam = np.full(B, -1, dtype=np.int32)
    tix = np.zeros(B, dtype=np.int32)
    stack = [S]  # trajectory 0: untransformed
    xs = np.empty((B, N + 1, 6))
    us = np.empty((B, N, 2))
    lb = np.array(cfg.x_lb[:])
    ub = np.array(cfg.x_ub[:])
    if families:
        cases = load_test_cases()
        trajs = []
        for c in cases:
            st = c["start"]
            Sm = se2_transform(S, (st["x"], st["y"], st["heading_rad"] + math.pi / 2.0))
            ok = (Sm[:, 2].min() >= lb[2] + 1e-2) and (Sm[:, 2].max() <= ub[2] - 1e-2)
            trajs.append(Sm if ok else None)
        for f, Sm in enumerate(trajs):
            sel = np.nonzero(np.arange(B) % len(cases) == f)[0]
            if sel.size == 0:
                continue
            src = S if Sm is None else Sm
            xs[sel], us[sel] = windows_batch(src, U, k[sel], N)
            fam[sel] = -1 if Sm is None else f
            if Sm is not None:
                tix[sel] = len(stack)
                stack.append(Sm)
    else:
        xs, us = windows_batch(S, U, k, N)
    x_init = xs[:, 0, :] + noise
    for i in range(6):
        if math.isfinite(lb[i]) or math.isfinite(ub[i]):
            x_init[:, i] = np.clip(x_init[:, i], lb[i] + 1e-3, ub[i] - 1e-3)
    return ScenarioBatch(np.ascontiguousarray(x_init), xs, us, k, fam, tix, np.ascontiguousarray(np.stack(stack)),
                         np.ascontiguousarray(np.broadcast_to(U, (len(stack),) + U.shape)))
