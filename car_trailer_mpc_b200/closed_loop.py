"""Headless closed-loop drivers: the reference's `while t <= T_sim` loops without matplotlib.

* :func:`simulate_single` -- one vehicle, any controller object with the reference's ``solve`` interface
  (the shims of this package, or a test double).  Mirrors simulation.py:484-560 / simulation_nmpc.py:192-255:
  float-accumulated time index, the three window regimes (D.1), RNG draw order (D.3), plant
  ``update`` with the disturbance model (D.5), failure policy of the NMPC driver, end-of-run metrics (D.6).
* :func:`simulate_batch` -- B vehicles at once on one GPU through :class:`BatchSolver` (shared-trajectory windows
  computed in-kernel, plant step kernel, state never leaves the device); per-scenario noise comes from a
  counter-based generator keyed on (seed, step, global scenario id) so results do not depend on how scenarios
  are sharded over GPUs.  This is the building block of the Monte-Carlo configuration.

"Jackknife" is builder-defined (the reference has no detector): ``|psi_t| > pi/3 + 1e-6`` at any step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import problem as pb
from .config import ST_ACCEPTABLE, ST_CONVERGED, Config

JACKKNIFE_LIMIT = math.pi / 3.0 + 1e-6

# DISTURBANCE_PARAMS of simulation.py:26-32
DEFAULT_DISTURBANCE = {
    "friction_coeff": 0.9,
    "slippage_coeff": 0.9,
    "process_noise_std": 0.02,
    "lateral_slip_gain": 0.01,
    "slip_angle_max": 0.0,
}


# ----------------------------------------------------------------------------- plant (host twin of the kernel)
def f_dyn(q: np.ndarray, u: np.ndarray, params: dict) -> np.ndarray:
    """simulation.py:34-48."""
    L1, L2, M = params["L1"], params["L2"], params["M"]
    _, _, theta, psi, phi, v = q
    a, omega = u
    return np.array([
        v * math.cos(theta), v * math.sin(theta), v * math.tan(phi) / L1,
        -v * math.tan(phi) / L1 * (1 + M / L2 * math.cos(psi)) - v * math.sin(psi) / L2, omega, a])


def plant_update(q, u, params, disturbance=None, state_noise=None, noise_scale=0.0):
    """``update`` of simulation.py:167-199; with ``state_noise``/``noise_scale = dt`` the variant of
    simulation_nmpc.py:94-105 (``q_ += state_noise * dt``)."""
    u = np.asarray(u, dtype=np.float64).copy()
    q = np.asarray(q, dtype=np.float64)
    if disturbance is not None:
        u[0] *= disturbance.get("friction_coeff", 1.0)
        u[1] *= disturbance.get("slippage_coeff", 1.0)
    qd = f_dyn(q, u, params)
    if disturbance is not None and "slip_angle_max" in disturbance:
        slip = 1.0 - min(abs(q[4]) * abs(q[5]) * disturbance["slip_angle_max"], 0.3)
        qd[2] *= slip
        qd[3] *= slip
    qn = q + qd * params["dt"]
    if state_noise is not None:
        qn = qn + np.asarray(state_noise) * noise_scale
    if disturbance is not None and "lateral_slip_gain" in disturbance:
        mag = disturbance["lateral_slip_gain"] * abs(q[5]) * abs(q[4])
        qn[0] += mag * math.cos(q[2] + math.pi / 2) * params["dt"]
        qn[1] += mag * math.sin(q[2] + math.pi / 2) * params["dt"]
    return qn


def wrap_angle(a):
    """simulation.py:578-580."""
    return (a + np.pi) % (2 * np.pi) - np.pi


# ----------------------------------------------------------------------------- LQR score (LQR_cost.py)
def euler_jacobians(cfg_or_params, x, u):
    """A = d(x + dt f)/dx, B = d(x + dt f)/du at (x,u): what ``ca.jacobian`` builds in LQR_cost.py:13-27."""
    g = (lambda k: getattr(cfg_or_params, k)) if isinstance(cfg_or_params, Config) else (lambda k: cfg_or_params[k])
    L1, L2, M, dt = g("L1"), g("L2"), g("M"), g("dt")
    th, psi, phi, v = x[2], x[3], x[4], x[5]
    t = math.tan(phi)
    s = 1 + t * t
    c = M / L2
    F = np.zeros((6, 6))
    F[0, 2], F[0, 5] = -v * math.sin(th), math.cos(th)
    F[1, 2], F[1, 5] = v * math.cos(th), math.sin(th)
    F[2, 4], F[2, 5] = v * s / L1, t / L1
    F[3, 3] = v * t / L1 * c * math.sin(psi) - v * math.cos(psi) / L2
    F[3, 4] = -v * s / L1 * (1 + c * math.cos(psi))
    F[3, 5] = -t / L1 * (1 + c * math.cos(psi)) - math.sin(psi) / L2
    A = np.eye(6) + dt * F
    B = np.zeros((6, 2))
    B[4, 1] = dt
    B[5, 0] = dt
    return A, B


def lqr_riccati(cfg_or_params, Q, R, x_goal, u_goal):
    """LQR_cost.py:7-34: DARE at the goal linearisation (shared by all scenarios of a trajectory)."""
    from scipy.linalg import solve_discrete_are

    A, B = euler_jacobians(cfg_or_params, np.asarray(x_goal, float), np.asarray(u_goal, float))
    P = solve_discrete_are(A, B, np.asarray(Q, float), np.asarray(R, float))
    return 0.5 * (P + P.T)


def lqr_distance(x, x_goal, P):
    """LQR_cost.py:37-41, vectorised over leading axes."""
    dx = np.asarray(x) - np.asarray(x_goal)
    return np.einsum("...i,ij,...j->...", dx, P, dx)


# ----------------------------------------------------------------------------- single vehicle
@dataclass
class EpisodeResult:
    states: np.ndarray  # [steps+1, 6] closed-loop trajectory
    controls: np.ndarray  # [steps, 2]
    k_index: np.ndarray
    failures: int
    iterations: list = field(default_factory=list)
    aborted: bool = False

    def metrics(self, goal: np.ndarray) -> dict:
        x = self.states
        fin = x[-1]
        return {
            "distance_error": float(np.hypot(fin[0] - goal[0], fin[1] - goal[1])),
            "heading_error": float(wrap_angle(fin[2] - goal[2])),
            "hitch_error": float(wrap_angle(fin[3] - goal[3])),
            "max_abs_psi": float(np.abs(x[:, 3]).max()),
            "max_abs_phi": float(np.abs(x[:, 4]).max()),
            "jackknife": bool((np.abs(x[:, 3]) > JACKKNIFE_LIMIT).any()),
            "failures": int(self.failures),
            "steps": int(len(self.controls)),
        }


def simulate_single(controller, S, U, x0, T_sim: float, dt: float, horizon: int, params: dict,
                    disturbance: dict | None = None, rng: np.random.Generator | np.random.RandomState | None = None,
                    variant: str = "tracking") -> EpisodeResult:
    """One closed-loop episode.  ``variant='tracking'``: simulation.py (measurement noise to the controller, plant
    noise drawn and discarded); ``variant='nmpc'``: simulation_nmpc.py (no measurement noise, ``noise*dt`` added to the
    plant, zero control on failure, abort after >20 consecutive failures)."""
    N = horizon
    state = np.asarray(x0, dtype=np.float64).copy()
    xs_hist, us_hist = [state.copy()], []
    ks = pb.time_indices(T_sim, dt)
    ref_s = np.zeros((6, N + 1))  # reused buffers mutated in place, as in simulation.py:463-464
    ref_u = np.zeros((2, N))
    failures = consecutive = 0
    iters = []
    aborted = False
    std = disturbance.get("process_noise_std", 0.0) if disturbance else 0.0
    for k in ks:
        xs, us = pb.window(S, U, int(k), N)
        ref_s[:, :] = xs.T
        ref_u[:, :] = us.T
        meas = state
        if variant == "tracking" and disturbance is not None:
            meas = state + rng.normal(0, std, 6)  # generate_measurement_noise, simulation.py:514
        states, inputs = controller.solve(meas, ref_s, ref_u)
        if getattr(controller, "last_iterations", None) is not None:
            iters.append(controller.last_iterations)
        if states is None or inputs is None:  # simulation_nmpc.py:207-216
            failures += 1
            consecutive += 1
            u_con = np.zeros(2)
            if consecutive > 20:
                aborted = True
                us_hist.append(u_con)
                break
        else:
            if getattr(controller, "last_status", 0) not in (ST_CONVERGED, ST_ACCEPTABLE):
                failures += 1
            consecutive = 0
            u_con = np.array(inputs[:, 0], dtype=np.float64)
        if disturbance is not None:
            noise = rng.normal(0, std, 6)  # apply_disturbances draws 6 normals (simulation.py:85)
            if variant == "nmpc":
                state = plant_update(state, u_con, params, disturbance, noise, params["dt"])
            else:
                state = plant_update(state, u_con, params, disturbance)  # the draw is discarded (simulation.py:181)
        else:
            state = plant_update(state, u_con, params)
        xs_hist.append(state.copy())
        us_hist.append(u_con)
    return EpisodeResult(np.array(xs_hist), np.array(us_hist), ks, failures, iters, aborted)


# ----------------------------------------------------------------------------- batched, on device
def _s64(c: int) -> int:
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(z, s: int):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix64(z):
    """splitmix64 finaliser on int64 tensors (multiplications wrap modulo 2^64)."""
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def counter_normal(seed: int, step: int, ids, n: int):
    """Counter-based standard normals ``[len(ids), n]`` keyed on (seed, step, scenario id, component): a splitmix64
    hash -> two uniforms -> Box-Muller, all in torch integer/float ops on the tensor's device.  Independent of batch
    composition, so any sharding of scenarios over GPUs reproduces the same noise."""
    import torch

    dev = ids.device
    comp = torch.arange(n, device=dev, dtype=torch.int64)[None, :]
    key = _s64(int(seed) * 0x2545F4914F6CDD1D + int(step) * 0x1B03738712FAD5C9)
    x = ids.to(torch.int64)[:, None] * _s64(0x9E3779B97F4A7C15) + comp * _s64(0x632BE59BD9B4E019) + key
    a = _mix64(x)
    b = _mix64(a ^ _s64(0x5851F42D4C957F2D))
    u1 = (_lsr(a, 11).to(torch.float64) + 0.5) * (1.0 / (1 << 53))
    u2 = (_lsr(b, 11).to(torch.float64) + 0.5) * (1.0 / (1 << 53))
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def simulate_batch(solver, S, U, x0, T_sim: float, dt: float, disturbance: dict | None = None, seed: int = 0,
                   scenario_ids=None, variant: str = "tracking", warm_start: bool = False, record_every: int = 0):
    """B closed-loop episodes on the solver's GPU.  ``x0`` ``[B,6]`` torch CUDA tensor (float64).  Returns a dict of
    per-scenario metric tensors (on device) and, if ``record_every > 0``, sub-sampled state history."""
    import torch

    cfg = solver.cfg
    N = cfg.horizon
    dev = x0.device
    B = x0.shape[0]
    S_d = torch.as_tensor(S, dtype=torch.float64, device=dev).contiguous()
    U_d = torch.as_tensor(U, dtype=torch.float64, device=dev).contiguous()
    ids = torch.arange(B, device=dev, dtype=torch.int64) if scenario_ids is None else scenario_ids.to(dev)
    ks = pb.time_indices(T_sim, dt)
    T = U_d.shape[0]
    state = x0.clone()
    max_psi = state[:, 3].abs().clone()
    max_phi = state[:, 4].abs().clone()
    fails = torch.zeros(B, dtype=torch.int32, device=dev)
    iters_sum = torch.zeros(B, dtype=torch.int64, device=dev)
    iters_max = torch.zeros(B, dtype=torch.int32, device=dev)
    sq_err = torch.zeros(B, dtype=torch.float64, device=dev)
    std = disturbance.get("process_noise_std", 0.0) if disturbance else 0.0
    kidx = torch.empty(B, dtype=torch.int32, device=dev)
    z_prev = None
    hist = []
    consec = torch.zeros(B, dtype=torch.int32, device=dev)   # consecutive failed solves (simulation_nmpc.py:207-216)
    stopped = torch.zeros(B, dtype=torch.bool, device=dev)
    for step, k in enumerate(ks):
        kidx.fill_(int(k))
        meas = state
        if variant == "tracking" and disturbance is not None and std > 0:
            meas = state + std * counter_normal(seed, 2 * step, ids, 6)
        zw = None
        if warm_start and z_prev is not None:
            zw = solver.shift_warm_start(z_prev, reference_bug=False)
        r = solver.solve_shared(meas, kidx, S_d, U_d, z_warm=zw, want_z=warm_start)
        ok = r["status"] <= ST_ACCEPTABLE
        u = r["u0"]
        if variant == "nmpc":
            u = torch.where(ok[:, None], u, torch.zeros_like(u))  # zero control on failure (simulation_nmpc.py:211)
        if warm_start:
            z_prev = r["z"] if z_prev is None else torch.where(ok[:, None], r["z"], z_prev)
        noise = None
        scale = 0.0
        if variant == "nmpc" and disturbance is not None and std > 0:
            noise = (std * counter_normal(seed, 2 * step + 1, ids, 6)).contiguous()
            scale = dt
        new_state = solver.plant_step(state, u.contiguous(), disturbance, noise, scale)
        live = ~stopped   # a stopped run ("Too many consecutive NMPC failures", simulation_nmpc.py:212-216) stays where it is
        state = torch.where(live[:, None], new_state, state)
        fails += (live & ~ok).to(torch.int32)
        iters_sum += torch.where(live, r["iters"], torch.zeros_like(r["iters"])).to(torch.int64)
        iters_max = torch.maximum(iters_max, r["iters"])
        max_psi = torch.maximum(max_psi, state[:, 3].abs())
        max_phi = torch.maximum(max_phi, state[:, 4].abs())
        ref_next = S_d[min(int(k) + 1, T)]
        sq_err += torch.where(live, ((state[:, :2] - ref_next[:2]) ** 2).sum(1), torch.zeros_like(sq_err))
        if variant == "nmpc":
            consec = torch.where(ok, torch.zeros_like(consec), consec + live.to(torch.int32))
            stopped = stopped | (consec > 20)
        if record_every and step % record_every == 0:
            hist.append(state.clone())
    goal = S_d[T]
    two_pi = 2 * math.pi
    out = {
        "final_state": state,
        "distance_error": torch.hypot(state[:, 0] - goal[0], state[:, 1] - goal[1]),
        "heading_error": torch.remainder(state[:, 2] - goal[2] + math.pi, two_pi) - math.pi,
        "hitch_error": torch.remainder(state[:, 3] - goal[3] + math.pi, two_pi) - math.pi,
        "max_abs_psi": max_psi,
        "max_abs_phi": max_phi,
        "jackknife": max_psi > JACKKNIFE_LIMIT,
        "failures": fails,
        "mean_iters": iters_sum.to(torch.float64) / len(ks),
        "max_iters": iters_max,
        "rms_tracking_error": torch.sqrt(sq_err / len(ks)),
        "steps": len(ks),
    }
    if hist:
        out["history"] = torch.stack(hist)
    return out
