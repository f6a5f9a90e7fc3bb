"""Collision test of a predicted horizon against the axis-aligned obstacle rectangles -- the caller-side check that
``simulation.py`` uses to pick its controller for the next solve (python-files/simulation.py:222-385:
``check_obb_aabb_collision`` / ``check_state_collision`` / ``check_trajectory_collision``).

Same decision as the reference (separating-axis test between each body rectangle and each obstacle; a contact without
gap counts as a collision because the reference's gap test is strict, simulation.py:298), evaluated for all stages,
both bodies and all obstacles at once in numpy instead of three nested Python loops.  Body geometry as in
``get_vehicle_center_np`` / ``get_trailer_center_np`` (simulation.py:305-317): the vehicle rectangle (L1 x W1) is
centred half a wheelbase ahead of the rear axle, the trailer rectangle (L2 x W2) half a trailer length behind the hitch,
which sits M behind the axle.
"""
from __future__ import annotations

import numpy as np


def _obstacle_array(obstacle_list) -> np.ndarray:
    """[n_obs, 4] = (cx, cy, half width, half height) from the reference's list of dicts (get_obstacles.py:5-33)."""
    if len(obstacle_list) == 0:
        return np.zeros((0, 4))
    return np.array([[o["center"][0], o["center"][1], 0.5 * o["width"], 0.5 * o["height"]] for o in obstacle_list],
                    dtype=np.float64)


def _body_frames(states: np.ndarray, params: dict):
    """states [6 or 4.., K] (column k = stage k, the layout ``solve`` returns) ->
    centres [K, 2, 2], unit length-axes [K, 2, 2], half extents [2, 2] for (vehicle, trailer)."""
    x, y, th, psi = (np.asarray(states[i], dtype=np.float64) for i in range(4))
    L1, L2, M = float(params["L1"]), float(params["L2"]), float(params["M"])
    al = th + psi
    cv = np.stack([x + np.cos(th) * (0.5 * L1), y + np.sin(th) * (0.5 * L1)], axis=-1)
    ct = np.stack([x - np.cos(th) * M - np.cos(al) * (0.5 * L2), y - np.sin(th) * M - np.sin(al) * (0.5 * L2)], axis=-1)
    centres = np.stack([cv, ct], axis=1)
    axes = np.stack([np.stack([np.cos(th), np.sin(th)], axis=-1), np.stack([np.cos(al), np.sin(al)], axis=-1)], axis=1)
    half = np.array([[0.5 * L1, 0.5 * float(params["W1"])], [0.5 * L2, 0.5 * float(params["W2"])]])
    return centres, axes, half


def stage_collisions(states, params: dict, obstacle_list) -> np.ndarray:
    """bool [K]: stage k's vehicle or trailer rectangle overlaps (or touches) some obstacle."""
    states = np.asarray(states, dtype=np.float64)
    if states.ndim == 1:
        states = states[:, None]
    K = states.shape[1]
    obs = _obstacle_array(obstacle_list)
    if obs.shape[0] == 0:
        return np.zeros(K, dtype=bool)
    c, u, half = _body_frames(states, params)             # [K,2,2], [K,2,2], [2,2]
    v = np.stack([-u[..., 1], u[..., 0]], axis=-1)        # width axes
    oc, oh = obs[:, :2], obs[:, 2:]                       # [O,2], [O,2]
    # the two obstacle axes (x, y): body extent along a world axis = |u_i| * hl + |v_i| * hw
    ext = np.abs(u) * half[None, :, 0:1] + np.abs(v) * half[None, :, 1:2]          # [K,2,2] (per world axis)
    gap_w = np.abs(c[:, :, None, :] - oc[None, None, :, :]) - ext[:, :, None, :] - oh[None, None, :, :]   # [K,2,O,2]
    hit = (gap_w <= 0.0).all(-1)
    # the two body axes: obstacle extent along a unit axis a = |a_x| * hw + |a_y| * hh
    d = oc[None, None, :, :] - c[:, :, None, :]                                     # [K,2,O,2]
    for ax, h_body in ((u, half[:, 0]), (v, half[:, 1])):
        proj = np.abs((d * ax[:, :, None, :]).sum(-1))                              # centre distance along the axis
        oext = (np.abs(ax)[:, :, None, :] * oh[None, None, :, :]).sum(-1)
        hit &= (proj - oext - h_body[None, :, None]) <= 0.0
    return hit.any(axis=(1, 2))


def check_state_collision(state, params: dict, obstacle_list) -> bool:
    """simulation.py:319-361 for one state ``[x, y, theta, psi, ...]``."""
    return bool(stage_collisions(np.asarray(state, dtype=np.float64).reshape(-1)[:4, None], params, obstacle_list)[0])


def check_trajectory_collision(states, params: dict, obstacle_list) -> bool:
    """simulation.py:363-385: does any column of ``states [num_state, horizon+1]`` collide?"""
    if len(obstacle_list) == 0:
        return False
    return bool(stage_collisions(states, params, obstacle_list).any())
