"""`BatchSolver`: Python front end of the C ABI (include/ttmpc.h) -- the new batched driver surface.

One instance = one CUDA device.  Inputs may be

* **torch CUDA tensors** (float64, contiguous): zero-copy, pointers go straight to the kernels and the call
  is asynchronous on the current torch stream; outputs are torch tensors on the same device;
* **numpy arrays**: the library stages them through its own device buffer (host-pointer mode, the path the
  B=1 shim classes use); outputs are numpy arrays; the call is synchronous.

Array layouts are the reference's (stage-major): ``x_init [B,6]``, ``ref_states [B,N+1,6]``,
``ref_inputs [B,N,2]``, ``z [B,8N+6]`` (trajectory_planning.py:38-60).  There is no CPU compute path here.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .config import FLAG_ASYNC_HOST, FLAG_HOST_POINTERS, Config, Obstacles


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class BatchSolver:
    def __init__(self, cfg: Config, device: int = 0):
        self.cfg = cfg.copy()
        self.device = int(device)
        self._L = _lib.load()
        self._h = {}  # flags -> handle
        self._handle(0)  # fail early and loudly when no CUDA device is usable

    # ------------------------------------------------------------------ handles
    def _handle(self, flags: int):
        h = self._h.get(flags)
        if h is None:
            c = self.cfg.copy()
            c.flags = flags
            h = ctypes.c_void_p()
            rc = self._L.ttmpc_create(ctypes.byref(c), self.device, ctypes.byref(h))
            if rc == _lib.E_NODEV:
                raise _lib.TTMPCError(
                    f"ttmpc_create: no usable CUDA device {self.device} (rc={rc}); this solver has no CPU fallback"
                )
            if rc != 0:
                raise _lib.TTMPCError(f"ttmpc_create failed (rc={rc}): invalid configuration")
            self._h[flags] = h
        return h

    def close(self):
        for h in self._h.values():
            self._L.ttmpc_destroy(h)
        self._h = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, h, rc: int, what: str):
        if rc != 0:
            msg = self._L.ttmpc_last_error(h)
            raise _lib.TTMPCError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")

    # ------------------------------------------------------------------ accounting
    def launch_count(self) -> int:
        return int(sum(self._L.ttmpc_launch_count(h) for h in self._h.values()))

    def last_solve_lanes(self, host: bool = False) -> int:
        """Kernel flavour of the last plain solve: 8/16/32 = warp-cooperative team kernel with that many lanes per
        problem, 0 = lane-per-problem kernel (include/ttmpc.h ttmpc_last_solve_lanes)."""
        h = self._h.get(FLAG_HOST_POINTERS if host else 0)
        return int(self._L.ttmpc_last_solve_lanes(h)) if h is not None else 0

    def kernel_launches(self) -> dict:
        out = {}
        for h in self._h.values():
            i = 0
            while True:
                n = ctypes.c_int64()
                name = self._L.ttmpc_kernel_name(h, i, ctypes.byref(n))
                if not name:
                    break
                out[name.decode()] = out.get(name.decode(), 0) + n.value
                i += 1
        return out

    def measure_fp64_peak(self) -> float:
        """Non-tensor FP64 FMA peak of this device in GFLOP/s (roofline denominator)."""
        v = self._L.ttmpc_measure_fp64_peak(self._handle(0), None)
        if v < 0:
            raise _lib.TTMPCError(f"ttmpc_measure_fp64_peak failed ({v})")
        return float(v)

    # ------------------------------------------------------------------ solve
    def solve(self, x_init, ref_states, ref_inputs, z_warm=None, want_z: bool = True, stream=None,
              q_weights=None, r_weights=None, host_async: bool = False, out=None) -> dict:
        """Solve B problems with per-problem reference windows (MPCTrackingControl.solve semantics).

        ``q_weights [B,6]`` / ``r_weights [B,2]`` (optional, together): per-problem weight scalings
        ``Q_w = diag(q) Q diag(q)``, ``R_w = diag(r) R diag(r)`` as in mpc_control_fuzzy.py:21-31 (diagonal Q, R only).

        Host arrays (numpy, or torch CPU tensors -- page-locked ones make the copies asynchronous) go through the
        library's copy-in | solve | copy-out pipeline.  ``host_async=True`` returns as soon as the work is queued (up to
        three solves in flight); the outputs are valid after :meth:`sync`.  ``out``: dict of preallocated output arrays
        to reuse (same keys as the result)."""
        return self._solve(x_init, ref_states, ref_inputs, None, None, None, z_warm, want_z, stream, q_weights, r_weights,
                           host_async=host_async, out=out)

    def solve_shared(self, x_init, k_index, traj_states, traj_inputs, z_warm=None, want_z: bool = True, stream=None,
                     traj_index=None, host_async: bool = False, out=None, want_kkt: bool = True) -> dict:
        """All problems track one trajectory ``traj_states [T+1,6]``, ``traj_inputs [T,2]``; problem i uses the
        window starting at ``k_index[i]`` with the padding rules of simulation.py:485-499.

        With ``traj_index [B]`` (int32) the trajectories are a stack ``traj_states [F,T+1,6]``, ``traj_inputs [F,T,2]``
        and problem i tracks trajectory ``traj_index[i]`` (the batch driver's scenario families).  ``want_z=False,
        want_kkt=False`` is the compact contract: 56 B in, 24 B out per problem."""
        return self._solve(x_init, None, None, k_index, traj_states, traj_inputs, z_warm, want_z, stream,
                           traj_index=traj_index, host_async=host_async, out=out, want_kkt=want_kkt)

    def solve_obca(self, obstacles, x_init, ref_states, ref_inputs, want_z: bool = True, stream=None) -> dict:
        """Solve B obstacle-aware problems (``MPCTrackingControlObs.solve``, mpc_control_obs.py:282-322).

        ``obstacles``: :class:`config.Obstacles` or the reference's obstacle list (dicts with center/width/height,
        get_obstacles.py:5-33).  Always a cold start (mpc_control_obs.py:216-239); ``z`` holds states and inputs."""
        return self._solve(x_init, ref_states, ref_inputs, None, None, None, None, want_z, stream, obstacles=obstacles)

    def solve_obca_shared(self, obstacles, x_init, k_index, traj_states, traj_inputs, want_z: bool = True, stream=None) -> dict:
        """:meth:`solve_obca` with the windows taken from one shared trajectory (simulation.py:485-499)."""
        return self._solve(x_init, None, None, k_index, traj_states, traj_inputs, None, want_z, stream, obstacles=obstacles)

    def plan(self, obstacles, x_init, goal, terminal_weight: float = 100.0, terminal_box: float = 1e-2, z_guess=None,
             want_z: bool = True, stream=None) -> dict:
        """The offline planner's NLP (``TrajectoryOptimization.plan``, trajectory_optimization.py:311-331) for B start
        states ``x_init [B,6]`` towards one ``goal [6]``: goal-tracking cost with terminal weight ``terminal_weight * Q``
        (:175-183), ``|x_N - goal| <= terminal_box`` (:168-173), all collision rows; ``z_guess [B,8N+6]``: the initial
        trajectory (states and inputs in the z layout; None = every state at the goal)."""
        return self._solve(x_init, None, None, None, None, None, z_guess, want_z, stream, obstacles=obstacles,
                           plan=(np.ascontiguousarray(goal, dtype=np.float64).reshape(6), float(terminal_weight), float(terminal_box)))

    def sync(self) -> float:
        """Wait for every ``host_async`` solve queued so far (``ttmpc_sync``); returns the device-side duration of that
        burst in ms (first copy-in to last copy-out, CUDA events on the library's copy streams)."""
        h = self._h.get(FLAG_HOST_POINTERS | FLAG_ASYNC_HOST)
        if h is None:
            return 0.0
        self._check(h, self._L.ttmpc_sync(h), "ttmpc_sync")
        return float(self._L.ttmpc_host_pipeline_ms(h))

    def _solve(self, x_init, ref_states, ref_inputs, k_index, traj_states, traj_inputs, z_warm, want_z, stream,
               q_weights=None, r_weights=None, obstacles=None, traj_index=None, host_async=False, out=None, want_kkt=True,
               plan=None):
        if obstacles is not None and not isinstance(obstacles, Obstacles):
            obstacles = Obstacles.from_list(obstacles)
        if (q_weights is None) != (r_weights is None):
            raise ValueError("q_weights and r_weights go together")
        N = self.cfg.horizon
        nz = 8 * N + 6
        shared = ref_states is None and plan is None
        if plan is not None and obstacles is None:
            raise ValueError("the planner needs an obstacle set")
        if shared and (k_index is None or traj_states is None or traj_inputs is None):
            raise ValueError("shared-trajectory mode needs k_index, traj_states and traj_inputs")
        if traj_index is not None and (not shared or obstacles is not None or q_weights is not None):
            raise ValueError("traj_index is only available for the plain shared-trajectory solve")
        on_device = _is_torch(x_init) and x_init.device.type == "cuda"
        keep = []
        if on_device:
            import torch

            dev = x_init.device
            if dev.index != self.device:
                raise ValueError(f"tensors must live on the solver's device cuda:{self.device}, got {dev}")
            f64, i32 = torch.float64, torch.int32

            def arg(t, shape, dtype=f64, name="array"):
                if not (_is_torch(t) and t.dtype == dtype and t.is_contiguous() and t.device == dev and tuple(t.shape) == tuple(shape)):
                    raise ValueError(f"{name}: expected a contiguous {dtype} tensor of shape {tuple(shape)} on {dev}, got "
                                     f"{getattr(t, 'dtype', type(t))} {tuple(getattr(t, 'shape', ()))} on {getattr(t, 'device', '?')}")
                return t.data_ptr()

            def new(shape, dtype=f64):
                return torch.empty(shape, dtype=dtype, device=dev)

            if stream is None:
                stream = torch.cuda.current_stream(dev).cuda_stream
            h = self._handle(0)
        else:
            f64, i32 = np.float64, np.int32

            def arg(t, shape, dtype=f64, name="array"):
                if _is_torch(t):  # torch CPU tensor (page-locked or not): used in place
                    import torch

                    want = torch.float64 if dtype is np.float64 else torch.int32
                    if not (t.device.type == "cpu" and t.dtype == want and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
                        raise ValueError(f"{name}: expected a contiguous CPU {dtype.__name__} tensor of shape {tuple(shape)}")
                    keep.append(t)
                    return t.data_ptr()
                a = np.ascontiguousarray(t, dtype=dtype)
                if a.size != int(np.prod(shape)):
                    raise ValueError(f"{name}: expected {int(np.prod(shape))} elements (shape {tuple(shape)}), got {a.shape}")
                keep.append(a)
                return a.ctypes.data

            def new(shape, dtype=f64):
                return np.empty(shape, dtype=dtype)

            stream = None
            h = self._handle(FLAG_HOST_POINTERS | (FLAG_ASYNC_HOST if host_async else 0))
        B = int(x_init.shape[0]) if hasattr(x_init, "shape") and len(x_init.shape) == 2 else int(np.asarray(x_init).size // 6)
        px = arg(x_init, (B, 6), name="x_init")
        F = 1
        if shared:
            ts_shape = tuple(traj_states.shape) if hasattr(traj_states, "shape") else np.asarray(traj_states).shape
            if traj_index is not None:
                if len(ts_shape) != 3:
                    raise ValueError("with traj_index the trajectories are stacks traj_states [F,T+1,6], traj_inputs [F,T,2]")
                F, T = int(ts_shape[0]), int(ts_shape[1]) - 1
                pts, ptu = arg(traj_states, (F, T + 1, 6), name="traj_states"), arg(traj_inputs, (F, T, 2), name="traj_inputs")
                pti = arg(traj_index, (B,), i32, name="traj_index")
            else:
                T = int(ts_shape[0]) - 1
                pts, ptu = arg(traj_states, (T + 1, 6), name="traj_states"), arg(traj_inputs, (T, 2), name="traj_inputs")
            pk = arg(k_index, (B,), i32, name="k_index")
        elif plan is None:
            prs, pru = arg(ref_states, (B, N + 1, 6), name="ref_states"), arg(ref_inputs, (B, N, 2), name="ref_inputs")
        pzw = arg(z_warm, (B, nz), name="z_warm") if z_warm is not None else None
        pqw = arg(q_weights, (B, 6), name="q_weights") if q_weights is not None else None
        prw = arg(r_weights, (B, 2), name="r_weights") if r_weights is not None else None
        shapes = dict(z=((B, nz), f64), u0=((B, 2), f64), obj=((B,), f64), kkt=((B, 3), f64), iters=((B,), i32), status=((B,), i32))
        res, ptrs = {}, {}
        for key, (shape, dtype) in shapes.items():
            if (key == "z" and not want_z) or (key in ("kkt", "obj") and not want_kkt):
                res[key], ptrs[key] = None, None
                continue
            buf = out.get(key) if out is not None else None
            if buf is None:
                buf = new(shape, dtype)
            ptrs[key] = arg(buf, shape, dtype, name=f"out[{key}]")
            res[key] = buf
        o = [ptrs[k] for k in ("z", "u0", "obj", "kkt", "iters", "status")]
        if plan is not None:
            keep.append(plan[0])
            rc = self._L.ttmpc_plan_batch(h, ctypes.byref(obstacles), B, px, plan[0].ctypes.data, plan[1], plan[2], pzw, *o, stream)
        elif obstacles is not None and shared:
            rc = self._L.ttmpc_obca_solve_batch_shared(h, ctypes.byref(obstacles), B, px, pk, pts, ptu, T, *o, stream)
        elif obstacles is not None:
            rc = self._L.ttmpc_obca_solve_batch(h, ctypes.byref(obstacles), B, px, prs, pru, *o, stream)
        elif shared and traj_index is not None:
            rc = self._L.ttmpc_solve_batch_multi(h, B, px, pk, pti, pts, ptu, F, T, pzw, *o, stream)
        elif shared:
            rc = self._L.ttmpc_solve_batch_shared(h, B, px, pk, pts, ptu, T, pzw, *o, stream)
        elif pqw is not None:
            rc = self._L.ttmpc_solve_batch_weighted(h, B, px, prs, pru, pqw, prw, pzw, *o, stream)
        else:
            rc = self._L.ttmpc_solve_batch(h, B, px, prs, pru, pzw, *o, stream)
        self._check(h, rc, "ttmpc_solve_batch")
        if host_async and not on_device:
            self._inflight = (getattr(self, "_inflight", []) + [keep])[-4:]  # inputs must outlive the queued copies
        return res

    # ------------------------------------------------------------------ helpers of the closed-loop drivers
    def shift_warm_start(self, z, reference_bug: bool = False, stream=None):
        """``TruckTrailerNMPC._shift_solution`` (mpc_control_nmpc.py:69-88) for a batch ``z [B,8N+6]``."""
        nz = 8 * self.cfg.horizon + 6
        if _is_torch(z):
            import torch

            if not (z.dtype == torch.float64 and z.is_contiguous() and z.shape[-1] == nz and z.device.type == "cuda"
                    and z.device.index == self.device):
                raise ValueError(f"shift_warm_start: z must be a contiguous float64 tensor [..., {nz}] on cuda:{self.device}")
            out = torch.empty_like(z)
            if stream is None:
                stream = torch.cuda.current_stream(z.device).cuda_stream
            h = self._handle(0)
            rc = self._L.ttmpc_shift_warm_start(h, z.numel() // nz, z.data_ptr(), out.data_ptr(), int(reference_bug), stream)
        else:
            z = np.ascontiguousarray(z, dtype=np.float64)
            out = np.empty_like(z)
            h = self._handle(FLAG_HOST_POINTERS)
            rc = self._L.ttmpc_shift_warm_start(h, z.size // nz, z.ctypes.data, out.ctypes.data, int(reference_bug), None)
        self._check(h, rc, "ttmpc_shift_warm_start")
        return out

    def episodes(self, x0, traj_states, traj_inputs, k_seq, disturb: Optional[dict] = None, variant: str = "tracking",
                 seed: int = 0, scenario_ids=None, stream=None) -> dict:
        """B closed-loop episodes entirely on the device (``ttmpc_episode_batch``): ``x0 [B,6]`` torch CUDA tensor,
        ``k_seq`` the host-built window-index sequence (``problem.time_indices``).  Returns ``metrics [B,8]`` (columns:
        distance error, |heading error|, |hitch error|, max|psi|, jackknife, failed solves, mean iterations, RMS
        tracking error) and ``final_state [B,6]``."""
        import torch

        dev = x0.device
        assert x0.dtype == torch.float64 and x0.is_contiguous() and x0.shape[1] == 6
        B = x0.shape[0]
        S = torch.as_tensor(traj_states, dtype=torch.float64, device=dev).contiguous()
        U = torch.as_tensor(traj_inputs, dtype=torch.float64, device=dev).contiguous()
        ks = torch.as_tensor(np.asarray(k_seq, dtype=np.int32), device=dev).contiguous()
        ids = None if scenario_ids is None else scenario_ids.to(device=dev, dtype=torch.int64).contiguous()
        d = None
        if disturb is not None:
            d = (ctypes.c_double * 5)(disturb.get("friction_coeff", 1.0), disturb.get("slippage_coeff", 1.0),
                                      disturb.get("lateral_slip_gain", 0.0), disturb.get("slip_angle_max", 0.0),
                                      disturb.get("process_noise_std", 0.0))
        metrics = torch.empty((B, 8), dtype=torch.float64, device=dev)
        final = torch.empty((B, 6), dtype=torch.float64, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        h = self._handle(0)
        rc = self._L.ttmpc_episode_batch(
            h, B, x0.data_ptr(), None if ids is None else ids.data_ptr(), S.data_ptr(), U.data_ptr(), U.shape[0], ks.data_ptr(),
            int(ks.numel()), ctypes.cast(d, ctypes.c_void_p) if d is not None else None, 0 if variant == "tracking" else 1,
            ctypes.c_uint64(int(seed) & ((1 << 64) - 1)), metrics.data_ptr(), final.data_ptr(), stream)
        self._check(h, rc, "ttmpc_episode_batch")
        self._keepalive = (S, U, ks, ids)
        return {"metrics": metrics, "final_state": final}

    def plant_step(self, q, u, disturb: Optional[dict] = None, noise=None, noise_scale: float = 0.0, stream=None):
        """One Euler plant step ``update(q,u)`` of simulation.py:167-199 for ``q [B,6]``, ``u [B,2]``.

        ``disturb``: dict with keys friction_coeff, slippage_coeff, lateral_slip_gain, slip_angle_max
        (DISTURBANCE_PARAMS of simulation.py:26-32) or None for the nominal plant.
        """
        d = None
        if disturb is not None:
            d = (ctypes.c_double * 4)(
                disturb.get("friction_coeff", 1.0), disturb.get("slippage_coeff", 1.0),
                disturb.get("lateral_slip_gain", 0.0), disturb.get("slip_angle_max", 0.0))
        dptr = ctypes.cast(d, ctypes.c_void_p) if d is not None else None
        if _is_torch(q):
            import torch

            B = q.shape[0] if q.dim() == 2 else -1
            for name, t, shape in (("q", q, (B, 6)), ("u", u, (B, 2)), ("noise", noise, (B, 6))):
                if t is None and name == "noise":
                    continue
                if not (_is_torch(t) and t.dtype == torch.float64 and t.is_contiguous() and t.device == q.device
                        and tuple(t.shape) == shape and t.device.type == "cuda" and t.device.index == self.device):
                    raise ValueError(f"plant_step: {name} must be a contiguous float64 tensor of shape {shape} on cuda:{self.device}")
            out = torch.empty_like(q)
            if stream is None:
                stream = torch.cuda.current_stream(q.device).cuda_stream
            h = self._handle(0)
            rc = self._L.ttmpc_plant_step(h, q.shape[0], q.data_ptr(), u.data_ptr(), dptr,
                                          None if noise is None else noise.data_ptr(), noise_scale, out.data_ptr(), stream)
        else:
            q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 6)
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 2)
            if noise is not None:
                noise = np.ascontiguousarray(noise, dtype=np.float64).reshape(-1, 6)
            out = np.empty_like(q)
            h = self._handle(FLAG_HOST_POINTERS)
            rc = self._L.ttmpc_plant_step(h, q.shape[0], q.ctypes.data, u.ctypes.data, dptr,
                                          None if noise is None else noise.ctypes.data, noise_scale, out.ctypes.data, None)
        self._check(h, rc, "ttmpc_plant_step")
        return out
