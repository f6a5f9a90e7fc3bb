"""Known answers from an INDEPENDENT algorithm for the obstacle-aware NLP: SciPy SLSQP (active-set SQP) on the literal NLP
of mpc_control_obs.py -- decision variables (x_k, u_k, mu_k, lam_k) with x_0 pinned by an equality, the collision rows of
:65-139 as INEQUALITIES (no slacks, no barrier), bounds of :141-176 -- written here from the reference's formulas with
numpy only.  Neither the oracle nor the kernels are involved; gradients are SLSQP's own finite differences except for the
cost.   python tools/make_golden_slsqp_obca.py   -> tests/golden/obca_slsqp.npz
Cases: two of tests/golden/obca_cases.npz's problems (inputs copied from there: x_init, window, rectangles) -- a
2-obstacle case without active rows and the case whose solution is pressed against a blocking obstacle (SLSQP stops there
with "positive directional derivative for linesearch", i.e. at the resolution of its finite differences: feasible to 6e-13,
objective within 1.1e-6 of the oracle's).  The third small case with active rows, n12_k200_blocked3 (three obstacles), is
not a known answer: from the same start SLSQP slides into a second, cheaper local minimum (215.56 vs 217.63, both
feasible); started at the oracle's states and inputs with the closed-form duals of the poses it stays there (60 iterations:
|dx| 3e-6, objective 217.630) -- the oracle's point is a local minimiser, the NLP is not convex."""
import os
import subprocess
import sys
import time

import numpy as np
import scipy
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from car_trailer_mpc_b200 import tracking_preset  # noqa: E402

L1, L2, M, W1, W2, DT, D_MIN = 7.05, 12.45, 0.15, 3.05, 2.95, 0.05, 0.2
A = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])  # mpc_control_obs.py:49-54, truck_trailer_model.py:31-58


def f(q, u):  # truck_trailer_model.py:8-24
    x, y, th, psi, phi, v = q
    return np.array([v * np.cos(th), v * np.sin(th), v * np.tan(phi) / L1,
                     -v * np.tan(phi) / L1 * (1 + M / L2 * np.cos(psi)) - v * np.sin(psi) / L2, u[1], u[0]])


class Nlp:
    def __init__(self, N, rects, x_init, xs, us, cfg):
        self.N, self.P, self.x_init, self.xs, self.us = N, len(rects), x_init, xs, us
        self.b = [np.array([w / 2, h / 2, w / 2, h / 2]) + A @ np.array([cx, cy]) for cx, cy, w, h in rects]  # :55-63
        self.nv = 16 * self.P
        self.n = (N + 1) * (6 + self.nv) + 2 * N
        self.cfg = cfg

    def split(self, z):
        N, nv = self.N, self.nv
        per = 8 + nv
        X = np.stack([z[k * per:k * per + 6] for k in range(N)] + [z[N * per:N * per + 6]])
        U = np.stack([z[k * per + 6:k * per + 8] for k in range(N)])
        V = [z[k * per + 8:k * per + 8 + nv] for k in range(N)] + [z[N * per + 6:N * per + 6 + nv]]
        return X, U, V

    def cost(self, z):  # mpc_control.py:17-25
        X, U, _ = self.split(z)
        Q, R = np.eye(6), 10.0 * np.eye(2)
        d = X - self.xs
        du = U - self.us
        return float(np.einsum("ki,ij,kj->", d, Q, d) + np.einsum("ki,ij,kj->", du, R, du))

    def eq(self, z):  # trajectory_planning.py:28-36
        X, U, _ = self.split(z)
        out = [X[0] - self.x_init]
        for k in range(self.N):
            out.append(X[k + 1] - X[k] - DT * f(X[k], U[k]))
        return np.concatenate(out)

    def ineq(self, z):  # SLSQP wants g(z) >= 0: the rows of mpc_control_obs.py:98-137 with their bounds
        X, _, V = self.split(z)
        out = []
        for k in range(self.N + 1):
            x, y, th, psi = X[k][:4]
            pv = np.array([x + np.cos(th) * L1 / 2, y + np.sin(th) * L1 / 2])  # truck_trailer_model.py:61-64
            pt = np.array([x - np.cos(th) * M - np.cos(th + psi) * L2 / 2, y - np.sin(th) * M - np.sin(th + psi) * L2 / 2])  # :66-72
            for i in range(self.P):
                v = V[k][16 * i:16 * i + 16]
                for body, (pc, al, g) in enumerate(((pv, th, np.array([L1 / 2, W1 / 2, L1 / 2, W1 / 2])),
                                                    (pt, th + psi, np.array([L2 / 2, W2 / 2, L2 / 2, W2 / 2])))):
                    mu, lam = v[8 * body:8 * body + 4], v[8 * body + 4:8 * body + 8]
                    Rm = np.array([[np.cos(al), -np.sin(al)], [np.sin(al), np.cos(al)]])
                    c1 = g @ mu - (A @ pc - self.b[i]) @ lam + D_MIN
                    c2 = A.T @ mu + Rm.T @ A.T @ lam
                    c3 = np.linalg.norm(A.T @ lam) - 1.0
                    out += [-c1, 1e-5 - c2[0], c2[0] + 1e-5, 1e-5 - c2[1], c2[1] + 1e-5, -c3]
        return np.array(out)

    def bounds(self):
        cfg = self.cfg
        lo, up = [], []
        for k in range(self.N + 1):
            lo += list(cfg.x_lb)
            up += list(cfg.x_ub)
            if k < self.N:
                lo += list(cfg.u_lb)
                up += list(cfg.u_ub)
            lo += [0.0] * self.nv
            up += [np.inf] * self.nv
        return [(None if not np.isfinite(a) else a, None if not np.isfinite(b) else b) for a, b in zip(lo, up)]

    def start(self, warm):
        """SLSQP is a local method: it is started at the oracle's states/inputs perturbed by 1e-3 with duals from the
        geometry-free constants scaled down -- the known answer is the minimiser it converges to, not the path."""
        N, per = self.N, 8 + self.nv
        z = np.zeros(self.n)
        for k in range(N + 1):
            o = k * per if k < N else N * per
            z[o:o + 6] = warm[0][k]
            if k < N:
                z[o + 6:o + 8] = warm[1][k]
            z[(o + 8 if k < N else o + 6):(o + 8 if k < N else o + 6) + self.nv] = 0.5
        return z


if __name__ == "__main__":
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    g = np.load(os.path.join(ROOT, "tests", "golden", "obca_cases.npz"))
    out = {}
    rng = np.random.default_rng(77)
    for name in ("n6_k300_2obs", "n12_k60_blocked"):
        N = int(g[name + "/horizon"])
        cfg = tracking_preset(N)
        p = Nlp(N, [tuple(r) for r in g[name + "/rects"]], g[name + "/x_init"], g[name + "/ref_states"], g[name + "/ref_inputs"], cfg)
        # start: the reference window (the controllers' cold start) shifted off the obstacle by the oracle's solution +
        # noise; SLSQP never sees the oracle's duals
        warm = (g[name + "/states"] + rng.normal(0, 1e-3, (N + 1, 6)), g[name + "/inputs"] + rng.normal(0, 1e-3, (N, 2)))
        warm[0][0] = g[name + "/x_init"]
        t = time.time()
        res = minimize(p.cost, p.start(warm), method="SLSQP", bounds=p.bounds(),
                       constraints=[{"type": "eq", "fun": p.eq}, {"type": "ineq", "fun": p.ineq}], options={"ftol": 1e-13, "maxiter": 600})
        X, U, _ = p.split(res.x)
        viol = max(float(np.abs(p.eq(res.x)).max()), float(-min(0.0, p.ineq(res.x).min())))
        print(f"{name}: success {res.success} nit {res.nit} obj {res.fun:.10e} (oracle {float(g[name + '/obj']):.10e}) viol {viol:.1e} "
              f"dx {np.abs(X - g[name + '/states']).max():.1e} du {np.abs(U - g[name + '/inputs']).max():.1e} ({time.time() - t:.0f} s)", flush=True)
        out[name + "/states"], out[name + "/inputs"], out[name + "/obj"] = X, U, res.fun
        out[name + "/success"] = int(res.status in (0, 8) and viol < 1e-7)  # 8: positive directional derivative (precision reached)
        out[name + "/slsqp_status"] = int(res.status)
        out[name + "/nit"] = res.nit
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "obca_slsqp.npz"), scipy_version=scipy.__version__,
                        generator="tools/make_golden_slsqp_obca.py", git_commit=commit, seed=77, **out)
