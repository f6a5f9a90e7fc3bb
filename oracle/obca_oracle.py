"""CPU oracle for the obstacle-aware (OBCA) controller -- TEST INFRASTRUCTURE ONLY.

Only tests/ (and tools that generate fixtures for tests/) import this module; the product package never does.

PARITY UNPINNED: the reference solves this NLP with CasADi -> Ipopt -> MUMPS (``ca.nlpsol('solver','ipopt',...)``,
python-files/mpc_control_obs.py:211); neither CasADi nor Ipopt is installable in this image and the reference has no
tests or golden vectors for this path.  This file restates

  * the NLP of ``MPCTrackingControlObs`` exactly: decision layout ``[x_k, u_k, mu_k, lam_k]`` per stage and
    ``[x_N, mu_N, lam_N]`` (mpc_control_obs.py:141-176), tracking cost (:32-40), multiple-shooting equalities
    (trajectory_planning.py:28-36), collision rows (:65-139: for every stage, obstacle and body the three OBCA row
    groups ``g'mu - (A_o p_c - b_o)'lam + d_min <= 0``, ``G'mu + R(alpha)'A_o'lam in [-1e-5, 1e-5]^2``,
    ``||A_o'lam||_2 - 1 <= 0``), variable bounds (state/input box, duals >= 0), H-representations and body centres
    (truck_trailer_model.py:31-72, mpc_control_obs.py:42-63), initial guess (:216-239: reference window, mu = 100,
    lam = (100,105,110,115) per body);
  * Ipopt's published algorithm (Waechter & Biegler 2006) the same way oracle/ttmpc_oracle.c does for the plain
    tracking NLP -- inequality rows get slack variables ``d(w) - s = 0`` with bounds on ``s``; monotone barrier
    update, fraction-to-boundary rule, filter line search, inertia-correcting regularisation, scaled termination test
    -- with ONE difference in the linear algebra: here the full KKT matrix is assembled DENSE and factorised with
    LAPACK's symmetric-indefinite LDL' (inertia read off the block diagonal), i.e. no structure is exploited.  The
    CUDA path eliminates the per-(stage, obstacle, body) blocks and runs a Riccati recursion; agreement between the
    two is the parity test.

x_0 is data (x_0 = x_init is eliminated), as in oracle/ttmpc_oracle.c.  Sizes the dense path finishes in seconds:
horizon <= 12 with <= 11 obstacles, or horizon 40 with 1-2 obstacles.

``solve(..., linear_solver="banded")`` is the same algorithm with a linear solver that reaches the reference's own size
(simulation.py:390: horizon 50, the 11 rectangles of obstacles.json -- an 18 664 x 18 664 KKT matrix): the unknowns are
ordered stage by stage, which makes the KKT matrix block tridiagonal (366 x 366 blocks), and a generic block LDL' is run
over it -- every pivot block (a Schur complement) is factorised dense by LAPACK, the inertia of the whole matrix is the
sum of the pivot blocks' inertias (Haynsworth).  Still no knowledge of the (obstacle, body) pair structure, no Riccati
recursion, nothing shared with the CUDA path; tests/test_obca_cpu.py checks that it walks the dense path's iterates.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

NX, NU = 6, 2
BOUND_RELAX, BOUND_PUSH, BOUND_FRAC = 1e-8, 1e-2, 1e-2
KAPPA_EPS, KAPPA_MU, THETA_MU, TAU_MIN, S_MAX, KAPPA_SIGMA = 10.0, 0.2, 1.5, 0.99, 100.0, 1e10
DUAL_INF_TOL, CONSTR_VIOL_TOL, COMPL_INF_TOL = 1.0, 1e-4, 1e-4
ACC_DUAL_INF_TOL, ACC_CONSTR_VIOL_TOL, ACC_COMPL_INF_TOL = 1e10, 1e-2, 1e-2
GAMMA_THETA, GAMMA_PHI, ETA_PHI, S_THETA, S_PHI, DELTA_SW = 1e-5, 1e-8, 1e-8, 1.1, 2.3, 1.0
THETA_MAX_FACT, THETA_MIN_FACT, ALPHA_RED, MAX_BACKTRACK = 1e4, 1e-4, 0.5, 30
EPS = 2.220446049250313e-16
A_O = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])  # mpc_control_obs.py:49-54 (= Gv = Gt)
D_MIN = 0.2  # mpc_control_obs.py:67
C2_HALF_WIDTH = 1e-5  # mpc_control_obs.py:120-123
MU_GUESS = 100.0
LAM_GUESS = np.array([100.0, 105.0, 110.0, 115.0])  # mpc_control_obs.py:226-237
MAX_RESTARTS = 3  # recoveries from an exhausted line search per solve
SEP_TOL = 1e-4    # a start is colliding when a body is closer than d_min - SEP_TOL to an obstacle


def model_f(q, u, L1, L2, M):
    """truck_trailer_model.py:8-24"""
    x, y, th, psi, phi, v = q
    t = np.tan(phi)
    return np.array([
        v * np.cos(th), v * np.sin(th), v * t / L1,
        -v * t / L1 * (1.0 + M / L2 * np.cos(psi)) - v * np.sin(psi) / L2, u[1], u[0]])


def model_jac(q, L1, L2, M):
    x, y, th, psi, phi, v = q
    t = np.tan(phi)
    sec2 = 1.0 + t * t
    F = np.zeros((6, 6))
    F[0, 2], F[0, 5] = -v * np.sin(th), np.cos(th)
    F[1, 2], F[1, 5] = v * np.cos(th), np.sin(th)
    F[2, 4], F[2, 5] = v * sec2 / L1, t / L1
    k = 1.0 + M / L2 * np.cos(psi)
    F[3, 3] = v * t / L1 * (M / L2) * np.sin(psi) - v * np.cos(psi) / L2
    F[3, 4] = -v * sec2 / L1 * k
    F[3, 5] = -t / L1 * k - np.sin(psi) / L2
    return F


def model_hess(q, lam, L1, L2, M):
    """sum_i lam_i * Hessian of f_i"""
    x, y, th, psi, phi, v = q
    t = np.tan(phi)
    sec2 = 1.0 + t * t
    H = np.zeros((6, 6))

    def add(i, j, val):
        H[i, j] += val
        if i != j:
            H[j, i] += val

    add(2, 2, lam[0] * (-v * np.cos(th)) + lam[1] * (-v * np.sin(th)))
    add(2, 5, lam[0] * (-np.sin(th)) + lam[1] * np.cos(th))
    add(4, 4, lam[2] * v * 2 * t * sec2 / L1)
    add(4, 5, lam[2] * sec2 / L1)
    k = 1.0 + M / L2 * np.cos(psi)
    dk = -M / L2 * np.sin(psi)
    add(3, 3, lam[3] * (v * t / L1 * (M / L2) * np.cos(psi) + v * np.sin(psi) / L2))
    add(3, 4, lam[3] * (-v * sec2 / L1 * dk))
    add(3, 5, lam[3] * (-t / L1 * dk - np.cos(psi) / L2))
    add(4, 4, lam[3] * (-v * 2 * t * sec2 / L1 * k))
    add(4, 5, lam[3] * (-sec2 / L1 * k))
    return H


class Pair:
    """One (obstacle, body) block of a stage: 8 local variables v = (mu[4], lam[4]) and 4 rows
    (c1, c2a, c2b, c3); the rows depend on xt = (x, y, theta, psi) of the same stage only."""

    def __init__(self, body, obstacle, geom):
        self.body = body  # 0 vehicle, 1 trailer
        cx, cy = obstacle["center"]
        w, h = obstacle["width"], obstacle["height"]
        self.b = np.array([w / 2, h / 2, w / 2, h / 2]) + A_O @ np.array([cx, cy])  # mpc_control_obs.py:55-63
        L1, L2, M, W1, W2 = geom
        self.g = np.array([L1 / 2, W1 / 2, L1 / 2, W1 / 2]) if body == 0 else np.array([L2 / 2, W2 / 2, L2 / 2, W2 / 2])
        self.L1, self.L2, self.M = L1, L2, M

    def centre(self, xt):
        """body centre, its Jacobian (2x4) and second derivatives (2 x 4 x 4) wrt xt = (x, y, theta, psi)"""
        x, y, th, psi = xt
        J = np.zeros((2, 4))
        H = np.zeros((2, 4, 4))
        J[0, 0] = J[1, 1] = 1.0
        if self.body == 0:  # truck_trailer_model.py:61-64
            a = self.L1 / 2
            pc = np.array([x + np.cos(th) * a, y + np.sin(th) * a])
            J[:, 2] = [-np.sin(th) * a, np.cos(th) * a]
            H[:, 2, 2] = [-np.cos(th) * a, -np.sin(th) * a]
            alpha = th
        else:  # truck_trailer_model.py:66-72
            hl, M = self.L2 / 2, self.M
            al = th + psi
            pc = np.array([x - np.cos(th) * M - np.cos(al) * hl, y - np.sin(th) * M - np.sin(al) * hl])
            J[:, 2] = [np.sin(th) * M + np.sin(al) * hl, -np.cos(th) * M - np.cos(al) * hl]
            J[:, 3] = [np.sin(al) * hl, -np.cos(al) * hl]
            H[:, 2, 2] = [np.cos(th) * M + np.cos(al) * hl, np.sin(th) * M + np.sin(al) * hl]
            H[:, 2, 3] = H[:, 3, 2] = H[:, 3, 3] = [np.cos(al) * hl, np.sin(al) * hl]
            alpha = al
        return pc, J, H, alpha

    def rows(self, xt, v):
        mu, lam = v[:4], v[4:]
        pc, _, _, alpha = self.centre(xt)
        ell = A_O.T @ lam
        m = A_O.T @ mu
        c, s = np.cos(alpha), np.sin(alpha)
        c1 = self.g @ mu - (A_O @ pc - self.b) @ lam + D_MIN
        c2 = m + np.array([c * ell[0] + s * ell[1], -s * ell[0] + c * ell[1]])
        c3 = np.sqrt(ell @ ell) - 1.0
        return np.array([c1, c2[0], c2[1], c3])

    def corners(self, xt):
        """the body's corners (truck_trailer_model.py:31-72) and the obstacle's, as rows"""
        pc, _, _, alpha = self.centre(xt)
        R = np.array([[np.cos(alpha), -np.sin(alpha)], [np.sin(alpha), np.cos(alpha)]])
        loc = np.array([[sx * self.g[0], sy * self.g[1]] for sy in (-1, 1) for sx in (-1, 1)])
        box = np.array([[x, y] for y in (-self.b[3], self.b[1]) for x in (-self.b[2], self.b[0])])
        return pc + loc @ R.T, box, R

    def restore(self, xt):
        """Recovery of the pair's block for a fixed pose: the duals are the multipliers of the
        distance problem between the two rectangles.  For a unit vector n, sep(n) = min over the body of n.q - max over
        the obstacle of n.p; it is largest (= the distance when the rectangles are disjoint) for a face normal of either
        rectangle or a vertex-to-vertex direction.  With ell = A_o' lam = kappa n and m = G' mu = -R' ell in their
        minimal non-negative representation, rows c2 vanish and c1 = d_min - kappa sep; kappa is centred between
        c3 (kappa <= 1) and c1 (kappa >= d_min / sep).  Returns (v, sep)."""
        body, box, R = self.corners(xt)
        cand = [np.array(a, float) for a in ((1, 0), (-1, 0), (0, 1), (0, -1))]
        cand += [sg * R[:, i] for i in (0, 1) for sg in (1.0, -1.0)]
        for q in body:
            for pt in box:
                d = q - pt
                if np.linalg.norm(d) > 1e-12:
                    cand.append(d / np.linalg.norm(d))
        seps = [(body @ n).min() - (box @ n).max() for n in cand]
        i = int(np.argmax(seps))
        sep, n = seps[i], cand[i]
        kappa = 0.5 * (1.0 + D_MIN / sep) if sep > D_MIN else 1.0
        ell = kappa * n
        m = -R.T @ ell
        pos = lambda a: np.array([max(a[0], 0.0), max(a[1], 0.0), max(-a[0], 0.0), max(-a[1], 0.0)])
        return np.concatenate([pos(m), pos(ell)]), sep

    def jac(self, xt, v):
        """Jx (4x4) wrt xt, Jv (4x8) wrt (mu, lam)"""
        mu, lam = v[:4], v[4:]
        pc, Jp, _, alpha = self.centre(xt)
        ell = A_O.T @ lam
        c, s = np.cos(alpha), np.sin(alpha)
        dal = np.array([0.0, 0.0, 1.0, 0.0 if self.body == 0 else 1.0])
        Jx = np.zeros((4, 4))
        Jv = np.zeros((4, 8))
        Jx[0] = -(ell @ Jp)
        Jv[0, :4] = self.g
        Jv[0, 4:] = -(A_O @ pc - self.b)
        Jx[1] = (-s * ell[0] + c * ell[1]) * dal
        Jx[2] = (-c * ell[0] - s * ell[1]) * dal
        Jv[1, :4] = A_O[:, 0]
        Jv[2, :4] = A_O[:, 1]
        Jv[1, 4:] = c * A_O[:, 0] + s * A_O[:, 1]
        Jv[2, 4:] = -s * A_O[:, 0] + c * A_O[:, 1]
        nrm = np.sqrt(ell @ ell)
        Jv[3, 4:] = (A_O @ ell) / nrm
        return Jx, Jv

    def hess(self, xt, v, y):
        """sum_r y_r * Hessian of row r: blocks Wxx (4x4), Wxv (4x8), Wvv (8x8)"""
        lam = v[4:]
        pc, Jp, Hp, alpha = self.centre(xt)
        ell = A_O.T @ lam
        c, s = np.cos(alpha), np.sin(alpha)
        dal = np.array([0.0, 0.0, 1.0, 0.0 if self.body == 0 else 1.0])
        Wxx = np.zeros((4, 4))
        Wxv = np.zeros((4, 8))
        Wvv = np.zeros((8, 8))
        # c1 = ... - pc(xt) . ell(lam)
        Wxx += y[0] * -(ell[0] * Hp[0] + ell[1] * Hp[1])
        Wxv[:, 4:] += y[0] * -(Jp.T @ A_O.T)
        # c2 = m + R(alpha)' ell
        d2 = np.array([-c * ell[0] - s * ell[1], s * ell[0] - c * ell[1]])
        Wxx += (y[1] * d2[0] + y[2] * d2[1]) * np.outer(dal, dal)
        dlam_a = -s * A_O[:, 0] + c * A_O[:, 1]
        dlam_b = -c * A_O[:, 0] - s * A_O[:, 1]
        Wxv[:, 4:] += np.outer(dal, y[1] * dlam_a + y[2] * dlam_b)
        # c3 = ||ell||
        nrm = np.sqrt(ell @ ell)
        eh = ell / nrm
        Wvv[4:, 4:] += y[3] * (A_O @ (np.eye(2) - np.outer(eh, eh)) @ A_O.T) / nrm
        return Wxx, Wxv, Wvv


class ObcaNlp:
    """Index bookkeeping + evaluation of the OBCA NLP with slacks (all variables of stage k contiguous)."""

    def __init__(self, N, dt, L1, L2, M, W1, W2, Q, R, x_lb, x_ub, u_lb, u_ub, obstacles, terminal_weight=1.0,
                 terminal_box=None):
        """terminal_weight, terminal_box: the offline planner's variant of the NLP (trajectory_optimization.py): terminal
        cost weight Q_f = terminal_weight * Q (:181: 100) and, with terminal_box = (goal, half_width), the final-state
        constraint |x_N - goal| <= half_width (:168-173: 1e-2) -- a linear range row whose slack equals x_N - goal, stated
        here as bounds on x_N (intersected with the state bounds)."""
        self.N, self.dt, self.L1, self.L2, self.M = N, dt, L1, L2, M
        self.term_w = float(terminal_weight)
        self.Q = 0.5 * (np.asarray(Q, float).reshape(6, 6) + np.asarray(Q, float).reshape(6, 6).T)
        self.R = 0.5 * (np.asarray(R, float).reshape(2, 2) + np.asarray(R, float).reshape(2, 2).T)
        geom = (L1, L2, M, W1, W2)
        self.pairs = []
        for ob in obstacles:  # mpc_control_obs.py:98-111: vehicle rows then trailer rows per obstacle
            self.pairs.append(Pair(0, ob, geom))
            self.pairs.append(Pair(1, ob, geom))
        self.P = len(self.pairs)
        P = self.P
        # stage k >= 1: x(6) [k<N: u(2)] then per pair v(8), s(4);  stage 0: u(2) then pairs
        self.off = []
        n = 0
        for k in range(N + 1):
            self.off.append(n)
            n += (NX if k >= 1 else 0) + (NU if k < N else 0) + 12 * P
        self.n = n
        self.m = NX * N + 4 * P * (N + 1)
        lo = np.full(n, -np.inf)
        up = np.full(n, np.inf)
        for k in range(N + 1):
            if k >= 1:
                lo[self.ix(k)] = x_lb
                up[self.ix(k)] = x_ub
                if k == N and terminal_box is not None:
                    goal, half = np.asarray(terminal_box[0], float), float(terminal_box[1])
                    lo[self.ix(k)] = np.maximum(np.asarray(x_lb, float), goal - half)
                    up[self.ix(k)] = np.minimum(np.asarray(x_ub, float), goal + half)
            if k < N:
                lo[self.iu(k)] = u_lb
                up[self.iu(k)] = u_ub
            for j in range(P):
                lo[self.iv(k, j)] = 0.0
                s = self.isl(k, j)
                up[s[0]] = 0.0
                lo[s[1]] = lo[s[2]] = -C2_HALF_WIDTH
                up[s[1]] = up[s[2]] = C2_HALF_WIDTH
                up[s[3]] = 0.0
        self.has_lo = np.isfinite(lo) & (lo > -1e19)
        self.has_up = np.isfinite(up) & (up < 1e19)
        self.lo = np.where(self.has_lo, lo - BOUND_RELAX * np.maximum(1.0, np.abs(lo)), -np.inf)
        self.up = np.where(self.has_up, up + BOUND_RELAX * np.maximum(1.0, np.abs(up)), np.inf)
        self.x_lb, self.x_ub = np.asarray(x_lb, float), np.asarray(x_ub, float)

    def Qk(self, k):
        return self.term_w * self.Q if k == self.N else self.Q

    def ix(self, k):
        assert k >= 1
        return np.arange(self.off[k], self.off[k] + NX)

    def iu(self, k):
        o = self.off[k] + (NX if k >= 1 else 0)
        return np.arange(o, o + NU)

    def iv(self, k, j):
        o = self.off[k] + (NX if k >= 1 else 0) + (NU if k < self.N else 0) + 12 * j
        return np.arange(o, o + 8)

    def isl(self, k, j):
        o = self.iv(k, j)[0] + 8
        return np.arange(o, o + 4)

    def ic_dyn(self, k):  # rows of x_k - x_{k-1} - dt f(x_{k-1}, u_{k-1}) = 0, k = 1..N
        return np.arange(NX * (k - 1), NX * k)

    def ic_pair(self, k, j):
        o = NX * self.N + 4 * (k * self.P + j)
        return np.arange(o, o + 4)

    def state(self, w, k, x_init):
        return x_init if k == 0 else w[self.ix(k)]

    def push_inside(self, w):
        w = w.copy()
        lo = np.where(self.has_lo, self.lo, 0.0)
        up = np.where(self.has_up, self.up, 0.0)
        both = self.has_lo & self.has_up
        span = np.where(both, up - lo, np.inf)
        pl = np.minimum(BOUND_PUSH * np.maximum(1.0, np.abs(lo)), BOUND_FRAC * span)
        pu = np.minimum(BOUND_PUSH * np.maximum(1.0, np.abs(up)), BOUND_FRAC * span)
        w = np.where(self.has_lo & (w < lo + pl), lo + pl, w)
        w = np.where(self.has_up & (w > up - pu), up - pu, w)
        return w

    def initial_point(self, x_init, ref_states, ref_inputs, z_warm=None):
        """mpc_control_obs.py:216-239; slacks start at the row values (Ipopt), everything pushed inside its bounds"""
        w = np.zeros(self.n)
        gx, gu = (ref_states, ref_inputs) if z_warm is None else z_warm  # (states [N+1,6], inputs [N,2]) guess of the caller
        for k in range(self.N + 1):
            if k >= 1:
                w[self.ix(k)] = gx[k]
            if k < self.N:
                w[self.iu(k)] = gu[k]
            for j in range(self.P):
                w[self.iv(k, j)] = np.concatenate([np.full(4, MU_GUESS), LAM_GUESS])
        w = self.push_inside(w)
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                w[self.isl(k, j)] = pr.rows(xt, w[self.iv(k, j)])
        return self.push_inside(w)

    def restored_point(self, w, x_init):
        """Recovery from an exhausted line search (where Ipopt would enter feasibility restoration): states and inputs
        stay, the OBCA duals of every pair are replaced by the exact minimiser of their rows' violation for the
        current pose (Pair.restore), slacks are re-seated on their rows, everything is pushed inside like a starting
        point."""
        w = self.push_inside(w)
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                w[self.iv(k, j)] = pr.restore(xt)[0]
        w = self.push_inside(w)
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                w[self.isl(k, j)] = pr.rows(xt, w[self.iv(k, j)])
        return self.push_inside(w)

    def start_clearance(self, x_init):
        """smallest body-obstacle distance at x_init: stage 0 is data, so below d_min the NLP has no feasible point"""
        return min(pr.restore(np.asarray(x_init, float)[:4])[1] for pr in self.pairs)

    def objective(self, w, x_init, ref_states, ref_inputs):
        J = 0.0
        for k in range(1, self.N + 1):  # the k = 0 term is a constant of the data x_init (kept for reporting below)
            d = w[self.ix(k)] - ref_states[k]
            J += d @ self.Qk(k) @ d
        for k in range(self.N):
            d = w[self.iu(k)] - ref_inputs[k]
            J += d @ self.R @ d
        d0 = x_init - ref_states[0]
        return J + d0 @ self.Q @ d0

    def grad(self, w, ref_states, ref_inputs):
        g = np.zeros(self.n)
        for k in range(1, self.N + 1):
            g[self.ix(k)] = 2.0 * self.Qk(k) @ (w[self.ix(k)] - ref_states[k])
        for k in range(self.N):
            g[self.iu(k)] = 2.0 * self.R @ (w[self.iu(k)] - ref_inputs[k])
        return g

    def constraints(self, w, x_init):
        c = np.zeros(self.m)
        for k in range(1, self.N + 1):
            xp = self.state(w, k - 1, x_init)
            c[self.ic_dyn(k)] = w[self.ix(k)] - xp - self.dt * model_f(xp, w[self.iu(k - 1)], self.L1, self.L2, self.M)
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                c[self.ic_pair(k, j)] = pr.rows(xt, w[self.iv(k, j)]) - w[self.isl(k, j)]
        return c

    def jacobian(self, w, x_init):
        Jm = np.zeros((self.m, self.n))
        B = np.zeros((6, 2))
        B[4, 1] = B[5, 0] = self.dt
        for k in range(1, self.N + 1):
            r = self.ic_dyn(k)
            Jm[np.ix_(r, self.ix(k))] = np.eye(6)
            xp = self.state(w, k - 1, x_init)
            if k >= 2:
                Jm[np.ix_(r, self.ix(k - 1))] = -(np.eye(6) + self.dt * model_jac(xp, self.L1, self.L2, self.M))
            Jm[np.ix_(r, self.iu(k - 1))] = -B
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                r = self.ic_pair(k, j)
                Jx, Jv = pr.jac(xt, w[self.iv(k, j)])
                if k >= 1:
                    Jm[np.ix_(r, self.ix(k)[:4])] = Jx
                Jm[np.ix_(r, self.iv(k, j))] = Jv
                Jm[np.ix_(r, self.isl(k, j))] = -np.eye(4)
        return Jm

    def hessian(self, w, y, x_init):
        H = np.zeros((self.n, self.n))
        for k in range(1, self.N + 1):
            H[np.ix_(self.ix(k), self.ix(k))] += 2.0 * self.Qk(k)
        for k in range(self.N):
            H[np.ix_(self.iu(k), self.iu(k))] += 2.0 * self.R
        for k in range(2, self.N + 1):  # dynamics rows k depend nonlinearly on x_{k-1}
            xp = w[self.ix(k - 1)]
            H[np.ix_(self.ix(k - 1), self.ix(k - 1))] -= self.dt * model_hess(xp, y[self.ic_dyn(k)], self.L1, self.L2, self.M)
        for k in range(self.N + 1):
            xt = self.state(w, k, x_init)[:4]
            for j, pr in enumerate(self.pairs):
                Wxx, Wxv, Wvv = pr.hess(xt, w[self.iv(k, j)], y[self.ic_pair(k, j)])
                iv = self.iv(k, j)
                H[np.ix_(iv, iv)] += Wvv
                if k >= 1:
                    ixt = self.ix(k)[:4]
                    H[np.ix_(ixt, ixt)] += Wxx
                    H[np.ix_(ixt, iv)] += Wxv
                    H[np.ix_(iv, ixt)] += Wxv.T
        return H


def stage_blocks(nlp: "ObcaNlp", w, y, x_init):
    """Per stage k: (H_k, Jp_k, Jd_k) = Hessian of the Lagrangian w.r.t. the stage's variables, Jacobian of the stage's
    collision rows, Jacobian of the dynamics rows c_{k+1} w.r.t. the stage's variables (None for k = N).  The same
    formulas as ObcaNlp.jacobian / hessian, written into stage-local matrices."""
    out = []
    B = np.zeros((6, 2))
    B[4, 1] = B[5, 0] = nlp.dt
    P = nlp.P
    for k in range(nlp.N + 1):
        o, nk = nlp.off[k], (nlp.off[k + 1] if k < nlp.N else nlp.n) - nlp.off[k]
        H = np.zeros((nk, nk))
        Jp = np.zeros((4 * P, nk))
        lx = nlp.ix(k) - o if k >= 1 else None
        if k >= 1:
            H[np.ix_(lx, lx)] += 2.0 * nlp.Qk(k)
        if k < nlp.N:
            lu_ = nlp.iu(k) - o
            H[np.ix_(lu_, lu_)] += 2.0 * nlp.R
        Jd = None
        if k < nlp.N:
            Jd = np.zeros((6, nk))
            xp = nlp.state(w, k, x_init)
            if k >= 1:
                Jd[:, lx] = -(np.eye(6) + nlp.dt * model_jac(xp, nlp.L1, nlp.L2, nlp.M))
                H[np.ix_(lx, lx)] -= nlp.dt * model_hess(xp, y[nlp.ic_dyn(k + 1)], nlp.L1, nlp.L2, nlp.M)
            Jd[:, lu_] = -B
        xt = nlp.state(w, k, x_init)[:4]
        for j, pr in enumerate(nlp.pairs):
            iv, isl = nlp.iv(k, j) - o, nlp.isl(k, j) - o
            v = w[nlp.iv(k, j)]
            Jx, Jv = pr.jac(xt, v)
            r = np.arange(4 * j, 4 * j + 4)
            Jp[np.ix_(r, iv)] = Jv
            Jp[np.ix_(r, isl)] = -np.eye(4)
            Wxx, Wxv, Wvv = pr.hess(xt, v, y[nlp.ic_pair(k, j)])
            H[np.ix_(iv, iv)] += Wvv
            if k >= 1:
                Jp[np.ix_(r, lx[:4])] = Jx
                H[np.ix_(lx[:4], lx[:4])] += Wxx
                H[np.ix_(lx[:4], iv)] += Wxv
                H[np.ix_(iv, lx[:4])] += Wxv.T
        out.append((H, Jp, Jd))
    return out


class BandedKKT:
    """Block-tridiagonal LDL' of K = [[H + diag(d), J'], [J, 0]] with the unknowns ordered stage by stage:
    block k = (w_k, y_pair_k, y_dyn_k) -- the multipliers of c_k = x_k - x_{k-1} - dt f(x_{k-1}, u_{k-1}) sit with x_k,
    on which that row has an identity Jacobian, so every pivot block has full-rank constraint rows.  Consecutive
    blocks are coupled through d c_k / d w_{k-1} only."""

    def __init__(self, nlp: "ObcaNlp", blocks, diag):
        self.nlp = nlp
        N, P = nlp.N, nlp.P
        self.wi, self.yi, self.lu, self.Jd = [], [], [], []
        neg = zero = 0
        for k in range(N + 1):
            H, Jp, Jd = blocks[k]
            o = nlp.off[k]
            nk = H.shape[0]
            wi = np.arange(o, o + nk)
            yi = np.concatenate([nlp.ic_pair(k, j) for j in range(P)] + ([nlp.ic_dyn(k)] if k >= 1 else []))
            if k >= 1:
                E = np.zeros((6, nk))
                E[:, nlp.ix(k) - o] = np.eye(6)
                Jk = np.vstack([Jp, E])
            else:
                Jk = Jp
            mk = Jk.shape[0]
            Dk = np.zeros((nk + mk, nk + mk))
            Dk[:nk, :nk] = H + np.diag(diag[wi])
            Dk[:nk, nk:] = Jk.T
            Dk[nk:, :nk] = Jk
            if k >= 1:  # Schur complement of the previous pivot block: lands on the (y_dyn_k, y_dyn_k) corner
                Jprev = self.Jd[k - 1]
                npv = Jprev.shape[1]
                rhs = np.zeros((self.lu[k - 1][0].shape[0], 6))
                rhs[:npv, :] = Jprev.T
                X = sla.lu_solve(self.lu[k - 1], rhs)[:npv, :]
                corr = Jprev @ X
                Dk[-6:, -6:] -= 0.5 * (corr + corr.T)
            l_, d_, perm = sla.ldl(Dk, lower=True)
            ng, zr = _inertia(l_, d_, perm)
            neg += ng
            zero += zr
            self.wi.append(wi)
            self.yi.append(yi)
            self.lu.append(sla.lu_factor(Dk))
            self.Jd.append(Jd)
        self.neg, self.zero = neg, zero

    def solve(self, rhs):
        """rhs, solution in the global ordering (w; y)"""
        nlp = self.nlp
        n, N = nlp.n, nlp.N
        b = [np.concatenate([rhs[self.wi[k]], rhs[n + self.yi[k]]]) for k in range(N + 1)]
        for k in range(1, N + 1):  # forward: b_k -= C_k S_{k-1}^-1 b_{k-1}
            Jprev = self.Jd[k - 1]
            t = sla.lu_solve(self.lu[k - 1], b[k - 1])
            b[k][-6:] -= Jprev @ t[:Jprev.shape[1]]
        x = [None] * (N + 1)
        x[N] = sla.lu_solve(self.lu[N], b[N])
        for k in range(N - 1, -1, -1):  # backward: x_k = S_k^-1 (b_k - C_{k+1}' x_{k+1})
            r = b[k].copy()
            r[:self.Jd[k].shape[1]] -= self.Jd[k].T @ x[k + 1][-6:]
            x[k] = sla.lu_solve(self.lu[k], r)
        sol = np.zeros_like(rhs)
        for k in range(N + 1):
            nk = len(self.wi[k])
            sol[self.wi[k]] = x[k][:nk]
            sol[n + self.yi[k]] = x[k][nk:]
        return sol


def banded_matvec(nlp: "ObcaNlp", blocks, diag, vec):
    """K @ vec without assembling K"""
    n, N, P = nlp.n, nlp.N, nlp.P
    out = np.zeros_like(vec)
    for k in range(N + 1):
        H, Jp, Jd = blocks[k]
        o = nlp.off[k]
        nk = H.shape[0]
        wi = np.arange(o, o + nk)
        yp = np.concatenate([nlp.ic_pair(k, j) for j in range(P)])
        out[wi] += (H + np.diag(diag[wi])) @ vec[wi] + Jp.T @ vec[n + yp]
        out[n + yp] += Jp @ vec[wi]
        if Jd is not None:
            yd = nlp.ic_dyn(k + 1)
            out[wi] += Jd.T @ vec[n + yd]
            out[n + yd] += Jd @ vec[wi] + vec[nlp.ix(k + 1)]
            out[nlp.ix(k + 1)] += vec[n + yd]
    return out


def banded_jt_times(nlp: "ObcaNlp", blocks, y):
    """J' y"""
    out = np.zeros(nlp.n)
    N, P = nlp.N, nlp.P
    for k in range(N + 1):
        H, Jp, Jd = blocks[k]
        wi = np.arange(nlp.off[k], nlp.off[k] + H.shape[0])
        yp = np.concatenate([nlp.ic_pair(k, j) for j in range(P)])
        out[wi] += Jp.T @ y[yp]
        if Jd is not None:
            yd = nlp.ic_dyn(k + 1)
            out[wi] += Jd.T @ y[yd]
            out[nlp.ix(k + 1)] += y[yd]
    return out


def _inertia(lu, d, piv):
    """number of negative / zero eigenvalues of the block-diagonal D of an LDL' factorisation"""
    ev = np.linalg.eigvalsh(d) if False else None
    n = d.shape[0]
    neg = zero = 0
    i = 0
    while i < n:
        if i + 1 < n and d[i + 1, i] != 0.0:
            e = np.linalg.eigvalsh(d[i:i + 2, i:i + 2])
            neg += int((e < 0).sum())
            zero += int((e == 0).sum())
            i += 2
        else:
            neg += d[i, i] < 0
            zero += d[i, i] == 0
            i += 1
    return neg, zero


def solve(nlp: ObcaNlp, x_init, ref_states, ref_inputs, tol=1e-8, acc_tol=1e-6, acc_iter=15, max_iter=5000,
          mu_init=0.1, verbose=False, linear_solver="dense", recover=True, guess=None, geometric_start=False):
    """Returns dict(states[N+1,6], inputs[N,2], obj, iters, status, kkt=(dual_inf, constr_viol, compl), w)."""
    x_init = np.asarray(x_init, float)
    ref_states = np.asarray(ref_states, float).reshape(nlp.N + 1, 6)
    ref_inputs = np.asarray(ref_inputs, float).reshape(nlp.N, 2)
    n, m = nlp.n, nlp.m
    hl, hu = nlp.has_lo, nlp.has_up
    w = nlp.initial_point(x_init, ref_states, ref_inputs, guess)
    if geometric_start:  # TTMPC_OBCA_GEOMETRIC_START: the duals start at the distance problems' multipliers (not the reference's)
        w = nlp.restored_point(w, x_init)
    y = np.zeros(m)
    zl = np.where(hl, 1.0, 0.0)
    zu = np.where(hu, 1.0, 0.0)
    n_b = int(hl.sum() + hu.sum())
    mu = mu_init
    tau = max(TAU_MIN, 1.0 - mu)
    mu_floor = min(tol, COMPL_INF_TOL) / (KAPPA_EPS + 1.0)
    filt = []
    delta_last = 0.0
    acc_count = ls_fail = 0
    status = -1
    theta_max = theta_min = 0.0
    x0_bad = bool(np.any(x_init < nlp.x_lb) or np.any(x_init > nlp.x_ub))
    x0_bad = x0_bad or nlp.start_clearance(x_init) < D_MIN - SEP_TOL  # stage 0 collides: no feasible point
    restarts = 0
    reinit = True

    def slacks(wv):
        return np.where(hl, wv - nlp.lo, 1.0), np.where(hu, nlp.up - wv, 1.0)

    def barrier(wv, muv):
        sl, su = slacks(wv)
        if np.any(sl[hl] <= 0) or np.any(su[hu] <= 0):
            return np.inf
        return nlp.objective(wv, x_init, ref_states, ref_inputs) - muv * (np.log(sl[hl]).sum() + np.log(su[hu]).sum())

    it = 0
    while True:
        c = nlp.constraints(w, x_init)
        g = nlp.grad(w, ref_states, ref_inputs)
        banded = linear_solver == "banded"
        if banded:
            blocks = stage_blocks(nlp, w, y, x_init)
            Jty = banded_jt_times(nlp, blocks, y)
        else:
            Jm = nlp.jacobian(w, x_init)
            Jty = Jm.T @ y
        sl, su = slacks(w)
        theta = np.abs(c).sum()
        if not np.isfinite(theta) or not np.all(np.isfinite(w)):
            status = 4
            break
        if reinit:
            theta_max = THETA_MAX_FACT * max(1.0, theta)
            theta_min = THETA_MIN_FACT * max(1.0, theta)
            reinit = False
        rd = g + Jty - zl + zu
        rd_inf = np.abs(rd).max()
        cinf = np.abs(c).max()
        comp = np.concatenate([(sl * zl)[hl], (su * zu)[hu]])
        cmax, cmin = comp.max(), comp.min()
        y1, z1 = np.abs(y).sum(), zl.sum() + zu.sum()
        s_d = max(S_MAX, (y1 + z1) / (m + n_b)) / S_MAX
        s_c = max(S_MAX, z1 / n_b) / S_MAX

        def E(muv):
            return max(rd_inf / s_d, cinf, max(cmax - muv, muv - cmin) / s_c)

        E0 = E(0.0)
        if verbose:
            print(f"it {it:3d} mu {mu:.1e} obj {nlp.objective(w, x_init, ref_states, ref_inputs):.9e} theta {theta:.2e} "
                  f"rd {rd_inf:.2e} compl {cmax:.2e} E0 {E0:.2e}")
        if E0 <= tol and rd_inf <= DUAL_INF_TOL and cinf <= CONSTR_VIOL_TOL and cmax <= COMPL_INF_TOL:
            status = 0
            break
        if E0 <= acc_tol and rd_inf <= ACC_DUAL_INF_TOL and cinf <= ACC_CONSTR_VIOL_TOL and cmax <= ACC_COMPL_INF_TOL:
            acc_count += 1
        else:
            acc_count = 0
        if acc_iter > 0 and acc_count >= acc_iter:
            status = 1
            break
        if it >= max_iter:
            status = 2
            break
        if x0_bad and it >= 30:
            status = 5
            break
        while mu > mu_floor and E(mu) <= KAPPA_EPS * mu:
            mu = max(mu_floor, min(KAPPA_MU * mu, mu ** THETA_MU))
            tau = max(TAU_MIN, 1.0 - mu)
            filt = []
        # ---- KKT system, dense LDL', inertia-correcting regularisation ----
        sig = np.where(hl, zl / sl, 0.0) + np.where(hu, zu / su, 0.0)
        gphi = g - np.where(hl, mu / sl, 0.0) + np.where(hu, mu / su, 0.0)
        rhs = -np.concatenate([gphi, c])
        if not banded:
            H = nlp.hessian(w, y, x_init)
            K = np.zeros((n + m, n + m))
            K[:n, n:] = Jm.T
            K[n:, :n] = Jm
        delta = 0.0
        ok = False
        for _ in range(40):
            if banded:
                fac = BandedKKT(nlp, blocks, sig + delta)
                neg, zero = fac.neg, fac.zero
            else:
                K[:n, :n] = H + np.diag(sig + delta)
                lu, d, perm = sla.ldl(K, lower=True)
                neg, zero = _inertia(lu, d, perm)
            if neg == m and zero == 0:
                ok = True
                break
            if delta == 0.0:
                delta = 1e-4 if delta_last == 0.0 else max(1e-20, delta_last / 3.0)
            else:
                delta *= 100.0 if delta_last == 0.0 else 8.0
        if not ok:
            status = 4
            break
        if delta > 0:
            delta_last = delta
        if banded:
            sol = fac.solve(rhs)
            sol += fac.solve(rhs - banded_matvec(nlp, blocks, sig + delta, sol))  # one step of iterative refinement
        else:
            sol = np.linalg.solve(K, rhs)
            sol += np.linalg.solve(K, rhs - K @ sol)  # one step of iterative refinement
        dw, yp = sol[:n], sol[n:]
        dzl = np.where(hl, mu / sl - zl - zl / sl * dw, 0.0)
        dzu = np.where(hu, mu / su - zu + zu / su * dw, 0.0)
        a_pr = a_du = 1.0
        mlo = hl & (dw < 0)
        if mlo.any():
            a_pr = min(a_pr, (-tau * sl[mlo] / dw[mlo]).min())
        mup = hu & (dw > 0)
        if mup.any():
            a_pr = min(a_pr, (tau * su[mup] / dw[mup]).min())
        for z, dz in ((zl, dzl), (zu, dzu)):
            mk = dz < 0
            if mk.any():
                a_du = min(a_du, (-tau * z[mk] / dz[mk]).min())
        gphi_d = gphi @ dw
        phi = barrier(w, mu)
        roundoff = theta <= 1e-2 * tol and abs(gphi_d) <= max(100 * EPS * max(1.0, abs(phi)), theta * y1)
        accepted = roundoff
        alpha = a_pr
        if not roundoff:
            for bt in range(MAX_BACKTRACK + 1):
                wt = w + alpha * dw
                phit = barrier(wt, mu)
                tht = np.abs(nlp.constraints(wt, x_init)).sum()
                good = np.isfinite(phit) and np.isfinite(tht) and tht <= theta_max and \
                    all(not (tht >= ft and phit >= fp) for ft, fp in filt)
                if good:
                    switching = gphi_d < 0 and alpha * (-gphi_d) ** S_PHI > DELTA_SW * theta ** S_THETA
                    if theta <= theta_min and switching:
                        okstep = phit - phi - 10 * EPS * abs(phi) <= ETA_PHI * alpha * gphi_d
                        ftype = 1
                    else:
                        okstep = (tht - (1 - GAMMA_THETA) * theta <= 10 * EPS * abs(theta)) or \
                            (phit - (phi - GAMMA_PHI * theta) <= 10 * EPS * abs(phi))
                        ftype = 0
                    if okstep:
                        if not ftype:
                            filt.append(((1 - GAMMA_THETA) * theta, phi - GAMMA_PHI * theta))
                        accepted = True
                        break
                alpha *= ALPHA_RED
        if verbose:
            print(f"      delta {delta:.1e} a_pr {a_pr:.3e} a_du {a_du:.3e} alpha {alpha:.3e} gphi_d {gphi_d:.3e} acc {accepted}")
        if not accepted and recover and restarts < MAX_RESTARTS:
            # the line search is exhausted: Ipopt would enter feasibility restoration.  Recovery: restored point, all
            # multipliers and the barrier parameter back to their initial values, empty filter.
            restarts += 1
            w = nlp.restored_point(w, x_init)
            y = np.zeros(m)
            zl = np.where(hl, 1.0, 0.0)
            zu = np.where(hu, 1.0, 0.0)
            mu = mu_init
            tau = max(TAU_MIN, 1.0 - mu)
            delta_last = 0.0
            filt = []
            acc_count = ls_fail = 0
            reinit = True
            it += 1
            continue
        if not accepted:
            ls_fail += 1
            if ls_fail >= 3:
                status = 3
                break
            alpha = a_pr * ALPHA_RED ** MAX_BACKTRACK
            filt = []
        else:
            ls_fail = 0
        w = w + alpha * dw
        y = y + alpha * (yp - y)
        zl = zl + a_du * dzl
        zu = zu + a_du * dzu
        sl, su = slacks(w)
        zl = np.where(hl, np.maximum(np.minimum(zl, KAPPA_SIGMA * mu / sl), mu / (KAPPA_SIGMA * sl)), 0.0)
        zu = np.where(hu, np.maximum(np.minimum(zu, KAPPA_SIGMA * mu / su), mu / (KAPPA_SIGMA * su)), 0.0)
        it += 1

    if x0_bad and status >= 2:
        status = 5
    states = np.vstack([x_init] + [w[nlp.ix(k)] for k in range(1, nlp.N + 1)])
    inputs = np.vstack([w[nlp.iu(k)] for k in range(nlp.N)])
    return dict(states=states, inputs=inputs, obj=nlp.objective(w, x_init, ref_states, ref_inputs), iters=it,
                status=status, kkt=(rd_inf, cinf, cmax), w=w, y=y, restarts=restarts)
