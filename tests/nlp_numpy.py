"""Independent numpy statement of the reference NLP (x_0 kept as a decision variable, exactly as
trajectory_planning.py:28-60 builds it) used to cross-check the oracle with SciPy SLSQP and to
compute solver-independent KKT certificates.  Test helper, not product code."""
import numpy as np


def unpack(z, N):
    body = z[: 8 * N].reshape(N, 8)
    X = np.vstack([body[:, :6], z[8 * N :][None]])
    return X, body[:, 6:]


def f(cfg, X, U):
    th, psi, phi, v = X[:, 2], X[:, 3], X[:, 4], X[:, 5]
    t = np.tan(phi)
    return np.stack(
        [v * np.cos(th), v * np.sin(th), v * t / cfg.L1,
         -v * t / cfg.L1 * (1 + cfg.M / cfg.L2 * np.cos(psi)) - v * np.sin(psi) / cfg.L2, U[:, 1], U[:, 0]], axis=1)


def cost(cfg, z, Xr, Ur):
    N = cfg.horizon
    X, U = unpack(z, N)
    Q, R = cfg.Qm(), cfg.Rm()
    dx, du = X - Xr, U - Ur
    return np.einsum("ki,ij,kj->", dx, Q, dx) + np.einsum("ki,ij,kj->", du, R, du)


def cost_grad(cfg, z, Xr, Ur):
    N = cfg.horizon
    X, U = unpack(z, N)
    Q, R = cfg.Qm(), cfg.Rm()
    g = np.zeros_like(z)
    gb = g[: 8 * N].reshape(N, 8)
    gx = 2 * (X - Xr) @ Q.T
    gb[:, :6] = gx[:N]
    gb[:, 6:] = 2 * (U - Ur) @ R.T
    g[8 * N :] = gx[N]
    return g


def constraints(cfg, z, x_init):
    """g of trajectory_planning.py:28-36: [x_0 - x_init; x_{k+1} - x_k - dt f(x_k,u_k)]."""
    N = cfg.horizon
    X, U = unpack(z, N)
    c = X[1:] - X[:-1] - cfg.dt * f(cfg, X[:-1], U)
    return np.concatenate([X[0] - x_init, c.ravel()])


def fx(cfg, x):
    th, psi, phi, v = x[2], x[3], x[4], x[5]
    t = np.tan(phi); s = 1 + t * t; c = cfg.M / cfg.L2
    F = np.zeros((6, 6))
    F[0, 2] = -v * np.sin(th); F[0, 5] = np.cos(th)
    F[1, 2] = v * np.cos(th); F[1, 5] = np.sin(th)
    F[2, 4] = v * s / cfg.L1; F[2, 5] = t / cfg.L1
    F[3, 3] = v * t / cfg.L1 * c * np.sin(psi) - v * np.cos(psi) / cfg.L2
    F[3, 4] = -v * s / cfg.L1 * (1 + c * np.cos(psi))
    F[3, 5] = -t / cfg.L1 * (1 + c * np.cos(psi)) - np.sin(psi) / cfg.L2
    return F


def constraints_jac(cfg, z, x_init):
    N = cfg.horizon
    X, U = unpack(z, N)
    n = 8 * N + 6
    J = np.zeros((6 * (N + 1), n))
    J[:6, :6] = np.eye(6)
    Bm = np.zeros((6, 2)); Bm[4, 1] = cfg.dt; Bm[5, 0] = cfg.dt
    for k in range(N):
        r = 6 * (k + 1)
        A = np.eye(6) + cfg.dt * fx(cfg, X[k])
        J[r : r + 6, 8 * k : 8 * k + 6] = -A
        J[r : r + 6, 8 * k + 6 : 8 * k + 8] = -Bm
        J[r : r + 6, 8 * (k + 1) : 8 * (k + 1) + 6] = np.eye(6)
    return J


def bounds(cfg):
    N = cfg.horizon
    lb = np.concatenate([np.tile(np.r_[cfg.x_lb[:], cfg.u_lb[:]], N), cfg.x_lb[:]])
    ub = np.concatenate([np.tile(np.r_[cfg.x_ub[:], cfg.u_ub[:]], N), cfg.x_ub[:]])
    return lb, ub


def kkt_certificate(cfg, z, x_init, Xr, Ur, active_tol=1e-6, with_compl=False):
    """Solver-independent first-order check: with the active set read off z, find multipliers by
    least squares and return (stationarity residual inf-norm, constraint violation, bound violation,
    most negative bound multiplier sign violation [, complementarity max(multiplier * slack)]).

    An interior-point solution at barrier parameter mu carries a multiplier mu/slack on every bounded variable, so
    with a tight `active_tol` a variable that sits 3e-6 off its bound leaves mu/3e-6 ~ 3e-4 in the stationarity
    residual; a wider active set (1e-3) lets such variables have a multiplier and `with_compl` then checks that
    multiplier * slack is of the order of mu."""
    lb, ub = bounds(cfg)
    g = cost_grad(cfg, z, Xr, Ur)
    J = constraints_jac(cfg, z, x_init)
    act_l = np.isfinite(lb) & (z - lb <= active_tol)
    act_u = np.isfinite(ub) & (ub - z <= active_tol)
    n = z.size
    E = np.zeros((n, int(act_l.sum() + act_u.sum())))
    sign = []
    col = 0
    for i in np.nonzero(act_l)[0]:
        E[i, col] = -1.0; col += 1; sign.append(1)
    for i in np.nonzero(act_u)[0]:
        E[i, col] = 1.0; col += 1; sign.append(1)
    M = np.hstack([J.T, E])
    sol, *_ = np.linalg.lstsq(M, -g, rcond=None)
    stat = np.abs(g + M @ sol).max()
    mult_b = sol[J.shape[0] :]
    neg = float(min(0.0, mult_b.min())) if mult_b.size else 0.0
    viol = np.abs(constraints(cfg, z, x_init)).max()
    bviol = max(0.0, float((lb - z)[np.isfinite(lb)].max()), float((z - ub)[np.isfinite(ub)].max()))
    if with_compl:
        slack = np.concatenate([(z - lb)[act_l], (ub - z)[act_u]])
        compl = float(np.max(np.abs(mult_b) * np.maximum(slack, 0.0))) if mult_b.size else 0.0
        return stat, viol, bviol, neg, compl
    return stat, viol, bviol, neg
