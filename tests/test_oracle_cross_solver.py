"""Pins the oracle's SOLVE (a9) in the absence of a runnable Ipopt ("parity unpinned"): the same NLP,
stated independently in numpy with x_0 as a bounded decision variable exactly like
trajectory_planning.py:28-60, is solved by SciPy SLSQP (a different algorithm: active-set SQP) and must
reach the oracle's minimiser; known answers are the survey probe values of SURVEY.md Appendix E.2."""
import numpy as np
import pytest
from scipy.optimize import minimize

import nlp_numpy as nlp
from car_trailer_mpc_b200 import problem as pb
from oracle import oracle

KNOWN = {  # k0: (objective, u0)  -- SURVEY.md Appendix E.2 (seed 0, k0 = 0, 100, 200 in that order)
    0: (0.342818049, (-4.980490, 1.516745)),
    100: (0.173029518, (0.014492, -0.035013)),
    200: (0.191797632, (0.874506, 0.049064)),
}


@pytest.fixture(scope="module")
def cases(traj, cfg40):
    S, U = traj
    rng = np.random.default_rng(0)
    out = []
    for k0 in (0, 100, 200):
        x0 = S[k0] + rng.normal(0, 0.02, 6)
        xs, us = pb.window(S, U, k0, 40)
        out.append((k0, x0, xs, us))
    return out


def test_known_answers(cases, cfg40):
    for k0, x0, xs, us in cases:
        r = oracle.solve(cfg40, x0, xs, us)
        obj, u0 = KNOWN[k0]
        assert r["status"] == 0
        assert abs(r["obj"] - obj) < 5e-9
        assert np.abs(r["u0"] - np.array(u0)).max() < 2e-6
        assert r["iters"] <= 10
        assert np.array_equal(r["z"][:6], x0)  # states[:,0] == x_init


def test_slsqp_agrees(cases, cfg40):
    lb, ub = nlp.bounds(cfg40)
    bnds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(lb, ub)]
    for k0, x0, xs, us in cases:
        z0 = pb.pack_z(xs, us)
        res = minimize(
            lambda z: nlp.cost(cfg40, z, xs, us), z0, jac=lambda z: nlp.cost_grad(cfg40, z, xs, us),
            method="SLSQP", bounds=bnds,
            constraints=[{"type": "eq", "fun": lambda z: nlp.constraints(cfg40, z, x0),
                          "jac": lambda z: nlp.constraints_jac(cfg40, z, x0)}],
            options={"ftol": 1e-12, "maxiter": 400},
        )
        assert res.success, res.message
        r = oracle.solve(cfg40, x0, xs, us)
        assert abs(res.fun - r["obj"]) < 1e-6 * max(1.0, abs(r["obj"]))
        assert np.abs(res.x[6:8] - r["u0"]).max() < 1e-5
        assert np.abs(res.x - r["z"]).max() < 1e-4


def test_kkt_certificate_is_solver_independent(cfg40):
    sc = pb.make_scenarios(cfg40, 24, seed=7)
    r = oracle.solve_batch(cfg40, sc.x_init, sc.ref_states, sc.ref_inputs)
    assert (r["status"] == 0).all()
    for i in range(24):
        stat, viol, bviol, neg = nlp.kkt_certificate(cfg40, r["z"][i], sc.x_init[i], sc.ref_states[i], sc.ref_inputs[i])
        assert stat < 1e-5, (i, stat)
        assert viol < 1e-9
        assert bviol < 2e-8          # Ipopt-style bound relaxation 1e-8*max(1,|b|)
        assert neg > -1e-6
        assert abs(nlp.cost(cfg40, r["z"][i], sc.ref_states[i], sc.ref_inputs[i]) - r["obj"][i]) < 1e-12
