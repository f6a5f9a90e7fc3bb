"""Shared parity assertions: the north-star tolerances (BASELINE.json) written down once."""
import numpy as np

from car_trailer_mpc_b200 import problem as pb

U0_ABS_TOL = 1e-4       # first-step controls within 1e-4 absolute
OBJ_REL_TOL = 1e-6      # objective within 1e-6 relative
VIOL_TOL = 1e-6         # constraint violation <= 1e-6


def assert_parity(cfg, got, ref, x_init=None, check_z=True, label=""):
    """`got` = solver under test, `ref` = oracle (dicts with z,u0,obj,status,iters as numpy arrays)."""
    gs, rs = np.asarray(got["status"]), np.asarray(ref["status"])
    assert np.array_equal(gs, rs), f"{label} status mismatch: {np.bincount(gs, minlength=6)} vs {np.bincount(rs, minlength=6)}"
    du0 = np.abs(np.asarray(got["u0"]) - ref["u0"]).max()
    assert du0 <= U0_ABS_TOL, f"{label} |du0| = {du0}"
    rel = np.abs(np.asarray(got["obj"]) - ref["obj"]) / np.maximum(np.abs(ref["obj"]), 1e-12)
    assert rel.max() <= OBJ_REL_TOL, f"{label} objective rel err {rel.max()}"
    if check_z and got.get("z") is not None:
        N = cfg.horizon
        z = np.asarray(got["z"])
        X, U = pb.unpack_z(z, N)
        viol = np.abs(pb.dynamics_defect(cfg, X, U)).max()
        assert viol <= VIOL_TOL, f"{label} dynamics defect {viol}"
        if x_init is not None:
            assert np.array_equal(X[:, 0, :], np.asarray(x_init).reshape(-1, 6)), "states[:,0] must equal x_init"
        lbx, ubx = np.array(cfg.x_lb[:]), np.array(cfg.x_ub[:])
        lbu, ubu = np.array(cfg.u_lb[:]), np.array(cfg.u_ub[:])
        slack = 1e-8 * 11  # bound_relax_factor * max(1,|b|)
        assert (X[:, 1:, :] >= lbx - slack).all() and (X[:, 1:, :] <= ubx + slack).all(), f"{label} state bound violated"
        assert (U >= lbu - slack).all() and (U <= ubu + slack).all(), f"{label} input bound violated"
        assert np.abs(z - ref["z"]).max() <= 1e-4, f"{label} |dz| = {np.abs(z - ref['z']).max()}"
    return du0, rel.max()
