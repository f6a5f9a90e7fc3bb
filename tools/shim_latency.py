"""Wall-clock latency of the drop-in shims as a user of the reference calls them (host arrays in, host arrays out, one
problem): MPCTrackingControl.solve, TruckTrailerNMPC.solve, MPCTrackingControlObs.solve (reference start and the opt-in
geometric start), TrajectoryOptimization.plan.  usage: shim_latency.py [repeats]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from car_trailer_mpc_b200 import MPCTrackingControl, MPCTrackingControlObs, TruckTrailerModel, TruckTrailerNMPC, TrajectoryOptimization
from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200.config import parking_lot_obstacles

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20
S, U = pb.load_reference_trajectory()
pi = np.pi
sb = {"lb": [-np.inf, -np.inf, -pi, -pi / 3, -pi / 4, -10.0], "ub": [np.inf, np.inf, pi, pi / 3, pi / 4, 10.0]}
ib = {"lb": [-5.0, -pi / 2], "ub": [5.0, pi / 2]}


def timeit(f, n=R):
    f(); f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t)), float(np.max(t))


def window(N, k=100):
    rs, ru = pb.window(S, U, k, N)
    return rs[0] + np.array([0.05, -0.05, 0.01, 0.0, 0.0, 0.0]), rs.T.copy(), ru.T.copy()


for N in (40,):
    params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": N}
    args = (TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib)
    x0, rs, ru = window(N)
    c = MPCTrackingControl(*args)
    print("MPCTrackingControl.solve        N=%d: median %.3f ms, max %.3f ms (%d iterations)" % (N, *timeit(lambda: c.solve(x0, rs, ru)), c.last_iterations))
    c = TruckTrailerNMPC(*args)
    print("TruckTrailerNMPC.solve          N=%d: median %.3f ms, max %.3f ms (%d iterations)" % (N, *timeit(lambda: c.solve(x0, rs, ru)), c.last_iterations))
N = 50
params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": N}
args = (TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib)
x0, rs, ru = window(N)
for geo in (False, True):
    c = MPCTrackingControlObs(*args, obstacle_list=parking_lot_obstacles(), geometric_start=geo)
    print("MPCTrackingControlObs.solve     N=%d, 11 obstacles, %s start: median %.3f ms, max %.3f ms (%d iterations)"
          % (N, "geometric" if geo else "reference", *timeit(lambda: c.solve(x0, rs, ru)), c.last_iterations))
N = 200
St = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T
params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.1, "horizon": N}
goal = St[N].copy(); goal[4:] = 0.0
for geo in (False, True):
    pl = TrajectoryOptimization(TruckTrailerModel(params), params, np.eye(6), np.eye(2), sb, ib, parking_lot_obstacles(),
                                waypoints={"Positions": St[::20, 0:2].tolist(), "Headings": (St[::20, 2] - np.pi / 2.0).tolist(),
                                           "HitchAngles": St[::20, 3].tolist()}, geometric_start=geo)
    try:
        print("TrajectoryOptimization.plan     N=%d, 11 obstacles, %s start: median %.3f ms, max %.3f ms (%s iterations, status %s)"
              % (N, "geometric" if geo else "reference", *timeit(lambda: pl.plan(St[0], goal), 5), pl.last_iterations, pl.last_status))
    except Exception as e:
        print("planner:", type(e).__name__, e)
