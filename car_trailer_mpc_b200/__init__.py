"""B200-native batched NMPC solver for the truck-trailer tracking controller of
Avan1ko/car-trailer-mpc (drop-in for ``controller.solve`` of mpc_control.py / mpc_control_nmpc.py).

Public surface:
  * :class:`BatchSolver`           -- thin ctypes front end of the C ABI in ``include/ttmpc.h`` (CUDA only).
  * :class:`MPCTrackingControl`, :class:`TruckTrailerNMPC`, :class:`MPCTrackingControlFuzzy`, :class:`MPCTrackingControlObs`,
    :class:`TrajectoryOptimization` (the offline planner), :class:`TruckTrailerModel` -- shims with
    the reference's constructor / ``solve`` signatures.
  * :mod:`problem`                 -- trajectory, windows, layouts, synthetic scenario batches.
  * :mod:`closed_loop`, :mod:`batch_driver` -- headless closed loop and the sweep/CSV driver.
"""
from .config import (  # noqa: F401
    Config,
    Obstacles,
    STATUS_NAMES,
    nmpc_preset,
    parking_lot_obstacles,
    planner_preset,
    tracking_preset,
)

__all__ = ["Config", "Obstacles", "tracking_preset", "nmpc_preset", "planner_preset", "parking_lot_obstacles", "STATUS_NAMES"]


def __getattr__(name):  # lazy: importing the package must not require the CUDA library
    if name == "BatchSolver":
        from .solver import BatchSolver

        return BatchSolver
    if name == "MPCTrackingControl":
        from .mpc_control import MPCTrackingControl

        return MPCTrackingControl
    if name == "TruckTrailerNMPC":
        from .mpc_control_nmpc import TruckTrailerNMPC

        return TruckTrailerNMPC
    if name == "MPCTrackingControlFuzzy":
        from .mpc_control_fuzzy import MPCTrackingControlFuzzy

        return MPCTrackingControlFuzzy
    if name == "MPCTrackingControlObs":
        from .mpc_control_obs import MPCTrackingControlObs

        return MPCTrackingControlObs
    if name == "TrajectoryOptimization":
        from .trajectory_optimization import TrajectoryOptimization

        return TrajectoryOptimization
    if name == "SwitchingController":
        from .mpc_control_switch import SwitchingController

        return SwitchingController
    if name == "TruckTrailerModel":
        from .truck_trailer_model import TruckTrailerModel

        return TruckTrailerModel
    raise AttributeError(name)
