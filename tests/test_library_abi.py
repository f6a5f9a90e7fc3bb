"""The C-ABI shared library loads on a CPU-only machine, exports every symbol include/ttmpc.h declares,
and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from car_trailer_mpc_b200 import _lib, build, tracking_preset
from car_trailer_mpc_b200.config import Config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build.build_library()
    return _lib.load()


def test_header_symbols_are_exported(L):
    hdr = open(os.path.join(ROOT, "include", "ttmpc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ttmpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_config_struct_matches_header(L):
    c = Config()
    L.ttmpc_default_config(ctypes.byref(c), 40)
    ref = tracking_preset(40)
    assert ctypes.sizeof(Config) == 16 + 8 * (4 + 36 + 4 + 12 + 4 + 3)
    for f, _ in Config._fields_:
        a, b = getattr(c, f), getattr(ref, f)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), f
        else:
            assert a == b, f


def test_obstacle_struct_matches_header():
    """struct ttmpc_obstacles: {int32 count, int32 flags, double rect[MAX][4], double W1, W2, d_min}."""
    from car_trailer_mpc_b200.config import MAX_OBSTACLES, Obstacles, parking_lot_obstacles
    hdr = open(os.path.join(ROOT, "include", "ttmpc.h")).read()
    assert int(re.search(r"#define TTMPC_MAX_OBSTACLES (\d+)", hdr).group(1)) == MAX_OBSTACLES
    assert ctypes.sizeof(Obstacles) == 8 + MAX_OBSTACLES * 4 * 8 + 3 * 8
    assert Obstacles.rect.offset == 8 and Obstacles.W1.offset == 8 + MAX_OBSTACLES * 32
    o = Obstacles.from_list(parking_lot_obstacles())
    assert o.count == 11 and o.as_list() == parking_lot_obstacles() and (o.W1, o.W2, o.d_min) == (3.05, 2.95, 0.2)
    # flag values of the header == the Python constants; the reference's behaviour (recovery on, reference start) is 0
    from car_trailer_mpc_b200.config import OBCA_GEOMETRIC_START, OBCA_NO_RECOVERY
    assert int(re.search(r"#define TTMPC_OBCA_NO_RECOVERY (\d+)", hdr).group(1)) == OBCA_NO_RECOVERY
    assert int(re.search(r"#define TTMPC_OBCA_GEOMETRIC_START (\d+)", hdr).group(1)) == OBCA_GEOMETRIC_START
    assert o.flags == 0
    assert Obstacles.from_list(parking_lot_obstacles(), recover=False).flags == OBCA_NO_RECOVERY
    assert Obstacles.from_list(parking_lot_obstacles(), geometric_start=True).flags == OBCA_GEOMETRIC_START
    assert Obstacles.from_list(parking_lot_obstacles(), recover=False, geometric_start=True).flags == 3


def test_invalid_config_is_rejected(L):
    h = ctypes.c_void_p()
    c = tracking_preset(40); c.horizon = 0
    assert L.ttmpc_create(ctypes.byref(c), 0, ctypes.byref(h)) == -22
    c = tracking_preset(40); c.R[0] = -1.0
    assert L.ttmpc_create(ctypes.byref(c), 0, ctypes.byref(h)) == -22
    c = tracking_preset(40); c.x_lb[3] = 2.0
    assert L.ttmpc_create(ctypes.byref(c), 0, ctypes.byref(h)) == -22


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    h = ctypes.c_void_p()
    c = tracking_preset(40)
    assert L.ttmpc_create(ctypes.byref(c), 0, ctypes.byref(h)) == _lib.E_NODEV
    from car_trailer_mpc_b200 import BatchSolver
    with pytest.raises(_lib.TTMPCError):
        BatchSolver(c)


def test_product_never_imports_the_oracle():
    """only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/ -- the package must not."""
    pkg = os.path.join(ROOT, "car_trailer_mpc_b200")
    pat = re.compile(r"^\s*(import|from)\s+\S*oracle|#\s*include.*oracle|CDLL\(.*oracle|dlopen\(.*oracle", re.I)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                for ln in open(os.path.join(dp, f)).read().splitlines():
                    assert not pat.search(ln), (f, ln)
