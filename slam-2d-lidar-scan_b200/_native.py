"""ctypes binding of the C ABI in include/slam2d_b200.h (libslam2d_b200.so, built by __graft_entry__.build()).

There is no CPU fallback: if the shared library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLAM2D_B200_LIB") or os.path.join(_HERE, "libslam2d_b200.so")   # env: A/B builds of the kernels

MAX_BLUR_RADIUS = 8
MAX_BEAMS = 512

ST_WINDOW_OUTSIDE_MAP = 1
ST_INDEX_OUT_OF_FIELD = 2
ST_NAN_SCORE = 4
ST_SCAN_OUTSIDE_MAP = 8
ST_HEADING_MISSING = 16

c_double_p = C.POINTER(C.c_double)


class Geometry(C.Structure):
    _fields_ = [("G", C.c_int32), ("pitch", C.c_int32), ("K", C.c_int32), ("L", C.c_int32),
                ("numSpokes", C.c_int32), ("spokesStartIdx", C.c_int32),
                ("unit", C.c_double), ("mapX0", C.c_double), ("mapX1", C.c_double),
                ("mapY0", C.c_double), ("mapY1", C.c_double), ("fovHalf", C.c_double),
                ("maxRange", C.c_double), ("wallHalf", C.c_double),
                ("d_gridX", C.c_void_p), ("d_gridY", C.c_void_p), ("d_sector", C.c_void_p),
                ("d_radius", C.c_void_p), ("d_localAxis", C.c_void_p)]


class StageDesc(C.Structure):
    _fields_ = [("unitLength", C.c_double), ("logMiss", C.c_double), ("blurRadius", C.c_int32),
                ("blurW", C.c_double * (2 * MAX_BLUR_RADIUS + 1)), ("nHalf", C.c_int32), ("nTheta", C.c_int32),
                ("h_thetas", c_double_p), ("h_cos", c_double_p), ("h_sin", c_double_p)]


class MatcherDesc(C.Structure):
    _fields_ = [("windowRadius", C.c_double), ("coarse", StageDesc), ("fine", StageDesc)]


class MatchDebug(C.Structure):
    _fields_ = [("d_prob", C.c_void_p * 2), ("d_probDims", C.c_void_p * 2), ("d_vol", C.c_void_p * 2)]


# every symbol include/slam2d_b200.h declares: name -> (restype, argtypes)
_V, _I, _D, _Z = C.c_void_p, C.c_int32, C.c_double, C.c_size_t
SYMBOLS = {
    "slam_last_error": (C.c_char_p, []),
    "slam_version": (C.c_int, []),
    "slam_matcher_create": (C.c_int, [C.POINTER(Geometry), C.POINTER(MatcherDesc), C.POINTER(_V)]),
    "slam_matcher_destroy": (None, [_V]),
    "slam_matcher_workspace_bytes": (_Z, [_V]),
    "slam_matcher_workspace_bytes_n": (_Z, [_V, _I]),
    "slam_matcher_field_side": (C.c_int, [_V, C.c_int]),
    "slam_matcher_num_poses": (C.c_int, [_V, C.c_int]),
    "slam_match_scan": (C.c_int, [_V, _V, _I, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _Z, C.POINTER(MatchDebug), _V]),
    "slam_motion_priors": (C.c_int, [_I, _I, _D, _V, _V, _V, _V]),
    "slam_grid_init": (C.c_int, [C.POINTER(Geometry), _V, _I, _V]),
    "slam_update_grid": (C.c_int, [C.POINTER(Geometry), _V, _I, _V, _V, _V, _V, _Z, _V]),
    "slam_update_workspace_bytes": (_Z, [_I]),
    "slam_update_grid_slots": (C.c_int, [C.POINTER(Geometry), _V, _V, _I, _V, _V, _V, _V, _Z, _V]),
    "slam_match_scan_slots": (C.c_int, [_V, _V, _V, _I, _I, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _Z,
                                        C.POINTER(MatchDebug), _V]),
    "slam_copy_lattices": (C.c_int, [C.POINTER(Geometry), _V, _I, _V, _V, _V]),
    "slam_field_build": (C.c_int, [_V, _I, _V, _I, _V, _V, _V, _V, _V, _Z, _V]),
    "slam_correlate": (C.c_int, [_V, _I, _I, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _Z, _V]),
    "slam_blur_clamp": (C.c_int, [_V, _I, _I, _V, _I, _V, _V, _V]),
    "slam_propose_poses": (C.c_int, [_I, _V, _D, _D, _I, _D, _V, _V, _V, _V, _V, _V, _V]),
    "slam_finish_step": (C.c_int, [_I, _V, _V, _V, _V, _V, _V, _V]),
    "slam_normalize_weights": (C.c_int, [_I, _V, _V, _V]),
    "slam_step_trigger": (C.c_int, [_I, _V, _V, _V, _I, _V, _V]),
    "slam_status_reduce": (C.c_int, [_I, _V, _V, _V]),
    "slam_resample_indices": (C.c_int, [_I, _V, _V, _V, _V, _V]),
    "slam_gather_particles": (C.c_int, [C.POINTER(Geometry), _I, _V, _V, _V, _V, _V, _I, _V, _V]),
}
# test / profiling hooks exported by the library but outside the reference-facing surface
_EXTRA = {
    "slam_matcher_set_debug": (None, [_V, _V, C.c_int]),
    "slam_matcher_num_ctas": (C.c_int, [_V]),
    "slam_matcher_plan": (C.c_int, [_V, C.c_int, C.POINTER(C.c_int * 8)]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "slam-2d-lidar-scan_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in list(SYMBOLS.items()) + list(_EXTRA.items()):
    _fn = getattr(lib, _name)          # AttributeError here = the .so is stale: rebuild
    _fn.restype = _res
    _fn.argtypes = _args


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NativeError("libslam2d_b200: error %d: %s" % (rc, lib.slam_last_error().decode()))
