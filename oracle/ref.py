"""ctypes front end of oracle/_ref/libmot_ref.so -- TEST INFRASTRUCTURE ONLY.

libmot_ref.so is the REFERENCE's own tracker and IHGP source compiled where it lies under /root/reference against the
stand-in headers of oracle/shim (see oracle/Makefile target `_ref` and oracle/ref_harness.cpp).  It exists in the
development container only (and, prebuilt, on a GPU box it was shipped to); nothing in the product imports it.  Its two
jobs: checking the oracle's restatements against the code they restate (tests/test_ref_pin.py) and generating the golden
vectors under tests/golden/ (tests/golden/make_ref_fixtures.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libmot_ref.so")
REFERENCE_ROOT = os.environ.get("MOT_REFERENCE_ROOT", "/root/reference")
_LIB = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


def can_build():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "src", "multiple_object_tracking_lidar.cpp"))


def build(force=False):
    """Compiles the reference sources (only possible where /root/reference exists).  Returns the .so path or None."""
    if not can_build():
        return SO if os.path.exists(SO) else None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE_ROOT, "_ref"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return SO


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.POINTER(C.c_char_p), _f64p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_now.argtypes = [C.c_double]
        L.ref_set_map.argtypes = [C.c_void_p, _i8p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, _f64p]
        L.ref_yaw_from_quat.restype = C.c_float
        L.ref_yaw_from_quat.argtypes = [C.c_void_p, _f64p]
        L.ref_remove_static.restype = C.c_int64
        L.ref_remove_static.argtypes = [C.c_void_p, _f32p, C.c_int64, _f32p]
        L.ref_get_centroid.argtypes = [C.c_void_p, _f32p, C.c_int64, _i32p, _i32p, C.c_int, C.c_double, C.c_double, _f32p]
        L.ref_cluster_point_cloud.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_double, _f32p, C.c_int]
        L.ref_cloud_callback.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_double, _i32p, _f32p, C.c_int]
        L.ref_tracks.argtypes = [C.c_void_p, _i32p, _f32p, _f64p, C.c_int]
        L.ref_next_obj_num.argtypes = [C.c_void_p]
        L.ref_ihgp_constants.restype = None
        L.ref_ihgp_constants.argtypes = [C.c_double, _f64p, _f64p]
        L.ref_call_ihgp.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f64p, _f32p]
        _LIB = L
    return _LIB


def ihgp_constants(dt, sigma2, magn_sigma2, length_scale):
    """A[4], AKHA[4], K[2], G[4], S, lambda of an InfiniteHorizonGP built as registerNewObstacle builds it."""
    out = np.zeros(16)
    lib().ref_ihgp_constants(float(dt), np.array([sigma2, magn_sigma2, length_scale], dtype=np.float64), out)
    return out


class Reference:
    """One ObstacleTrack object of the reference.  params: launch-file names -> values (floats)."""

    def __init__(self, **params):
        names = (C.c_char_p * len(params))(*[k.encode() for k in params])
        vals = np.array([float(v) for v in params.values()], dtype=np.float64) if params else np.zeros(1)
        self.L = int(params.get("data_length", 10))
        self.h = lib().ref_create(names, vals, len(params))
        assert self.h, "ObstacleTrack::initialize failed"

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_map(self, occ, resolution, origin_xy, quat_xyzw=(0, 0, 0, 1)):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        H, W = occ.shape
        lib().ref_set_map(self.h, occ, W, H, np.float32(resolution), float(origin_xy[0]), float(origin_xy[1]),
                          np.ascontiguousarray(quat_xyzw, dtype=np.float64))

    def yaw_from_quat(self, quat_xyzw):
        return float(lib().ref_yaw_from_quat(self.h, np.ascontiguousarray(quat_xyzw, dtype=np.float64)))

    def remove_static(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        out = np.zeros((max(len(pts), 1), 4), dtype=np.float32)
        k = lib().ref_remove_static(self.h, pts, len(pts), out)
        if k == -2:
            raise IndexError("the reference indexed the map out of bounds (undefined behaviour, MOT.cpp:686)")
        return out[:k].copy()

    def get_centroid(self, pts, off, idx, stamp=0.0, time_init=0.0):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        K = len(off) - 1
        out = np.zeros((max(K, 1), 4), dtype=np.float32)
        idx = np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32)
        k = lib().ref_get_centroid(self.h, pts, len(pts), np.ascontiguousarray(off, np.int32), idx, K, float(stamp), float(time_init), out)
        return out[:k].copy()

    def cluster_point_cloud(self, pts, stamp, cap=4096):
        """clusterPointCloud (MOT.cpp:438-505) on one frame; returns the K x 4 centroids."""
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        out = np.zeros((cap, 4), dtype=np.float32)
        k = lib().ref_cluster_point_cloud(self.h, pts, len(pts), float(stamp), out, cap)
        if k == -2:
            raise IndexError("the reference indexed the map out of bounds (undefined behaviour, MOT.cpp:686)")
        return out[:k].copy()

    def cloud_callback(self, pts, stamp, cap=4096):
        """One frame through cloudCallback.  Returns None if nothing was published, else (ids, pos_vel T x 8)."""
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        ids = np.zeros(cap, dtype=np.int32)
        pv = np.zeros((cap, 8), dtype=np.float32)
        t = lib().ref_cloud_callback(self.h, pts, len(pts), float(stamp), ids, pv, cap)
        if t == -2:
            raise IndexError("the reference indexed the map out of bounds (undefined behaviour, MOT.cpp:686)")
        if t == 0:
            return None
        return ids[:t].copy(), pv[:t].copy()

    def tracks(self, cap=4096):
        ids = np.zeros(cap, dtype=np.int32)
        rings = np.zeros((cap, self.L, 4), dtype=np.float32)
        m = np.zeros((cap, 4))
        t = lib().ref_tracks(self.h, ids, rings, m, cap)
        return ids[:t].copy(), rings[:t].copy(), m[:t].copy()

    def next_obj_num(self):
        return int(lib().ref_next_obj_num(self.h))

    def call_ihgp(self, rings, m_state=None):
        """callIHGP over all T tracks (registered on first use).  rings T x L x 4; returns (pos_vel T x 8, m T x 4)."""
        rings = np.ascontiguousarray(rings, dtype=np.float32)
        T, L, _ = rings.shape
        assert L == self.L
        m = np.zeros((T, 4)) if m_state is None else np.ascontiguousarray(m_state, dtype=np.float64).copy()
        pv = np.zeros((T, 8), dtype=np.float32)
        t = lib().ref_call_ihgp(self.h, rings, T, int(m_state is not None), m, pv)
        assert t == T, "track count changed between calls"
        return pv, m
