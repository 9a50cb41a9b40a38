"""mot_b200 -- Python (ctypes) host-side mirror of the B200-native tracker hot path.

The product is `libmot_b200.so` (CUDA kernels for sm_100a behind the C ABI in include/mot_b200.h) plus the
header-only C++ adapter include/mot_b200_pcl.hpp.  This module only binds the C ABI so that the parity tests and
bench.py can drive it with numpy arrays; it mirrors the reference's call sites in
ObstacleTrack::clusterPointCloud / callIHGP (reference src/multiple_object_tracking_lidar.cpp:444-505, 621-662):

    reference call site                               here
    ------------------------------------------------  -----------------------------------------
    mapCallback(OccupancyGrid)            :235-251     Tracker.set_map(occ, resolution, origin, quat, static_tolarance)
    cloud_2 = removeStatic(cloud_1)       :461         Tracker.remove_static(cloud)
    ec.setClusterTolerance/Min/Max        :481-483     Tracker.set_cluster_params(tol, min, max)
    ec.extract(cluster_indices)           :488         Tracker.extract(cloud) -> (offsets, indices)  [CSR of PointIndices]
    getCentroid(cluster_indices, cloud..) :491         Tracker.get_centroid(stamp_minus_time_init)
    callIHGP(this_objIDs)                 :223,:621    Tracker.ihgp_step(rings, m_state)

There is no CPU fallback: if the shared library is missing or no CUDA device is present every call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("MOT_B200_LIB", os.path.join(_HERE, "libmot_b200.so"))  # override: A/B runs of two builds
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]

MOT_OK = 0
MOT_WARN_TRACKS_FULL = 1
MOT_WARN_VOXEL_OVERFLOW = 2
ERRORS = {-1: "MOT_ERR_INVALID", -2: "MOT_ERR_CUDA", -3: "MOT_ERR_CAPACITY", -4: "MOT_ERR_NO_MAP", -5: "MOT_ERR_STATE",
          -6: "MOT_ERR_NONFINITE"}

OBSTACLE_DTYPE = np.dtype([("id", np.int32), ("radius", np.float32), ("x", np.float32), ("y", np.float32), ("vx", np.float32), ("vy", np.float32),
                           ("vel_cov", np.float32, 6)])
STAT_DTYPE = np.dtype([("count", np.int32), ("mean", np.float32, 3), ("bbox_min", np.float32, 3), ("bbox_max", np.float32, 3)])


class MotError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("remove_static_ms", "grid_build_ms", "union_find_ms", "cluster_table_ms", "reduce_ms", "total_ms")]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + [os.path.join(_ROOT, "include", "mot_b200.h")]


def build_lib(force=False, verbose=False):
    """Compile csrc/mot_b200.cu for sm_100a into libmot_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in _sources()):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "mot_b200.cu")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return LIB_PATH


_LIB = None
_f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i8 = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")

# every symbol include/mot_b200.h declares: name -> (restype, argtypes)
_SIZE = C.c_size_t
_H = C.c_void_p
SYMBOLS = {
    "mot_version": (C.c_char_p, []),
    "mot_create": (C.c_int, [C.c_int, _SIZE, _SIZE, C.POINTER(_H)]),
    "mot_destroy": (C.c_int, [_H]),
    "mot_last_error": (C.c_char_p, [_H]),
    "mot_set_map": (C.c_int, [_H, _i8, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, _f64, C.c_int]),
    "mot_set_cluster_params": (C.c_int, [_H, C.c_float, C.c_int, C.c_int]),
    "mot_remove_static": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.POINTER(_SIZE)]),
    "mot_unpack_pointcloud2": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, _SIZE,
                                         C.POINTER(_SIZE)]),
    "mot_tracks_reset": (C.c_int, [_H]),
    "mot_assoc_match_below": (C.c_double, [C.c_float]),
    "mot_tracks_step": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32)]),
    "mot_tracks_get": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, _SIZE, C.POINTER(C.c_int32)]),
    "mot_ihgp_step_obstacles": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mot_voxel_grid": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_float, C.c_float, C.c_float, C.c_void_p, _SIZE, C.POINTER(_SIZE)]),
    "mot_cluster": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.POINTER(C.c_int32)]),
    "mot_cluster_stats": (C.c_int, [_H, C.c_void_p, _SIZE]),
    "mot_get_centroid": (C.c_int, [_H, C.c_double, C.c_void_p, _SIZE]),
    "mot_frame": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_double, C.c_void_p, _SIZE, C.POINTER(_SIZE), C.c_void_p, _SIZE, C.c_void_p, _SIZE,
                            C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, _SIZE]),
    "mot_frame_device": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_int, C.c_int, C.c_double]),
    "mot_result_counts": (C.c_int, [_H, C.POINTER(_SIZE), C.POINTER(C.c_int32), C.POINTER(_SIZE)]),
    "mot_result_grid": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mot_result_counters": (C.c_int, [_H, _i32, C.c_int]),
    "mot_grid_plan": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mot_small_frames": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mot_small_frame_phases": (C.c_int, [_H, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), C.c_int]),
    "mot_debug_stats": (C.c_int, [_H, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), C.c_int]),
    "mot_result_device_ptrs": (C.c_int, [_H] + [C.POINTER(C.c_void_p)] * 5),
    "mot_result_fetch": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.c_void_p, C.c_void_p, _SIZE]),
    "mot_result_labels": (C.c_int, [_H, C.c_void_p, _SIZE]),
    "mot_last_timings": (C.c_int, [_H, C.POINTER(Timings)]),
    "mot_last_launches": (C.c_int, [_H]),
    "mot_set_profiling": (C.c_int, [_H, C.c_int]),
    "mot_profile_kernels": (C.c_int, []),
    "mot_profile_kernel_name": (C.c_char_p, [C.c_int]),
    "mot_profile_read": (C.c_int, [_H, _f32, _i32, C.c_int]),
    "mot_timer_start": (C.c_int, [_H]),
    "mot_timer_stop": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "mot_host_register": (C.c_int, [C.c_void_p, _SIZE]),
    "mot_host_unregister": (C.c_int, [C.c_void_p]),
    "mot_cluster_batch": (C.c_int, [_H, C.c_void_p, _i64, C.c_int, C.c_void_p, C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.POINTER(C.c_int32)]),
    "mot_cluster_batch_device": (C.c_int, [_H, C.c_void_p, _i64, C.c_int]),
    "mot_frame_batch": (C.c_int, [_H, C.c_void_p, C.c_int, _i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _SIZE, C.c_void_p, _SIZE,
                                  C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, _SIZE]),
    "mot_frame_batch_device": (C.c_int, [_H, C.c_void_p, _i64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mot_batch_run": (C.c_int, [C.POINTER(_H), C.c_int, C.c_void_p, C.c_int, _i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _SIZE,
                                C.c_void_p, _SIZE, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, _SIZE]),
    "mot_cluster_pointcloud2": (C.c_int, [_H, C.c_void_p, _SIZE, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_int, C.c_double,
                                          C.c_void_p, _SIZE, C.POINTER(_SIZE), C.c_void_p, _SIZE, C.c_void_p, _SIZE, C.POINTER(C.c_int32), C.c_void_p,
                                          C.c_void_p, _SIZE]),
    "mot_ihgp_configure": (C.c_int, [_H, C.c_double, C.c_float, _f64, _f64, C.c_int]),
    "mot_ihgp_constants": (C.c_int, [_H, C.c_int, _f64]),
    "mot_ihgp_step": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
}


def load():
    """Load libmot_b200.so.  Raises if it has not been built -- there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                               "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("clouds are N x 4 float32 arrays (x, y, z, pad) -- the layout of pcl::PointXYZ")
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Tracker:
    """One handle = one GPU stream worth of the reference's per-frame hot path (not re-entrant)."""

    def __init__(self, device=0, max_points=1 << 20, max_tracks=1024):
        self.lib = load()
        self.h = _H()
        self.max_points = int(max_points)
        rc = self.lib.mot_create(int(device), self.max_points, int(max_tracks), C.byref(self.h))
        if rc != MOT_OK:
            self.h = None
            raise MotError(rc, "mot_create failed (no CUDA device? there is no CPU fallback)")
        self.data_length = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.mot_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        """negative = error (raises); positive = completed with a warning (kept in self.last_warning)."""
        self.last_warning = rc if rc > 0 else 0
        if rc < 0:
            raise MotError(rc, self.lib.mot_last_error(self.h).decode())

    # -- configuration -------------------------------------------------------------------------------------
    def set_map(self, occ, resolution, origin_xy, quat_xyzw=(0.0, 0.0, 0.0, 1.0), static_tolarance=2):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        H, W = occ.shape
        self._ck(self.lib.mot_set_map(self.h, occ, W, H, np.float32(resolution), float(origin_xy[0]), float(origin_xy[1]),
                                      np.ascontiguousarray(quat_xyzw, dtype=np.float64), int(static_tolarance)))

    def set_cluster_params(self, cluster_tolerance, min_cluster_size, max_cluster_size):
        self._ck(self.lib.mot_set_cluster_params(self.h, np.float32(cluster_tolerance), int(min_cluster_size), int(max_cluster_size)))

    # -- per-frame calls (host buffers in, host buffers out) --------------------------------------------
    def remove_static(self, cloud):
        cloud = _cloud(cloud)
        out = np.empty_like(cloud)
        m = _SIZE(0)
        self._ck(self.lib.mot_remove_static(self.h, _ptr(cloud), len(cloud), _ptr(out), len(out), C.byref(m)))
        return out[: m.value]

    def unpack_pointcloud2(self, data, n_points, point_step, off_xyz, is_bigendian=False, drop_nonfinite=False):
        """pcl::fromROSMsg(*input, input_cloud) for a sensor_msgs/PointCloud2 payload (reference MOT.cpp:448-449)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty((max(n_points, 1), 4), dtype=np.float32)
        m = _SIZE(0)
        self._ck(self.lib.mot_unpack_pointcloud2(self.h, _ptr(data), int(n_points), int(point_step), int(off_xyz[0]), int(off_xyz[1]), int(off_xyz[2]),
                                                 int(is_bigendian), int(drop_nonfinite), _ptr(out), len(out), C.byref(m)))
        return out[: m.value]

    def voxel_grid(self, cloud, leaf_xyz):
        """vg.setLeafSize(lx, ly, lz); vg.filter(out)  (reference MOT.cpp:452-456)."""
        cloud = _cloud(cloud)
        out = np.empty_like(cloud)
        m = _SIZE(0)
        self._ck(self.lib.mot_voxel_grid(self.h, _ptr(cloud), len(cloud), np.float32(leaf_xyz[0]), np.float32(leaf_xyz[1]), np.float32(leaf_xyz[2]),
                                         _ptr(out), len(out), C.byref(m)))
        return out[: m.value]

    def extract(self, cloud):
        """ec.extract(): returns (cluster_offsets[K+1], point_indices) -- the CSR form of vector<PointIndices>."""
        cloud = _cloud(cloud)
        m = len(cloud)
        off = np.empty(m + 1, dtype=np.int32)
        idx = np.empty(max(m, 1), dtype=np.int32)
        k = C.c_int32(0)
        self._ck(self.lib.mot_cluster(self.h, _ptr(cloud), m, _ptr(off), len(off), _ptr(idx), len(idx), C.byref(k)))
        K = k.value
        return off[: K + 1].copy(), idx[: off[K]].copy()

    def cluster_stats(self):
        m, K, total = self.result_counts()
        st = np.zeros(max(K, 1), dtype=STAT_DTYPE)
        self._ck(self.lib.mot_cluster_stats(self.h, _ptr(st), len(st)))
        return st[:K]

    def get_centroid(self, stamp_minus_time_init=0.0):
        m, K, total = self.result_counts()
        out = np.zeros((max(K, 1), 4), dtype=np.float32)
        self._ck(self.lib.mot_get_centroid(self.h, float(stamp_minus_time_init), _ptr(out), len(out)))
        return out[:K]

    def frame(self, cloud, stamp_minus_time_init=0.0, want_kept=True, want_stats=True, want_centroids=True):
        """Fused removeStatic -> extract -> tables.  Returns a dict."""
        cloud = _cloud(cloud)
        n = len(cloud)
        kept = np.empty_like(cloud) if want_kept else None
        off = np.empty(n + 1, dtype=np.int32)
        idx = np.empty(max(n, 1), dtype=np.int32)
        # table capacity: a cluster has >= 1 point
        st = np.zeros(max(n, 1), dtype=STAT_DTYPE) if want_stats else None
        cen = np.zeros((max(n, 1), 4), dtype=np.float32) if want_centroids else None
        m = _SIZE(0)
        k = C.c_int32(0)
        self._ck(self.lib.mot_frame(self.h, _ptr(cloud), n, float(stamp_minus_time_init), _ptr(kept), n, C.byref(m), _ptr(off), len(off),
                                    _ptr(idx), len(idx), C.byref(k), _ptr(st), _ptr(cen), max(n, 1)))
        K, M = k.value, m.value
        return dict(m=M, K=K, kept=kept[:M] if want_kept else None, offsets=off[: K + 1].copy(), indices=idx[: off[K]].copy(),
                    stats=st[:K].copy() if want_stats else None, centroids=cen[:K].copy() if want_centroids else None)

    def frame_buffers(self, n):
        """Caller-owned output buffers for frame_into(), allocated (and touched) once -- what a C++ caller of mot_frame keeps."""
        return dict(n=int(n), kept=np.zeros((n, 4), dtype=np.float32), off=np.zeros(n + 1, dtype=np.int32), idx=np.zeros(max(n, 1), dtype=np.int32),
                    st=np.zeros(max(n, 1), dtype=STAT_DTYPE), cen=np.zeros((max(n, 1), 4), dtype=np.float32))

    def frame_into(self, cloud, bufs, stamp_minus_time_init=0.0):
        """mot_frame into preallocated buffers (no allocation, no copies of the results): returns (M, K); the arrays of `bufs` hold
        kept[:M], off[:K + 1], idx[:off[K]], st[:K], cen[:K]."""
        n = len(cloud)
        assert n <= bufs["n"] and cloud.dtype == np.float32 and cloud.flags.c_contiguous
        m = _SIZE(0)
        k = C.c_int32(0)
        self._ck(self.lib.mot_frame(self.h, _ptr(cloud), n, float(stamp_minus_time_init), _ptr(bufs["kept"]), bufs["n"], C.byref(m), _ptr(bufs["off"]),
                                    len(bufs["off"]), _ptr(bufs["idx"]), len(bufs["idx"]), C.byref(k), _ptr(bufs["st"]), _ptr(bufs["cen"]), len(bufs["st"])))
        return m.value, k.value

    # -- device-resident path ------------------------------------------------------------------------------
    def frame_device(self, d_ptr, n, do_remove_static=False, with_centroids=False, stamp_minus_time_init=0.0):
        self._ck(self.lib.mot_frame_device(self.h, C.c_void_p(int(d_ptr)), int(n), int(do_remove_static), int(with_centroids),
                                           float(stamp_minus_time_init)))

    def result_counts(self):
        m, k, t = _SIZE(0), C.c_int32(0), _SIZE(0)
        self._ck(self.lib.mot_result_counts(self.h, C.byref(m), C.byref(k), C.byref(t)))
        return m.value, k.value, t.value

    def result_grid(self):
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.mot_result_grid(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(fine_cells=a.value, coarse_cells=b.value, key_bits=c.value)

    def result_counters(self):
        out = np.zeros(16, dtype=np.int32)
        self._ck(self.lib.mot_result_counters(self.h, out, 16))
        return out

    def grid_plan(self, enable=None):
        """Switch the speculative grid plan (None: leave) and return (hits, misses) since the handle was created."""
        hits, misses = C.c_int(0), C.c_int(0)
        self._ck(self.lib.mot_grid_plan(self.h, -1 if enable is None else int(bool(enable)), C.byref(hits), C.byref(misses)))
        return hits.value, misses.value

    def small_frames(self, max_points=None):
        """Limit of the single-launch small-frame path (None: leave); returns (cluster CTAs, hits, misses)."""
        hits, misses = C.c_int(0), C.c_int(0)
        rc = self.lib.mot_small_frames(self.h, -1 if max_points is None else int(max_points), C.byref(hits), C.byref(misses))
        if rc < 0:
            self._ck(rc)
        return rc, hits.value, misses.value

    def small_frame_phases(self):
        """Nanoseconds spent in each phase of the last small-frame launch (dict phase -> ns)."""
        t = np.zeros(32, dtype=np.uint64)
        self._ck(self.lib.mot_small_frame_phases(self.h, t, 32))
        names = ("front: A rs+compact", "front: B hash insert", "front: C scan", "front: D place", "(gap)", "pairs", "tables: F flatten", "tables: G kept list",
                 "tables: H rank", "tables: I offsets", "tables: J segments", "tables: K order", "(gap)", "farthest pair", "finish")
        t = t.astype(np.int64)
        out = {f"{i:02d} {n}": int(t[i + 1] - t[i]) for i, n in enumerate(names) if t[i + 1] >= t[i] > 0}
        # sub-steps of the shared-memory union-find inside "tables: F": nanoseconds since the start of k_fs_tables
        sub = ("forest loaded and flattened", "cell pairs joined", "points flattened")
        out.update({f"F+ {n}": int(t[16 + i] - t[6]) for i, n in enumerate(sub) if t[16 + i] >= t[6] > 0})
        out["cell pairs"] = int(t[31])
        return out

    def debug_stats(self):
        """Union-find counters of a -DMOT_UF_STATS build (None for the product build)."""
        out = np.zeros(16, dtype=np.uint64)
        rc = self.lib.mot_debug_stats(self.h, out, 16)
        if rc < 0:
            self._ck(rc)
        names = ("finds", "hops", "unites", "cas_retry", "fine_pairs", "witness_tests", "cross_pairs", "root_skips", "cbox_rejects", "accepts",
                 "rejects", "local_pairs", "max_hops")
        return dict(zip(names, (int(v) for v in out))) if rc == 1 else None

    def result_device_ptrs(self):
        ps = [C.c_void_p() for _ in range(5)]
        self._ck(self.lib.mot_result_device_ptrs(self.h, *[C.byref(p) for p in ps]))
        return dict(zip(("kept", "offsets", "indices", "stats", "centroids"), [p.value for p in ps]))

    def result_fetch(self, want_kept=False, want_stats=False, want_centroids=False):
        M, K, total = self.result_counts()
        kept = np.empty((max(M, 1), 4), dtype=np.float32) if want_kept else None
        off = np.empty(K + 1, dtype=np.int32)
        idx = np.empty(max(total, 1), dtype=np.int32)
        st = np.zeros(max(K, 1), dtype=STAT_DTYPE) if want_stats else None
        cen = np.zeros((max(K, 1), 4), dtype=np.float32) if want_centroids else None
        self._ck(self.lib.mot_result_fetch(self.h, _ptr(kept), max(M, 1), _ptr(off), len(off), _ptr(idx), len(idx), _ptr(st), _ptr(cen), max(K, 1)))
        return dict(m=M, K=K, kept=kept[:M] if want_kept else None, offsets=off, indices=idx[:total], stats=st[:K] if want_stats else None,
                    centroids=cen[:K] if want_centroids else None)

    def result_labels(self):
        M, K, total = self.result_counts()
        lab = np.empty(max(M, 1), dtype=np.int32)
        self._ck(self.lib.mot_result_labels(self.h, _ptr(lab), len(lab)))
        return lab[:M]

    def last_launches(self):
        return int(self.lib.mot_last_launches(self.h))

    def set_profiling(self, on):
        self._ck(self.lib.mot_set_profiling(self.h, int(on)))

    def profile(self):
        """{kernel name: (total ms, launches)} accumulated since set_profiling(True)."""
        n = self.lib.mot_profile_kernels()
        ms = np.zeros(n, dtype=np.float32)
        cnt = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.mot_profile_read(self.h, ms, cnt, n))
        return {self.lib.mot_profile_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}

    def timer_start(self):
        self._ck(self.lib.mot_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self.lib.mot_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def timings(self):
        t = Timings()
        self._ck(self.lib.mot_last_timings(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    # -- batches ------------------------------------------------------------------------------------------------
    def extract_batch(self, clouds):
        """clouds: list of N_f x 4 arrays.  Returns (frame_cluster_offsets[F+1], cluster_offsets[K+1], indices)."""
        fo = np.zeros(len(clouds) + 1, dtype=np.int64)
        fo[1:] = np.cumsum([len(c) for c in clouds])
        allp = _cloud(np.concatenate([_cloud(c) for c in clouds])) if fo[-1] else np.zeros((0, 4), np.float32)
        total = int(fo[-1])
        fco = np.zeros(len(clouds) + 1, dtype=np.int32)
        off = np.empty(total + 1, dtype=np.int32)
        idx = np.empty(max(total, 1), dtype=np.int32)
        k = C.c_int32(0)
        self._ck(self.lib.mot_cluster_batch(self.h, _ptr(allp), fo, len(clouds), _ptr(fco), _ptr(off), len(off), _ptr(idx), len(idx), C.byref(k)))
        K = k.value
        return fco, off[: K + 1].copy(), idx[: off[K]].copy()

    def frame_batch(self, clouds, do_remove_static=False, stamps=None, want_stats=True, want_centroids=True, packed12=False):
        """mot_frame_batch: the whole per-frame path (removeStatic -> extract -> tables) on a list of clouds.  Returns a dict."""
        return frame_batch_call(self.lib, None, self, clouds, do_remove_static, stamps, want_stats, want_centroids, packed12)

    def frame_batch_device(self, d_ptr, frame_offsets, do_remove_static=False, with_centroids=False, stamps=None):
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        st = np.ascontiguousarray(stamps, dtype=np.float32) if stamps is not None else None
        self._ck(self.lib.mot_frame_batch_device(self.h, C.c_void_p(int(d_ptr)), fo, len(fo) - 1, int(do_remove_static), int(with_centroids), _ptr(st)))

    def cluster_pointcloud2(self, data, n_points, point_step, off_xyz, is_bigendian=False, voxel_leaf_size=0.0, do_remove_static=False,
                            stamp_minus_time_init=0.0, want_centroids=True):
        """ObstacleTrack::clusterPointCloud in one call (reference MOT.cpp:444-505).  Returns a dict like frame()."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = int(n_points)
        kept = np.empty((max(n, 1), 4), dtype=np.float32)
        off = np.empty(n + 1, dtype=np.int32)
        idx = np.empty(max(n, 1), dtype=np.int32)
        st = np.zeros(max(n, 1), dtype=STAT_DTYPE)
        cen = np.zeros((max(n, 1), 4), dtype=np.float32) if want_centroids else None
        m, k = _SIZE(0), C.c_int32(0)
        self._ck(self.lib.mot_cluster_pointcloud2(self.h, _ptr(data), n, int(point_step), int(off_xyz[0]), int(off_xyz[1]), int(off_xyz[2]),
                                                  int(is_bigendian), np.float32(voxel_leaf_size), int(do_remove_static), float(stamp_minus_time_init),
                                                  _ptr(kept), len(kept), C.byref(m), _ptr(off), len(off), _ptr(idx), len(idx), C.byref(k), _ptr(st),
                                                  _ptr(cen), max(n, 1)))
        K, M = k.value, m.value
        return dict(m=M, K=K, kept=kept[:M], offsets=off[: K + 1].copy(), indices=idx[: off[K]].copy(), stats=st[:K].copy(),
                    centroids=cen[:K].copy() if want_centroids else None)

    def cluster_batch_device(self, d_ptr, frame_offsets):
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        self._ck(self.lib.mot_cluster_batch_device(self.h, C.c_void_p(int(d_ptr)), fo, len(fo) - 1))

    # -- IHGP ---------------------------------------------------------------------------------------------------
    def ihgp_configure(self, dt, lpf_tau, hyp_x, hyp_y, data_length):
        """hyp = (sigma2, magnSigma2, lengthScale), already exponentiated (reference MOT.cpp:524-530)."""
        self._ck(self.lib.mot_ihgp_configure(self.h, float(dt), np.float32(lpf_tau), np.ascontiguousarray(hyp_x, dtype=np.float64),
                                             np.ascontiguousarray(hyp_y, dtype=np.float64), int(data_length)))
        self.data_length = int(data_length)

    def ihgp_constants(self, axis):
        out = np.zeros(16, dtype=np.float64)
        self._ck(self.lib.mot_ihgp_constants(self.h, int(axis), out))
        return out

    def ihgp_step(self, rings, m_state):
        """rings T x L x 4 float32 (x, y, z, time); m_state T x 4 float64, updated in place.  Returns T x 8 float32."""
        rings = np.ascontiguousarray(rings, dtype=np.float32)
        T, L, four = rings.shape
        if four != 4 or L != self.data_length:
            raise ValueError("rings must be T x data_length x 4")
        if m_state.dtype != np.float64 or m_state.shape != (T, 4) or not m_state.flags.c_contiguous:
            raise ValueError("m_state must be a contiguous T x 4 float64 array")
        out = np.zeros((T, 8), dtype=np.float32)
        self._ck(self.lib.mot_ihgp_step(self.h, _ptr(rings), T, _ptr(m_state), _ptr(out)))
        return out

    def tracks_reset(self):
        self._ck(self.lib.mot_tracks_reset(self.h))

    def tracks_step(self, centroids, now, id_threshold, frequency):
        """The association / lifecycle / callIHGP part of cloudCallback (reference MOT.cpp:176-233) on the device.
        centroids: K x 4 float32 (x, y, z, intensity).  Returns dict(produced, ids, pos_vel, obstacles, n_tracks)."""
        cen = np.ascontiguousarray(centroids, dtype=np.float32).reshape(-1, 4)
        K = len(cen)
        ids = np.full(max(K, 1), -1, dtype=np.int32)
        pv = np.zeros((max(K, 1), 8), dtype=np.float32)
        obs = np.zeros(max(K, 1), dtype=OBSTACLE_DTYPE)
        n, prod = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.mot_tracks_step(self.h, _ptr(cen), K, float(now), np.float32(id_threshold), np.float32(frequency), _ptr(ids), _ptr(pv),
                                          _ptr(obs), C.byref(n), C.byref(prod)))
        return dict(produced=bool(prod.value), ids=ids[:K], pos_vel=pv[:K], obstacles=obs[:K], n_tracks=n.value)

    def tracks_get(self):
        n = C.c_int32(0)
        self._ck(self.lib.mot_tracks_get(self.h, None, None, None, 1 << 30, C.byref(n)))
        T, L = n.value, self.data_length
        ids = np.zeros(max(T, 1), dtype=np.int32)
        rings = np.zeros((max(T, 1), L, 4), dtype=np.float32)
        m = np.zeros((max(T, 1), 4), dtype=np.float64)
        self._ck(self.lib.mot_tracks_get(self.h, _ptr(ids), _ptr(rings), _ptr(m), max(T, 1), C.byref(n)))
        return ids[:T], rings[:T], m[:T]

    def ihgp_step_obstacles(self, rings, m_state, track_ids=None):
        """ihgp_step + the packed ObstacleMsg table (publishObstacles, reference MOT.cpp:253-295)."""
        rings = np.ascontiguousarray(rings, dtype=np.float32)
        T = rings.shape[0]
        out = np.zeros((T, 8), dtype=np.float32)
        obs = np.zeros(max(T, 1), dtype=OBSTACLE_DTYPE)
        ids = np.ascontiguousarray(track_ids, dtype=np.int32) if track_ids is not None else None
        self._ck(self.lib.mot_ihgp_step_obstacles(self.h, _ptr(rings), T, _ptr(ids), _ptr(m_state), _ptr(out), _ptr(obs)))
        return out, obs[:T]


def frame_batch_call(lib, handles, trk, clouds, do_remove_static, stamps, want_stats, want_centroids, packed12):
    """Shared by Tracker.frame_batch (one handle) and batch_run (several handles / GPUs)."""
    F = len(clouds)
    fo = np.zeros(F + 1, dtype=np.int64)
    fo[1:] = np.cumsum([len(c) for c in clouds])
    total = int(fo[-1])
    allp = _cloud(np.concatenate([_cloud(c) for c in clouds])) if total else np.zeros((0, 4), np.float32)
    src = np.ascontiguousarray(allp[:, :3]) if packed12 else allp
    fko = np.zeros(F + 1, dtype=np.int32)
    fco = np.zeros(F + 1, dtype=np.int32)
    off = np.empty(total + 1, dtype=np.int32)
    idx = np.empty(max(total, 1), dtype=np.int32)
    st = np.zeros(max(total, 1), dtype=STAT_DTYPE) if want_stats else None
    cen = np.zeros((max(total, 1), 4), dtype=np.float32) if want_centroids else None
    stp = np.ascontiguousarray(stamps, dtype=np.float32) if stamps is not None else None
    k = C.c_int32(0)
    tail = (_ptr(src), 12 if packed12 else 16, fo, F, int(do_remove_static), _ptr(stp), _ptr(fko), _ptr(fco), _ptr(off), len(off), _ptr(idx), len(idx),
            C.byref(k), _ptr(st), _ptr(cen), max(total, 1))
    if handles is None:
        trk._ck(lib.mot_frame_batch(trk.h, *tail))
    else:
        arr = (_H * len(handles))(*[t.h for t in handles])
        handles[0]._ck(lib.mot_batch_run(arr, len(handles), *tail))
    K = k.value
    return dict(K=K, frame_kept_offsets=fko, frame_cluster_offsets=fco, offsets=off[: K + 1].copy(), indices=idx[: off[K]].copy(),
                stats=st[:K].copy() if want_stats else None, centroids=cen[:K].copy() if want_centroids else None)


def batch_run(trackers, clouds, do_remove_static=False, stamps=None, want_stats=True, want_centroids=True, packed12=False):
    """mot_batch_run: the frames of `clouds` sharded over the handles in `trackers` (one per GPU, or several per GPU)."""
    return frame_batch_call(trackers[0].lib, list(trackers), None, clouds, do_remove_static, stamps, want_stats, want_centroids, packed12)


from . import synth  # noqa: E402,F401
