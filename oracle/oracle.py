"""ctypes front end of the CPU oracle (oracle/mot_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.  Parity status: removeStatic, getCentroid and
IHGP are pinned to the reference's own sources (oracle/_ref, tests/test_ref_pin.py); the PCL pieces (clustering,
VoxelGrid, fromROSMsg) are restated and unpinned -- see the header of mot_oracle.cpp and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libmot_oracle.so")
    src = os.path.join(_HERE, "mot_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmot_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmot_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_yaw_from_quat.restype = C.c_float
        L.orc_yaw_from_quat.argtypes = [_f64p]
        L.orc_remove_static.restype = C.c_int64
        L.orc_remove_static.argtypes = [_f32p, C.c_int64, _i8p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double,
                                        _f64p, C.c_int, _u8p]
        L.orc_labels_bruteforce.restype = None
        L.orc_labels_bruteforce.argtypes = [_f32p, C.c_int, C.c_float, _i32p]
        L.orc_labels_grid.restype = None
        L.orc_labels_grid.argtypes = [_f32p, C.c_int, C.c_float, _i32p]
        L.orc_cluster_kdtree.restype = C.c_int
        L.orc_cluster_kdtree.argtypes = [_f32p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.orc_csr_from_labels.restype = C.c_int
        L.orc_csr_from_labels.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.orc_get_centroid.restype = None
        L.orc_get_centroid.argtypes = [_f32p, _i32p, _i32p, C.c_int, C.c_double, _f32p]
        L.orc_cluster_stats.restype = None
        L.orc_cluster_stats.argtypes = [_f32p, _i32p, _i32p, C.c_int, _f32p]
        L.orc_ihgp_setup.restype = None
        L.orc_ihgp_setup.argtypes = [C.c_double, _f64p, _f64p]
        L.orc_ihgp_step.restype = None
        L.orc_ihgp_step.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, _f64p, _f64p, _f64p, _f32p]
        L.orc_voxel_grid.restype = C.c_int64
        L.orc_voxel_grid.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, C.c_float, _f32p]
        _LIB = L
    return _LIB


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "points are N x 4 float32 (x, y, z, pad) == pcl::PointXYZ"
    return a


def yaw_from_quat(q_xyzw):
    return float(lib().orc_yaw_from_quat(np.ascontiguousarray(q_xyzw, dtype=np.float64)))


def remove_static(pts, occ, resolution, origin_xy, quat_xyzw=(0, 0, 0, 1), static_tolerance=2):
    """Returns the kept points (order preserved) and the keep mask.  MOT.cpp:664-706."""
    pts = _pts(pts)
    occ = np.ascontiguousarray(occ, dtype=np.int8)
    H, W = occ.shape
    keep = np.zeros(len(pts), dtype=np.uint8)
    lib().orc_remove_static(pts, len(pts), occ, W, H, np.float32(resolution), float(origin_xy[0]), float(origin_xy[1]),
                            np.ascontiguousarray(quat_xyzw, dtype=np.float64), int(static_tolerance), keep)
    return pts[keep.astype(bool)], keep


def labels_bruteforce(pts, tol):
    pts = _pts(pts)
    lab = np.empty(len(pts), dtype=np.int32)
    lib().orc_labels_bruteforce(pts, len(pts), np.float32(tol), lab)
    return lab


def labels_grid(pts, tol):
    pts = _pts(pts)
    lab = np.empty(len(pts), dtype=np.int32)
    lib().orc_labels_grid(pts, len(pts), np.float32(tol), lab)
    return lab


def cluster_kdtree(pts, tol, min_size, max_size, build_twice=True):
    """The reference path: PCL EuclideanClusterExtraction restated.  Returns (offsets[K+1], indices)."""
    pts = _pts(pts)
    m = len(pts)
    off = np.zeros(m + 2, dtype=np.int32)
    idx = np.zeros(max(m, 1), dtype=np.int32)
    k = lib().orc_cluster_kdtree(pts, m, np.float32(tol), int(min_size), int(max_size), int(build_twice), off, idx)
    return off[: k + 1].copy(), idx[: off[k]].copy()


def csr_from_labels(lab, min_size, max_size):
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    m = len(lab)
    off = np.zeros(m + 2, dtype=np.int32)
    idx = np.zeros(max(m, 1), dtype=np.int32)
    k = lib().orc_csr_from_labels(lab, m, int(min_size), int(max_size), off, idx)
    return off[: k + 1].copy(), idx[: off[k]].copy()


def labels_from_csr(off, idx, m):
    """Canonical per-point label (min index of the point's kept cluster, -1 if dropped)."""
    lab = np.full(m, -1, dtype=np.int32)
    for c in range(len(off) - 1):
        seg = idx[off[c]:off[c + 1]]
        lab[seg] = seg.min() if len(seg) else -1
    return lab


def get_centroid(pts, off, idx, stamp_minus_time_init=0.0):
    pts = _pts(pts)
    K = len(off) - 1
    out = np.zeros((max(K, 1), 4), dtype=np.float32)
    lib().orc_get_centroid(pts, np.ascontiguousarray(off, np.int32), np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32),
                           K, float(stamp_minus_time_init), out)
    return out[:K]


def cluster_stats(pts, off, idx):
    pts = _pts(pts)
    K = len(off) - 1
    out = np.zeros((max(K, 1), 10), dtype=np.float32)
    lib().orc_cluster_stats(pts, np.ascontiguousarray(off, np.int32), np.ascontiguousarray(idx, np.int32) if len(idx) else np.zeros(1, np.int32), K, out)
    return out[:K]


def ihgp_setup(dt, sigma2, magn_sigma2, length_scale):
    """Returns 16 doubles: A[4], AKHA[4], K[2], G[4], S, lambda (row-major 2x2)."""
    consts = np.zeros(16, dtype=np.float64)
    lib().orc_ihgp_setup(float(dt), np.array([sigma2, magn_sigma2, length_scale], dtype=np.float64), consts)
    return consts


def ihgp_step(rings, m_state, dt_gp, lpf_tau, consts_x, consts_y):
    """rings T x L x 4 float32; m_state T x 4 float64 (updated in place).  Returns pos_vel T x 8 float32."""
    rings = np.ascontiguousarray(rings, dtype=np.float32)
    T, L, four = rings.shape
    assert four == 4 and m_state.dtype == np.float64 and m_state.shape == (T, 4) and m_state.flags.c_contiguous
    out = np.zeros((T, 8), dtype=np.float32)
    lib().orc_ihgp_step(rings, T, L, np.float32(dt_gp), np.float32(lpf_tau), consts_x, consts_y, m_state, out)
    return out


def voxel_grid(pts, leaf_xyz):
    pts = _pts(pts)
    out = np.zeros((max(len(pts), 1), 4), dtype=np.float32)
    k = lib().orc_voxel_grid(pts, len(pts), np.float32(leaf_xyz[0]), np.float32(leaf_xyz[1]), np.float32(leaf_xyz[2]), out)
    return out[:k].copy()


def load_map_server_trinary(pgm_path, yaml_path):
    """ROS map_server trinary rule restated: occ=(255-pix)/255; > occupied_thresh -> 100;
    < free_thresh -> 0; else -1; image row j -> grid row H-1-j.  Returns (occ int8 HxW, resolution, origin xyz)."""
    import yaml

    with open(yaml_path) as f:
        meta = yaml.safe_load(f)
    with open(pgm_path, "rb") as f:
        data = f.read()
    # P5 header: magic, width, height, maxval separated by whitespace, '#' comments allowed
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        s = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        toks.append(data[s:pos])
    pos += 1
    assert toks[0] == b"P5"
    W, H, maxval = int(toks[1]), int(toks[2]), int(toks[3])
    pix = np.frombuffer(data, dtype=np.uint8, count=W * H, offset=pos).reshape(H, W).astype(np.float64)
    if meta.get("negate", 0):
        occp = pix / 255.0
    else:
        occp = (255.0 - pix) / 255.0
    occ = np.full((H, W), -1, dtype=np.int8)
    occ[occp > meta["occupied_thresh"]] = 100
    occ[occp < meta["free_thresh"]] = 0
    occ = occ[::-1].copy()  # image row j -> grid row H-1-j
    return occ, float(meta["resolution"]), [float(v) for v in meta["origin"]]


def unpack_pointcloud2(data, n_points, point_step, off_xyz, is_bigendian=False, drop_nonfinite=False):
    """pcl::fromROSMsg restated for the x/y/z FLOAT32 fields of a sensor_msgs/PointCloud2 (reference MOT.cpp:448-449):
    a field-mapped copy into pcl::PointXYZ (pad = 1); optionally pcl::removeNaNFromPointCloud on top."""
    raw = np.frombuffer(np.ascontiguousarray(data, dtype=np.uint8).tobytes(), dtype=np.uint8).reshape(n_points, point_step)
    out = np.ones((n_points, 4), dtype=np.float32)
    dt = np.dtype(">f4" if is_bigendian else "<f4")
    for d, off in enumerate(off_xyz):
        out[:, d] = np.ascontiguousarray(raw[:, off:off + 4]).view(dt).reshape(n_points).astype(np.float32)
    if drop_nonfinite:
        out = out[np.isfinite(out[:, :3]).all(axis=1)]
    return out


def obstacle_table(pos_vel, ids):
    """publishObstacles (reference MOT.cpp:253-295): id, radius 0.3, polygon point = position, twist.linear = velocity,
    velocity covariance diagonal (.1, .1, 1e9, 1e9, 1e9, .1).  Returns a T x 12 float64 array (id first)."""
    T = len(pos_vel)
    out = np.zeros((T, 12))
    out[:, 0] = ids
    out[:, 1] = np.float32(0.3)
    out[:, 2:4] = pos_vel[:, 0:2]
    out[:, 4:6] = pos_vel[:, 4:6]
    out[:, 6:12] = np.array([.1, .1, 1e9, 1e9, 1e9, .1], dtype=np.float32)
    return out
