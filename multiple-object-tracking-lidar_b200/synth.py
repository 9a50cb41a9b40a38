"""Deterministic synthetic inputs for the five BASELINE.json configs.

Everything is generated on the host with numpy's PCG64 (seed = 0xB2000000 + 1000*config + frame_id), so the
oracle and the GPU library are fed identical bytes.  Points are N x 4 float32 (x, y, z, pad=1) -- the memory
layout of pcl::PointXYZ -- so `&cloud.points[0]` and these arrays are interchangeable.
"""
import numpy as np

SEED_BASE = 0xB2000000


def _rng(config, frame=0):
    return np.random.Generator(np.random.PCG64(SEED_BASE + 1000 * config + frame))


def _as_cloud(xyz):
    out = np.ones((len(xyz), 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


# ----------------------------------------------------------------------------------------------------
# c1: 64k points over a 40 m x 40 m occupancy grid with 40 boxes (removeStatic + clustering, tol 0.3)
# ----------------------------------------------------------------------------------------------------
def make_map_c1(cells=800, resolution=0.05):
    """Trinary occupancy grid (-1 unknown / 0 free / 100 occupied): unknown margin, wall band, free inside."""
    occ = np.zeros((cells, cells), dtype=np.int8)
    m_unknown, m_wall = 20, 30  # 1.0 m unknown margin, then a 0.5 m wall band
    occ[:m_wall, :] = 100
    occ[-m_wall:, :] = 100
    occ[:, :m_wall] = 100
    occ[:, -m_wall:] = 100
    occ[:m_unknown, :] = -1
    occ[-m_unknown:, :] = -1
    occ[:, :m_unknown] = -1
    occ[:, -m_unknown:] = -1
    origin = (-cells * resolution / 2.0, -cells * resolution / 2.0, 0.0)
    return occ, np.float32(resolution), origin


def _sample_box_surface(rng, centre, edge, height, n):
    """n points on the four vertical faces and the top of an axis-aligned box."""
    ex, ey = edge
    areas = np.array([ey * height, ey * height, ex * height, ex * height, ex * ey])
    face = rng.choice(5, size=n, p=areas / areas.sum())
    u, v = rng.random(n), rng.random(n)
    x = np.where(face == 0, -ex / 2, np.where(face == 1, ex / 2, (u - 0.5) * ex))
    y = np.where(face == 2, -ey / 2, np.where(face == 3, ey / 2, np.where(face < 2, (u - 0.5) * ey, (v - 0.5) * ey)))
    z = np.where(face == 4, height, v * height)
    return np.stack([x + centre[0], y + centre[1], z], axis=1)


def make_frame_c1(n_points=65536, n_boxes=40, frame=0, box_shift=(0.0, 0.0)):
    """Returns (cloud N x 4, box_centres).  ~20 % of the points sit on boxes in free space (per-box counts
    straddle min_cluster_size=5 and max_cluster_size=300), ~1 % are isolated noise, the rest lie on the wall
    band / unknown margin and are removed by removeStatic (static_tolarance 2)."""
    rng = _rng(1, frame)
    layout = _rng(1, 999)  # box layout is frame independent; frames differ by sampling noise and box_shift
    lattice = [(i, j) for i in range(7) for j in range(6)]
    pick = layout.permutation(len(lattice))[:n_boxes]
    parts, centres = [], []
    for b in pick:
        i, j = lattice[b]
        cx = -15.0 + 5.0 * i + layout.uniform(-1.0, 1.0) + box_shift[0]
        cy = -12.5 + 5.0 * j + layout.uniform(-1.0, 1.0) + box_shift[1]
        edge = layout.uniform(0.3, 1.5, size=2)
        height = layout.uniform(0.5, 2.0)
        cnt = int(layout.integers(3, 601))
        cnt = max(3, int(cnt * min(1.0, n_points / 65536.0)))  # smaller frames keep the same layout, thinner boxes
        parts.append(_sample_box_surface(rng, (cx, cy), edge, height, cnt))
        centres.append((cx, cy))
    n_noise = n_points // 100
    noise = np.stack([rng.uniform(-17.5, 17.5, n_noise), rng.uniform(-17.5, 17.5, n_noise), rng.uniform(0, 2, n_noise)], 1)
    parts.append(noise)
    n_static = n_points - sum(len(p) for p in parts)
    # static returns: on the wall band / unknown margin (|x| or |y| in [18.5, 20))
    side = rng.integers(0, 4, n_static)
    along = rng.uniform(-19.9, 19.9, n_static)
    depth = rng.uniform(18.6, 19.9, n_static)
    sx = np.where(side == 0, depth, np.where(side == 1, -depth, along))
    sy = np.where(side == 2, depth, np.where(side == 3, -depth, np.where(side < 2, along, along)))
    parts.append(np.stack([sx, sy, rng.uniform(0, 2.5, n_static)], 1))
    xyz = np.concatenate(parts)
    xyz = xyz[rng.permutation(len(xyz))]
    return _as_cloud(xyz), np.array(centres)


C1_PARAMS = dict(cluster_tolerance=0.3, min_cluster_size=5, max_cluster_size=300, static_tolerance=2)


# ----------------------------------------------------------------------------------------------------
# c2 / c3: ray-cast LiDAR frames (boxes + enclosing walls, no ground returns)
# ----------------------------------------------------------------------------------------------------
def _make_scene(rng, n_boxes, half_extent):
    """Axis-aligned boxes scattered in a square room [-half, half]^2; returns (lo, hi) arrays B x 3."""
    lo, hi = [], []
    tries = 0
    centres = []
    while len(lo) < n_boxes and tries < 100000:
        tries += 1
        c = rng.uniform(-half_extent + 4, half_extent - 4, size=2)
        if np.hypot(*c) < 3.0:
            continue
        if centres and np.min(np.hypot(*(np.array(centres) - c).T)) < 4.0:
            continue
        e = rng.uniform(0.5, 2.5, size=2)
        h = rng.uniform(0.8, 3.0)
        centres.append(c)
        lo.append([c[0] - e[0] / 2, c[1] - e[1] / 2, 0.0])
        hi.append([c[0] + e[0] / 2, c[1] + e[1] / 2, h])
    return np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)


def _raycast(origin, elev, az, lo, hi, half_extent):
    """Nearest hit range for the (beam, azimuth) ray fan of one sensor against the boxes and the room's
    walls (hit from inside).  Each box is only tested against the azimuth columns its footprint can span."""
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    ca, sa = np.cos(az)[None, :], np.sin(az)[None, :]
    d = np.stack([ce * ca, ce * sa, se * np.ones_like(ca)], axis=-1)  # beams x az x 3
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        tx = np.where(d[..., 0] > 0, (half_extent - origin[0]) * inv[..., 0], (-half_extent - origin[0]) * inv[..., 0])
        ty = np.where(d[..., 1] > 0, (half_extent - origin[1]) * inv[..., 1], (-half_extent - origin[1]) * inv[..., 1])
    t = np.minimum(np.where(np.isfinite(tx), tx, np.inf), np.where(np.isfinite(ty), ty, np.inf))
    n_az = len(az)
    step = 2 * np.pi / n_az
    for b in range(len(lo)):
        cx = np.array([lo[b, 0], lo[b, 0], hi[b, 0], hi[b, 0]]) - origin[0]
        cy = np.array([lo[b, 1], hi[b, 1], lo[b, 1], hi[b, 1]]) - origin[1]
        ang = np.arctan2(cy, cx)
        mid = np.arctan2(cy.mean(), cx.mean())
        rel = (ang - mid + np.pi) % (2 * np.pi) - np.pi
        a0, a1 = mid + rel.min() - 2 * step, mid + rel.max() + 2 * step
        k0 = int(np.floor((a0 - az[0]) / step))
        k1 = int(np.ceil((a1 - az[0]) / step))
        cols = np.arange(k0, k1 + 1) % n_az
        dd, ii = d[:, cols, :], inv[:, cols, :]
        with np.errstate(invalid="ignore"):
            t1 = (lo[b] - origin) * ii
            t2 = (hi[b] - origin) * ii
        tn = np.nanmax(np.minimum(t1, t2), axis=-1)
        tf = np.nanmin(np.maximum(t1, t2), axis=-1)
        hit = (tf >= tn) & (tn > 0)
        t[:, cols] = np.where(hit, np.minimum(t[:, cols], tn), t[:, cols])
    return d.reshape(-1, 3), t.reshape(-1)


class LidarScene:
    """A static scene ray-cast once; frames are cheap variations (azimuth offset + fresh range noise)."""

    def __init__(self, config, n_beams, n_azimuth, n_boxes, half_extent, sensors=((0.0, 0.0, 1.5),), elev=(-25.0, 15.0)):
        rng = _rng(config, 900)
        self.config = config
        self.half = half_extent
        self.lo, self.hi = _make_scene(rng, n_boxes, half_extent)
        self.sensors = np.array(sensors, dtype=np.float64)
        self.n_beams, self.n_az = n_beams, n_azimuth
        self.elev = np.deg2rad(np.linspace(elev[0], elev[1], n_beams))

    def frame(self, frame=0, noise_sigma=0.02, n_points=None):
        rng = _rng(self.config, frame)
        clouds = []
        for origin in self.sensors:
            az0 = rng.uniform(0, 2 * np.pi)
            az = az0 + np.arange(self.n_az) * (2 * np.pi / self.n_az)
            dirs, t = _raycast(origin, self.elev, az, self.lo, self.hi, self.half)
            t = t + rng.normal(0.0, noise_sigma, size=t.shape)
            clouds.append(origin[None, :] + dirs * t[:, None])
        xyz = np.concatenate(clouds)
        if n_points is not None:
            xyz = xyz[:n_points]
        return _as_cloud(xyz)


def scene_c2():
    """c2: one 2^20-point frame, 128 beams x 8192 azimuths, ~300 boxes + walls, tol 0.5, min 5, max 100000."""
    return LidarScene(2, n_beams=128, n_azimuth=8192, n_boxes=300, half_extent=60.0)


C2_PARAMS = dict(cluster_tolerance=0.5, min_cluster_size=5, max_cluster_size=100000)


def scene_c3():
    """c3: 130,000-point frames = four merged 32-beam scans (4 x 32 x 1016 = 130,048, truncated), tol 0.3."""
    sensors = ((1.0, 0.6, 1.6), (1.0, -0.6, 1.6), (-1.0, 0.6, 1.6), (-1.0, -0.6, 1.6))
    return LidarScene(3, n_beams=32, n_azimuth=1016, n_boxes=120, half_extent=40.0, sensors=sensors, elev=(-15.0, 15.0))


C3_PARAMS = dict(cluster_tolerance=0.3, min_cluster_size=5, max_cluster_size=100000)
C3_POINTS = 130000


# ----------------------------------------------------------------------------------------------------
# c4: dense 4M-point frame, 2,000 Gaussian blobs on a jittered 2-D lattice (cell-occupancy stress)
# ----------------------------------------------------------------------------------------------------
def make_frame_c4(n_points=1 << 22, n_blobs=2000, sigma=0.4, frame=0):
    rng = _rng(4, frame)
    side = int(np.ceil(np.sqrt(n_blobs)))
    ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)[:n_blobs]
    centres = ij * 7.0 + rng.uniform(-0.5, 0.5, size=(n_blobs, 2))  # 7 m lattice, +-0.5 m jitter: >= 6 m apart
    centres = np.concatenate([centres - centres.mean(0), rng.uniform(0.8, 1.6, size=(n_blobs, 1))], axis=1)
    owner = np.arange(n_points) % n_blobs
    xyz = centres[owner] + np.clip(rng.normal(0.0, sigma, size=(n_points, 3)), -3 * sigma, 3 * sigma)
    xyz = xyz[rng.permutation(n_points)]
    return _as_cloud(xyz)


C4_TOLERANCES = (0.1, 0.2, 0.3, 0.5, 0.7, 1.0)
C4_PARAMS = dict(min_cluster_size=5, max_cluster_size=1000000)


# ----------------------------------------------------------------------------------------------------
# c5: track rings for the batched IHGP step (T tracks, L = data_length samples each)
# ----------------------------------------------------------------------------------------------------
def make_rings_c5(n_tracks=1000, data_length=40, dt=0.1, frame=0):
    """T x L x 4 float32 rings (x, y, z=0, intensity=time): smooth motion <= 1.5 m/s plus centroid jitter."""
    rng = _rng(5, frame)
    t = np.arange(data_length) * dt
    p0 = rng.uniform(-15, 15, size=(n_tracks, 2))
    v = rng.uniform(-1.5, 1.5, size=(n_tracks, 2))
    w = rng.uniform(-0.5, 0.5, size=(n_tracks, 1))
    x = p0[:, :1] + v[:, :1] * t + 0.3 * np.sin(w * t * 2 * np.pi)
    y = p0[:, 1:] + v[:, 1:] * t + 0.3 * np.cos(w * t * 2 * np.pi)
    jitter = rng.normal(0, 0.01, size=(n_tracks, data_length, 2))
    rings = np.zeros((n_tracks, data_length, 4), dtype=np.float32)
    rings[:, :, 0] = x + jitter[:, :, 0]
    rings[:, :, 1] = y + jitter[:, :, 1]
    rings[:, :, 3] = t[None, :] + 100.0
    return rings


# launch/simTracker.launch:13-37 (the values the reference intends to run with)
LAUNCH = dict(frequency=10.0, cluster_tolerance=0.15, min_cluster_size=5, max_cluster_size=300, voxel_leaf_size=0.1,
              static_tolarance=2, id_threshold=0.4, lpf_tau=0.03, logSigma2=-5.5, logMagnSigma2=-3.5, logLengthScale=0.75,
              data_length=40)
