"""Hand-built known-answer inputs shared by the oracle tests (CPU) and the parity tests (GPU)."""
import numpy as np


def cloud(xyz):
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.ones((len(xyz), 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


def pair_at_distance(d):
    return cloud([[0, 0, 0], [d, 0, 0]])


def chain(n, step, start=(0.0, 0.0, 0.0), axis=0):
    p = np.zeros((n, 3), dtype=np.float64)
    p[:] = start
    p[:, axis] += np.arange(n) * step
    return cloud(p)


def blob(rng, centre, n, radius):
    return cloud(np.asarray(centre) + rng.uniform(-radius, radius, size=(n, 3)))


def kat_cases():
    """name -> (cloud, tol, min, max, expected number of clusters or None)"""
    rng = np.random.default_rng(1234)
    tol = np.float32(0.3)
    below = np.nextafter(tol, np.float32(0))
    cases = {}
    cases["empty"] = (cloud(np.zeros((0, 3))), 0.3, 1, 100, 0)
    cases["single_point"] = (cloud([[1, 2, 3]]), 0.3, 1, 100, 1)
    cases["single_point_min2"] = (cloud([[1, 2, 3]]), 0.3, 2, 100, 0)
    # strict '<': two points exactly tol apart are NOT joined; one ulp closer they are
    cases["pair_exactly_tol"] = (pair_at_distance(tol), float(tol), 1, 100, 2)
    cases["pair_tol_minus_ulp"] = (pair_at_distance(below), float(tol), 1, 100, 1)
    cases["pair_tol_plus"] = (pair_at_distance(np.nextafter(tol, np.float32(1))), float(tol), 1, 100, 2)
    # chain bridging two blobs
    a = blob(rng, (0, 0, 0), 40, 0.1)
    b = blob(rng, (3, 0, 0), 40, 0.1)
    bridge = chain(14, 0.2, start=(0.15, 0, 0))
    cases["bridged_blobs"] = (np.concatenate([a, b, bridge]), 0.3, 1, 1000, 1)
    cases["unbridged_blobs"] = (np.concatenate([a, b, bridge[:7]]), 0.3, 1, 1000, None)
    # size filter edges: components of size min-1, min, max, max+1
    comps = []
    for k, n in enumerate([4, 5, 20, 21]):
        comps.append(chain(n, 0.1, start=(0, 2.0 * k, 0)))
    cases["size_filter_edges"] = (np.concatenate(comps), 0.3, 5, 20, 2)
    # duplicates
    d = blob(rng, (0, 0, 0), 30, 0.2)
    cases["duplicates"] = (np.concatenate([d, d, d[:5]]), 0.3, 1, 1000, 1)
    cases["all_duplicates"] = (np.repeat(cloud([[0.5, 0.5, 0.5]]), 100, axis=0), 0.05, 1, 1000, 1)
    # all singletons (spacing > tol) and all one cluster
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(4), indexing="ij"), -1).reshape(-1, 3) * 1.0
    cases["all_singletons"] = (cloud(g), 0.5, 1, 10, len(g))
    cases["all_singletons_min2"] = (cloud(g), 0.5, 2, 10, 0)
    cases["one_cluster_lattice"] = (cloud(g * 0.4), 0.5, 1, 100000, 1)
    cases["oversized_dropped"] = (cloud(g * 0.4), 0.5, 1, 100, 0)
    # negative coordinates, large offset from the origin (cell coordinates far from zero)
    far = blob(rng, (-512.3, 977.1, -3.2), 200, 0.5)
    cases["far_from_origin"] = (far, 0.25, 1, 1000, None)
    # flat clouds (z constant -> zero cell bits on an axis) and a line
    flat = cloud(np.c_[rng.uniform(0, 10, (500, 2)), np.zeros(500)])
    cases["flat_z"] = (flat, 0.4, 2, 1000, None)
    cases["line_x"] = (chain(300, 0.05), 0.06, 1, 1000, 1)
    cases["line_x_gaps"] = (chain(300, 0.05), 0.05, 1, 1000, None)
    # equal-size clusters: order pinned by smallest index
    eq = np.concatenate([chain(10, 0.1, start=(0, 5.0 * k, 0)) for k in range(6)])
    eq = eq[rng.permutation(len(eq))]
    cases["equal_sizes"] = (eq, 0.3, 1, 1000, 6)
    # huge extent with a small tolerance: the voxel key needs more than 32 bits (64-bit key path)
    wide = np.concatenate([blob(rng, c, 60, 0.08) for c in rng.uniform([-1500, -1500, -20], [1500, 1500, 30], size=(40, 3))])
    cases["wide_extent_u64_keys"] = (wide, 0.05, 2, 1000, None)
    # random sparse / dense mixtures
    cases["uniform_sparse"] = (cloud(rng.uniform(0, 20, (4000, 3))), 0.5, 1, 100000, None)
    cases["uniform_dense"] = (cloud(rng.uniform(0, 4, (6000, 3))), 0.3, 3, 100000, None)
    return cases
