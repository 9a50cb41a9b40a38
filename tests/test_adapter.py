"""The header-only C++ adapter (include/mot_b200_pcl.hpp): the reference's clusterPointCloud call sequence
(MOT.cpp:461-491) compiled against it.  CPU: it compiles against the PCL shim and fails loudly without a GPU.
GPU: its output equals the oracle's on a c1 frame."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "adapter_demo")


def build_demo(mot):
    pkg = os.path.dirname(mot.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(CPP, "adapter_demo.cpp"),
                           "-L", pkg, "-lmot_b200", f"-Wl,-rpath,{pkg}", "-o", EXE])


def write_inputs(tmp_path, synth):
    occ, res, origin = synth.make_map_c1()
    cloud, _ = synth.make_frame_c1(n_points=32768)
    cloud.tofile(tmp_path / "cloud.bin")
    occ.tofile(tmp_path / "map.bin")
    H, W = occ.shape
    args = [str(tmp_path / "cloud.bin"), str(tmp_path / "map.bin"), str(W), str(H), repr(float(res)), repr(origin[0]), repr(origin[1]), "0.3", "5", "300"]
    return occ, res, origin, cloud, args


def test_adapter_compiles_and_fails_loudly_without_gpu(mot, synth, tmp_path):
    build_demo(mot)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    _, _, _, _, args = write_inputs(tmp_path, synth)
    r = subprocess.run([EXE] + args, capture_output=True, text=True)
    assert r.returncode != 0 and "mot_create failed" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_adapter_matches_oracle(mot, oracle, synth, tmp_path):
    build_demo(mot)
    occ, res, origin, cloud, args = write_inputs(tmp_path, synth)
    r = subprocess.run([EXE] + args, capture_output=True, text=True, check=True)
    lines = r.stdout.strip().splitlines()
    M, K = map(int, lines[0].split())
    kept, _ = oracle.remove_static(cloud, occ, res, origin[:2], static_tolerance=2)
    off, idx = oracle.cluster_kdtree(kept, 0.3, 5, 300)
    cen = oracle.get_centroid(kept, off, idx, 2.5)
    assert M == len(kept) and K == len(off) - 1 and K > 5
    for k, line in enumerate(lines[1:]):
        size, first, cx, cy, inten = line.split()
        assert int(size) == off[k + 1] - off[k] and int(first) == idx[off[k]]
        np.testing.assert_allclose([float(cx), float(cy), float(inten)], cen[k, [0, 1, 3]], rtol=1e-5, atol=1e-6)
