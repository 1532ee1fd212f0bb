"""The per-ray core of ps_generate_rays (presight_b200/csrc/raygen_core.h) compiled for the host and checked against
the live reference's fixture (tests/golden/rays.npz).  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import Fixture, assert_close

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("raygen") / "libraygen_host.so")
    src = os.path.join(HERE, "native", "raygen_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.raygen_host.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float] + \
        [ctypes.c_void_p] * 4
    return lib


def test_raygen_core_matches_reference(host_lib):
    fx = Fixture("rays.npz")
    f32 = lambda k: np.ascontiguousarray(fx.np(k), dtype=np.float32)
    c2w, fxs, fys, cxs, cys = f32("c2w"), f32("fx"), f32("fy"), f32("cx"), f32("cy")
    idx = np.ascontiguousarray(fx.np("ray_indices"), dtype=np.int64)
    N = idx.shape[0]
    o, d = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    area, norm = np.zeros(N, np.float32), np.zeros(N, np.float32)
    rc = host_lib.raygen_host(c2w.ctypes.data, fxs.ctypes.data, fys.ctypes.data, cxs.ctypes.data, cys.ctypes.data,
                              c2w.shape[0], idx.ctypes.data, N, 0.5, o.ctypes.data, d.ctypes.data, area.ctypes.data,
                              norm.ctypes.data)
    assert rc == 0
    assert np.array_equal(o, fx.np("origins"))
    assert_close(torch.from_numpy(d), fx["directions"], 1e-6, "directions")
    assert_close(torch.from_numpy(norm), fx["directions_norm"][:, 0], 1e-6, "directions_norm")
    assert_close(torch.from_numpy(area), fx["pixel_area"][:, 0], 1e-4, "pixel_area")
    # pixel_area is a product of two differences of nearly equal unit vectors: bound every ray, not only the largest
    rel = np.abs(area - fx.np("pixel_area")[:, 0]) / fx.np("pixel_area")[:, 0]
    assert float(rel.max()) < 1e-3, float(rel.max())
