"""The per-ray core of the z-anti-aliased interlevel loss (presight_b200/csrc/zaa_core.h — the code the CUDA kernel
runs per thread) compiled for the host and checked against the live reference's fixture.  CPU only."""
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
    out = str(tmp_path_factory.mktemp("zaa") / "libzaa_host.so")
    src = os.path.join(HERE, "native", "zaa_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.zaa_loss_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def run_host(lib, c, w, cp, wp, r):
    c, w, cp, wp = (np.ascontiguousarray(t.numpy(), dtype=np.float32) for t in (c, w, cp, wp))
    N, S, Sp = w.shape[0], w.shape[1], wp.shape[1]
    loss = ctypes.c_double(0.0)
    grad = np.zeros((N, Sp), dtype=np.float32)
    rc = lib.zaa_loss_host(c.ctypes.data, w.ctypes.data, N, S, cp.ctypes.data, wp.ctypes.data, Sp, float(r),
                           ctypes.byref(loss), grad.ctypes.data)
    assert rc == 0
    return loss.value / (N * Sp), torch.from_numpy(grad) / (N * Sp)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_zaa_core_matches_reference(host_lib, case):
    fx = Fixture("zaa.npz")
    pulse = [float(v) for v in fx.np("pulse_width")]
    n = int(fx.np(f"{case}/n_levels"))
    c, w = fx[f"{case}/c"], fx[f"{case}/w"]
    total = 0.0
    for i in range(n):
        loss_i, grad_i = run_host(host_lib, c, w, fx[f"{case}/t{i}"], fx[f"{case}/w{i}"], pulse[i])
        total += loss_i
        assert_close(grad_i, fx[f"{case}/g{i}"], 2e-5, f"grad level {i}")
    assert abs(total - float(fx[f"{case}/loss"])) <= 1e-5 * abs(float(fx[f"{case}/loss"])), (total, float(fx[f"{case}/loss"]))


def test_parallel_formulation_matches_reference():
    """The loop-free formulation planned for the warp-per-ray kernel (tools/zaa_parallel_prototype.py: merge by rank,
    scans, binary-search interval and flat-run lookup) reproduces the live reference's fixture."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import zaa_parallel_prototype as proto
    assert proto.check(os.path.join(HERE, "golden", "zaa.npz")) < 2e-5
