"""The per-element core of ps_adam_step (presight_b200/csrc/adam_core.h) compiled for the host and checked against
torch.optim.Adam — which IS the reference's optimiser (engine/optimizers.py:133-140 instantiates it with
configs/method_configs.py:115's lr 1e-2, eps 1e-15, weight_decay 1e-5).  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import assert_close

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("adam") / "libadam_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                    os.path.join(HERE, "native", "adam_host.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.adam_host.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] + [ctypes.c_double] * 5 + [ctypes.c_int64]
    lib.adam_host.restype = None
    return lib


@pytest.mark.parametrize("wd", [1e-5, 0.0])
def test_adam_core_matches_torch_adam(host_lib, wd):
    g = torch.Generator().manual_seed(0)
    n = 10007
    p0 = (torch.rand(n, generator=g) * 2 - 1) * 1e-3                    # hash-table init scale
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15, weight_decay=wd, foreach=False)
    p = p0.numpy().copy()
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 8):
        grad = torch.randn(n, generator=g) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))
        grad[torch.rand(n, generator=g) < 0.3] = 0.0                    # untouched table entries have exactly zero gradient
        ref.grad = grad.clone()
        opt.step()
        gn = np.ascontiguousarray(grad.numpy())
        host_lib.adam_host(p.ctypes.data, gn.ctypes.data, m.ctypes.data, v.ctypes.data, n, 1e-2, 0.9, 0.999, 1e-15, wd, step)
        st = opt.state[ref]
        assert_close(torch.from_numpy(m), st["exp_avg"], 1e-6, f"exp_avg step {step}")
        assert_close(torch.from_numpy(v), st["exp_avg_sq"], 1e-6, f"exp_avg_sq step {step}")
        assert_close(torch.from_numpy(p), ref.detach(), 1e-6, f"param step {step}")
