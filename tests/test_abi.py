"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
argument validation reports errors through ps_last_error, and the product refuses to run without CUDA tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "presight_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from presight_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
    return _lib.load()


def test_header_symbols_exported(lib):
    from presight_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/presight_b200.h but not exported"
    # the Python binding declares a prototype for every entry point of the header, and nothing else
    assert sorted(_lib.exported_symbols()) == names


def test_abi_version_and_errors(lib):
    assert lib.ps_abi_version() == 1
    # invalid arguments are reported, never thrown: L = 0
    scal = (ctypes.c_float * 1)(16.0)
    rc = lib.ps_hash_fwd(None, 4, None, scal, 0, 2, 19, None, None)
    assert rc != 0 and b"num_levels" in lib.ps_last_error()
    rc = lib.ps_hash_fwd(None, 4, None, scal, 1, 3, 19, None, None)
    assert rc != 0                      # null pointers / F=3
    dims = (ctypes.c_int * 3)(32, 40, 1)  # hidden width not a multiple of 16
    rc = lib.ps_mlp_fwd(None, 0, None, None, dims, 2, 0, 1, None, None)
    assert rc != 0 and b"multiple of 16" in lib.ps_last_error()
    # empty inputs are a no-op
    assert lib.ps_hash_fwd(None, 0, None, scal, 1, 2, 19, None, None) == 0


def test_no_cpu_fallback():
    from presight_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.hash_encode(torch.zeros(4, 3), torch.zeros(32 * 4, 2), [16.0, 32.0, 64.0, 128.0], 5)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.get_weights(torch.ones(2, 4), torch.ones(2, 4))
    # the loss stack is kernels only as well
    from presight_b200 import losses
    c, w = torch.linspace(0, 1, 5).repeat(2, 1), torch.full((2, 4, 1), 0.25)
    cp, wp = torch.linspace(0, 1, 9).repeat(2, 1), torch.full((2, 8, 1), 0.125, requires_grad=True)
    for fn in (lambda: losses.interlevel_loss([wp, w], [cp, c]),
               lambda: losses.z_anti_aliasing_interlevel_loss([wp, w], [cp, c], (0.03,)),
               lambda: losses.distortion_loss([w], [c]),
               lambda: losses.render_losses({"rgb": torch.zeros(2, 3), "accumulation": torch.zeros(2, 1)},
                                            {"rgb": torch.zeros(2, 3), "sky": torch.zeros(2, 1)}, True, False)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "presight_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
