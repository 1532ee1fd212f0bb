"""Shared helpers for the parity tests: fixture loading and oracle-model construction."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIELD_META_KEYS = ("num_levels", "base_res", "max_res", "log2_hashmap_size", "features_per_level",
                   "geo_feat_dim", "semantic_dim", "appearance_embedding_dim")
PROP_META_KEYS = ("num_levels", "base_res", "max_res", "log2_hashmap_size", "features_per_level", "hidden_dim")

# the fixture generator's constants (tests/golden/make_golden.py)
THR, NEAR, FAR = 100.0 * 0.05, 0.1 * 0.05, 1000.0 * 0.05


class Fixture:
    """npz wrapper returning torch tensors; `sub(prefix)` yields a state-dict view."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name))
        self.keys = list(self.z.keys())

    def __contains__(self, k):
        return k in self.z

    def np(self, k):
        return self.z[k]

    def __getitem__(self, k):
        return torch.from_numpy(np.ascontiguousarray(self.z[k]))

    def sub(self, prefix):
        n = len(prefix)
        return {k[n:]: self[k] for k in self.keys if k.startswith(prefix)}


def field_meta(fx, key="meta/field"):
    m = dict(zip(FIELD_META_KEYS, (int(v) for v in fx.np(key))))
    m["use_semantics"] = m["semantic_dim"] > 0
    return m


def prop_meta(fx, key):
    return dict(zip(PROP_META_KEYS, (int(v) for v in fx.np(key))))


def rel_err(a, b):
    """max |a-b| / (max|b| + tiny): scale-relative error used for the fp32 / bf16 tolerances."""
    a = torch.as_tensor(a).detach().double()
    b = torch.as_tensor(b).detach().double()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: scale-relative error {e:.3e} > {tol:.1e}"


def elem_rel_err(a, b, floor):
    """max over elements of |a-b| / max(|b|, floor): the per-element relative error with an absolute floor — the metric
    BASELINE.json's "within 1e-3 relative" reads as for every element that is not itself rounding noise (`floor` is set
    per quantity, a few orders below its scale)."""
    a = torch.as_tensor(a).detach().double()
    b = torch.as_tensor(b).detach().double()
    if a.numel() == 0:
        return 0.0
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def assert_close_elem(a, b, tol, floor, what=""):
    e = elem_rel_err(a, b, floor)
    assert e <= tol, f"{what}: per-element relative error {e:.3e} > {tol:.1e} (floor {floor:.1e})"


def ms_count(sd, prefix="fields."):
    n = 0
    while any(k.startswith(f"{prefix}{n}.") for k in sd):
        n += 1
    return n
