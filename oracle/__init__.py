"""CPU oracle for the PreSight NeRF inner loop — TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a torch-CPU restatement of the reference's *PyTorch* path
(`implementation="torch"`), function by function, each citing the reference
file:line it follows (paths relative to /root/reference/nerfstudio-0.3.3/nerfstudio).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker or the reported
CPU baseline.  Nothing under `presight_b200/` imports it.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the live reference
modules in the build container (with the two import-only shims of SURVEY §8c),
runs them on seeded inputs and commits the inputs/outputs as fixtures under
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every oracle function
against those fixtures (bit-exact for indices, exact-or-1e-6 for fp32 values) and
against the reference's own known answers (SURVEY §8c KATs,
tests/cameras/test_rays.py:11-30, tests/utils/test_math.py:8-16).
"""
from .nerf_oracle import *  # noqa: F401,F403
