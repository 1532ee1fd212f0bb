"""Both implementations of the z-anti-aliased interlevel loss against the reference fixture, with a C2-sized timing:

  PS_ZAA_WARP=1 (default)  warp-per-ray kernel (csrc/losses.cu:zaa_interlevel_warp_kernel)
  PS_ZAA_WARP=0            thread-per-ray kernel (zaa_interlevel_kernel), the first version

Each runs in its own process because the library reads the switch once.
"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, time, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from helpers import Fixture, assert_close
from presight_b200 import losses, ops
fx = Fixture("zaa.npz")
pulse = [float(v) for v in fx.np("pulse_width")]
for case in "abc":
    n = int(fx.np(f"{case}/n_levels"))
    c, w = fx[f"{case}/c"].cuda(), fx[f"{case}/w"].cuda()
    ws = [fx[f"{case}/w{i}"].cuda().requires_grad_(True) for i in range(n)]
    ts = [fx[f"{case}/t{i}"].cuda() for i in range(n)]
    loss = losses.z_anti_aliasing_interlevel_loss([x[..., None] for x in ws] + [w[..., None]], ts + [c], pulse)
    assert_close(loss.cpu(), fx[f"{case}/loss"], 1e-5, "loss " + case)
    loss.backward()
    for i in range(n):
        assert_close(ws[i].grad.cpu(), fx[f"{case}/g{i}"], 2e-5, f"grad {case} level {i}")
# C2-sized timing: 65 536 rays, 64 final samples, 128 proposal samples
g = torch.Generator().manual_seed(0)
N, S, Sp = 65536, 64, 128
c = torch.rand(N, S + 1, generator=g).sort(-1).values.cuda(); w = (torch.rand(N, S, generator=g) ** 3 * 0.05).cuda()
t = torch.rand(N, Sp + 1, generator=g).sort(-1).values.cuda(); we = (torch.rand(N, Sp, generator=g) ** 3 * 0.05).cuda().requires_grad_(True)
for _ in range(3):
    ops.zaa_interlevel_loss_level(c, w, t, we, 0.03)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.zaa_interlevel_loss_level(c, w, t, we, 0.03)
b.record(); torch.cuda.synchronize()
print(f"zaa level, 65536 rays: {a.elapsed_time(b) / 10:.3f} ms (PS_ZAA_WARP=%(flag)s)")
"""


@pytest.mark.parametrize("flag", ["1", "0"])
def test_zaa_warp_per_ray_kernel(flag):
    env = dict(os.environ, PS_ZAA_WARP=flag)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "flag": flag}], env=env, capture_output=True,
                       text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stderr[-2000:]
