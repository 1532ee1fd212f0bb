"""Size-independent properties of the oracle (the checker itself must be sane before it is trusted as one): index
ranges, corner aliasing, monotone samplers, conservation of the compositing weights, signs and zeros of the loss terms.
CPU only; complements the golden-fixture pinning of tests/test_oracle_golden.py."""
import numpy as np
import pytest
import torch

import oracle as O

torch.set_num_threads(1)


@pytest.mark.parametrize("L,lo,hi,log2T", [(16, 16, 2048, 22), (8, 16, 4096, 20), (5, 16, 128, 17), (10, 16, 16384, 20)])
def test_hash_rows_stay_inside_their_level(L, lo, hi, log2T):
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.rand(4000, 3, generator=g), torch.rand(500, 3, generator=g) * 3 - 1,      # outside the unit cube
                   torch.randint(0, 17, (500, 3), generator=g).float() / 16.0])                       # lattice points
    s = O.hash_scalings(L, lo, hi)
    idx, off = O.hash_corner_indices(x, s, log2T)
    T = 1 << log2T
    level = torch.arange(L)[None, :, None]
    assert idx.dtype == torch.int64 and bool(((idx >= level * T) & (idx < (level + 1) * T)).all())
    assert bool(((off >= 0) & (off < 1)).all())
    # a scaled coordinate that is an exact integer has ceil == floor: the x-ceil and x-floor corners alias
    exact_x = (x[:, None, 0] * s[None, :]) == torch.floor(x[:, None, 0] * s[None, :])
    assert bool((idx[..., 0][exact_x] == idx[..., 3][exact_x]).all()) and bool(exact_x.any())


def test_hash_encode_is_linear_in_the_table_and_local():
    g = torch.Generator().manual_seed(1)
    s = O.hash_scalings(4, 16, 128)
    x = torch.rand(300, 3, generator=g)
    t1, t2 = torch.randn(4 << 10, 2, generator=g), torch.randn(4 << 10, 2, generator=g)
    e = lambda t: O.hash_encode(x, O.HashGrid(t, s, 10))
    assert torch.allclose(e(t1 + 2 * t2), e(t1) + 2 * e(t2), atol=1e-5)
    # a constant table encodes to that constant (the trilinear weights sum to one)
    assert torch.allclose(e(torch.full((4 << 10, 2), 0.75)), torch.full((300, 8), 0.75), atol=1e-6)


@pytest.mark.parametrize("S_in,S_out", [(128, 64), (64, 64), (256, 96), (5, 9)])
def test_samplers_are_monotone_and_stay_in_range(S_in, S_out):
    g = torch.Generator().manual_seed(2)
    N = 200
    nears, fars = torch.full((N, 1), 0.005), torch.full((N, 1), 50.0)
    sp, eu = O.spaced_bins(nears, fars, S_in, 5.0, torch.rand(N, 1, generator=g))
    assert bool((sp[:, 1:] >= sp[:, :-1]).all()) and bool((eu[:, 1:] >= eu[:, :-1]).all())
    assert float(sp.min()) >= 0.0 and float(sp.max()) <= 1.0
    assert float(eu.min()) >= 0.005 - 1e-6 and float(eu.max()) <= 50.0 * (1 + 1e-5)
    w = torch.rand(N, S_in, generator=g) ** 4
    w[torch.rand(N, S_in, generator=g) < 0.3] = 0.0
    w[:3] = 0.0                                                        # rays with no weight at all
    bins, inds, cdf, u = O.pdf_resample(w, sp, S_out, torch.rand(N, 1, generator=g), eps=float(np.finfo(np.float32).eps))
    assert bins.shape == (N, S_out + 1) and bool((bins[:, 1:] >= bins[:, :-1]).all())
    assert float(bins.min()) >= 0.0 and float(bins.max()) <= 1.0
    assert bool((cdf[:, 1:] >= cdf[:, :-1]).all()) and bool((inds >= 0).all()) and bool((inds <= S_in + 1).all())


def test_compositing_weights_are_a_sub_probability():
    g = torch.Generator().manual_seed(3)
    deltas = torch.rand(500, 64, 1, generator=g) * 0.5
    dens = torch.exp(torch.randn(500, 64, 1, generator=g) * 2)
    dens[:5] = 0.0
    dens[5:8] = float("inf")                                            # nan_to_num path (rays.py:148)
    w = O.get_weights(deltas, dens)
    assert bool(torch.isfinite(w).all()) and float(w.min()) >= 0.0
    total = w.sum(1)
    assert float(total.max()) <= 1.0 + 1e-5 and float(total[:5].abs().max()) == 0.0
    # transmittance telescopes: sum of weights = 1 - exp(-sum(delta * sigma))
    ok = torch.isfinite(dens).all(1).squeeze(-1)
    want = 1 - torch.exp(-(deltas * dens).sum(1))
    assert torch.allclose(total[ok], want[ok], atol=1e-5)


def test_loss_terms_signs_and_zeros():
    g = torch.Generator().manual_seed(4)
    N, S, Sp = 50, 32, 64
    c = torch.rand(N, S + 1, generator=g).sort(-1).values
    c[:, 0], c[:, -1] = 0.0, 1.0
    w = torch.rand(N, S, generator=g)
    w = w / w.sum(-1, keepdim=True)
    # distortion: non-negative, zero for an empty ray, and equal to width / 3 for a single unit spike
    assert float(O.lossfun_distortion(c, w).min()) >= 0.0
    assert float(O.lossfun_distortion(c, torch.zeros(N, S)).abs().max()) == 0.0
    spike = torch.zeros(N, S)
    spike[:, 7] = 1.0
    assert torch.allclose(O.lossfun_distortion(c, spike), (c[:, 8] - c[:, 7]) / 3, atol=1e-7)
    # proposal losses vanish when the proposal histogram already bounds the final one: identical bins, weights x 1.5
    big = (w * 1.5)[..., None]
    assert float(O.interlevel_loss([big, w[..., None]], [c, c])) == 0.0
    # ... and are positive when the proposal puts no mass where the final level has some
    assert float(O.interlevel_loss([torch.zeros(N, S, 1), w[..., None]], [c, c])) > 0.0
    cp = torch.linspace(0, 1, Sp + 1).repeat(N, 1)
    assert float(O.z_anti_aliasing_interlevel_loss([torch.ones(N, Sp, 1), w[..., None]], [cp, c], (0.03,))) == 0.0
    assert float(O.z_anti_aliasing_interlevel_loss([torch.zeros(N, Sp, 1), w[..., None]], [cp, c], (0.03,))) > 0.0
    # the blurred histogram keeps the mass of the original one (its integral ends at sum(w) up to rounding)
    xr, yr = O.blur_stepfun(c, w / (c[:, 1:] - c[:, :-1]), 0.03)
    area = (0.5 * (yr[:, 1:] + yr[:, :-1]) * (xr[:, 1:] - xr[:, :-1])).sum(-1)
    assert torch.allclose(area, w.sum(-1), atol=2e-3)
    # sky loss: near zero when the accumulation agrees with the mask, large when it contradicts it
    acc, sky = torch.tensor([[0.0], [1.0]]), torch.tensor([[1.0], [0.0]])
    assert float(O.sky_loss(acc, sky)) < 1e-6 and float(O.sky_loss(1 - acc, sky)) > 10.0
