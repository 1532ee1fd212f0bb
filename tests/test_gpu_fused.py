"""The level-fused fast paths (presight_b200/fused.py) against the modular drop-in path and the CPU oracle."""
import copy

import pytest
import torch

import oracle as O
from oracle import state as OS
from helpers import FAR, NEAR, THR, Fixture, assert_close, field_meta, prop_meta

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_single_field_model(impl, samples=(48, 32, 32)):
    """nf = 1 model with the golden fixture's network sizes (base 12->64->80, sem 64x3, rgb 47->64->64->3)."""
    from presight_b200.model import NerfactoNuscMSModel, NerfactoNuscMSModelConfig
    ffx = Fixture("fields.npz")
    m = field_meta(ffx)
    p0, p1 = prop_meta(ffx, "meta/prop0"), prop_meta(ffx, "meta/prop1")

    def pargs(p):
        return {"features_per_level": p["features_per_level"], "log2_hashmap_size": p["log2_hashmap_size"],
                "num_levels": p["num_levels"], "base_res": p["base_res"], "max_res": p["max_res"],
                "hidden_dim": p["hidden_dim"], "use_linear": False}
    cfg = NerfactoNuscMSModelConfig(
        near_plane=NEAR, far_plane=FAR, piecewise_sampler_threshold=THR, num_levels=m["num_levels"],
        base_res=m["base_res"], max_res=m["max_res"], log2_hashmap_size=m["log2_hashmap_size"],
        features_per_level=m["features_per_level"], num_proposal_samples_per_ray=samples[:2],
        num_nerf_samples_per_ray=samples[2], proposal_net_args_list=[pargs(p0), pargs(p1)], implementation=impl,
        appearance_embed_dim=4, video_embed_dim=12, sky_mlp_dims=32, semantic_dim=m["semantic_dim"])
    torch.manual_seed(5)
    aabb = torch.tensor([[[-1.0, -1.0, -0.25], [1.0, 1.0, 0.75]]])
    model = NerfactoNuscMSModel(cfg, torch.zeros(1, 3), aabb, num_train_cameras=7, num_train_videos=3)
    with torch.no_grad():
        for f in model.field.fields:
            f.mlp_base_grid.hash_table.mul_(300.0)
        for p in model.proposal_networks:
            for f in p.fields:
                f.encoding.hash_table.mul_(300.0)
    return model.to(DEV), cfg


def make_batch(n, seed=0):
    from presight_b200.cameras.rays import RayBundle
    from presight_b200.model import VIDEO_ID
    g = torch.Generator().manual_seed(seed)
    o = (torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([2.0, 2.0, 0.1])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * torch.tensor([1.0, 1.0, 0.3]), dim=-1)
    cam = torch.randint(0, 7, (n, 1), generator=g)
    vid = torch.randint(0, 3, (n, 1), generator=g)
    jit = [torch.rand(n, 1, generator=g) for _ in range(3)]
    tgt = {"rgb": torch.rand(n, 3, generator=g), "sem": torch.rand(n, 64, generator=g),
           "gw0": torch.randn(n, 48, 1, generator=g) * 0.1, "gw1": torch.randn(n, 32, 1, generator=g) * 0.1}

    def bundle():
        return RayBundle(origins=o.to(DEV), directions=d.to(DEV), camera_indices=cam.to(DEV),
                         metadata={VIDEO_ID: vid.to(DEV)})
    return bundle, [j.to(DEV) for j in jit], {k: v.to(DEV) for k, v in tgt.items()}, (o, d, cam, vid, jit)


def loss_of(out, tgt, n):
    wl = out["weights_list"]
    return ((out["rgb"] - tgt["rgb"]) ** 2).mean() + 0.5 * ((out["semantics"] - tgt["sem"]) ** 2).mean() \
        + 0.1 * out["expected_depth"].mean() + 0.01 * out["accumulation"].mean() \
        + (wl[0] * tgt["gw0"]).sum() / n + (wl[1] * tgt["gw1"]).sum() / n


@pytest.mark.parametrize("impl", ["b200+fp32", "b200"])
def test_fused_matches_modular(impl):
    n = 300
    fused_model, _ = build_single_field_model(impl)
    modular = copy.deepcopy(fused_model)
    modular.use_fused = False
    modular.proposal_sampler.use_fused = False
    bundle, jit, tgt, _ = make_batch(n)
    for m in (fused_model, modular):
        m.train()
    out_f = fused_model(bundle(), jitters=jit)
    out_m = modular(bundle(), jitters=jit)
    # fp32: same kernels, same inputs, differences are only fp32 summation order.  bf16: the fused levels run the
    # tcgen05 kernels, the modular path the mma.sync ones — same parity class, values on a bf16 rounding boundary flip
    tol = 1e-5 if impl == "b200+fp32" else 3e-3
    for k in ("rgb", "accumulation", "expected_depth", "semantics", "depth"):
        assert_close(out_f[k], out_m[k], tol, k)
    for i in range(3):
        assert_close(out_f["weights_list"][i], out_m["weights_list"][i], tol, f"weights {i}")
    loss_of(out_f, tgt, n).backward()
    loss_of(out_m, tgt, n).backward()
    pf, pm = dict(fused_model.named_parameters()), dict(modular.named_parameters())
    checked = 0
    for k, p in pm.items():
        if p.grad is None:
            assert pf[k].grad is None or float(pf[k].grad.abs().max()) == 0.0, k
            continue
        assert pf[k].grad is not None, f"fused path produced no gradient for {k}"
        assert rel_l2(pf[k].grad, p.grad) < (1e-4 if impl == "b200+fp32" else 2e-2), \
            f"{k}: {rel_l2(pf[k].grad, p.grad):.2e}"
        checked += 1
    assert checked >= 25


def test_fused_matches_oracle_fp32():
    n = 200
    model, cfg = build_single_field_model("b200+fp32")
    model.train()
    bundle, jit, tgt, (o, d, cam, vid, jit_cpu) = make_batch(n, seed=3)
    out = model(bundle(), jitters=jit)
    # oracle with the same weights
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ffx = Fixture("fields.npz")
    fm = field_meta(ffx)
    field = OS.ngp_from_state(sd, "field.fields.0.", fm, True)
    props = [[OS.prop_from_state(sd, f"proposal_networks.{i}.fields.0.", prop_meta(ffx, f"meta/prop{i}"), True)]
             for i in range(2)]
    sky = [OS.sky_from_state(sd, "sky_model.fields.0.", True, True)]
    ocfg = O.ModelCfg(num_proposal_samples=(48, 32), num_nerf_samples=32, near=NEAR, far=FAR, piecewise_thr=THR)
    om = O.Model(ocfg, torch.zeros(1, 3), [field], props, sky)
    app = torch.cat([sd["appearance_embedding.embedding.weight"][cam[:, 0]],
                     sd["video_embedding.embedding.weight"][vid[:, 0]]], dim=-1)
    oo = O.model_outputs(om, o, d, app, jit_cpu)
    for k in ("rgb", "accumulation", "expected_depth", "semantics"):
        assert_close(out[k].detach().cpu(), oo[k], 1e-3, k)
    for i in range(3):
        assert_close(out["weights_list"][i].detach().cpu(), oo["weights_list"][i], 1e-3, f"weights {i}")
    tgt_c = {k: v.cpu() for k, v in tgt.items()}
    loss_of(out, tgt, n).backward()
    loss_of(oo, tgt_c, n).backward()
    g = dict(model.named_parameters())
    assert rel_l2(g["field.fields.0.mlp_base_grid.hash_table"].grad.cpu(), field.grid.table.grad) < 5e-3
    assert rel_l2(g["field.fields.0.rgb_head.layers.0.weight"].grad.cpu(), field.rgb.weights[0].grad) < 5e-3
    assert rel_l2(g["field.fields.0.semantic_head.layers.2.weight"].grad.cpu(), field.sem.weights[2].grad) < 5e-3
    assert rel_l2(g["field.fields.0.mlp_base_mlp.layers.0.weight"].grad.cpu(), field.base.weights[0].grad) < 5e-3
    assert rel_l2(g["proposal_networks.0.fields.0.encoding.hash_table"].grad.cpu(), props[0][0].grid.table.grad) < 5e-3
    assert rel_l2(g["proposal_networks.1.fields.0.mlp_base.1.layers.0.weight"].grad.cpu(),
                  props[1][0].net.weights[0].grad) < 5e-3


def test_tc5_model_matches_oracle_bf16():
    """The default path — both proposal levels and the final level on the fused tcgen05 kernels — against the CPU
    oracle on the same weights and jitters: 1e-2 (bf16 parity class) on every output, gradients by rel-L2."""
    from presight_b200 import fused
    n = 256
    model, cfg = build_single_field_model("b200", samples=(64, 32, 32))
    model.train()
    bundle, jit, tgt, (o, d, cam, vid, jit_cpu) = make_batch(n, seed=4)
    tgt["gw0"] = torch.randn(n, 64, 1, generator=torch.Generator().manual_seed(1)).to(DEV) * 0.1
    out = model(bundle(), jitters=jit)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ffx = Fixture("fields.npz")
    fm = field_meta(ffx)
    field = OS.ngp_from_state(sd, "field.fields.0.", fm, True)
    props = [[OS.prop_from_state(sd, f"proposal_networks.{i}.fields.0.", prop_meta(ffx, f"meta/prop{i}"), True)]
             for i in range(2)]
    sky = [OS.sky_from_state(sd, "sky_model.fields.0.", True, True)]
    # the fused kernels must actually be the ones that ran
    f0, p0 = model.field.fields[0], model.proposal_networks[0].fields[0]
    assert fused.USE_TC5_FIELD and fused.USE_TC5_PROP
    assert f0.mlp_base_mlp.layers[0].weight.shape[1] <= 48 and p0.encoding.features_per_level in (1, 2)
    ocfg = O.ModelCfg(num_proposal_samples=(64, 32), num_nerf_samples=32, near=NEAR, far=FAR, piecewise_thr=THR)
    om = O.Model(ocfg, torch.zeros(1, 3), [field], props, sky)
    app = torch.cat([sd["appearance_embedding.embedding.weight"][cam[:, 0]],
                     sd["video_embedding.embedding.weight"][vid[:, 0]]], dim=-1)
    oo = O.model_outputs(om, o, d, app, jit_cpu)
    for k in ("rgb", "accumulation", "expected_depth", "semantics"):
        assert_close(out[k].detach().cpu(), oo[k], 1e-2, k)
    # weights: identical sample positions are not guaranteed once a bf16 proposal weight moves a PDF bin, so compare
    # the first level (same bins by construction) tightly and the rendered quantities above for the rest
    assert_close(out["weights_list"][0].detach().cpu(), oo["weights_list"][0], 1e-2, "weights 0")
    tgt_c = {k: v.cpu() for k, v in tgt.items()}
    loss_of(out, tgt, n).backward()
    loss_of(oo, tgt_c, n).backward()
    g = dict(model.named_parameters())
    assert rel_l2(g["field.fields.0.mlp_base_grid.hash_table"].grad.cpu(), field.grid.table.grad) < 5e-2
    assert rel_l2(g["field.fields.0.rgb_head.layers.0.weight"].grad.cpu(), field.rgb.weights[0].grad) < 5e-2
    assert rel_l2(g["field.fields.0.semantic_head.layers.2.weight"].grad.cpu(), field.sem.weights[2].grad) < 5e-2
    assert rel_l2(g["proposal_networks.0.fields.0.encoding.hash_table"].grad.cpu(), props[0][0].grid.table.grad) < 5e-2


def test_mlp_segments_and_density_epilogue():
    """ps_mlp_*_ex: [per-ray | strided window | per-ray] inputs equal torch.cat of the pieces; the per-ray input
    gradient equals the sum over the ray's samples; the density epilogue equals trunc_exp * selector."""
    from presight_b200 import fused, ops
    g = torch.Generator().manual_seed(12)
    N, S, hd = 37, 32, 80
    P = N * S
    sh = torch.randn(N, 16, generator=g)
    h = torch.randn(P, hd, generator=g)
    app = torch.randn(N, 16, generator=g)
    ws = [torch.randn(64, 47, generator=g) / 7, torch.randn(64, 64, generator=g) / 8, torch.randn(3, 64, generator=g) / 8]
    bs = [torch.randn(64, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1, torch.randn(3, generator=g) * 0.1]
    dy = torch.randn(P, 3, generator=g)
    # oracle: explicit concatenation
    shc, hc, appc = sh.clone().requires_grad_(True), h.clone().requires_grad_(True), app.clone().requires_grad_(True)
    x = torch.cat([shc[:, None, :].expand(N, S, 16).reshape(P, 16), hc[:, 1:16],
                   appc[:, None, :].expand(N, S, 16).reshape(P, 16)], dim=-1)
    wc = [w.clone().requires_grad_(True) for w in ws]
    yc = O.mlp_forward(x, O.Mlp(wc, [b.clone() for b in bs], "sigmoid"))
    yc.backward(dy)
    # kernel
    meta = fused.MlpMeta((47, 64, 64, 3), ops.ACT_SIGMOID)
    shg, hg, appg = sh.to(DEV), h.to(DEV), app.to(DEV)
    wg, bg = [w.to(DEV) for w in ws], [b.to(DEV) for b in bs]
    y = torch.empty(P, 3, device=DEV)
    segs = [(shg, None, 16, 0, 16, S), (hg, None, hd, 1, 15, 1), (appg, None, 16, 0, 16, S)]
    fused._mlp_fwd(segs, P, wg, bg, meta, 0, y)
    assert_close(y.cpu(), yc, 1e-3, "segmented forward")
    dh = torch.zeros(P, hd, device=DEV)
    dapp = torch.zeros(N, 16, device=DEV)
    dW, db = [torch.zeros_like(w) for w in wg], [torch.zeros_like(b) for b in bg]
    segs_b = [(shg, None, 16, 0, 16, S), (hg, dh, hd, 1, 15, 1), (appg, dapp, 16, 0, 16, S)]
    fused._mlp_bwd(segs_b, dy.to(DEV), P, wg, bg, meta, 0, dW, db)
    assert_close(dh.cpu(), hc.grad, 1e-3, "strided-window input gradient")
    assert_close(dapp.cpu(), appc.grad, 1e-3, "per-ray input gradient")
    assert_close(dW[0].cpu(), wc[0].grad, 1e-3, "dW0")
    # density epilogue on an 80-wide base MLP
    ws2 = [torch.randn(64, 32, generator=g) / 6, torch.randn(80, 64, generator=g) / 8]
    bs2 = [torch.randn(64, generator=g) * 0.1, torch.randn(80, generator=g) * 0.1]
    feat = torch.randn(P, 32, generator=g)
    sel = torch.rand(P, generator=g) > 0.1
    gd, gh = torch.randn(P, generator=g), torch.randn(P, 80, generator=g)
    fc = feat.clone().requires_grad_(True)
    hh = O.mlp_forward(fc, O.Mlp([w.clone() for w in ws2], [b.clone() for b in bs2]))
    dens = O.trunc_exp(hh[:, 0]) * sel
    gh0 = gh.clone()
    gh0[:, 0] = 0
    ((dens * gd).sum() + (hh * gh0).sum()).backward()
    meta2 = fused.MlpMeta((32, 64, 80), ops.ACT_NONE)
    w2, b2 = [w.to(DEV) for w in ws2], [b.to(DEV) for b in bs2]
    fg, selg = feat.to(DEV), sel.to(DEV).to(torch.uint8)
    hout, dout = torch.empty(P, 80, device=DEV), torch.empty(P, device=DEV)
    fused._mlp_fwd([(fg, None, 32, 0, 32, 1)], P, w2, b2, meta2, 0, hout, selg, dout)
    assert_close(hout.cpu(), hh, 1e-3, "h")
    assert_close(dout.cpu(), dens, 1e-3, "density epilogue")
    dfeat = torch.empty(P, 32, device=DEV)
    dW2, db2 = [torch.zeros_like(w) for w in w2], [torch.zeros_like(b) for b in b2]
    gh_dev = gh.to(DEV)      # column 0 deliberately NOT zeroed: the kernel must ignore it when d_density is given
    fused._mlp_bwd([(fg, dfeat, 32, 0, 32, 1)], gh_dev, P, w2, b2, meta2, 0, dW2, db2, selg, gd.to(DEV))
    assert_close(dfeat.cpu(), fc.grad, 1e-3, "d features through the density epilogue")


def test_fused_prior_query_matches_modular_and_oracle():
    model, cfg = build_single_field_model("b200+fp32")
    model.eval()
    modular = copy.deepcopy(model)
    modular.use_fused = False
    g = torch.Generator().manual_seed(4)
    pts = (torch.rand(5000, 3, generator=g) - 0.5) * torch.tensor([3.0, 2.5, 1.0])
    mean_f, feats_f = model.query_priors(pts.to(DEV))
    mean_m, feats_m = modular.query_priors(pts.to(DEV))
    assert feats_f.dtype == torch.float16 and feats_f.shape == (5000, 64)
    assert_close(mean_f, mean_m, 1e-5, "mean density fused vs modular")
    assert_close(feats_f.float(), feats_m.float(), 1e-3, "features fused vs modular")
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ffx = Fixture("fields.npz")
    field = OS.ngp_from_state(sd, "field.fields.0.", field_meta(ffx))
    props = [[OS.prop_from_state(sd, f"proposal_networks.{i}.fields.0.", prop_meta(ffx, f"meta/prop{i}"))]
             for i in range(2)]
    om = O.Model(O.ModelCfg(), torch.zeros(1, 3), [field], props, None)
    mean_o, feats_o = O.prior_query(om, pts)
    assert_close(mean_f.cpu(), mean_o, 1e-3, "mean density vs oracle")
    assert_close(feats_f.float().cpu(), feats_o.float(), 2e-3, "features vs oracle")
    # empty query
    m0, f0 = model.query_priors(torch.zeros(0, 3, device=DEV))
    assert m0.shape == (0,) and f0.shape == (0, 64)


# ------------------------------------------------------------------------------------------ tcgen05 field level
def _tc5_case(n, S, A=16, L=16, F=2, seed=0):
    """Random field of the reference architecture + rays; returns everything both paths need."""
    from presight_b200 import fused, ops
    g = torch.Generator().manual_seed(seed)
    log2T = 12
    scal = O.hash_scalings(L, 16, 512).tolist()
    table = ((torch.rand(L << log2T, F, generator=g) * 2 - 1) * 0.5).to(DEV)
    dims = {"base": (L * F, 64, 80), "sem": (64, 64, 64, 64), "rgb": (31 + A, 64, 64, 3)}
    nets = {}
    for k, dd in dims.items():
        ws = [(torch.randn(dd[i + 1], dd[i], generator=g) / dd[i] ** 0.5).to(DEV).requires_grad_(True) for i in range(len(dd) - 1)]
        bs = [(torch.randn(dd[i + 1], generator=g) * 0.1).to(DEV).requires_grad_(True) for i in range(len(dd) - 1)]
        nets[k] = (ws, bs)
    o = ((torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([1.0, 1.0, 0.1])).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    eu = (torch.rand(n, S + 1, generator=g) * 0.08 + 0.002).cumsum(-1).to(DEV)
    app = torch.randn(n, A, generator=g).to(DEV).requires_grad_(True) if A else None
    grid = fused.GridMeta(tuple(scal), log2T, F)
    metas = (fused.MlpMeta(dims["base"], ops.ACT_NONE), fused.MlpMeta(dims["sem"], ops.ACT_NONE),
             fused.MlpMeta(dims["rgb"], ops.ACT_SIGMOID))
    aabb = [-1.0, -1.0, -0.5, 1.0, 1.0, 0.5]
    return dict(o=o, d=d, eu=eu, app=app, table=table.requires_grad_(True), grid=grid, metas=metas, nets=nets,
                aabb=aabb, A=A)


@pytest.mark.parametrize("n,S,A,L,F", [(1000, 64, 16, 16, 2), (333, 32, 16, 10, 4), (130, 128, 7, 16, 2),
                                       (65, 96, 0, 6, 2), (1, 64, 16, 16, 2)])
def test_tc5_field_forward_matches_modular(n, S, A, L, F):
    """ps_field_level_fwd (tcgen05, one kernel) vs the chain of stand-alone bf16 kernels on the same inputs."""
    from presight_b200 import fused, ops
    c = _tc5_case(n, S, A, L, F)
    base, sem, rgb = c["metas"]
    assert fused.tc5_field_supported(c["grid"], base, sem, rgb, 15, ops.PREC_BF16, S, A)
    with torch.no_grad():
        ws = [*c["nets"]["base"][0], *c["nets"]["sem"][0], *c["nets"]["rgb"][0]]
        bs = [*c["nets"]["base"][1], *c["nets"]["sem"][1], *c["nets"]["rgb"][1]]
        got = fused.tc5_field_forward(c["o"], c["d"], c["eu"], None if c["app"] is None else c["app"].detach(),
                                      c["table"].detach(), c["aabb"], True, c["grid"], [w.detach() for w in ws],
                                      [b.detach() for b in bs], A, 0.5)
        _, _, _, w, rgb_o, acc, dexp, dthr, sem_o, tmm = got
        want = fused._FieldLevel.apply(c["o"], c["d"], c["eu"], c["app"], c["table"], c["aabb"], True, c["grid"], base,
                                       sem, rgb, 15, ops.PREC_BF16, 0.5, *c["nets"]["base"][0], *c["nets"]["base"][1],
                                       *c["nets"]["sem"][0], *c["nets"]["sem"][1], *c["nets"]["rgb"][0],
                                       *c["nets"]["rgb"][1])
    ww, wrgb, wacc, wdexp, wdthr, wsem, wtmm = want
    assert_close(w, ww[..., 0], 3e-3, "weights")
    assert_close(rgb_o, wrgb, 3e-3, "rgb")
    assert_close(acc, wacc, 3e-3, "accumulation")
    assert_close(dexp, wdexp, 3e-3, "expected depth")
    assert_close(sem_o, wsem, 3e-3, "semantics")
    assert torch.equal(tmm, wtmm)
    # the threshold depth is an index decision: identical unless the cumulative weight sits on the threshold
    assert float((dthr != wdthr).float().mean()) < 0.02


@pytest.mark.parametrize("n,S,A,L,F", [(700, 64, 16, 16, 2), (333, 32, 16, 10, 4), (130, 128, 7, 16, 2),
                                       (65, 96, 0, 6, 2), (40000, 64, 16, 16, 2)])
def test_tc5_field_backward_matches_modular(n, S, A, L, F):
    """ps_field_level_bwd (one tcgen05 kernel: recompute + dgrad + wgrad + compositing backward) vs the chain of
    stand-alone bf16 kernels: same loss, same inputs, every gradient."""
    from presight_b200 import fused, ops
    c = _tc5_case(n, S, A, L, F, seed=1)
    base, sem, rgb = c["metas"]
    g = torch.Generator().manual_seed(7)
    tgt = {k: v.to(DEV) for k, v in dict(rgb=torch.rand(n, 3, generator=g), sem=torch.rand(n, 64, generator=g),
                                         gw=torch.randn(n, S, 1, generator=g) * 0.05).items()}
    ws = [*c["nets"]["base"][0], *c["nets"]["sem"][0], *c["nets"]["rgb"][0]]
    bs = [*c["nets"]["base"][1], *c["nets"]["sem"][1], *c["nets"]["rgb"][1]]
    leaves = [c["table"], *ws, *bs] + ([c["app"]] if A else [])

    def loss_of(out):
        w, rgb_o, acc, dexp, _, sem_o, _ = out
        return ((rgb_o - tgt["rgb"]) ** 2).mean() + 0.5 * ((sem_o - tgt["sem"]) ** 2).mean() + 0.1 * dexp.mean() \
            + 0.01 * acc.mean() + (w * tgt["gw"]).sum() / n

    out_t = fused._FieldLevelTc5.apply(c["o"], c["d"], c["eu"], c["app"], c["table"], c["aabb"], True, c["grid"], 0.5,
                                       *ws, *bs)
    g_t = torch.autograd.grad(loss_of(out_t), leaves)
    out_m = fused._FieldLevel.apply(c["o"], c["d"], c["eu"], c["app"], c["table"], c["aabb"], True, c["grid"], base, sem,
                                    rgb, 15, ops.PREC_BF16, 0.5, *c["nets"]["base"][0], *c["nets"]["base"][1],
                                    *c["nets"]["sem"][0], *c["nets"]["sem"][1], *c["nets"]["rgb"][0],
                                    *c["nets"]["rgb"][1])
    g_m = torch.autograd.grad(loss_of(out_m), leaves)
    names = ["table"] + [f"W{i}" for i in range(8)] + [f"b{i}" for i in range(8)] + (["app"] if A else [])
    for name, a, b in zip(names, g_t, g_m):
        assert torch.isfinite(a).all(), name
        e = rel_l2(a, b)
        assert e < 2e-2, f"{name}: rel-L2 {e:.3e}"


# ------------------------------------------------------------------------------------------ tcgen05 proposal level
@pytest.mark.parametrize("n,S,L,F,H,log2T", [(900, 128, 8, 1, 64, 14), (901, 64, 8, 1, 64, 14), (333, 32, 5, 2, 16, 12),
                                            (70, 96, 8, 2, 64, 12), (1, 64, 8, 1, 16, 10), (30000, 64, 8, 1, 64, 20)])
def test_tc5_prop_level_matches_modular(n, S, L, F, H, log2T):
    """ps_prop_level_fwd / _bwd (one kernel each) vs the chain ray_points -> hash -> mma.sync MLP -> compositing."""
    from presight_b200 import fused, ops
    g = torch.Generator().manual_seed(n + S)
    scal = tuple(O.hash_scalings(L, 16, 1024).tolist())
    table = ((torch.rand(L << log2T, F, generator=g) * 2 - 1) * 2.0).to(DEV).requires_grad_(True)
    dims = (L * F, H, 1)
    ws = [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(DEV).requires_grad_(True) for i in range(2)]
    bs = [(torch.randn(dims[i + 1], generator=g) * 0.1).to(DEV).requires_grad_(True) for i in range(2)]
    o = ((torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([1.0, 1.0, 0.1])).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    eu = (torch.rand(n, S + 1, generator=g) * 0.05 + 0.001).cumsum(-1).to(DEV)
    gw = (torch.randn(n, S, 1, generator=g) * 0.1).to(DEV)
    grid = fused.GridMeta(scal, log2T, F)
    meta = fused.MlpMeta(dims, ops.ACT_NONE)
    aabb = [-1.0, -1.0, -0.5, 1.0, 1.0, 0.5]
    assert fused.tc5_prop_supported(grid, meta, ops.PREC_BF16, S)
    leaves = [table, *ws, *bs]
    w_t = fused._PropLevelTc5.apply(o, d, eu, table, aabb, True, grid, None, ws[0], bs[0], ws[1], bs[1])
    g_t = torch.autograd.grad((w_t * gw).sum(), leaves)
    w_m = fused._PropLevel.apply(o, d, eu, table, aabb, True, grid, meta, ops.PREC_BF16, *ws, *bs)
    g_m = torch.autograd.grad((w_m * gw).sum(), leaves)
    assert_close(w_t, w_m, 3e-3, "weights")
    for name, a, b in zip(["table", "W0", "W1", "b0", "b1"], g_t, g_m):
        assert torch.isfinite(a).all(), name
        e = rel_l2(a, b)
        # b1 is ONE number, the sum of all d_raw (heavy cancellation): the stand-alone kernel sums bf16-rounded values,
        # the fused one fp32 values, so they differ by the rounding noise of that sum rather than by 1e-2 of its value
        assert e < (1e-1 if name == "b1" else 2e-2), f"{name}: rel-L2 {e:.3e}"
    # no-grad forward (eval / non-update steps) takes the same kernel without saving features
    with torch.no_grad():
        w_n = fused._PropLevelTc5.apply(o, d, eu, table, aabb, True, grid, None, ws[0], bs[0], ws[1], bs[1])
    assert torch.equal(w_n, w_t)
