#!/usr/bin/env python
"""Generate golden fixtures from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Imports the reference's own modules from /root/reference/nerfstudio-0.3.3 with the two
import-only shims of SURVEY §8c (stub `nerfacc`, stub `nerfstudio.configs.base_config`),
runs them on seeded inputs on CPU (`implementation="torch"`) and stores inputs, parameters
(as state dicts) and outputs/gradients.  The GPU box has no /root/reference: tests only read
the committed .npz files.
"""
import dataclasses
import os
import sys
import types

sys.dont_write_bytecode = True
REF = "/root/reference/nerfstudio-0.3.3"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    na = types.ModuleType("nerfacc")
    na.OccGridEstimator = type("OccGridEstimator", (), {})
    na.accumulate_along_rays = None
    sys.modules["nerfacc"] = na
    bc = types.ModuleType("nerfstudio.configs.base_config")

    @dataclasses.dataclass
    class PrintableConfig:
        pass

    @dataclasses.dataclass
    class InstantiateConfig(PrintableConfig):
        _target: type = None

        def setup(self, **kw):
            return self._target(self, **kw)

    bc.PrintableConfig = PrintableConfig
    bc.InstantiateConfig = InstantiateConfig
    sys.modules["nerfstudio.configs.base_config"] = bc
    sys.path.insert(0, REF)


install_shims()
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # noqa: E402
from nerfstudio.field_components.encodings import HashEncoding, SHEncoding  # noqa: E402
from nerfstudio.field_components.field_heads import FieldHeadNames  # noqa: E402
from nerfstudio.field_components.spatial_distortions import SceneContraction  # noqa: E402
from nerfstudio.fields.PreSight.ingp_field import iNGPField  # noqa: E402
from nerfstudio.fields.PreSight.ingp_field_ms import iNGPFieldMS  # noqa: E402
from nerfstudio.fields.PreSight.prop_density_field import PropNetDensityField  # noqa: E402
from nerfstudio.fields.PreSight.prop_density_field_ms import PropNetDensityFieldMS  # noqa: E402
from nerfstudio.fields.PreSight.sky_field import SkyField  # noqa: E402
from nerfstudio.fields.PreSight.sky_field_ms import SkyFieldMS  # noqa: E402
from nerfstudio.model_components.ray_samplers import PDFSampler, ProposalNetworkSampler, SpacedSampler  # noqa: E402
from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer  # noqa: E402
from nerfstudio.model_components.scene_colliders import NearFarCollider  # noqa: E402


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **npify(d))
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB, {len(d)} arrays)")


def special_points(n, gen):
    """Unit-cube points incl. exact grid integers, 0, 1, and out-of-range values."""
    x = torch.rand(n, 3, generator=gen)
    x[0] = torch.tensor([0.0, 0.0, 0.0])
    x[1] = torch.tensor([1.0, 1.0, 1.0])
    x[2] = torch.tensor([0.5, 0.25, 0.125])        # exact integers on every power-of-two level
    x[3] = torch.tensor([-0.2, 0.5, 1.3])          # outside the cube (negative coords hashed too)
    x[4] = torch.tensor([0.3, 0.7, 0.1])
    x[5] = torch.tensor([0.123456, 0.654321, 0.987654])
    x[6] = torch.tensor([0.0625, 0.9375, 0.5])
    x[7] = torch.tensor([-1.5, 2.5, -0.75])
    return x


def ref_corner_indices(enc, x):
    """Call the reference's own hash_fn on the 8 corner combinations of ENC:354-361."""
    t = x[..., None, :]
    scaled = t * enc.scalings.view(-1, 1)
    c = torch.ceil(scaled).type(torch.int32)
    f = torch.floor(scaled).type(torch.int32)
    pick = [(c, c, c), (c, f, c), (f, f, c), (f, c, c), (c, c, f), (c, f, f), (f, f, f), (f, c, f)]
    cols = [enc.hash_fn(torch.cat([a[..., 0:1], b[..., 1:2], d[..., 2:3]], dim=-1)) for a, b, d in pick]
    return torch.stack(cols, dim=-1), scaled - f


# ------------------------------------------------------------------------------------
def gen_hash():
    g = torch.Generator().manual_seed(1234)
    out = {}
    configs = {  # name: (L, min, max, log2T, F)
        "kat4": (4, 16, 1024, 5, 2),
        "main16": (16, 16, 2048, 19, 2),
        "presight10": (10, 16, 16384, 20, 4),
        "prop8a": (8, 16, 1024, 20, 1),
        "prop8b": (8, 16, 4096, 20, 1),
        "prop5a": (5, 16, 128, 17, 2),
        "prop5b": (5, 16, 256, 17, 2),
        "main16_c2": (16, 16, 2048, 22, 2),
        "f8": (3, 16, 64, 8, 8),
    }
    x = special_points(96, g)
    out["x"] = x
    for name, (L, lo, hi, log2T, F) in configs.items():
        # scalings / indices do not need a real table: build a tiny one and override sizes
        enc = HashEncoding(num_levels=L, min_res=lo, max_res=hi, log2_hashmap_size=min(log2T, 4),
                           features_per_level=F, implementation="torch")
        enc.log2_hashmap_size = log2T
        enc.hash_table_size = 2 ** log2T
        enc.hash_offset = torch.arange(L) * enc.hash_table_size
        idx, off = ref_corner_indices(enc, x)
        out[f"{name}/cfg"] = np.array([L, lo, hi, log2T, F])
        out[f"{name}/scalings"] = enc.scalings
        out[f"{name}/idx"] = idx
        out[f"{name}/offset"] = off
    # value + gradient fixtures on tables small enough to commit
    for name, (L, lo, hi, log2T, F) in {"v_l4f2": (4, 16, 1024, 5, 2), "v_l8f1": (8, 16, 1024, 10, 1),
                                        "v_l10f4": (10, 16, 16384, 9, 4), "v_l16f2": (16, 16, 2048, 11, 2),
                                        "v_l3f8": (3, 16, 64, 8, 8)}.items():
        torch.manual_seed(sum(map(ord, name)) + 7)
        enc = HashEncoding(num_levels=L, min_res=lo, max_res=hi, log2_hashmap_size=log2T,
                           features_per_level=F, implementation="torch")
        with torch.no_grad():
            enc.hash_table.mul_(1000.0)      # O(1) values so tolerances are meaningful
        xin = x.clone().requires_grad_(True)
        y = enc(xin)
        dout = torch.randn(y.shape, generator=g)
        y.backward(dout)
        out[f"{name}/cfg"] = np.array([L, lo, hi, log2T, F])
        out[f"{name}/table"] = enc.hash_table
        out[f"{name}/out"] = y
        out[f"{name}/dout"] = dout
        out[f"{name}/dtable"] = enc.hash_table.grad
        out[f"{name}/dx"] = xin.grad
    save("hash.npz", out)


# ------------------------------------------------------------------------------------
FIELD_META = dict(num_levels=6, base_res=16, max_res=512, log2_hashmap_size=9, features_per_level=2,
                  hidden_dim=64, hidden_dim_color=64, geo_feat_dim=15, use_semantics=True, semantic_dim=64,
                  appearance_embedding_dim=16)
PROP_META = [dict(num_levels=5, base_res=16, max_res=128, log2_hashmap_size=8, features_per_level=1,
                  hidden_dim=64, use_linear=False),
             dict(num_levels=5, base_res=16, max_res=256, log2_hashmap_size=8, features_per_level=2,
                  hidden_dim=16, use_linear=False)]


def make_field(aabb, meta=FIELD_META):
    f = iNGPField(aabb, hidden_dim=meta["hidden_dim"], num_levels=meta["num_levels"], max_res=meta["max_res"],
                  base_res=meta["base_res"], features_per_level=meta["features_per_level"],
                  log2_hashmap_size=meta["log2_hashmap_size"], hidden_dim_color=meta["hidden_dim_color"],
                  spatial_distortion=SceneContraction(order=float("inf")), use_semantics=meta["use_semantics"],
                  semantic_dim=meta["semantic_dim"], appearance_embedding_dim=meta["appearance_embedding_dim"],
                  implementation="torch")
    with torch.no_grad():
        f.mlp_base_grid.hash_table.mul_(300.0)
    return f


def make_prop(aabb, meta, use_linear=False):
    p = PropNetDensityField(aabb, hidden_dim=meta["hidden_dim"], spatial_distortion=SceneContraction(order=float("inf")),
                            use_linear=use_linear, num_levels=meta["num_levels"], max_res=meta["max_res"],
                            base_res=meta["base_res"], log2_hashmap_size=meta["log2_hashmap_size"],
                            features_per_level=meta["features_per_level"], implementation="torch")
    with torch.no_grad():
        p.encoding.hash_table.mul_(300.0)
    return p


def grads_of(module, prefix):
    return {f"{prefix}{k}": v.grad for k, v in module.named_parameters() if v.grad is not None}


def gen_fields():
    torch.manual_seed(77)
    g = torch.Generator().manual_seed(78)
    out = {}
    aabb = torch.tensor([[-1.0, -1.5, -0.25], [1.0, 1.5, 0.75]])
    out["meta/field"] = np.array([FIELD_META[k] for k in ("num_levels", "base_res", "max_res", "log2_hashmap_size",
                                                          "features_per_level", "geo_feat_dim", "semantic_dim",
                                                          "appearance_embedding_dim")])
    out["meta/prop0"] = np.array([PROP_META[0][k] for k in ("num_levels", "base_res", "max_res", "log2_hashmap_size",
                                                            "features_per_level", "hidden_dim")])
    out["meta/prop1"] = np.array([PROP_META[1][k] for k in ("num_levels", "base_res", "max_res", "log2_hashmap_size",
                                                            "features_per_level", "hidden_dim")])
    # points: inside the aabb, in the contracted shell, far away; one exactly on the aabb centre
    pos = (torch.rand(160, 3, generator=g) * 2 - 1) * torch.tensor([3.0, 4.0, 2.0])
    pos[0] = torch.tensor([0.0, 0.0, 0.25])
    pos[1] = torch.tensor([50.0, -20.0, 3.0])
    pos[2] = torch.tensor([1.0, 1.5, 0.75])
    dirs = torch.nn.functional.normalize(torch.randn(160, 3, generator=g), dim=-1)
    app = torch.randn(160, 16, generator=g)
    out.update({"pos": pos, "dirs": dirs, "app": app})

    # --- single iNGPField forward/backward
    f = make_field(aabb)
    for k, v in f.state_dict().items():
        out[f"field/{k}"] = v
    rs = RaySamples(frustums=Frustums(origins=pos, directions=dirs, starts=torch.zeros(160, 1),
                                      ends=torch.zeros(160, 1), pixel_area=torch.ones(160, 1)))
    fo = f(rs, appearance_embedding=app)
    den, rgb, sem = fo[FieldHeadNames.DENSITY], fo[FieldHeadNames.RGB], fo[FieldHeadNames.SEMANTICS]
    gd, gr, gs = torch.randn(den.shape, generator=g), torch.randn(rgb.shape, generator=g), torch.randn(sem.shape, generator=g)
    (den * gd).sum().add((rgb * gr).sum()).add((sem * gs).sum()).backward()
    out.update({"field_out/density": den, "field_out/rgb": rgb, "field_out/semantics": sem,
                "field_out/g_density": gd, "field_out/g_rgb": gr, "field_out/g_semantics": gs})
    out.update(grads_of(f, "field_grad/"))
    with torch.no_grad():
        out["field_out/semantic_fn"] = f.semantic_fn(pos)
        d2, emb = f.density_fn(pos)
        out["field_out/embedding"] = emb

    # --- proposal fields (MLP 64, MLP 16, linear)
    for name, meta, lin in (("prop0", PROP_META[0], False), ("prop1", PROP_META[1], False), ("proplin", PROP_META[0], True)):
        p = make_prop(aabb, meta, lin)
        for k, v in p.state_dict().items():
            out[f"{name}/{k}"] = v
        d = p.density_fn(pos)
        gdd = torch.randn(d.shape, generator=g)
        (d * gdd).sum().backward()
        out[f"{name}_out/density"] = d
        out[f"{name}_out/g_density"] = gdd
        out.update(grads_of(p, f"{name}_grad/"))

    # --- sky field
    sky = SkyField(mlp_num_layers=3, mlp_layer_width=32, appearance_embedding_dim=16, use_semantics=True,
                   semantic_dim=64, implementation="torch")
    for k, v in sky.state_dict().items():
        out[f"sky/{k}"] = v
    so = sky.get_outputs(dirs, app)
    out["sky_out/rgb"] = so[FieldHeadNames.RGB]
    out["sky_out/semantics"] = so[FieldHeadNames.SEMANTICS]

    # --- SH encoding alone
    out["sh_out"] = SHEncoding(levels=4, implementation="torch")((dirs + 1) / 2)

    # --- multi-sub-field routing (nf = 3)
    centroids = torch.tensor([[-1.5, 0.0, 0.0], [1.0, 2.0, 0.5], [0.5, -2.5, 0.0]])
    aabbs = [aabb + c for c in centroids]
    fms = iNGPFieldMS([make_field(a) for a in aabbs], centroids)
    pms = PropNetDensityFieldMS([make_prop(a, PROP_META[0]) for a in aabbs], centroids)
    for k, v in fms.state_dict().items():
        out[f"ms_field/{k}"] = v
    for k, v in pms.state_dict().items():
        out[f"ms_prop/{k}"] = v
    out["ms/centroids"] = centroids
    out["ms/assign"] = torch.cdist(pos, centroids).argmin(dim=1)
    rs = RaySamples(frustums=Frustums(origins=pos, directions=dirs, starts=torch.zeros(160, 1),
                                      ends=torch.zeros(160, 1), pixel_area=torch.ones(160, 1)))
    with torch.no_grad():
        mo = fms(rs, appearance_embedding=app)
        out["ms_out/density"] = mo[FieldHeadNames.DENSITY]
        out["ms_out/rgb"] = mo[FieldHeadNames.RGB]
        out["ms_out/semantics"] = mo[FieldHeadNames.SEMANTICS]
        out["ms_out/prop_density"] = pms.density_fn(pos)
        out["ms_out/semantic_fn"] = fms.semantic_fn(pos)
    save("fields.npz", out)


# ------------------------------------------------------------------------------------
THR, NEAR, FAR = 100.0 * 0.05, 0.1 * 0.05, 1000.0 * 0.05


def make_spaced(single_jitter=True):
    thr = THR
    return SpacedSampler(
        spacing_fn=lambda x: torch.where(x < thr, x / (2 * thr), 1 - 1 / (2 * x / thr)),
        spacing_fn_inv=lambda x: torch.where(x < 0.5, x * (2 * thr), thr / (2 - 2 * x)),
        single_jitter=single_jitter)


def make_bundle(n, g):
    o = (torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([2.0, 2.0, 0.1])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * torch.tensor([1.0, 1.0, 0.3]), dim=-1)
    rb = RayBundle(origins=o, directions=d, pixel_area=torch.ones(n, 1) * 1e-6,
                   camera_indices=torch.zeros(n, 1, dtype=torch.long))
    return rb


def gen_sampler_render():
    g = torch.Generator().manual_seed(4321)
    out = {}
    n = 48
    rb = make_bundle(n, g)
    rb.nears = torch.ones(n, 1) * NEAR
    rb.fars = torch.ones(n, 1) * FAR
    out.update({"origins": rb.origins, "dirs": rb.directions, "nears": rb.nears, "fars": rb.fars})

    # --- spaced sampler, train (single jitter) and eval
    for S in (128, 256, 48):
        sp = make_spaced()
        sp.train()
        torch.manual_seed(100 + S)
        rs = sp(rb, num_samples=S)
        torch.manual_seed(100 + S)
        t_rand = torch.rand((n, 1))
        out[f"spaced{S}/t_rand"] = t_rand
        out[f"spaced{S}/sp_starts"] = rs.spacing_starts[..., 0]
        out[f"spaced{S}/sp_ends"] = rs.spacing_ends[..., 0]
        out[f"spaced{S}/starts"] = rs.frustums.starts[..., 0]
        out[f"spaced{S}/ends"] = rs.frustums.ends[..., 0]
        out[f"spaced{S}/positions"] = rs.frustums.get_positions()
        sp.eval()
        rs_e = sp(rb, num_samples=S)
        out[f"spaced{S}/eval_starts"] = rs_e.frustums.starts[..., 0]
        out[f"spaced{S}/eval_ends"] = rs_e.frustums.ends[..., 0]

    # --- get_weights + renderers on random densities (incl. zero and huge densities)
    sp = make_spaced()
    sp.train()
    torch.manual_seed(7)
    rs = sp(rb, num_samples=64)
    dens = torch.exp(torch.randn(n, 64, 1, generator=g) * 1.5)
    dens[0] = 0.0
    dens[1] = 1e6
    dens[2, 10:] = 0.0
    dens = dens.requires_grad_(True)
    rgb = torch.rand(n, 64, 3, generator=g).requires_grad_(True)
    sem = torch.randn(n, 64, 8, generator=g).requires_grad_(True)
    w = rs.get_weights(dens)
    r_rgb = RGBRenderer(background_color="black")
    r_rgb.train()
    img = r_rgb(rgb=rgb, weights=w)
    acc = AccumulationRenderer()(weights=w)
    dexp = DepthRenderer(method="expected")(weights=w, ray_samples=rs)
    dthr = DepthRenderer(method="threshold")(weights=w, ray_samples=rs)
    semo = torch.sum(sem * w, dim=-2)
    g_img, g_acc = torch.randn(img.shape, generator=g), torch.randn(acc.shape, generator=g)
    g_dexp, g_sem, g_w = torch.randn(dexp.shape, generator=g), torch.randn(semo.shape, generator=g), torch.randn(w.shape, generator=g)
    ((img * g_img).sum() + (acc * g_acc).sum() + (dexp * g_dexp).sum() + (semo * g_sem).sum() + (w * g_w).sum()).backward()
    cw = torch.cumsum(w[..., 0], dim=-1)
    out.update({"render/starts": rs.frustums.starts[..., 0], "render/ends": rs.frustums.ends[..., 0],
                "render/deltas": rs.deltas[..., 0], "render/density": dens, "render/rgb": rgb, "render/sem": sem,
                "render/weights": w, "render/img": img, "render/acc": acc, "render/depth_expected": dexp,
                "render/depth_threshold": dthr, "render/sem_out": semo,
                "render/depth_index": torch.clamp(torch.searchsorted(cw, torch.ones(n, 1) * 0.5, side="left"), 0, 63),
                "render/g_img": g_img, "render/g_acc": g_acc, "render/g_dexp": g_dexp, "render/g_sem": g_sem,
                "render/g_w": g_w, "render/d_density": dens.grad, "render/d_rgb": rgb.grad, "render/d_sem": sem.grad})

    # --- PDF sampler: train single jitter / eval; several (S_in, S_out); PreSight eps
    eps = torch.finfo(torch.float32).eps
    for S_in, S_out in ((128, 64), (64, 64), (256, 96), (96, 48)):
        sp = make_spaced()
        sp.train()
        torch.manual_seed(11 + S_in)
        rs = sp(rb, num_samples=S_in)
        wts = torch.exp(torch.randn(n, S_in, 1, generator=g) * 2.0)
        wts = wts / wts.sum(dim=1, keepdim=True) * torch.rand(n, 1, 1, generator=g)
        wts[0] = 0.0                       # all-zero weights -> padding path
        wts[1, :, 0] = 0.0
        wts[1, 5, 0] = 1.0                 # delta distribution
        pdf = PDFSampler(include_original=False, single_jitter=True)
        for mode in ("train", "eval"):
            pdf.train(mode == "train")
            torch.manual_seed(200 + S_in + S_out)
            ns = pdf(rb, rs, wts, num_samples=S_out, eps=eps)
            torch.manual_seed(200 + S_in + S_out)
            rand = torch.rand((n, 1))
            key = f"pdf{S_in}_{S_out}/{mode}"
            out[f"{key}/sp_starts"] = ns.spacing_starts[..., 0]
            out[f"{key}/sp_ends"] = ns.spacing_ends[..., 0]
            out[f"{key}/starts"] = ns.frustums.starts[..., 0]
            out[f"{key}/ends"] = ns.frustums.ends[..., 0]
            if mode == "train":
                out[f"pdf{S_in}_{S_out}/rand"] = rand
        out[f"pdf{S_in}_{S_out}/weights"] = wts[..., 0]
        out[f"pdf{S_in}_{S_out}/existing_sp"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1)
        # the reference's intermediate cdf / inds for the train case (restating RS:305-345 on the reference tensors)
        w2 = wts[..., 0] + 0.01
        ws = torch.sum(w2, dim=-1, keepdim=True)
        pad = torch.relu(eps - ws)
        w2 = w2 + pad / w2.shape[-1]
        ws = ws + pad
        cdf = torch.min(torch.ones_like(w2), torch.cumsum(w2 / ws, dim=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        nb = S_out + 1
        u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb).expand(n, nb) + out[f"pdf{S_in}_{S_out}/rand"] / nb
        out[f"pdf{S_in}_{S_out}/cdf"] = cdf
        out[f"pdf{S_in}_{S_out}/u"] = u.contiguous()
        out[f"pdf{S_in}_{S_out}/inds"] = torch.searchsorted(cdf, u.contiguous(), side="right")
    save("sampler_render.npz", out)


# ------------------------------------------------------------------------------------
def gen_model():
    """Restated get_outputs (MODEL:452-546) over the reference's own components, nf = 2, train mode."""
    torch.manual_seed(2024)
    g = torch.Generator().manual_seed(2025)
    out = {}
    n = 40
    centroids = torch.tensor([[-0.6, 0.0, 0.0], [0.6, 0.1, 0.0]])
    base_aabb = torch.tensor([[-1.0, -1.0, -0.25], [1.0, 1.0, 0.75]])
    aabbs = [base_aabb + c for c in centroids]
    field = iNGPFieldMS([make_field(a) for a in aabbs], centroids)
    props = torch.nn.ModuleList([PropNetDensityFieldMS([make_prop(a, PROP_META[i]) for a in aabbs], centroids)
                                 for i in range(2)])
    sky = SkyFieldMS([SkyField(mlp_num_layers=3, mlp_layer_width=32, appearance_embedding_dim=16,
                               use_semantics=True, semantic_dim=64, implementation="torch") for _ in aabbs], centroids)
    sampler = ProposalNetworkSampler(num_nerf_samples_per_ray=24, num_proposal_samples_per_ray=(48, 32),
                                     num_proposal_network_iterations=2, single_jitter=True,
                                     update_sched=lambda s: 1, initial_sampler=make_spaced())
    collider = NearFarCollider(near_plane=NEAR, far_plane=FAR)
    for m in (field, props, sky, sampler):
        m.train()
    collider.training = True
    r_rgb = RGBRenderer(background_color="black")
    r_rgb.train()
    r_depth, r_exp, r_acc = DepthRenderer("threshold"), DepthRenderer("expected"), AccumulationRenderer()

    rb = make_bundle(n, g)
    app = torch.randn(n, 16, generator=g)
    out.update({"origins": rb.origins, "dirs": rb.directions, "app": app, "centroids": centroids,
                "samples": np.array([48, 32, 24])})
    for k, v in field.state_dict().items():
        out[f"field/{k}"] = v
    for i, p in enumerate(props):
        for k, v in p.state_dict().items():
            out[f"prop{i}/{k}"] = v
    for k, v in sky.state_dict().items():
        out[f"sky/{k}"] = v

    rb = collider(rb)
    torch.manual_seed(555)
    ray_samples, weights_list, ray_samples_list = sampler(rb, density_fns=[p.density_fn for p in props])
    torch.manual_seed(555)
    for i in range(3):
        out[f"jitter{i}"] = torch.rand((n, 1))
    S = ray_samples.frustums.starts.shape[1]
    app_s = app[:, None, :].expand(n, S, 16)
    fo = field.forward(ray_samples, appearance_embedding=app_s)
    weights = ray_samples.get_weights(fo[FieldHeadNames.DENSITY])
    weights_list.append(weights)
    ray_samples_list.append(ray_samples)
    rgb = r_rgb(rgb=fo[FieldHeadNames.RGB], weights=weights)
    with torch.no_grad():
        depth = r_depth(weights=weights, ray_samples=ray_samples)
    expected = r_exp(weights=weights, ray_samples=ray_samples)
    acc = torch.clamp(r_acc(weights=weights), min=0.0, max=1.0)
    so = sky(ray_samples, appearance_embedding=app_s)
    rgb = rgb + (1.0 - acc) * so[FieldHeadNames.RGB]
    sem = torch.sum(fo[FieldHeadNames.SEMANTICS] * weights, dim=-2) + (1.0 - acc) * so[FieldHeadNames.SEMANTICS]
    out.update({"out/rgb": rgb, "out/depth": depth, "out/expected_depth": expected, "out/accumulation": acc,
                "out/semantics": sem})
    for i in range(3):
        out[f"out/weights{i}"] = weights_list[i][..., 0]
        rs = ray_samples_list[i]
        out[f"out/sp_bins{i}"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1)
        out[f"out/eu_bins{i}"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[..., -1:, 0]], -1)
    for i in range(2):
        out[f"out/prop_depth_{i}"] = r_depth(weights=weights_list[i], ray_samples=ray_samples_list[i])

    # a loss that seeds every differentiable output (stand-in for MODEL:558-645)
    tgt_rgb, tgt_sem = torch.rand(n, 3, generator=g), torch.rand(n, 64, generator=g)
    gw = [torch.randn(weights_list[i].shape, generator=g) * 0.1 for i in range(2)]
    loss = ((rgb - tgt_rgb) ** 2).mean() + 0.5 * ((sem - tgt_sem) ** 2).mean() + 0.1 * expected.mean() \
        + 0.01 * acc.mean() + (weights_list[0] * gw[0]).sum() / n + (weights_list[1] * gw[1]).sum() / n
    loss.backward()
    out.update({"loss/tgt_rgb": tgt_rgb, "loss/tgt_sem": tgt_sem, "loss/gw0": gw[0], "loss/gw1": gw[1],
                "loss/value": loss})
    out.update(grads_of(field, "grad/field/"))
    for i, p in enumerate(props):
        out.update(grads_of(p, f"grad/prop{i}/"))
    out.update(grads_of(sky, "grad/sky/"))

    # eval-mode depth (MODEL:688-708) and the prior query (XP:130-138)
    for m in (field, props, sky, sampler):
        m.eval()
    collider.training = False
    with torch.no_grad():
        rb2 = collider(make_bundle(n, torch.Generator().manual_seed(2025)))
        rsamp, _, _ = sampler(rb2, density_fns=[p.density_fn for p in props])
        den, _ = field.get_density(rsamp)
        w = rsamp.get_weights(den)
        out["depth_eval/depth"] = r_depth(weights=w, ray_samples=rsamp, threshold=0.5)
        out["depth_eval/expected_depth"] = r_exp(weights=w, ray_samples=rsamp)
        pts = (torch.rand(200, 3, generator=g) - 0.5) * torch.tensor([3.0, 2.5, 1.0])
        dl = [p.density_fn(pts).squeeze(-1) for p in props]
        dl.append(field.density_fn(pts)[0].squeeze(-1))
        out["query/points"] = pts
        out["query/density_mean"] = torch.stack(dl, dim=0).mean(dim=0)
        out["query/features"] = field.semantic_fn(pts).clip(0.0, 1.0).to(torch.float16)
    save("model.npz", out)


if __name__ == "__main__":
    torch.set_num_threads(1)          # deterministic reductions
    gen_hash()
    gen_fields()
    gen_sampler_render()
    gen_model()
