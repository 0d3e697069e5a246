"""The reference's OWN host Python (models/neus.py, models/split_mixed_occ.py, models/geometry.py, models/texture.py,
models/network_utils.py, models/volrend.py, lib/pbr/light.py -- imported unmodified from the reference tree) run on the
CPU over the oracle's restatements of its third-party imports (oracle/ref_host.py, backend="oracle"), against the
hand-written restatements oracle/neus.py and oracle/split.py.  This is what pins those two files -- the yardsticks
of every end-to-end GPU parity test -- to the reference's Python (models/neus.py:227-327,
models/split_mixed_occ.py:224-456).  Runs where the reference tree exists (the build container); on the GPU box the
same facts travel as tests/golden/ref_host_*.npz (tests/golden/make_ref_host_golden.py, tests/test_gpu_ref_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle import fields as ofields
from oracle import neus as oneus
from oracle import ref_host
from oracle import split as osplit

pytestmark = pytest.mark.skipif(ref_host.reference_root() is None, reason="reference tree not present on this box")


@pytest.fixture(autouse=True)
def _cpu_division():
    """`x / scalar` in models/utils.py:109-114 is a true division on the CPU (a reciprocal multiply on CUDA, which the
    oracle follows by default): compare like with like."""
    ofields.CUDA_SCALAR_DIV = False
    yield
    ofields.CUDA_SCALAR_DIV = True


def aabb_contraction(model):
    """models/neus.py:57 has `self.geometry.contraction_type = ContractionType.AABB` commented out and no system
    assigns it ("assigned in system", models/geometry.py:71), so the shipped neus model raises in
    contract_to_unisphere; the assignment models/split_mixed_occ.py:66 makes is applied by hand."""
    import sys
    model.geometry.contraction_type = sys.modules["models.geometry"].ContractionType.AABB


def neus_state(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        model.geometry.encoding.encoding.params.copy_(
            (torch.rand(model.geometry.encoding.encoding.params.shape, generator=g) * 2 - 1) * 0.05)
        w = model.geometry.network.layers[0].weight_v
        w[:, 3:] = torch.randn(w[:, 3:].shape, generator=g) * 0.05


def params_of(model, radius, sh_degree):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    enc = model.geometry.encoding.encoding
    geo = ofields.mlp_layers_from_state(sd, "geometry.network.")
    tex = ofields.mlp_layers_from_state(sd, "texture.network.")
    return oneus.NeusParams(sd["geometry.encoding.encoding.params"], geo, tex, sd["variance.variance"], enc.meta,
                            radius=radius, sh_degree=sh_degree)


def test_reference_neus_forward_and_grads_equal_oracle_neus():
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.neus import neus_blender_config
    cfg = ref_host.ref_config(neus_blender_config())
    with ref_host.reference_modules("oracle") as models:
        torch.manual_seed(0)
        m = models.make("neus", cfg)
        aabb_contraction(m)
        neus_state(m)
        m.train()
        m.randomized = False
        m.cos_anneal_ratio = 0.37
        grid = syn.analytic_grid("ball")
        m.occupancy_grid.binaries = grid[None]
        m.render_step_size = 1.732 * 2 * 1.5 / 192
        rays, rgb, fg, bg = syn.training_rays(96, seed=2)
        m.background_color = bg
        out = m(rays)
        loss, parts = oneus.loss({**out, "rays_valid": out["rays_valid_full"]}, rgb, fg)
        loss.backward()

        P = params_of(m, 1.5, 4)
        for t in P.tensors():
            t.requires_grad_(True)
        ref = oneus.forward(P, rays, grid.numpy(), m.render_step_size, 0.37, background=bg, training=True,
                            create_graph=True)
        rloss, rparts = oneus.loss(ref, rgb, fg)
        rloss.backward()

        assert int(out["num_samples"]) == ref["num_samples"] > 1000
        assert torch.equal(out["ray_indices"], ref["ray_indices"])
        for k in ("comp_rgb", "comp_normal", "opacity", "depth", "comp_rgb_full", "sdf_samples", "sdf_grad_samples",
                  "weights"):
            a, b = out[k].detach(), ref[k].detach()
            assert a.shape == b.shape, k
            assert float((a - b).abs().max()) <= 2e-6 * max(float(b.abs().max()), 1.0), (k, float((a - b).abs().max()))
        assert abs(float(loss) - float(rloss)) <= 1e-6 * abs(float(rloss))
        named = dict(m.named_parameters())
        checks = [("geometry.encoding.encoding.params", P.table), ("variance.variance", P.variance)]
        for i, layer in enumerate(P.geo_mlp):
            checks += [(f"geometry.network.layers.{2 * i}.{n}", t) for n, t in layer.items()]
        for i, layer in enumerate(P.tex_mlp):
            checks += [(f"texture.network.layers.{2 * i}.{n}", t) for n, t in layer.items()]
        for name, t in checks:
            g, r = named[name].grad, t.grad
            assert g is not None and r is not None, name
            err = float((g - r).norm() / r.norm().clamp_min(1e-30))
            assert err <= 2e-5, (name, err)


def test_reference_neus_eval_chunks_and_occupancy_update_equal_oracle():
    """models/neus.py:319-327 (eval: chunk_batch + .cpu()) and :90-122 (update_step: cos anneal + occupancy update)."""
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.neus import neus_blender_config
    cfg = ref_host.ref_config(neus_blender_config())
    cfg["ray_chunk"] = 40
    with ref_host.reference_modules("oracle") as models:
        torch.manual_seed(0)
        m = models.make("neus", cfg)
        aabb_contraction(m)
        neus_state(m, seed=3)
        m.train()
        jit = torch.rand(128 ** 3, 3, generator=torch.Generator().manual_seed(9))
        m.occupancy_grid.update_jitter = jit
        m.update_step(0, 0)                                   # warm-up branch: all cells
        assert m.cos_anneal_ratio == 0.0
        P = params_of(m, 1.5, 4)
        occs, binary = oneus.grid_update(torch.zeros(128 ** 3), 0, oneus.occ_eval_fn(P, m.render_step_size),
                                         [-1.5] * 3 + [1.5] * 3, occ_thre=0.001, jitter=jit)
        assert float((m.occupancy_grid.occs - occs).abs().max()) <= 1e-6
        # cells whose occupancy sits within fp32 rounding of the threshold (the oracle evaluates the 2 M points in
        # 256 k chunks, the reference in one batch: different GEMM blocking) may land on either side
        assert int((m.occupancy_grid.binaries[0] != binary).sum()) <= 16
        binary = m.occupancy_grid.binaries[0].clone()
        assert 0.05 < float(binary.float().mean()) < 0.95
        m.eval()
        m.render_step_size = 1.732 * 2 * 1.5 / 128
        rays, _, _, bg = syn.training_rays(100, seed=5)
        m.background_color = bg
        with torch.no_grad():
            out = m(rays)
        ref = oneus.forward(P, rays, binary.numpy(), m.render_step_size, 0.0, background=bg)
        assert int(out["num_samples"].sum()) == ref["num_samples"]
        for k in ("comp_rgb", "opacity", "depth", "comp_rgb_full"):
            a, b = out[k].detach(), ref[k].detach()
            assert float((a - b).abs().max()) <= 2e-6 * max(float(b.abs().max()), 1.0), k
        # comp_normal = normalize(sum w n) (models/neus.py:277) amplifies rounding by 1/opacity on rays that graze the
        # surface (40-ray chunks vs one batch: different GEMM blocking) -> error scaled back by min(1, opacity / 1e-2)
        cond = (ref["opacity"].detach() / 1e-2).clamp(max=1.0)
        en = (out["comp_normal"] - ref["comp_normal"].detach()).abs() * cond
        assert float(en.max()) <= 5e-5 and float(en.median()) <= 5e-7, (float(en.max()), float(en.median()))


# ------------------------------------------------------------------------------------------------ split-mixed-occ
def split_params_of(model):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    pb = model.geometry.encoding.encoding                       # the reference's ProgressiveBandHashGrid
    geo = ofields.mlp_layers_from_state(sd, "geometry.network.")
    nets = {n: ofields.mlp_layers_from_state(sd, f"texture.{n}_network.")
            for n in ("albedo", "roughness", "metallic", "env", "secondary")}
    return osplit.SplitParams(sd["geometry.encoding.encoding.encoding.params"], pb.encoding.meta, geo, nets,
                              sd["variance.variance"], sd["texture.FG_LUT"], sd["emitter.base"],
                              radius=model.config.radius, level_mask=pb.mask.detach().clone(),
                              fd_eps=model.geometry._finite_difference_eps)


@pytest.fixture
def small_pyramid(monkeypatch):
    """The dense CPU prefilter is O(texels^2): a 32 -> 16 -> 8 pyramid (three levels, like 512 -> ... -> 16 has six)
    keeps the host logic under test and the run time in seconds."""
    monkeypatch.setattr(osplit, "LIGHT_MIN_RES", 8)
    return 32


def make_split(models, base_res, seed=0):
    import sys
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.split_mixed_occ import split_mixed_occ_config
    cfg = ref_host.ref_config(split_mixed_occ_config())
    cfg["light"]["envlight_config"]["base_res"] = base_res
    ref_host.write_bsdf_lut(syn.bsdf_lut())
    sys.modules["lib.pbr.light"].EnvironmentLightMipCube.LIGHT_MIN_RES = osplit.LIGHT_MIN_RES
    torch.manual_seed(seed)
    m = models.make("split-mixed-occ", cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        p = m.geometry.encoding.encoding.encoding.params
        p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.02)
        w = m.geometry.network.layers[0].weight_v
        w[:, 3:] = torch.randn(w[:, 3:].shape, generator=g) * 0.05
        m.variance.variance.fill_(0.5)
    m.eval()
    m.update_step(0, 20000)                  # eval mode: schedule only (stage 1, all levels, eps), no occupancy update
    assert m.stage == 1 and float(m.geometry.encoding.encoding.mask.min()) == 1.0
    m.occupancy_grid.binaries = syn.analytic_grid("ball")[None]
    m.render_step_size = 1.732 * 2 * 1.5 / 192
    return m


@pytest.mark.parametrize("relighting", [False, True])
def test_reference_split_eval_render_equals_oracle_split(relighting, small_pyramid):
    """models/split_mixed_occ.py:224-456 in eval mode (chunk_batch, + the third bounce when relighting) ==
    oracle.split.forward.  Both sides in float64 (ref_host `double=True`): finite-difference normals amplify any fp32
    rounding difference between two implementations by 1/(2 eps) ~ 1400 (a 1-ulp difference in sdf is 2.4e-4 in the
    normal), which would hide a logic difference; in double the two agree to 1e-12."""
    from rise_sdf_b200 import synthetic as syn
    with ref_host.reference_modules("oracle", double=True) as models:
        m = make_split(models, small_pyramid)
        rays, _, _, bg = syn.training_rays(72, seed=3)
        rays, bg = rays.double(), bg.double()
        m.double()
        m.background_color = bg
        m.config["ray_chunk"] = 50
        with torch.no_grad():
            m.emitter.build_mips()
            out = m(rays, relighting)
        P = split_params_of(m)
        with torch.no_grad():
            osplit.build_mips(P)
        for a, b in zip(m.emitter.specular + [m.emitter.diffuse], P.specular + [P.diffuse]):
            assert a.dtype == torch.float64 and float((a - b).abs().max()) <= 1e-12
        ref = osplit.forward(P, rays, syn.analytic_grid("ball").numpy(), m.render_step_size, stage=1,
                             relighting=relighting, background=bg, dtype=torch.float64)
        assert int(out["num_samples"].sum()) == ref["num_samples"] > 200
        assert len(ref["valid_indices"]) > 10
        for k in ("comp_rgb", "comp_rgb_phys", "comp_normal", "opacity", "depth", "comp_albedo", "comp_roughness",
                  "comp_metallic", "comp_spec_rgb", "comp_spec_rgb_phys", "comp_diffuse_rgb", "comp_diffuse_rgb_phys",
                  "comp_blend", "comp_rgb_full", "comp_rgb_phys_full"):
            e = float((out[k] - ref[k]).abs().max())
            assert out[k].dtype == torch.float64 and e <= 1e-9 * max(float(ref[k].abs().max()), 1.0), (k, e)


def test_reference_split_training_step_equals_oracle_split(small_pyramid):
    """Training mode: curvature probe, normal-orientation map, reflection bounce with gradients, every loss term of
    systems/split_occ.py:163-225 and the gradient of every parameter group incl. emitter.base (through build_mips
    with the reference's own cubemap_mip backward).  float64 on both sides, as above."""
    from rise_sdf_b200 import synthetic as syn
    with ref_host.reference_modules("oracle", double=True) as models:
        m = make_split(models, small_pyramid)
        m.double()
        m.train()
        m.randomized = False
        rays, rgb, fg, bg = (t.double() for t in syn.training_rays(64, seed=4))
        m.background_color = bg
        dirs = []
        real = torch.rand_like
        torch.rand_like = lambda x, *a, **k: (dirs.append(real(x, *a, **k)), dirs[-1])[1]
        try:
            m.emitter.build_mips()
            out = m(rays)
        finally:
            torch.rand_like = real
        assert len(dirs) == 1 and dirs[0].shape == (int(out["num_samples"]), 3)
        loss, parts = osplit.loss(out, rgb, fg, stage=1)
        loss.backward()

        P = split_params_of(m)
        for t in P.named_tensors().values():
            t.requires_grad_(True)
        osplit.build_mips(P)
        ref = osplit.forward_train(P, rays, syn.analytic_grid("ball").numpy(), m.render_step_size, dirs[0], stage=1,
                                   background=bg, dtype=torch.float64)
        rloss, rparts = osplit.loss(ref, rgb, fg, stage=1)
        rloss.backward()
        assert int(out["num_samples"]) == ref["num_samples"] and torch.equal(out["ray_indices"], ref["ray_indices"])
        for k in ("comp_rgb", "comp_rgb_phys", "comp_normal", "opacity", "depth", "sdf_samples", "sdf_grad_samples",
                  "sdf_laplace_samples", "weights", "normals_orientation_loss_map", "comp_rgb_full", "comp_rgb_phys_full"):
            e = float((out[k].detach() - ref[k].detach()).abs().max())
            assert e <= 1e-9 * max(float(ref[k].detach().abs().max()), 1.0), (k, e)
        for k in rparts:
            assert abs(float(parts[k]) - float(rparts[k])) <= 1e-10 * max(abs(float(rparts[k])), 1e-3), k
        named = dict(m.named_parameters())
        named["emitter.base"] = m.emitter.base
        for name, t in P.named_tensors().items():
            g, r = named[name].grad, t.grad
            assert g is not None and r is not None, name
            err = float((g - r).norm() / r.norm().clamp_min(1e-30))
            # emitter.base: the oracle's cubemap_mip backward takes its texel directions from an fp32 table
            assert float(r.norm()) > 0 and err <= (1e-6 if name == "emitter.base" else 1e-8), (name, err)
