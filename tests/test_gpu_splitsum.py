"""K5 parity: texture lookups vs the torch oracle (Appendix A.3), cube-map prefilter vs the
reference's own compiled kernels (oracle/_ref/renderutils_plugin.so) and the numpy restatement."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref as oref
from oracle import textures as ot
from rise_sdf_b200 import nvdiffrast as dr
from rise_sdf_b200 import renderutils as ru

pytestmark = pytest.mark.gpu


def unit_dirs(n, seed=0, edges=True):
    g = torch.Generator().manual_seed(seed)
    d = F.normalize(torch.randn(n, 3, generator=g), dim=-1)
    if edges:   # face centres, edges and corners
        sp = torch.tensor([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                           [1, 1, 0.01], [1, -0.999, 0.3], [-1, 0.2, 1.0001], [1, 1, 1], [-1, 1, -1], [0.999, 1, -1],
                           [1, 0.97, 0.99], [-0.98, -1, 0.99]], dtype=torch.float32)
        d[: len(sp)] = F.normalize(sp, dim=-1)
    return d


def test_tex2d_fwd_bwd():
    g = torch.Generator().manual_seed(1)
    tex = torch.rand(1, 256, 256, 2, generator=g)
    uv = torch.rand(4096, 2, generator=g)
    uv[:8] = torch.tensor([[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 0.5], [0.001, 0.999], [1 / 512, 1 / 512], [0.25, 1.0]])
    tc, uc = tex.cuda().requires_grad_(True), uv.cuda().requires_grad_(True)
    out = dr.texture(tc, uc.view(1, -1, 1, 2), filter_mode="linear", boundary_mode="clamp").view(-1, 2)
    t0, u0 = tex[0].double().requires_grad_(True), uv.double().requires_grad_(True)
    ref = ot.tex2d(t0, u0)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-6
    go = torch.randn(4096, 2, generator=g)
    (out * go.cuda()).sum().backward(); (ref * go.double()).sum().backward()
    assert float((tc.grad[0].cpu() - t0.grad).abs().max()) <= 1e-4 * float(t0.grad.abs().max())
    interior = ((uv * 256 - 0.5) - torch.floor(uv * 256 - 0.5) - 0.5).abs().max(-1).values < 0.49
    assert float((uc.grad.cpu() - u0.grad).abs()[interior].max()) <= 1e-3 * float(u0.grad.abs().max())


def test_tex2d_wrap_fwd_bwd():
    """boundary_mode='wrap' (nvdiffrast's default; the lat-long -> cube conversion,
    lib/pbr/utils/light_utils.py:126-139): coordinates outside [0,1) and taps across the u = 0/1 seam."""
    g = torch.Generator().manual_seed(2)
    tex = torch.rand(1, 32, 64, 3, generator=g)
    uv = torch.rand(4096, 2, generator=g) * 3.0 - 1.0
    uv[:8] = torch.tensor([[0, 0], [1, 1], [0.999, 0.5], [0.001, 0.5], [0.5, 0.001], [0.5, 0.999], [-0.25, 1.75],
                           [1 / 128, 1 / 64]])
    tc, uc = tex.cuda().requires_grad_(True), uv.cuda().requires_grad_(True)
    out = dr.texture(tc, uc.view(1, -1, 1, 2), filter_mode="linear").view(-1, 3)          # default boundary mode
    t0, u0 = tex[0].double().requires_grad_(True), uv.double().requires_grad_(True)
    ref = ot.tex2d(t0, u0, wrap=True)
    clamp = ot.tex2d(tex[0].double(), uv.double() - torch.floor(uv.double()))
    assert float((ref.detach() - clamp).abs().max()) > 0.05           # the seam taps really differ from clamping
    frac = (uv * torch.tensor([64.0, 32.0]) - 0.5) - torch.floor(uv * torch.tensor([64.0, 32.0]) - 0.5)
    interior = (frac - 0.5).abs().max(-1).values < 0.49               # fp32 uv next to a texel centre may pick the other cell
    assert float((out.detach().cpu() - ref.detach()).abs()[interior].max()) <= 1e-5
    go = torch.randn(4096, 3, generator=g) * interior[:, None]
    (out * go.cuda()).sum().backward(); (ref * go.double()).sum().backward()
    assert float((tc.grad[0].cpu() - t0.grad).abs().max()) <= 1e-4 * float(t0.grad.abs().max())
    assert float((uc.grad.cpu() - u0.grad).abs()[interior].max()) <= 1e-3 * float(u0.grad.abs().max())


@pytest.mark.parametrize("N", [16, 64])
def test_cube_linear_fwd_bwd(N):
    g = torch.Generator().manual_seed(N)
    tex = torch.rand(6, N, N, 3, generator=g)
    d = unit_dirs(5000, seed=N)
    tc, dc = tex.cuda().requires_grad_(True), d.cuda().requires_grad_(True)
    out = dr.texture(tc[None], dc.view(1, -1, 1, 3), filter_mode="linear", boundary_mode="cube").view(-1, 3)
    t0, d0 = tex.double().requires_grad_(True), d.double().requires_grad_(True)
    ref = ot.cube_linear(t0, d0)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 1e-5
    go = torch.randn(5000, 3, generator=g)
    (out * go.cuda()).sum().backward(); (ref * go.double()).sum().backward()
    assert float((tc.grad.cpu() - t0.grad).abs().max()) <= 1e-4 * float(t0.grad.abs().max())
    face, u, v = ot.dir_to_face(d)
    tx, ty = (u + 1) * 0.5 * N - 0.5, (v + 1) * 0.5 * N - 0.5
    fr = torch.stack([tx - torch.floor(tx), ty - torch.floor(ty)], -1)
    interior = ((fr - 0.5).abs().max(-1).values < 0.49) & (torch.stack([u, v], -1).abs().max(-1).values < 1 - 2.0 / N)
    err = (dc.grad.cpu() - d0.grad).abs()[interior].max()
    assert float(err) <= 2e-3 * float(d0.grad.abs()[interior].max())


def test_cube_mip_trilinear_fwd_bwd():
    g = torch.Generator().manual_seed(5)
    levels = [torch.rand(6, r, r, 3, generator=g) for r in (64, 32, 16)]
    d = unit_dirs(4000, seed=9)
    bias = torch.rand(4000, generator=g) * 2.4 - 0.2          # also exercises both clamps
    lc = [t.cuda().requires_grad_(True) for t in levels]
    bc = bias.cuda().requires_grad_(True)
    out = dr.texture(lc[0][None], d.cuda().view(1, -1, 1, 3), mip=[t[None] for t in lc[1:]],
                     mip_level_bias=bc.view(1, -1, 1), filter_mode="linear-mipmap-linear", boundary_mode="cube").view(-1, 3)
    l0 = [t.double().requires_grad_(True) for t in levels]
    b0 = bias.double().requires_grad_(True)
    ref = ot.cube_sample(l0, d.double(), b0)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 1e-5
    go = torch.randn(4000, 3, generator=g)
    (out * go.cuda()).sum().backward(); (ref * go.double()).sum().backward()
    for a, b in zip(lc, l0):
        assert float((a.grad.cpu() - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())
    assert float((bc.grad.cpu() - b0.grad).abs().max()) <= 1e-4 * float(b0.grad.abs().max())


def test_face_convention_roundtrip():
    """texel centres of face s map back to face s / the same texel (pins cube_to_dir <-> lookup)."""
    N = 16
    tex = torch.arange(6 * N * N, dtype=torch.float32).view(6, N, N, 1).expand(6, N, N, 3).contiguous()
    dirs = torch.from_numpy(ot.texel_dirs(N)).reshape(-1, 3)
    out = dr.texture(tex.cuda()[None], dirs.cuda().view(1, -1, 1, 3), filter_mode="linear", boundary_mode="cube")
    assert float((out.view(-1, 3)[:, 0].cpu() - torch.arange(6 * N * N)).abs().max()) <= 1e-2


@pytest.mark.parametrize("N", [16])
def test_diffuse_cubemap_vs_reference(N):
    g = torch.Generator().manual_seed(3)
    cube = (torch.rand(6, N, N, 3, generator=g) * 0.5 + 0.25)
    cc = cube.cuda().requires_grad_(True)
    out = ru.diffuse_cubemap(cc)
    go = torch.randn(6, N, N, 3, generator=g)
    (out * go.cuda()).sum().backward()
    ref = ot.diffuse_cubemap(cube.double())
    assert float((out.detach().cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    W = torch.from_numpy(ot.diffuse_weights(N))
    gref = (W.T @ go.double().reshape(-1, 3)).reshape(6, N, N, 3)
    assert float((cc.grad.cpu() - gref).abs().max()) <= 1e-4 * float(gref.abs().max())
    P = oref.renderutils_plugin()
    if P is not None:
        r_out = P.diffuse_cubemap_fwd(cube.cuda())
        r_g = P.diffuse_cubemap_bwd(cube.cuda(), go.cuda().contiguous())
        assert float((out.detach() - r_out).abs().max()) <= 1e-5 * float(r_out.abs().max())
        assert float((cc.grad - r_g).abs().max()) <= 3e-4 * float(r_g.abs().max())  # fp32 atomics (order varies) vs gather


@pytest.mark.parametrize("cached", [False, True])
@pytest.mark.parametrize("N,roughness", [(16, 1.0), (32, 0.5), (32, 0.08), (64, 0.185), (128, 0.29), (256, 0.08)])
def test_specular_cubemap_vs_reference(N, roughness, cached, monkeypatch):
    """`cached`: the prefilter as a cached sparse operator (pair weights evaluated once, streamed every step) against
    the direct kernel: both must match the reference's compiled kernels."""
    monkeypatch.setattr(ru, "OPERATOR_CACHE", cached)
    monkeypatch.setattr(ru, "OPERATOR_CACHE_MIN_RES", 16)
    g = torch.Generator().manual_seed(N)
    cube = (torch.rand(6, N, N, 3, generator=g) * 0.5 + 0.25)
    cutoff = ru.ndf_cutoff(roughness, 0.99)
    assert cutoff == ot.ndf_cutoff(roughness, 0.99)
    bounds = ru.specular_bounds(N, cutoff, "cuda")
    cc = cube.cuda().requires_grad_(True)
    out = ru.specular_cubemap(cc, roughness, 0.99)
    go = torch.randn(6, N, N, 3, generator=g)
    (out * go.cuda()).sum().backward()
    P = oref.renderutils_plugin()
    if P is not None:
        rb = P.specular_bounds(N, cutoff)
        assert torch.equal(bounds.cpu(), rb.cpu()), "lobe bounds differ from the reference kernel"
        rc = cube.cuda().requires_grad_(True)
        raw = P.specular_cubemap_fwd(rc.detach(), rb, roughness, cutoff)
        r_out = raw[..., 0:3] / raw[..., 3:]
        assert float((out.detach() - r_out).abs().max()) <= 1e-5 * float(r_out.abs().max())
        # reference backward wrt the raw 4-channel output: chain the normalisation by hand
        g4 = torch.cat([go.cuda() / raw[..., 3:], torch.zeros(6, N, N, 1, device="cuda")], -1).contiguous()
        r_g = P.specular_cubemap_bwd(rc.detach(), rb, g4, roughness, cutoff)
        assert float((cc.grad - r_g).abs().max()) <= 3e-4 * float(r_g.abs().max())  # fp32 atomics (order varies) vs gather
    if N <= 32:
        ob = ot.specular_bounds(N, cutoff)
        assert np.array_equal(bounds.cpu().numpy(), ob), "lobe bounds differ from the numpy restatement"
        c0 = cube.double().requires_grad_(True)
        ref = ot.specular_cubemap(c0, roughness, 0.99, bounds=ob)
        (ref * go.double()).sum().backward()
        assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-5 * float(ref.abs().max())
        assert float((cc.grad.cpu() - c0.grad).abs().max()) <= 1e-4 * float(c0.grad.abs().max())


# ------------------------------------------------------------------ end-to-end split-sum render
def _split_model(base_res=64, seed=0):
    from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
    torch.manual_seed(0)
    cfg = split_mixed_occ_config()
    cfg["light"]["envlight_config"]["base_res"] = base_res
    m = SplitMixedOCCModel(cfg).cuda()
    with torch.no_grad():
        m.geometry.network.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
        m.variance.variance.fill_(0.5)
        m.geometry.encoding.encoding.encoding.params.uniform_(-0.02, 0.02)
    return m


def _sample_set_difference(m, rays, ref, keep):
    """Samples the product keeps and the fp64 oracle drops (or vice versa), with the oracle's transmittance at each."""
    from oracle import fields as ofields
    ro, rd = rays[:, :3].contiguous().cuda(), rays[:, 3:].contiguous().cuda()
    with torch.no_grad():
        ri, ts, te = m.occupancy_grid.sampling(ro, rd, alpha_fn=m._alpha_fn(ro, rd), render_step_size=m.render_step_size,
                                               stratified=False)
    ours = set(zip(ri.cpu().tolist(), ts.cpu().numpy().view(np.uint32).tolist()))
    theirs = set(zip(ref["ray_indices"].tolist(), ref["t_starts"].numpy().view(np.uint32).tolist()))
    _, T = ofields.render_weight_from_alpha(keep["alphas"].double(), keep["ri"], rays.shape[0])
    cand = {(int(r), int(b)): float(t) for r, b, t in zip(keep["ri"].tolist(), keep["ts"].numpy().view(np.uint32).tolist(),
                                                            T.tolist())}
    return ours, theirs, cand


@pytest.mark.parametrize("relighting", [False, True])
def test_split_sum_render_vs_oracle(relighting):
    """configs[2]/[3] shape at 192 rays, stage 1 (+ third bounce when relighting), eval mode (fused tcgen05 MLPs), full
    512^2 environment light, against oracle/split.py evaluated in float64.
    The north-star's 1e-4 is not reachable by ANY fp32 evaluation of this configuration: finite-difference normals
    amplify the fp32 rounding of the SDF by 1/(2 eps) ~ 1400 and the reflection bounce starts at the surface.  The
    bounds are therefore derived here, per channel, from an fp32 twin of the oracle (same code, float32):
    `floor` = |oracle32 - oracle64|; the product must stay within 1e-4 or 3x floor (mean / p99), 5x floor (max).
    Sample set: identical to the fp64 oracle's except for candidates whose transmittance sits at the early-stop
    threshold (checked: every differing sample has T within a factor 2 of 1e-4)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import split_oracle_params
    from oracle import split as osplit
    from rise_sdf_b200 import synthetic as syn
    m = _split_model(base_res=512).eval()
    m.update_step(0, 20000)
    assert m.stage == 1
    with torch.no_grad():
        m.emitter.build_mips()
    rays, _, _, bg = syn.training_rays(192, seed=3)
    m.background_color = bg.cuda()
    grid = syn.analytic_grid("ball")
    m.occupancy_grid.binaries = grid[None].cuda()
    m.render_step_size = 1.732 * 2 * 1.5 / 256
    with torch.no_grad():
        out = m(rays.cuda(), relighting=relighting)

    def oracle(dtype, keep=None):
        P = split_oracle_params(m).to(dtype)
        # prefiltered levels from the product (pinned to the reference's compiled kernels by the prefilter tests; the
        # dense CPU prefilter cannot run at 512^2)
        P.specular = [t.detach().cpu().to(dtype) for t in m.emitter.specular]
        P.diffuse = m.emitter.diffuse.detach().cpu().to(dtype)
        return osplit.forward(P, rays, grid.numpy(), m.render_step_size, stage=1, relighting=relighting, background=bg,
                              dtype=dtype, keep=keep)

    keep = {}
    ref = oracle(torch.float64, keep)
    ref32 = oracle(torch.float32)
    ours, theirs, cand = _sample_set_difference(m, rays, ref, keep)
    diff = ours ^ theirs
    assert int(out["num_samples"].sum()) == len(ours)
    assert len(diff) <= 8 and all(0.5e-4 <= cand[k] <= 2e-4 for k in diff), sorted(cand[k] for k in diff)
    report = []
    for k in ("comp_normal", "opacity", "depth", "comp_albedo", "comp_roughness", "comp_metallic", "comp_rgb",
              "comp_rgb_phys", "comp_rgb_full", "comp_rgb_phys_full"):
        scale = max(float(ref[k].abs().max()), 1.0)
        e = (out[k].cpu().double() - ref[k]).abs().max(-1).values / scale
        f = (ref32[k].double() - ref[k]).abs().max(-1).values / scale
        q = lambda x: (float(x.mean()), float(torch.quantile(x, 0.99)), float(x.max()))
        (em, e99, ex), (fm, f99, fx) = q(e), q(f)
        report.append(f"{k}: product mean/p99/max {em:.1e}/{e99:.1e}/{ex:.1e}   fp32-oracle floor {fm:.1e}/{f99:.1e}/{fx:.1e}")
        assert em <= max(1e-4, 3 * fm) and e99 <= max(1e-4, 3 * f99) and ex <= max(1e-4, 5 * fx), report[-1]
    print("\n".join(report))
    assert out["comp_rgb_phys"].shape == (192, 3) and out["comp_rgb_full"].min() >= 0 and out["comp_rgb_full"].max() <= 1


def test_split_sum_training_step_runs_and_reaches_every_parameter_group():
    from rise_sdf_b200 import synthetic as syn
    m = _split_model().train()
    m.update_step(0, 20000)
    m.randomized = False
    m.emitter.build_mips()
    rays, rgb, _, bg = syn.training_rays(256, seed=4)
    m.background_color = bg.cuda()
    m.occupancy_grid.binaries = syn.analytic_grid("ball")[None].cuda()
    m.render_step_size = 1.732 * 2 * 1.5 / 256
    out = m(rays.cuda())
    loss = ((out["comp_rgb_full"] - rgb.cuda()) ** 2).mean() + ((out["comp_rgb_phys_full"] - rgb.cuda()) ** 2).mean() \
        + 0.1 * ((out["sdf_grad_samples"].norm(dim=-1) - 1) ** 2).mean() + 0.01 * out["sdf_laplace_samples"].mean() \
        + out["normals_orientation_loss_map"].mean()
    loss.backward()
    for name in ("geometry.encoding.encoding.encoding.params", "emitter.base", "variance.variance",
                 "texture.albedo_network.layers.0.weight", "texture.roughness_network.layers.0.weight",
                 "texture.env_network.layers.0.weight", "texture.metallic_network.layers.0.weight",
                 "texture.secondary_network.layers.0.weight", "geometry.network.layers.0.weight_v"):
        g = dict(m.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0, name


def test_relighting_reuse_is_bit_identical(monkeypatch):
    """configs[3]: (a) a no-grad render gathers the surviving samples' field values from the visibility pass
    instead of evaluating the field again, (b) a tile rendered under several env maps computes everything that
    does not depend on the map once, (c) the visibility pass walks the candidates front to back and stops at
    rays that are already opaque.  All three must reproduce the evaluate-everything-again order of the reference
    (models/split_mixed_occ.py:197-240, systems/split_occ.py:331-458) bit for bit."""
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.relight import EnvSet, render_frame_shard
    m = _split_model().eval()
    m.update_step(0, 20000)
    m.background_color = torch.ones(3, device="cuda")
    m.occupancy_grid.binaries = syn.analytic_grid("ball")[None].cuda()
    m.render_step_size = 1.732 * 2 * 1.5 / 256
    g = torch.Generator().manual_seed(5)
    envs = EnvSet(m, [torch.rand(32, 64, 3, generator=g) * 2.0, torch.rand(32, 64, 3, generator=g) ** 4 * 20.0])
    rays = syn.training_rays(600, seed=9)[0].cuda()
    keys = ("comp_rgb_phys_full", "comp_rgb_full", "comp_normal", "opacity", "depth", "comp_spec_rgb_phys",
            "comp_roughness")
    from rise_sdf_b200 import nerfacc as rn
    chunks = rn.VISIBILITY_CHUNKS
    m.reuse_sampling_pass = False
    monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", ())           # one-shot visibility pass: the reference's order
    base, tiles = render_frame_shard(m, rays, envs, tile=256, keys=keys, share_across_envs=False)
    assert len(tiles) == 3
    assert not torch.equal(base[0]["comp_rgb_phys_full"], base[1]["comp_rgb_phys_full"])    # the maps do differ
    for front_to_back, reuse, share in ((False, True, False), (False, False, True), (True, False, False),
                                        (True, True, True)):
        monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", (8, 16) if front_to_back else ())
        m.reuse_sampling_pass = reuse
        out, _ = render_frame_shard(m, rays, envs, tile=256, keys=keys, share_across_envs=share)
        assert m._tile_cache is None
        for e in base:
            for k in keys:
                assert torch.equal(out[e][k], base[e][k]), (front_to_back, reuse, share, e, k)
    monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", chunks)
    # multi-GPU sharding (run here rank by rank): the interleaved shards assemble into the single-GPU frame
    for world in (2, 3):
        frame = torch.empty_like(base[1]["comp_rgb_phys_full"])
        for rank in range(world):
            shard, _ = render_frame_shard(m, rays, envs, rank=rank, world=world, tile=128, keys=keys)
            frame[rank::world] = shard[1]["comp_rgb_phys_full"]
        assert torch.equal(frame, base[1]["comp_rgb_phys_full"]), world
    # the training-mode secondary bounce (no-grad inside a grad-enabled step) reuses its alphas the same way
    m.train(); m.randomized = False
    outs = []
    for reuse in (False, True):
        m.reuse_sampling_pass = reuse
        torch.manual_seed(11)                     # the curvature probe draws random tangents
        outs.append(m(rays[:200]))
    for k in ("comp_rgb_full", "comp_rgb_phys_full", "opacity"):
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_relighting_recombination_follows_the_reference_chunks():
    """models/split_mixed_occ.py:332 sets spec_rgb_phys = spec_ref_map * spec_light_map for EVERY ray of a batch that
    holds a valid ray (opacity > 0.5) and leaves a batch without one alone; the reference's batches are its `ray_chunk`
    chunks (models/utils.py:14-51).  A large tile must reproduce the chunked render bit for bit -- including the
    silhouette rays of chunks that never see a valid ray."""
    from rise_sdf_b200 import synthetic as syn
    from rise_sdf_b200.neus import chunk_batch
    from rise_sdf_b200.relight import EnvSet
    m = _split_model().eval()
    with torch.no_grad():
        m.variance.variance.fill_(0.3)                     # a soft surface: plenty of rays with 0 < opacity <= 0.5
    m.update_step(0, 20000)
    m.background_color = torch.ones(3, device="cuda")
    m.occupancy_grid.binaries = syn.analytic_grid("ball")[None].cuda()
    m.render_step_size = 1.732 * 2 * 1.5 / 128
    envs = EnvSet(m, [torch.rand(32, 64, 3, generator=torch.Generator().manual_seed(5)) * 2.0])
    envs.use(0)
    # a 24-pixel-wide vertical strip of a frame, row-major: chunks of 4 rows walk across the silhouette
    frame = syn.frame_rays(3, syn.camera_poses(), syn.ray_directions()).view(800, 800, 6)
    rays = frame[::4, 388:412].reshape(-1, 6).contiguous().cuda()
    rc = 96
    m.config["ray_chunk"] = rc
    with torch.no_grad():
        chunked = chunk_batch(m.forward_, rc, False, rays, True)           # the reference's order
        whole = m.forward_(rays, relighting=True)
    op = chunked["opacity"][:, 0]
    per_chunk_valid = torch.nn.functional.pad(op > 0.5, (0, -len(op) % rc)).view(-1, rc).any(1)
    lonely = (~per_chunk_valid.repeat_interleave(rc)[:len(op)]) & (op > 0)
    assert int(lonely.sum()) > 0, "the strip must contain silhouette rays in a chunk without a valid ray"
    for k in ("comp_rgb_phys", "comp_spec_rgb_phys", "comp_rgb_phys_full", "comp_rgb", "opacity", "depth"):
        assert torch.equal(chunked[k], whole[k]), k
    # and the rule matters: recombining those rays as well gives different pixels
    m.config["ray_chunk"] = 1 << 30
    with torch.no_grad():
        naive = m.forward_(rays, relighting=True)
    assert not torch.equal(naive["comp_spec_rgb_phys"][lonely], whole["comp_spec_rgb_phys"][lonely])
