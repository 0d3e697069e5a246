"""§8f f3: FlatAdam (one `rsdf_adam_step` launch over flat buffers) against `torch.optim.Adam`, the optimizer
the reference builds in systems/utils.py:309-320, and the warm-up + exponential schedule of
configs/neus-blender.yaml:104-119 against torch's own `SequentialLR`."""
import math

import pytest
import torch


def _params(device, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1001, 2), (128, 35), (128, 1), (128,), (), (6, 9, 9, 3), (3,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in shapes]


def _groups(ps):
    return [{"params": ps[:4], "lr": 0.01}, {"params": ps[4:5], "lr": 0.001}, {"params": ps[5:], "lr": 0.02}]


@pytest.mark.gpu
@pytest.mark.parametrize("betas,eps", [((0.9, 0.99), 1e-15), ((0.9, 0.999), 1e-12)])
def test_flat_adam_matches_torch_adam(betas, eps):
    from rise_sdf_b200.optim import FlatAdam, warmup_exponential_scheduler
    a, b = _params("cuda"), _params("cuda")
    ref = torch.optim.Adam(_groups(a), lr=0.01, betas=betas, eps=eps, foreach=False)
    opt = FlatAdam(_groups(b), lr=0.01, betas=betas, eps=eps)
    sa = warmup_exponential_scheduler(ref, warmup_steps=5, max_steps=20)
    sb = warmup_exponential_scheduler(opt, warmup_steps=5, max_steps=20)
    assert all(p.data_ptr() == opt.flat_p.data_ptr() + 4 * off for p, (off, _) in zip(b, opt._slices))
    assert all(off % 64 == 0 for off, _ in opt._slices)
    g = torch.Generator().manual_seed(1)
    for it in range(12):
        opt.zero_grad()
        ref.zero_grad(set_to_none=True)
        for pa, pb in zip(a, b):
            # gradients over 12 decades, some exactly zero (untouched hash-table entries)
            gr = torch.randn(pa.shape, generator=g) * 10.0 ** float(torch.randint(-9, 3, (1,), generator=g))
            gr = torch.where(torch.rand(pa.shape, generator=g) < 0.3, torch.zeros(()), gr).cuda()
            pa.grad = gr.clone()
            pb.grad.copy_(gr)
        v0 = b[0]._version
        ref.step(); opt.step()
        sa.step(); sb.step()
        assert b[0]._version > v0                     # packed-weight caches key on the version counter
        assert [g_["lr"] for g_ in ref.param_groups] == [g_["lr"] for g_ in opt.param_groups]
        for pa, pb in zip(a, b):
            scale = float(pa.abs().max()) + 1e-12
            assert float((pa - pb).abs().max()) <= 2e-6 * scale + 1e-5 * 0.02 * (it + 1), it   # + ulps of each update
    sd_a, sd_b = ref.state_dict(), opt.state_dict()
    assert sd_a["state"].keys() == sd_b["state"].keys()
    for k in sd_a["state"]:
        assert float(sd_a["state"][k]["step"]) == float(sd_b["state"][k]["step"]) == 12.0
        for key in ("exp_avg", "exp_avg_sq"):
            x, y = sd_a["state"][k][key], sd_b["state"][k][key]
            assert x.shape == y.shape
            assert float((x - y).abs().max()) <= 2e-6 * (float(x.abs().max()) + 1e-30)


@pytest.mark.gpu
def test_flat_adam_state_dict_round_trip_and_external_grads():
    from rise_sdf_b200.optim import FlatAdam
    a, b, c = _params("cuda"), _params("cuda"), _params("cuda")
    ref = torch.optim.Adam(_groups(a), lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    o1 = FlatAdam(_groups(b), lr=0.01, betas=(0.9, 0.99), eps=1e-15, zero_grad_in_step=True)
    g = torch.Generator().manual_seed(3)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in a] for _ in range(6)]
    for it in range(3):
        for pa, pb, gr in zip(a, b, grads[it]):
            pa.grad = gr.clone()
            pb.grad = gr.clone()              # a gradient tensor from outside the bucket: copied in by step()
        ref.step(); o1.step()
        assert float(o1.bucket.flat.abs().max()) == 0.0        # cleared in the same pass
    # resume in a fresh optimizer from torch.optim.Adam's own state dict (Lightning `optimizer_states`)
    with torch.no_grad():
        for pc, pa in zip(c, a):
            pc.copy_(pa)
    o2 = FlatAdam(_groups(c), lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    o2.load_state_dict(ref.state_dict())
    assert o2._t == 3 and o2.state[c[0]]["exp_avg"].data_ptr() == o2.flat_m.data_ptr()
    for it in range(3, 6):
        for pa, pc, gr in zip(a, c, grads[it]):
            pa.grad = gr.clone()
            pc.grad.copy_(gr)
        ref.step(); o2.step()
    for pa, pc in zip(a, c):
        assert float((pa - pc).abs().max()) <= 2e-6 * (float(pa.abs().max()) + 1e-12)


@pytest.mark.gpu
def test_flat_adam_keeps_empty_and_frozen_parameters_in_its_groups():
    """The reference's groups come from `module.parameters()` (systems/utils.py:309-320) and carry the empty tcnn
    SphericalHarmonics `params` (texture group) -- and may carry frozen tensors.  FlatAdam must keep them in
    `param_groups` so that group sizes and state-dict indices are torch.optim.Adam's: a Lightning `optimizer_states`
    entry loads, and what FlatAdam saves loads into torch.optim.Adam."""
    from rise_sdf_b200.optim import FlatAdam

    def make():
        ps = _params("cuda", seed=5)
        empty = torch.nn.Parameter(torch.zeros(0, device="cuda"))
        frozen = torch.nn.Parameter(torch.randn(7, device="cuda"), requires_grad=False)
        groups = [{"params": ps[:2] + [empty] + ps[2:4], "lr": 0.01}, {"params": [frozen] + ps[4:], "lr": 0.002}]
        return ps, groups

    (a, ga), (b, gb), (c, gc) = make(), make(), make()
    ref = torch.optim.Adam(ga, lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    opt = FlatAdam(gb, lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    assert [len(g["params"]) for g in opt.param_groups] == [len(g["params"]) for g in ref.param_groups] == [5, 4]
    g = torch.Generator().manual_seed(2)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in a] for _ in range(4)]
    for it in range(2):
        for pa, pb, gr in zip(a, b, grads[it]):
            pa.grad = gr.clone()
            pb.grad.copy_(gr)
        ref.step(); opt.step()
    sd_ref, sd_opt = ref.state_dict(), opt.state_dict()
    assert [g_["params"] for g_ in sd_ref["param_groups"]] == [g_["params"] for g_ in sd_opt["param_groups"]]
    assert sd_ref["state"].keys() == sd_opt["state"].keys() and 2 not in sd_opt["state"] and 5 not in sd_opt["state"]
    # torch's state dict -> FlatAdam, FlatAdam's -> torch: both continue identically
    with torch.no_grad():
        for pc, pa in zip(c, a):
            pc.copy_(pa)
    o2 = FlatAdam(gc, lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    o2.load_state_dict(sd_ref)
    r2 = torch.optim.Adam(make()[1], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    r2.load_state_dict(sd_opt)                              # must not raise "doesn't match the size"
    for it in range(2, 4):
        for pa, pc, gr in zip(a, c, grads[it]):
            pa.grad = gr.clone()
            pc.grad.copy_(gr)
        ref.step(); o2.step()
    for pa, pc in zip(a, c):
        assert float((pa - pc).abs().max()) <= 2e-6 * (float(pa.abs().max()) + 1e-12)


@pytest.mark.gpu
def test_flat_adam_late_activated_group_matches_torch():
    """`emitter.base` gets no gradient during stage 0 (10 000 steps): torch.optim.Adam skips it (grad None) and its
    first real update runs with step = 1 bias correction.  `step(inactive_groups=...)` reproduces that; the saved
    per-parameter `step` counts match torch's."""
    from rise_sdf_b200.optim import FlatAdam
    a, b = _params("cuda", seed=7), _params("cuda", seed=7)
    ref = torch.optim.Adam(_groups(a), lr=0.01, betas=(0.9, 0.999), eps=1e-12, foreach=False)
    opt = FlatAdam(_groups(b), lr=0.01, betas=(0.9, 0.999), eps=1e-12)
    g = torch.Generator().manual_seed(4)
    late = {id(p) for p in a[5:]}
    for it in range(7):
        off = it < 4                                           # group 2 is off the graph for the first four steps
        ref.zero_grad(set_to_none=True); opt.zero_grad()
        for pa, pb in zip(a, b):
            if off and id(pa) in late:
                continue
            gr = torch.randn(pa.shape, generator=g).cuda()
            pa.grad = gr.clone()
            pb.grad.copy_(gr)
        ref.step(); opt.step(inactive_groups=(2,) if off else ())
        for pa, pb in zip(a, b):
            assert float((pa - pb).abs().max()) <= 2e-6 * (float(pa.abs().max()) + 1e-12), it
    sd_a, sd_b = ref.state_dict(), opt.state_dict()
    assert {k: float(v["step"]) for k, v in sd_a["state"].items()} == {k: float(v["step"]) for k, v in sd_b["state"].items()}
    assert float(sd_b["state"][5]["step"]) == 3.0 and float(sd_b["state"][0]["step"]) == 7.0


def test_flat_adam_refuses_cpu_parameters():
    from rise_sdf_b200.optim import FlatAdam
    with pytest.raises(NotImplementedError):
        FlatAdam(_groups(_params("cpu")), lr=0.01)


def test_schedule_values():
    """LinearLR 0.01 -> 1 over `warmup_steps`, then lr * gamma^k with gamma^(max_steps - warmup_steps) = 0.1."""
    from rise_sdf_b200.optim import exp_lr_decay_rate, warmup_exponential_scheduler
    p = [torch.nn.Parameter(torch.zeros(3))]
    opt = torch.optim.Adam([{"params": p, "lr": 0.01}, {"params": [torch.nn.Parameter(torch.zeros(()))], "lr": 0.001}])
    sch = warmup_exponential_scheduler(opt, warmup_steps=500, max_steps=30000)
    gamma = exp_lr_decay_rate(0.1, 29500)
    assert math.isclose(gamma ** 29500, 0.1, rel_tol=1e-9)
    lrs = []
    for step in range(1000):
        lrs.append([g["lr"] for g in opt.param_groups])
        opt.step(); sch.step()
    assert math.isclose(lrs[0][0], 0.01 * 0.01, rel_tol=1e-9) and math.isclose(lrs[0][1], 0.001 * 0.01, rel_tol=1e-9)
    assert math.isclose(lrs[250][0], 0.01 * (0.01 + 0.99 * 250 / 500), rel_tol=1e-6)
    assert math.isclose(lrs[500][0], 0.01, rel_tol=1e-6)
    assert math.isclose(lrs[999][0], 0.01 * gamma ** 499, rel_tol=1e-6)
    assert math.isclose(lrs[999][1], 0.001 * gamma ** 499, rel_tol=1e-6)


@pytest.mark.parametrize("betas,eps", [((0.9, 0.99), 1e-15), ((0.9, 0.999), 1e-12)])
def test_adam_restatement_matches_torch_cpu(betas, eps):
    """oracle/adam.py (the arithmetic of csrc/optim.cu, operation by operation) against torch.optim.Adam on the host:
    gradients over 12 decades with exact zeros, 12 steps."""
    import numpy as np
    from oracle import adam as oadam
    g_ = torch.Generator().manual_seed(0)
    p0 = torch.randn(5000, generator=g_)
    p0[:100] *= 0.01
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=0.01, betas=betas, eps=eps, foreach=False)
    p, m, v = p0.numpy().copy(), np.zeros(5000, np.float32), np.zeros(5000, np.float32)
    for t in range(1, 13):
        gr = torch.randn(5000, generator=g_) * 10.0 ** float(torch.randint(-9, 3, (1,), generator=g_))
        gr = torch.where(torch.rand(5000, generator=g_) < 0.3, torch.zeros(()), gr)
        pt.grad = gr.clone()
        opt.step()
        p, m, v = oadam.adam_step(p, gr.numpy(), m, v, t, 0.01, betas[0], betas[1], eps)
        assert np.abs(p - pt.detach().numpy()).max() <= 5e-7            # a few ulps of |p| <= 4
    st = opt.state[pt]
    assert np.abs(m - st["exp_avg"].numpy()).max() <= 2e-7 * np.abs(m).max()
    assert np.abs(v - st["exp_avg_sq"].numpy()).max() <= 2e-7 * np.abs(v).max()


def test_masked_losses_equal_boolean_indexing():
    """train.masked_mse / masked_l1 (no host sync) against `F.mse_loss(a[valid], b[valid])` of systems/neus.py:103."""
    import torch.nn.functional as F
    from rise_sdf_b200.train import masked_l1, masked_mse
    g = torch.Generator().manual_seed(0)
    a = torch.randn(1000, 3, generator=g).requires_grad_(True)
    b = torch.randn(1000, 3, generator=g)
    valid = torch.rand(1000, generator=g) < 0.6
    for ours, ref in ((masked_mse, F.mse_loss), (masked_l1, F.l1_loss)):
        x, y = ours(a, b, valid), ref(a[valid], b[valid])
        assert abs(float(x) - float(y)) <= 1e-6 * abs(float(y))
        gx, = torch.autograd.grad(x, a)
        gy, = torch.autograd.grad(y, a)
        assert float((gx - gy).abs().max()) <= 1e-8 and float(gx[~valid].abs().max()) == 0.0
    assert torch.isnan(masked_mse(a, b, torch.zeros(1000, dtype=torch.bool)))      # mean of an empty selection
