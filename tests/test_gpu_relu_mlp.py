"""Per-layer ReLU-MLP training kernels (csrc/relu_mlp.cu) against fp64 torch: forward values and
first-order gradients w.r.t. inputs, weights and biases for every texture-network shape of the two
configs (models/texture.py:15-41,234-434)."""
import pytest
import torch
import torch.nn.functional as F

from rise_sdf_b200 import relu_mlp
from rise_sdf_b200.network_utils import VanillaMLP

pytestmark = pytest.mark.gpu


def make(dim_in, dim_out, hidden, seed=0):
    torch.manual_seed(seed)
    m = VanillaMLP(dim_in, dim_out, {"n_neurons": 128, "n_hidden_layers": hidden, "sphere_init": False,
                                     "weight_norm": False, "output_activation": "none"}).cuda()
    with torch.no_grad():
        for lin in m.linears():
            lin.bias.normal_(0.0, 0.05)
    return m


def ref64(m, x):
    ws = [(W.double(), b.double()) for W, b in m.effective_weights()]
    h = x
    for i, (W, b) in enumerate(ws):
        h = F.linear(h, W, b)
        if i + 1 < len(ws):
            h = F.relu(h)
    return h


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / max(float(b.abs().max()), 1e-30))


def bad_rows(a, b, tol):
    """Fraction of sample rows whose error exceeds tol * max|b|.  A hidden unit whose pre-activation is
    within fp32 rounding of 0 takes the other ReLU branch than in the fp64 reference (~1e-6 of the units:
    a handful per test); such a flip changes that one sample's input gradient by O(10 %) and is inherent
    to ANY fp32 evaluation, so per-sample quantities are checked as 'all but a few rows'."""
    e = (a.double().cpu() - b.double().cpu()).abs().amax(dim=-1)
    return float((e > tol * float(b.abs().max())).double().mean())


@pytest.mark.parametrize("dim_in,dim_out,hidden,S", [
    (67, 3, 4, 5000),            # neus radiance
    (84, 6, 4, 1000),            # split albedo
    (84, 1, 2, 777),             # split roughness
    (73, 3, 4, 64 * 148 * 2 + 5),   # split env, several tiles per CTA
    (76, 3, 4, 63),              # split secondary, single ragged tile
    (16, 2, 1, 300),
])
def test_relu_mlp_forward_backward(dim_in, dim_out, hidden, S):
    m = make(dim_in, dim_out, hidden, seed=S)
    g = torch.Generator().manual_seed(S)
    x = torch.randn(S, dim_in, generator=g).cuda().requires_grad_(True)
    c = (torch.randn(S, dim_out, generator=g) / S).cuda()
    out = relu_mlp.relu_mlp(m, x)
    params = list(m.parameters())
    got = torch.autograd.grad((torch.sigmoid(out) * c).sum(), [x] + params)
    torch.cuda.synchronize()
    x64 = x.detach().double().requires_grad_(True)
    ro = ref64(m, x64)
    want = torch.autograd.grad((torch.sigmoid(ro) * c.double()).sum(), [x64] + params)
    assert rel(out, ro.detach()) <= 3e-6
    assert bad_rows(got[0], want[0], 2e-5) <= 2e-3, ("x", bad_rows(got[0], want[0], 2e-5))
    for n, a, b in zip([n for n, _ in m.named_parameters()], got[1:], want[1:]):
        assert rel(a, b) <= 2e-3, (n, rel(a, b))      # sums over samples: a flipped unit moves them by O(1/S)
    # and away from flipped units the sums are fp32-exact: compare on the median element
    for a, b in zip(got[1:], want[1:]):
        e = (a.double().cpu() - b.double().cpu()).abs() / max(float(b.abs().max()), 1e-30)
        assert float(e.median()) <= 2e-6


def test_relu_mlp_segments_and_affine():
    m = make(67, 3, 4, seed=3)
    g = torch.Generator().manual_seed(3)
    S = 2000
    feat = torch.randn(S, 48, generator=g).cuda().requires_grad_(True)
    d = torch.rand(S, 16, generator=g).cuda()
    n = torch.randn(S, 3, generator=g).cuda().requires_grad_(True)
    out = relu_mlp.relu_mlp(m, [feat, d, n], scales=[1.0, 2.0, 1.0], shifts=[0.0, -1.0, 0.0])
    gf, gn = torch.autograd.grad(out.sum(), [feat, n])
    f64, n64 = feat.detach().double().requires_grad_(True), n.detach().double().requires_grad_(True)
    ro = ref64(m, torch.cat([f64, d.double() * 2 - 1, n64], -1))
    rf, rn = torch.autograd.grad(ro.sum(), [f64, n64])
    assert rel(out, ro.detach()) <= 3e-6
    assert bad_rows(gf, rf, 2e-5) <= 2e-3 and bad_rows(gn, rn, 2e-5) <= 2e-3
    assert relu_mlp.relu_mlp(m, [feat[:0], d[:0], n[:0]]).shape == (0, 3)
