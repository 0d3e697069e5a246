"""tcgen05 building blocks: every operand role the fused MLP kernels use, checked against
fp64 matmul.  The 3-term bf16 split must deliver ~fp32 accuracy (<= 2e-6 of the row scale)."""
import numpy as np
import pytest
import torch

from rise_sdf_b200 import _lib as L

pytestmark = pytest.mark.gpu


def pad16(n):
    return (n + 15) // 16 * 16


def pack(W):
    N, K = W.shape
    blob = torch.empty(4 * pad16(N) * pad16(K), dtype=torch.uint8, device="cuda")
    L.call("rsdf_mlp_pack_weight", L.ptr(W), N, K, pad16(N), pad16(K), L.ptr(blob), L.stream())
    return blob


def err(C, ref):
    return float((C.double().cpu() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("S,K,N", [(128, 128, 128), (1000, 35, 128), (333, 128, 48), (4096, 67, 128), (257, 128, 3)])
def test_linear_forward_role(S, K, N):
    g = torch.Generator().manual_seed(S + K + N)
    A = torch.randn(S, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    C = torch.full((S, N), float("nan"), device="cuda")
    L.call("rsdf_tc_gemm_test", 0, L.ptr(A), L.ptr(pack(W)), None, L.ptr(C), S, K, N, pad16(N), pad16(K), 8, L.stream())
    torch.cuda.synchronize()
    assert err(C, A.double().cpu() @ W.double().cpu().T) <= 2e-6


@pytest.mark.parametrize("S,K,N", [(128, 128, 128), (777, 35, 128), (640, 128, 48)])
def test_transposed_weight_role(S, K, N):
    """C[S,K] = G[S,N] @ W[N,K] through the MN-major view of the SAME blob (backward data path)."""
    g = torch.Generator().manual_seed(7 * S + K + N)
    G = torch.randn(S, N, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    C = torch.full((S, K), float("nan"), device="cuda")
    L.call("rsdf_tc_gemm_test", 1, L.ptr(G), L.ptr(pack(W)), None, L.ptr(C), S, N, K, pad16(N), pad16(K), 8, L.stream())
    torch.cuda.synchronize()
    assert err(C, G.double().cpu() @ W.double().cpu()) <= 2e-6


@pytest.mark.parametrize("S,Fb", [(128, 128), (1000, 48), (5000, 80), (3000, 16)])
def test_weight_gradient_role(S, Fb):
    """C[128,Fb] = X[S,128]^T @ Y[S,Fb]: both operands MN-major, fp32 accumulation in TMEM across
    tiles, several CTAs reduced with atomics."""
    g = torch.Generator().manual_seed(S + Fb)
    X = torch.randn(S, 128, generator=g).cuda()
    Y = torch.randn(S, Fb, generator=g).cuda()
    C = torch.zeros(128, Fb, device="cuda")
    L.call("rsdf_tc_gemm_test", 2, L.ptr(X), None, L.ptr(Y), L.ptr(C), S, 128, Fb, 0, 0, 4, L.stream())
    torch.cuda.synchronize()
    assert err(C, X.double().cpu().T @ Y.double().cpu()) <= 1e-5   # fp32 accumulation over up to 40 tiles


# ------------------------------------------------------------------ fused forward MLP chain
def _mlp(dim_in, dim_out, hidden, sphere, n_neurons=128):
    from rise_sdf_b200.network_utils import VanillaMLP
    torch.manual_seed(dim_in + dim_out)
    m = VanillaMLP(dim_in, dim_out, {"n_neurons": n_neurons, "n_hidden_layers": hidden, "sphere_init": sphere,
                                     "weight_norm": sphere, "output_activation": "none"}).cuda()
    if sphere:
        with torch.no_grad():
            m.layers[0].weight_v[:, 3:].normal_(0.0, 0.05)
    return m


@pytest.mark.parametrize("dim_in,dim_out,hidden,sphere,S,width", [
    (35, 48, 2, True, 5000, 128),     # geometry MLP (Softplus-100, weight-norm)
    (67, 3, 4, False, 4097, 128),     # neus radiance MLP
    (84, 6, 4, False, 1000, 128),     # split albedo
    (84, 1, 2, False, 777, 128),      # split roughness
    (73, 3, 4, False, 300, 128),      # split env
    (35, 48, 2, True, 640, 64),       # width-64 variant named by the north star
])
def test_fused_mlp_forward(dim_in, dim_out, hidden, sphere, S, width):
    from rise_sdf_b200.fused_mlp import PackedMLP
    m = _mlp(dim_in, dim_out, hidden, sphere, width)
    g = torch.Generator().manual_seed(S)
    x = (torch.rand(S, dim_in, generator=g) * 2 - 1).cuda()
    if sphere:
        x[:, 3:] *= 0.1
    with torch.no_grad():
        ref = m.double()(x.double()).double().cpu() if False else None
    m64 = _mlp(dim_in, dim_out, hidden, sphere, width).double()
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    with torch.no_grad():
        h = x.double()
        for layer in m64.layers:
            h = layer(h)
        ref = h.cpu()
    out = PackedMLP(m)(x)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((out.double().cpu() - ref).abs().max()) <= 2e-6 * max(scale, 1.0)
    # segmented inputs + affine staging + sigmoid output == cat + torch
    a, b = x[:, :10].contiguous(), x[:, 10:].contiguous()
    out2 = PackedMLP(m, out_act="sigmoid")([a, b], scales=[2.0, 1.0], shifts=[-1.0, 0.0])
    with torch.no_grad():
        h = torch.cat([a.double() * 2 - 1, b.double()], -1)
        for layer in m64.layers:
            h = layer(h)
        ref2 = torch.sigmoid(h).cpu()
    assert float((out2.double().cpu() - ref2).abs().max()) <= 2e-6


# ------------------------------------------------------------------ training-path GEMMs + autograd
def test_mm_primitives_first_and_second_order():
    """mm_nt / mm_nn / mm_tn vs fp64 torch matmul: values, gradients and gradients of gradients."""
    from rise_sdf_b200 import tc_autograd as tca
    g = torch.Generator().manual_seed(0)
    S, K, N = 3000, 35, 128
    x = torch.randn(S, K, generator=g); W = torch.randn(N, K, generator=g) * 0.3
    c = torch.randn(S, K, generator=g); gy = torch.randn(S, N, generator=g)

    def run(xx, WW, f_nt):
        y = f_nt(xx, WW)
        (gx,) = torch.autograd.grad((torch.tanh(y) * gy.to(y)).sum(), xx, create_graph=True)
        loss = (gx * c.to(gx)).sum() + (y ** 2).sum() * 0.1
        gW, gx2 = torch.autograd.grad(loss, [WW, xx])
        return y.detach(), gx.detach(), gW, gx2

    xc, Wc = x.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
    got = run(xc, Wc, tca.mm_nt)
    x0, W0 = x.double().requires_grad_(True), W.double().requires_grad_(True)
    want = run(x0, W0, lambda a, b: a @ b.T)
    for a, b, name in zip(got, want, ("y", "gx", "gW(2nd order)", "gx(2nd order)")):
        e = float((a.double().cpu() - b).abs().max() / b.abs().max())
        assert e <= 5e-6, (name, e)
    # mm_tn and mm_nn directly
    A = torch.randn(5000, 128, generator=g).cuda(); B = torch.randn(5000, 48, generator=g).cuda()
    assert err(tca.mm_tn(A, B), A.double().cpu().T @ B.double().cpu()) <= 1e-5
    assert err(tca.mm_tn(B, A), B.double().cpu().T @ A.double().cpu()) <= 1e-5     # Fa < 128 (padded M)
    Wn = torch.randn(48, 128, generator=g).cuda()
    assert err(tca.mm_nn(B, Wn), B.double().cpu() @ Wn.double().cpu()) <= 2e-6
    assert tca.mm_nt(torch.zeros(0, 35, device="cuda"), Wc).shape == (0, 128)


def test_mm_stream_ragged_and_large():
    from rise_sdf_b200 import tc_autograd as tca
    g = torch.Generator().manual_seed(1)
    for S in (1, 127, 129, 148 * 128 + 5, 300000):
        x = torch.randn(S, 67, generator=g).cuda(); W = torch.randn(128, 67, generator=g).cuda()
        assert err(tca.mm_nt(x, W), x.double().cpu() @ W.double().cpu().T) <= 2e-6, S


@pytest.mark.parametrize("S,K,N", [(128, 128, 128), (1000, 35, 128), (333, 128, 48), (257, 64, 16)])
def test_a_operand_from_tensor_memory(S, K, N):
    """tcgen05.mma with the A operand in TMEM (lane = row, 32-bit column j = fp16 pair k = 2j, 2j+1, written with
    tcgen05.st): same product as the shared-memory form, C = A W^T, to fp32-class accuracy."""
    g = torch.Generator().manual_seed(3 * S + K + N)
    A = torch.randn(S, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    C = torch.full((S, N), float("nan"), device="cuda")
    L.call("rsdf_tc_gemm_test", 3, L.ptr(A), L.ptr(pack(W)), None, L.ptr(C), S, K, N, pad16(N), pad16(K), 8, L.stream())
    torch.cuda.synchronize()
    assert err(C, A.double().cpu() @ W.double().cpu().T) <= 2e-6
