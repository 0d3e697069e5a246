"""tcgen05 building blocks: every operand role the fused MLP kernels use, checked against
fp64 matmul.  The 3-term bf16 split must deliver ~fp32 accuracy (<= 3e-5 of the row scale)."""
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
    assert err(C, A.double().cpu() @ W.double().cpu().T) <= 3e-5


@pytest.mark.parametrize("S,K,N", [(128, 128, 128), (777, 35, 128), (640, 128, 48)])
def test_transposed_weight_role(S, K, N):
    """C[S,K] = G[S,N] @ W[N,K] through the MN-major view of the SAME blob (backward data path)."""
    g = torch.Generator().manual_seed(7 * S + K + N)
    G = torch.randn(S, N, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    C = torch.full((S, K), float("nan"), device="cuda")
    L.call("rsdf_tc_gemm_test", 1, L.ptr(G), L.ptr(pack(W)), None, L.ptr(C), S, N, K, pad16(N), pad16(K), 8, L.stream())
    torch.cuda.synchronize()
    assert err(C, G.double().cpu() @ W.double().cpu()) <= 3e-5


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
    assert err(C, X.double().cpu().T @ Y.double().cpu()) <= 3e-5
