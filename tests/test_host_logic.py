"""Host-side logic of the render path that needs no device: the front-to-back visibility rounds of
`nerfacc.ray_marching` (index bookkeeping, with the transmittance scan replaced by a serial Python product) and the
vectorised frequency encoding against the reference's per-band loop (models/network_utils.py:27-33)."""
import pytest
import torch

from rise_sdf_b200 import nerfacc as rn
from rise_sdf_b200.network_utils import VanillaFrequency


def _serial_transmittance(packed, alphas):
    a = alphas.reshape(-1)
    T = torch.empty_like(a)
    for b, c in packed.tolist():
        t = 1.0
        for j in range(b, b + c):
            T[j] = t
            t = t * (1.0 - float(a[j]))
    return T


@pytest.mark.parametrize("chunks", [(4,), (4, 8), (3, 5, 7), (64,)])
def test_front_to_back_rounds_bookkeeping(chunks, monkeypatch):
    """Ragged rays (empty ones included): whatever the round sizes, the samples the final mask keeps have been
    evaluated, carry the alpha a one-shot pass would give them, and `rows` points at them in call order."""
    monkeypatch.setattr(rn, "_transmittance", _serial_transmittance)
    monkeypatch.setattr(rn, "VISIBILITY_CHUNKS", chunks)
    g = torch.Generator().manual_seed(0)
    n_rays, eps = 60, 1e-2
    count = torch.randint(0, 40, (n_rays,), generator=g)
    count[3] = 0
    count[-1] = 0
    base = torch.cumsum(count, 0) - count
    packed = torch.stack([base, count], 1).int()
    S0 = int(count.sum())
    ri = torch.repeat_interleave(torch.arange(n_rays), count)
    ts = torch.arange(S0, dtype=torch.float32)[:, None]          # t_start doubles as the sample's identity
    te = ts + 0.5
    true_alpha = torch.rand(S0, generator=g) * 0.9
    true_alpha[base[5]:base[5] + count[5]] = 0.0                 # a ray that never goes opaque
    calls = []

    def alpha_fn(t0, t1, r):
        idx = t0[:, 0].long()
        assert torch.equal(r, ri[idx]) and torch.equal(t1, te[idx])
        calls.append(idx)
        return true_alpha[idx].reshape(-1, 1)

    alphas, rows = rn._alphas_front_to_back(alpha_fn, packed, ri, ts, te, eps)
    vis = _serial_transmittance(packed, true_alpha) >= eps
    assert torch.equal(_serial_transmittance(packed, alphas) >= eps, vis)          # same survivors
    assert torch.equal(alphas[vis, 0], true_alpha[vis]) and int(rows[vis].min()) >= 0
    evaluated = torch.cat(calls)
    assert torch.equal(evaluated[rows[vis]], torch.nonzero(vis)[:, 0])            # rows index the call-order concat
    assert evaluated.numel() == evaluated.unique().numel() <= S0                  # nothing evaluated twice
    assert torch.equal(rows >= 0, torch.zeros(S0, dtype=torch.bool).index_fill_(0, evaluated, True))
    assert len(calls) <= len(chunks) + 1
    if chunks[0] <= 8:
        assert evaluated.numel() < S0                                              # opaque rays left the later rounds


@pytest.mark.parametrize("cfg,step", [({"n_frequencies": 10}, None),
                                      ({"n_frequencies": 6, "n_masking_step": 100, "x_scale": 2.0, "x_offset": -1.0}, 37)])
def test_frequency_encoding_matches_the_per_band_loop(cfg, step):
    enc = VanillaFrequency(3, cfg)
    if step is not None:
        enc.update_step(0, step)
    x = torch.randn(500, 3, generator=torch.Generator().manual_seed(1)).requires_grad_(True)
    y = enc(x)
    xs = x * enc.x_scale + enc.x_offset
    out = []
    for freq, mask in zip(enc.freq_bands.tolist(), enc.mask.tolist()):           # models/network_utils.py:27-33
        out += [torch.sin(freq * xs) * mask, torch.cos(freq * xs) * mask]
    ref = torch.cat(out, -1)
    assert y.shape == (500, enc.n_output_dims) and torch.equal(y, ref)            # bit-identical forward
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    g1, = torch.autograd.grad(y, x, go)
    g2, = torch.autograd.grad(ref, x, go)
    assert float((g1 - g2).abs().max()) <= 1e-5 * float(g2.abs().max())
    if step is not None:                                                           # the mask follows update_step
        enc.update_step(0, 100)
        assert float((enc(x) - y).abs().max()) > 0


def test_fold_once_shares_the_folded_weights_inside_a_step_only():
    """VanillaMLP.effective_weights(): inside a `fold_once()` scope (one training step) the weight-norm fold is computed
    once and its gradient accumulates over the uses; outside a scope every call folds again (a graph kept past its
    backward could not be reused)."""
    from rise_sdf_b200.network_utils import VanillaMLP, fold_once
    torch.manual_seed(0)
    m = VanillaMLP(35, 48, {"n_neurons": 128, "n_hidden_layers": 2, "sphere_init": True, "weight_norm": True,
                            "output_activation": "none"})
    a = m.effective_weights()
    b = m.effective_weights()
    assert a[0][0] is not b[0][0] and torch.equal(a[0][0], b[0][0])
    x = torch.randn(7, 35)

    def loss_of(ws):
        (W1, b1), (W2, b2), (W3, b3) = ws
        return (torch.relu(torch.relu(x @ W1.T + b1) @ W2.T + b2) @ W3.T + b3).sum()

    ref = [None]
    for shared in (False, True):
        m.zero_grad(set_to_none=True)
        if shared:
            with fold_once():
                w1, w2 = m.effective_weights(), m.effective_weights()
                assert w1[0][0] is w2[0][0]
                (loss_of(w1) + loss_of(w2)).backward()
            with fold_once():                                   # a new step: folded afresh (the old graph is gone)
                w3 = m.effective_weights()
                assert w3[0][0] is not w1[0][0]
                loss_of(w3).backward()
        else:
            (loss_of(m.effective_weights()) + loss_of(m.effective_weights())).backward()
            loss_of(m.effective_weights()).backward()
        grads = [p.grad.clone() for p in m.parameters()]
        if ref[0] is None:
            ref[0] = grads
        else:
            for g, r in zip(grads, ref[0]):
                assert torch.allclose(g, r, rtol=1e-5, atol=1e-7)
    assert m.effective_weights()[0][0] is not m.effective_weights()[0][0]       # no sharing once the scope is left
