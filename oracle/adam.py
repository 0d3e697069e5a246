"""CPU restatement of the flat Adam kernel's arithmetic (rise_sdf_b200/csrc/optim.cu `adam_one`) -- TEST
INFRASTRUCTURE ONLY, never imported by the product.

Follows torch/optim/adam.py `_single_tensor_adam` (amsgrad off, maximize off, weight_decay = L2), the update the
reference runs through `parse_optimizer` (systems/utils.py:309-320): every product rounded to fp32 on its own, in the
order lerp_ / mul_.addcmul_ / sqrt / div / add / addcdiv_, with `1 - beta` rounded from the host's double.
Pinned: tests/test_optim.py::test_adam_restatement_matches_torch_cpu runs it against torch.optim.Adam itself.
"""
import math

import numpy as np

f32 = np.float32


def adam_step(p, g, m, v, t, lr, beta1, beta2, eps, weight_decay=0.0):
    """One update of fp32 arrays p, g, m, v (step count t >= 1); returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, dtype=f32) for a in (p, g, m, v))
    step_size = f32(lr / (1.0 - beta1 ** t))
    bias2_sqrt = f32(math.sqrt(1.0 - beta2 ** t))
    if weight_decay:
        g = (f32(weight_decay) * p + g).astype(f32)      # the kernel fuses this one multiply-add (torch: add(alpha=))
    m = (m + ((g - m).astype(f32) * f32(1.0 - beta1)).astype(f32)).astype(f32)
    v = ((v * f32(beta2)).astype(f32) + ((f32(1.0 - beta2) * g).astype(f32) * g).astype(f32)).astype(f32)
    denom = ((np.sqrt(v).astype(f32) / bias2_sqrt).astype(f32) + f32(eps)).astype(f32)
    p = (p - (step_size * (m / denom).astype(f32)).astype(f32)).astype(f32)
    return p, m, v
