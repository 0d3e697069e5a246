"""Optimizer and learning-rate schedule of the training step.

Reference: `systems/utils.py:309-320` (`parse_optimizer` -> `torch.optim.Adam` over named parameter
groups with their own lr) and `systems/utils.py:323-346` (`parse_scheduler` -> `SequentialLR(LinearLR
warm-up, ExponentialLR)`, configs/neus-blender.yaml:104-119, configs/split-mixed-occ-tensoir.yaml:167-182).

`FlatAdam` keeps the call surface of `torch.optim.Adam` (param groups, `step`, `zero_grad`, `state_dict`
with `step / exp_avg / exp_avg_sq` per parameter, works under `torch.optim.lr_scheduler.*`) but owns ONE
flat fp32 buffer each for parameters, gradients (the data-parallel exchange bucket, `train.FlatGradBucket`),
`exp_avg` and `exp_avg_sq`.  Every `nn.Parameter` is re-pointed at its slice, so the whole update is a single
`rsdf_adam_step` launch (csrc/optim.cu; 28 B/parameter of HBM traffic) instead of a multi-tensor sweep, and
the gradient bucket can be cleared in the same pass.  CUDA only: there is no CPU path.
"""
import ctypes
import math

import torch

from . import _lib as L


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, bucket=None,
                 zero_grad_in_step=False):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) > L.ADAM_MAX_GROUPS:
            raise ValueError(f"FlatAdam supports at most {L.ADAM_MAX_GROUPS} parameter groups")
        self.zero_grad_in_step = bool(zero_grad_in_step)
        # `param_groups` keeps EVERY parameter it was given, like torch.optim.Adam (the reference's groups come from
        # `module.parameters()` and include the empty tcnn SphericalHarmonics `params` and anything frozen), so group
        # lengths and the state-dict index mapping are the reference's and a Lightning `optimizer_states` entry loads.
        # Only the flat buffers leave out what can never receive a gradient.
        self._in_flat = lambda p: p.requires_grad and p.numel() > 0
        self._flat_params = [p for g in self.param_groups for p in g["params"] if self._in_flat(p)]
        if not self._flat_params:
            raise ValueError("FlatAdam got no trainable parameters")
        L.require_cuda(*self._flat_params)
        for p in self._flat_params:
            if p.dtype != torch.float32:
                raise TypeError("FlatAdam keeps fp32 master parameters only")
        if bucket is None:
            from .train import FlatGradBucket
            bucket = FlatGradBucket(self._flat_params)
        if [id(p) for p in bucket.params] != [id(p) for p in self._flat_params]:
            raise ValueError("gradient bucket and optimizer must list the same parameters in the same order")
        self.bucket = bucket
        dev = bucket.flat.device
        self.n = n = bucket.flat.numel()
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self._slices, self._group_end, self._group_t = [], [], []
        i = 0
        with torch.no_grad():
            for g in self.param_groups:
                for p in g["params"]:
                    if not self._in_flat(p):
                        continue
                    off, k = bucket.offsets[i], p.numel()
                    self.flat_p[off:off + k].copy_(p.data.reshape(-1))
                    p.data = self.flat_p[off:off + k].view(p.shape)
                    self._slices.append((off, k))
                    self.state[p] = {"step": torch.tensor(0.0),
                                     "exp_avg": self.flat_m[off:off + k].view(p.shape),
                                     "exp_avg_sq": self.flat_v[off:off + k].view(p.shape)}
                    i += 1
                # a group ends where the next one's first slice starts (padding belongs to the group before:
                # p = g = m = v = 0 there, which Adam leaves at 0)
                self._group_end.append(bucket.offsets[i] if i < len(bucket.offsets) else n)
                self._group_t.append(0)
        self._t = 0

    # ------------------------------------------------------------------ gradients
    def _grad_view(self, i):
        off, k = self._slices[i]
        return self.bucket.flat[off:off + k].view(self._flat_params[i].shape)

    def _attach_grads(self):
        """A gradient produced outside the bucket (after `zero_grad(set_to_none=True)` or an external
        assignment) is copied into its slice; every `.grad` ends up a view of the bucket again."""
        for i, p in enumerate(self._flat_params):
            off, _ = self._slices[i]
            want = self.bucket.flat.data_ptr() + 4 * off
            if p.grad is None or p.grad.data_ptr() != want:
                view = self._grad_view(i)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view

    def zero_grad(self, set_to_none=False):
        """Clears the flat bucket; the `.grad` views stay attached (set_to_none is ignored on purpose:
        dropping the views would break the single-buffer exchange)."""
        self.bucket.zero()
        self._attach_grads()

    # ------------------------------------------------------------------ update
    def _groups_struct(self, skip):
        G = L.AdamGroupsC()
        G.n_groups = len(self.param_groups)
        for k, g in enumerate(self.param_groups):
            b1, b2 = g["betas"]
            t = max(self._group_t[k], 1)
            G.end[k] = self._group_end[k]
            G.step_size[k] = float(g["lr"]) / (1.0 - b1 ** t)          # torch/optim/adam.py: lr / bias_correction1
            G.one_minus_beta1[k], G.beta2[k], G.one_minus_beta2[k], G.eps[k] = 1.0 - b1, b2, 1.0 - b2, g["eps"]
            G.bias2_sqrt[k] = math.sqrt(1.0 - b2 ** t)
            G.weight_decay[k] = g["weight_decay"]
            G.skip[k] = 1 if k in skip else 0
        return G

    @torch.no_grad()
    def step(self, closure=None, inactive_groups=()):
        """One Adam update of every group not listed in `inactive_groups`.

        torch.optim.Adam skips a parameter whose `.grad` is None -- no moment decay, no step count -- which is how the
        reference's `emitter.base` sits out the 10 000 stage-0 steps (the env light is not on the graph before
        `split_sum_kick_in_step`) and then starts with step = 1 and full bias correction.  Gradients here are always
        views of the flat bucket (zero, never None), so the caller names the groups that are off the graph; each group
        carries its own step count."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._attach_grads()
        skip = {int(k) for k in inactive_groups}
        self._t += 1
        for k in range(len(self.param_groups)):
            if k not in skip:
                self._group_t[k] += 1
        G = self._groups_struct(skip)
        L.call("rsdf_adam_step", L.ptr(self.flat_p), L.ptr(self.bucket.flat), L.ptr(self.flat_m), L.ptr(self.flat_v),
               self.n, ctypes.addressof(G), int(self.zero_grad_in_step), L.stream())
        # the kernel wrote through raw pointers: tell autograd / the packed-weight caches
        torch.autograd.graph.increment_version(self._flat_params)
        for k, g in enumerate(self.param_groups):
            if k in skip:
                continue
            for p in g["params"]:
                if self._in_flat(p):
                    self.state[p]["step"] += 1
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout.  A parameter that has not been stepped yet has no entry there (Adam creates
        state lazily, on the first gradient): same here, so a checkpoint written by either optimizer loads in both."""
        sd = super().state_dict()
        sd["state"] = {i: st for i, st in sd["state"].items() if float(st["step"]) > 0}
        return sd

    # ------------------------------------------------------------------ checkpoints
    def load_state_dict(self, state_dict):
        """Accepts a `torch.optim.Adam` state dict of the same parameter layout (the `optimizer_states`
        entry of a Lightning checkpoint): moments are copied into the flat buffers."""
        super().load_state_dict(state_dict)
        t = 0
        group_of = {id(p): k for k, g in enumerate(self.param_groups) for p in g["params"]}
        self._group_t = [0] * len(self.param_groups)
        with torch.no_grad():
            for i, p in enumerate(self._flat_params):
                off, k = self._slices[i]
                st = self.state.get(p, {})
                for key, flat in (("exp_avg", self.flat_m), ("exp_avg_sq", self.flat_v)):
                    view = flat[off:off + k].view(p.shape)
                    if key in st:
                        if st[key].data_ptr() != view.data_ptr():
                            view.copy_(st[key])
                    else:                               # never stepped when the checkpoint was written
                        view.zero_()
                    st[key] = view
                step = st.get("step", torch.tensor(0.0))
                st["step"] = torch.tensor(float(step))
                t = max(t, int(float(step)))
                gk = group_of[id(p)]
                self._group_t[gk] = max(self._group_t[gk], int(float(step)))
                self.state[p] = st
            for p in [p for g in self.param_groups for p in g["params"] if not self._in_flat(p)]:
                self.state.pop(p, None)
        self._t = t


def exp_lr_decay_rate(factor, n):
    """OmegaConf resolver `calc_exp_lr_decay_rate` (utils/misc.py:7): gamma with gamma^n = factor."""
    return factor ** (1.0 / n)


def warmup_exponential_scheduler(optimizer, warmup_steps=500, max_steps=30000, start_factor=0.01, end_factor=1.0,
                                 final_factor=0.1):
    """`SequentialLR([LinearLR, ExponentialLR], milestones=[warmup_steps])` stepped once per optimizer step
    (configs/neus-blender.yaml:104-119, systems/utils.py:323-346)."""
    from torch.optim import lr_scheduler as S
    return S.SequentialLR(
        optimizer,
        [S.LinearLR(optimizer, start_factor=start_factor, end_factor=end_factor, total_iters=warmup_steps),
         S.ExponentialLR(optimizer, gamma=exp_lr_decay_rate(final_factor, max_steps - warmup_steps))],
        milestones=[warmup_steps])
