"""The training step around the render hot path: loss block of systems/neus.py:98-135
(rgb MSE on valid rays, eikonal, mask BCE, sparsity; weights from configs/neus-blender.yaml:83-91),
Adam with the per-group learning rates of configs/neus-blender.yaml:92-104, and -- for N>1
GPUs -- the DDP-equivalent gradient exchange of launch.py:84-97: ONE flat fp32 bucket
(hash table || MLPs || variance) all-reduced with NCCL and averaged.
"""
import torch
import torch.distributed as dist

from .network_utils import fold_once
from .optim import FlatAdam


def masked_mean(err, valid):
    """`err[valid].mean()` for err [R, C] and a row mask valid [R] (systems/neus.py:103, systems/split_occ.py:163-177
    index with the boolean mask, which costs a host sync in the middle of the step): the masked sum over the row
    count.  No valid row -> 0/0 = nan, like the mean of an empty selection."""
    v = valid.to(err.dtype)
    return (err * v[:, None]).sum() / (v.sum() * err.shape[1])


def masked_mse(a, b, valid):
    return masked_mean((a - b) ** 2, valid)


def masked_l1(a, b, valid):
    return masked_mean((a - b).abs(), valid)


def binary_cross_entropy(inp, target):
    """systems/criterions.py:155-159."""
    return -(target * torch.log(inp) + (1 - target) * torch.log(1 - inp)).mean()


class _SdfRegularisers(torch.autograd.Function):
    """(sdf_grad [S,3], sdf [S]) -> (mean (|sdf_grad| - 1)^2, mean exp(-scale |sdf|)): one reduction launch
    forward, one elementwise launch backward (instead of ~8 + ~10 torch kernels)."""

    @staticmethod
    def forward(ctx, sdf_grad, sdf, scale):
        from . import _lib as L
        sdf_grad, sdf = sdf_grad.contiguous().float(), sdf.contiguous().float()
        n = sdf.shape[0]
        out = torch.empty(2, device=sdf.device, dtype=torch.float32)
        partials = torch.empty(2 * L.SDF_REG_BLOCKS, device=sdf.device, dtype=torch.float32)
        L.call("rsdf_sdf_reg_fwd", L.ptr(sdf_grad), L.ptr(sdf), n, float(scale), L.ptr(out), L.ptr(partials), L.stream())
        ctx.save_for_backward(sdf_grad, sdf)
        ctx.scale = float(scale)
        out = out / max(n, 1)
        return out[0], out[1]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ge, gs):
        from . import _lib as L
        sdf_grad, sdf = ctx.saved_tensors
        n = sdf.shape[0]
        cot = torch.stack([ge, gs]).float() / max(n, 1)
        g1, g2 = torch.empty_like(sdf_grad), torch.empty_like(sdf)
        L.call("rsdf_sdf_reg_bwd", L.ptr(sdf_grad), L.ptr(sdf), n, ctx.scale, L.ptr(cot.contiguous()), L.ptr(g1), L.ptr(g2),
               L.stream())
        return g1, g2, None


def sdf_regularisers(sdf_grad, sdf, sparsity_scale=1.0):
    """-> (eikonal, sparsity) loss terms (systems/neus.py:117-131)."""
    if sdf.is_cuda and sdf.shape[0] > 0:
        return _SdfRegularisers.apply(sdf_grad, sdf, sparsity_scale)
    eik = ((torch.linalg.norm(sdf_grad, ord=2, dim=-1) - 1.0) ** 2).mean()
    return eik, torch.exp(-sparsity_scale * sdf.abs()).mean()


def ray_terms(comp_rgb_full, opacity, rgb, fg_mask):
    """-> (masked rgb MSE, mask BCE) of systems/neus.py:103,123-125: one fused pass over the rays on CUDA."""
    if comp_rgb_full.is_cuda and comp_rgb_full.shape[0] > 0:
        from .glue import ray_loss_terms
        return ray_loss_terms(comp_rgb_full, opacity, rgb, fg_mask)
    op = torch.clamp(opacity.squeeze(-1), 1e-3, 1 - 1e-3)
    return masked_mse(comp_rgb_full, rgb, opacity.reshape(-1) > 0), binary_cross_entropy(op, fg_mask.float())


def neus_loss(out, rgb, fg_mask, lambda_rgb_mse=10.0, lambda_mask=0.1, lambda_eikonal=0.1,
              lambda_sparsity=0.01, sparsity_scale=1.0):
    loss_rgb, loss_mask = ray_terms(out["comp_rgb_full"], out["opacity"], rgb, fg_mask)
    loss_eik, loss_sparse = sdf_regularisers(out["sdf_grad_samples"], out["sdf_samples"], sparsity_scale)
    loss = (loss_rgb * lambda_rgb_mse + loss_eik * lambda_eikonal + loss_mask * lambda_mask
            + loss_sparse * lambda_sparsity)
    return loss, {"rgb_mse": loss_rgb, "eikonal": loss_eik, "mask": loss_mask, "sparsity": loss_sparse}


SPLIT_LAMBDAS = dict(lambda_rgb_mse=10.0, lambda_rgb_l1=0.0, lambda_rgb_phys_mse=10.0, lambda_rgb_phys_l1=0.0,
                     lambda_mask=0.1, lambda_eikonal=0.05, lambda_sparsity=0.01, lambda_curvature=1.0,
                     lambda_opaque=0.0, lambda_normal_orientation=0.05, lambda_emitter_distillation=0.0,
                     sparsity_scale=1.0)      # configs/split-mixed-occ-tensoir.yaml:139-152


def split_loss(model, out, rgb, fg_mask, has_mask=True, **overrides):
    """Loss block of systems/split_occ.py:163-225 (distortion terms excluded: their lambdas are 0 in the
    config and the helper library is not part of the path)."""
    lam = dict(SPLIT_LAMBDAS, **overrides)
    valid = out["rays_valid_full"][..., 0]
    parts = {}
    parts["rgb_mse"], parts["mask"] = ray_terms(out["comp_rgb_full"], out["opacity"], rgb, fg_mask)
    loss = parts["rgb_mse"] * lam["lambda_rgb_mse"]
    if lam["lambda_rgb_l1"]:
        loss = loss + masked_l1(out["comp_rgb_full"], rgb, valid) * lam["lambda_rgb_l1"]
    if model.stage != 0:
        parts["rgb_phys_mse"] = masked_mse(out["comp_rgb_phys_full"], rgb, valid)
        loss = loss + parts["rgb_phys_mse"] * lam["lambda_rgb_phys_mse"]
        if lam["lambda_rgb_phys_l1"]:
            loss = loss + masked_l1(out["comp_rgb_phys_full"], rgb, valid) * lam["lambda_rgb_phys_l1"]
    parts["eikonal"], parts["sparsity"] = sdf_regularisers(out["sdf_grad_samples"], out["sdf_samples"],
                                                            lam["sparsity_scale"])
    loss = loss + parts["eikonal"] * lam["lambda_eikonal"]
    loss = loss + parts["mask"] * (lam["lambda_mask"] if has_mask else 0.0)
    if lam["lambda_opaque"]:
        opacity = torch.clamp(out["opacity"].squeeze(-1), 1e-3, 1 - 1e-3)
        loss = loss + binary_cross_entropy(opacity, opacity) * lam["lambda_opaque"]
    loss = loss + parts["sparsity"] * lam["lambda_sparsity"]
    if lam["lambda_curvature"] > 0:
        assert "sdf_laplace_samples" in out, "Need geometry.grad_type='finite_difference' to get SDF Laplace samples"
        parts["curvature"] = out["sdf_laplace_samples"].abs().mean()
        loss = loss + parts["curvature"] * lam["lambda_curvature"]
    if lam["lambda_emitter_distillation"] > 0 and model.stage != 0:
        loss = loss + masked_mse(out["comp_spec_rgb_full"], out["comp_spec_rgb_phys_full"], valid) \
            * lam["lambda_emitter_distillation"]
    for name, value in model.geometry.regularizations(out).items():      # normal_orientation
        parts[name] = value
        loss = loss + value * lam[f"lambda_{name}"]
    return loss, parts


class FlatGradBucket:
    """All parameters' gradients as views into one contiguous fp32 buffer, so the data-parallel
    exchange is a single NCCL all-reduce instead of DDP's 25 MB buckets.  Every
    parameter's slice starts on a 256-byte boundary (`ALIGN` elements; the padding stays zero), which
    keeps the vectorised kernels' alignment assumptions when `optim.FlatAdam` lays the parameters out
    the same way."""

    ALIGN = 64

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad and p.numel() > 0]
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())


class _Trainer:
    """The per-batch sequence of the reference's Lightning loop, without Lightning:
        on_train_batch_start  -> update_module_step(model, epoch, global_step)   (systems/base.py:100-103): occupancy
                                 refresh every 16 steps, cos-anneal ratio, progressive hash-level mask + FD eps, stage
        training_step         -> render, losses (systems/neus.py:98-135, systems/split_occ.py:150-237)
        backward, (DDP all-reduce), optimizer.step(), scheduler.step() [interval: step]
    `step(..., update=False)` leaves the schedule to the caller (parity tests pin the state by hand)."""

    max_steps, warmup_steps = 30000, 500

    def _setup(self, groups, lr, betas, eps, scheduler):
        from .optim import warmup_exponential_scheduler
        self.opt = FlatAdam(groups, lr=lr, betas=betas, eps=eps)
        self.bucket = self.opt.bucket
        self.sched = warmup_exponential_scheduler(self.opt, self.warmup_steps, self.max_steps) if scheduler else None
        self.global_step = 0

    def _inactive_groups(self):
        return ()

    def _finish(self, loss, optimize):
        loss.backward()
        self.bucket.all_reduce_mean()
        if optimize:
            self.opt.step(inactive_groups=self._inactive_groups())
            if self.sched is not None:
                self.sched.step()
        self.global_step += 1


class NeusTrainer(_Trainer):
    """configs/neus-blender.yaml:92-119: Adam(betas 0.9/0.99, eps 1e-15), lr 0.01 (variance 0.001), 500-step linear
    warm-up then exponential decay to 0.1x at trainer.max_steps = 30000."""

    def __init__(self, model, lr=0.01, lr_variance=0.001, scheduler=True):
        self.model = model
        groups = [
            {"params": list(model.geometry.parameters()), "lr": lr},
            {"params": list(model.texture.parameters()), "lr": lr},
            {"params": list(model.variance.parameters()), "lr": lr_variance},
        ]
        self._setup(groups, lr, (0.9, 0.99), 1e-15, scheduler)

    def step(self, rays, rgb, fg_mask, background, optimize=True, update=True):
        m = self.model
        if update:
            m.update_step(0, self.global_step)
        m.background_color = background
        self.bucket.zero()
        with fold_once():
            out = m(rays)
            loss, parts = neus_loss(out, rgb, fg_mask)
            self._finish(loss, optimize)
        return loss, out


class SplitTrainer(_Trainer):
    """One split-mixed-occ training step (systems/split_occ.py:150-237): rebuild the env-light mip pyramid
    (`emitter.build_mips()`, the `base` cube map is learnable), render, loss, backward, gradient exchange,
    Adam with the per-group learning rates of configs/split-mixed-occ-tensoir.yaml:153-166 and its schedule
    (:167-182, trainer.max_steps = 80000)."""

    max_steps = 80000
    EMITTER_GROUP = 3

    def __init__(self, model, lr=0.005, lr_variance=0.001, lr_emitter=0.01, scheduler=True):
        self.model = model
        groups = [
            {"params": list(model.geometry.parameters()), "lr": lr},
            {"params": list(model.texture.parameters()), "lr": lr},
            {"params": list(model.variance.parameters()), "lr": lr_variance},
            {"params": list(model.emitter.parameters()), "lr": lr_emitter},
        ]
        self._setup(groups, lr, (0.9, 0.999), 1e-12, scheduler)

    def _inactive_groups(self):
        # stage 0 (models/texture.py:323-324 returns before any emitter lookup): `emitter.base.grad` stays None in the
        # reference and torch.optim.Adam skips it -- no moment decay, its step count starts at the stage switch
        return (self.EMITTER_GROUP,) if self.model.stage == 0 else ()

    def step(self, rays, rgb, fg_mask, background, optimize=True, update=True):
        m = self.model
        if update:
            m.update_step(0, self.global_step)
        m.background_color = background
        self.bucket.zero()
        with fold_once():
            m.emitter.build_mips()
            out = m(rays)
            loss, parts = split_loss(m, out, rgb, fg_mask)
            self._finish(loss, optimize)
        return loss, out
