"""Generates tests/golden/ref_host_neus.npz by running the REFERENCE's own models/neus.py (+ geometry.py, texture.py,
network_utils.py; imported unmodified from /root/reference) on the CPU over the oracle's stand-ins for its third-party
imports (oracle/ref_host.py, backend="oracle").  Run in the build container (the reference tree cannot travel):
    python tests/golden/make_ref_host_golden.py
tests/test_gpu_ref_golden.py then checks the CUDA product against these outputs on the GPU box.
The 50 MB hash table is not stored: both sides draw it from the same seeded CPU generator.  `fp32 tensor / python scalar`
follows the CUDA semantics (reciprocal multiply) the reference always runs under: see oracle/ref_host.py.  (A float64
run is no better a yardstick here: its sample positions differ from the fp32 ones by 1e-7, which this random
high-frequency field turns into 1e-4 in the colours.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import neus as oneus  # noqa: E402
from oracle import ref_host  # noqa: E402
from rise_sdf_b200 import synthetic as syn  # noqa: E402
from rise_sdf_b200.neus import neus_blender_config  # noqa: E402

N_RAYS, STEP, RATIO, TABLE_SEED = 192, 1.732 * 2 * 1.5 / 256, 0.37, 123


def table(n):
    return (torch.rand(n, generator=torch.Generator().manual_seed(TABLE_SEED)) * 2 - 1) * 0.005


def fields_meta_params():
    from oracle import fields
    return fields.HashGridMeta(base_resolution=16).n_params


def main():
    out_path = os.path.join(ROOT, "tests", "golden", "ref_host_neus.npz")
    cfg = ref_host.ref_config(neus_blender_config())
    tab = table(fields_meta_params())                       # fp32 draws, made OUTSIDE the float64 context
    w_hash = torch.randn(128, 32, generator=torch.Generator().manual_seed(5)) * 0.05
    batch = syn.training_rays(N_RAYS, seed=6)
    with ref_host.reference_modules("oracle", cuda_scalar_division=True) as models:
        torch.manual_seed(0)
        m = models.make("neus", cfg)
        m.geometry.contraction_type = sys.modules["models.geometry"].ContractionType.AABB     # models/neus.py:57
        with torch.no_grad():
            p = m.geometry.encoding.encoding.params
            assert p.numel() == tab.numel()
            p.copy_(tab)
            m.geometry.network.layers[0].weight_v[:, 3:] = w_hash
        m.train()
        m.randomized = False
        m.cos_anneal_ratio = RATIO
        m.occupancy_grid.binaries = syn.analytic_grid("ball")[None]
        m.render_step_size = STEP
        rays, rgb, fg, bg = batch
        m.background_color = bg
        out = m(rays)
        loss, parts = oneus.loss({**out, "rays_valid": out["rays_valid_full"]}, rgb, fg)
        loss.backward()
        rec = {"loss": np.float64(float(loss))}
        for k in ("comp_rgb", "comp_normal", "opacity", "depth", "comp_rgb_full", "sdf_samples", "sdf_grad_samples", "weights",
                  "ray_indices"):
            rec["out." + k] = out[k].detach().numpy() if k == "ray_indices" else out[k].detach().float().numpy()
        for k, v in parts.items():
            rec["loss." + k] = np.float64(float(v))
        for k, v in m.state_dict().items():
            if "encoding.encoding.params" in k or "occupancy_grid" in k or v.numel() == 0:
                continue
            rec["state." + k] = v.detach().float().numpy() if v.is_floating_point() else v.detach().numpy()
        for k, v in m.named_parameters():
            if v.numel() == 0:
                continue
            g = v.grad.detach()
            if "encoding.encoding.params" in k:
                rec["grad_norm." + k] = np.float64(float(g.double().norm()))
                rec["grad_sub." + k] = g[::997].float().numpy()
            else:
                rec["grad." + k] = g.float().numpy()
    np.savez_compressed(out_path, **rec)
    print("wrote", out_path, os.path.getsize(out_path), "bytes;", int(out["num_samples"]), "samples")


if __name__ == "__main__":
    main()
