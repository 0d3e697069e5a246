"""Generates tests/golden/ref_host_neus.npz by running the REFERENCE's own models/neus.py (+ geometry.py, texture.py,
network_utils.py; imported unmodified from /root/reference) on the CPU over the oracle's stand-ins for its third-party
imports (oracle/ref_host.py, backend="oracle").  Run in the build container (the reference tree cannot travel):
    python tests/golden/make_ref_host_golden.py
tests/test_gpu_ref_golden.py then checks the CUDA product against these outputs on the GPU box.
The 50 MB hash table is not stored: both sides draw it from the same seeded CPU generator."""
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
    return (torch.rand(n, generator=torch.Generator().manual_seed(TABLE_SEED)) * 2 - 1) * 0.05


def main():
    out_path = os.path.join(ROOT, "tests", "golden", "ref_host_neus.npz")
    cfg = ref_host.ref_config(neus_blender_config())
    with ref_host.reference_modules("oracle") as models:
        torch.manual_seed(0)
        m = models.make("neus", cfg)
        m.geometry.contraction_type = sys.modules["models.geometry"].ContractionType.AABB     # models/neus.py:57
        with torch.no_grad():
            p = m.geometry.encoding.encoding.params
            p.copy_(table(p.numel()))
            w = m.geometry.network.layers[0].weight_v
            w[:, 3:] = torch.randn(w[:, 3:].shape, generator=torch.Generator().manual_seed(5)) * 0.05
        m.train()
        m.randomized = False
        m.cos_anneal_ratio = RATIO
        m.occupancy_grid.binaries = syn.analytic_grid("ball")[None]
        m.render_step_size = STEP
        rays, rgb, fg, bg = syn.training_rays(N_RAYS, seed=6)
        m.background_color = bg
        out = m(rays)
        loss, parts = oneus.loss({**out, "rays_valid": out["rays_valid_full"]}, rgb, fg)
        loss.backward()
        rec = {"loss": np.float64(float(loss))}
        for k in ("comp_rgb", "comp_normal", "opacity", "depth", "comp_rgb_full", "sdf_samples", "sdf_grad_samples", "weights",
                  "ray_indices"):
            rec["out." + k] = out[k].detach().numpy()
        for k, v in parts.items():
            rec["loss." + k] = np.float64(float(v))
        for k, v in m.state_dict().items():
            if "encoding.encoding.params" in k or "occupancy_grid" in k or v.numel() == 0:
                continue
            rec["state." + k] = v.detach().numpy()
        for k, v in m.named_parameters():
            if v.numel() == 0:
                continue
            g = v.grad.detach()
            if "encoding.encoding.params" in k:
                rec["grad_norm." + k] = np.float64(float(g.double().norm()))
                rec["grad_sub." + k] = g[::997].numpy()
            else:
                rec["grad." + k] = g.numpy()
    np.savez_compressed(out_path, **rec)
    print("wrote", out_path, os.path.getsize(out_path), "bytes;", int(out["num_samples"]), "samples")


if __name__ == "__main__":
    main()
