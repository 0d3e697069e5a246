"""Shared test plumbing: build a product NeuSModel and the matching oracle parameter set."""
import numpy as np
import torch

from oracle import fields as ofields
from oracle import neus as oneus


def oracle_params_from_model(model):
    """Copy every learnable tensor of a rise_sdf_b200.neus.NeuSModel to CPU oracle containers."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    enc = model.geometry.encoding.encoding
    cfg = enc.encoding_config
    meta = ofields.HashGridMeta(cfg["n_levels"], cfg["n_features_per_level"], cfg["log2_hashmap_size"],
                                cfg["base_resolution"], cfg["per_level_scale"])
    geo = ofields.mlp_layers_from_state(sd, "geometry.network.")
    tex = ofields.mlp_layers_from_state(sd, "texture.network.")
    return oneus.NeusParams(sd["geometry.encoding.encoding.params"], geo, tex, sd["variance.variance"], meta,
                            radius=model.config.radius, sh_degree=model.texture.encoding.encoding.degree)


def rel_err(a, b, floor=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def split_oracle_params(model):
    """Copy a rise_sdf_b200.split_mixed_occ.SplitMixedOCCModel to the CPU oracle container."""
    from oracle import split as osplit
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    enc = model.geometry.encoding.encoding          # ProgressiveBandHashGrid
    cfg = enc.encoding.encoding_config
    meta = ofields.HashGridMeta(cfg["n_levels"], cfg["n_features_per_level"], cfg["log2_hashmap_size"],
                                cfg["base_resolution"], cfg["per_level_scale"])
    geo = ofields.mlp_layers_from_state(sd, "geometry.network.")
    nets = {n: ofields.mlp_layers_from_state(sd, f"texture.{n}_network.")
            for n in ("albedo", "roughness", "metallic", "env", "secondary")}
    return osplit.SplitParams(sd["geometry.encoding.encoding.encoding.params"], meta, geo, nets, sd["variance.variance"],
                              sd["texture.FG_LUT"], sd["emitter.base"], radius=model.config.radius,
                              level_mask=enc.mask.detach().cpu().clone(), fd_eps=model.geometry._finite_difference_eps)
