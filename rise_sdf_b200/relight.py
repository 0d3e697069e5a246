"""Relighting render (BASELINE configs[3]; systems/split_occ.py:331-458): 800x800 frames of a
split-sum model under new environment maps, pixels sharded across GPUs with NO communication
(SURVEY.md §8e): the frame is cut into `tile`-ray tiles dealt round-robin to ranks (the object
is centred, so row blocks would be imbalanced); every rank holds the full model, occupancy grid,
prefiltered mip pyramids and LUT, renders its tiles and keeps the result sharded (an optional
all_gather assembles the frame)."""
import torch

from . import synthetic as syn
from .light import blender_latlong_to_cubemap


class EnvSet:
    """Prefiltered pyramids for a list of lat-long HDR maps (built once per map: `base` is frozen
    when relighting, unlike training where build_mips runs every step)."""

    def __init__(self, model, latlongs):
        self.model = model
        self.maps = []
        em = model.emitter
        with torch.no_grad():
            for img in latlongs:
                em.base.data = blender_latlong_to_cubemap(img.to(em.base.device).float(), [512, 512])
                em.build_mips()
                self.maps.append(([t.detach().clone() for t in em.specular], em.diffuse.detach().clone()))

    def use(self, i):
        self.model.emitter.specular, self.model.emitter.diffuse = self.maps[i]


def my_tiles(n_rays, tile, rank, world):
    return [(s, min(s + tile, n_rays)) for k, s in enumerate(range(0, n_rays, tile)) if k % world == rank]


def balanced_tile(n_rays, world, max_tile=32768, per_rank=4):
    """Tile size such that every rank gets the same NUMBER of tiles (>= per_rank of them when the frame is
    large enough): 640 000 rays -> 32 000 for 1/2/4 ranks (20 tiles), 20 032 for 8 ranks (32 tiles, 4 each)."""
    t = min(max_tile, -(-n_rays // (world * per_rank)))
    t = -(-t // 64) * 64
    n = -(-n_rays // t)
    n = -(-n // world) * world                # round the tile count up to a multiple of the ranks
    return -(-(-(-n_rays // n)) // 64) * 64


@torch.no_grad()
def render_frame_shard(model, rays, envs, rank=0, world=1, tile=None, keys=("comp_rgb_phys_full",),
                       share_across_envs=True):
    """Render this rank's tiles of one frame under every env map.  rays: [H*W, 6] on the device.
    Returns {env_index: {key: [n_my_rays, C]}} plus the tile list.

    share_across_envs: a tile is rendered under all maps back to back with a tile cache installed on the model
    (`SplitMixedOCCModel._memo`): sampling, field evaluations, material networks and the secondary bounce run once
    per tile, only the emitter lookups, compositing and the third-bounce shading run per map.  The frames are
    bit-identical to rendering every map from scratch (share_across_envs=False, the reference's loop order)."""
    if tile is None:
        tile = balanced_tile(rays.shape[0], world)
    tiles = my_tiles(rays.shape[0], tile, rank, world)
    n_env = len(envs.maps)
    parts = {e: {k: [] for k in keys} for e in range(n_env)}
    try:
        if share_across_envs:
            for a, b in tiles:
                model._tile_cache = {}
                for e in range(n_env):
                    envs.use(e)
                    o = model.forward_(rays[a:b], relighting=True)
                    for k in keys:
                        parts[e][k].append(o[k])
        else:
            for e in range(n_env):
                envs.use(e)
                for a, b in tiles:
                    o = model.forward_(rays[a:b], relighting=True)
                    for k in keys:
                        parts[e][k].append(o[k])
    finally:
        model._tile_cache = None
    out = {e: {k: torch.cat(v) if v else torch.zeros(0, 3, device=rays.device) for k, v in parts[e].items()}
           for e in range(n_env)}
    return out, tiles


def synthetic_envs():
    return [syn.env_latlong("bridge", seed=1), syn.env_latlong("city", seed=2)]
