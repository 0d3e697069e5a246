"""Drive the reference's OWN host Python -- models/{neus,geometry,texture,network_utils,split_mixed_occ,volrend}.py,
lib/pbr/light.py -- UNMODIFIED, from where it lies, over stand-ins for its third-party imports.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and by tests/golden/make_ref_host_golden.py.

Two backends behind the names the reference imports:

  backend="product"  nerfacc / lib.nerfacc / tinycudann / nvdiffrast.torch / lib.renderutils -> the rise_sdf_b200 shims
                     (CUDA only).  Proves the drop-in claim: the reference's models run on librsdf_b200.so with no
                     source change, and agree with the repo's host mirrors.
  backend="oracle"   the same names -> CPU restatements from oracle/{fields,march,textures}.py.  Runs in the build
                     container (no GPU): pins the hand-written oracle/neus.py and oracle/split.py to the reference's
                     Python and produces tests/golden/ref_host_*.npz for the GPU box (where /root/reference is absent).

Also stubbed (host-only packages that are not installed: requirements.txt:1-2): pytorch_lightning.utilities.rank_zero,
omegaconf.OmegaConf (resolvers + to_container over the plain `Cfg` below), imageio, pyexr.  Nothing is copied: the
reference root is put on sys.path.
"""
import contextlib
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CANDIDATES = (os.environ.get("RSDF_REFERENCE_ROOT"), "/root/reference")
_REF_MODULES = ("models", "systems", "utils", "lib", "datasets")


def reference_root():
    for p in REF_CANDIDATES:
        if p and os.path.exists(os.path.join(p, "models", "neus.py")):
            return p
    return None


# ------------------------------------------------------------------------------------------------------------------
# config stand-in for OmegaConf nodes: attribute access, .get, .copy, `in`
# ------------------------------------------------------------------------------------------------------------------
class Cfg(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Cfg(dict.copy(self))


def to_primitive(c):
    if isinstance(c, dict):
        return {k: to_primitive(v) for k, v in c.items()}
    if isinstance(c, (list, tuple)):
        return [to_primitive(v) for v in c]
    return c


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _host_stubs():
    class OmegaConf:
        register_new_resolver = staticmethod(lambda *a, **k: None)
        to_container = staticmethod(lambda cfg, resolve=True: to_primitive(cfg))
        create = staticmethod(lambda d=None: Cfg(d or {}))

    noop = lambda *a, **k: None
    rz = _module("pytorch_lightning.utilities.rank_zero", rank_zero_info=noop, rank_zero_debug=noop, rank_zero_warn=noop,
                 rank_zero_only=lambda f: f)
    ut = _module("pytorch_lightning.utilities", rank_zero=rz)
    pl = _module("pytorch_lightning", utilities=ut)
    return {"omegaconf": _module("omegaconf", OmegaConf=OmegaConf), "pytorch_lightning": pl,
            "pytorch_lightning.utilities": ut, "pytorch_lightning.utilities.rank_zero": rz,
            "imageio": _module("imageio"), "pyexr": _module("pyexr")}


# ------------------------------------------------------------------------------------------------------------------
# backend "oracle": CPU restatements under the third-party names
# ------------------------------------------------------------------------------------------------------------------
def _oracle_backend():
    from enum import Enum

    from torch import nn

    from . import fields, march
    from . import neus as oneus
    from . import textures as ot

    class Encoding(nn.Module):
        """tinycudann.Encoding over oracle.fields (HashGrid, SphericalHarmonics)."""

        def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
            super().__init__()
            cfg = dict(encoding_config)
            self.n_input_dims, self.encoding_config = n_input_dims, cfg
            if cfg["otype"] in ("HashGrid", "Grid"):
                self.kind = "hashgrid"
                self.meta = fields.HashGridMeta(cfg.get("n_levels", 16), cfg.get("n_features_per_level", 2),
                                                cfg.get("log2_hashmap_size", 19), cfg.get("base_resolution", 16),
                                                cfg.get("per_level_scale", 2.0))
                self.n_output_dims = self.meta.n_output_dims
                g = torch.Generator().manual_seed(seed)
                self.params = nn.Parameter((torch.rand(self.meta.n_params, generator=g) * 2 - 1) * 1e-4)
            elif cfg["otype"] == "SphericalHarmonics":
                self.kind, self.degree = "sh", int(cfg["degree"])
                self.n_output_dims = self.degree ** 2
                self.params = nn.Parameter(torch.zeros(0))
            else:
                raise NotImplementedError(cfg["otype"])

        def forward(self, x):
            if self.kind == "hashgrid":
                return fields.hash_encode(x, self.params, self.meta)
            return fields.sh_encode(x, self.degree)

    tcnn = _module("tinycudann", Encoding=Encoding, free_temporary_memory=lambda: None)

    class ContractionType(Enum):
        AABB = 0
        UN_BOUNDED_TANH = 1
        UN_BOUNDED_SPHERE = 2

    class OccGridEstimator(nn.Module):
        """nerfacc 0.5.3 OccGridEstimator over oracle.march (sampling == the in-tree 0.3.5 march against the
        estimator's box) and oracle.neus.grid_update."""

        def __init__(self, roi_aabb, resolution=128, levels=1):
            super().__init__()
            roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
            self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
            self.register_buffer("aabbs", roi[None].clone())
            self.register_buffer("occs", torch.zeros(int(resolution) ** 3))
            self.register_buffer("binaries", torch.zeros(1, resolution, resolution, resolution, dtype=torch.bool))
            self._res = int(resolution)
            self.jitter = None           # explicit U[0,1) draws for `stratified` (tests share them with the product)

        @torch.no_grad()
        def sampling(self, rays_o, rays_d, sigma_fn=None, alpha_fn=None, near_plane=0.0, far_plane=1e10, t_min=None,
                     t_max=None, render_step_size=1e-3, early_stop_eps=1e-4, alpha_thre=0.0, stratified=False,
                     cone_angle=0.0):
            assert sigma_fn is None and t_min is None and t_max is None
            roi = self.aabbs[0].numpy()
            fn = None
            if alpha_fn is not None:
                fn = lambda ts, te, ri: alpha_fn(torch.from_numpy(ts), torch.from_numpy(te),
                                                 torch.from_numpy(ri)).detach().reshape(-1).numpy()
            jit = None
            if stratified:
                jit = self.jitter if self.jitter is not None else torch.rand(rays_o.shape[0])
            ri, ts, te = march.ray_marching(rays_o.detach().numpy(), rays_d.detach().numpy(), scene_aabb=roi, grid_roi=roi,
                                            grid_binary=self.binaries[0].numpy(), alpha_fn=fn,
                                            early_stop_eps=early_stop_eps, alpha_thre=alpha_thre, near_plane=near_plane,
                                            far_plane=far_plane, render_step_size=render_step_size,
                                            jitter=None if jit is None else jit.numpy(), cone_angle=cone_angle)
            return torch.from_numpy(ri), torch.from_numpy(ts), torch.from_numpy(te)

        @torch.no_grad()
        def update_every_n_steps(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
            if step % n == 0 and self.training:
                occs, binary = oneus.grid_update(self.occs, step, occ_eval_fn, self.aabbs[0], self._res, occ_thre,
                                                 ema_decay, warmup_steps, jitter=self.update_jitter)
                self.occs.copy_(occs)
                self.binaries = binary[None]

        update_jitter = None

    def render_weight_from_alpha(alphas, *, packed_info=None, ray_indices=None, n_rays=None):
        shape = alphas.shape
        w, T = fields.render_weight_from_alpha(alphas.reshape(-1), ray_indices, n_rays)
        return w.view(shape), T.view(shape)

    def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
        return fields.accumulate_along_rays(weights.reshape(-1), values, ray_indices, n_rays)

    def _unsupported(*a, **k):
        raise NotImplementedError("learned-background branch: out of scope (SURVEY.md section 8b)")

    nerfacc = _module("nerfacc", OccGridEstimator=OccGridEstimator, render_weight_from_alpha=render_weight_from_alpha,
                      accumulate_along_rays=accumulate_along_rays, render_weight_from_density=_unsupported,
                      ray_aabb_intersect=_unsupported, ContractionType=ContractionType, OccupancyGrid=OccGridEstimator,
                      ray_marching=_unsupported)

    def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode="auto", boundary_mode="wrap"):
        """nvdiffrast.torch.texture over oracle.textures, the three modes of SURVEY.md section 8b."""
        if boundary_mode == "cube":
            levels = [tex[0]] + [m[0] for m in (mip or [])]
            d = uv.reshape(-1, 3)
            out = ot.cube_sample(levels, d, None if mip_level_bias is None else mip_level_bias.reshape(-1))
            return out.reshape(*uv.shape[:-1], tex.shape[-1])
        out = ot.tex2d(tex[0], uv.reshape(-1, 2), wrap=(boundary_mode == "wrap"))
        return out.reshape(*uv.shape[:-1], tex.shape[-1])

    drt = _module("nvdiffrast.torch", texture=texture)
    ru = _module("lib.renderutils", diffuse_cubemap=lambda c, use_python=False: ot.diffuse_cubemap(c),
                 specular_cubemap=lambda c, roughness, cutoff=0.99, use_python=False: ot.specular_cubemap(c, roughness, cutoff))
    return {"tinycudann": tcnn, "nerfacc": nerfacc, "nerfacc.volrend": nerfacc, "lib.nerfacc": nerfacc,
            "nvdiffrast": _module("nvdiffrast", torch=drt), "nvdiffrast.torch": drt, "lib.renderutils": ru}


def _product_backend():
    import rise_sdf_b200.nerfacc as rn
    import rise_sdf_b200.nvdiffrast as rdr
    import rise_sdf_b200.renderutils as rru
    import rise_sdf_b200.tinycudann as rt
    return {"tinycudann": rt, "nerfacc": rn, "nerfacc.volrend": rn, "lib.nerfacc": rn,
            "nvdiffrast": _module("nvdiffrast", torch=rdr), "nvdiffrast.torch": rdr, "lib.renderutils": rru}


_FACTORIES = ("zeros", "ones", "empty", "full", "rand", "randn", "linspace", "arange", "tensor", "as_tensor", "eye",
              "zeros_like", "ones_like", "rand_like", "empty_like", "full_like")


@contextlib.contextmanager
def _cuda_to_cpu(double=False, cuda_scalar_division=False):
    """backend="oracle": the reference allocates with device='cuda' / `.cuda()` / device=rank in a few places
    (lib/pbr/light.py:139, lib/pbr/utils/light_utils.py:99-133, models/texture.py:294, models/network_utils.py:56);
    on the CPU those requests are redirected.  Plain attribute patches on `torch` (not a TorchFunctionMode: custom
    autograd backwards, e.g. cubemap_mip's, run outside the mode stack).
    double=True additionally runs the reference in float64: its explicit `.float()` casts
    (models/network_utils.py:122, models/geometry.py:217 ...) and dtype=torch.float32 requests become float64.  That
    removes the 1/(2 eps) ~ 1400x amplification of fp32 rounding that finite-difference normals carry, so host LOGIC
    can be compared at 1e-7."""
    saved = {n: getattr(torch, n) for n in _FACTORIES}
    saved_cuda, saved_float, saved_dev = torch.Tensor.cuda, torch.Tensor.float, torch.cuda.device
    saved_default = torch.get_default_dtype()
    saved_div = torch.Tensor.__truediv__

    def wrap(fn):
        def inner(*args, **kwargs):
            d = kwargs.get("device")
            if d is not None and (isinstance(d, int) or "cuda" in str(d)):
                kwargs["device"] = "cpu"
            if double and kwargs.get("dtype") is torch.float32:
                kwargs["dtype"] = torch.float64
            return fn(*args, **kwargs)
        return inner

    try:
        for n, fn in saved.items():
            setattr(torch, n, wrap(fn))
        torch.Tensor.cuda = lambda self, *a, **k: self
        if double:
            torch.Tensor.float = lambda self, *a, **k: self.double()
            # dtype-less factories too (models/neus.py:41 `torch.ones([len(x), 1]) * inv_s`: a 0-dim float64 operand
            # does not promote an fp32 tensor)
            torch.set_default_dtype(torch.float64)
        torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
        if cuda_scalar_division:
            # The reference only ever runs on CUDA, where `fp32 tensor / python scalar` is a multiplication by the fp32
            # reciprocal (scripts/probe_div.py; oracle.fields.cuda_scalar_div).  `scale_anything` (models/utils.py:
            # 109-114) feeds the hash-grid cell lookup, where that last bit is visible in the render at 1e-4: goldens
            # for the GPU box are generated under the CUDA semantics.
            def cuda_div(self, other):
                if isinstance(other, (int, float)) and self.dtype == torch.float32 and other != 0:
                    return self * float(np.float32(1.0) / np.float32(other))
                return saved_div(self, other)
            torch.Tensor.__truediv__ = cuda_div
        yield
    finally:
        for n, fn in saved.items():
            setattr(torch, n, fn)
        torch.Tensor.cuda, torch.Tensor.float, torch.cuda.device = saved_cuda, saved_float, saved_dev
        torch.set_default_dtype(saved_default)
        torch.Tensor.__truediv__ = saved_div


@contextlib.contextmanager
def reference_modules(backend, double=False, cuda_scalar_division=False):
    """Context: sys.modules / sys.path arranged so that `import models` imports the reference's package.  Yields the
    `models` module.  Everything is undone on exit (the repo's own tests must not see the aliases)."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference tree (set RSDF_REFERENCE_ROOT)")
    assert backend in ("product", "oracle")
    stubs = dict(_host_stubs())
    stubs.update(_product_backend() if backend == "product" else _oracle_backend())
    saved = {k: sys.modules.get(k) for k in stubs}
    saved_ref = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_MODULES}
    for k in saved_ref:
        del sys.modules[k]
    sys.modules.update(stubs)
    sys.path.insert(0, root)
    # `from systems.utils import update_module_step` (models/*.py) would run systems/__init__.py, which pulls in the
    # Lightning training systems (pytorch_lightning, torch_efficient_distloss, torchmetrics ...: the caller of the path,
    # out of scope).  A bare package object with the real __path__ lets `systems.utils` itself import unmodified.
    pkg = types.ModuleType("systems")
    pkg.__path__ = [os.path.join(root, "systems")]
    sys.modules["systems"] = pkg
    assert backend == "oracle" or not double
    mode = _cuda_to_cpu(double, cuda_scalar_division) if backend == "oracle" else contextlib.nullcontext()
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="rsdf_ref_host_")
    try:
        os.chdir(tmp)                    # models/texture.py:285 reads load/bsdf/bsdf_256_256.bin relative to the cwd
        with mode:
            models = importlib.import_module("models")
            yield models
    finally:
        os.chdir(cwd)
        sys.path.remove(root)
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_MODULES]:
            del sys.modules[k]
        sys.modules.update(saved_ref)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def write_bsdf_lut(fg_lut):
    """models/texture.py:285 loads `load/bsdf/bsdf_256_256.bin` (a README.md:66-71 download): write the synthetic
    table under the (temporary) working directory in that raw float32 layout."""
    os.makedirs("load/bsdf", exist_ok=True)
    np.asarray(fg_lut.detach().cpu().float().numpy(), dtype=np.float32).reshape(256, 256, 2).tofile(
        "load/bsdf/bsdf_256_256.bin")


def ref_config(cfg):
    """A repo config dict (rise_sdf_b200.neus.neus_blender_config / split_mixed_occ_config) -> the node the reference's
    constructors expect: adds the keys its yaml carries that the hot path never reads (`isosurface: null`)."""
    c = Cfg(to_primitive(cfg))
    c["geometry"]["isosurface"] = None
    return c
