"""Build recipe for oracle/_ref: the reference's OWN in-tree CUDA sources, compiled
unmodified from where they lie under /root/reference (never copied into this repo).

TEST INFRASTRUCTURE ONLY.  Outputs go to oracle/_ref/ (git-ignored, shipped to the GPU
box by gpurun).  Two pybind modules are produced:

  oracle/_ref/nerfacc_cuda.so       <- /root/reference/lib/nerfacc/cuda/csrc/*.cu
        (ray_marching, ray_aabb_intersect, grid_query, weight/transmittance_from_alpha_*)
  oracle/_ref/renderutils_plugin.so <- /root/reference/lib/renderutils/c_src/*.{cu,cpp}
        (diffuse_cubemap_fwd/bwd, specular_bounds, specular_cubemap_fwd/bwd)

Flags mirror the reference's JIT recipe: lib/nerfacc/cuda/_backend.py:43-44 ("-O3") and
lib/renderutils/ops.py:40-66 ("-DNVDR_TORCH", -lcuda -lnvrtc), with the arch forced to
sm_100a.  Run:  python oracle/build_ref.py [nerfacc|renderutils|all]
"""
import glob
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def _load(name, sources, extra_cflags, extra_cuda_cflags, extra_ldflags):
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    load(
        name=name,
        sources=sources,
        extra_cflags=extra_cflags,
        extra_cuda_cflags=extra_cuda_cflags,
        extra_ldflags=extra_ldflags,
        build_directory=bdir,
        is_python_module=False,
        verbose=True,
    )
    shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
    print("built", os.path.join(OUT, name + ".so"))


def build_nerfacc():
    src = sorted(glob.glob(os.path.join(REF, "lib/nerfacc/cuda/csrc/*.cu")))
    _load("nerfacc_cuda", src, ["-O3"], ["-O3"], [])


def build_renderutils():
    d = os.path.join(REF, "lib/renderutils/c_src")
    src = sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cpp")))
    stubs = "/usr/local/cuda/lib64/stubs"
    _load(
        "renderutils_plugin",
        src,
        ["-DNVDR_TORCH"],
        ["-DNVDR_TORCH"],
        ["-L" + stubs, "-lcuda", "-lnvrtc"],
    )


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("no /root/reference here; oracle/_ref must be prebuilt")
        sys.exit(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("nerfacc", "all"):
        build_nerfacc()
    if what in ("renderutils", "all"):
        build_renderutils()
