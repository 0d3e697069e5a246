"""rise_sdf_b200 -- B200-native (sm_100a) implementation of RISE-SDF's ray-marched neural-SDF
rendering hot path, behind the reference's own operator interface.

Layout:
  csrc/ + librsdf_b200.so   hand-written CUDA kernels behind a C ABI (include/rsdf_b200.h)
  _lib.py                   ctypes binding (fails loudly when the library is missing)
  nerfacc.py                drop-in for `nerfacc` / `lib.nerfacc` (march, scan, accumulate, occ grid)
  tinycudann.py             drop-in for `tinycudann.Encoding` (HashGrid, SphericalHarmonics)
  network_utils.py, geometry.py, texture.py, neus.py
                            host-side mirrors of the reference's models/*.py for this path
  synthetic.py              seeded synthetic rays / grids / env maps
"""
__version__ = "0.1.0"

import os as _os

# The sample count of a step is data dependent (occupancy grid, jitter, visibility filter), so torch's caching allocator
# sees slightly different sizes for the same [S, ...] tensors every step.  With fixed-size segments it answers by
# cudaMalloc-ing new multi-GB blocks whenever S grows (measured on the split-sum training step: 72 -> 84 GB reserved,
# 25-55 ms outlier steps); expandable segments grow in place (30 GB reserved, outliers gone).  Takes effect when set
# before the first CUDA allocation of the process; an explicit user setting wins.
_os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
