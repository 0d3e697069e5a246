"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under rise_sdf_b200/ imports it (checked by
tests/test_no_oracle_in_product.py).

Parity status (also stated in DESIGN.md):
  * march / AABB / scan:   pinned against the reference's own compiled kernels
                           (oracle/_ref/nerfacc_cuda.so run on a B200 -> tests/golden/march_*.npz)
                           and the docstring known-answers of lib/nerfacc/vol_rendering.py:430-434,496-500.
  * cubemap prefilter:     pinned against oracle/_ref/renderutils_plugin.so (tests/golden/cubemap_*.npz).
  * hash grid / SH (tiny-cuda-nn), OccGridEstimator.sampling (nerfacc 0.5.3), texture()
    (nvdiffrast): third-party sources absent from /root/reference and not installable here
    -> "parity unpinned": restated from their published algorithms.
"""
