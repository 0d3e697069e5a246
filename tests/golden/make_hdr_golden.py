"""Golden vector for the Radiance .hdr reader: a picture WRITTEN by OpenCV (new-style RLE scanlines) and the array
OpenCV itself DECODES from it -- cv2.imdecode is what the reference's read_hdr calls
(lib/pbr/utils/nvdiffrecmc_util.py:380-392).  Run in the build container (cv2 4.13):
    python tests/golden/make_hdr_golden.py
"""
import os

import cv2
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
h, w = 24, 48
img = rng.uniform(0.0, 1.5, size=(h, w, 3)).astype(np.float32)
img[3:6, 10:30] = 40.0                 # sun: runs
img[16:, :] = 0.03                     # ground: runs
img[0, :5] = 0.0                       # black pixels
img[1, 1] = (1e-5, 2e3, 0.5)
path = os.path.join(here, "env_cv2_rle.hdr")
assert cv2.imwrite(path, np.ascontiguousarray(img[..., ::-1]))
with open(path, "rb") as f:
    bgr = cv2.imdecode(np.frombuffer(f.read(), np.uint8), cv2.IMREAD_UNCHANGED)
np.save(os.path.join(here, "env_cv2_rle_decoded.npy"), np.ascontiguousarray(bgr[..., ::-1]))
print(path, os.path.getsize(path), "bytes")
