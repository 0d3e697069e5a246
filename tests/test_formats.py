"""§8f f4: on-disk formats (Lightning checkpoint key layout, raw BSDF table, Radiance .hdr) -- host code,
runs without a GPU.  The .hdr codec is pinned to OpenCV, the decoder the reference itself calls
(lib/pbr/utils/nvdiffrecmc_util.py:380-392)."""
import os
import warnings

import numpy as np
import pytest
import torch

from rise_sdf_b200 import formats as fm


def _hdr_image(h=37, w=64, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0.0, 2.0, size=(h, w, 3)).astype(np.float32)
    img[5:9, 10:30] = 50.0                 # sun disc: long runs for the RLE coder
    img[20:, :] = 0.05                     # dark ground
    img[0, 0] = 0.0                        # exact black (exponent byte 0)
    img[1, 1] = (1e-6, 3e4, 1.0)           # wide per-pixel range
    return img


@pytest.mark.parametrize("rle", [True, False])
def test_hdr_write_read_round_trip(tmp_path, rle):
    img = _hdr_image()
    p = fm.write_hdr(str(tmp_path / "env.hdr"), img, rle=rle)
    back = fm.read_hdr(p)
    assert back.shape == img.shape and back.dtype == np.float32
    peak = img.max(axis=-1, keepdims=True)
    assert np.all(np.abs(back - img) <= peak / 128.0 + 1e-30)      # 8-bit mantissa shared per pixel
    assert np.array_equal(back[0, 0], [0, 0, 0])
    assert np.array_equal(fm.load_image(p), back)


def test_hdr_reader_and_writer_agree_with_opencv(tmp_path):
    cv2 = pytest.importorskip("cv2")
    img = _hdr_image(seed=1)
    ours = fm.write_hdr(str(tmp_path / "ours.hdr"), img)
    with open(ours, "rb") as f:
        bgr = cv2.imdecode(np.frombuffer(f.read(), np.uint8), cv2.IMREAD_UNCHANGED)
    assert bgr is not None
    assert np.array_equal(fm.read_hdr(ours), bgr[..., ::-1])        # the reference's read_hdr on our file
    theirs = str(tmp_path / "cv.hdr")
    assert cv2.imwrite(theirs, np.ascontiguousarray(img[..., ::-1]))
    assert np.array_equal(fm.read_hdr(theirs), cv2.imread(theirs, cv2.IMREAD_UNCHANGED)[..., ::-1])


def test_hdr_reader_against_committed_opencv_golden():
    """tests/golden/env_cv2_rle.hdr was written by OpenCV (RLE scanlines) and env_cv2_rle_decoded.npy is what
    cv2.imdecode -- the reference's decoder -- returns for it (tests/golden/make_hdr_golden.py): bit-exact."""
    here = os.path.join(os.path.dirname(__file__), "golden")
    got = fm.read_hdr(os.path.join(here, "env_cv2_rle.hdr"))
    want = np.load(os.path.join(here, "env_cv2_rle_decoded.npy"))
    assert got.dtype == np.float32 and got.shape == want.shape == (24, 48, 3)
    assert np.array_equal(got, want)
    assert np.array_equal(got[0, :5], np.zeros((5, 3), np.float32)) and float(got.max()) >= 1e3


def test_hdr_rejects_other_files(tmp_path):
    p = tmp_path / "x.hdr"
    p.write_bytes(b"P6\n1 1\n255\n\0\0\0")
    with pytest.raises(ValueError):
        fm.read_hdr(str(p))
    with pytest.raises(NotImplementedError):
        fm.load_image(str(tmp_path / "x.png"))


def test_bsdf_lut_round_trip(tmp_path):
    from rise_sdf_b200.synthetic import bsdf_lut
    lut = bsdf_lut(res=256, n_samples=8)
    p = fm.save_bsdf_lut(str(tmp_path / "bsdf_256_256.bin"), lut)
    assert os.path.getsize(p) == 256 * 256 * 2 * 4
    back = fm.load_bsdf_lut(p)
    assert back.shape == (1, 256, 256, 2) and torch.equal(back, lut.reshape(1, 256, 256, 2).float())
    raw = np.fromfile(p, dtype=np.float32).reshape(1, 256, 256, 2)             # models/texture.py:285
    assert np.array_equal(raw, back.numpy())
    with open(p, "ab") as f:
        f.write(b"\0\0\0\0")
    with pytest.raises(ValueError):
        fm.load_bsdf_lut(p)


def _neus():
    from rise_sdf_b200.neus import NeuSModel, neus_blender_config
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return NeuSModel(neus_blender_config())


def test_checkpoint_key_layout_and_round_trip(tmp_path):
    torch.manual_seed(0)
    a, b = _neus(), _neus()
    with torch.no_grad():
        a.geometry.encoding.encoding.params.uniform_(-1e-4, 1e-4)
        a.variance.variance.fill_(0.41)
        a.occupancy_grid.binaries[0, 3, 4, 5] = True
    opt = torch.optim.Adam(a.parameters(), lr=0.01)
    path = fm.save_checkpoint(str(tmp_path / "epoch=0-step=7.ckpt"), a, optimizer=opt, epoch=0, global_step=7)
    ckpt = torch.load(path, weights_only=False)
    keys = set(ckpt["state_dict"])
    for k in ("model.geometry.encoding.encoding.params", "model.geometry.network.layers.0.weight_g",
              "model.geometry.network.layers.4.weight_v", "model.texture.network.layers.8.weight",
              "model.variance.variance", "model.occupancy_grid.binaries", "model.occupancy_grid.occs",
              "model.occupancy_grid.aabbs"):                      # the reference system's key names
        assert k in keys, k
    assert ckpt["global_step"] == 7 and "optimizer_states" in ckpt
    info = fm.load_checkpoint(path, b, weights_only=True, strict=True)
    assert info["global_step"] == 7 and not info["missing"] and not info["unexpected"]
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka


def test_reference_state_dict_quirks():
    """fp16 tiny-cuda-nn parameters, uint8 occupancy bytes, buffers of nerfacc 0.5.3 we do not keep."""
    a, b = _neus(), _neus()
    with torch.no_grad():
        a.geometry.encoding.encoding.params.uniform_(-1e-4, 1e-4)
    sd = fm.to_reference_state_dict(a)
    sd["model.geometry.encoding.encoding.params"] = sd["model.geometry.encoding.encoding.params"].half()
    sd["model.occupancy_grid.binaries"] = sd["model.occupancy_grid.binaries"].to(torch.uint8)
    sd["model.occupancy_grid.grid_coords"] = torch.zeros(8, 3, dtype=torch.long)
    del sd["model.variance.variance"]
    missing, unexpected = fm.load_reference_state_dict(b, sd)
    assert missing == ["variance.variance"] and unexpected == ["model.occupancy_grid.grid_coords"]
    p = b.geometry.encoding.encoding.params
    assert p.dtype == torch.float32 and torch.equal(p, a.geometry.encoding.encoding.params.half().float())
    assert b.occupancy_grid.binaries.dtype == torch.bool
    with pytest.raises(RuntimeError):
        fm.load_reference_state_dict(b, sd, strict=True)
    sd["model.geometry.network.layers.0.bias"] = torch.zeros(7)
    with pytest.raises(RuntimeError):
        fm.load_reference_state_dict(b, sd)


def test_png_writer_round_trips_through_opencv(tmp_path):
    """formats.save_png (stdlib zlib) is decoded by cv2.imread to the same 8-bit pixels."""
    import cv2
    import numpy as np
    from rise_sdf_b200 import formats
    g = np.random.default_rng(0)
    img = g.random((37, 53, 3)).astype(np.float32)
    p = str(tmp_path / "frame.png")
    formats.save_png(p, img)
    back = cv2.imread(p, cv2.IMREAD_COLOR)[..., ::-1]
    assert back.shape == (37, 53, 3)
    assert np.array_equal(back, (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8))
    formats.save_png(str(tmp_path / "depth.png"), img[..., 0])
    assert cv2.imread(str(tmp_path / "depth.png"), cv2.IMREAD_GRAYSCALE).shape == (37, 53)
