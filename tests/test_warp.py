"""Crop / warp: oracle properties and host logic on the CPU, kernel-vs-oracle parity on the GPU.

The reference's arithmetic lives in scikit-image, which is absent here (parity unpinned, see oracle/warp_ref.py); the
CPU tests pin the restatement to what can be checked independently: identity and integer shifts are exact, the
interior agrees with scipy.ndimage.map_coordinates(order=1), Umeyama recovers a known similarity, crop_face maps the
landmark square onto the crop corners.  The GPU tests demand bit-exact uint8 agreement with the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import warp_ref
from smirk_b200 import crop

DEV = "cuda:0"


def _frame(rng, H, W, lo=0, hi=256):
    return rng.integers(lo, hi, size=(H, W, 3), dtype=np.uint8)


def _similarity(scale, theta, tx, ty):
    c, s = np.cos(theta) * scale, np.sin(theta) * scale
    return np.array([[c, -s, tx], [s, c, ty], [0.0, 0.0, 1.0]])


# ---------------------------------------------------------------------------------------------- CPU
def test_umeyama_recovers_a_similarity():
    rng = np.random.default_rng(0)
    T = _similarity(1.7, 0.3, 12.5, -40.25)
    src = rng.normal(size=(7, 2)) * 50
    dst = (np.hstack([src, np.ones((7, 1))]) @ T.T)[:, :2]
    for est in (warp_ref.umeyama(src, dst, True), crop.estimate_transform("similarity", src, dst).params):
        np.testing.assert_allclose(est, T, atol=1e-9)


def test_landmark_box_transform_maps_the_landmark_square_to_the_crop():
    lm = np.array([[300.0, 200.0], [500.0, 260.0], [420.0, 420.0], [310.0, 400.0]])
    t = crop.landmark_box_transform(lm, scale=1.4, image_size=224)
    np.testing.assert_allclose(t.params, warp_ref.crop_face_ref((720, 1280, 3), lm, 1.4, 224), atol=1e-12)
    size = int(((500 - 300) + (420 - 200)) / 2 * 1.4)
    cx, cy = 500 - 100.0, 420 - 110.0
    corners = np.array([[cx - size / 2, cy - size / 2], [cx - size / 2, cy + size / 2], [cx + size / 2, cy - size / 2]])
    np.testing.assert_allclose(t(corners), [[0, 0], [0, 223], [223, 0]], atol=1e-9)
    np.testing.assert_allclose(t.inverse.params @ t.params, np.eye(3), atol=1e-12)
    with pytest.raises(NotImplementedError):
        crop.estimate_transform("affine", corners, corners)


def test_warp_oracle_identity_and_integer_shift_are_exact():
    rng = np.random.default_rng(1)
    img = _frame(rng, 37, 53)
    assert np.array_equal(warp_ref.warp_ref(img, np.eye(3), (37, 53)), img)
    M = np.array([[1.0, 0, 5], [0, 1.0, -3], [0, 0, 1]])          # output (c, r) reads input (c + 5, r - 3)
    out = warp_ref.warp_ref(img, M, (37, 53))
    assert np.array_equal(out[3:, :48], img[:34, 5:])
    assert not out[:3].any() and not out[:, 48:].any()              # outside the source: cval = 0


def test_warp_oracle_interior_matches_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(2)
    img = _frame(rng, 90, 120)
    M = _similarity(0.8, 0.2, 20.0, 15.0)
    Ho, Wo = 40, 50
    out = warp_ref.warp_ref(img, M, (Ho, Wo))
    tfc, tfr = np.meshgrid(np.arange(Wo, dtype=np.float64), np.arange(Ho, dtype=np.float64))
    c = M[0, 0] * tfc + M[0, 1] * tfr + M[0, 2]
    r = M[1, 0] * tfc + M[1, 1] * tfr + M[1, 2]
    inside = (r >= 1) & (r <= 88) & (c >= 1) & (c <= 118)
    assert inside.mean() > 0.9
    for ch in range(3):
        ref = ndi.map_coordinates(img[..., ch].astype(np.float64), [r, c], order=1, mode="constant", cval=0.0)
        d = np.abs(out[..., ch].astype(np.float64) - np.floor(ref))[inside]
        assert (d <= 1).all() and (d == 0).mean() > 0.999          # same bilinear up to float64 rounding at integers


def test_warp_oracle_clip_rule_keeps_exact_zeros_only():
    img = np.full((8, 8, 3), 100, np.uint8)                          # min = max = 100 > cval = 0
    M = np.array([[1.0, 0, -0.5], [0, 1.0, 0], [0, 0, 1]])         # column 0 blends half cval, half image
    out = warp_ref.warp_ref(img, M, (8, 10))
    assert (out[:, 0] == 100).all()                                  # 50 is clipped up to the source minimum
    assert (out[:, 1:8] == 100).all() and (out[:, 9] == 0).all()     # fully outside: exact cval survives


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,lo", [(270, 480, 0), (1080, 1920, 0), (200, 333, 17)])
def test_crop_to_tensor_matches_oracle(native_lib, H, W, lo):
    rng = np.random.default_rng(10 + H)
    B = 3
    frames = np.stack([_frame(rng, H, W, lo) for _ in range(B)])
    tforms = []
    for i in range(B):
        cx, cy = W * (0.3 + 0.2 * i), H * (0.4 + 0.1 * i)
        half = min(H, W) * (0.2 + 0.15 * i)                           # the last crop reaches outside the frame
        lm = np.array([[cx - half, cy - half], [cx + half, cy + half * 0.9], [cx, cy]])
        tforms.append(crop.landmark_box_transform(lm, scale=1.4, image_size=224))
    tforms[1] = crop.SimilarityTransform(tforms[1].params @ _similarity(1.0, 0.35, 0.0, 0.0))   # a rotated crop as well
    got = crop.crop_to_tensor(torch.from_numpy(frames).to(DEV), tforms, 224).cpu().numpy()
    for i in range(B):
        ref = warp_ref.crop_to_tensor_ref(frames[i], tforms[i].params, 224)[0]
        assert np.array_equal(got[i], ref), "frame %d: %d of %d values differ" % (i, int((got[i] != ref).sum()), ref.size)


@pytest.mark.gpu
def test_warp_back_matches_oracle(native_lib):
    rng = np.random.default_rng(20)
    B, S, H, W = 2, 224, 360, 640
    rendered = rng.random((B, 3, S, S), dtype=np.float32)
    rendered[:, :, :40] = 0.0                                         # background rows, like a rendered mesh
    tforms = [crop.SimilarityTransform(_similarity(0.9, 0.1 * i, -150.0 - 30 * i, -60.0)) for i in range(B)]
    got = crop.warp_back(torch.from_numpy(rendered).to(DEV), tforms, (H, W)).cpu().numpy()
    for i in range(B):
        ref = warp_ref.warp_back_ref(rendered[i], tforms[i].params, (H, W))
        assert np.array_equal(got[i], ref), "frame %d: %d values differ" % (i, int((got[i] != ref).sum()))


@pytest.mark.gpu
def test_crop_rejects_bad_inputs(native_lib):
    f = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        crop.crop_to_tensor(f.float(), [np.eye(3)])
    with pytest.raises(ValueError):
        crop.crop_to_tensor(f, [np.eye(3), np.eye(3)])
    with pytest.raises(ValueError):
        crop.crop_to_tensor(f, [np.array([[1.0, 0, 0], [0, 1, 0], [1e-3, 0, 1]])])      # projective row
    assert crop.crop_to_tensor(f[:0], []).shape == (0, 3, 224, 224)
