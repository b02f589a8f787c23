"""TEST INFRASTRUCTURE ONLY — CPU restatement of the crop / warp steps around the per-frame path.

Reference call sites: ``demo.py:16-34`` (``crop_face``), ``demo.py:97,103-105`` and ``demo_video.py:128,134-137``
(crop + BGR->RGB + /255), ``demo_video.py:147-150`` (warp back to the frame).  The arithmetic lives in a third-party
dependency that is absent from this image and from /root/reference: **scikit-image** (``skimage.transform.warp`` and
``estimate_transform``; the reference's requirements do not pin a version — 0.19…0.24 share this code).  Its published
algorithm is restated here:

* ``estimate_transform('similarity', src, dst)`` = Umeyama's least-squares similarity (``skimage/transform/
  _geometric.py::_umeyama`` with ``estimate_scale=True``).
* ``warp(image, inverse_map, output_shape, preserve_range=True)`` with the defaults ``order=1`` (for non-bool images),
  ``mode='constant'``, ``cval=0``, ``clip=True``: the image is converted to float64, each channel goes through
  ``_warp_fast`` (``skimage/transform/_warps_cy.pyx``): ``(c, r) = M (tfc, tfr, 1)`` with ``M = inv(tform.params)``
  for ``tform.inverse`` (``M = tform.params`` when the transform itself is passed), then
  ``bilinear_interpolation`` (``skimage/_shared/interpolation.pxd``): floor/ceil neighbours, out-of-image neighbours
  read ``cval``, ``top = (1-dc) tl + dc tr``, ``bottom = (1-dc) bl + dc br``, ``v = (1-dr) top + dr bottom``;
  finally ``_clip_warp_output`` clips to ``[min(image), max(image)]`` but keeps exact ``cval`` pixels when ``cval`` is
  outside that range.  ``.astype(np.uint8)`` truncates.

**Parity unpinned**: skimage cannot be imported here, so nothing ties this restatement to the real library beyond
the properties in tests/test_oracle_golden.py (identity, integer shifts, agreement with scipy.ndimage.map_coordinates
away from the border, Umeyama recovering a known similarity).
"""
import numpy as np


def umeyama(src, dst, estimate_scale=True):
    src = np.asarray(src, np.float64); dst = np.asarray(dst, np.float64)
    num, dim = src.shape
    src_mean, dst_mean = src.mean(axis=0), dst.mean(axis=0)
    src_demean, dst_demean = src - src_mean, dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d) if estimate_scale else 1.0
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


def crop_face_ref(frame_shape, landmarks, scale=1.0, image_size=224):
    """demo.py:16-34 — returns the 3x3 params of the similarity transform frame -> crop."""
    landmarks = np.asarray(landmarks)
    left, right = np.min(landmarks[:, 0]), np.max(landmarks[:, 0])
    top, bottom = np.min(landmarks[:, 1]), np.max(landmarks[:, 1])
    old_size = (right - left + bottom - top) / 2
    center = np.array([right - (right - left) / 2.0, bottom - (bottom - top) / 2.0])
    size = int(old_size * scale)
    src_pts = np.array([[center[0] - size / 2, center[1] - size / 2], [center[0] - size / 2, center[1] + size / 2],
                        [center[0] + size / 2, center[1] - size / 2]])
    dst_pts = np.array([[0, 0], [0, image_size - 1], [image_size - 1, 0]])
    return umeyama(src_pts, dst_pts, True)


def warp_ref(image, M, output_shape):
    """image uint8 [H,W,C]; M 3x3 float64 map from output (col,row,1) to input (x,y,1); returns uint8 [Ho,Wo,C]."""
    img = np.asarray(image).astype(np.float64)
    H, W, C = img.shape
    Ho, Wo = output_shape
    M = np.asarray(M, np.float64)
    tfc, tfr = np.meshgrid(np.arange(Wo, dtype=np.float64), np.arange(Ho, dtype=np.float64))
    c = M[0, 0] * tfc + M[0, 1] * tfr + M[0, 2]
    r = M[1, 0] * tfc + M[1, 1] * tfr + M[1, 2]
    minr, minc = np.floor(r).astype(np.int64), np.floor(c).astype(np.int64)
    maxr, maxc = np.ceil(r).astype(np.int64), np.ceil(c).astype(np.int64)
    dr, dc = r - minr, c - minc

    def px(rr, cc):
        ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
        v = img[np.clip(rr, 0, H - 1), np.clip(cc, 0, W - 1)]
        return np.where(ok[..., None], v, 0.0)

    top = (1 - dc)[..., None] * px(minr, minc) + dc[..., None] * px(minr, maxc)
    bottom = (1 - dc)[..., None] * px(maxr, minc) + dc[..., None] * px(maxr, maxc)
    out = (1 - dr)[..., None] * top + dr[..., None] * bottom
    lo, hi = img.min(), img.max()
    preserve = not (lo <= 0.0 <= hi)
    mask = out == 0.0
    out = np.clip(out, lo, hi)
    if preserve:
        out[mask] = 0.0
    return out.astype(np.uint8)


def crop_to_tensor_ref(frame_bgr, tform_params, image_size=224):
    """demo.py:97,103-105: warp(image, tform.inverse) -> BGR2RGB -> [1,3,S,S] float32 / 255 (as a numpy array)."""
    crop = warp_ref(frame_bgr, np.linalg.inv(tform_params), (image_size, image_size))
    rgb = crop[..., ::-1]
    return (np.ascontiguousarray(rgb.transpose(2, 0, 1))[None].astype(np.float32) / np.float32(255.0))


def warp_back_ref(rendered, tform_params, out_hw):
    """demo_video.py:147-149: rendered float32 [3,S,S] -> uint8 [H,W,3] in frame coordinates."""
    u8 = (np.asarray(rendered, np.float32).transpose(1, 2, 0) * np.float32(255.0)).astype(np.uint8)
    return warp_ref(u8, np.asarray(tform_params, np.float64), out_hw)
