"""Masking between Renderer and SmirkGenerator on the GPU — WORK IN PROGRESS, not yet validated on
a GPU (branch wip/masking-kernels).

Mirrors ``src/utils/masking.py``: ``mesh_based_mask_uniform_faces`` and ``masking`` keep the reference's names,
arguments and return values.  Random draws (multinomial, rand, randn, bernoulli) are made with torch on the tensors'
device, exactly where the reference makes them; the deterministic parts run in ``csrc/masking.cu``.
"""
import ctypes as C

import torch

from . import _lib


class SmkMaskingDesc(C.Structure):
    _fields_ = [("n_verts", C.c_int), ("n_faces", C.c_int), ("faces", C.c_void_p)]


class MaskingContext:
    """Native handle for one mesh topology (the reference passes ``flame.faces_tensor`` on every call)."""

    def __init__(self, faces, n_verts):
        f = faces.detach().to("cpu", torch.int32).contiguous()
        self._faces_host = f
        self.n_verts, self.n_faces = int(n_verts), int(f.shape[0])
        L = _lib.lib()
        self._h = C.c_void_p()
        desc = SmkMaskingDesc(self.n_verts, self.n_faces, f.data_ptr())
        _lib.check(L.smk_masking_create(C.byref(desc), C.byref(self._h)), "smk_masking_create")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().smk_masking_destroy(self._h)
        except Exception:
            pass

    def workspace(self, B, S, device):
        n = int(_lib.lib().smk_masking_workspace_bytes(self._h, B, S))
        return torch.empty(n, dtype=torch.uint8, device=device), n


_CTX = {}


def _context(flame_faces, n_verts):
    key = (flame_faces.data_ptr(), int(flame_faces.shape[0]), int(n_verts))
    if key not in _CTX:
        _CTX[key] = MaskingContext(flame_faces, n_verts)
    return _CTX[key]


def random_barycentric(num=1, device="cpu"):
    """masking.py:55-68 (drawn on ``device``)."""
    u, v = torch.rand(num, device=device), torch.rand(num, device=device)
    outside = u + v > 1
    u[outside], v[outside] = 1 - u[outside], 1 - v[outside]
    return torch.stack((1 - (u + v), u, v), dim=1)


def face_weights(flame_trans_verts, flame_faces, face_probabilities):
    """masking.py:146-160: sampling weight of every face, [B,F]."""
    _lib.require_cuda(flame_trans_verts, "flame_trans_verts")
    tv = flame_trans_verts.float().contiguous()
    B, V = tv.shape[:2]
    ctx = _context(flame_faces, V)
    L = _lib.lib()
    w = torch.empty(B, ctx.n_faces, dtype=torch.float32, device=tv.device)
    if B == 0:
        return w
    ws, n = ctx.workspace(B, 224, tv.device)
    bp = face_probabilities.to(tv.device, torch.float32).contiguous()
    _lib.check(L.smk_masking_face_weights(ctx._h, tv.data_ptr(), bp.data_ptr(), B, w.data_ptr(), ws.data_ptr(), n,
                                          _lib.stream_ptr(tv.device)), "smk_masking_face_weights")
    return w


def mesh_based_mask_uniform_faces(flame_trans_verts, flame_faces, face_probabilities, mask_ratio=0.1, coords=None, IMAGE_SIZE=224):
    """masking.py:132-181 — same arguments and return value ``(npoints, coords)``."""
    _lib.require_cuda(flame_trans_verts, "flame_trans_verts")
    tv = flame_trans_verts.float().contiguous()
    B, V = tv.shape[:2]
    num = int(mask_ratio * IMAGE_SIZE * IMAGE_SIZE)
    if coords is None:
        w = face_weights(tv, flame_faces, face_probabilities)
        idx = torch.multinomial(w, num, replacement=True)
        bary = random_barycentric(B * num, tv.device).view(B, num, 3)
    else:
        idx, bary = coords["sampled_faces_indices"], coords["barycentric_coords"]
    idx = idx.to(tv.device, torch.int64).contiguous()
    bary = bary.to(tv.device, torch.float32).contiguous()
    ctx = _context(flame_faces, V)
    npoints = torch.empty(B, idx.shape[1], 2, dtype=torch.int64, device=tv.device)
    if B and idx.shape[1]:
        _lib.check(_lib.lib().smk_masking_points(ctx._h, tv.data_ptr(), idx.data_ptr(), bary.data_ptr(), B, idx.shape[1], IMAGE_SIZE,
                                                 npoints.data_ptr(), _lib.stream_ptr(tv.device)), "smk_masking_points")
    return npoints, {"sampled_faces_indices": idx, "barycentric_coords": bary}


def masking_from_points(img, mask, npoints, rbound, wr=15, rendered_mask=None, noise_mult=None, random_centres=None, flame_faces=None, n_verts=5023):
    """demo.py:154-163 + masking.py:71-102 fused: point mask of the first ``rbound[b]`` points, dilation, noise, random
    patches, composite.  ``noise_mult`` / ``random_centres`` are the two random tensors of ``masking()`` (None = off)."""
    _lib.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    ctx = _context(flame_faces, n_verts) if flame_faces is not None else next(iter(_CTX.values()))
    out = torch.empty_like(img)
    if B == 0:
        return out
    P = lambda t: 0 if t is None else t.to(img.device, torch.float32).contiguous().data_ptr()
    keep = [t.to(img.device, torch.float32).contiguous() if t is not None else None for t in (mask, rendered_mask, noise_mult, random_centres)]
    npts = npoints.to(img.device, torch.int64).contiguous()
    rb = rbound.to(img.device, torch.int64).contiguous()
    ws, n = ctx.workspace(B, S, img.device)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    _lib.check(_lib.lib().smk_masking_compose(ctx._h, img.data_ptr(), ptr(keep[0]), npts.data_ptr(), rb.data_ptr(), npts.shape[1],
                                              ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), int(wr), B, S, out.data_ptr(),
                                              ws.data_ptr(), n, _lib.stream_ptr(img.device)), "smk_masking_compose")
    return out
