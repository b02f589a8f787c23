"""Masking between Renderer and SmirkGenerator on the GPU.

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
        h = C.c_void_p()
        desc = SmkMaskingDesc(self.n_verts, self.n_faces, f.data_ptr())
        _lib.check(L.smk_masking_create(C.byref(desc), C.byref(h)), "smk_masking_create")
        self._h = _lib.NativeHandle(h, "smk_masking_destroy")

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
    npts = npoints[..., :2].to(img.device, torch.int64).contiguous()      # the reference's npoints may carry a third (z) column
    rb = rbound.to(img.device, torch.int64).contiguous()
    ws, n = ctx.workspace(B, S, img.device)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    _lib.check(_lib.lib().smk_masking_compose(ctx._h, img.data_ptr(), ptr(keep[0]), npts.data_ptr(), rb.data_ptr(), npts.shape[1], 0,
                                              ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), int(wr), B, S, out.data_ptr(),
                                              ws.data_ptr(), n, _lib.stream_ptr(img.device)), "smk_masking_compose")
    return out


def transfer_pixels(img, points1, points2, rbound=None):
    """masking.py:116-129 — same arguments and result.  With duplicate targets the last pair in index order wins."""
    _lib.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    p1 = points1[..., :2].to(img.device, torch.int64).contiguous()
    p2 = points2[..., :2].to(img.device, torch.int64).contiguous()
    rb = rbound.to(img.device, torch.int64).contiguous() if rbound is not None else None
    out = torch.empty_like(img)
    if B == 0:
        return out
    ws = torch.empty(B * S * S, dtype=torch.int32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().smk_masking_transfer_pixels(img.data_ptr(), p1.data_ptr(), p2.data_ptr(), rb.data_ptr() if rb is not None else 0,
                                                          B, p1.shape[1], S, out.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                                          _lib.stream_ptr(img.device)), "smk_masking_transfer_pixels")
    return out


def masking(img, mask, extra_points, wr=15, rendered_mask=None, extra_noise=True, random_mask=0.01, flame_faces=None, n_verts=5023):
    """masking.py:71-102 — same arguments and result; the two random draws are made with torch on the image's device
    (as the reference does), the dilation / composite runs in ``csrc/masking.cu``."""
    _lib.require_cuda(img, "img")
    img = img.float().contiguous()
    B, _, S, _ = img.shape
    ctx = _context(flame_faces, n_verts) if flame_faces is not None else (next(iter(_CTX.values())) if _CTX else MaskingContext(torch.tensor([[0, 1, 2]]), 3))
    noise = (torch.randn_like(img) * 0.05 + 1) if extra_noise else None
    centres = torch.bernoulli(torch.ones((B, 1, S, S), device=img.device) * random_mask) if random_mask > 0 else None
    f = lambda t: None if t is None else t.to(img.device, torch.float32).contiguous()
    hull, extra, rmask = f(mask), f(extra_points), f(rendered_mask)
    out = torch.empty_like(img)
    if B == 0:
        return out
    ptr = lambda t: 0 if t is None else t.data_ptr()
    with torch.cuda.device(img.device):
        ws, n = ctx.workspace(B, S, img.device)
        _lib.check(_lib.lib().smk_masking_compose(ctx._h, img.data_ptr(), hull.data_ptr(), 0, 0, 0, extra.data_ptr(), ptr(rmask), ptr(noise), ptr(centres),
                                                  int(wr), B, S, out.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr(img.device)), "smk_masking_compose")
    return out


class MaskingStage:
    """The masking step of the full cycle (``demo.py:138-165``) as one capturable device call:
    ``rendered_img, transformed_vertices, img, hull_mask -> masked_img`` with every random draw made on the device by a
    counter-based generator (``csrc/masking.cu``, Philox4x32-10).  ``rng_state`` = (seed, call counter) lives in device
    memory; each call advances the counter on the stream, so a CUDA-graph replay draws fresh samples.

    Parameters follow the script: ``mask_ratio = 0.01``, ``mask_ratio_mul = 5`` (upper bound on the sampled points),
    ``mask_dilation_radius = 10``; ``extra_noise`` / ``random_mask`` are ``masking()``'s defaults (masking.py:71)."""

    def __init__(self, flame_faces, face_probabilities, n_verts=5023, mask_ratio=0.01, mask_ratio_mul=5, mask_dilation_radius=10,
                 extra_noise=True, random_mask=0.01, image_size=224, seed=0):
        self.faces = flame_faces.detach().to("cpu", torch.int64)
        self.n_verts, self.S = int(n_verts), int(image_size)
        self.base_prob_host = face_probabilities.detach().to("cpu", torch.float32).contiguous()
        self.ratio_mul, self.wr = float(mask_ratio_mul), int(mask_dilation_radius)
        self.N = int(mask_ratio * mask_ratio_mul * image_size * image_size)
        self.extra_noise, self.p_centre, self.seed = bool(extra_noise), float(random_mask), int(seed)
        self._ctx, self._dev, self._ws = None, {}, _lib.Workspace()

    def __deepcopy__(self, memo):
        new = MaskingStage.__new__(MaskingStage)
        new.__dict__.update(self.__dict__)
        new._ctx, new._dev, new._ws = None, {}, _lib.Workspace()
        return new

    @property
    def _handle(self):                          # for SmirkPipeline's keep-alive list
        return self._ctx

    def _state(self, device):
        if self._ctx is None:
            with torch.cuda.device(device):
                self._ctx = MaskingContext(self.faces, self.n_verts)
        if device not in self._dev:
            rng = torch.tensor([self.seed, 0], dtype=torch.int64, device=device)
            self._dev[device] = (self.base_prob_host.to(device), rng)
        return self._dev[device]

    def reseed(self, seed, counter=0):
        self.seed = int(seed)
        for _, rng in self._dev.values():
            rng.copy_(torch.tensor([self.seed, int(counter)], dtype=torch.int64))

    @torch.no_grad()
    def forward(self, img, hull_mask, transformed_vertices, rendered_img, debug=False):
        _lib.require_cuda(img, "img")
        dev = img.device
        img, hull = _lib.dev_f32(img, "img"), _lib.dev_f32(hull_mask, "hull_mask")
        tv, rend = _lib.dev_f32(transformed_vertices, "transformed_vertices"), _lib.dev_f32(rendered_img, "rendered_img")
        B, S, N = img.shape[0], self.S, self.N
        if tuple(img.shape[1:]) != (3, S, S) or tuple(rend.shape) != tuple(img.shape) or hull.numel() != B * S * S or tuple(tv.shape) != (B, self.n_verts, 3):
            raise RuntimeError("smirk_b200.MaskingStage: expected img/rendered [B,3,%d,%d], hull [B,1,%d,%d], transformed_vertices [B,%d,3]" % (S, S, S, S, self.n_verts))
        base_prob, rng = self._state(dev)
        out = torch.empty_like(img)
        dbg = {}
        if debug:
            i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=dev)
            f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            dbg = dict(sampled_faces_indices=i64(B, N), barycentric_coords=f32(B, N, 3), npoints=i64(B, N, 2), rbound=i64(B),
                       noise_mult=f32(B, 3, S, S), random_centres=f32(B, 1, S, S))
        L = _lib.lib()
        P = lambda k: _lib.ptr(dbg.get(k))
        with torch.cuda.device(dev):
            n = int(L.smk_masking_forward_workspace_bytes(self._ctx._h, B, S, N))
            ws = self._ws.get(n, dev)
            _lib.check(L.smk_masking_forward(self._ctx._h, _lib.ptr(img), _lib.ptr(hull), _lib.ptr(tv), _lib.ptr(rend), _lib.ptr(base_prob),
                                             B, S, N, self.wr, self.ratio_mul, self.p_centre, 1 if self.extra_noise else 0, _lib.ptr(rng),
                                             _lib.ptr(out), P("sampled_faces_indices"), P("barycentric_coords"), P("npoints"), P("rbound"),
                                             P("noise_mult"), P("random_centres"), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                       "smk_masking_forward")
        return (out, dbg) if debug else out

    __call__ = forward
