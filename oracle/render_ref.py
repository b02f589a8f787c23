"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU fp32 restatement of ``Renderer.forward`` (src/renderer/renderer.py:100-207,239-250) and
``src/renderer/util.py``.  The torch glue is Tier A (checked against the reference class itself);
the rasteriser is the Tier-B C restatement in ``raster_ref.c`` (pytorch3d is third-party and absent).
"""
import ctypes
import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from . import build as _build

_lib = None


def _raster_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def rasterize_ref(face_verts, N, Fm, H=224, W=224, brute=False):
    """face_verts: float32 [N*F,3,3] in pytorch3d NDC -> (pix_to_face i64 [N,H,W,1], zbuf, bary [N,H,W,1,3], dists)."""
    fv = np.ascontiguousarray(face_verts.detach().cpu().numpy().astype(np.float32))
    p2f = np.empty((N, H, W), np.int64)
    zb = np.empty((N, H, W), np.float32)
    bc = np.empty((N, H, W, 3), np.float32)
    ds = np.full((N, H, W), -1, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib = _raster_lib()
    if brute:
        lib.smk_oracle_rasterize_bruteforce(P(fv), N, Fm, H, W, P(p2f), P(zb), P(bc))
    else:
        lib.smk_oracle_rasterize(P(fv), N, Fm, H, W, P(p2f), P(zb), P(bc), P(ds))
    return (torch.from_numpy(p2f)[..., None], torch.from_numpy(zb)[..., None],
            torch.from_numpy(bc)[:, :, :, None, :], torch.from_numpy(ds)[..., None])


def parse_obj(path):
    """What the reference takes from pytorch3d.io.load_obj (renderer.py:54-57): verts, verts_idx,
    textures_idx, verts_uvs."""
    v, vt, f, ft = [], [], [], []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith("v "):
                v.append([float(x) for x in ln.split()[1:4]])
            elif ln.startswith("vt "):
                vt.append([float(x) for x in ln.split()[1:3]])
            elif ln.startswith("f "):
                t = [s.split("/") for s in ln.split()[1:]]
                f.append([int(s[0]) - 1 for s in t[:3]])
                ft.append([int(s[1]) - 1 for s in t[:3]])
    return (torch.tensor(v, dtype=torch.float32), torch.tensor(f, dtype=torch.int64),
            torch.tensor(ft, dtype=torch.int64), torch.tensor(vt, dtype=torch.float32))


class RenderConstants:
    """renderer.py:50-98 with render_full_head=False: crop the topology to the FLAME 'face' mask."""

    def __init__(self, root="."):
        A = os.path.join(root, "assets")
        _, faces, _, _ = parse_obj(os.path.join(A, "head_template.obj"))
        with open(os.path.join(A, "FLAME_masks", "FLAME_masks.pkl"), "rb") as fh:
            masks = pickle.load(fh, encoding="latin1")
        self.final_mask = masks["face"].tolist()
        keep = torch.unique(torch.tensor(self.final_mask, dtype=torch.long))          # renderer.py:24-28
        nv = int(faces.max()) + 1
        remap = torch.full((nv,), -1, dtype=torch.long)
        remap[keep] = torch.arange(len(keep))
        ok = (remap[faces] != -1).all(1)
        self.faces = remap[faces[ok]][None]                                           # [1,3408,3]
        self.image_size = 224


def vertex_normals_ref(vertices, faces):
    """util.py:30-62 (area-weighted, each corner adds the same un-normalised face normal)."""
    B, nv = vertices.shape[:2]
    fl = (faces + (torch.arange(B) * nv)[:, None, None]).reshape(-1, 3)
    vf = vertices.reshape(B * nv, 3)[fl]
    n = torch.zeros(B * nv, 3)
    n.index_add_(0, fl[:, 1], torch.cross(vf[:, 2] - vf[:, 1], vf[:, 0] - vf[:, 1], dim=-1))
    n.index_add_(0, fl[:, 2], torch.cross(vf[:, 0] - vf[:, 2], vf[:, 1] - vf[:, 2], dim=-1))
    n.index_add_(0, fl[:, 0], torch.cross(vf[:, 1] - vf[:, 0], vf[:, 2] - vf[:, 0], dim=-1))
    return F.normalize(n, eps=1e-6, dim=1).reshape(B, nv, 3)


def orth_proj_ref(X, cam):
    """util.py:64-78 followed by the y/z sign flip of renderer.py:102."""
    cam = cam.view(-1, 1, 3)
    Xt = torch.cat([X[:, :, :2] + cam[:, :, 1:], X[:, :, 2:]], 2)
    Xn = cam[:, :, 0:1] * Xt
    Xn[:, :, 1:] = -Xn[:, :, 1:]
    return Xn


LIGHT_DIRS = torch.tensor([[-1., 1, 1], [1, 1, 1], [-1, -1, 1], [1, -1, 1], [0, 0, 1]])   # renderer.py:127-135


def render_forward_ref(rc, vertices, cam, brute=False, **landmarks):
    """Renderer.forward (renderer.py:100-118) -> dict incl. the rasteriser's raw outputs."""
    B = vertices.shape[0]
    tv = orth_proj_ref(vertices, cam)
    out = {k: orth_proj_ref(v, cam)[..., :2] for k, v in landmarks.items()}
    tvm = tv[:, rc.final_mask, :].clone()                  # copy: the +10 below does not leak (renderer.py:140-144)
    vm = vertices[:, rc.final_mask, :]
    tvm[:, :, 2] = tvm[:, :, 2] + 10
    faces = rc.faces.expand(B, -1, -1)
    normals = vertex_normals_ref(vm, faces)                # on UNtransformed vertices (renderer.py:147)
    nv = vm.shape[1]
    fl = (faces + (torch.arange(B) * nv)[:, None, None])
    face_normals = normals.reshape(B * nv, 3)[fl]          # [B,F,3,3]
    fixed = tvm.clone()
    fixed[..., :2] = -fixed[..., :2]                       # renderer.py:172-173
    face_verts = fixed.reshape(B * nv, 3)[fl].reshape(-1, 3, 3)
    Fm = faces.shape[1]
    p2f, zbuf, bary, dists = rasterize_ref(face_verts, B, Fm, rc.image_size, rc.image_size, brute=brute)
    # attribute interpolation, renderer.py:194-207 (colour is the constant 180/255)
    attr = torch.cat([torch.full((B * Fm, 3, 3), 180.0 / 255.0), face_normals.reshape(B * Fm, 3, 3)], -1)
    mask = p2f == -1
    idx = p2f.clone()
    idx[mask] = 0
    vals = attr[idx.view(-1)].view(B, 224, 224, 1, 3, 6)
    pix = (bary[..., None] * vals).sum(-2)
    pix[mask] = 0
    pix = pix[:, :, :, 0].permute(0, 3, 1, 2)
    albedo, nimg = pix[:, :3], pix[:, 3:6]
    # shading, renderer.py:158-166,239-250
    nrm = nimg.permute(0, 2, 3, 1).reshape(B, -1, 3)
    ld = F.normalize(LIGHT_DIRS[None, :, None, :].expand(B, -1, nrm.shape[1], -1), dim=3)
    ndl = torch.clamp((nrm[:, None] * ld).sum(3), 0., 1.)
    shading = (ndl[..., None] * 1.7).expand(-1, -1, -1, 3).mean(1)
    shading = shading.reshape(B, 224, 224, 3).permute(0, 3, 1, 2)
    out.update(rendered_img=albedo * shading, transformed_vertices=tv,
               pix_to_face=p2f[..., 0], bary=bary[:, :, :, 0], zbuf=zbuf[..., 0], normals=normals)
    return out
