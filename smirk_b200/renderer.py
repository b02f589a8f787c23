"""``Renderer`` — drop-in for the reference class ``src/renderer/renderer.py:49-207`` (forward only).

Same constructor arguments, registered buffers (``state_dict`` keys: faces, face_colors, raw_uvcoords,
uvcoords, uvfaces, face_uvcoords, constant_factor) and ``forward`` output dict; the vertex stage,
normals, rasterisation (pytorch3d's ``rasterize_meshes`` in the reference) and shading run in
``csrc/render.cu`` through ``smk_renderer_forward``.
"""
import ctypes as C
import pickle

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def load_obj(path):
    """Minimal OBJ reader for what renderer.py:54-57 takes from pytorch3d.io.load_obj."""
    v, vt, f, ft = [], [], [], []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith("v "):
                v.append(ln.split()[1:4])
            elif ln.startswith("vt "):
                vt.append(ln.split()[1:3])
            elif ln.startswith("f "):
                corners = [c.split("/") for c in ln.split()[1:]]
                for k in range(1, len(corners) - 1):          # fan-triangulate polygons
                    tri = (corners[0], corners[k], corners[k + 1])
                    f.append([int(c[0]) - 1 for c in tri])
                    ft.append([int(c[1]) - 1 if len(c) > 1 and c[1] else -1 for c in tri])
    return (torch.tensor(np.array(v, dtype=np.float32)), torch.tensor(np.array(f, dtype=np.int64)),
            torch.tensor(np.array(ft, dtype=np.int64)), torch.tensor(np.array(vt, dtype=np.float32)))


def keep_vertices_and_update_faces(faces, vertices_to_keep):
    """renderer.py:11-47: drop faces touching removed vertices, renumber the rest."""
    keep = torch.unique(torch.as_tensor(vertices_to_keep, dtype=torch.long))
    n = int(faces.max()) + 1
    remap = torch.full((n,), -1, dtype=torch.long)
    remap[keep] = torch.arange(len(keep))
    return remap[faces[(remap[faces] != -1).all(1)]]


def _face_vertices(vertices, faces):
    bs, nv = vertices.shape[:2]
    return vertices.reshape(bs * nv, -1)[(faces + (torch.arange(bs) * nv)[:, None, None]).long()]


class Renderer(nn.Module):
    def __init__(self, render_full_head=False, obj_filename="assets/head_template.obj"):
        super().__init__()
        self.image_size = 224
        verts, faces, uvfaces, uvcoords = load_obj(obj_filename)
        uvcoords, uvfaces, faces = uvcoords[None], uvfaces[None], faces[None]
        self.render_full_head = render_full_head
        self.n_verts = verts.shape[0]
        colors = torch.tensor([180, 180, 180])[None, None, :].repeat(1, int(faces.max()) + 1, 1).float() / 255.
        with open("assets/FLAME_masks/FLAME_masks.pkl", "rb") as fh:
            self.flame_masks = pickle.load(fh, encoding="latin1")
        if not render_full_head:
            self.final_mask = self.flame_masks["face"].tolist()
            faces = keep_vertices_and_update_faces(faces[0], self.final_mask).unsqueeze(0)
            colors = colors[:, self.final_mask, :]
        else:
            self.final_mask = list(range(self.n_verts))
        self.register_buffer("faces", faces)
        self.register_buffer("face_colors", _face_vertices(colors, faces))
        self.register_buffer("raw_uvcoords", uvcoords)
        uvcoords = torch.cat([uvcoords, uvcoords[:, :, 0:1] * 0. + 1.], -1)
        uvcoords = uvcoords * 2 - 1
        uvcoords[..., 1] = -uvcoords[..., 1]
        self.register_buffer("uvcoords", uvcoords)
        self.register_buffer("uvfaces", uvfaces)
        self.register_buffer("face_uvcoords", _face_vertices(uvcoords, uvfaces))
        pi = np.pi
        self.register_buffer("constant_factor", torch.tensor(
            [1 / np.sqrt(4 * pi), ((2 * pi) / 3) * (np.sqrt(3 / (4 * pi))), ((2 * pi) / 3) * (np.sqrt(3 / (4 * pi))),
             ((2 * pi) / 3) * (np.sqrt(3 / (4 * pi))), (pi / 4) * (3) * (np.sqrt(5 / (12 * pi))),
             (pi / 4) * (3) * (np.sqrt(5 / (12 * pi))), (pi / 4) * (3) * (np.sqrt(5 / (12 * pi))),
             (pi / 4) * (3 / 2) * (np.sqrt(5 / (12 * pi))), (pi / 4) * (1 / 2) * (np.sqrt(5 / (4 * pi)))]).float())
        self._handle, self._handle_dev, self._ws = None, None, _lib.Workspace()

    def _native(self, device):
        sig = (str(device), self.faces._version, self.faces.data_ptr(), self.image_size)
        if self._handle is not None and self._handle_dev == sig:
            return self._handle
        self._release()
        L = _lib.lib()
        mask_np, mask_p = _lib.i32(np.asarray(self.final_mask))
        faces_np, faces_p = _lib.i32(self.faces[0])
        d = _lib.SmkRendererDesc()
        d.n_verts, d.n_mask, d.mask_ids = self.n_verts, len(self.final_mask), mask_p
        d.n_faces, d.faces, d.image_size = faces_np.shape[0], faces_p, self.image_size
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.smk_renderer_create(C.byref(d), C.byref(h)), "smk_renderer_create")
        self._handle, self._handle_dev = _lib.NativeHandle(h, "smk_renderer_destroy"), sig
        return self._handle

    def _release(self):
        self._handle = None                    # the native object dies with its last reference (_lib.NativeHandle)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        nn.Module.__init__(new)
        for k, v in self.__dict__.items():
            if k not in ("_handle", "_handle_dev", "_ws"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._handle, new._handle_dev, new._ws = None, None, _lib.Workspace()
        return new

    @torch.no_grad()
    def forward(self, vertices, cam_params, **landmarks):
        return self.render_full(vertices, cam_params, raw=False, **landmarks)

    @torch.no_grad()
    def render_full(self, vertices, cam_params, raw=True, **landmarks):
        """forward() plus, with ``raw=True``, the rasteriser's own outputs (pix_to_face int64 [B,S,S]
        packed like pytorch3d, bary [B,S,S,3], zbuf [B,S,S]) and the vertex normals [B,n_mask,3]."""
        _lib.require_cuda(vertices, "vertices")
        dev = vertices.device
        L = _lib.lib()
        h = self._native(dev)
        verts, cam = _lib.dev_f32(vertices, "vertices"), _lib.dev_f32(cam_params, "cam_params")
        B, S = verts.shape[0], self.image_size
        if verts.shape[1] != self.n_verts or cam.shape != (B, 3):
            raise RuntimeError("smirk_b200.Renderer: expected vertices [B,%d,3] and cam [B,3]" % self.n_verts)
        o = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        rendered, tverts = o(B, 3, S, S), o(B, self.n_verts, 3)
        p2f = torch.empty(B, S, S, dtype=torch.int64, device=dev) if raw else None
        bary, zbuf = (o(B, S, S, 3), o(B, S, S)) if raw else (None, None)
        normals = o(B, len(self.final_mask), 3) if raw else None
        out = {"rendered_img": rendered, "transformed_vertices": tverts}
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            ws = self._ws.get(L.smk_renderer_workspace_bytes(h, B), dev)
            _lib.check(L.smk_renderer_forward(h, _lib.ptr(verts), _lib.ptr(cam), B, _lib.ptr(rendered), _lib.ptr(tverts),
                                              _lib.ptr(p2f), _lib.ptr(bary), _lib.ptr(zbuf), _lib.ptr(normals),
                                              _lib.ptr(ws), ws.numel(), st), "smk_renderer_forward")
            for k, pts in landmarks.items():                                     # renderer.py:104-108
                pts = _lib.dev_f32(pts, k)
                xy = o(B, pts.shape[1], 2)
                _lib.check(L.smk_project_points(_lib.ptr(pts), _lib.ptr(cam), B, pts.shape[1], _lib.ptr(xy), st),
                           "smk_project_points")
                out[k] = xy
        if self.render_full_head:
            # renderer.py:140-144: with the full head the fancy-index is skipped, so the in-place `z += 10` of render()
            # also lands in the tensor the reference returns as `transformed_vertices`
            tverts[..., 2] += 10.0
        if raw:
            out.update(pix_to_face=p2f, bary=bary, zbuf=zbuf, normals=normals)
        return out
