"""ORACLE — TEST INFRASTRUCTURE ONLY.  Runs ONLY in the build container (needs /root/reference).

Imports the reference's own hot-path modules so the restatements in this package and the committed
golden vectors can be pinned to what the reference's Python actually computes:

  * NumPy-2 shim for the aliases FLAME.py:19-25 touches (np.float_, np.complex_, np.unicode_);
  * stub ``pytorch3d`` exposing exactly the three symbols renderer.py:5-7 imports: ``Meshes`` (a
    record), ``load_obj`` (our OBJ parser) and ``rasterize_meshes`` (-> oracle/raster_ref.c);
  * stub ``timm`` whose ``create_model`` returns the restated backbone (oracle/encoder_ref.py);
  * the materialised asset tree (synthetic FLAME pkl + the shipped topology) as cwd, because the
    reference constructors use hard-coded relative paths.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SMIRK_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "src"))


def _install_stubs():
    for a, b in (("float_", np.float64), ("complex_", np.complex128), ("unicode_", np.str_)):
        if not hasattr(np, a):
            setattr(np, a, b)
    if "pytorch3d" not in sys.modules:
        from . import render_ref

        class Meshes:
            def __init__(self, verts, faces):
                self.verts, self.faces = verts, faces

        def load_obj(path):
            v, f, ft, vt = render_ref.parse_obj(path)
            return v, types.SimpleNamespace(verts_idx=f, textures_idx=ft), types.SimpleNamespace(verts_uvs=vt)

        def rasterize_meshes(meshes, image_size=224, blur_radius=0.0, faces_per_pixel=1, bin_size=None,
                             max_faces_per_bin=None, perspective_correct=False, **kw):
            assert blur_radius == 0.0 and faces_per_pixel == 1 and not perspective_correct
            B, V = meshes.verts.shape[:2]
            Fm = meshes.faces.shape[1]
            fl = meshes.faces + (torch.arange(B) * V)[:, None, None]
            fv = meshes.verts.reshape(B * V, 3)[fl].reshape(-1, 3, 3)
            rasterize_meshes.last_call = dict(image_size=image_size, bin_size=bin_size)
            return render_ref.rasterize_ref(fv, B, Fm, image_size, image_size)

        p3 = types.ModuleType("pytorch3d")
        st = types.ModuleType("pytorch3d.structures"); st.Meshes = Meshes
        io = types.ModuleType("pytorch3d.io"); io.load_obj = load_obj
        rd = types.ModuleType("pytorch3d.renderer")
        ms = types.ModuleType("pytorch3d.renderer.mesh"); ms.rasterize_meshes = rasterize_meshes
        rd.mesh = ms; p3.structures = st; p3.io = io; p3.renderer = rd
        sys.modules.update({"pytorch3d": p3, "pytorch3d.structures": st, "pytorch3d.io": io,
                            "pytorch3d.renderer": rd, "pytorch3d.renderer.mesh": ms})
    if "timm" not in sys.modules:
        from . import encoder_ref

        def create_model(name, pretrained=True, features_only=True):
            m = encoder_ref.BackboneRef(name)
            m.feature_info = [{"num_chs": m.num_chs}]
            return m

        tm = types.ModuleType("timm"); tm.create_model = create_model
        sys.modules["timm"] = tm
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")       # renderer/util.py:5 imports it, never uses it


@contextlib.contextmanager
def reference(asset_root):
    """Context manager: cwd = asset_root, /root/reference importable, stubs installed.
    Yields a namespace with the reference classes."""
    if not available():
        raise RuntimeError("reference checkout not available at %s" % REF)
    _install_stubs()
    old_cwd, old_dwb = os.getcwd(), sys.dont_write_bytecode
    sys.dont_write_bytecode = True            # /root/reference is read-only
    sys.path.insert(0, REF)
    os.chdir(asset_root)
    try:
        from src.FLAME.FLAME import FLAME
        from src.FLAME import lbs
        from src.renderer.renderer import Renderer
        from src.renderer import util
        from src.smirk_generator import SmirkGenerator
        from src.smirk_encoder import SmirkEncoder
        yield types.SimpleNamespace(FLAME=FLAME, lbs=lbs, Renderer=Renderer, util=util,
                                    SmirkGenerator=SmirkGenerator, SmirkEncoder=SmirkEncoder)
    finally:
        os.chdir(old_cwd)
        sys.path.remove(REF)
        sys.dont_write_bytecode = old_dwb
