"""GPU parity of the masking kernels against the reference-pinned golden fixture — WORK IN PROGRESS
(branch wip/masking-kernels): these have not run on a B200 yet.

Tolerances: the face weights go through vertex normals (1e-5 relative, and the `mean z < 0.05` threshold may flip for
faces within 1e-5 of it); the integer pixel coordinates come from a truncated fp32 value (a point may differ by one
pixel only where that value is within 1e-4 of an integer); the composite is exact given the same inputs.
"""
import os

import numpy as np
import pytest
import torch

from smirk_b200 import synth_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masking.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def faces(asset_root):
    from oracle import flame_ref
    return flame_ref.FlameConstants(asset_root).faces_tensor


def T(a, dt=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dt) if dt is not None else t


def test_face_weights(native_lib, g, faces):
    from smirk_b200 import masking
    w = masking.face_weights(T(g["trans_verts"]).to(DEV), faces, T(g["base_prob"])).cpu()
    ref = T(g["face_weights"])
    flipped = (w == 0) != (ref == 0)
    assert int(flipped.sum()) <= 3, "threshold flips: %d" % int(flipped.sum())
    ok = ~flipped
    assert torch.allclose(w[ok], ref[ok], rtol=1e-4, atol=1e-9)


def test_points_from_coords(native_lib, g, faces):
    from smirk_b200 import masking
    coords = {"sampled_faces_indices": T(g["sampled_faces_indices"], torch.long), "barycentric_coords": T(g["barycentric_coords"])}
    pts, _ = masking.mesh_based_mask_uniform_faces(T(g["trans_verts"]).to(DEV), faces, T(g["base_prob"]), mask_ratio=0.05, coords=coords)
    ref = T(g["npoints"], torch.long)
    d = (pts.cpu() - ref).abs()
    assert int(d.max()) <= 1 and int((d > 0).sum()) <= 4, "pixel mismatches: %d" % int((d > 0).sum())


def test_compose_matches_reference(native_lib, g, faces):
    from oracle import masking_ref
    from smirk_b200 import masking
    npoints, rbound = T(g["npoints"], torch.long), T(g["rbound"])
    img = synth_inputs.images(npoints.shape[0], int(g["seeds"][2]))
    hull, rmask = T(g["hull"], torch.float32), T(g["rendered_img_nonzero"], torch.float32)
    out = masking.masking_from_points(img.to(DEV), hull, npoints, rbound, wr=10, flame_faces=faces).cpu()
    assert torch.equal(out[:, :, ::2, 1::2], T(g["masked_plain_sub"]))
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn(img.shape, generator=gen) * 0.05 + 1
    centres = torch.bernoulli(torch.ones(img.shape[0], 1, 224, 224) * 0.01, generator=gen)
    out = masking.masking_from_points(img.to(DEV), hull, npoints, rbound, wr=10, rendered_mask=rmask, noise_mult=noise,
                                      random_centres=centres, flame_faces=faces).cpu()
    ref = masking_ref.masking_ref(img, hull, img * masking_ref.point_mask_ref(npoints, rbound, 224), 10, rendered_mask=rmask,
                                  noise_mult=noise, random_centres=centres)
    assert torch.equal(out, ref)
