"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/masking.npz by running the REFERENCE's own
``src/utils/masking.py`` functions (imported from /root/reference through oracle/ref_harness.py) on seeded inputs.
Re-run in the build container:  ``python -m oracle.make_golden_masking``.

The sequence follows ``demo.py:138-167``: FLAME -> Renderer -> ``mesh_based_mask_uniform_faces`` (first with its own
random draws to obtain ``coords``, then through the deterministic ``coords`` path) -> point mask -> ``masking``.
Random inputs that the oracle takes explicitly are regenerated in the tests from the seeds stored here (torch's
CPU generator is deterministic for a given torch version; the tests skip if the stored probe value disagrees).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from smirk_b200 import synth_assets, synth_inputs  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
N = lambda t: t.detach().cpu().numpy()
SEED_SAMPLE, SEED_MASK, SEED_IMG = 7001, 7002, 7003


def main():
    root = synth_assets.materialize(os.path.join(tempfile.gettempdir(), "smk_assets_golden"))
    with ref_harness.reference(root) as R, torch.no_grad():
        import src.utils.masking as M
        flame, rend = R.FLAME(), R.Renderer()
        B, S = 2, 224
        p = synth_inputs.flame_params(B, 501)
        fo = flame.forward(p)
        ro = rend.forward(fo["vertices"], p["cam"], landmarks_fan=fo["landmarks_fan"], landmarks_mp=fo["landmarks_mp"])
        tv, faces = ro["transformed_vertices"], flame.faces_tensor
        g = torch.Generator().manual_seed(11)
        base_prob = (torch.rand(faces.shape[0], generator=g) > 0.6).float() * torch.tensor([0.5, 1.0])[torch.randint(0, 2, (faces.shape[0],), generator=g)]
        # 1. the reference's own sampling (global RNG) -> coords
        torch.manual_seed(SEED_SAMPLE)
        npoints_a, coords = M.mesh_based_mask_uniform_faces(tv, flame_faces=faces, face_probabilities=base_prob, mask_ratio=0.05)
        # 2. the deterministic path the oracle / kernels implement
        npoints, _ = M.mesh_based_mask_uniform_faces(tv, flame_faces=faces, face_probabilities=base_prob, mask_ratio=0.05, coords=coords)
        assert torch.equal(npoints, npoints_a)
        # face weights as the reference computes them before multinomial (masking.py:146-160), re-evaluated with its helpers
        fe = faces.expand(B, -1, -1)
        nrm = R.util.vertex_normals(tv, fe)
        fnz = R.util.face_vertices(nrm, fe)[:, :, :, 2].mean(dim=-1)
        w = torch.where(fnz < 0.05, base_prob.repeat(B, 1), torch.zeros_like(fnz)) * M.triangle_area(R.util.face_vertices(tv, fe))
        # 3. point mask (demo.py:154-160) with fixed rbound, masking() with seeded draws
        rbound = torch.tensor([npoints.shape[1] // 3, npoints.shape[1]])
        pmask = torch.zeros(B, 1, S, S)
        for bi in range(B):
            pmask[bi, :, npoints[bi, :rbound[bi], 1], npoints[bi, :rbound[bi], 0]] = 1
        img = synth_inputs.images(B, SEED_IMG)
        rendered_mask = 1 - (ro["rendered_img"] == 0).all(dim=1, keepdim=True).float()
        hull = torch.ones(B, 1, S, S)
        hull[0, :, 40:190, 50:180] = 0
        hull[1, :, 30:200, 60:170] = 0
        extra = img * pmask
        torch.manual_seed(SEED_MASK)
        masked = M.masking(img, hull, extra, 10, rendered_mask=rendered_mask)
        masked_plain = M.masking(img, hull, extra, 10, rendered_mask=None, extra_noise=False, random_mask=0)
        # 4. transfer_pixels with duplicate targets
        p1 = npoints[:, :400]
        p2 = torch.flip(npoints[:, :400], dims=[1]) // 2
        tp = M.transfer_pixels(img, p1, p2)
        tpb = M.transfer_pixels(img, p1, p2, rbound=torch.tensor([100, 400]))
        torch.manual_seed(SEED_MASK)
        probe = float(torch.randn(3)[2])
        np.savez_compressed(os.path.join(GOLD, "masking.npz"),
                            params_seed=np.int64(501), base_prob=N(base_prob), trans_verts=N(tv), rendered_img_nonzero=N(rendered_mask).astype(np.uint8),
                            face_weights=N(w), sampled_faces_indices=N(coords["sampled_faces_indices"]).astype(np.int32),
                            barycentric_coords=N(coords["barycentric_coords"]), npoints=N(npoints).astype(np.int16),
                            rbound=N(rbound), hull=N(hull).astype(np.uint8),
                            masked_sub=N(masked[:, :, 1::2, ::2]), masked_sum=N(masked.double().sum((2, 3))),
                            masked_plain_sub=N(masked_plain[:, :, ::2, 1::2]), masked_plain_sum=N(masked_plain.double().sum((2, 3))),
                            transfer=N(tp), transfer_rbound=N(tpb), p1=N(p1).astype(np.int16), p2=N(p2).astype(np.int16),
                            seeds=np.array([SEED_SAMPLE, SEED_MASK, SEED_IMG]), rng_probe=np.float64(probe))
    print("masking.npz", os.path.getsize(os.path.join(GOLD, "masking.npz")))


if __name__ == "__main__":
    main()
