"""Build tests/golden/flame_topology.npz from the reference's shipped assets.

TEST INFRASTRUCTURE (runs only in the build container, where /root/reference is mounted).
The GPU box has no /root/reference, so the mesh topology and landmark embeddings the hot
path needs are carried as one compact derived blob:

  reference asset (read here)                                  -> blob key(s)
  assets/head_template.obj (5023 v / 5118 vt / 9976 f)          -> verts, uvs, faces, uvfaces
  assets/FLAME_masks/FLAME_masks.pkl ['face'] (renderer.py:66-71)-> face_mask
  assets/l_eyelid.npy, r_eyelid.npy (FLAME.py:81-82)            -> {l,r}_eyelid_{idx,val}
  assets/landmark_embedding.npy (FLAME.py:94-101)               -> static_*, dynamic_*, full_*
  assets/mediapipe_landmark_embedding/*.npz (FLAME.py:111-113)  -> mp_*

smirk_b200.synth_assets.materialize() expands the blob back into an `assets/` tree in the
formats the reference constructors read, plus a *synthetic* FLAME2020/generic_model.pkl
(the licensed model is not available anywhere in this build).
"""
import os, pickle, sys
import numpy as np

REF = os.environ.get("SMIRK_REFERENCE", "/root/reference")


def parse_obj(path):
    v, vt, f, ft = [], [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("vt "):
                vt.append([float(t) for t in line.split()[1:3]])
            elif line.startswith("f "):
                toks = [t.split("/") for t in line.split()[1:4]]
                f.append([int(t[0]) - 1 for t in toks])
                ft.append([int(t[1]) - 1 for t in toks])
    return (np.asarray(v, np.float32), np.asarray(vt, np.float32),
            np.asarray(f, np.int16), np.asarray(ft, np.int16))


def sparse_rows(a):
    a = np.asarray(a).astype(np.float32)      # the reference casts f64 -> f32 (FLAME.py:81)
    idx = np.nonzero(np.abs(a).sum(1) > 0)[0].astype(np.int16)
    return idx, a[idx]


def main(out):
    A = os.path.join(REF, "assets")
    verts, uvs, faces, uvfaces = parse_obj(os.path.join(A, "head_template.obj"))
    masks = pickle.load(open(os.path.join(A, "FLAME_masks/FLAME_masks.pkl"), "rb"), encoding="latin1")
    emb = np.load(os.path.join(A, "landmark_embedding.npy"), allow_pickle=True, encoding="latin1")[()]
    mp = np.load(os.path.join(A, "mediapipe_landmark_embedding/mediapipe_landmark_embedding.npz"))
    li, lv = sparse_rows(np.load(os.path.join(A, "l_eyelid.npy")))
    ri, rv = sparse_rows(np.load(os.path.join(A, "r_eyelid.npy")))
    t = lambda x: x.numpy() if hasattr(x, "numpy") else np.asarray(x)
    blob = dict(
        verts=verts, uvs=uvs, faces=faces, uvfaces=uvfaces,
        face_mask=np.asarray(masks["face"]).astype(np.int16),
        l_eyelid_idx=li, l_eyelid_val=lv, r_eyelid_idx=ri, r_eyelid_val=rv,
        static_lmk_faces_idx=t(emb["static_lmk_faces_idx"]).astype(np.int16),
        static_lmk_bary_coords=t(emb["static_lmk_bary_coords"]).astype(np.float64),
        dynamic_lmk_faces_idx=t(emb["dynamic_lmk_faces_idx"]).astype(np.int16),
        dynamic_lmk_bary_coords=t(emb["dynamic_lmk_bary_coords"]).astype(np.float32),
        full_lmk_faces_idx=t(emb["full_lmk_faces_idx"]).astype(np.int16),
        full_lmk_bary_coords=t(emb["full_lmk_bary_coords"]).astype(np.float32),
        mp_lmk_face_idx=mp["lmk_face_idx"].astype(np.int16),
        mp_lmk_b_coords=mp["lmk_b_coords"].astype(np.float64),
        mp_landmark_indices=mp["landmark_indices"].astype(np.int16),
    )
    assert verts.shape == (5023, 3) and faces.shape == (9976, 3) and uvs.shape == (5118, 2)
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else
         os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "flame_topology.npz"))
