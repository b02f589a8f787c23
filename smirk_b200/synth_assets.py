"""Materialise an ``assets/`` tree that the reference constructors (and ours) can read.

The reference reads its constants from hard-coded relative paths (``FLAME.py:54,81-82,94,111``,
``renderer.py:54,66-68``).  Two things are missing on a GPU box: the reference checkout itself
and the licensed ``assets/FLAME2020/generic_model.pkl``.  ``materialize()`` rebuilds the former
from the compact derived blob ``tests/golden/flame_topology.npz`` (see
``oracle/make_topology_blob.py``) and *synthesises* the latter: a smooth, deterministic FLAME-like
model (same keys, shapes and dtypes as the real pickle: ``f, v_template, shapedirs[5023,3,400],
posedirs[5023,3,36], J_regressor[5,5023], kintree_table[2,5], weights[5023,5]``) so every stage
of the hot path runs on realistic geometry.  Used by tests, ``bench.py`` and ``smoke()``.
"""
import os
import pickle

import numpy as np

_DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                             "flame_topology.npz")

N_VERTS, N_FACES, N_JOINTS = 5023, 9976, 5


def load_blob(blob_path=None):
    z = np.load(blob_path or _DEFAULT_BLOB)
    return {k: z[k] for k in z.files}


def synthetic_flame_model(verts, faces, seed=0):
    """A smooth random FLAME-like model on the real topology.

    Blendshape directions are low-frequency sinusoids of the template position so deformed meshes
    stay smooth surfaces (triangle sizes and overlaps stay face-like for the rasteriser); amplitudes
    decay with the component index like a PCA basis.  Skinning weights / joint regressor are
    softmaxes of the distance to five plausible joint centres (root, neck, jaw, two eyes).
    """
    rng = np.random.RandomState(seed)
    v = verts.astype(np.float64)
    v = v - v.mean(0, keepdims=True)               # the OBJ is not origin-centred; FLAME's template is
    nb = 400
    freq = rng.uniform(4.0, 40.0, size=(nb, 3, 3))  # [component, out-axis, in-axis]
    phase = rng.uniform(0, 2 * np.pi, size=(nb, 3))
    amp = 3e-3 / (1.0 + np.arange(nb) / 12.0)
    amp[300:] = 2.5e-3 / (1.0 + np.arange(100) / 10.0)   # expression block (FLAME.py:68 slices 300:350)
    arg = np.einsum("vi,lki->vkl", v, freq) + phase.T[None]          # [V,3,nb]
    shapedirs = (np.sin(arg) * amp[None, None, :])
    pfreq = rng.uniform(4.0, 30.0, size=(36, 3, 3))
    pphase = rng.uniform(0, 2 * np.pi, size=(36, 3))
    posedirs = np.sin(np.einsum("vi,lki->vkl", v, pfreq) + pphase.T[None]) * 1.5e-3
    lo, hi = v.min(0), v.max(0)
    ext = hi - lo
    centres = np.array([
        [0.0, lo[1] + 0.10 * ext[1], lo[2] + 0.35 * ext[2]],     # root
        [0.0, lo[1] + 0.22 * ext[1], lo[2] + 0.40 * ext[2]],     # neck
        [0.0, lo[1] + 0.38 * ext[1], lo[2] + 0.55 * ext[2]],     # jaw
        [-0.16 * ext[0], lo[1] + 0.66 * ext[1], lo[2] + 0.80 * ext[2]],   # eyes
        [+0.16 * ext[0], lo[1] + 0.66 * ext[1], lo[2] + 0.80 * ext[2]],
    ])
    d2 = ((v[:, None, :] - centres[None]) ** 2).sum(-1)              # [V,5]
    w = np.exp(-d2 / (0.05 ** 2) * np.array([0.6, 1.0, 1.6, 6.0, 6.0])[None])
    w[:, 1] += 0.05
    weights = w / w.sum(1, keepdims=True)
    jr = np.exp(-d2.T / (0.03 ** 2))                                  # [5,V]
    jr = jr / jr.sum(1, keepdims=True)
    kintree = np.array([[2 ** 32 - 1, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=np.int64)
    return dict(
        f=faces.astype(np.uint32), v_template=v, shapedirs=shapedirs, posedirs=posedirs,
        J_regressor=jr, kintree_table=kintree, weights=weights, bs_type="lrotmin", bs_style="lbs",
    )


def _dense(idx, val, n):
    a = np.zeros((n, 3), np.float64)
    a[idx.astype(np.int64)] = val.astype(np.float64)
    return a


def write_obj(path, verts, uvs, faces, uvfaces):
    with open(path, "w") as fh:
        fh.write("# smirk-b200 materialised head topology\n")
        for x, y, z in verts:
            fh.write("v %.6f %.6f %.6f\n" % (x, y, z))
        for u, w in uvs:
            fh.write("vt %.6f %.6f\n" % (u, w))
        fh.write("s off\n")
        for (a, b, c), (ta, tb, tc) in zip(faces.astype(np.int64) + 1, uvfaces.astype(np.int64) + 1):
            fh.write("f %d/%d %d/%d %d/%d\n" % (a, ta, b, tb, c, tc))


def materialize(root, blob_path=None, seed=0, force=False):
    """Write ``<root>/assets/...``; returns ``root``.  Idempotent unless ``force``."""
    import torch
    A = os.path.join(root, "assets")
    stamp = os.path.join(A, ".materialized_seed_%d" % seed)
    if os.path.exists(stamp) and not force:
        return root
    b = load_blob(blob_path)
    os.makedirs(os.path.join(A, "FLAME2020"), exist_ok=True)
    os.makedirs(os.path.join(A, "FLAME_masks"), exist_ok=True)
    os.makedirs(os.path.join(A, "mediapipe_landmark_embedding"), exist_ok=True)
    write_obj(os.path.join(A, "head_template.obj"), b["verts"], b["uvs"], b["faces"], b["uvfaces"])
    with open(os.path.join(A, "FLAME_masks", "FLAME_masks.pkl"), "wb") as fh:
        pickle.dump({"face": b["face_mask"].astype(np.int64)}, fh, protocol=2)
    np.save(os.path.join(A, "l_eyelid.npy"), _dense(b["l_eyelid_idx"], b["l_eyelid_val"], N_VERTS))
    np.save(os.path.join(A, "r_eyelid.npy"), _dense(b["r_eyelid_idx"], b["r_eyelid_val"], N_VERTS))
    emb = {
        "static_lmk_faces_idx": b["static_lmk_faces_idx"].astype(np.int64),
        "static_lmk_bary_coords": b["static_lmk_bary_coords"].astype(np.float64),
        # the shipped file stores these two as torch tensors (FLAME.py:98-99 call .long()/.to())
        "dynamic_lmk_faces_idx": torch.from_numpy(b["dynamic_lmk_faces_idx"].astype(np.int64)),
        "dynamic_lmk_bary_coords": torch.from_numpy(b["dynamic_lmk_bary_coords"].astype(np.float32)),
        "full_lmk_faces_idx": b["full_lmk_faces_idx"].astype(np.int64),
        "full_lmk_bary_coords": b["full_lmk_bary_coords"].astype(np.float32),
    }
    np.save(os.path.join(A, "landmark_embedding.npy"), np.array(emb, dtype=object), allow_pickle=True)
    np.savez(os.path.join(A, "mediapipe_landmark_embedding", "mediapipe_landmark_embedding.npz"),
             lmk_face_idx=b["mp_lmk_face_idx"].astype(np.uint32),
             lmk_b_coords=b["mp_lmk_b_coords"].astype(np.float64),
             landmark_indices=b["mp_landmark_indices"].astype(np.int64))
    model = synthetic_flame_model(b["verts"], b["faces"].astype(np.int64), seed=seed)
    with open(os.path.join(A, "FLAME2020", "generic_model.pkl"), "wb") as fh:
        pickle.dump(model, fh, protocol=2)
    open(stamp, "w").write("ok\n")
    return root
