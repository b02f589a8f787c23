"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU fp32 restatement of the reference FLAME forward: ``src/FLAME/FLAME.py:232-315`` and the
functions of ``src/FLAME/lbs.py`` it reaches.  Tier A: checked against the reference's own classes
by ``oracle/make_golden.py`` / ``tests/test_oracle_vs_golden.py``.
"""
import math
import os
import pickle

import numpy as np
import torch


class FlameConstants:
    """Buffers the reference builds in ``FLAME.__init__`` (FLAME.py:50-113), same names/shapes."""

    def __init__(self, root=".", n_shape=300, n_exp=50):
        A = os.path.join(root, "assets")
        with open(os.path.join(A, "FLAME2020", "generic_model.pkl"), "rb") as fh:
            m = pickle.load(fh, encoding="latin1")
        f32 = lambda a: torch.from_numpy(np.array(a, dtype=np.float32))
        self.faces_tensor = torch.from_numpy(np.array(m["f"], dtype=np.int64))            # FLAME.py:61
        self.v_template = f32(m["v_template"])                                             # :64
        sd = f32(m["shapedirs"])
        self.shapedirs = torch.cat([sd[:, :, :n_shape], sd[:, :, 300:300 + n_exp]], 2)     # :67-69
        pd = np.asarray(m["posedirs"])
        self.posedirs = f32(np.reshape(pd, [-1, pd.shape[-1]]).T)                          # :71-73
        self.J_regressor = f32(m["J_regressor"])                                           # :75
        parents = torch.from_numpy(np.array(m["kintree_table"][0], dtype=np.float32)).long()
        parents[0] = -1                                                                    # :76
        self.parents = parents
        self.lbs_weights = f32(m["weights"])                                               # :78
        self.l_eyelid = torch.from_numpy(np.load(os.path.join(A, "l_eyelid.npy"))).float()[None]   # :81
        self.r_eyelid = torch.from_numpy(np.load(os.path.join(A, "r_eyelid.npy"))).float()[None]   # :82
        e = np.load(os.path.join(A, "landmark_embedding.npy"), allow_pickle=True, encoding="latin1")[()]
        tt = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))
        self.lmk_faces_idx = tt(e["static_lmk_faces_idx"]).long()                          # :96
        self.lmk_bary_coords = tt(e["static_lmk_bary_coords"]).float()
        self.dynamic_lmk_faces_idx = tt(e["dynamic_lmk_faces_idx"]).long()
        self.dynamic_lmk_bary_coords = tt(e["dynamic_lmk_bary_coords"]).float()
        self.full_lmk_faces_idx = tt(e["full_lmk_faces_idx"]).long()
        self.full_lmk_bary_coords = tt(e["full_lmk_bary_coords"]).float()
        self.neck_kin_chain = torch.tensor([1, 0])                                         # :103-108
        mp = np.load(os.path.join(A, "mediapipe_landmark_embedding", "mediapipe_landmark_embedding.npz"))
        self.mp_lmk_faces_idx = torch.from_numpy(mp["lmk_face_idx"].astype("int32")).long()   # :112
        self.mp_lmk_bary_coords = torch.from_numpy(mp["lmk_b_coords"]).float()
        self.n_shape, self.n_exp = n_shape, n_exp


def rodrigues(r):
    """lbs.py:274-305 — note the 1e-8 is added to the vector before the norm only."""
    n = r.shape[0]
    angle = torch.norm(r + 1e-8, dim=1, keepdim=True)
    d = r / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], 1).view(n, 3, 3)
    return torch.eye(3)[None] + s * K + (1 - c) * torch.bmm(K, K)


def rigid_chain(R, J, parents):
    """lbs.py:321-378: world transforms along the kinematic tree, then remove the rest pose."""
    B, nj = J.shape[:2]
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    T = torch.zeros(B, nj, 4, 4)
    T[:, :, :3, :3] = R
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], T[:, i]))
    G = torch.stack(chain, 1)
    Jh = torch.cat([J, torch.zeros(B, nj, 1)], 2)[..., None]               # [B,nj,4,1]
    corr = torch.matmul(G, Jh)                                             # [B,nj,4,1]
    A = G.clone()
    A[..., 3:4] = G[..., 3:4] - corr
    return G[:, :, :3, 3], A


def lbs_ref(betas, pose, c):
    """lbs.py:140-227."""
    B = betas.shape[0]
    v_shaped = c.v_template[None] + torch.einsum("bl,mkl->bmk", betas, c.shapedirs)      # :184
    J = torch.einsum("bik,ji->bjk", v_shaped, c.J_regressor)                             # :188
    R = rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)                                 # :194
    pf = (R[:, 1:] - torch.eye(3)).reshape(B, -1)                                        # :197
    v_posed = torch.matmul(pf, c.posedirs).view(B, -1, 3) + v_shaped                     # :199-208
    Jt, A = rigid_chain(R, J, c.parents)                                                 # :210
    T = torch.matmul(c.lbs_weights[None].expand(B, -1, -1), A.view(B, -1, 16)).view(B, -1, 4, 4)   # :217
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1)], 2)
    v = torch.matmul(T, vh[..., None])[:, :, :3, 0]                                      # :223-225
    return v, Jt


def landmarks_ref(verts, faces, fidx, bary):
    """lbs.py:101-137: fidx [B,L] long, bary [B,L,3]."""
    B, V = verts.shape[:2]
    tri = faces[fidx.reshape(-1)].view(B, -1, 3) + (torch.arange(B) * V).view(-1, 1, 1)
    pts = verts.reshape(-1, 3)[tri].view(B, -1, 3, 3)
    return torch.einsum("blfi,blf->bli", pts, bary)


def dynamic_contour_ref(full_pose, c):
    """FLAME.py:117-159 (the method FLAME.forward uses; NO negation of the yaw, unlike lbs.py:85)."""
    B = full_pose.shape[0]
    aa = full_pose.view(B, -1, 3)[:, c.neck_kin_chain]
    R = rodrigues(aa.reshape(-1, 3)).view(B, -1, 3, 3)
    rel = torch.eye(3)[None].expand(B, -1, -1)
    for i in range(len(c.neck_kin_chain)):
        rel = torch.bmm(R[:, i], rel)
    sy = torch.sqrt(rel[:, 0, 0] * rel[:, 0, 0] + rel[:, 1, 0] * rel[:, 1, 0])       # lbs.py:26-32
    yaw = torch.atan2(-rel[:, 2, 0], sy)
    y = torch.round(torch.clamp(yaw * 180.0 / np.pi, max=39)).long()
    neg = y.lt(0).long()
    m = y.lt(-39).long()
    y = neg * (m * 78 + (1 - m) * (39 - y)) + (1 - neg) * y
    return c.dynamic_lmk_faces_idx[y], c.dynamic_lmk_bary_coords[y], y


def flame_forward_ref(c, params, zero_expression=False, zero_shape=False, zero_pose=False):
    """FLAME.forward (FLAME.py:232-315).  ``params`` is the reference's param_dictionary."""
    shape = params["shape_params"].float()
    expr = params["expression_params"].float()
    pose = params.get("pose_params")
    jaw = params.get("jaw_params")
    eye = params.get("eye_pose_params")
    neck = params.get("neck_pose_params")
    eyelid = params.get("eyelid_params")
    B = shape.shape[0]
    if expr.shape[1] < c.n_exp:
        expr = torch.cat([expr, torch.zeros(B, c.n_exp - expr.shape[1])], 1)
    if shape.shape[1] < c.n_shape:
        shape = torch.cat([shape, torch.zeros(B, c.n_shape - shape.shape[1])], 1)
    if zero_expression:
        expr, jaw = torch.zeros_like(expr), torch.zeros_like(jaw)
    if zero_shape:
        shape = torch.zeros_like(shape)
    if zero_pose:
        pose = torch.zeros_like(pose)
        pose[..., 0], pose[..., 1] = 0.2, -0.7
    eye = torch.zeros(B, 6) if eye is None else eye
    neck = torch.zeros(B, 3) if neck is None else neck
    betas = torch.cat([shape, expr], 1)
    full_pose = torch.cat([pose, neck, jaw, eye], 1)
    v, _ = lbs_ref(betas, full_pose, c)
    if eyelid is not None:                                                   # FLAME.py:284-286
        v = v + c.r_eyelid.expand(B, -1, -1) * eyelid[:, 1:2, None]
        v = v + c.l_eyelid.expand(B, -1, -1) * eyelid[:, 0:1, None]
    dfi, dbc, yaw_idx = dynamic_contour_ref(full_pose, c)
    fi = torch.cat([dfi, c.lmk_faces_idx[None].expand(B, -1)], 1)
    bc = torch.cat([dbc, c.lmk_bary_coords[None].expand(B, -1, -1)], 1)
    out = {
        "vertices": v,
        "landmarks_fan": landmarks_ref(v, c.faces_tensor, fi, bc),
        "landmarks_fan_3d": landmarks_ref(v, c.faces_tensor, c.full_lmk_faces_idx.repeat(B, 1),
                                          c.full_lmk_bary_coords.repeat(B, 1, 1)),
        "landmarks_mp": landmarks_ref(v, c.faces_tensor, c.mp_lmk_faces_idx.repeat(B, 1),
                                      c.mp_lmk_bary_coords.repeat(B, 1, 1)),
        "_dyn_idx": yaw_idx,
    }
    return out
