"""TEST INFRASTRUCTURE ONLY — end-to-end parity report of the device pipeline against the CPU oracle.

Used by tests/test_gpu_parity.py, ``__graft_entry__.smoke()`` and bench.py's in-run ``parity`` block (the checker,
never the thing measured).  Given one batch of images and what the device produced for it, it restates the path on
the CPU (oracle/encoder_ref -> flame_ref -> render_ref, i.e. the reference's own arithmetic: torch fp32 ops and the
C restatement of pytorch3d's rasteriser) and reports, in the north_star's terms:

  params_rel         max |regressed parameter - oracle| / max(|oracle|, 1), worst of the six heads
  vertices_rel       max |v - v_oracle| / max |v_oracle|                      (north_star: <= 1e-4)
  tverts_rel         the same for Renderer's transformed_vertices
  pixels_stage_abs   max |rendered - oracle rendered| with the ORACLE rasterising the DEVICE's own vertices / cam
                     (isolates the rasteriser + shading: north_star <= 1e-4; rendered values lie in [0, 1.7])
  p2f_stage_mismatch pixels whose face index differs, same setting (north_star: bit-exact -> 0)
  pixels_e2e_abs_p999 / p2f_e2e_mismatch_frac
                     image -> pixels through the whole chain on both sides.  Face indices and silhouette pixels are a
                     discontinuous function of the vertices, so ANY two fp32 evaluation orders (the reference's CPU
                     and GPU builds included) disagree on a few edge pixels; reported, not asserted.
"""
import torch

from . import encoder_ref, flame_ref, render_ref

_CONST = {}


def _constants(root):
    if root not in _CONST:
        _CONST[root] = (flame_ref.FlameConstants(root), render_ref.RenderConstants(root))
    return _CONST[root]


def pipeline_report(root, enc_state_dict, img_cpu, dev_out):
    """dev_out: dict of device (or CPU) tensors with keys params_dict (the encoder's dict), vertices, rendered_img,
    transformed_vertices and optionally pix_to_face (int64 [B,S,S])."""
    fc, rc = _constants(root)
    cpu = lambda t: t.detach().float().cpu()
    sd = {k: v.detach().cpu() for k, v in enc_state_dict.items()}
    with torch.no_grad():
        pe = encoder_ref.encoder_forward_ref(sd, img_cpu)
        pe = {k: v for k, v in pe.items() if not k.startswith("_")}
        fo = flame_ref.flame_forward_ref(fc, pe)
        ro = render_ref.render_forward_ref(rc, fo["vertices"], pe["cam"])
        pd = {k: cpu(v) for k, v in dev_out["params_dict"].items()}
        v_dev = cpu(dev_out["vertices"])
        ro_stage = render_ref.render_forward_ref(rc, v_dev, pd["cam"])
    rep = {}
    rep["params_rel"] = max(float((pd[k] - pe[k]).abs().max()) / max(float(pe[k].abs().max()), 1.0) for k in pe)
    rep["vertices_rel"] = float((v_dev - fo["vertices"]).abs().max()) / float(fo["vertices"].abs().max())
    tv = cpu(dev_out["transformed_vertices"])
    rep["tverts_rel"] = float((tv - ro["transformed_vertices"]).abs().max()) / float(ro["transformed_vertices"].abs().max())
    img_dev = cpu(dev_out["rendered_img"])
    rep["pixels_stage_abs"] = float((img_dev - ro_stage["rendered_img"]).abs().max())
    d = (img_dev - ro["rendered_img"]).abs().flatten()
    rep["pixels_e2e_abs_p999"] = float(torch.quantile(d[:: max(1, d.numel() // 2_000_000)], 0.999))
    rep["pixels_e2e_mean_abs"] = float(d.mean())
    if dev_out.get("pix_to_face") is not None:
        p2f = dev_out["pix_to_face"].detach().cpu()
        rep["p2f_stage_mismatch"] = int((p2f != ro_stage["pix_to_face"]).sum())
        rep["p2f_e2e_mismatch_frac"] = float((p2f != ro["pix_to_face"]).float().mean())
    rep["faces"] = int(img_cpu.shape[0])
    rep["coverage"] = float((ro["pix_to_face"] >= 0).float().mean())
    return rep
