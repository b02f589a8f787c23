"""Seeded synthetic inputs and weights.

No dataset, licensed FLAME model or pretrained checkpoint is reachable from this build, so tests,
``bench.py``, ``smoke()`` and the golden-vector generator all draw inputs and weights from the recipes
below.  Everything is generated on the CPU with an explicit ``torch.Generator`` so the same seed gives
the same bytes in the build container and on the GPU box.
"""
import math

import torch


def flame_params(B, seed):
    """shape/expr ~ N(0,1); pose ~ N(0,0.2^2); jaw/eyelid inside the encoder's clamps
    (smirk_encoder.py:105-108); cam = (s~U(6,9), tx,ty~U(-.03,.03)) (scale 7 is the encoder's init,
    smirk_encoder.py:30-31)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    n = lambda *s: torch.randn(*s, generator=g)
    return {
        "shape_params": n(B, 300), "expression_params": n(B, 50), "pose_params": n(B, 3) * 0.2,
        "jaw_params": torch.cat([r(B, 1) * 0.3, r(B, 2) * 0.4 - 0.2], 1),
        "eyelid_params": r(B, 2),
        "cam": torch.cat([r(B, 1) * 3 + 6, r(B, 2) * 0.06 - 0.03], 1),
    }


def images(B, seed):
    """RGB NCHW fp32 in [0,1) (demo.py:102-105)."""
    return torch.rand(B, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


def masked_images(B, seed):
    """Stand-in for masking() output (demo.py:165): sparse random pixels, ~30 % kept."""
    g = torch.Generator().manual_seed(seed)
    m = (torch.rand(B, 1, 224, 224, generator=g) < 0.3).float()
    return torch.rand(B, 3, 224, 224, generator=g) * m


def hull_masks(B, seed):
    """Stand-in for ``create_mask(landmarks)`` (datasets/base_dataset.py:9-15, demo.py:142): [B,1,224,224], 0 inside a
    face-sized ellipse (the landmark hull), 1 outside."""
    g = torch.Generator().manual_seed(seed)
    cx, cy = 112 + (torch.rand(B, generator=g) - 0.5) * 16, 116 + (torch.rand(B, generator=g) - 0.5) * 16
    ax, ay = 58 + torch.rand(B, generator=g) * 14, 76 + torch.rand(B, generator=g) * 14
    y, x = torch.meshgrid(torch.arange(224.0), torch.arange(224.0), indexing="ij")
    inside = ((x[None] - cx[:, None, None]) / ax[:, None, None]) ** 2 + ((y[None] - cy[:, None, None]) / ay[:, None, None]) ** 2 < 1
    return (~inside).float()[:, None].contiguous()


def face_probabilities(n_faces, seed=11):
    """Stand-in for ``load_probabilities_per_FLAME_triangle()`` (masking.py:11-38; its asset is not shipped): per-triangle
    base sampling weights in {0, 0.5, 1} like the reference's area weights."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n_faces, generator=g) > 0.6).float() * torch.tensor([0.5, 1.0])[torch.randint(0, 2, (n_faces,), generator=g)]


def random_state_dict(template, seed=7):
    """Fill a ``state_dict``-shaped template (name -> tensor) deterministically, by key name, so the
    reference modules, the oracle and the product modules get identical weights regardless of their
    construction order.  Convs / linears: U(-b, b), b = sqrt(3 / fan_in) * gain; BN: gamma~U(.5,1.5),
    beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5)."""
    out = {}
    for k in sorted(template.keys()):
        t = template[k]
        g = torch.Generator().manual_seed((seed * 1000003 + sum(ord(c) * (i + 1) for i, c in enumerate(k))) % (2 ** 31))
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            out[k] = torch.rand(t.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            out[k] = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1 and (".bn" in k or "norm" in k or ".conv_block.2." in k or ".conv_block.6." in k):
            out[k] = (torch.rand(t.shape, generator=g) + 0.5) if k.endswith("weight") else torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() >= 2:
            fan_in = t[0].numel() if not k.startswith("upconv") else t.shape[0]
            linear = any(s in k for s in ("conv_pwl", "conv_block.5", "upconv", "blocks.0.0.conv_pw", "_layers.")) \
                or k == "conv.weight"
            b = math.sqrt(3.0 / fan_in) * (1.0 if linear else math.sqrt(2.0))
            if "conv_pwl" in k or "conv_block.5" in k:
                b *= 0.5                         # residual branches: keep x + f(x) from doubling the variance
            w = (torch.rand(t.shape, generator=g) * 2 - 1) * b
            if "pose_cam_layers" in k:
                w = w * 0.1                      # keeps cam scale near its init of 7 (smirk_encoder.py:26-31)
            out[k] = w
        elif "pose_cam_layers" in k:             # bias: (pose 3, cam scale, tx, ty)
            out[k] = torch.tensor([0.05, -0.1, 0.02, 7.5, 0.01, -0.01])
        else:                                   # conv / linear bias
            out[k] = torch.randn(t.shape, generator=g) * 0.05
    return out
