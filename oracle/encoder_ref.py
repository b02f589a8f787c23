"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU fp32 restatement of ``SmirkEncoder`` (src/smirk_encoder.py:14-133).  The three backbones come
from timm (``timm==0.9.16``, requirements.txt:10; ``create_model(name, features_only=True)`` at
smirk_encoder.py:7-12) which is NOT vendored in /root/reference and not installed here, so the two
architectures ``tf_mobilenetv3_small_minimal_100`` / ``tf_mobilenetv3_large_minimal_100`` are restated
from timm's published definitions (MobileNetV3 "minimal": ReLU everywhere, no squeeze-excite, 3x3
depthwise only, BN eps 1e-3, TF-"SAME" padding, features_only => stops after the final 1x1 ``cn``
stage).  PARITY UNPINNED for the backbones (Tier B); the heads/clamps (:34-45,66-73,95-110) are
checked against the reference classes through a ``timm`` stub (``oracle/ref_harness.py``).

Module / parameter names follow timm's so that ``state_dict`` keys match a real checkpoint:
``conv_stem, bn1, blocks.<stage>.<i>.{conv_dw,bn1,conv_pw,bn2 | conv_pw,bn1,conv_dw,bn2,conv_pwl,bn3 | conv,bn1}``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-3

# (kind, stride, expansion, out_channels) per block, grouped per stage.
ARCH = {
    "tf_mobilenetv3_large_minimal_100": [
        [("ds", 1, 1.0, 16)],
        [("ir", 2, 4.0, 24), ("ir", 1, 3.0, 24)],
        [("ir", 2, 3.0, 40), ("ir", 1, 3.0, 40), ("ir", 1, 3.0, 40)],
        [("ir", 2, 6.0, 80), ("ir", 1, 2.5, 80), ("ir", 1, 2.3, 80), ("ir", 1, 2.3, 80)],
        [("ir", 1, 6.0, 112), ("ir", 1, 6.0, 112)],
        [("ir", 2, 6.0, 160), ("ir", 1, 6.0, 160), ("ir", 1, 6.0, 160)],
        [("cn", 1, 1.0, 960)],
    ],
    "tf_mobilenetv3_small_minimal_100": [
        [("ds", 2, 1.0, 16)],
        [("ir", 2, 4.5, 24), ("ir", 1, 3.67, 24)],
        [("ir", 2, 4.0, 40), ("ir", 1, 6.0, 40), ("ir", 1, 6.0, 40)],
        [("ir", 1, 3.0, 48), ("ir", 1, 3.0, 48)],
        [("ir", 2, 6.0, 96), ("ir", 1, 6.0, 96), ("ir", 1, 6.0, 96)],
        [("cn", 1, 1.0, 576)],
    ],
}


def make_divisible(v, divisor=8, min_value=None, round_limit=0.9):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def conv_same(x, w, stride, groups=1):
    """TF 'SAME': symmetric for stride 1; for stride 2 the extra pixel goes bottom/right."""
    k = w.shape[-1]
    if k == 1:
        return F.conv2d(x, w, stride=stride, groups=groups)
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / stride) - 1) * stride + (k - 1) + 1 - ih, 0)
    pw = max((math.ceil(iw / stride) - 1) * stride + (k - 1) + 1 - iw, 0)
    x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
    return F.conv2d(x, w, stride=stride, groups=groups)


class _BN(nn.BatchNorm2d):
    def __init__(self, c):
        super().__init__(c, eps=BN_EPS)


class _Conv(nn.Conv2d):
    def __init__(self, cin, cout, k, stride=1, groups=1):
        super().__init__(cin, cout, k, stride=stride, groups=groups, bias=False)

    def forward(self, x):
        return conv_same(x, self.weight, self.stride[0], self.groups)


class DS(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv_dw, self.bn1 = _Conv(cin, cin, 3, stride, groups=cin), _BN(cin)
        self.conv_pw, self.bn2 = _Conv(cin, cout, 1), _BN(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = F.relu(self.bn1(self.conv_dw(x)))
        y = self.bn2(self.conv_pw(y))
        return x + y if self.skip else y


class IR(nn.Module):
    def __init__(self, cin, cout, stride, exp):
        super().__init__()
        mid = make_divisible(cin * exp)
        self.conv_pw, self.bn1 = _Conv(cin, mid, 1), _BN(mid)
        self.conv_dw, self.bn2 = _Conv(mid, mid, 3, stride, groups=mid), _BN(mid)
        self.conv_pwl, self.bn3 = _Conv(mid, cout, 1), _BN(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = F.relu(self.bn1(self.conv_pw(x)))
        y = F.relu(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return x + y if self.skip else y


class CN(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv, self.bn1 = _Conv(cin, cout, 1), _BN(cout)

    def forward(self, x):
        return F.relu(self.bn1(self.conv(x)))


class BackboneRef(nn.Module):
    """Returns a 1-element list holding the last feature map, so ``backbone(img)[-1]`` works as
    the reference uses it (smirk_encoder.py:35,67,96)."""

    def __init__(self, name):
        super().__init__()
        self.conv_stem, self.bn1 = _Conv(3, 16, 3, 2), _BN(16)
        stages, cin = [], 16
        for stage in ARCH[name]:
            blocks = []
            for kind, s, e, c in stage:
                blocks.append(DS(cin, c, s) if kind == "ds" else IR(cin, c, s, e) if kind == "ir" else CN(cin, c))
                cin = c
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.num_chs = cin

    def forward(self, x):
        return [self.blocks(F.relu(self.bn1(self.conv_stem(x))))]


def encoder_forward_ref(sd, img, n_exp=50):
    """Functional forward from a reference-format SmirkEncoder ``state_dict`` (smirk_encoder.py:123-133)."""
    out = {}
    feats = {}
    for enc, name in (("pose_encoder", "tf_mobilenetv3_small_minimal_100"),
                      ("shape_encoder", "tf_mobilenetv3_large_minimal_100"),
                      ("expression_encoder", "tf_mobilenetv3_large_minimal_100")):
        bb = BackboneRef(name).eval()
        bb.load_state_dict({k[len(enc) + 9:]: v.detach().float().cpu() for k, v in sd.items()
                            if k.startswith(enc + ".encoder.")})
        with torch.no_grad():
            f = bb(img)[-1]
        feats[enc] = F.adaptive_avg_pool2d(f, (1, 1)).squeeze(-1).squeeze(-1)
    g = lambda k: sd[k].detach().float().cpu()
    pc = F.linear(feats["pose_encoder"], g("pose_encoder.pose_cam_layers.0.weight"), g("pose_encoder.pose_cam_layers.0.bias"))
    out["pose_params"], out["cam"] = pc[..., :3], pc[..., 3:]
    out["shape_params"] = F.linear(feats["shape_encoder"], g("shape_encoder.shape_layers.0.weight"),
                                   g("shape_encoder.shape_layers.0.bias"))
    p = F.linear(feats["expression_encoder"], g("expression_encoder.expression_layers.0.weight"),
                 g("expression_encoder.expression_layers.0.bias"))
    out["expression_params"] = p[..., :n_exp]
    out["eyelid_params"] = torch.clamp(p[..., n_exp:n_exp + 2], 0, 1)
    out["jaw_params"] = torch.cat([F.relu(p[..., n_exp + 2].unsqueeze(-1)),
                                   torch.clamp(p[..., n_exp + 3:n_exp + 5], -.2, .2)], -1)
    out["_features"] = feats
    return out
