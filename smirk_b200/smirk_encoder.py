"""``SmirkEncoder`` — drop-in for the reference ``src/smirk_encoder.py`` (forward only).

Same class names, constructor arguments, sub-module / parameter names (``state_dict`` keys follow the
timm ``features_only`` MobileNetV3 layout the reference checkpoints use) and output dicts.  The
``nn.Conv2d`` / ``nn.BatchNorm2d`` objects below are parameter containers only — they are never
called; the forward pass runs in ``csrc/encoder.cu`` through ``smk_encoder_forward``.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-3          # timm tf_* models (BN_EPS_TF_DEFAULT)

# (kind, stride, expansion, out_channels) per block, grouped per stage — timm 0.9.16
# tf_mobilenetv3_{large,small}_minimal_100 with features_only=True (stops after the `cn` stage).
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


def _make_divisible(v, divisor=8, round_limit=0.9):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < round_limit * v:
        new_v += divisor
    return new_v


def _conv(cin, cout, k, stride=1, groups=1):
    return nn.Conv2d(cin, cout, k, stride=stride, groups=groups, bias=False)


def _bn(c):
    return nn.BatchNorm2d(c, eps=BN_EPS)


class _Block(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("smirk_b200: backbone blocks are parameter containers; call SmirkEncoder.forward")


class _DS(_Block):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv_dw, self.bn1 = _conv(cin, cin, 3, stride, cin), _bn(cin)
        self.conv_pw, self.bn2 = _conv(cin, cout, 1), _bn(cout)


class _IR(_Block):
    def __init__(self, cin, cout, stride, exp):
        super().__init__()
        mid = _make_divisible(cin * exp)
        self.conv_pw, self.bn1 = _conv(cin, mid, 1), _bn(mid)
        self.conv_dw, self.bn2 = _conv(mid, mid, 3, stride, mid), _bn(mid)
        self.conv_pwl, self.bn3 = _conv(mid, cout, 1), _bn(cout)


class _CN(_Block):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv, self.bn1 = _conv(cin, cout, 1), _bn(cout)


class _Backbone(_Block):
    """Parameter tree of a timm MobileNetV3Features model (conv_stem, bn1, blocks.<stage>.<i>...)."""

    def __init__(self, name):
        super().__init__()
        self.conv_stem, self.bn1 = _conv(3, 16, 3, 2), _bn(16)
        stages, cin = [], 16
        for stage in ARCH[name]:
            blocks = []
            for kind, s, e, c in stage:
                blocks.append(_DS(cin, c, s) if kind == "ds" else _IR(cin, c, s, e) if kind == "ir" else _CN(cin, c))
                cin = c
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.feature_dim = cin

    def tensor_list(self):
        """fp32 tensors in state_dict order without num_batches_tracked (the C ABI's contract)."""
        return [v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")]


def create_backbone(backbone_name, pretrained=True):
    """Signature of smirk_encoder.py:7-12.  ``pretrained`` is accepted and ignored: there is no network
    access here and the reference always overwrites the weights from its checkpoint (demo.py:55-58)."""
    bb = _Backbone(backbone_name)
    return bb, bb.feature_dim


class PoseEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder, feature_dim = create_backbone("tf_mobilenetv3_small_minimal_100")
        self.pose_cam_layers = nn.Sequential(nn.Linear(feature_dim, 6))
        self.init_weights()

    def init_weights(self):                     # smirk_encoder.py:26-31
        self.pose_cam_layers[-1].weight.data *= 0.001
        self.pose_cam_layers[-1].bias.data *= 0.001
        self.pose_cam_layers[-1].weight.data[3] = 0
        self.pose_cam_layers[-1].bias.data[3] = 7


class ShapeEncoder(nn.Module):
    def __init__(self, n_shape=300):
        super().__init__()
        self.encoder, feature_dim = create_backbone("tf_mobilenetv3_large_minimal_100")
        self.shape_layers = nn.Sequential(nn.Linear(feature_dim, n_shape))
        self.init_weights()

    def init_weights(self):                     # smirk_encoder.py:61-63
        self.shape_layers[-1].weight.data *= 0
        self.shape_layers[-1].bias.data *= 0


class ExpressionEncoder(nn.Module):
    def __init__(self, n_exp=50):
        super().__init__()
        self.encoder, feature_dim = create_backbone("tf_mobilenetv3_large_minimal_100")
        self.expression_layers = nn.Sequential(nn.Linear(feature_dim, n_exp + 2 + 3))
        self.n_exp = n_exp
        self.init_weights()

    def init_weights(self):                     # smirk_encoder.py:90-92
        self.expression_layers[-1].weight.data *= 0.1
        self.expression_layers[-1].bias.data *= 0.1


class SmirkEncoder(nn.Module):
    def __init__(self, n_exp=50, n_shape=300):
        super().__init__()
        self.pose_encoder = PoseEncoder()
        self.shape_encoder = ShapeEncoder(n_shape=n_shape)
        self.expression_encoder = ExpressionEncoder(n_exp=n_exp)
        self.n_exp, self.n_shape = n_exp, n_shape
        self.precision = 0
        self._handle, self._sig, self._ws = None, None, _lib.Workspace()

    # -- native handle (re-packed whenever a parameter / buffer is modified or moved) ----------------
    def _signature(self, device):
        s = [str(device), self.precision]
        for t in list(self.parameters()) + list(self.buffers()):
            s.append(t._version)
            s.append(t.data_ptr())
        return tuple(s)

    def _native(self, device):
        sig = self._signature(device)
        if self._handle is not None and self._sig == sig:
            return self._handle
        self._release()
        if self.training:
            raise RuntimeError("smirk_b200.SmirkEncoder: train-mode BatchNorm is not implemented (forward/eval only)")
        L = _lib.lib()
        keep = []
        d = _lib.SmkEncoderDesc()
        heads = (self.pose_encoder.pose_cam_layers[0], self.shape_encoder.shape_layers[0],
                 self.expression_encoder.expression_layers[0])
        for i, enc in enumerate((self.pose_encoder, self.shape_encoder, self.expression_encoder)):
            ts = enc.encoder.tensor_list()
            arr = (_lib.c_f32p * len(ts))()
            for j, t in enumerate(ts):
                a, p = _lib.f32(t)
                keep.append(a)
                arr[j] = p
            keep.append(arr)
            d.tensors[i] = C.cast(arr, C.POINTER(_lib.c_f32p))
            d.n_tensors[i] = len(ts)
            a, p = _lib.f32(heads[i].weight); keep.append(a); d.head_w[i] = p
            a, p = _lib.f32(heads[i].bias); keep.append(a); d.head_b[i] = p
        d.n_shape, d.n_exp, d.precision = self.n_shape, self.n_exp, self.precision
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.smk_encoder_create(C.byref(d), C.byref(h)), "smk_encoder_create")
        self._handle, self._sig = h, sig
        return h

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.lib().smk_encoder_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._release()

    def __deepcopy__(self, memo):               # base_trainer.py:237 deep-copies the encoder
        import copy
        new = self.__class__.__new__(self.__class__)
        nn.Module.__init__(new)
        for k, v in self.__dict__.items():
            if k not in ("_handle", "_sig", "_ws"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._handle, new._sig, new._ws = None, None, _lib.Workspace()
        return new

    @torch.no_grad()
    def forward(self, img):
        _lib.require_cuda(img, "img")
        dev = img.device
        L = _lib.lib()
        h = self._native(dev)
        x = _lib.dev_f32(img, "img")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise RuntimeError("smirk_b200.SmirkEncoder: expected img [B,3,224,224], got %s" % (tuple(x.shape),))
        B, ne = x.shape[0], self.n_exp
        o = lambda n: torch.empty(B, n, dtype=torch.float32, device=dev)
        pose_cam, shape, expr = o(6), o(self.n_shape), o(ne + 5)
        with torch.cuda.device(dev):
            ws = self._ws.get(L.smk_encoder_workspace_bytes(h, B), dev)
            _lib.check(L.smk_encoder_forward(h, _lib.ptr(x), B, _lib.ptr(pose_cam), _lib.ptr(shape), _lib.ptr(expr),
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "smk_encoder_forward")
        return {
            "pose_params": pose_cam[..., :3], "cam": pose_cam[..., 3:],
            "shape_params": shape,
            "expression_params": expr[..., :ne], "eyelid_params": expr[..., ne:ne + 2],
            "jaw_params": expr[..., ne + 2:ne + 5],
        }
