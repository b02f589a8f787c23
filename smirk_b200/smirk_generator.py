"""``SmirkGenerator`` — drop-in for the reference ``src/smirk_generator.py`` (forward only, eval BN).

Same constructor, sub-module names (``state_dict`` keys such as ``encoder1.enc1conv1.weight``,
``resnet_blocks.0.conv_block.1.weight``, ``upconv4.bias``) and forward signature.  The torch modules are
parameter containers; the forward pass runs in ``csrc/generator.cu`` through ``smk_generator_forward``.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib


class ResnetBlock(nn.Module):
    """Parameter layout of smirk_generator.py:121-171 (indices 1,2 / 5,6 of ``conv_block``)."""

    def __init__(self, dim, padding_type="reflect", norm_layer=nn.BatchNorm2d, use_dropout=False, use_bias=False):
        super().__init__()
        if padding_type != "reflect" or use_dropout:
            raise NotImplementedError("smirk_b200.ResnetBlock: only reflect padding without dropout (as built at smirk_generator.py:25)")
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3, padding=0, bias=use_bias), norm_layer(dim), nn.ReLU(True),
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3, padding=0, bias=use_bias), norm_layer(dim))


class SmirkGenerator(nn.Module):
    def __init__(self, in_channels=3, out_channels=1, init_features=16, res_blocks=3):
        super().__init__()
        f = init_features
        self.encoder1 = SmirkGenerator._block(in_channels, f, name="enc1")
        self.pool1 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder2 = SmirkGenerator._block(f, f * 2, name="enc2")
        self.pool2 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder3 = SmirkGenerator._block(f * 2, f * 4, name="enc3")
        self.pool3 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.encoder4 = SmirkGenerator._block(f * 4, f * 8, name="enc4")
        self.pool4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.bottleneck = SmirkGenerator._block(f * 8, f * 16, name="bottleneck")
        self.resnet_blocks = nn.ModuleList([ResnetBlock(f * 16) for _ in range(res_blocks)])
        self.upconv4 = nn.ConvTranspose2d(f * 16, f * 8, kernel_size=2, stride=2)
        self.decoder4 = SmirkGenerator._block(f * 16, f * 8, name="dec4")
        self.upconv3 = nn.ConvTranspose2d(f * 8, f * 4, kernel_size=2, stride=2)
        self.decoder3 = SmirkGenerator._block(f * 8, f * 4, name="dec3")
        self.upconv2 = nn.ConvTranspose2d(f * 4, f * 2, kernel_size=2, stride=2)
        self.decoder2 = SmirkGenerator._block(f * 4, f * 2, name="dec2")
        self.upconv1 = nn.ConvTranspose2d(f * 2, f, kernel_size=2, stride=2)
        self.decoder1 = SmirkGenerator._block(f * 2, f, name="dec1")
        self.conv = nn.Conv2d(in_channels=f, out_channels=out_channels, kernel_size=1)
        self._cfg = (in_channels, out_channels, init_features, res_blocks)
        self.precision = 0
        self._handle, self._sig, self._ws = None, None, _lib.Workspace()

    @staticmethod
    def _block(in_channels, features, name):
        return nn.Sequential(OrderedDict([
            (name + "conv1", nn.Conv2d(in_channels, features, kernel_size=3, padding=1, bias=False)),
            (name + "norm1", nn.BatchNorm2d(num_features=features)),
            (name + "relu1", nn.ReLU(inplace=True)),
            (name + "conv2", nn.Conv2d(features, features, kernel_size=3, padding=1, bias=False)),
            (name + "norm2", nn.BatchNorm2d(num_features=features)),
            (name + "relu2", nn.ReLU(inplace=True)),
        ]))

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
        L = _lib.lib()
        keep = []
        ts = [v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")]
        arr = (_lib.c_f32p * len(ts))()
        for j, t in enumerate(ts):
            a, p = _lib.f32(t)
            keep.append(a)
            arr[j] = p
        d = _lib.SmkGeneratorDesc()
        d.in_channels, d.out_channels, d.init_features, d.res_blocks = self._cfg
        d.tensors, d.n_tensors, d.precision = C.cast(arr, C.POINTER(_lib.c_f32p)), len(ts), self.precision
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.smk_generator_create(C.byref(d), C.byref(h)), "smk_generator_create")
        self._handle, self._sig = _lib.NativeHandle(h, "smk_generator_destroy"), sig
        return self._handle

    def _release(self):
        self._handle = None                    # the native object dies with its last reference (_lib.NativeHandle)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        nn.Module.__init__(new)
        for k, v in self.__dict__.items():
            if k not in ("_handle", "_sig", "_ws"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._handle, new._sig, new._ws = None, None, _lib.Workspace()
        return new

    @torch.no_grad()
    def forward(self, x):
        _lib.require_cuda(x, "x")
        if self.training:                      # checked on every call: .train() after the first forward must not silently run eval BN
            raise RuntimeError("smirk_b200.SmirkGenerator: train-mode BatchNorm is not implemented (forward/eval only)")
        dev = x.device
        L = _lib.lib()
        h = self._native(dev)
        x = _lib.dev_f32(x, "x")
        cin, cout = self._cfg[0], self._cfg[1]
        if x.dim() != 4 or tuple(x.shape[1:]) != (cin, 224, 224):
            raise RuntimeError("smirk_b200.SmirkGenerator: expected x [B,%d,224,224], got %s" % (cin, tuple(x.shape)))
        B = x.shape[0]
        y = torch.empty(B, cout, 224, 224, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = self._ws.get(L.smk_generator_workspace_bytes(h, B), dev)
            _lib.check(L.smk_generator_forward(h, _lib.ptr(x), B, _lib.ptr(y), _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr(dev)), "smk_generator_forward")
        return y
