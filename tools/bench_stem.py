import sys, torch
sys.path.insert(0, '/root/repo')
from smirk_b200 import _lib
lib = _lib.lib(); dev = torch.device('cuda:0'); P = lambda t: t.data_ptr()
B = 32
img = torch.rand(B, 3, 224, 224, device=dev)
t = [torch.randn(27, 16, device=dev), torch.rand(16, device=dev), torch.rand(16, device=dev), torch.randn(9, 16, device=dev), torch.rand(16, device=dev),
     torch.rand(16, device=dev), torch.randn(16, 16, device=dev), torch.rand(16, device=dev), torch.rand(16, device=dev)]
st = torch.cuda.current_stream().cuda_stream
for stride in (1, 2):
    out = torch.empty(B, 112 // stride, 112 // stride, 16, device=dev)
    for _ in range(3):
        assert lib.smk_debug_stem_ds(P(img), B, 224, 224, *[P(x) for x in t], stride, 1, P(out), st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.smk_debug_stem_ds(P(img), B, 224, 224, *[P(x) for x in t], stride, 1, P(out), st)
    e1.record(); torch.cuda.synchronize()
    print("stride", stride, "us", e0.elapsed_time(e1) * 100)
