"""Micro-benchmark of the fused expand+depthwise kernel (xdw_tc.cu) on the encoder's real layer shapes.

    python tools/bench_xdw.py [--batch 32] [--reps 20] [--only K160]

Every shape is the (H, Cin, mid, stride) of an inverted-residual block of tf_mobilenetv3_{large,small}_minimal_100
at 224x224 input (reference src/smirk_encoder.py:7-12).  Times are CUDA-event means over `reps` launches with an
L2 flush (write of a 256 MiB buffer) between launches; GB/s is over the algorithmic bytes x + d + params.
Also the target for `ncu --set full --import-source on -k regex:xdw_kernel`.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smirk_b200 import _lib  # noqa: E402

# (H, Cin, mid, stride, launches per encoder pass)
SHAPES = [
    (112, 16, 64, 2, 2), (56, 24, 72, 1, 2), (56, 24, 72, 2, 2), (28, 40, 120, 1, 4), (28, 40, 240, 2, 4),
    (14, 80, 200, 1, 2), (14, 80, 184, 1, 4), (14, 80, 480, 1, 2), (14, 112, 672, 1, 2), (14, 112, 672, 2, 2),
    (7, 160, 960, 1, 4),
    (56, 16, 72, 2, 1), (28, 24, 88, 1, 1), (28, 24, 96, 2, 1), (14, 40, 240, 1, 2), (14, 40, 120, 1, 1),
    (14, 48, 144, 1, 1), (14, 48, 288, 2, 1), (7, 96, 576, 1, 2),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--x3", action="store_true", help="the 3xTF32 variant (smk_debug_xdw3x: weights split into TF32 heads and tails)")
    a = ap.parse_args()
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    P = lambda t: t.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    tot = 0.0
    for (H, Cin, mid, stride, n) in SHAPES:
        name = "H%d_K%d_N%d_s%d" % (H, Cin, mid, stride)
        if a.only and a.only not in name:
            continue
        B = a.batch
        Ho = (H + stride - 1) // stride
        x = torch.randn(B, H, H, Cin, device=dev)
        w1 = torch.randn(mid, Cin, device=dev) / Cin ** 0.5
        s1, b1 = torch.rand(mid, device=dev) + 0.5, torch.randn(mid, device=dev) * 0.2
        wd = torch.randn(9, mid, device=dev) / 3.0
        s2, b2 = torch.rand(mid, device=dev) + 0.5, torch.randn(mid, device=dev) * 0.2
        out = torch.empty(B, Ho, Ho, mid, device=dev)

        w1_hi = (w1.view(torch.int32) & -8192).view(torch.float32)
        w1_lo = w1 - w1_hi

        def run():
            if a.x3:
                rc = lib.smk_debug_xdw3x(P(x), B, H, H, Cin, P(w1_hi), P(w1_lo), P(s1), P(b1), mid, P(wd), P(s2), P(b2), stride, P(out), st)
                assert rc == 0, lib.smk_last_error()
                return
            rc = lib.smk_debug_xdw(P(x), B, H, H, Cin, P(w1), P(s1), P(b1), mid, P(wd), P(s2), P(b2), stride, 1, P(out), st)
            assert rc == 0, lib.smk_last_error()

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(a.reps):
            if not a.no_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        us = ms / a.reps * 1e3
        nbytes = 4.0 * (x.numel() + out.numel() + mid * (Cin + 13))
        print("%-22s x%d  %7.1f us  %7.1f GB/s   (%.1f MB)" % (name, n, us, nbytes / us * 1e-3, nbytes / 1e6))
        tot += us * n
    print("sum over an encoder pass: %.1f us" % tot)


if __name__ == "__main__":
    main()
