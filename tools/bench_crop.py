"""Device time of the crop / warp front end (configs[3]-style: 1080p BGR frames -> 224x224 RGB tensors) and of the
warp back to the frame.  CUDA-event mean over `--reps` calls after warm-up; inputs resident on the device."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smirk_b200 import crop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, W = a.batch, 1080, 1920
    frames = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(0)
    tf = []
    for i in range(B):
        cx, cy, half = 960 + rng.uniform(-300, 300), 540 + rng.uniform(-150, 150), rng.uniform(120, 260)
        tf.append(crop.landmark_box_transform(np.array([[cx - half, cy - half], [cx + half, cy + half]]), 1.4, 224))
    rendered = torch.rand(B, 3, 224, 224, device=dev)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    t1 = timed(lambda: crop.crop_to_tensor(frames, tf, 224))
    t2 = timed(lambda: crop.warp_back(rendered, tf, (H, W)))
    print("crop_to_tensor  B=%d 1080p -> 224: %.3f ms/batch  (%.0f frames/s; includes host-side matrix inverse + H2D of %d matrices)" % (B, t1, B / t1 * 1e3, B))
    print("warp_back       B=%d 224 -> 1080p: %.3f ms/batch  (%.0f frames/s, %.1f GB/s written)" % (B, t2, B / t2 * 1e3, B * H * W * 3 / t2 / 1e6))


if __name__ == "__main__":
    main()
