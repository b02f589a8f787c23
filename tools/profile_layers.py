#!/usr/bin/env python
"""Per-layer device times of one eager pass of the hot path, from the library's built-in event
profiler at detail level 2 (GEMM / depthwise launches tagged with their M, K, N).

    python tools/profile_layers.py [--batch 32] [--generator] [--precision tf32x3] [--out profiles/x.json]

Prints one line per (kernel, shape): launches per step, microseconds per launch, achieved GB/s on
algorithmic bytes and TFLOP/s — the table the roofline work is steered by."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--generator", action="store_true")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32", "tf32-unfused", "tf32x3"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = os.path.abspath(args.out) if args.out else None
    import torch
    import smirk_b200
    from smirk_b200 import _lib, synth_assets, synth_inputs
    from smirk_b200.pipeline import SmirkPipeline
    root = synth_assets.materialize(os.path.join(tempfile.gettempdir(), "smk_assets_prof"))
    os.chdir(root)
    dev = torch.device("cuda:0")
    enc = smirk_b200.SmirkEncoder()
    enc.load_state_dict(synth_inputs.random_state_dict(enc.state_dict(), seed=7))
    enc = enc.eval().to(dev)
    enc.precision = {"fp32": 0, "tf32-unfused": 1, "tf32": 2, "tf32x3": 3}[args.precision]
    gen = None
    if args.generator:
        gen = smirk_b200.SmirkGenerator(6, 3, 32, 5)
        gen.load_state_dict(synth_inputs.random_state_dict(gen.state_dict(), seed=7))
        gen = gen.eval().to(dev)
        gen.precision = min(enc.precision, 1)
    fl = smirk_b200.FLAME().to(dev)
    stage = None
    if gen is not None:                     # the real masking step between renderer and generator, as in bench.py
        from smirk_b200.masking import MaskingStage
        stage = MaskingStage(fl.faces_tensor, synth_inputs.face_probabilities(fl.faces_tensor.shape[0]), seed=1234)
    pipe = SmirkPipeline(enc, fl, smirk_b200.Renderer().to(dev), gen, device=dev, masking=stage)
    B = args.batch
    imgs = [synth_inputs.images(B, 100 + i).to(dev) for i in range(4)]
    masks = [synth_inputs.hull_masks(B, 200 + i).to(dev) for i in range(4)] if gen is not None else None
    L = _lib.lib()
    for i in range(3):
        pipe.forward(imgs[i % 4], masks[i % 4] if masks else None)
    torch.cuda.synchronize()
    L.smk_profiler_reset(); L.smk_profiler_enable(2)
    for i in range(args.steps):
        pipe.forward(imgs[i % 4], masks[i % 4] if masks else None)
    torch.cuda.synchronize()
    rep = _lib.profiler_report()
    L.smk_profiler_enable(0); L.smk_profiler_reset()
    tot = sum(v["ms"] for v in rep.values())
    rows = []
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        rows.append(dict(kernel=k, launches_per_step=v["launches"] / args.steps, us_per_launch=1e3 * v["ms"] / v["launches"],
                         us_per_step=1e3 * v["ms"] / args.steps, share=v["ms"] / tot, gbs=v["bytes"] / v["ms"] / 1e6,
                         tflops=v["flops"] / v["ms"] / 1e9, mbytes_per_launch=v["bytes"] / v["launches"] / 1e6))
    print("batch %d  precision %s  generator %s  eager device time %.3f ms/step" % (B, args.precision, bool(gen), tot / args.steps))
    for r in rows:
        print("%-44s x%-4.1f %9.1f us/launch %9.1f us/step %5.1f%%  %8.1f GB/s %8.2f TFLOP/s  %8.2f MB" % (
            r["kernel"], r["launches_per_step"], r["us_per_launch"], r["us_per_step"], 100 * r["share"], r["gbs"], r["tflops"], r["mbytes_per_launch"]))
    if out:
        with open(out, "w") as fh:
            json.dump(dict(batch=B, precision=args.precision, generator=bool(gen), ms_per_step=tot / args.steps, layers=rows), fh, indent=1)


if __name__ == "__main__":
    main()
