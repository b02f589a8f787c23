#!/usr/bin/env python
"""bench.py — faces/sec of the SMIRK hot path (encode -> FLAME -> render @224^2) on N B200s.

Contract:
  python bench.py --gpus N --steps K --warmup W            product arm (under torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port) arm

A step = one pass of the hot path over one batch of synthetic 224x224 RGB faces per GPU
(BASELINE.json configs[1]: encoder+FLAME+raster, batch 32 per GPU; `--generator` switches to
configs[2], the full cycle with SmirkGenerator; `--batch` changes the per-GPU batch).
  value  : faces/s with the input batches already resident in HBM, CUDA-graph replay, timed with CUDA
           events over exactly K steps between barriers, max over ranks.  Inputs rotate over a set of
           batches larger than L2 so no step re-reads its input from cache.
  e2e    : the same metric through SmirkPipeline.run_host(): pinned host images -> H2D -> graph ->
           D2H of rendered image + vertices + FLAME parameters into pinned host memory, every step.
  roofline: dominant kernel (largest share of device time) from the library's built-in event
           profiler, live in this run: algorithmic bytes (or FLOPs) per launch / mean launch time.
  cpu_baseline: the oracle port of the reference (torch CPU ops + C rasteriser) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "faces/sec encode->FLAME->render @224^2"
UNIT = "faces/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops", 1590.0)), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(root, sample, with_generator, threads=None):
    """One pass of the reference's CPU path (oracle port) over `sample` faces; returns seconds."""
    import torch
    from smirk_b200 import synth_inputs
    from oracle import encoder_ref, flame_ref, render_ref, generator_ref
    if threads:
        torch.set_num_threads(threads)
    st = cpu_reference_pass.state
    if not st:
        import smirk_b200
        enc = smirk_b200.SmirkEncoder()
        st["enc_sd"] = synth_inputs.random_state_dict(enc.state_dict(), seed=7)
        st["fc"], st["rc"] = flame_ref.FlameConstants(root), render_ref.RenderConstants(root)
        if with_generator:
            gen = smirk_b200.SmirkGenerator(6, 3, 32, 5)
            st["gen_sd"] = synth_inputs.random_state_dict(gen.state_dict(), seed=7)
    img = synth_inputs.images(sample, 9001)
    t0 = time.perf_counter()
    with torch.no_grad():
        p = encoder_ref.encoder_forward_ref(st["enc_sd"], img)
        p = {k: v for k, v in p.items() if not k.startswith("_")}
        fo = flame_ref.flame_forward_ref(st["fc"], p)
        ro = render_ref.render_forward_ref(st["rc"], fo["vertices"], p["cam"], landmarks_fan=fo["landmarks_fan"],
                                           landmarks_mp=fo["landmarks_mp"])
        if with_generator:
            generator_ref.generator_forward_ref(st["gen_sd"], torch.cat([ro["rendered_img"], synth_inputs.masked_images(sample, 9002)], 1))
    return time.perf_counter() - t0


cpu_reference_pass.state = {}


def pick_cpu_threads(root, with_generator):
    """torch's intra-op pool is pathological at some thread counts on small-kernel workloads (SURVEY.md
    §6: lbs() 1.4 ms at 1 thread vs 27 ms at 8).  Time a 2-face pass at a few thread counts and keep
    the fastest, so the CPU arm gets the best configuration the host offers."""
    import torch
    n = os.cpu_count() or 1
    cands = sorted({1, max(1, n // 4), max(1, n // 2), n, min(n, 16), min(n, 32)})
    cpu_reference_pass(root, 1, with_generator, threads=1)                 # build state, warm caches
    best = None
    for c in cands:
        sec = cpu_reference_pass(root, 2, with_generator, threads=c)
        if best is None or sec < best[1]:
            best = (c, sec)
    torch.set_num_threads(best[0])
    return best[0]


def run_reference_arm(args, root, rank, world):
    """--impl reference: the reference's own CPU implementation (oracle port; the reference is pure
    Python + third-party wheels and cannot travel to the GPU box).  Rank 0 only."""
    if rank != 0:
        return
    import torch
    sample = args.cpu_sample
    threads = pick_cpu_threads(root, args.generator)
    cpu_reference_pass(root, sample, args.generator, threads=threads)
    t = [cpu_reference_pass(root, sample, args.generator, threads=threads) for _ in range(max(1, min(args.steps, 5)))]
    sec = sum(t) / len(t)
    val = sample / sec
    wl = workload_name(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(t),
        "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "faces_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d faces per step of the same synthetic workload (oracle port: torch CPU ops + C rasteriser)" % sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    if args.generator:
        return "configs[2]: full cycle encoder->FLAME->raster->smirk_generator 224x224, batch %d per GPU" % args.batch
    return "configs[1]: encoder+FLAME+raster 224x224, batch %d per GPU" % args.batch


# --------------------------------------------------------------------------------------- product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="smirk_b200", choices=["smirk_b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="faces per GPU per step (configs[1] = 32)")
    ap.add_argument("--generator", action="store_true", help="include SmirkGenerator (configs[2], full cycle)")
    ap.add_argument("--cpu-sample", type=int, default=16, help="faces per CPU-baseline pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slots", type=int, default=4, help="pipeline lanes: consecutive batches alternate over this many stream/graph replicas")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel breakdown JSON here")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "tf32-unfused"],
                    help="tf32: 1x1/3x3/transposed convs on tcgen05 tensor cores (the reference's own cuDNN default); "
                         "fp32: every conv on the exact fp32 CUDA-core path")
    args = ap.parse_args()
    if args.profile_out:
        args.profile_out = os.path.abspath(args.profile_out)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from smirk_b200 import synth_assets
    root = synth_assets.materialize(os.path.join(tempfile.gettempdir(), "smk_assets_bench_%d" % rank))
    os.chdir(root)

    if args.impl == "reference":
        run_reference_arm(args, root, rank, world)
        return

    import torch
    import smirk_b200
    from smirk_b200 import _lib, synth_inputs
    from smirk_b200.pipeline import SmirkPipeline
    assert torch.cuda.is_available(), "bench.py (product arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    enc = smirk_b200.SmirkEncoder()
    enc.load_state_dict(synth_inputs.random_state_dict(enc.state_dict(), seed=7))
    enc = enc.eval().to(dev)
    # tf32 = tensor-core convs + fused expand/depthwise blocks in the encoder; tf32-unfused keeps one kernel per layer
    enc.precision = {"fp32": 0, "tf32-unfused": 1, "tf32": 2}[args.precision]
    gen = None
    if args.generator:
        gen = smirk_b200.SmirkGenerator(6, 3, 32, 5)
        gen.load_state_dict(synth_inputs.random_state_dict(gen.state_dict(), seed=7))
        gen = gen.eval().to(dev)
        gen.precision = min(enc.precision, 1)
    pipe = SmirkPipeline(enc, smirk_b200.FLAME().to(dev), smirk_b200.Renderer().to(dev), gen, device=dev, slots=args.slots)

    # rotating input set larger than L2 (126 MB): R batches of B x 602 KB
    per = B * 3 * 224 * 224 * 4 * (2 if gen is not None else 1)
    R = max(2, -(-160_000_000 // per))
    host_imgs = [synth_inputs.images(B, 5000 + rank * 100 + i).pin_memory() for i in range(R)]
    host_masks = [synth_inputs.masked_images(B, 6000 + rank * 100 + i).pin_memory() for i in range(R)] if gen is not None else None
    dev_imgs = [h.to(dev) for h in host_imgs]
    dev_masks = [h.to(dev) for h in host_masks] if gen is not None else None
    rec = pipe.capture(B)
    for lane in range(1, pipe.slots):
        pipe.capture(B, lane)

    def dev_step(i):
        # batch i goes to lane i % slots (own stream + graph replica): consecutive batches overlap
        pipe.submit(i, dev_imgs[i % R], dev_masks[i % R] if gen is not None else None)

    # ---- device-resident throughput (value) ----
    for i in range(W):
        dev_step(i)
    pipe.join()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_before = _lib.lib().smk_launch_count()
    barrier()
    e0.record()
    for i in range(K):
        dev_step(W + i)
    pipe.join()                                  # the timing stream waits for every lane
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- end to end through the public host API (e2e) ----
    keys = ("rendered_img", "vertices", "params") + (("reconstructed_img",) if gen is not None else ())
    h2d, d2h = pipe.bytes_per_step(B, keys)
    for i in range(W):
        pipe.run_host(host_imgs[i % R], i, host_masks[i % R] if gen is not None else None, keys)
    pipe.join()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for i in range(K):
        last = pipe.run_host(host_imgs[(W + i) % R], W + i, host_masks[(W + i) % R] if gen is not None else None, keys)
    pipe.join()                                  # includes the copy streams: every D2H has landed
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms_e2e = float(t2.item())
    checksum = float(last["params"].double().abs().sum())                       # touch the host result

    # ---- per-kernel breakdown (eager, event-bracketed launches; rank 0) ----
    breakdown, roof = None, None
    if rank == 0:
        L = _lib.lib()
        for i in range(2):
            pipe.forward(dev_imgs[i % R], dev_masks[i % R] if gen is not None else None)
        torch.cuda.synchronize(dev)
        L.smk_profiler_reset(); L.smk_profiler_enable(1)
        NP = 5
        for i in range(NP):
            pipe.forward(dev_imgs[(i + 3) % R], dev_masks[(i + 3) % R] if gen is not None else None)
        torch.cuda.synchronize(dev)
        breakdown = _lib.profiler_report()
        L.smk_profiler_enable(0); L.smk_profiler_reset()
        peaks = load_peaks()
        tot = sum(v["ms"] for v in breakdown.values())
        for v in breakdown.values():
            v["share"] = v["ms"] / tot
            v["ms_per_launch"] = v["ms"] / v["launches"]
            v["gbs"] = v["bytes"] / v["ms"] / 1e6
            v["tflops"] = v["flops"] / v["ms"] / 1e9
        top_tag, top = max(breakdown.items(), key=lambda kv: kv[1]["ms"])
        ai = top["flops"] / max(top["bytes"], 1.0)
        ridge = peaks["tensor"] * 1e3 / peaks["hbm"]
        if "tc" in top_tag and ai > ridge:
            roof = {"bound": "tensor", "achieved": top["tflops"], "peak": peaks["tensor"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm"], "unit": "GB/s"}
        # DRAM traffic per launch of that kernel from the committed ncu capture of this same workload
        # (profiles/*ncu_dram_traffic*.json: dram__bytes_read.sum + dram__bytes_write.sum); null if the
        # capture does not cover this configuration.
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_dram_traffic_c2_b32_%s.json" % args.precision)
        if os.path.exists(tpath) and B == 32 and gen is None:
            tk = json.load(open(tpath)).get("kernels", {}).get(top_tag.split(":")[0])
            if tk:
                traffic = tk["traffic_bytes_per_launch"]
        roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=top_tag, share_of_step=top["share"],
                    launches_per_step=top["launches"] / NP, us_per_launch=top["ms_per_launch"] * 1e3,
                    peak_source=peaks["source"],
                    algorithmic_bytes_per_launch=top["bytes"] / top["launches"],
                    algorithmic_flops_per_launch=top["flops"] / top["launches"])
        if args.profile_out:
            with open(args.profile_out, "w") as fh:
                json.dump({"batch": B, "steps_profiled": NP, "eager_ms_per_step": tot / NP, "kernels": breakdown}, fh, indent=1)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        threads = pick_cpu_threads(root, args.generator)                 # also warms up (weights, C oracle)
        sec = cpu_reference_pass(root, args.cpu_sample, args.generator, threads=threads)
        cpu = {"value": args.cpu_sample / sec, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
               "sample": "%d faces, one pass of the same workload (oracle port: torch CPU ops + C rasteriser), %.1f s" % (args.cpu_sample, sec)}

    if rank == 0:
        faces = B * world * K
        line = {
            "metric": METRIC, "value": faces / (ms_total / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 convs (fp32 accumulate), f32 elsewhere" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "precision": args.precision, "global_batch": B * world, "faces_per_gpu_per_step": B,
                       "image": "224x224 RGB fp32", "parallelism": "frame-shard dp%d" % world,
                       "l2_policy": "inputs rotate over %d batches (%.0f MB > 126 MB L2)" % (R, R * per / 1e6),
                       "execution": "CUDA graph replay, %d kernels per step, %d-lane software pipeline over consecutive steps" % (rec["launches"], pipe.slots)},
            "e2e": {"value": faces / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "api": "SmirkPipeline.run_host (pinned host in/out, %d lanes, copies on a second stream)" % pipe.slots,
                    "host_checksum": checksum},
            "gpu_launches": int(rec["launches"]) * K,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        if breakdown:
            line["kernel_shares"] = {k: round(v["share"], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])[:8]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
