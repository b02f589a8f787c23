"""torchrun --nproc-per-node N tools/check_gather.py — the in-pipeline all-gather (both backends) against a plain NCCL
all_gather of the same outputs; prints one OK line per backend on rank 0."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import smirk_b200
from smirk_b200 import synth_assets, synth_inputs
from smirk_b200.pipeline import SmirkPipeline

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
root = synth_assets.materialize(os.path.join(tempfile.gettempdir(), "smk_assets_chk_%d" % rank)); os.chdir(root)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
enc = smirk_b200.SmirkEncoder(); enc.load_state_dict(synth_inputs.random_state_dict(enc.state_dict(), seed=7)); enc = enc.eval().to(dev); enc.precision = 3
keys = ("rendered_img", "vertices", "params")
for backend in ("nccl", "p2p"):
    pipe = SmirkPipeline(enc, smirk_b200.FLAME().to(dev), smirk_b200.Renderer().to(dev), None, device=dev, slots=2)
    pipe.enable_gather(keys, backend=backend)
    B = 4
    imgs = [synth_inputs.images(B, 900 + 10 * rank + i).to(dev) for i in range(4)]
    ok = True
    for i in range(4):
        o = pipe.submit(i, imgs[i])
        pipe.gather_sync()
        for k in keys:
            ref = torch.empty((dist.get_world_size() * B,) + tuple(o[k].shape[1:]), device=dev)
            dist.all_gather_into_tensor(ref, o[k].contiguous())
            ok = ok and torch.equal(pipe.gathered(i, k), ref)
        dist.barrier()
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("gather backend %s (resolved %s): %s" % (backend, pipe._gather_backend, "OK" if t.item() else "MISMATCH"), flush=True)
    del pipe
dist.destroy_process_group()
