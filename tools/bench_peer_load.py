"""torchrun --nproc-per-node 2 tools/bench_peer_load.py — what the in-pipeline gather costs the encoder pipeline (B = 32) as a
function of how many 21 MB pushes follow each batch and how large they are: separates the host-side submission cost (tiny
pushes) from the memory-system cost (full pushes).  Emulates the 8-GPU traffic on 2 GPUs (k = 7 pushes to the one peer)."""
import ctypes as C
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import smirk_b200
from smirk_b200 import _lib, synth_assets, synth_inputs
from smirk_b200.pipeline import SmirkPipeline, _PeerBuffer

rank, local, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
root = synth_assets.materialize(os.path.join(tempfile.gettempdir(), "smk_assets_peer_%d" % rank)); os.chdir(root)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
shard = 21 << 20
buf = _PeerBuffer(8 * shard, dev)
handles = [None] * ws
dist.all_gather_object(handles, buf.handle)
buf.map_peers(handles, rank, 0)
peer = buf.ptrs[(rank + 1) % ws]
fan = C.c_void_p(); _lib.check(L.smk_peer_fan_create(8, C.byref(fan)), "fan")
enc = smirk_b200.SmirkEncoder(); enc.load_state_dict(synth_inputs.random_state_dict(enc.state_dict(), seed=7)); enc = enc.eval().to(dev); enc.precision = 3
pipe = SmirkPipeline(enc, smirk_b200.FLAME().to(dev), smirk_b200.Renderer().to(dev), None, device=dev, slots=4)
imgs = [synth_inputs.images(32, 900 + i).to(dev) for i in range(4)]
comm = torch.cuda.Stream(device=dev)
stage = torch.empty(shard, dtype=torch.uint8, device=dev)
mode = {"k": 0, "bytes": shard, "local": False}


def gather(Ln, rec):
    k = mode["k"]
    if k == 0:
        return
    with torch.cuda.stream(comm):
        comm.wait_event(Ln.computed)
        base = buf.ptr if mode["local"] else peer
        dsts = (C.c_void_p * k)(*[base + j * shard for j in range(k)])
        _lib.check(L.smk_peer_fan_push(fan, dsts, k, stage.data_ptr(), mode["bytes"], comm.cuda_stream), "fan push")
        Ln.gathered.record(comm)


pipe._gather = gather
pipe._gather_keys = ("x",)
for i in range(12):
    pipe.submit(i, imgs[i % 4])
pipe.join(); torch.cuda.synchronize()


def run(name, k, nbytes, local_dst=False, steps=300):
    mode.update(k=k, bytes=nbytes, local=local_dst)
    dist.barrier()
    for i in range(20):
        pipe.submit(i, imgs[i % 4])
    pipe.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pipe.submit(i, imgs[i % 4])
    pipe.join()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-52s %7.4f ms/step  %8.0f faces/s/GPU" % (name, t.item(), 32 / t.item() * 1e3), flush=True)


run("no pushes", 0, 0)
run("7 pushes of 256 B to the peer (host cost only)", 7, 256)
run("1 push of 21 MB to the peer", 1, shard)
run("7 pushes of 21 MB to the peer (8-GPU traffic)", 7, shard)
run("7 copies of 21 MB inside this GPU", 7, shard, local_dst=True)
run("no pushes (again)", 0, 0)
dist.destroy_process_group()
