"""torchrun --nproc-per-node N tools/bench_peer.py — copy-engine push bandwidth into CUDA-IPC mapped peer buffers
(csrc/peer.cu), idle and under the encoder pipeline's load.  Rank 0 prints one line per case."""
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
buf = _PeerBuffer(ws * shard, dev)
handles = [None] * ws
dist.all_gather_object(handles, buf.handle)
buf.map_peers(handles, rank, shard)
src = torch.empty(shard, dtype=torch.uint8, device=dev)
fan = C.c_void_p(); _lib.check(L.smk_peer_fan_create(8, C.byref(fan)), "fan")
comm = torch.cuda.Stream(device=dev)
peer = buf.ptrs[(rank + 1) % ws] + rank * shard
dsts8 = (C.c_void_p * 8)(*[peer] * 8)


def timed(fn, reps=10):
    with torch.cuda.stream(comm):
        fn(); comm.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(comm)
        for _ in range(reps):
            fn()
        e1.record(comm)
    comm.synchronize()
    return e0.elapsed_time(e1) / reps


def serial():
    for _ in range(8):
        _lib.check(L.smk_peer_push(peer, src.data_ptr(), shard, comm.cuda_stream), "push")


def fanned():
    _lib.check(L.smk_peer_fan_push(fan, dsts8, 8, src.data_ptr(), shard, comm.cuda_stream), "fan push")


def allpeers():
    _lib.check(L.smk_peer_fan_push(fan, buf.dsts[0], ws - 1, src.data_ptr(), shard, comm.cuda_stream), "fan push")


def report(name, ms, nbytes):
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-44s %8.3f ms  %7.1f GB/s per rank" % (name, t.item(), nbytes / t.item() / 1e6), flush=True)


dist.barrier()
report("idle: 8 x 21 MB to one peer, one stream", timed(serial), 8 * shard)
report("idle: 8 x 21 MB to one peer, 8 streams", timed(fanned), 8 * shard)
report("idle: 21 MB to each of %d peers, fan" % (ws - 1), timed(allpeers), (ws - 1) * shard)
enc = smirk_b200.SmirkEncoder(); enc.load_state_dict(synth_inputs.random_state_dict(enc.state_dict(), seed=7)); enc = enc.eval().to(dev); enc.precision = 3
pipe = SmirkPipeline(enc, smirk_b200.FLAME().to(dev), smirk_b200.Renderer().to(dev), None, device=dev, slots=4)
imgs = [synth_inputs.images(32, 900 + i).to(dev) for i in range(4)]
for i in range(8):
    pipe.submit(i, imgs[i % 4])
pipe.join(); torch.cuda.synchronize()
dist.barrier()
for name, fn, n in (("loaded: 8 x 21 MB to one peer, one stream", serial, 8 * shard), ("loaded: 8 x 21 MB to one peer, 8 streams", fanned, 8 * shard),
                    ("loaded: 21 MB to each of %d peers, fan" % (ws - 1), allpeers, (ws - 1) * shard)):
    for i in range(60):                                   # ~55 ms of encoder work queued behind the copies' start
        pipe.submit(i, imgs[i % 4])
    ms = timed(fn, reps=5)
    pipe.join(); torch.cuda.synchronize()
    report(name, ms, n)
    dist.barrier()
dist.destroy_process_group()
