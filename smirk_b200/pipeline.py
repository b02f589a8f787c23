"""``SmirkPipeline`` — the whole per-frame hot path as one call: encode -> FLAME -> render (-> generator).

This is the public entry point ``bench.py`` and the demos use for batched work:

* ``forward(img)``            device tensors in, dict of device tensors out (what demo.py:107-112 does
                              with three module calls);
* ``capture(B)`` / ``replay`` the same work recorded once into a CUDA graph (one launch per batch instead
                              of ~250 kernel launches — the launch-bound regime of small batches);
* ``run_host(img_pinned)``    end-to-end from pinned host memory: H2D copy, graph replay, D2H of the
                              results into pinned buffers, double-buffered over two streams so the
                              copies of batch i+1 overlap the kernels of batch i;
* ``shard`` / ``all_gather``  frame-shard data parallelism (one process per GPU, contiguous split of the
                              batch; a single NCCL all-gather of the final outputs — SURVEY.md §8e).
"""
import torch

from . import _lib


def shard_bounds(n_frames, world_size, rank):
    """Contiguous split of ``n_frames`` over ``world_size`` ranks: rank g owns
    [lo, hi).  The first ``n_frames % world_size`` ranks take one extra frame."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_frames(t, n_frames=None, group=None):
    """Gather per-rank frame shards (dim 0) into the full batch on every rank.  Ragged shards are padded
    to the largest shard for one equal-count ``all_gather_into_tensor`` and trimmed afterwards."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    if ws == 1:
        return t
    counts = [shard_bounds(n_frames, ws, r)[1] - shard_bounds(n_frames, ws, r)[0] for r in range(ws)] \
        if n_frames is not None else [t.shape[0]] * ws
    mx = max(counts)
    if t.shape[0] < mx:
        pad = torch.zeros((mx - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], 0)
    out = torch.empty((ws * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx:r * mx + c] for r, c in enumerate(counts)], 0)


class SmirkPipeline:
    OUT_KEYS = ("rendered_img", "vertices", "transformed_vertices", "landmarks_fan", "landmarks_mp", "params")

    def __init__(self, encoder, flame, renderer, generator=None, device="cuda:0"):
        self.device = torch.device(device)
        self.encoder, self.flame, self.renderer, self.generator = encoder, flame, renderer, generator
        self._graphs = {}
        self._host = {}

    # ---- plain forward (device in, device out) -----------------------------------------------------
    @torch.no_grad()
    def forward(self, img, masked_img=None):
        p = self.encoder(img)
        fo = self.flame.forward(p)
        ro = self.renderer.forward(fo["vertices"], p["cam"], landmarks_fan=fo["landmarks_fan"],
                                   landmarks_mp=fo["landmarks_mp"])
        out = {
            "rendered_img": ro["rendered_img"], "vertices": fo["vertices"],
            "transformed_vertices": ro["transformed_vertices"], "landmarks_fan": ro["landmarks_fan"],
            "landmarks_mp": ro["landmarks_mp"],
            "params": torch.cat([p["pose_params"], p["cam"], p["shape_params"], p["expression_params"],
                                 p["eyelid_params"], p["jaw_params"]], 1),                      # [B,361]
        }
        if self.generator is not None:
            if masked_img is None:
                raise RuntimeError("SmirkPipeline: the generator stage needs `masked_img` (demo.py:165-167)")
            out["reconstructed_img"] = self.generator(torch.cat([ro["rendered_img"], masked_img], 1))
        return out

    # ---- CUDA graph -----------------------------------------------------------------------------------
    def capture(self, B):
        """Record forward() for batch size B into a CUDA graph with static input/output buffers."""
        if B in self._graphs:
            return self._graphs[B]
        dev = self.device
        static_in = torch.zeros(B, 3, 224, 224, device=dev)
        static_mask = torch.zeros(B, 3, 224, 224, device=dev) if self.generator is not None else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up: builds handles, sizes workspaces
            for _ in range(2):
                self.forward(static_in, static_mask)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.lib().smk_launch_count()
        with torch.cuda.graph(g):
            static_out = self.forward(static_in, static_mask)
        launches = _lib.lib().smk_launch_count() - n0
        rec = dict(graph=g, img=static_in, mask=static_mask, out=static_out, launches=int(launches))
        self._graphs[B] = rec
        return rec

    def replay(self, img, masked_img=None):
        rec = self.capture(img.shape[0])
        rec["img"].copy_(img, non_blocking=True)
        if rec["mask"] is not None:
            rec["mask"].copy_(masked_img, non_blocking=True)
        rec["graph"].replay()
        return rec["out"]

    # ---- host-to-host (pinned) ------------------------------------------------------------------------
    def host_buffers(self, B, keys=("rendered_img", "vertices", "params")):
        """Two sets of pinned staging buffers (double buffering) for run_host()."""
        key = (B, tuple(keys))
        if key not in self._host:
            rec = self.capture(B)
            sets = []
            for _ in range(2):
                sets.append(dict(
                    dev_in=torch.empty(B, 3, 224, 224, device=self.device),
                    dev_mask=torch.empty(B, 3, 224, 224, device=self.device) if rec["mask"] is not None else None,
                    dev_out={k: torch.empty_like(rec["out"][k]) for k in keys},
                    out={k: torch.empty(rec["out"][k].shape, dtype=rec["out"][k].dtype).pin_memory() for k in keys},
                    copy_stream=torch.cuda.Stream(device=self.device),
                    done=torch.cuda.Event(), staged=torch.cuda.Event(), computed=torch.cuda.Event()))
            self._host[key] = sets
        return self._host[key]

    def run_host(self, img_pinned, slot, masked_pinned=None, keys=("rendered_img", "vertices", "params")):
        """One batch, host to host.  ``slot`` alternates 0/1 between consecutive calls so that the H2D
        copy of this batch and the D2H copy of the previous one overlap compute.  Returns the pinned
        output dict of this slot; call ``sets[slot]['done'].synchronize()`` before reading it."""
        B = img_pinned.shape[0]
        s = self.host_buffers(B, keys)[slot]
        rec = self.capture(B)
        main = torch.cuda.current_stream(self.device)
        cs = s["copy_stream"]
        with torch.cuda.stream(cs):                         # H2D on the copy stream
            cs.wait_event(s["done"])                        # previous use of this slot fully drained
            s["dev_in"].copy_(img_pinned, non_blocking=True)
            if s["dev_mask"] is not None:
                s["dev_mask"].copy_(masked_pinned, non_blocking=True)
            s["staged"].record(cs)
        main.wait_event(s["staged"])
        rec["img"].copy_(s["dev_in"], non_blocking=True)
        if rec["mask"] is not None:
            rec["mask"].copy_(s["dev_mask"], non_blocking=True)
        rec["graph"].replay()
        for k in keys:                                      # detach results from the graph's static buffers
            s["dev_out"][k].copy_(rec["out"][k], non_blocking=True)
        s["computed"].record(main)
        with torch.cuda.stream(cs):                         # D2H on the copy stream
            cs.wait_event(s["computed"])
            for k in keys:
                s["out"][k].copy_(s["dev_out"][k], non_blocking=True)
            s["done"].record(cs)
        return s["out"]

    def bytes_per_step(self, B, keys=("rendered_img", "vertices", "params")):
        rec = self.capture(B)
        h2d = B * 3 * 224 * 224 * 4 * (2 if rec["mask"] is not None else 1)
        d2h = sum(rec["out"][k].numel() * rec["out"][k].element_size() for k in keys)
        return h2d, d2h
