"""Rank -> NUMA placement for the host side of the pipeline (pinned staging buffers, the Python launch thread).

On an 8-GPU HGX box half of the GPUs hang off each CPU socket.  A rank whose launch thread or pinned buffers live on
the far socket pays the inter-socket hop on every H2D / D2H copy, and with eight ranks streaming ~40 MB per step each
that hop is the end-to-end limiter (round 1: e2e scaling efficiency 0.84 at N = 8 with device efficiency 0.99).
``pin_to_gpu`` binds the calling process to the CPUs NVML reports as local to the GPU *before* any pinned allocation is
made, so first-touch places the staging buffers on the GPU's own socket.
"""
import os
import re
import subprocess


def _cpus_from_nvml(index):
    import pynvml
    pynvml.nvmlInit()
    try:
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        words = (n + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return [c for c in cpus if c < n]
    finally:
        pynvml.nvmlShutdown()


def _cpus_from_topo(index):
    out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    for ln in out.splitlines():
        if re.match(r"^GPU%d\s" % index, ln):
            m = re.search(r"\s(\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*)\s+\d+", ln)
            if m:
                cpus = []
                for part in m.group(1).split(","):
                    a, _, b = part.partition("-")
                    cpus.extend(range(int(a), int(b or a) + 1))
                return cpus
    return []


def pin_to_gpu(index, visible=None):
    """Bind this process to the CPUs local to physical GPU ``index``; returns a small report dict.  ``visible`` maps a
    logical CUDA ordinal through CUDA_VISIBLE_DEVICES when that variable lists plain indices."""
    phys = index
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES") if visible is None else visible
    if cvd:
        ids = [x.strip() for x in cvd.split(",") if x.strip()]
        if index < len(ids) and ids[index].isdigit():
            phys = int(ids[index])
    rep = {"gpu": phys, "cpus": None, "source": None}
    for name, fn in (("nvml", _cpus_from_nvml), ("nvidia-smi topo", _cpus_from_topo)):
        try:
            cpus = fn(phys)
        except Exception:
            cpus = []
        if cpus:
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                rep.update(cpus="%d CPUs (%d..%d)" % (len(allowed), allowed[0], allowed[-1]), source=name, n_cpus=len(allowed))
                return rep
    rep["source"] = "unavailable (affinity unchanged)"
    return rep
