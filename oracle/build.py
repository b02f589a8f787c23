"""Compile the C part of the oracle (raster_ref.c) -> oracle/libsmk_oracle.so.  Test infrastructure."""
import os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsmk_oracle.so")

def build(force=False):
    src = os.path.join(HERE, "raster_ref.c")
    if (not force) and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(src):
        return SO
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-o", SO, src, "-lm"])
    return SO

if __name__ == "__main__":
    print(build(force=True))
