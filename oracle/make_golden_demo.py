"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/demo.npz by running the REFERENCE's own, unmodified
``demo.py`` (/root/reference) with the reference's own classes on the CPU — third-party pieces (timm, pytorch3d, skimage,
mediapipe) replaced by the stand-ins of oracle/ref_harness.py and tests/dropin_support.py — on the synthetic input image
and synthetic checkpoint of tests/dropin_support.py.  Re-run in the build container: ``python -m oracle.make_golden_demo``.

Stored: the grid the script writes (demo.py:170-182) without ``--use_smirk_generator`` (cropped image | rendered mesh).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import cv2
    import dropin_support as ds
    from smirk_b200 import synth_assets
    from oracle import ref_harness
    work = tempfile.mkdtemp(prefix="smk_demo_golden_")
    root = synth_assets.materialize(os.path.join(work, "assets_root"))
    img = ds.synthetic_image(os.path.join(work, "face.png"))
    ck = ds.write_checkpoint(os.path.join(work, "ck.pt"), with_generator=False)
    out = os.path.join(work, "out")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "dropin_support.py"), "--mode", "reference", "--script",
                           os.path.join(ref_harness.REF, "demo.py"), "--cwd", root, "--", "--input_path", img, "--device", "cpu",
                           "--checkpoint", ck, "--out_path", out])
    grid = cv2.imread(os.path.join(out, "face.png"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "demo.npz"), grid=grid)
    print("demo.npz:", grid.shape, "rendered coverage %.3f" % float((grid[:, 224:] > 0).any(-1).mean()))


if __name__ == "__main__":
    main()
