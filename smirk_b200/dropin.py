"""Run the reference's own entry scripts on smirk_b200 without modifying them.

    cd <reference checkout>;  python -m smirk_b200.dropin demo.py --input_path ... [script args]

``install()`` registers our modules under the four names the reference imports its hot-path classes
from (``demo.py:5-7,62``, ``src/smirk_trainer.py:3-6``):

    src.smirk_encoder      -> smirk_b200.smirk_encoder    (SmirkEncoder, create_backbone, ...)
    src.FLAME.FLAME        -> smirk_b200.flame            (FLAME)
    src.renderer.renderer  -> smirk_b200.renderer         (Renderer)
    src.smirk_generator    -> smirk_b200.smirk_generator  (SmirkGenerator, ResnetBlock)

Everything else under ``src`` (``src.FLAME.lbs``, ``src.renderer.util``, ``src.utils.*``, losses, trainers)
still resolves to the reference's files: the real ``src`` package is imported first and only these four
submodules are replaced in ``sys.modules`` (and as attributes of their parent packages).
"""
import importlib
import os
import runpy
import sys
import types

_MAP = {
    "src.smirk_encoder": "smirk_b200.smirk_encoder",
    "src.FLAME.FLAME": "smirk_b200.flame",
    "src.renderer.renderer": "smirk_b200.renderer",
    "src.smirk_generator": "smirk_b200.smirk_generator",
}


def _ensure_package(name, path_hint):
    """Import the reference's package if it exists on sys.path; otherwise make an empty namespace so the
    aliases still work (e.g. in tests that have no reference checkout)."""
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pkg = types.ModuleType(name)
        pkg.__path__ = [path_hint] if path_hint and os.path.isdir(path_hint) else []
        sys.modules[name] = pkg
        return pkg


def install(reference_root=None):
    """Alias the four hot-path modules.  ``reference_root`` (default: cwd) is put on sys.path so the
    rest of the reference package keeps importing normally."""
    root = os.path.abspath(reference_root or os.getcwd())
    if root not in sys.path:
        sys.path.insert(0, root)
    for parent, sub in (("src", "src"), ("src.FLAME", "src/FLAME"), ("src.renderer", "src/renderer")):
        pkg = _ensure_package(parent, os.path.join(root, sub))
        if "." in parent:
            setattr(sys.modules[parent.rsplit(".", 1)[0]], parent.rsplit(".", 1)[1], pkg)
    for ref_name, ours in _MAP.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        parent, leaf = ref_name.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mod)
    return sorted(_MAP)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m smirk_b200.dropin <reference script.py> [script args...]")
    script = argv[0]
    install(os.path.dirname(os.path.abspath(script)) or os.getcwd())
    sys.argv = argv
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
