"""Builds ``oracle/_ref/``: the UNMODIFIED reference package, importable where /root/reference is
not (the GPU box), so that ``bench.py --impl reference`` and ``--impl cudnn`` run the reference's
own code (``cpu_baseline.kind = "reference"``) instead of the oracle's restatement.

    python oracle/make_ref.py            # needs /root/reference (the build container)

What it does, and nothing else:
  * copies the reference's pure-Python package directory ``/root/reference/openglottal`` to
    ``oracle/_ref/openglottal`` byte for byte (no build system is run; the package has no compiled
    code), recording a SHA-256 of every file in ``oracle/_ref/MANIFEST.json``;
  * writes ``oracle/_ref/ultralytics/__init__.py``, a three-line stand-in for the one third-party
    import the package makes at import time that is absent from this image
    (openglottal/models/detector.py:6, ``from ultralytics import YOLO``; the YOLO pipelines are out
    of scope and never constructed).

``oracle/_ref/`` is listed in .gitignore (reference sources never enter the history) and NOT in
.gpurunignore, so it travels to the GPU box like the built ``.so``. TEST/BENCH INFRASTRUCTURE ONLY:
nothing under ``openglottal_b200/`` imports it.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/openglottal")
DST = HERE / "_ref"


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is in place (freshly built or already there)."""
    if not REF_SRC.is_dir():
        return (DST / "openglottal" / "__init__.py").exists()
    if DST.exists():
        shutil.rmtree(DST)
    shutil.copytree(REF_SRC, DST / "openglottal", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    (DST / "ultralytics").mkdir()
    (DST / "ultralytics" / "__init__.py").write_text(
        '"""Stand-in for the absent third-party package (see oracle/make_ref.py)."""\n'
        "class YOLO:  # pragma: no cover - never constructed on the unet-only path\n"
        "    def __init__(self, *a, **k):\n"
        '        raise RuntimeError("ultralytics is not installed: YOLO pipelines are out of scope")\n')
    manifest = {str(p.relative_to(DST)): hashlib.sha256(p.read_bytes()).hexdigest()
                for p in sorted(DST.rglob("*.py"))}
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files from {REF_SRC}")
    return True


def import_reference():
    """``import openglottal`` from oracle/_ref (None when it has not been built)."""
    if not (DST / "openglottal" / "__init__.py").exists():
        return None
    if str(DST) not in sys.path:
        sys.path.insert(0, str(DST))
    import openglottal  # noqa: F401  (the reference package)

    return openglottal


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
