#!/usr/bin/env python
"""Put an UNMODIFIED copy of the reference where the GPU box can see it.

    python tools/install_reference.py [--src /root/reference]

The reference (X-rayLaser/Implementing-JPEG-compression) is plain Python without a setup.py / pyproject, so
there is nothing for pip to install: this recipe copies its source tree verbatim to ``baseline/_ref/reference/``
and writes ``baseline/_ref/MANIFEST.json`` (sha256 per file).  ``baseline/_ref/`` is git-ignored (reference
sources never enter this repository's history) but is NOT gpurun-ignored, so the copy travels to the GPU box
with the snapshot, like the built ``.so`` files.  There it serves

* ``bench.py --impl reference`` and the ``cpu_baseline`` leg (``kind: "reference"``): the reference's own
  ``compress_band`` / ``decompress_band`` (pipeline/__init__.py:71-88) timed on the box's host cores;
* ``tests/test_reference_dropin.py`` (``-m gpu``): the stock ``Jpeg.decompress`` decodes the GPU's streams, and
  the reference's own integration tests run over the C-ABI stub of INTEGRATION.md.

The two shims the reference needs in this image (no ``bitarray`` wheel, numpy >= 1.24 without ``np.float``)
stay in ``oracle/_shim`` / ``oracle/load_reference.py``; the copy itself is byte-identical to the source.
``__graft_entry__.build()`` runs this whenever the source tree is present.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(src="/root/reference", dest=DEST, quiet=False):
    """Copy `src` to `dest`/reference; returns the number of files, or 0 when `src` is absent."""
    if not os.path.isfile(os.path.join(src, "pipeline", "__init__.py")):
        return 0
    target = os.path.join(dest, "reference")
    if os.path.isdir(target):
        shutil.rmtree(target)
    os.makedirs(dest, exist_ok=True)
    shutil.copytree(src, target, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(target):
        for name in sorted(files):
            path = os.path.join(base, name)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, target)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print("installed %d reference files into %s" % (len(manifest), target))
    return len(manifest)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    args = ap.parse_args()
    sys.exit(0 if install(args.src) else 1)
