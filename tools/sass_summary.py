#!/usr/bin/env python
"""Per-kernel SASS evidence of the built libjpegb200.so (cuobjdump -sass): which kernels use TMA (UTMALDG / UTMASTG),
mbarriers (SYNCS), byte-SIMD dot products (IDP.4A), 128-bit global accesses, programmatic dependent launch, fp64, and how
large they are.  Writes a text table; no GPU needed.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "implementing-jpeg-compression_b200", "libjpegb200.so")
CSRC = os.path.join(ROOT, "implementing-jpeg-compression_b200", "csrc")

PATTERNS = collections.OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"),
    ("IDP.4A", r"\bIDP\.4A"), ("LDG.128", r"\bLDG\.E\.(?:\w+\.)*128"), ("STG.128", r"\bSTG\.E\.(?:\w+\.)*128"),
    ("LDS.128", r"\bLDS\.128"), ("SHFL", r"\bSHFL"), ("ATOMS/RED", r"\b(?:ATOMS|ATOMG|REDG|RED)\b"),
    ("FFMA", r"\bFFMA"), ("DFMA/DADD/DMUL", r"\b(?:DFMA|DADD|DMUL)"), ("PDL(ACQBULK/PREEXIT)", r"\b(?:ACQBULK|PREEXIT)"),
    ("LDL/STL", r"\b(?:LDL|STL)\b"),
])


def build_id():
    """sha256 over the kernel sources (names + contents): identifies the code a profile belongs to."""
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h", ".cpp")) or name == "Makefile":
            h.update(name.encode())
            with open(os.path.join(CSRC, name), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "jpegb200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:12]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["instructions"] += 1
            for key, pat in PATTERNS.items():
                if re.search(pat, line):
                    kernels[cur][key] += 1
    demangled = {}
    try:
        names = list(kernels)
        res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(names, res))
    except OSError:
        pass
    print("libjpegb200.so  build id %s (tools/sass_summary.py: sha256 of csrc/ + include/)  arch sm_100a" % build_id())
    cols = ["instructions"] + list(PATTERNS)
    print("%-78s %s" % ("kernel", " ".join("%9s" % c[:9] for c in cols)))
    totals = collections.Counter()
    for name, cnt in kernels.items():
        short = re.sub(r"\(.*", "", demangled.get(name, name))[:78]
        print("%-78s %s" % (short, " ".join("%9d" % cnt[c] for c in cols)))
        totals.update(cnt)
    print("%-78s %s" % ("TOTAL (%d kernels)" % len(kernels), " ".join("%9d" % totals[c] for c in cols)))


if __name__ == "__main__":
    main()
