#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins the SASS page of an `ncu --set full --import-source on` report
(instructions executed + stall samples per SASS address) with the line table `nvdisasm -g` prints for the same kernel
in the built library.  The library must be the build the report was taken from (same SASS); the tool checks that the
instruction at every joined offset has the same mnemonic.

    python tools/line_profile.py report.ncu-rep kernel_regex [top_n] [--ranges file:lo-hi,...]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "implementing-jpeg-compression_b200", "libjpegb200.so")


def sass_page(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-name", "regex:" + kernel_regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = None
    table = []
    header = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if name is not None:
                break                                           # first matching launch only
            name = r[1]
        elif r and r[0] == "Address":
            header = r
        elif header and len(r) >= len(header) - 2 and r[0].startswith("0x"):
            table.append(r)
    ix = {h: i for i, h in enumerate(header)}
    base = int(table[0][0], 16)
    res = []
    for r in table:
        res.append((int(r[0], 16) - base, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0),
                    int(r[ix["# Samples"]] or 0)))
    return name, res


def line_table(kernel_name):
    """offset -> (file, line) for the kernel whose demangled name is kernel_name"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    want = re.sub(r"\(\w+\)", "", kernel_name.split("(CUtensorMap")[0].split("(Jb")[0])      # void f<1, 0>
    want = want.replace("void ", "").replace(" ", "")
    for cubin in sorted(os.listdir(tmp)):
        path = os.path.join(tmp, cubin)
        txt = subprocess.run(["nvdisasm", "-g", "-c", path], capture_output=True, text=True).stdout
        funcs = re.split(r"\n//-+ \.text\.(\S+) -+\n", txt)
        for i in range(1, len(funcs), 2):
            mangled = funcs[i]
            dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\((bool|int|unsigned int)\)", "", dem).replace("void ", "").replace(" ", "")
            dem = dem.replace("true", "1").replace("false", "0").split("(")[0]
            dem = re.sub(r"(\d)u\b", r"\1", dem)
            if dem != want:
                continue
            cur = ("?", 0)
            table = {}
            for line in funcs[i + 1].split("\n"):
                m = re.search(r'//## File "([^"]+)", line (\d+)', line)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
                if m:
                    table[int(m.group(1), 16)] = (cur, m.group(2).strip())
            return table
    raise SystemExit("kernel %r not found in %s" % (want, LIB))


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 40
    name, sass = sass_page(rep, regex)
    table = line_table(name)
    per_line = collections.Counter()
    samples = collections.Counter()
    mismatch = 0
    for off, src, n, s in sass:
        (loc, ins) = table.get(off, (("?", 0), ""))
        if ins.split(" ")[0].lstrip("@!P0123456789 ") != src.split(" ")[0].lstrip("@!P0123456789 ") and ins[:12] != src[:12]:
            mismatch += 1
        per_line[loc] += n
        samples[loc] += s
    total = sum(per_line.values())
    tot_s = sum(samples.values())
    print("%s\n%d warp instructions, %d samples, %d/%d SASS lines differ from the library (0 = same build)" %
          (name, total, tot_s, mismatch, len(sass)))
    src_cache = {}

    def text(loc):
        f, l = loc
        if f not in src_cache:
            p = os.path.join(ROOT, "implementing-jpeg-compression_b200", "csrc", f)
            src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        lines = src_cache[f]
        return lines[l - 1].strip()[:90] if 0 < l <= len(lines) else ""
    print("%-26s %9s %6s %6s  %s" % ("file:line", "warp-inst", "inst%", "smpl%", "source"))
    for loc, n in per_line.most_common(top):
        print("%-26s %9d %6.2f %6.2f  %s" % ("%s:%d" % loc, n, 100.0 * n / total, 100.0 * samples[loc] / max(tot_s, 1), text(loc)))
    for a in sys.argv[3:]:
        if a.startswith("--ranges"):
            spec = a.split("=", 1)[1] if "=" in a else sys.argv[sys.argv.index(a) + 1]
            for part in spec.split(","):
                f, r = part.split(":")
                lo, hi = map(int, r.split("-"))
                n = sum(v for (ff, ll), v in per_line.items() if ff == f and lo <= ll <= hi)
                s = sum(v for (ff, ll), v in samples.items() if ff == f and lo <= ll <= hi)
                print("range %-30s %9d %6.2f%% inst %6.2f%% samples" % (part, n, 100.0 * n / total, 100.0 * s / max(tot_s, 1)))


if __name__ == "__main__":
    main()
