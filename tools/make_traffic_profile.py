#!/usr/bin/env python
"""profiles/r2_traffic.json from two `ncu --set full` captures of the fused kernels (bench.py --images N --no-graph):
DRAM bytes read + written per launch, per image, together with the build id of the kernel sources they were
measured on.  bench.py puts them into roofline.traffic and refuses them once the sources have changed.

    python tools/make_traffic_profile.py fwd.ncu-rep inv.ncu-rep N_IMAGES
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_summary  # noqa: E402


def raw(rep, rx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    fwd_rep, inv_rep, n_img = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rec = {"build_id": sass_summary.build_id(), "images": n_img,
           "how": "ncu --set full --clock-control none, one launch of each fused kernel, %d x 1920x1080 images" % n_img}
    for key, rep, rx in (("fwd", fwd_rep, "jb_fwd_fast"), ("inv", inv_rep, "jb_inv_fast")):
        m = raw(rep, rx)
        rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
        rec[key + "_dram_bytes_read"] = rd
        rec[key + "_dram_bytes_write"] = wr
        rec[key + "_dram_bytes_per_image"] = (rd + wr) / n_img
        rec[key + "_kernel_us_under_ncu"] = float(m["gpu__time_duration.sum"][0].replace(",", "")) * \
            {"ns": 1e-3, "us": 1.0, "ms": 1e3}[m["gpu__time_duration.sum"][1]]
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    json.dump(rec, open(path, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
