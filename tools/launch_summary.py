"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel, launches and median microseconds."""
import collections
import csv
import statistics
import sys


def summarise(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    return agg


if __name__ == "__main__":
    for path in sys.argv[1:]:
        print(path)
        for k, v in summarise(path).items():
            print("  %-70s n=%3d median %9.1f us  last %9.1f us" % (k, len(v), statistics.median(v), v[-1]))
