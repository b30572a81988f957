"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean and total time per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, rows = r, rows[i + 1:]
            break
    ki, mi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) <= mi:
            continue
        try:
            v = float(r[mi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ki].split("(")[0][-48:], []).append(v)
    for k, v in agg.items():
        print(f"{k:48s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:10.1f} us total={sum(v) / 1e3:10.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
