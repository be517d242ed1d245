"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import sys
from collections import defaultdict


def main(path, note=""):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v   # -> us
        tot[r[ki][:110]] += v
        cnt[r[ki][:110]] += 1
    S = sum(tot.values())
    print(f"# {note}")
    print(f"# {sum(cnt.values())} launches, {S / 1e3:.2f} ms of device time (cold-cache, serialised under ncu: compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v / 1e3:10.3f} ms {100 * v / S:5.1f}% n={cnt[k]:4d} {k}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
