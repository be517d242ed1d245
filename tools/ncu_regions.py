"""Instruction and stall-sample shares per source REGION (file + line range) of one kernel in an .ncu-rep.
usage: ncu_regions.py rep kernel-regex file:lo-hi=name ..."""
import csv
import subprocess
import sys


def main(path, kernel, specs):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    regions = []
    for s in specs:
        loc, name = s.split("=")
        f, rng = loc.split(":")
        lo, hi = rng.split("-")
        regions.append((f, int(lo), int(hi), name))
    acc = {r[3]: [0, 0] for r in regions}
    acc["(other)"] = [0, 0]
    fname, hdr = "", None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit() or r[2] != "-":
            continue
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            n, s, line = int(r[ii]), int(r[si]), int(r[0])
        except ValueError:
            continue
        for f, lo, hi, name in regions:
            if fname == f and lo <= line <= hi:
                acc[name][0] += n; acc[name][1] += s
                break
        else:
            acc["(other:%s)" % fname] = [acc.get("(other:%s)" % fname, [0, 0])[0] + n, acc.get("(other:%s)" % fname, [0, 0])[1] + s]
    ti = sum(v[0] for v in acc.values()) or 1
    ts = sum(v[1] for v in acc.values()) or 1
    for k, v in sorted(acc.items(), key=lambda x: -x[1][0]):
        if v[0] or v[1]:
            print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% samples  {k}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
