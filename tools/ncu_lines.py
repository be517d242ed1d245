"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    items, fname, hdr = [], "", None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0].isdigit() and r[2] == "-":   # a CUDA-C line (aggregated over its SASS)
            ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            try:
                items.append((int(r[ii]), int(r[si]), fname, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot_i = sum(i[0] for i in items) or 1
    tot_s = sum(i[1] for i in items) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    for n, s, f, l, src in sorted(items, reverse=True)[:top]:
        print(f"{100 * n / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% smpl  {f}:{l}: {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
