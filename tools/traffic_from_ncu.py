#!/usr/bin/env python3
"""profiles/roofline_traffic.json from a tools/ncu_summary.py text summary: DRAM bytes (read + write) per launch, averaged
over the captured launches of each kernel, keyed by the name the library's per-kernel timer uses (bench.py reads it)."""
import json, re, sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
NAMES = ["sort_query_kernel", "sketch2_kernel", "seed_lookup_kernel", "seed_count_kernel", "anchor_fill_kernel", "chain_group_kernel",
         "index_post_kernel", "compact_kernel", "mz_flt_kernel", "key32_kernel", "sketch_kernel", "pack_kernel", "owner_merge_kernel"]


def main(src, workload, out):
    acc = defaultdict(list)
    cur = None
    for line in open(src):
        if line.startswith("== "):
            cur = next((n for n in NAMES if re.search(r"\b" + n + r"\b", line)), None)
            if cur:
                acc[cur].append(0.0)
            continue
        m = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m and cur:
            acc[cur][-1] += float(m.group(2)) * UNIT[m.group(3)]
    doc = {"workload": workload,
           "source": f"{src} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches)",
           "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items()},
           "launches_captured": {k: len(v) for k, v in acc.items()}}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc["dram_bytes_per_launch"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "C2", sys.argv[3] if len(sys.argv) > 3 else "profiles/roofline_traffic.json")
