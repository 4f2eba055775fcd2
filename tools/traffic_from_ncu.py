#!/usr/bin/env python
"""Record the DRAM traffic of one kernel launch from an `ncu --set full` report into profiles/traffic.json,
together with the hash of the kernel's sources (bench.py reports the figure only while the sources are unchanged).

    python tools/traffic_from_ncu.py <report.ncu-rep> <key> "<capture description>" <source.cu> [<source.cuh> ...]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import TRAFFIC_FILE, source_hash  # noqa: E402


def main():
    rep, key, capture, sources = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    totals, times = [], []
    for r in launches:
        b = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[name]]) * scale[units[col[name]]]
        totals.append(b)
        times.append(float(r[col["gpu__time_duration.sum"]]))
    entry = {"dram_bytes": int(sum(totals) / len(totals)), "launches_in_capture": len(totals), "kernel": launches[0][col["Kernel Name"]],
             "duration_under_ncu": "%s %s" % (sum(times) / len(times), units[col["gpu__time_duration.sum"]]),
             "capture": capture, "sources": sources, "source_sha256": source_hash(sources)}
    data = json.load(open(TRAFFIC_FILE)) if os.path.exists(TRAFFIC_FILE) else {}
    data[key] = entry
    json.dump(data, open(TRAFFIC_FILE, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(entry))


if __name__ == "__main__":
    main()
