"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the total.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        rows.append((r[ik], us))
    rows = rows[skip:]
    agg, cnt = defaultdict(float), defaultdict(int)
    for k, us in rows:
        k = re.sub(r"<unnamed>::|\(.*", "", k)
        k = re.sub(r"void ", "", k)[:70]
        agg[k] += us; cnt[k] += 1
    tot = sum(agg.values())
    print(f"{len(rows)} launches, {tot / 1000:.3f} ms total")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
        print(f"{v / 1000:9.3f} ms {100 * v / tot:5.1f}%  x{cnt[k]:4d}  avg {v / cnt[k]:8.1f} us  {k}")


if __name__ == "__main__":
    main()
