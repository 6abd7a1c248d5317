"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/launch_list_summary.py gpurun_out/launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<unnamed>::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    tot[name] += v_us
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k] / 1000:10.3f} {tot[k] / cnt[k]:9.1f} {100 * tot[k] / total:6.1f}%")
print(f"{'TOTAL':70s} {sum(cnt.values()):8d} {total / 1000:10.3f}")
