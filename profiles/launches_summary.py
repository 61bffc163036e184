"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown).
Usage: python profiles/launches_summary.py launches.csv "command line" > out.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = r[ik].split("(")[0][:150]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"# ncu launch list of `{sys.argv[2]}` (per-launch times are cold-cache and serialised: compare SHARES)")
print(f"total device time {T / 1e3:.3f} ms over {sum(cnt.values())} launches\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / T:.1f}% |")
