"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name."""
import collections
import csv
import io
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {s:.2f} ms total (cold-cache, serialised: compare shares)")
print(f"{'kernel':64s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:64]:64s} {cnt[k]:8d} {v:10.3f} {100 * v / s:6.1f}%")
