"""Per-kernel table from an ncu launch list taken with
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv
Usage: python tools/ncu_dram_summary.py gpurun_out/launches.csv profiles/r02_ncu_dram_per_kernel  (-> .txt and .json)
Template instances of one kernel are merged under the bare name (`dl::igemm_kernel`); bench.py reads the JSON for
`roofline.traffic` (DRAM bytes per igemm launch) when the launch count matches its own."""
import collections
import csv
import io
import json
import re
import sys

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
UNIT_T = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for row in csv.DictReader(io.StringIO("".join(lines))):
    name = re.sub(r"^void ", "", row["Kernel Name"].split("(")[0])
    name = re.sub(r"<.*", "", name)
    m, v, u = row.get("Metric Name"), float(row["Metric Value"].replace(",", "")), row.get("Metric Unit", "")
    if m == "gpu__time_duration.sum":
        agg[name]["ms"] += v * UNIT_T.get(u, 1e-6)
        cnt[name] += 1
    elif m == "dram__bytes_read.sum":
        agg[name]["dram_read_bytes"] += v * UNIT_B.get(u, 1.0)
    elif m == "dram__bytes_write.sum":
        agg[name]["dram_write_bytes"] += v * UNIT_B.get(u, 1.0)
    elif m == "lts__t_bytes.sum":
        agg[name]["l2_bytes"] += v * UNIT_B.get(u, 1.0)
res = {k: dict(launches=cnt[k], **{f: agg[k].get(f, 0.0) for f in ("ms", "dram_read_bytes", "dram_write_bytes", "l2_bytes")})
       for k in agg}
json.dump(res, open(out + ".json", "w"), indent=1)
tot = sum(v["ms"] for v in res.values())
with open(out + ".txt", "w") as f:
    f.write(f"# {path}: {sum(cnt.values())} launches, {tot:.2f} ms (ncu: cold-cache, serialised - compare shares)\n")
    f.write(f"{'kernel':40s} {'launches':>8s} {'ms':>9s} {'share':>6s} {'DRAM rd GB':>11s} {'DRAM wr GB':>11s} {'L2 GB':>9s} {'DRAM GB/s':>10s}\n")
    for k, v in sorted(res.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = (v["dram_read_bytes"] + v["dram_write_bytes"]) / (v["ms"] * 1e-3) / 1e9 if v["ms"] else 0.0
        f.write(f"{k[:40]:40s} {v['launches']:8d} {v['ms']:9.3f} {100 * v['ms'] / tot:5.1f}% {v['dram_read_bytes'] / 1e9:11.3f} "
                f"{v['dram_write_bytes'] / 1e9:11.3f} {v['l2_bytes'] / 1e9:9.2f} {gbs:10.0f}\n")
print(open(out + ".txt").read())
