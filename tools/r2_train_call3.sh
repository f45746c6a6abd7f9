#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/train_call3.txt
: > $out
echo "== bench --workload train (micro-batch 4, 640x512)" >> $out
timeout 600 python bench.py --workload train --steps 3 --warmup 3 2>&1 | tail -3 >> $out
echo "== per-symbol time shares of one step" >> $out
timeout 600 python tools/train_profile.py 2>&1 | tail -60 >> $out
echo "== ncu launch list of one step" >> $out
EVENTS=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py > gpurun_out/train_ncu.log 2>&1
python - <<'PY' >> $out 2>&1
import csv, collections, re
rows = []
with open("gpurun_out/train_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    key = re.sub(r"\(.*", "", name)[:90]
    agg[key][0] += 1
    agg[key][1] += us
    tot += us
print(f"{sum(a[0] for a in agg.values())} launches, {tot/1000:.1f} ms of kernel time (ncu, serialised)")
for k, (n, us) in sorted(agg.items(), key=lambda t: -t[1][1])[:45]:
    print(f"  {us/1000:9.2f} ms {100*us/tot:5.1f} % {n:5d} x {k}")
lib = sum(us for k, (n, us) in agg.items() if "imagd" in k)
print(f"library (imagd::) share of kernel time: {100*lib/tot:.1f} %")
PY
cat $out
