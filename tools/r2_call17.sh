#!/bin/bash
# Evidence call: smoke(), ncu --set full of the training kernels, ncu launch lists (one inference step, one training step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/call17.txt
: > $out
echo "== smoke()" >> $out
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> $out
echo "== ncu --set full: training kernels at the level-0 shapes of configs[4]" >> $out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_train_kernels -f python tools/kernel_ncu3.py > gpurun_out/ncu_train.log 2>&1
echo "ncu rc=$?" >> $out
python tools/ncu_summary.py gpurun_out/r02_train_kernels.ncu-rep > gpurun_out/r02_ncu_train_kernels.txt 2>&1
grep -E "Kernel Name|time_duration|tensor_cycles|pipe_xu|dram__bytes" gpurun_out/r02_ncu_train_kernels.txt | cut -c1-150 >> $out
echo "== ncu launch list: one inference step (B=1)" >> $out
B=1 bash tools/profile_launches.sh >> $out 2>&1
wc -l gpurun_out/launches_b1.csv >> $out
echo "== ncu launch list: one training step" >> $out
EVENTS=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_profile.py > gpurun_out/train_ncu.log 2>&1
python - <<'PY' >> $out 2>&1
import csv, collections, re
lines = [l for l in open("gpurun_out/train_launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    key = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    agg[key][0] += 1; agg[key][1] += us; tot += us
print(f"{sum(a[0] for a in agg.values())} launches, {tot/1000:.1f} ms of kernel time (ncu, serialised)")
for k, (n, us) in sorted(agg.items(), key=lambda t: -t[1][1])[:40]:
    print(f"  {us/1000:9.2f} ms {100*us/tot:5.1f} % {n:5d} x {k}")
lib = sum(us for k, (n, us) in agg.items() if "imagd" in k)
print(f"library (imagd::) share of kernel time: {100*lib/tot:.1f} %")
PY
cat $out | cut -c1-220
