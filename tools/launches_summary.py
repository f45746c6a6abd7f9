"""gpurun_out/launches_bN.csv (ncu launch list of one step) -> profiles/<name>.csv + .md share table."""
import collections
import csv
import re
import sys

src, name, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
lines = [l for l in open(src) if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
        rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]).replace("imagd::", "").replace("void ", ""), ns,
                     r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[2] for r in rows)
with open(f"profiles/{name}.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["launch_id", "kernel", "gpu__time_duration_ns", "grid", "block"])
    for r in rows:
        w.writerow([r[0], r[1], int(r[2]), r[3], r[4]])
agg = collections.defaultdict(lambda: [0, 0.0])
for _, k, ns, _, _ in rows:
    agg[k][0] += 1
    agg[k][1] += ns
with open(f"profiles/{name}.md", "w") as f:
    f.write(f"# ncu launch list of ONE denoising step — {note}\n\n")
    f.write("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/one_step.py` "
            "(tools/profile_launches.sh). Per-launch times are cold-cache and serialised: read the SHARES. "
            f"Raw rows: {name}.csv.\n\n")
    f.write(f"{len(rows)} launches, {tot / 1e6:.3f} ms of kernel time.\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {c} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.2f} |\n")
print(open(f"profiles/{name}.md").read())
