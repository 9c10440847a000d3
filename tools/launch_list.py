"""Post-process an `ncu --metrics gpu__time_duration.sum --csv` log of tools/step_once.py into the launch list of ONE step
(the last complete one: from the dropout-counter bump that opens a step to the next one).
    python tools/launch_list.py gpurun_out/launches_raw.csv profiles/r02_step_launches.csv"""
import csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
marks = [i for i, (k, _) in enumerate(rows) if "CUDAFunctorOnSelf_add<long>" in k]
a, b = (marks[-2], marks[-1]) if len(marks) >= 2 else (0, len(rows))
step = rows[a:b]
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "us"])
    for k, us in step:
        w.writerow([k, f"{us:.2f}"])
print(f"{len(step)} launches, {sum(u for _, u in step):.1f} us (serialised, cold cache)")
