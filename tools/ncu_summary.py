"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary json + markdown kept under profiles/.
    ncu -i gpurun_out/r02_hot.ncu-rep --page raw --csv | python tools/ncu_summary.py profiles/r02_hot_kernels_ncu"""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale=1.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3,
            "Tbyte/second": 1e12, "Gbyte/second": 1e9}.get(u, 1.0)
    return v * mult * scale
out, seen = [], {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("da::", "")
    k = seen.get(short, 0); seen[short] = k + 1
    stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for h, i in col.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "n/a")}
    tot = sum(stalls.values()) or 1.0
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out.append({
        "kernel": short, "instance": k, "ncu_us": val(r, "gpu__time_duration.sum"),
        "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]], "regs": r[col["launch__registers_per_thread"]],
        "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes": None if rd is None else rd + wr,
        "l2_to_sm_bytes": val(r, "l1tex__m_xbar2l1tex_read_bytes.sum"), "sm_to_l2_bytes": val(r, "l1tex__m_l1tex2xbar_write_bytes.sum"),
        "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "top_stalls_pct": {k_: round(100 * v / tot, 1) for k_, v in top},
    })
base = sys.argv[1]
json.dump({"source": "ncu --set full --clock-control none --import-source on, tools/prof_r02.py (warm launches at the step's sizes), B200, round 2",
           "launches": out}, open(base + ".json", "w"), indent=1)
with open(base + ".md", "w") as f:
    f.write("# r02 — `ncu --set full` of the hot kernels at the step's sizes (tools/prof_r02.py)\n\n")
    f.write("| kernel | # | us (cold, serialised) | grid x block, regs | DRAM rd / wr MB | L2->SM MB | SM->L2 MB | L2 hit % | tensor pipe % | top stalls |\n|---|---|---|---|---|---|---|---|---|---|\n")
    mb = lambda v: "-" if v is None else f"{v / 1e6:.1f}"
    pc = lambda v: "-" if v is None else f"{v:.1f}"
    for o in out:
        f.write(f"| `{o['kernel']}` | {o['instance']} | {pc(o['ncu_us'])} | {o['grid']} x {o['block']}, {o['regs']} | {mb(o['dram_read_bytes'])} / {mb(o['dram_write_bytes'])} | "
                f"{mb(o['l2_to_sm_bytes'])} | {mb(o['sm_to_l2_bytes'])} | {pc(o['l2_hit_pct'])} | {pc(o['tensor_pipe_active_pct'])} | "
                f"{', '.join(f'{k_} {v}' for k_, v in o['top_stalls_pct'].items())} |\n")
print(f"{len(out)} launches summarised")
