"""Summarise `ncu -i X.ncu-rep --page raw --csv` output as markdown: one column per captured launch.
usage: python tools/ncu_summary.py raw.csv [title] > profiles/NAME.md"""
import csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("cluster", "launch__cluster_size"),
    ("registers/thread", "launch__registers_per_thread"),
    ("dynamic smem/CTA", "launch__shared_mem_per_block_dynamic"),
    ("CTAs/SM limit (regs / smem / warps)", None),
    ("achieved warps per scheduler", "smsp__warps_active.avg.per_cycle_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("issue slots busy %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("FMA-heavy pipe (IMAD) busy % of elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("ALU pipe % of peak", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe busy % of elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM throughput %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("L1 hit rate %", "l1tex__t_sector_hit_rate.pct"),
    ("local (spill) load+store instr", None),
    ("smem bank-conflict wavefronts (ld / st)", None),
]
STALLS = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)]


def val(r, name):
    if name not in ix:
        return "n/a"
    v, u = r[ix[name]], units[ix[name]]
    try:
        f = float(v.replace(",", ""))
        v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.4e}"
    except ValueError:
        pass
    return f"{v} {u}".strip()


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("hb::", "")


print(f"# {sys.argv[2] if len(sys.argv) > 2 else 'ncu --set full summary'}\n")
print("Times under ncu are cold-cache and serialised: never bench values.\n")
names = [short(r[ix["Kernel Name"]]) for r in data]
print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|" + "---|" * len(names))
for label, m in METRICS:
    cells = []
    for r in data:
        if label.startswith("CTAs/SM"):
            cells.append(" / ".join(val(r, k).split()[0] for k in ("launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps")))
        elif label.startswith("local"):
            cells.append(" + ".join(val(r, k).split()[0] for k in ("smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum")))
        elif label.startswith("smem bank"):
            cells.append(" / ".join(val(r, k).split()[0] for k in ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum")))
        else:
            cells.append(val(r, m))
    print(f"| {label} | " + " | ".join(cells) + " |")
print("\nWarp stall reasons (warps stalled per issued instruction):\n")
print("| reason | " + " | ".join(f"`{n[:28]}`" for n in names) + " |")
print("|---|" + "---|" * len(names))
order = sorted(STALLS, key=lambda h: -max(float(r[ix[h]] or 0) for r in data))
for h in order[:10]:
    reason = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active", h).group(1)
    print(f"| {reason} | " + " | ".join(f"{float(r[ix[h]] or 0):.2f}" for r in data) + " |")
