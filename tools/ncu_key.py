"""Key metrics of an `ncu --page raw --csv` export, one line each.  usage: python tools/ncu_key.py raw.csv [ascans]"""
import csv, re, sys
r = list(csv.reader(open(sys.argv[1])))
h, u, v = r[0], r[1], r[2]
n = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576.0
want = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "datapipe%"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "all_wf"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("launch__registers_per_thread", "regs"),
        ("sass__inst_executed_local_loads", "LDL"), ("sass__inst_executed_local_stores", "STL"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
d = dict(zip(h, zip(u, v)))
out = []
for k, nm in want:
    if k in d:
        un, val = d[k]
        try:
            f = float(val.replace(",", ""))
        except ValueError:
            continue
        if nm in ("inst", "smem_wf", "all_wf", "bank_conf", "LDL", "STL"):
            out.append(f"{nm}/ascan={f / n:.1f}")
        else:
            out.append(f"{nm}={f:.4g}{un if un in ('ms','Gbyte','Mbyte') else ''}")
print("  ".join(out))
st = [(float(d[k][1]), k.split("issue_stalled_")[1].split("_per_")[0]) for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
print("stalls/issue: " + "  ".join(f"{n}={x:.2f}" for x, n in sorted(st, reverse=True)[:9]))
