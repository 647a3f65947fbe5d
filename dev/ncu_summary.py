#!/usr/bin/env python3
"""dev/ncu_summary.py <raw.csv> <tag> [directions] — condense one `ncu --page raw --csv` capture of k_maxmin_bp into
profiles/<tag>_ncu_summary_k_maxmin_bp.txt and refresh profiles/traffic.json (read by bench.py for roofline.traffic)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, tag = sys.argv[1], sys.argv[2]
ndir = int(sys.argv[3]) if len(sys.argv) > 3 else 4194304
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def f(name):
    return float(m[name][0].replace(",", ""))
def unit_bytes(name):
    v, u = m[name]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u.split("/")[0]]
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
]
keys += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
out = [f"{k} = {m[k][0]} {m[k][1]}" for k in keys if k in m]
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary_k_maxmin_bp.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
dram = unit_bytes("dram__bytes_read.sum") + unit_bytes("dram__bytes_write.sum")
pct = lambda k: round(f(k), 1) if k in m else None
tj = {
    "config": "cfg3", "dram_bytes_per_launch": dram, "directions_per_launch": ndir,
    "pipe_utilisation": {
        "alu_pct": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "lsu_pct": pct("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "fma_pct": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "fp64_pct": pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_slots_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions_per_direction": round(f("smsp__inst_executed.sum") / ndir),
        "registers": int(f("launch__registers_per_thread")),
        "smem_kb_per_cta": round(unit_bytes("launch__shared_mem_per_block_dynamic") / 1024, 1) if "launch__shared_mem_per_block_dynamic" in m else None,
    },
    "source": f"profiles/{tag}_ncu_summary_k_maxmin_bp.txt: dram__bytes_read.sum + dram__bytes_write.sum and pipe counters of the fused "
              f"k_maxmin_bp<10,2,GEN> launch, ncu --set full, {ndir} mid-lattice directions x 10^4 points (algorithmic bytes 0.8 MB "
              "for the point set; directions are generated in the kernel, the per-CTA scratch, pair and hit lists stay in L2)",
}
json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
