"""Turn an .ncu-rep of gp_lnlike_kernel into the committed summary under profiles/ (runs here, no GPU needed).
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_gp_lnlike.txt [profiles/traffic.json]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum", "sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
lines = ["# ncu --set full --clock-control none --import-source on : %s" % rep]
got = {}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        got[h] = (u, v)
        lines.append("%-90s %-14s %s" % (h, u, v))
def num(k):
    u, v = got[k]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
lines.append("dram traffic per launch (read+write) = %.4e bytes" % traffic)
# opcode mix of the stall samples (SASS page)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
from collections import Counter
c, tot = Counter(), 0
for r in srows[2:]:
    try:
        s = int(r[4])
    except Exception:
        continue
    ins = r[1].strip().split()
    op = ins[1] if ins and ins[0].startswith("@") else (ins[0] if ins else "?")
    c[op] += s
    tot += s
lines.append("warp stall samples by SASS opcode (top 12 of %d): %s" % (tot, ", ".join("%s %.1f%%" % (o, 100.0 * v / tot) for o, v in c.most_common(12))))
sass = " ".join(r[1] for r in srows[2:] if len(r) > 1)
lines.append("SASS evidence: DMMA.8x8x4 x%d, UBLKCP x%d, SYNCS (mbarrier) x%d" % (sass.count("DMMA.8x8x4"), sass.count("UBLKCP"), sass.count("SYNCS.")))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if len(sys.argv) > 3:
    json.dump({"dram_bytes_per_launch": traffic, "source": rep, "kernel": got.get("Kernel Name", ("", ""))[1]}, open(sys.argv[3], "w"))
