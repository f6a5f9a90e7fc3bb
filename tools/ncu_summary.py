"""Summarise an .ncu-rep (first kernel): key metrics, stall reasons, opcode mix.  usage: ncu_summary.py rep [out.txt]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
def run(*a): return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, unit, val = raw[0], raw[1], raw[2]
d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
print("kernel:", d.get("Kernel Name", ("?",))[0], file=out)
for k in keys:
    if k in d: print(f"{k:70s} {d[k][0]:>18s} {d[k][1]}", file=out)
print("\nstall reasons (warps per issue-active cycle):", file=out)
st = []
for k, (v, u) in d.items():
    m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
    if m:
        try: st.append((float(v.replace(",", "")), m.group(1)))
        except ValueError: pass
for v, n in sorted(st, reverse=True)[:10]: print(f"  {n:30s} {v:8.3f}", file=out)
src = list(csv.reader(io.StringIO(run("--page", "source", "--csv"))))
h = src[1]; ix = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); tot = 0; n_sass = 0
for r in src[2:]:
    if len(r) < len(h): continue
    n_sass += 1
    s = r[ix["Source"]].strip(); n = int(r[ix["Instructions Executed"]] or 0)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    ops[m.group(2).split(".")[0] if m else s[:8]] += n; tot += n
print(f"\nSASS instructions in kernel: {n_sass}; executed warp instructions: {tot}", file=out)
for op, n in ops.most_common(24): print(f"  {op:10s} {n:>14d} {100*n/max(tot,1):5.1f}%", file=out)
