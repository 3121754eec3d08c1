"""Summarise an ncu report (ncu --set full ... -o X) into a text file for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt ["note"]
Reads the report on the CPU box with `ncu -i ... --page raw/source --csv`."""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
L = [f"# ncu summary of {rep}", f"# {note}", ""]
for r in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
    L.append(f"== launch: {d.get('Kernel Name', ('', '?'))[1]}  grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
    for k in KEYS:
        if k in d:
            L.append(f"{k:78s} {d[k][0]:16s} {d[k][1]}")
    L.append("-- warp stall reasons (warps per issue-active cycle)")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and float(d[h][1] or 0) > 0.01:
            L.append(f"{h:78s} {d[h][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 3:
    sh, data = srows[1], [r for r in srows[2:] if len(r) == len(srows[1])]
    iex, ismp, isrc = sh.index("Instructions Executed"), sh.index("# Samples"), sh.index("Source")
    tot, tots = sum(int(r[iex]) for r in data), sum(int(r[ismp]) for r in data)
    mx = max(int(r[iex]) for r in data)
    L += ["", f"-- SASS regions by execution count (first launch): {len(data)} SASS instructions, {tot} warp-instructions, {tots} samples"]
    for name, cond in (("scan loop (exec > 50% of max)", lambda e: e > 0.5 * mx), ("per-ray code (2%..50%)", lambda e: 0.02 * mx < e <= 0.5 * mx), ("cold (<2%)", lambda e: e <= 0.02 * mx)):
        sel = [r for r in data if cond(int(r[iex]))]
        ops = collections.Counter()
        for r in sel:
            t = r[isrc].split(); op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += int(r[iex])
        L.append(f"{name:34s} static {len(sel):5d}  dyn {sum(int(r[iex]) for r in sel) / tot:6.3f}  samples {sum(int(r[ismp]) for r in sel) / max(tots, 1):6.3f}  top ops "
                 + ", ".join(f"{k} {v / tot:.3f}" for k, v in ops.most_common(8)))
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L[:60]))
