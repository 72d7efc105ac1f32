#!/usr/bin/env python
"""Summarise an .ncu-rep: key throughput metrics, stall reasons, executed instruction mix and the hottest SASS lines.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 15
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__occupancy_limit_registers|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed|sm__cycles_elapsed.max|smsp__thread_inst_executed.sum|sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active|l1tex__data_pipe_lsu_wavefronts.sum|l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum|l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum)$"
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for i, h in enumerate(hdr):
        if re.search(keys, h): print(f"  {h:75s} {r[i]:>16s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m:
            try: st.append((float(r[i]), m.group(1)))
            except ValueError: pass
    print("  stalls/issue:", "  ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    hdr = rows[hi[0]]; ia = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
    data = []
    for r in rows[hi[0] + 1:]:
        try: data.append((int(r[ia]), int(r[ie]), r[isrc].strip()))
        except (ValueError, IndexError): pass
    tot = sum(d[0] for d in data) or 1; totex = sum(d[1] for d in data) or 1
    mix = collections.Counter(); smp = collections.Counter()
    for s, e, t in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]; mix[op] += e; smp[op] += s
    print(f"  executed warp-instr {totex}; mix:", "  ".join(f"{op}:{100*c/totex:.1f}%" for op, c in mix.most_common(16)))
    print("  hottest instructions (share of stall samples):")
    for s, e, t in sorted(data, reverse=True)[:ntop]: print(f"    {100*s/tot:5.2f}%  ex={e:9d}  {t[:90]}")
