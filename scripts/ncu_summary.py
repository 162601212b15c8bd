#!/usr/bin/env python3
"""Prints the key metrics of an .ncu-rep (raw page) and the top stall lines of the source page."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_tensor.sum','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic',
 'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ,
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']
for r in rows[2:3]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w or h.endswith("." + w):
                print(f"{w:90s} {r[i]} {units[i]}")
                break
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if rows:
        h = rows[0]
        def col(name):
            for i, x in enumerate(h):
                if x.strip() == name: return i
            return None
        ci, ss, si = col("# Samples") or col("Warp Stall Sampling (All Samples)"), col("Source"), col("Address")
        samp = None
        for i, x in enumerate(h):
            if "Sampl" in x and "All" in x: samp = i; break
        if samp is None:
            for i, x in enumerate(h):
                if "Sampl" in x: samp = i; break
        print("columns:", h[:12], "...", "sampling col:", h[samp] if samp is not None else None)
        body = []
        for r in rows[1:]:
            try: v = float(r[samp].replace(',', ''))
            except Exception: continue
            body.append((v, r[ss] if ss is not None else r[1]))
        tot = sum(v for v, _ in body) or 1
        for v, t in sorted(body, reverse=True)[:int(sys.argv[2])]:
            print(f"{v:8.0f} {100*v/tot:5.1f}%  {t[:150]}")
