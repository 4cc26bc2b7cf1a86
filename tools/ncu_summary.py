#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (one launch) into the metric list kept under profiles/.
  tools/ncu_summary.py raw.csv "title line" > profiles/rNN_ncu_summary.txt"""
import csv
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum dram__bytes_write.sum.per_second
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio launch__registers_per_thread launch__grid_size launch__block_size
launch__shared_mem_per_block_dynamic sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
l1tex__t_requests_pipe_lsu_mem_global_op_st.sum lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed smsp__warps_eligible.avg.per_cycle_active""".split()
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print(sys.argv[2] if len(sys.argv) > 2 else "# ncu --set full --clock-control none")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== {name}")
    got = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
    for w in WANT:
        if w in got:
            print(f"{w:100s} {got[w][0]:>18s} {got[w][1]}")
    for h in sorted(got):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            print(f"{h:100s} {got[h][0]:>18s} {got[h][1]}")
    try:
        st = float(got["l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"][0])
        wr = float(got["dram__bytes_write.sum"][0])
        u = got["dram__bytes_write.sum"][1]
        wr *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        print(f"{'derived: global store sectors / (dram bytes written / 32)':100s} {st / (wr / 32.0):18.3f}")
    except (KeyError, ValueError):
        pass
