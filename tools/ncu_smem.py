#!/usr/bin/env python
"""Shared-memory / tensor-pipe counters per launch of an .ncu-rep (ncu --set full):  python tools/ncu_smem.py rep.ncu-rep
The counters VERDICT r1 asked for: LSU shared wavefronts (% of peak, ideal vs excessive), bank conflicts by op, the
data-bank read/write utilisation (which includes the tensor core's operand reads from shared memory), tensor pipe."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[0], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "derived__memory_l1_wavefronts_shared_excessive", "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_ideal.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
cols = [(w, hdr.index(w)) for w in want if w in hdr]
for r in data:
    print("; ".join(f"{w.split('.')[0] if w != 'Kernel Name' else 'kernel'}{'.pct' if 'pct' in w else ''}={r[i][:70]}" for w, i in cols))
