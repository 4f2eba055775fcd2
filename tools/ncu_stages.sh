#!/bin/bash
# usage: tools/ncu_stages.sh <rep> <kernel-substring>   -> key metrics + per-line/stage breakdown (needs the matching in-tree .so)
rep=$1; kern=${2:-extract_csr_kernelILb1}
ncu -i $rep --page raw --csv 2>/dev/null > /tmp/raw.csv
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.sum','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(w.replace('smsp__average_warps_issue_stalled_','stall_'), vals[i], units[i])
PY
ncu -i $rep --page source --csv --print-source sass > /tmp/sass.csv 2>/dev/null
( cd /tmp && rm -f *.cubin && cuobjdump -xelf all /root/repo/annealing-sign-problem_b200/libasp_b200.so >/dev/null 2>&1 && nvdisasm -c -g extract_fused.sm_100a.cubin > /tmp/kf.sass 2>/dev/null; rm -f /tmp/*.cubin )
python /root/repo/tools/ncu_by_line.py /tmp/sass.csv /tmp/kf.sass $kern extract_fused.cu ${3:-28}
