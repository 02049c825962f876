#!/bin/bash
# usage: tools/ncu_summary.sh <name> <pattern> <kernel regex> <outdir>  -- one `ncu --set full` capture of the scan kernel over an
# 8 GiB device-resident corpus; the (large) report stays in /tmp on the GPU box, the text summaries come back.
set -u
n=$1; p=$2; k=$3; O=$4
mkdir -p /tmp/ncu $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o /tmp/ncu/$n -f python tools/prof_one.py "$p" 8 2 > $O/ncu_$n.log 2>&1
python tools/ncu_top.py /tmp/ncu/$n.ncu-rep 40 > $O/ncu_$n.txt 2>&1
ncu -i /tmp/ncu/$n.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h,u,v=rows[0],rows[1],rows[2]
keep=('dram__bytes','gpu__time_duration','sm__inst_executed_pipe','smsp__inst_executed.sum','issue_active','l1tex__data_pipe','lsu_mem_shared','sm__warps_active','launch__','smsp__average_warp','stalled','sm__throughput','gpu__dram_throughput','l1tex__throughput','lts__throughput','smsp__cycles_active','sm__cycles')
for a,b,c in zip(h,u,v):
    if any(x in a for x in keep): print('%-100s %-14s %s'%(a,b,c))
" > $O/ncu_${n}_raw.txt 2>&1
ncu -i /tmp/ncu/$n.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/ncu_${n}_source.csv.gz
tail -1 $O/ncu_$n.log
