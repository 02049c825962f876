#!/bin/bash
# GPU session 11 (round 2): where the sparse hashed path spends its time -- ablation builds + one ncu capture.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s11
mkdir -p $O
for v in abl1 abl2 abl3; do
  GSCAN_LIB=$PWD/grab_b200/libgscan_$v.so timeout -s KILL 120 python tools/kbench.py --gib 8 --only lits100 --label $v --reps 4 --check-files 0 2>/dev/null | head -1 | tee -a $O/kbench_abl.jsonl
done
timeout -s KILL 400 bash tools/ncu_summary.sh lits100_sparse @lits100 HashEngine $O
head -60 $O/ncu_lits100_sparse.txt
