#!/bin/bash
# GPU session 14 (round 2): sparse hashed path with the tail-table verification.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s14
mkdir -p $O
timeout -s KILL 150 python tools/hash_check.py > $O/hash_check.txt 2>&1; echo "hash_check rc=$?"; tail -3 $O/hash_check.txt | cut -c1-200
timeout -s KILL 150 python tools/kbench.py --gib 8 --only lits100,lits100_16k,lits8 --label sparse > $O/kbench_sparse.jsonl 2>$O/kbench_sparse.err; echo "kbench rc=$?"; cat $O/kbench_sparse.jsonl; tail -3 $O/kbench_sparse.err
timeout -s KILL 400 bash tools/ncu_summary.sh lits100_sparse @lits100 scan_kernel $O
head -70 $O/ncu_lits100_sparse.txt
