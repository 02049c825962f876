#!/bin/bash
# GPU session 17 (round 2): hashed engine, straight-line first probe round (GS_HASH_NP = 2 / 3 / 4 positions per lane) against
# the looped one of session 16 (libgscan_base.so); command-line bench with the other BASELINE patterns.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s17
mkdir -p $O
timeout -s KILL 100 python tools/hash_check.py > $O/hash_check.txt 2>&1; echo "hash_check rc=$?"; tail -1 $O/hash_check.txt
for v in _base _np2 "" _np4; do
  GSCAN_LIB=$PWD/grab_b200/libgscan$v.so timeout -s KILL 100 python tools/kbench.py --gib 8 --only lits100,lits100_16k --label "lib$v" --reps 6 --check-files 3 2>/dev/null | head -2 | tee -a $O/kbench_np.jsonl
done
timeout -s KILL 500 python tools/cli_bench.py 8192 1 patterns > $O/cli_bench.txt 2>&1; echo "cli rc=$?"; cat $O/cli_bench.txt | cut -c1-220
