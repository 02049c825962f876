#!/bin/bash
# GPU session 15 (round 2): hashed engine, warps per CTA x table copies (more warps hide the probe chains; fewer copies free the shared memory for them)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s15
mkdir -p $O
timeout -s KILL 100 python tools/hash_check.py > $O/hash_check.txt 2>&1; echo "hash_check rc=$?"; tail -1 $O/hash_check.txt
for v in "" _w22c16 _w23c16 _w24c8; do
  for pre in 1 0; do
    GSCAN_HASH_PRE=$pre GSCAN_LIB=$PWD/grab_b200/libgscan$v.so timeout -s KILL 100 python tools/kbench.py --gib 8 --only lits100 --label "lib$v-pre$pre" --reps 5 --check-files 2 2>/dev/null | head -1 | tee -a $O/kbench_variants.jsonl
  done
done
