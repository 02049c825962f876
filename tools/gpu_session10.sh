#!/bin/bash
# GPU session 10 (round 2): first run of the sparse hashed path, every step under its own hard timeout.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s10
mkdir -p $O
timeout -s KILL 150 python tools/hash_check.py > $O/hash_check.txt 2>&1; echo "hash_check rc=$?"; tail -25 $O/hash_check.txt | cut -c1-200
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout -s KILL 150 python tools/kbench.py --gib 8 --only lits100,lits100_16k,lits8 --label sparse > $O/kbench_sparse.jsonl 2>$O/kbench_sparse.err; echo "kbench rc=$?"; cat $O/kbench_sparse.jsonl; tail -3 $O/kbench_sparse.err
GSCAN_HASH_PRE=0 timeout -s KILL 150 python tools/kbench.py --gib 8 --only lits100,lits100_16k --label dense > $O/kbench_dense.jsonl 2>$O/kbench_dense.err; echo "kbench rc=$?"; cat $O/kbench_dense.jsonl
