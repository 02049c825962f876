#!/bin/bash
# GPU session 16 (round 2): final sparse hashed path -- parity, kernel GB/s, text the class test cannot thin out, command-line bench.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s16
mkdir -p $O
timeout -s KILL 100 python tools/hash_check.py > $O/hash_check.txt 2>&1; echo "hash_check rc=$?"; tail -1 $O/hash_check.txt
timeout -s KILL 150 python tools/kbench.py --gib 16 --only lits100,lits100_16k,lits8 --label sparse > $O/kbench_sparse.jsonl 2>$O/kbench_sparse.err; echo "kbench rc=$?"; cat $O/kbench_sparse.jsonl
timeout -s KILL 200 python tools/hash_mix_bench.py > $O/hash_mix.jsonl 2>$O/hash_mix.err; echo "mix rc=$?"; cat $O/hash_mix.jsonl; tail -3 $O/hash_mix.err
timeout -s KILL 400 python tools/cli_bench.py 8192 1 > $O/cli_bench.txt 2>&1; echo "cli rc=$?"; cat $O/cli_bench.txt | cut -c1-220
