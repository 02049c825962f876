#!/bin/bash
# GPU session 9 (round 2): hashed engine, sparse (class-prefiltered) path -- parity of both paths, kernel GB/s with and without.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s9
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_hash_sparse.py tests/test_zz_gpu_large_sets.py -m gpu -x -q > $O/pytest_hash.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_hash.txt | cut -c1-300
timeout 300 python tools/kbench.py --gib 16 --only lits100,lits100_16k,lits8 --label sparse > $O/kbench_sparse.jsonl 2>$O/kbench_sparse.err; cat $O/kbench_sparse.jsonl
GSCAN_HASH_PRE=0 timeout 300 python tools/kbench.py --gib 16 --only lits100,lits100_16k --label dense > $O/kbench_dense.jsonl 2>$O/kbench_dense.err; cat $O/kbench_dense.jsonl
