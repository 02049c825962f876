#!/bin/bash
# GPU session 21 (round 2): dense general patterns on the chain path (attempts of all positions in parallel), shared per-unit
# VM budget, RUN / VM chain paths again after the budget change.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s21
mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -s -k "chain_resolve" > $O/pytest_chain.txt 2>&1; echo "chain rc=$?"; grep -a "one 64 MiB\|one 8 MiB\|passed\|failed\|Error\|assert" $O/pytest_chain.txt | cut -c1-220 | head -30
GSCAN_CHAIN=1 timeout -s KILL 200 python -m pytest tests/test_gpu_random_patterns.py -m gpu -q -x -k "0 or 3 or 7" > $O/pytest_forced.txt 2>&1; echo "forced rc=$?"; tail -3 $O/pytest_forced.txt | cut -c1-220
