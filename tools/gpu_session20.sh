#!/bin/bash
# GPU session 20 (round 2): chain path for class runs in LINE mode (entry positions) and for general patterns with start-free
# attempts (one VM attempt per candidate + chain): forced on small inputs, chosen on one 64 MiB / 1 GiB unit, and forced under
# the whole random-pattern differential and the known-answer tests.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s20
mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -s -k "chain_resolve" > $O/pytest_chain.txt 2>&1; echo "chain rc=$?"; grep -a "one 64 MiB\|passed\|failed\|Error\|assert" $O/pytest_chain.txt | cut -c1-220 | head -30
timeout -s KILL 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s -k "one_gib" > $O/pytest_big_unit.txt 2>&1; echo "big rc=$?"; grep -a "one 1 GiB unit\|passed\|failed\|Error" $O/pytest_big_unit.txt | cut -c1-220
GSCAN_CHAIN=1 timeout -s KILL 420 python -m pytest tests/test_gpu_random_patterns.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_forced.txt 2>&1; echo "forced rc=$?"; tail -4 $O/pytest_forced.txt | cut -c1-220
