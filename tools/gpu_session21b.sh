#!/bin/bash
# GPU session 21b (round 2): parallel VM attempts -- deterministic? limits reported? (diagnostic), then the chain tests again
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s21b
mkdir -p $O
timeout -s KILL 150 python tools/vm_par_diag.py > $O/vm_par_diag.txt 2>&1; echo "diag rc=$?"; tail -25 $O/vm_par_diag.txt | cut -c1-400
timeout -s KILL 200 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -s -k "chain_resolve" > $O/pytest_chain.txt 2>&1; echo "chain rc=$?"; grep -a "one 64 MiB\|one 8 MiB\|passed\|failed\|Error\|assert" $O/pytest_chain.txt | cut -c1-220 | head -30
