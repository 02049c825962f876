#!/bin/bash
# GPU session 8 (round 2): chain-path diagnostic, smoke(), chain / async tests, command-line bench.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s8
mkdir -p $O
timeout 300 python tools/chain_diag.py > $O/chain_diag.txt 2>&1; cat $O/chain_diag.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -6 $O/smoke.txt | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "chain or async" > $O/pytest_chain.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_chain.txt | cut -c1-200
timeout 600 python tools/cli_bench.py 8192 1 > $O/cli_bench.txt 2>&1; cat $O/cli_bench.txt
