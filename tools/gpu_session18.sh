#!/bin/bash
# GPU session 18 (round 2): descriptor feed (GSCAN_UNIT_FD) -- parity through the C ABI and the command line, then the
# command line's wall time / steady-state rate with the descriptor and the mapped feed on the same 8 GiB tree.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s18
mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_shapes.py tests/test_cli.py -m gpu -x -q -k "descriptor or cli" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt | cut -c1-200
timeout -s KILL 300 python tools/feed_bench.py 8192 > $O/feed_bench.txt 2>&1; echo "feed rc=$?"; cat $O/feed_bench.txt | cut -c1-260
