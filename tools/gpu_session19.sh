#!/bin/bash
# GPU session 19 (round 2): one 1 GiB unit with dense matches (flat / chain resolve against the oracle, scan and resolve times),
# command-line feed: staging-thread sweep, start-up phases, warm driver.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s19
mkdir -p $O
nvidia-smi --query-gpu=name,persistence_mode,pstate --format=csv,noheader
timeout -s KILL 420 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s -k "one_gib" > $O/pytest_big_unit.txt 2>&1; echo "pytest rc=$?"; grep -a "one 1 GiB unit\|passed\|failed\|Error" $O/pytest_big_unit.txt | cut -c1-220
timeout -s KILL 400 python tools/feed_bench.py 8192 > $O/feed_bench.txt 2>&1; echo "feed rc=$?"; cat $O/feed_bench.txt | cut -c1-260
