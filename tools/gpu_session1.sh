#!/bin/bash
# GPU session 1 (round 2): kernel numbers for every BASELINE shape with the new engines and their tuning variants,
# the GPU parity suite, and a first run of the multi-config bench.  Output: gpurun_out/s1/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python tools/kbench.py --gib 16 --label default > $O/kbench_default.jsonl 2> $O/kbench_default.err
GSCAN_NO_FIXEDB=1 timeout 300 python tools/kbench.py --gib 16 --only alt4 --label nofixedb > $O/kbench_nofixedb.jsonl 2>> $O/kbench_default.err
for v in fbA1 fbA3 pair16 pair24; do
  GSCAN_LIB=$PWD/grab_b200/libgscan_$v.so timeout 300 python tools/kbench.py --gib 16 --only alt4 --label $v > $O/kbench_$v.jsonl 2>> $O/kbench_default.err
done
for v in hsub0 hsub4; do
  GSCAN_LIB=$PWD/grab_b200/libgscan_$v.so timeout 300 python tools/kbench.py --gib 16 --only lits100,lits100_16k,lits8 --label $v > $O/kbench_$v.jsonl 2>> $O/kbench_default.err
done
cat $O/kbench_*.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
GSCAN_TILE_SHIFT=12 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_tile12.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_tile12.txt
tail -3 $O/pytest_tile12.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 6000 $O/bench.json; tail -20 $O/bench.err
