#!/bin/bash
# GPU session 6 (round 2): HASH variants, full GPU parity suite (bounded VM budgets), DRAM traffic per config by ncu,
# pageable feed after coalescing.  Output: gpurun_out/s6/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s6
mkdir -p $O
timeout 300 python tools/kbench.py --gib 16 --only lits100,lits100_16k,lits8,alt4,run16 --label default > $O/kbench_default.jsonl 2> $O/kbench.err
GSCAN_LIB=$PWD/grab_b200/libgscan_hpow2.so timeout 300 python tools/kbench.py --gib 16 --only lits100,lits100_16k,lits8 --label hpow2 > $O/kbench_hpow2.jsonl 2>> $O/kbench.err
cat $O/kbench_*.jsonl
timeout 1300 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -12 $O/pytest_gpu.txt
timeout 900 python tools/ncu_traffic.py 0 1 2 3 4 > $O/traffic.log 2>&1; tail -6 $O/traffic.log; cp gpurun_out/r02_traffic.json $O/ 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --only 1 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s6/bench.json'))
print('HEAD value',d['value'],'ms/step',d['ms_per_step'],'e2e',d.get('e2e'))
PY
du -sh gpurun_out
