#!/bin/bash
# GPU session 4 (round 2): chunked candidate reservation, dense VM walk, async ABI: parity, kernel numbers,
# ncu text summaries of the three issue-bound kernels; the full multi-config bench.  Output: gpurun_out/s4/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s4
mkdir -p $O
timeout 600 python tools/kbench.py --gib 16 --label default > $O/kbench_default.jsonl 2> $O/kbench.err
cat $O/kbench_default.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
bash tools/ncu_summary.sh run16 "[A-Za-z0-9_]{16,}" scan_kernel $O
bash tools/ncu_summary.sh run4 "[0-9]{4,}" scan_kernel $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -8 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s4/bench.json'))
print('HEAD value',d['value'],'ms/step',d['ms_per_step'],'median',d.get('ms_per_step_median'),'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
print('e2e',d.get('e2e')); print('cpu',d.get('cpu_baseline')); print('parity',d['parity'])
for e in d['configs']: print(e['config'],'value %.0f kernel %.0f frac %.2f ms/step %.2f resolve %.2f parity %s cpu %s'%(e['value'],e['kernel_gbs'],e['frac'],e['ms_per_step'],e['resolve_ms'],e['parity'],(e.get('cpu_baseline') or {}).get('value')))
PY
du -sh gpurun_out
