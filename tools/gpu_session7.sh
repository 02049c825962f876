#!/bin/bash
# GPU session 7 (round 2): merged slice+halo copies, look-arounds on the device VM: kernel numbers, full GPU suite, full bench.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s7
mkdir -p $O
timeout 600 python tools/kbench.py --gib 16 --label default > $O/kbench_default.jsonl 2> $O/kbench.err
cat $O/kbench_default.jsonl
timeout 1300 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -6 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s7/bench.json'))
print('HEAD value',d['value'],'ms/step',d['ms_per_step'],'median',d.get('ms_per_step_median'),'kernel_ms',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'traffic',d['roofline']['traffic'])
print('e2e',d.get('e2e')); print('cpu',d.get('cpu_baseline')); print('parity',d['parity'])
for e in d['configs']: print(e['config'],'value %.0f kernel %.0f frac %.2f ms/step %.2f resolve %.2f parity %s cpu %s traffic %s'%(e['value'],e['kernel_gbs'],e['frac'],e['ms_per_step'],e['resolve_ms'],e['parity'],(e.get('cpu_baseline') or {}).get('value'),e.get('traffic')))
PY
