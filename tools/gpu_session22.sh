#!/bin/bash
# GPU session 22 (round 2, final): the whole GPU suite, the full bench line, the launch list of a bench run, the command line's feed.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s22
mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -6 $O/pytest_gpu.txt | cut -c1-220
timeout -s KILL 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -4 $O/bench.err | cut -c1-220
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/s22/bench.json'))
    print('HEAD value',d.get('value'),'ms/step',d.get('ms_per_step'),'frac',d['roofline']['frac'],'traffic',d['roofline'].get('traffic'),'parity',d.get('parity'))
    print('e2e',d.get('e2e')); print('cpu',d.get('cpu_baseline'))
    for e in d['configs']: print(e['config'],'value %.0f kernel %.0f frac %.2f ms/step %.2f resolve %.2f parity %s cpu %s'%(e['value'],e['kernel_gbs'],e['frac'],e['ms_per_step'],e['resolve_ms'],e['parity'],(e.get('cpu_baseline') or {}).get('value')))
except Exception as ex: print('bench line unreadable', ex)
PY
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --quick --only 1 --corpus-gib 16 > $O/bench_under_ncu.log 2>&1; echo "ncu rc=$?"; wc -l $O/launches.csv
timeout -s KILL 120 python tools/feed_bench.py 8192 quick > $O/feed_bench.txt 2>&1; echo "feed rc=$?"; grep -a "wall" $O/feed_bench.txt | cut -c1-260
