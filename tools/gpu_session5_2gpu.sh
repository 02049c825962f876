#!/bin/bash
# GPU session 5 (round 2, 2 GPUs): the N > 1 path on hardware -- NCCL record gather test, the bench under torchrun as the
# driver launches it, the reference arm under torchrun.  Output: gpurun_out/s5/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s5
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_cli.py tests/test_gpu_shapes.py -m gpu -x -q -k 'nccl or cli or chain or async' > $O/pytest_nccl.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_nccl.txt; tail -4 $O/pytest_nccl.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -8 $O/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5/bench_n2.json'))
print('HEAD n_gpus',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'],'parity',d['parity'])
print('e2e',d.get('e2e'))
for e in d['configs']: print(e['config'],'value %.0f kernel %.0f frac %.2f ms/step %.2f parity %s %s'%(e['value'],e['kernel_gbs'],e['frac'],e['ms_per_step'],e['parity'],e.get('gathered_records','')))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref n2 rc=$?"; cat $O/ref_n2.json | cut -c1-600
