#!/bin/bash
# GPU session 2 (round 2): balanced pair kernel selected for alt4 + variants, ncu of the three issue-bound kernels,
# bench headline step time, staging-thread sweep for the pageable feed.  Output: gpurun_out/s2/
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s2
mkdir -p $O
timeout 600 python tools/kbench.py --gib 16 --only alt4,run16,run16_16k,lits100,lits8,icase,run4 --label default > $O/kbench_default.jsonl 2> $O/kbench.err
for v in fbA1 fbA3 pair16 pair24; do
  GSCAN_LIB=$PWD/grab_b200/libgscan_$v.so timeout 300 python tools/kbench.py --gib 16 --only alt4,lits8 --label $v > $O/kbench_$v.jsonl 2>> $O/kbench.err
done
cat $O/kbench_*.jsonl
for k in "alt4:foo|bar|baz|quux:FixedB" "run16:[A-Za-z0-9_]{16,}:RunEngine" "lits100:@lits100:HashEngine"; do
  n=${k%%:*}; r=${k#*:}; p=${r%%:*}; kn=${r#*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 1 -c 1 -o $O/ncu_$n -f python tools/prof_one.py "$p" 8 2 > $O/ncu_$n.log 2>&1
  tail -2 $O/ncu_$n.log
done
timeout 600 python bench.py --steps 20 --warmup 5 --only 1,2 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 400 python tools/stage_sweep.py > $O/stage_sweep.txt 2>&1; cat $O/stage_sweep.txt
