# Where the command line's wall time goes: timestamped milestones (GRAB_B200_TRACE=1) on a tmpfs tree.
# Usage: bash tools/cli_trace.sh [n_files]      (run from the repository root on a GPU box)
N=${1:-8192}
python - <<PY
import sys,os
sys.path.insert(0,os.getcwd())
import bench
d=bench.materialise(bench.baseline_configs()[1], $N)
open('/tmp/cli_dir','w').write(d)
PY
D=$(cat /tmp/cli_dir)
B=grab_b200/bin/grab-b200
echo "--- one 1 MiB file"; time GRAB_B200_TRACE=1 $B -O -l foobardoesexist $D/f000032
echo "--- $N x 1 MiB tree, 1 lane"; time GRAB_B200_TRACE=1 $B -r -O -l foobardoesexist $D 2>/tmp/tr.txt >/dev/null; head -8 /tmp/tr.txt; echo ...; tail -8 /tmp/tr.txt
echo "--- same, 2 lanes"; time GRAB_B200_LANES=2 GRAB_B200_TRACE=1 $B -r -O -l foobardoesexist $D 2>/tmp/tr.txt >/dev/null; head -8 /tmp/tr.txt; echo ...; tail -8 /tmp/tr.txt
cat $D/f* > /dev/shm/gscan_big.bin
echo "--- one file of $N MiB"; time GRAB_B200_TRACE=1 $B -O -l foobardoesexist /dev/shm/gscan_big.bin 2>/tmp/tr.txt >/dev/null; cat /tmp/tr.txt | head -40
echo "--- same, batches of 256 MiB windows (-L -L), 2 lanes"; time GRAB_B200_LANES=2 GRAB_B200_TRACE=1 $B -L -L -O -l foobardoesexist /dev/shm/gscan_big.bin 2>/tmp/tr.txt >/dev/null; head -12 /tmp/tr.txt; echo ...; tail -6 /tmp/tr.txt
rm -rf $D /dev/shm/gscan_big.bin
