python - <<'PY'
import sys,os
sys.path.insert(0,os.getcwd())
import bench
d=bench.materialise_sample(2048)
print(d)
open('/tmp/cli_dir','w').write(d)
PY
D=$(cat /tmp/cli_dir)
echo "--- one file"; time grab_b200/bin/grab-b200 -O -l foobardoesexist $D/f000032 
echo "--- 2 GiB tree, 1 thread"; time GRAB_B200_TRACE=1 grab_b200/bin/grab-b200 -r -O -l foobardoesexist $D > /dev/null
echo "--- 2 GiB tree, -n 8"; time GRAB_B200_TRACE=1 grab_b200/bin/grab-b200 -n 8 -r -O -l foobardoesexist $D 2>&1 >/dev/null | tail -12
rm -rf $D
