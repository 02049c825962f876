#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/s21c
mkdir -p $O
timeout -s KILL 100 python tools/tile_edge_diag.py > $O/diag_now.txt 2>&1; cat $O/diag_now.txt | cut -c1-300
GSCAN_LIB=$PWD/grab_b200/libgscan_s20.so timeout -s KILL 100 python tools/tile_edge_diag.py > $O/diag_s20.txt 2>&1; cat $O/diag_s20.txt | cut -c1-300
