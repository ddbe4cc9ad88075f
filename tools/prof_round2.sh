#!/bin/bash
# `ncu --set full` of the round-2 kernels at n = 128 (one GPU); report converted to CSV on the box.
set -e
OUT=${1:-gpurun_out/prof_round2}
python tools/prof_round2.py > ${OUT}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:quarter_gemm|mirror_fill|cyclic_fill|small_matmul|symmetry_check" -c 40 -o /tmp/prof_round2 python tools/prof_round2.py > ${OUT}_ncu.log 2>&1
ncu -i /tmp/prof_round2.ncu-rep --page raw --csv > ${OUT}_raw.csv
python tools/ncu_summary.py ${OUT}_raw.csv > ${OUT}_summary.csv
gzip -f ${OUT}_raw.csv
cut -c1-220 ${OUT}_summary.csv
