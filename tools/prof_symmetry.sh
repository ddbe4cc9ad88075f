#!/bin/bash
# `ncu --set full` of the symmetry-aware path (one GPU); report converted to CSV on the box.
set -e
OUT=${1:-gpurun_out/prof_symmetry}
python tools/prof_symmetry.py > ${OUT}_plain.log 2>&1
ncu --set full --clock-control none -k "regex:symmetry_check|mirror_fill|cyclic_fill|quarter_gemm" -c 24 -o /tmp/prof_sym python tools/prof_symmetry.py > ${OUT}_ncu.log 2>&1
ncu -i /tmp/prof_sym.ncu-rep --page raw --csv > ${OUT}_raw.csv
python tools/ncu_summary.py ${OUT}_raw.csv > ${OUT}_summary.csv
gzip -f ${OUT}_raw.csv
cat ${OUT}_summary.csv | cut -c1-200
