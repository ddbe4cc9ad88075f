#!/bin/bash
# builds qgemm_bench variants into tools/variants/ (git-ignored), each with different -D switches
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
SRC="tools/qgemm_bench.cu quantum_systems_b200/csrc/quarter_gemm.cu quantum_systems_b200/csrc/core.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
build() { name=$1; shift; nvcc $FLAGS "$@" -o tools/variants/$name $SRC & }
build base
build nostore -DQS_DBG_NOSTORE
build noload -DQS_DBG_NOLOAD
build noload_nostore -DQS_DBG_NOLOAD -DQS_DBG_NOSTORE
build nolds -DQS_DBG_NOLDS -DQS_DBG_NOSTORE
build nowait -DQS_DBG_NOWAIT -DQS_DBG_NOLOAD -DQS_DBG_NOSTORE
build nolds_nowait -DQS_DBG_NOLDS -DQS_DBG_NOWAIT -DQS_DBG_NOLOAD -DQS_DBG_NOSTORE
build sigend -DQS_SIGNAL_END
build sig1 -DQS_SIGNAL_NUM=1
build sig2 -DQS_SIGNAL_NUM=2
build sig3 -DQS_SIGNAL_NUM=3
build sig5 -DQS_SIGNAL_NUM=5
build sig6 -DQS_SIGNAL_NUM=6
build stages6 -DQS_STAGES=6
build l2none -DQS_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_NONE
build l2_128 -DQS_L2_PROMOTION=CU_TENSOR_MAP_L2_PROMOTION_L2_128B
wait
ls -la tools/variants
