#!/bin/bash
# Run the bring-up stages one process each under `timeout`, logs into gpurun_out/.
# usage: tools/gpu_check.sh [dbg|rel] stage...
mode=${1:-dbg}; shift
mkdir -p gpurun_out
if [ "$mode" = "dbg" ]; then export ARB_LIB_PATH=$PWD/arxiv_rag_b200/lib/libarxiv_rag_b200_dbg.so; fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi_$mode.txt 2>&1
for st in "$@"; do
  echo "=== $st ($mode)"
  timeout 300 python tools/gpu_check.py $st > gpurun_out/check_${mode}_$st.log 2>&1
  echo "exit $?" >> gpurun_out/check_${mode}_$st.log
  tail -n 40 gpurun_out/check_${mode}_$st.log
done
