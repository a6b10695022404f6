#!/bin/bash
# A/B of the encode step over library builds / schedules; one line per configuration in gpurun_out/ab.log
# usage: tools/ab_run.sh "dtype defer lib [extra args]" ...   (lib = default | <variant name>)
mkdir -p gpurun_out
L=$PWD/arxiv_rag_b200/lib
: > gpurun_out/ab.log
for cfg in "$@"; do
  set -- $cfg
  dtype=$1; defer=$2; lib=$3; shift 3
  if [ "$lib" = default ]; then unset ARB_LIB_PATH; else export ARB_LIB_PATH=$L/libarxiv_rag_b200_$lib.so; fi
  ARB_ATTN_DEFER=$defer timeout 300 python tools/ab_step.py --dtype $dtype "$@" >> gpurun_out/ab.log 2>&1 || echo "FAILED: $cfg" >> gpurun_out/ab.log
done
unset ARB_LIB_PATH
cat gpurun_out/ab.log
