#!/bin/bash
# A/B of the encode step over library builds / schedules; results appended to gpurun_out/ab.log
# usage: tools/ab_run.sh "dtype lib ENV=VAL[,ENV=VAL..]|- [ab_step args]" ...   (lib = default | <variant name>)
mkdir -p gpurun_out
L=$PWD/arxiv_rag_b200/lib
: > gpurun_out/ab.log
for cfg in "$@"; do
  set -- $cfg
  dtype=$1; lib=$2; envs=$3; shift 3
  (
    if [ "$lib" != default ]; then export ARB_LIB_PATH=$L/libarxiv_rag_b200_$lib.so; fi
    if [ "$envs" != "-" ]; then for kv in ${envs//,/ }; do export "$kv"; done; fi
    echo "## $cfg" >> gpurun_out/ab.log
    timeout 300 python tools/ab_step.py --dtype $dtype "$@" 2>&1 | grep -v -E "UserWarning|_warn_once" >> gpurun_out/ab.log || echo "FAILED: $cfg" >> gpurun_out/ab.log
  )
done
cat gpurun_out/ab.log
