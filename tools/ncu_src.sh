#!/bin/bash
# usage: tools/ncu_src.sh '<regex on the DEMANGLED kernel name>' <out-name> [skip] [count] [profile_step args...]
# One `ncu --set full --import-source on` capture of a kernel of the configs[2] micro-batch (run under gpurun, 1 GPU).
# The regex sees template arguments, e.g. 'gemm2_tn_kernel<\(int\)7>'.
mkdir -p gpurun_out
RE=$1; OUT=$2; SKIP=${3:-0}; CNT=${4:-1}; shift 4 2>/dev/null
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/$OUT -f python tools/profile_step.py --micro 1 "$@" > gpurun_out/$OUT.log 2>&1
tail -2 gpurun_out/$OUT.log
ls -la gpurun_out/$OUT.ncu-rep
