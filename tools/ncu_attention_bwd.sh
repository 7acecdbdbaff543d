#!/bin/bash
# one --set full capture of the two attention-backward kernels (after a plain run of the same command)
mkdir -p gpurun_out
NSEQ=${1:-32}
timeout 300 python tools/run_attention_bwd_once.py $NSEQ 1 > gpurun_out/bwd_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd -c 2 -f -o gpurun_out/attn_bwd \
    python tools/run_attention_bwd_once.py $NSEQ 1 > gpurun_out/bwd_ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/bwd_ncu_full.log
