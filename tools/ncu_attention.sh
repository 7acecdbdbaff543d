#!/bin/bash
mkdir -p gpurun_out
CMD="python -m pytest tests/test_gpu_attention.py -m gpu -q -k nseq2-3.0 -x"
CMD="python tools/run_attention_once.py"
$CMD > gpurun_out/attn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -f -o gpurun_out/attn $CMD > gpurun_out/attn_ncu.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/attn_ncu.log
