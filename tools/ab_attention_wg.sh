#!/bin/bash
# A/B of the forward attention kernel: BSEG_ATTN_WG=2 (default build) vs 1 (rebuilt on the box).  Runs the attention
# parity tests for both.  Output: gpurun_out/ab_attention_wg.txt
set -u
out=gpurun_out/ab_attention_wg.txt
mkdir -p gpurun_out
{
  echo "== WG=2 (default)"
  python tools/run_attention_once.py 64
  python tools/run_attention_once.py 128
  python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -2
  echo "== WG=1 (two CTAs per SM)"
  touch beach_seg_b200/csrc/attention.cu
  make -C beach_seg_b200/csrc EXTRA=-DBSEG_ATTN_WG=1 > /dev/null 2>&1 || echo "BUILD FAILED"
  timeout 300 python tools/run_attention_once.py 64
  timeout 300 python tools/run_attention_once.py 128
  timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -2
} > $out 2>&1
cat $out
