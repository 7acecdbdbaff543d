#!/bin/bash
# attention-backward iteration: parity tests, stand-alone timing, per-kernel durations, one-CTA timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -x -q > gpurun_out/bwd_tests.log 2>&1
echo "bwd tests rc=$?"; tail -4 gpurun_out/bwd_tests.log
timeout 300 python tools/run_attention_bwd_once.py 32 5 > gpurun_out/bwd_once.log 2>&1
echo "bwd once rc=$?"; tail -3 gpurun_out/bwd_once.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bwd_launches.csv \
    python tools/run_attention_bwd_once.py 32 1 > gpurun_out/bwd_ncu.log 2>&1
echo "ncu rc=$?"; python tools/launch_summary.py gpurun_out/bwd_launches.csv | head -6
[ -x tools/micro/abwd_trace ] && timeout 120 tools/micro/abwd_trace 32 > gpurun_out/abwd_trace.txt 2>&1
