#!/bin/bash
# round-2 profiling pass: native-ingest tests, bench, ncu launch list of the bench command, one --set full capture of the
# attention forward and of the native ingest kernel (each only after the same command ran plain with exit 0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_native_resolution.py -m gpu -x -q > gpurun_out/native_tests.log 2>&1
echo "native tests rc=$?"; tail -3 gpurun_out/native_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
LIGHT="--no-cpu-baseline --no-fast-path --no-fp32-check --no-train --no-scene --no-noprompt --no-latency --no-native"
timeout 600 python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 300 python tools/run_attention_once.py > gpurun_out/attn_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 1 -f -o gpurun_out/attn_r02 \
    python tools/run_attention_once.py > gpurun_out/attn_ncu.log 2>&1
echo "ncu attention rc=$?"; tail -2 gpurun_out/attn_ncu.log
