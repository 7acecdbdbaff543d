#!/bin/bash
# Runs every GPU test file in its own process (a CUDA trap poisons the context) under a timeout, logging to gpurun_out/.
# usage: tools/gpu_check.sh [test files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
files=("$@")
if [ ${#files[@]} -eq 0 ]; then
  files=(tests/test_gpu_gemm.py tests/test_gpu_glue.py tests/test_gpu_attention.py tests/test_gpu_decoder.py tests/test_gpu_model.py)
fi
rc_all=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout 600 python -m pytest "$f" -m gpu -q -s -x --timeout 300 > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log" | cut -c1-300
  grep -E "^\[|passed|failed|error" "gpurun_out/$name.log" | cut -c1-300 >> gpurun_out/summary.txt
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
