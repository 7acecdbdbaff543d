#!/bin/bash
# full GPU regression: parity suite, smoke(), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench.json 2>/dev/null | head -4
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_ref.json
