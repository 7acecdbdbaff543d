#!/bin/bash
# main step + kernel split + GEMM A/B legs only (about 25 s on the box): the iteration loop for a kernel change
python bench.py --no-cpu-baseline --no-train --no-fast-path --no-scene --no-noprompt --no-latency --no-native --no-fp32-check "$@"
