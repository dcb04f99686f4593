#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 120 ./scripts/tma_gather_probe.bin > $OUT/r02v_tma_gather_probe.txt 2>&1; cat $OUT/r02v_tma_gather_probe.txt
timeout 600 python -m pytest tests/test_gpu_pvrcnn.py -m gpu -q -p no:cacheprovider --timeout 300 -x > $OUT/r02v_tests.log 2>&1
tail -5 $OUT/r02v_tests.log
