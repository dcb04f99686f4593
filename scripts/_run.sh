set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py -m gpu -q -p no:cacheprovider --timeout 120 -k "conv or bench or engine" > $OUT/r02n_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  " $OUT/r02n_tests.log | head -20
timeout 200 python scripts/conv_variant.py 2>&1 | tail -1
