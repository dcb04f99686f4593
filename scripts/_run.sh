set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_pvrcnn.py -m gpu -q -p no:cacheprovider --timeout 600 -s > $OUT/r02h_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  |err " $OUT/r02h_tests.log | head -40
