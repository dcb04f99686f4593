set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_second.py -m gpu -q -p no:cacheprovider --timeout 120 -k "backward or invert or dropin" > $OUT/r02l_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  |err |Timeout" $OUT/r02l_tests.log | head -30
