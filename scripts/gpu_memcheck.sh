set -x
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_decode.py -x -q -m gpu -k "test_fast_and_svd_modes_agree" 2>&1 | grep -v "^$" | head -80
