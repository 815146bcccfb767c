timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rollout.py tests/test_gpu_soak.py -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh twosites default twosites default
