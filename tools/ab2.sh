timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rollout.py tests/test_gpu_soak.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh base default base default
