#!/bin/bash
# multi-GPU evidence (under gpurun --gpus N): bash tools/r2_multi.sh N tag
N=$1; TAG=${2:-r2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_nccl2.py -m gpu -x -q > $O/${TAG}_nccl2_test.log 2>&1; echo "nccl2 test rc=$?"; tail -3 $O/${TAG}_nccl2_test.log
fi
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus $N --config ant1m --steps 5 --warmup 2 > $O/${TAG}_bench_ant1m_${N}gpu.json 2> $O/${TAG}_bench_ant1m_${N}gpu.err; echo "ant1m rc=$?"
timeout 300 $TR tools/d2h_ceiling.py > $O/${TAG}_d2h_ceiling_${N}gpu.json 2> $O/${TAG}_d2h_ceiling_${N}gpu.err; echo "d2h ceiling rc=$?"
if [ "$N" = "8" ] && [ -z "$NO_SWEEP" ]; then
  timeout 900 $TR bench.py --gpus $N --config sweep --steps 3 --warmup 1 > $O/${TAG}_bench_sweep_${N}gpu.json 2> $O/${TAG}_bench_sweep_${N}gpu.err; echo "sweep rc=$?"
  if [ -n "$NOBIND" ]; then
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-numa-bind > $O/${TAG}_bench_${N}gpu_nobind.json 2> $O/${TAG}_bench_${N}gpu_nobind.err; echo "bench nobind rc=$?"
  fi
fi
