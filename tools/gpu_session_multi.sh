#!/bin/bash
# Multi-GPU session under `gpurun --gpus N`: the multi-rank parity test on min(N, 8) ranks and the bench line at N GPUs.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_session_multi.sh r02 8'
tag=${1:-r02}
n=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_n${n}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${tag}_n${n}_gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -rs 2>&1 | tail -40 ) > gpurun_out/${tag}_multi_gpu_pytest_n${n}.log
tail -5 gpurun_out/${tag}_multi_gpu_pytest_n${n}.log
if [ -n "$3" ]; then      # extra single-GPU test files to run in the same lease
  ( timeout 600 python -m pytest $3 -m gpu -q -rs 2>&1 | tail -30 ) > gpurun_out/${tag}_extra_pytest_n${n}.log
  tail -3 gpurun_out/${tag}_extra_pytest_n${n}.log
fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 3 --warmup 3 \
  > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "bench rc=$?"; tail -4 gpurun_out/${tag}_bench_n${n}.err; wc -c gpurun_out/${tag}_bench_n${n}.json
