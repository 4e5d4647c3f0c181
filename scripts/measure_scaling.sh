#!/bin/bash
# Developer tool: bench.py on N GPUs of one box (push and sweep workloads), as the driver launches it.
#   gpurun --gpus N -- 'bash scripts/measure_scaling.sh N'
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2f_push_${N}gpu.json 2> gpurun_out/r2f_push_${N}gpu.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload sweep > gpurun_out/r2f_sweep_${N}gpu.json 2> gpurun_out/r2f_sweep_${N}gpu.err
cut -c1-260 gpurun_out/r2f_push_${N}gpu.json; cut -c1-260 gpurun_out/r2f_sweep_${N}gpu.json
