#!/bin/bash
# N-GPU checks (run under gpurun --gpus N): tests that need >1 GPU, then the torchrun bench at N and at 1
N=${1:-2}
python -m pytest tests/test_host_and_multigpu_gpu.py -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; tail -c 400 gpurun_out/bench_n$N.err
