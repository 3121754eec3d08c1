#!/bin/bash
# N-GPU checks (run under gpurun --gpus N): tests that need >1 GPU, then the torchrun bench at N (headline config, optionally more)
# usage: tools/multigpu_check.sh N [extra config ...]
N=${1:-2}; shift
python -m pytest tests/test_host_and_multigpu_gpu.py -q -m gpu 2>&1 | tail -5
for cfg in cfg2 "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --config $cfg \
      > gpurun_out/bench_n${N}_$cfg.json 2> gpurun_out/bench_n${N}_$cfg.err
  cat gpurun_out/bench_n${N}_$cfg.json; tail -c 300 gpurun_out/bench_n${N}_$cfg.err
done
