#!/bin/bash
# The multi-GPU measurement table of one box (run under `gpurun --gpus 8`): tests, then bench.py lines for BASELINE configs[1] and [4]
# at 1/2/4/8 GPUs (one process per GPU under torchrun, gather inside librtiow_cuda.so), the NCCL-gather variant, and the one-process
# lines (rtiow_ctx_create(N)).  Every JSON line lands in gpurun_out/scale_<tag>_*.json.   Usage: tools/scale_run.sh <tag> [max_gpus]
tag=${1:-x}; maxn=${2:-8}; quick=${3:-0}     # quick=1: tests, cfg2 at 1/2/4/8, cfg5 and the one-process line at 8 only (GPU-minutes are charged x8)
run() { # name, n, extra args...
  name=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/scale_${tag}_${name}.json 2> gpurun_out/scale_${tag}_${name}.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/scale_${tag}_${name}.json 2> gpurun_out/scale_${tag}_${name}.err; fi
  python - "$name" gpurun_out/scale_${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:28s} n={d['n_gpus']} value {d['value']:9.1f} e2e {d['e2e']['value']:9.1f} Mpaths/s  ms/step {d['ms_per_step']:8.2f}  kernel_ms {d.get('roofline', {}).get('kernel_ms', 0):8.2f}  {d.get('impl_detail', {}).get('parallelism', '')[:90]}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_multirank_nccl_gpu.py tests/test_host_and_multigpu_gpu.py -q -m gpu -rs > gpurun_out/scale_${tag}_tests.log 2>&1; tail -6 gpurun_out/scale_${tag}_tests.log
for n in 1 2 4 8; do [ $n -le $maxn ] && run cfg2_n$n $n --steps 5 --warmup 3 --no-cpu-baseline; done
if [ "$quick" = 1 ]; then
  [ 8 -le $maxn ] && run cfg5_n8 8 --config cfg5 --steps 2 --warmup 3 --no-cpu-baseline
  [ 8 -le $maxn ] && { timeout 600 python bench.py --inproc --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_${tag}_inproc_n8.json 2> gpurun_out/scale_${tag}_inproc_n8.err; cut -c1-160 gpurun_out/scale_${tag}_inproc_n8.json; }
  exit 0
fi
[ 8 -le $maxn ] && run cfg2_n8_nccl 8 --steps 5 --warmup 3 --gather nccl
[ 2 -le $maxn ] && run cfg2_n2_nccl 2 --steps 5 --warmup 3 --gather nccl
for n in 1 2 4 8; do [ $n -le $maxn ] && run cfg5_n$n $n --config cfg5 --steps 2 --warmup 3 --no-cpu-baseline; done
run cfg4_n1 1 --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline
for n in 2 4 8; do if [ $n -le $maxn ]; then timeout 600 python bench.py --inproc --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_inproc_n$n.json 2> gpurun_out/scale_${tag}_inproc_n$n.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/scale_${tag}_inproc_n$n.json').read().strip().splitlines()[-1]); print('inproc n=%d value %.1f Mpaths/s ms/step %.2f %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['impl_detail']['parallelism'][:80]))"; fi; done
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/scale_${tag}_reference.json 2>/dev/null; cut -c1-200 gpurun_out/scale_${tag}_reference.json
