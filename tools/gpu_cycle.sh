#!/bin/bash
# one GPU iteration: parity tests, a quick render timing, optionally the full bench.  Usage: tools/gpu_cycle.sh [tag] [bench]
tag=${1:-x}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -4 gpurun_out/pytest_gpu_$tag.log
python tools/profile_render.py 50 2>&1 | tee gpurun_out/quick_$tag.log
if [ "$2" = "bench" ]; then python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -c 300 gpurun_out/bench_$tag.err; fi
