#!/bin/bash
# Everything DESIGN.md §6 quotes for one GPU, on one box: tools/final_measure.sh <tag>   (outputs under gpurun_out/)
# bench.py with default flags (+ the reference arm), then — each only after the plain command exited 0 — the ncu launch list of the
# bench command and one `--set full` capture of the bench-size render launch; the per-config kernel table and the tail probe.
tag=${1:-x}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -5 gpurun_out/bench_$tag.err; exit 1; }
cut -c1-400 gpurun_out/bench_$tag.json
python bench.py --impl reference > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_$tag.err; cut -c1-300 gpurun_out/bench_${tag}_reference.json
python tools/bench_configs.py 2 > gpurun_out/configs_$tag.jsonl 2>&1; cat gpurun_out/configs_$tag.jsonl | cut -c1-250
python tools/tail_probe.py > gpurun_out/tail_$tag.txt 2>&1; cat gpurun_out/tail_$tag.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_short.json 2>/dev/null &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_under_ncu.log 2>&1
python tools/profile_render.py 500 > gpurun_out/plain_$tag.log 2>&1 && tail -1 gpurun_out/plain_$tag.log &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_$tag \
    python tools/profile_render.py 500 > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/launches_$tag.csv
