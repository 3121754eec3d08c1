#!/bin/bash
# deterministic comparison of kernel variants: warp instructions and cycles of one 100-spp launch (run on the GPU box)
ncu --metrics smsp__inst_executed.sum,sm__cycles_elapsed.avg,smsp__inst_executed_pipe_fma.sum --clock-control none -k regex:render_kernel -s 1 -c 1 python tools/profile_render.py ${1:-100} 2>&1 | grep -E "smsp__inst_executed|sm__cycles_elapsed" 
