#!/usr/bin/env python
"""bench.py — Mpaths/s of the rtiow render hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1|cfg2|cfg3-*|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A *step* is one full frame of the workload: BASELINE.json configs[1], the RTIOW Part 1 final scene
(seeded random_scene, main.rs:59-102) at 1200x675, 500 spp, depth 50 — 405 M paths.  At N > 1 the same frame is
split into interleaved row tiles, one rank per GPU ("strong" scaling); the tiles reach rank 0 either fused into the epilogue
(stores into rank 0's frame over NVLink, torch symmetric memory — default when available, cross-checked against the other
path in the same run) or through tile buffers + an NCCL all-gather (--gather nccl).

  value     device-resident: scene already in HBM, tiles -> (all-gather) -> top-down frame left in HBM
  e2e       the reference-facing call with HOST buffers: scene upload (H2D) + render + frame to host (D2H)
  roofline  FP32 pipe: 17 FLOP x rays traced x spheres / render-kernel time, against an FFMA
            calibration kernel run in this process (MEASURED_PEAKS.json has no FP32 entry)
  cpu_baseline  the f64 CPU oracle (a port: the Rust reference cannot be built here) on a bounded sample
  --impl reference  times that same CPU restatement, all host threads, as the reference arm
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version banner must not land on stdout: ONE JSON line there

WORKLOAD = dict(name="RTIOW Part 1 final scene (seeded random_scene), 1200x675, 500 spp, depth 50", width=1200, height=675, spp=500,
                max_depth=50, t_min=1e-4, scene_seed=1, sample_seed=1)
# BASELINE.json configs by name; the default (and the line the driver records) is cfg2 = configs[1].  The others are parity-test
# cases (tests/test_full_size_gpu.py) that can also be timed: (workload name, width, height, spp, grid half-extent, material mode)
CONFIGS = {
    "cfg1": ("RTIOW Part 1 final scene (seeded random_scene), 400x225, 10 spp, depth 50", 400, 225, 10, 11, 0),
    "cfg2": (WORKLOAD["name"], 1200, 675, 500, 11, 0),
    "cfg3-lambertian": ("material isolation: all-Lambertian, 800x450, 100 spp, depth 50", 800, 450, 100, 11, 1),
    "cfg3-metal": ("material isolation: all-Metal (fuzz), 800x450, 100 spp, depth 50", 800, 450, 100, 11, 2),
    "cfg3-dielectric": ("material isolation: all-Dialectric + hollow glass shell, 800x450, 100 spp, depth 50", 800, 450, 100, 11, 3),
    "cfg4": ("random-spheres scene scaled to 10k spheres (grid -50..=50), 1920x1080, 256 spp, depth 50", 1920, 1080, 256, 50, 0),
    "cfg5": ("RTIOW Part 1 final scene (seeded random_scene), 3840x2160, 1024 spp, depth 50", 3840, 2160, 1024, 11, 0),
}
FLOP_PER_TEST = 17.0            # sphere.rs:18-25 with a and r^2 hoisted
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
# dram__bytes_read.sum + dram__bytes_write.sum of ONE render_kernel launch of this workload on one GPU, from the ncu --set full
# capture summarised in profiles/r1_p_render_kernel_ncu_bench_size.txt (19.56 MB read + 0.72 MB written): the scene, the
# per-sphere records and the fixed-point accumulators; the 9.7 TFLOP of the launch run out of shared memory and registers.
NCU_DRAM_BYTES_PER_LAUNCH = 20_274_176


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: configs[1], the headline)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--tile-rows", type=int, default=1)
    ap.add_argument("--cpu-sample-spp", type=int, default=0, help="spp of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1: 'fused' = every rank's epilogue stores its pixels into rank 0's frame over NVLink (torch symmetric memory) and a "
                         "device-side barrier follows; 'nccl' = tile buffers + NCCL all-gather + de-interleave; 'auto' = fused when available")
    a = ap.parse_args()
    name, w, h, spp, grid, mode = CONFIGS[a.config]
    a.workload = name if (a.width, a.height, a.spp) == (None, None, None) else f"{name} [overridden size]"
    a.width, a.height, a.spp = a.width or w, a.height or h, a.spp or spp
    a.grid, a.material_mode = grid, mode
    return a


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arms
def oracle_world(scene_arrays):
    from oracle import oracle as o            # the checker / CPU baseline: bench.py is one of the places allowed to load it
    return o, o.Scene(**scene_arrays)


def time_oracle(o, sc, W, H, spp, seed):
    cam = o.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    t0 = time.perf_counter()
    # n_threads explicit: torchrun exports OMP_NUM_THREADS=1, the reference arm must use every host core
    _, _, cnt = o.render(sc, cam, W, H, spp, seed=seed, sampler=o.SAMPLER_REJECTION, n_threads=o.host_threads())
    dt = time.perf_counter() - t0
    return W * H * spp / dt / 1e6, dt, cnt


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path, all host threads.
    rustc/cargo are absent, so this is the oracle port (oracle/rtiow_oracle.c, OpenMP rows like rayon-per-row)."""
    if rank != 0:
        return
    from rtiow_b200 import capi
    W, H = args.width, args.height
    o, sc = oracle_world(capi.random_scene(WORKLOAD["scene_seed"], args.grid, args.material_mode))
    spp = args.cpu_sample_spp or (4 if sc.n < 2000 else 1)
    cores = o.host_threads()
    for _ in range(args.warmup):
        time_oracle(o, sc, W, H, 1, 1)
    t0 = time.perf_counter()
    for k in range(args.steps):
        time_oracle(o, sc, W, H, spp, 1 + k)
    dt = time.perf_counter() - t0
    v = W * H * spp * args.steps / dt / 1e6
    sample = f"{W}x{H} @ {spp} spp per step ({W * H * spp / 1e6:.2f} M paths) of the {args.spp} spp workload; rate is spp-independent"
    emit_json({
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "width": W, "height": H, "spp": args.spp,
                                                          "max_depth": 50, "n_spheres": sc.n, "scene_seed": 1},
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of main.rs:122-145 (f64, gcc -O2, OpenMP rows); NOT `cargo run --release`: no rustc in this image"})


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from rtiow_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, spp = args.width, args.height, args.spp
    scene = capi.random_scene(WORKLOAD["scene_seed"], args.grid, args.material_mode)
    n_spheres = len(scene["radius"])
    ctx = capi.Context(device=local_rank)
    ctx.upload_scene(**scene)
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    prm = capi.default_params(width=W, height=H, spp=spp, max_depth=WORKLOAD["max_depth"], t_min=WORKLOAD["t_min"], seed=WORKLOAD["sample_seed"],
                              tile_rows=args.tile_rows)
    tile_bytes = ctx.tile_buffer_bytes(prm, world)
    dev = torch.device("cuda", local_rank)
    tiles = torch.empty(tile_bytes, dtype=torch.uint8, device=dev)
    gathered = torch.empty(tile_bytes * world, dtype=torch.uint8, device=dev) if world > 1 else tiles
    frame = torch.empty(W * H * 4, dtype=torch.uint8, device=dev)
    host_frame = torch.empty(W * H * 4, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > 126 MB L2
    # One dedicated stream for everything: the library's launches (it gets the raw handle), torch's fills and copies, NCCL's
    # stream dependencies and the timing events.  (The legacy default stream has handle 0, which the C ABI reads as "use the
    # context's own stream" — work there would not be ordered against torch's.)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0
    launches = [0]

    # N > 1, the gather.  Preferred: FUSED into the epilogue — rank 0's frame lives in symmetric memory (mapped into every rank's
    # address space over NVLink), finalize_to_frame_kernel of each rank stores its rows straight into it, and a device-side barrier
    # (signal pads, on the stream) tells rank 0 the frame is complete.  Fallback / cross-check: tile buffers + NCCL all-gather +
    # de-interleave kernel, which is what north_star names.
    fused, hdl, frame_sym, frame0_ptr, gather_note = False, None, None, 0, "single GPU"
    if world > 1:
        ok = 0
        if args.gather in ("auto", "fused"):
            try:
                import torch.distributed._symmetric_memory as symm
                frame_sym = symm.empty(W * H * 4, dtype=torch.uint8, device=dev)
                hdl = symm.rendezvous(frame_sym, dist.group.WORLD)
                frame0_ptr = int(hdl.buffer_ptrs[0])
                ok = 1
            except Exception as e:            # no P2P / symmetric-memory support on this box
                gather_note = f"NCCL all-gather (symmetric memory unavailable: {type(e).__name__})"
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        fused = bool(t_ok.item())
        if args.gather == "fused" and not fused:
            raise SystemExit("bench.py: --gather fused requested but torch symmetric memory is not available")
        if not fused and args.gather != "auto":
            gather_note = "NCCL all-gather"

    def step_nccl(stats=False):
        """tiles -> all-gather -> de-interleave, everything stays in HBM"""
        st = ctx.render_tiles_device(cam, prm, rank, world, tiles.data_ptr(), sp, want_stats=stats)
        launches[0] += 2
        if world > 1:
            dist.all_gather_into_tensor(gathered, tiles)
            ctx.deinterleave_device(gathered.data_ptr(), prm, world, frame.data_ptr(), sp)
            launches[0] += 1
        return st

    def step_fused(stats=False):
        """every rank's epilogue stores into rank 0's frame (peer memory), then a device-side barrier on the stream"""
        st = ctx.render_to_frame_device(cam, prm, rank, world, frame0_ptr, sp, want_stats=stats)
        launches[0] += 2
        hdl.barrier(channel=0)
        return st

    if fused:                                 # cross-check once: the fused frame must equal the NCCL-gathered one, byte for byte
        step_nccl(); step_fused(); torch.cuda.synchronize(dev)
        same = torch.tensor([1 if (rank != 0 or torch.equal(frame_sym, frame)) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if not bool(same.item()):             # never seen; if it happens the timed path is the one north_star names
            if rank == 0:
                a, b = frame_sym.view(H, W, 4), frame.view(H, W, 4)
                bad_rows = (a != b).any(dim=2).any(dim=1).nonzero().flatten().tolist()
                print(f"bench.py: fused gather != NCCL all-gather on {len(bad_rows)} rows (first {bad_rows[:8]}): falling back to NCCL", file=sys.stderr)
            if args.gather == "fused":
                raise SystemExit("bench.py: fused gather and NCCL all-gather disagree")
            fused = False
            gather_note = "NCCL all-gather (the fused gather failed its cross-check on this box)"
    if fused:
        gather_note = "epilogue stores into rank 0's frame over NVLink (torch symmetric memory) + device-side barrier; verified byte-identical to tiles + NCCL all-gather + de-interleave"
    step_device = step_fused if fused else step_nccl

    def step_e2e():
        """what a caller of the C ABI does per frame, from host buffers to host buffers"""
        ctx.upload_scene(**scene)                                                  # H2D: the scene SoA
        if world == 1:
            img, st = ctx.render(cam, prm, out=host_np)                            # D2H inside rtiow_render
            launches[0] += 2
            return st
        if fused:
            st = step_fused(stats=True)
            if rank == 0:
                host_frame.copy_(frame_sym, non_blocking=True)
            stream.synchronize()
            return st
        st = ctx.render_tiles_device(cam, prm, rank, world, tiles.data_ptr(), sp, want_stats=True)
        dist.all_gather_into_tensor(gathered, tiles)
        launches[0] += 2
        if rank == 0:
            ctx.deinterleave_device(gathered.data_ptr(), prm, world, frame.data_ptr(), sp)
            host_frame.copy_(frame, non_blocking=True)
            launches[0] += 1
        stream.synchronize()
        return st

    host_np = host_frame.numpy().reshape(H, W, 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # FP32-pipe calibration, same process, same clocks regime (a kernel of about the step's length)
    peak_tflops, _ = ctx.fp32_peak_probe(packed=True, target_ms=300.0)
    peak_scalar_tflops, _ = ctx.fp32_peak_probe(packed=False, target_ms=100.0)

    for _ in range(args.warmup):
        flush.fill_(1)
        step_device(stats=True)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, rays = [], []
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        flush.fill_(1)                                                             # evict L2 between timed iterations
        st = step_device(stats=True)
        kernel_ms.append(st["kernel_ms"]); rays.append(st["rays_traced"])
    e1.record(stream)
    barrier()
    total_ms = maxr(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    n_launch = launches[0]
    ms_per_step = total_ms / args.steps
    paths = W * H * spp
    value = paths / (ms_per_step * 1e-3) / 1e6
    kms = maxr(float(np.mean(kernel_ms)))
    rays_total = sumr(float(np.mean(rays)))                                        # whole frame, all ranks

    # end-to-end through the host-facing call
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st_e = step_e2e()
    torch.cuda.synchronize(dev)
    e2e_ms = maxr((time.perf_counter() - t0) * 1e3) / args.steps
    barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        o, sc = oracle_world(scene)
        cores = o.host_threads()
        s_spp = args.cpu_sample_spp or max(1, min(spp, int(2 * cores * (1200 * 675) / (W * H) * 530 / n_spheres)))   # ~10-30 s of CPU work
        v, dt, _ = time_oracle(o, sc, W, H, s_spp, 1)
        cpu = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port",
               "sample": f"{W}x{H} @ {s_spp} spp ({W * H * s_spp / 1e6:.1f} M paths, {dt:.1f} s) of the {spp} spp workload; f64 C restatement, OpenMP rows"}

    if rank == 0:
        flops = rays_total * n_spheres * FLOP_PER_TEST                             # per frame, all GPUs
        achieved = flops / (kms * 1e-3) / 1e12 / world                             # per GPU
        scene_bytes = sum(int(np.asarray(v).nbytes) for v in scene.values())
        out = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": W, "height": H, "spp": spp, "max_depth": WORKLOAD["max_depth"], "n_spheres": n_spheres,
                       "scene_seed": WORKLOAD["scene_seed"], "sample_seed": WORKLOAD["sample_seed"], "tile_rows": args.tile_rows,
                       "parallelism": f"interleaved row tiles x{world}" + (f", {gather_note}" if world > 1 else ""),
                       "l2": "256 MiB device buffer rewritten between timed steps (inside the timed region, ~0.1 ms)"},
            "clocks": clk,
            "e2e": {"value": paths / (e2e_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": scene_bytes + 176 + 48, "d2h_bytes_per_step": W * H * 4 + 16},
            "gpu_launches": n_launch,
            "roofline": {"bound": "fp32", "bound_note": "FP32 CUDA-core pipe (north_star's roofline): the scan is 7 packed FFMA2 per sphere pair out of shared memory; "
                                                        "HBM traffic is ~20 MB per 9.7 TFLOP launch and tensor cores do not apply (no contraction)",
                         "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (world == 1 and args.config == "cfg2" and (W, H, spp) == (1200, 675, 500)) else None, "traffic_unit": "bytes/launch (ncu)",
                         "kernel": "rt::render_kernel<float,true,256,3>" if n_spheres < 2500 else "rt::render_kernel<float,true,768,1>", "kernel_ms": kms, "flop_per_test": FLOP_PER_TEST,
                         "rays_per_path": rays_total / paths, "sphere_tests_per_launch": rays_total * n_spheres / world,
                         "peak_source": "FFMA2 calibration kernel in this run (rtiow_fp32_peak_probe, ~300 ms); MEASURED_PEAKS.json has no FP32 entry",
                         "peak_scalar_ffma": peak_scalar_tflops, "peak_nominal": FP32_NOMINAL_TFLOPS},
        }
        if cpu:
            out["cpu_baseline"] = cpu
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


_REAL_STDOUT = None


def claim_stdout():
    """ONE JSON line on stdout, whatever libraries print: fd 1 is pointed at stderr for the run (NCCL's version banner, torch
    warnings and the like land there) and the result line goes to the original stdout through emit_json()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
