"""One process per GPU with the gather inside librtiow_cuda.so — no torch, no MPI: ranks meet through a file.

    python tools/multirank_check.py --world N [--width W --height H --spp S] [--bench K]

Spawns N copies of itself (rank r on GPU r).  Rank 0 writes rtiow_nccl_unique_id() to a temp file, the others read it; every
rank creates Context(rank=, world=, nccl_id=) and renders with each gather (NCCL all-gather, fused IPC peer stores); rank 0
compares the frames byte for byte with a single-GPU render of the same frame and prints one JSON line.  This is what a Rust /
MPI caller of the C ABI does; tests/test_multirank_nccl_gpu.py runs it on >= 2 GPUs."""
import argparse, json, os, subprocess, sys, tempfile, time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def rank_main(a):
    from rtiow_b200 import capi
    rank, world = a.rank, a.world
    idf = Path(a.rendezvous) / "nccl_id"
    if rank == 0:
        nid = capi.nccl_unique_id() if world > 1 else b""
        tmp = idf.with_suffix(".tmp"); tmp.write_bytes(nid); os.replace(tmp, idf)
    else:
        t0 = time.time()
        while not idf.exists():
            if time.time() - t0 > 120: raise SystemExit("rank 0 never published the NCCL id")
            time.sleep(0.01)
        nid = idf.read_bytes()
    ctx = capi.Context(device=rank, rank=rank, world=world, nccl_id=nid if world > 1 else None)
    ctx.upload_scene(**capi.random_scene(1, 11, 0))
    W, H = a.width, a.height
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    prm = capi.default_params(width=W, height=H, spp=a.spp, seed=1, tile_rows=a.tile_rows)
    res = {"world": world, "width": W, "height": H, "spp": a.spp}
    frames = {}
    for name, mode in (("nccl", capi.GATHER_NCCL), ("fused", capi.GATHER_FUSED), ("auto", capi.GATHER_AUTO)):
        ctx.set_gather(mode)
        try:
            img, st = ctx.render_rank(cam, prm)
        except capi.RtiowError as e:
            if mode == capi.GATHER_FUSED and e.code == capi.ERR_UNSUPPORTED:
                res[name] = "unsupported"; continue
            raise
        if rank == 0:
            frames[name] = img.copy(); res[name + "_note"] = ctx.gather_info()
        if a.bench:
            t0 = time.perf_counter()
            for _ in range(a.bench):
                ctx.render_rank(cam, prm)
            dt = (time.perf_counter() - t0) / a.bench
            if rank == 0:
                res[name + "_ms_per_frame"] = dt * 1e3; res[name + "_mpaths_s"] = W * H * a.spp / dt / 1e6
    # frames back to back without a host round trip (rtiow_render_rank_enqueue x 3 + rtiow_ctx_synchronize), both gathers
    for name, mode in (("nccl", capi.GATHER_NCCL), ("fused", capi.GATHER_FUSED)):
        if res.get(name) == "unsupported":
            continue
        ctx.set_gather(mode)
        ptrs = [ctx.render_rank_enqueue(cam, prm) for _ in range(3)]
        st = ctx.synchronize()
        if rank == 0:
            frames["enqueue_" + name] = capi.device_to_host(ptrs[-1], W * H * 4).reshape(H, W, 4)
            res["enqueue_" + name + "_kernel_ms"] = st["kernel_ms"]
    if world > 1:                                    # NCCL gather with a host frame on every rank
        ctx.set_gather(capi.GATHER_NCCL)
        img_all, _ = ctx.render_rank(cam, prm, want_frame=True)
        if rank == 1:
            np.save(Path(a.rendezvous) / "rank1_frame.npy", img_all)
    ctx.close()
    if rank == 0:
        ref_ctx = capi.Context(device=0)
        ref_ctx.upload_scene(**capi.random_scene(1, 11, 0))
        ref, _ = ref_ctx.render(cam, prm)
        ref_ctx.close()
        for k, v in frames.items():
            res[k + "_identical_to_1gpu"] = bool(np.array_equal(v, ref))
        f1 = Path(a.rendezvous) / "rank1_frame.npy"
        t0 = time.time()
        while world > 1 and not f1.exists() and time.time() - t0 < 60: time.sleep(0.05)
        if f1.exists():
            time.sleep(0.2); res["rank1_frame_identical"] = bool(np.array_equal(np.load(f1), ref))
        print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2); ap.add_argument("--rank", type=int, default=None)
    ap.add_argument("--width", type=int, default=400); ap.add_argument("--height", type=int, default=225); ap.add_argument("--spp", type=int, default=10)
    ap.add_argument("--tile-rows", type=int, default=1); ap.add_argument("--bench", type=int, default=0); ap.add_argument("--rendezvous", default=None)
    a = ap.parse_args()
    if a.rank is not None:
        return rank_main(a)
    with tempfile.TemporaryDirectory() as td:
        procs = [subprocess.Popen([sys.executable, __file__, "--rank", str(r), "--rendezvous", td] + sys.argv[1:]) for r in range(a.world)]
        rcs = [p.wait(timeout=600) for p in procs]
    raise SystemExit(max(rcs))


if __name__ == "__main__":
    main()
