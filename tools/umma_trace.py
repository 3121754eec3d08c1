"""Cycle timeline of the tensor render kernel's loop (debug build: RTIOW_NVCC_EXTRA=-DRT_UMMA_TRACE python -m rtiow_b200.build --force).
    python tools/umma_trace.py [spp]
Block 0 / warp 0 stamps: 1 loop top, 2 work assigned, 3 group vote done, 4 per-ray code done (Philox, camera / scatter), 5 feature rows in
TMEM + a_full arrive, 6 large spheres (f64) done, 10+c chunk c's MMAs complete (full barrier passed), 7 all chunks collected, 8 candidates
drained, 9 back in the loop.  Prints the mean / median cycles of every phase over the iterations recorded."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtiow_b200 import capi

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 50
L = capi.lib()
W, H = 1200, 675
with capi.Context(1) as ctx:
    ctx.upload_scene(**capi.random_scene(1))
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    buf = (C.c_longlong * 4096)()
    for rep in range(2):
        img, st = ctx.render(cam, capi.default_params(width=W, height=H, spp=spp, seed=1))
        n = L.rtiow_debug_umma_trace(buf, 2040)
    a = np.array(buf[: 2 * n]).reshape(n, 2)
    print(f"kernel {st['kernel_ms']:.3f} ms; {n} stamps")
    tags, clk = a[:, 0], a[:, 1]
    starts = np.nonzero(tags == 1)[0]
    rows = []; fine = []
    for s, e in zip(starts[2:-1], starts[3:]):             # skip the first two iterations (cold)
        seg = {int(t): int(c) for t, c in zip(tags[s:e], clk[s:e])}
        if not all(k in seg for k in (1, 2, 3, 4, 5, 6, 10, 7, 8, 9)):
            continue
        chunks = [seg[k] for k in sorted(k for k in seg if 10 <= k < 40)]
        lds = [seg[k] for k in sorted(k for k in seg if 40 <= k < 70)]
        cols = [seg[k] for k in sorted(k for k in seg if 70 <= k < 100)]
        if len(lds) == len(chunks) == len(cols):
            fine.append([np.mean(np.array(lds) - np.array(chunks)), np.mean(np.array(cols) - np.array(lds)), np.mean(np.array(chunks[1:]) - np.array(cols[:-1]))])
        rows.append([seg[2] - seg[1], seg[3] - seg[2], seg[4] - seg[3], seg[5] - seg[4], seg[6] - seg[5], chunks[0] - seg[6],
                     (chunks[-1] - chunks[0]) / max(len(chunks) - 1, 1), seg[7] - chunks[-1], seg[8] - seg[7], seg[9] - seg[8], clk[e] - seg[1]])
    r = np.array(rows, float)
    names = ["assign_work", "group vote (wait for the slowest warp)", "per-ray code (Philox, camera/scatter)", "features -> TMEM, arrive", "large spheres f64",
             "wait for chunk 0", "per chunk (steady state)", "last chunk's sign collection", "candidate drain", "merge + loop tail", "WHOLE ITERATION"]
    for i, nm in enumerate(names):
        print(f"  {nm:42s} mean {r[:, i].mean():8.0f}  median {np.median(r[:, i]):8.0f}  p90 {np.percentile(r[:, i], 90):8.0f}")
    print(f"  iterations analysed: {len(r)}")
    if fine:
        f = np.array(fine, float).mean(axis=0)
        print(f"  inside a chunk (mean): full passed -> TMEM loads done {f[0]:.0f}; -> D handed back + signs collected + survivors listed {f[1]:.0f}; -> next full passed (waiting) {f[2]:.0f}")
