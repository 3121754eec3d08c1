"""Attribute an ncu report's per-SASS-instruction counts to CUDA source lines.
    python tools/ncu_lines.py report.ncu-rep [kernel-mangled-substring] [top]
Joins `ncu --page source --print-source sass` (execution counts, samples, in address order) with `nvdisasm -g` of the
cubin inside rtiow_b200/lib/librtiow_cuda.so (line info, same order).  The .so must be the build that was profiled."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "render_kernelIfLb1ELi256ELi4E"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "rtiow_b200/lib/librtiow_cuda.so")], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("capi.") and f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
lines, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
iex, ismp, iav = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
lines += [(0, ("callee", 0), r[1].strip()) for r in data[len(lines):]]
if len(data) != len(lines):
    print(f"WARNING: ncu has {len(data)} SASS instructions, nvdisasm {len(lines)}: is the .so the profiled build?")
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, (_, loc, _) in zip(data, lines):
    a = agg[loc]; a[0] += int(r[iex]); a[1] += int(r[ismp]); a[2] += int(r[iav])
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
srccache = {}
def text(loc):
    f, n = loc
    for d in ("rtiow_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in srccache: srccache[p] = open(p).read().splitlines()
            return srccache[p][n - 1].strip()[:110] if 0 < n <= len(srccache[p]) else ""
    return ""
print(f"{'inst%':>6} {'smp%':>6} {'thr':>5}  location")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / tot:6.2f} {100 * a[1] / max(tots, 1):6.2f} {a[2] / max(a[0], 1):5.1f}  {loc[0]}:{loc[1]}  {text(loc)}")
