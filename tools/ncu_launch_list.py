"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> per-kernel table for profiles/.
    python tools/ncu_launch_list.py gpurun_out/launches.csv profiles/name_launch_list.txt "command that was profiled" """
import collections, csv, sys

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "?"
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
n, t = collections.Counter(), collections.Counter()
for r in rows:
    if r[im] != "gpu__time_duration.sum":
        continue
    k = r[ik].split("(")[0]
    n[k] += 1; t[k] += float(r[iv].replace(",", "")) * 1e-6
tot = sum(t.values())
L = [f"# ncu launch list of `{cmd}` (1 x B200): --metrics gpu__time_duration.sum --clock-control none",
     "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
     f"# {'kernel':70s} launches   total ms   share"]
for k in sorted(t, key=lambda k: -t[k]):
    L.append(f"{k:72s} {n[k]:6d} {t[k]:10.3f} {100 * t[k] / tot:7.2f} %")
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L))
