"""ncu launch list (--metrics gpu__time_duration.sum, --csv) -> per-kernel totals of the LAST training step in the log.
    python tools/launch_list.py gpurun_out/r2p/launches_default.csv adamw_kernel > profiles/r2_launch_list_netflix_step.txt
The step boundary is the kernel named by argv[2] (the last launch of a step)."""
import collections
import csv
import sys

path, last = sys.argv[1], sys.argv[2]
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # boundary occurrences per step
rows = list(csv.reader(open(path, errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    t = float(r[mv].replace(",", ""))
    t *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[mu], 1.0)
    data.append((r[kn], t))
idx = [i for i, (k, _) in enumerate(data) if last in k]
step = data[idx[-1 - stride] + 1: idx[-1] + 1]
agg, tot = collections.OrderedDict(), 0.0
for k, t in step:
    k = k.split("(")[0][:78]
    a = agg.setdefault(k, [0.0, 0]); a[0] += t; a[1] += 1; tot += t
print(f"# {path}: last step = {len(step)} launches, {tot:.1f} us summed (ncu serialises launches and flushes caches: compare SHARES, not absolutes)")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:3d} launches  avg {t / c:9.1f} us  {k}")
