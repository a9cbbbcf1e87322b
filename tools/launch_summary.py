"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki]
    k = k.replace("void ", "").replace("unetk::", "")
    k = k.split("(")[0][:60]
    tot[k][0] += 1; tot[k][1] += v
s = sum(v[1] for v in tot.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in tot.values())} launches, {s/1e6:.3f} ms total, /{steps:g} steps = {s/1e6/steps:.3f} ms per step")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e6/steps:9.3f} ms/step {v[0]/steps:7.1f} launches/step {100*v[1]/s:5.1f}%  {k}")
