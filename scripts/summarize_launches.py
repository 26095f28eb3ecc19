"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into the per-kernel table kept
under profiles/ (one train step = the launches between two consecutive mpo_adam_step / advance_seed markers)."""
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
launches = [(r[ik], float(r[iv].replace(",", "")) / 1000.0) for r in rows[1:] if r[im] == "gpu__time_duration.sum"]
# one step: from an advance_seed_kernel to the next
marks = [i for i, (k, _) in enumerate(launches) if "advance_seed" in k]
if len(marks) >= 2:
    step = launches[marks[-2]:marks[-1]]
else:
    step = launches
agg = collections.OrderedDict()
for k, us in step:
    k = k.split("(")[0].replace("void ", "").replace("mpo::", "").replace("fused::", "")
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print("# launches in the step: %d   sum of kernel time: %.1f us" % (len(step), tot))
print("\n| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k[:70], v[0], v[1], 100 * v[1] / tot))
