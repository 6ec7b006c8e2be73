"""One steady-state frame out of an `ncu --metrics gpu__time_duration.sum` launch list of bench.py: launches between two
segment-stage k_voxelize launches, aggregated by kernel.  python tools/launches_frame.py raw.csv [frame_index] > frame.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
names = [r[ix["Kernel Name"]] for r in data]
ns = [float(r[ix["Metric Value"]]) for r in data]
vox = [i for i, n in enumerate(names) if "k_voxelize" in n]
starts = [v for v, nxt in zip(vox, vox[1:]) if nxt - v > 250]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a, b = starts[k], starts[k + 1]
agg = {}
for n, t in zip(names[a:b], ns[a:b]):
    short = re.sub(r"\(.*", "", n).replace("void ", "").replace("fsfb::", "")
    short = re.sub(r"<.*", "", short)
    e = agg.setdefault(short, [0, 0.0])
    e[0] += 1
    e[1] += t / 1e3
tot = sum(v[1] for v in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 1 --warmup 3 (300k pts x 6 cams frame)")
print(f"# one full frame = launches [{a}, {b}) between two segment-stage k_voxelize launches; durations are cold-cache, serialised")
print(f"# total {tot:.1f} us over {b - a} launches (kernels + memsets as ncu lists them)")
print("kernel,launches,total_us,share_pct")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{t:.1f},{100 * t / tot:.2f}")
