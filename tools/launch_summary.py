"""profiles/r2_launches_frame_summary.csv from the ncu launch list of one frame (profiles/r2_launches_frame.csv, written by
tools/round2_profile.sh): launches and summed gpu__time_duration per kernel, largest first."""
import collections, csv, os, re, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r2_launches_frame.csv")
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
c = {k: i for i, k in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[c["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[c["Kernel Name"]]).replace("void ", "").replace("fsfb::", "")
    v = float(r[c["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[c["Metric Unit"]], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
out = os.path.join(REPO, "profiles", "r2_launches_frame_summary.csv")
with open(out, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include frame/  python tools/frame_once.py\n")
    f.write("# one steady-state FULL-SCOPE frame (segment..combine + refine + boxes), 300 k points x 6 cameras; durations are cold-cache, serialised: compare shares\n")
    f.write(f"# total {tot:.1f} us over {n} launches\n")
    f.write("kernel,launches,total_us,share\n")
    for k, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{cnt},{t:.1f},{t / tot:.4f}\n')
print(open(out).read()[:2500])
