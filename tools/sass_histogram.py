"""SASS opcode histogram of libfsf_b200.so per kernel: the Blackwell-specific mnemonics (B200_PROFILING.md: tcgen05.mma → UTC*MMA,
tcgen05.ld/st → LDTM/STTM, bulk / tensor copies → UBLKCP / UTMALDG / UTMASTG, cp.async → LDGSTS, mbarrier → SYNCS).
  python tools/sass_histogram.py > profiles/r2_sass_histogram.md        (no GPU needed: cuobjdump on the in-tree library)"""
import collections, hashlib, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "fullysparsefusion_b200", "_lib", "libfsf_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "USETMAXREG", "HMMA", "LDG", "STG", "LDS", "STS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cur.replace("void ", "").replace("fsfb::", "")
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        per[cur]["_total"] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o in ("UTCHMMA",) and op.startswith("UTC") and op.endswith("MMA")):
                per[cur][o] += 1
digest = hashlib.sha256(open(LIB, "rb").read()).hexdigest()
print(f"# SASS opcode histogram — libfsf_b200.so (sha256 {digest[:16]}…), `python tools/sass_histogram.py`\n")
print("Blackwell evidence: `UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UBLKCP` = cp.async.bulk, `UTMALDG`/`UTMASTG` = tensor-map")
print("TMA copies (none: the operands are gathered rows, which a tiled tensor map cannot describe — DESIGN.md section 4a), `LDGSTS` = cp.async,")
print("`SYNCS` = mbarrier ops, `USETMAXREG` = setmaxnreg.  Only kernels with at least one of the first eight columns are listed.\n")
print("| kernel | instrs | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    if any(c[o] for o in OPS[:8]):
        print(f"| `{k[:70]}` | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
print(f"| **whole library ({len(per)} kernels)** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
