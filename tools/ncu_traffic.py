"""profiles/r2_ncu_traffic.json from an `ncu --set full` capture of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum
per launch, tied to the kernel source by sha256 (bench.py reports `roofline.traffic` from it and null when the source changed).
  python tools/ncu_traffic.py gpurun_out/prof_r2_ss.ncu-rep [launch index] "description" """
import csv, hashlib, io, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, idx, of = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def val(name, r):
    v, u = float(data[r][col[name]].replace(",", "")), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = val("dram__bytes_read.sum", idx), val("dram__bytes_write.sum", idx)
src = os.path.join(REPO, "fullysparsefusion_b200", "csrc", "gemm_ss.cu")
rec = {"kernel": data[idx][col["Kernel Name"]], "of": of, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
       "duration_us_under_ncu": float(data[idx][col["gpu__time_duration.sum"]].replace(",", "")),
       "tensor_pipe_active_pct": float(data[idx][col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
       "l2_hit_rate_pct": float(data[idx][col["lts__t_sector_hit_rate.pct"]]),
       "issue_active_pct": float(data[idx][col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
       "source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(), "report": os.path.basename(rep)}
json.dump(rec, open(os.path.join(REPO, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(rec, indent=1))
