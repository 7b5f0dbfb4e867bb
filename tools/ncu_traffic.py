"""Extracts per-launch DRAM traffic and tensor-pipe utilisation of the stage-4 conv kernels from an `ncu --set full`
capture of tools/profile_convs.py (4 launches: C=48, 96, 192, 384 at N=64) into profiles/r01_traffic.json.
  python tools/ncu_traffic.py gpurun_out/prof_convs.ncu-rep"""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
shapes = [(48, "96x72"), (96, "48x36"), (192, "24x18"), (384, "12x9")]
n = 64
out = {"source": os.path.basename(rep), "note": "per launch, ncu --set full --clock-control none, cold caches", "kernels": []}
def f(r, k):
    v = r[col[k]].replace(",", "")
    return float(v) if v else None
def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
for r, (c, m) in zip(data, shapes):
    h, w = map(int, m.split("x"))
    rd = to_bytes(f(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(f(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
    alg = n * h * w * c * 2 * 3 + 9 * c * c * 2   # read input + residual, write output (fp16) + weights
    out["kernels"].append({
        "kernel": r[col["Kernel Name"]].split("(")[0], "C": c, "map": m,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr, "algorithmic_bytes": alg,
        "duration_us": f(r, "gpu__time_duration.sum"),
        "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "xbar_to_sm_read_bytes": to_bytes(f(r, "l1tex__m_xbar2l1tex_read_bytes.sum"), units[col["l1tex__m_xbar2l1tex_read_bytes.sum"]]),
    })
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
