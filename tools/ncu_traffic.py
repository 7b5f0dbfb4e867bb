"""Per-launch DRAM traffic, duration and tensor-pipe utilisation of the kernels of one forward (W48, 64 crops) from the
`ncu --set full` captures of tools/gpu_final.sh -> profiles/r02_traffic.json (read by bench.py for `roofline.traffic`).
   python tools/ncu_traffic.py gpurun_out/prof_*.ncu-rep"""
import csv, io, json, os, subprocess, sys

reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
n = 64
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reports = []            # (header, units, rows) per report: the metric columns differ between captures
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) >= 3:
        reports.append((rows[0], rows[1], rows[2:]))
rep = ",".join(os.path.basename(r) for r in reps)
col, units = {}, []


def f(r, k):
    if k not in col:
        return None
    v = r[col[k]].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return None


def to_bytes(r, k):
    v = f(r, k)
    if v is None:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[col[k]], 1)


# algorithmic bytes per launch at W48 384x288 (fp16 NHWC activations), n persons: what each kernel must read + write once
def alg_bytes(name, idx_of_kind):
    px = {48: 96 * 72, 96: 48 * 36, 192: 24 * 18, 384: 12 * 9}
    if "conv_chain_patch" in name or "conv_chain_igemm" in name:
        return None      # filled from the grid: see below
    if "fuse_sum" in name:
        return None
    if "head_argmax_finish" in name:
        return n * 17 * 54 * 8 + n * 204
    if "head_c_kernel" in name:           # fused with the argmax: reads the 48-channel map, writes 54 candidates per joint
        return n * 96 * 72 * 48 * 2 + n * 17 * 54 * 8
    if "argmax" in name:
        return n * (4 * 17 * 96 * 72 + 204)
    if "head_conv1x1" in name:
        return n * 96 * 72 * (48 * 2 + 17 * 4)
    if "stem_conv3x3s2_tc" in name:
        return n * (3 * 384 * 288 * 4 + 64 * 192 * 144 * 2)
    return None


out = {"source": os.path.basename(rep), "note": "per launch, ncu --set full --clock-control none (serialised, cold caches)",
       "n": n, "kernels": []}
for hdr, units, data in reports:
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        dur = f(r, "gpu__time_duration.sum")
        if dur is not None and units[col["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            dur /= 1e3
        elif dur is not None and units[col["gpu__time_duration.sum"]] in ("ms", "msecond"):
            dur *= 1e3
        grid = r[col["Grid Size"]] if "Grid Size" in col else None
        k = {"kernel": name, "grid": grid, "duration_us": dur, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "traffic": (rd + wr) if rd is not None and wr is not None else None,
             "dram_gbs": round((rd + wr) / dur / 1e3, 1) if rd is not None and dur else None,
             "algorithmic_bytes": alg_bytes(name, 0),
             "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "sm_active_pct": f(r, "sm__cycles_active.avg.pct_of_peak_sustained_elapsed") if "sm__cycles_active.avg.pct_of_peak_sustained_elapsed" in col else None,
             "xbar_to_sm_read_bytes": to_bytes(r, "l1tex__m_xbar2l1tex_read_bytes.sum"),
             "registers_per_thread": f(r, "launch__registers_per_thread")}
        if k["algorithmic_bytes"] and k["duration_us"]:
            k["algorithmic_gbs"] = round(k["algorithmic_bytes"] / k["duration_us"] / 1e3, 1)
        out["kernels"].append(k)
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
for k in out["kernels"]:
    print(k["kernel"][:44].ljust(44), k["grid"], f"{k['duration_us']:.1f} us" if k["duration_us"] else "", "dram", k["traffic"], "GB/s", k["dram_gbs"],
          "tensor%", k["tensor_pipe_active_pct"])
