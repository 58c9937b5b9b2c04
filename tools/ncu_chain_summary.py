"""ncu --set full report of the MNIST step -> profiles/r02_ncu_chain_summary.json (per kernel: duration, DRAM bytes, pipe
utilisation, issue rate, instruction count), stamped with the commit the capture was taken at.  bench.py reads the DRAM
traffic of its dominant kernel from this file.
    python tools/ncu_chain_summary.py gpurun_out/r02_step_full.ncu-rep <commit> > profiles/r02_ncu_chain_summary.json"""
import csv
import json
import subprocess
import sys

rep, commit = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def val(r, name):
    if name not in ix:
        return None
    try:
        return float(r[ix[name]].replace(",", "")) * scale.get(units[ix[name]], 1.0)
    except ValueError:
        return None


label = {"ILi0ELi0ELi1": "chain_enc_fwd", "ILi0ELi0ELi2": "chain_dec_fwd", "ILi3ELi3ELi4": "chain_dec_bwd", "ILi3ELi3ELin1": "chain_enc_bwd"}
out = {"commit": commit, "source": rep, "how": "ncu --set full --clock-control none (one launch per kernel, third step of tools/step_prof.py bf16 4096)",
       "kernels": {}}
seen_chain = 0
for r in data:
    name = r[ix["Kernel Name"]]
    grid = r[ix["launch__grid_size"]] if "launch__grid_size" in ix else ""
    key = None
    if "chain_kernel" in name:
        import re
        m = re.search(r"chain_kernel<\(?(?:int\))?\s*(-?\d+),\s*(?:\(int\))?\s*(-?\d+),\s*(?:\(int\))?\s*(-?\d+)", name)
        kinds = tuple(int(x) for x in m.groups()) if m else None
        key = {(0, 0, 1): "chain_enc_fwd", (0, 0, 2): "chain_dec_fwd", (3, 3, 4): "chain_dec_bwd", (3, 3, -1): "chain_enc_bwd"}.get(kinds)
        if key is None:
            continue
    elif "tail_fwd" in name:
        key = "tail_forward"
    elif "tail_bwd" in name:
        key = "tail_backward"
    elif "adam_kernel" in name:
        key = "adam"
    elif "textdec" in name:
        key = "textdec"
    elif "textenc_fwd" in name:
        key = "textenc_forward"
    elif "textenc_bwd" in name:
        key = "textenc_backward"
    else:
        continue
    if key in out["kernels"]:
        continue
    out["kernels"][key] = {
        "kernel": name[:120], "grid": grid,
        "duration_us": val(r, "gpu__time_duration.sum"),
        "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
        "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": val(r, "smsp__inst_executed.sum"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
    }
print(json.dumps(out, indent=1))
