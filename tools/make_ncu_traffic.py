"""profiles/ncu_traffic.json from an `ncu --set full` report of tools/prof_layer.py (one layer call fwd+bwd at bench size):
average dram__bytes_read.sum + dram__bytes_write.sum of the fc1 / fc2 grouped-GEMM launches of the forward, keyed by
configuration and stamped with the sha256 of the kernel sources (bench.py's `roofline.traffic` reads it back and reports
null once the sources change).

    python tools/make_ncu_traffic.py gpurun_out/x.ncu-rep profiles/r2_ncu_all_kernels.md [B]
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, capture_name = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, c):
    return float(r[ix[c]].replace(",", ""))


def to_bytes(r, c, units_row=rows[1]):
    v, u = num(r, c), units_row[ix[c]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


# forward GEMMs of the training path = the first two gg_kernel launches of a layer call (fc1: bias + GELU, fc2: bias)
gg = [r for r in data if "gg_kernel" in r[ix["Kernel Name"]]]
fwd = gg[:2]
per = [to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in fwd]
T, D, H, E, K = B * bench.N_TOK, bench.D_MODEL, bench.D_HID, bench.N_EXP, bench.TOP_K
sources = ["m3vit_b200/csrc/ffn_bf16.cu", "m3vit_b200/csrc/tc_common.cuh", "m3vit_b200/csrc/common.cuh"]
h = hashlib.sha256()
for f in sources:
    with open(os.path.join(ROOT, f), "rb") as fh:
        h.update(fh.read())
out = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d[f"ffn_fwd_train:bf16:T{T}:D{D}:H{H}:E{E}:K{K}"] = {
    "dram_bytes_per_launch": sum(per) / len(per), "launches_averaged": len(per), "per_launch": per,
    "capture": capture_name, "sources": sources, "sources_sha256": h.hexdigest()}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
