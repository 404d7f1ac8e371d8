"""Per-launch DRAM traffic of the library's dominant kernels from an `ncu --set full` report, keyed by the C-ABI
labels bench.py profiles -> profiles/ncu_traffic.json.  Usage: python scripts/ncu_traffic.py REPORT.ncu-rep OUT.json"""
import csv, io, json, re, subprocess, sys
from collections import defaultdict

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = defaultdict(list)
for r in rows[2:]:
    m = re.search(r"(\w*kernel\w*)", r[col["Kernel Name"]])
    name = m.group(1) if m else r[col["Kernel Name"]].split("(")[0]
    b = sum(float(r[col[m]]) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per[name].append((b, float(r[col["gpu__time_duration.sum"]])))
kern = {k: {"dram_bytes_per_launch": sum(b for b, _ in v) / len(v), "us_per_launch_under_ncu": sum(t for _, t in v) / len(v),
            "launches": len(v)} for k, v in per.items()}
label_map = {
    "glam_triplet_edge_fwd": ["edge_win2_fwd_kernel"],
    "glam_triplet_edge_bwd_dst": ["edge_win_bwd_dst2_kernel"],
    "glam_triplet_edge_bwd_dst[single-buffered]": ["edge_win_bwd_dst_kernel"],
    "glam_triplet_edge_bwd_src": ["edge_win_bwd_src_kernel"],
    "glam_triplet_edge_fwd[gather]": ["edge_tile_fwd_kernel"],
    "glam_triplet_edge_bwd_dst[gather]": ["edge_dots_ep_kernel", "edge_softmax_bwd_kernel"],
    "glam_triplet_edge_bwd_src[gather]": ["edge_source_bwd_kernel"],
    "glam_gru_fused_fwd": ["tc_gru_fwd_kernel"],
    "glam_gru_gates_bwd_ex": ["gru_gates_bwd_vec_kernel"],
    "glam_gru_gates_fwd": ["gru_gates_fwd_vec_kernel"],
    "glam_gru_gates_bwd": ["gru_gates_bwd_vec_kernel"],
}
res = {"_kernels": kern}
for label, ks in label_map.items():
    if all(k in kern for k in ks):
        res[label] = sum(kern[k]["dram_bytes_per_launch"] for k in ks)
json.dump(res, open(out, "w"), indent=1)
for k, v in sorted(kern.items()):
    print(f"{k:36s} {v['dram_bytes_per_launch'] / 1e6:8.1f} MB/launch  {v['us_per_launch_under_ncu']:7.1f} us  x{v['launches']}")
