"""profiles/ncu_traffic.json + profiles/rN_ncu_step.md from the full ncu capture of one bench step
(scripts/gpu_profile_r1.sh -> gpurun_out/prof_step_r1.ncu-rep).  Runs in the CPU-only container."""
import csv, io, json, subprocess, sys

rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_step_r2.ncu-rep"
out_md = sys.argv[2] if len(sys.argv) > 2 else "profiles/r2_ncu_step.md"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}

def mbytes(r, name):
    v, u = float(r[col[name]]), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

# launch order of one step -> the names bench.py's kernel_breakdown uses
order = {"fps_kernel": ["fps_sa1", "fps_sa2", "fps_sa3", "fps_sa4"],
         "lin_tc": ["sa1_pointwise_l1", "sa3_pointwise_l1", "sa4_pointwise_l1"],
         "sa_tc_v3": ["sa1_fused_bf16", "sa2_fused_bf16", "sa3_fused_bf16", "sa4_fused_bf16"],
         "bq_grid_query": ["ball_query_sa1"],
         "ball_query_kernel": ["ball_query_sa2", "ball_query_sa3", "ball_query_sa4"],
         "three_nn_kernel": ["three_nn_fp1", "three_nn_fp2"],
         "fp_tc_kernel": ["fp1_fused_bf16", "fp2_fused_bf16"]}
seen = {k: 0 for k in order}
traffic, table = {}, []
for r in data:
    kname = r[col["Kernel Name"]]
    key = next((k for k in order if k in kname), None)
    if key is None or seen[key] >= len(order[key]):
        continue
    name = order[key][seen[key]]
    seen[key] += 1
    rd, wr = mbytes(r, "dram__bytes_read.sum"), mbytes(r, "dram__bytes_write.sum")
    traffic[name] = int(rd + wr)
    g = lambda m: r[col[m]] if m in col else ""
    table.append((name, kname.split("(")[0].replace("void pn2::", ""), g("gpu__time_duration.sum") + " " + units[col["gpu__time_duration.sum"]],
                  "%.2f" % (rd / 1e6), "%.2f" % (wr / 1e6), g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                  g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), g("launch__registers_per_thread"),
                  g("launch__grid_size"), g("launch__block_size"), g("smsp__issue_active.avg.pct_of_peak_sustained_active")))
# the throughput-mode sampling kernel of the same level (fused.sampling_mode("throughput"): csrc/fps_bucket.cu at 40 000
# points, 8 scenes) from its own capture, scripts/gpu_r2_final.sh -> gpurun_out/prof_fpsb40k_r2.ncu-rep
import os
extra_rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_fpsb40k_r2.ncu-rep"
try:
    prev = json.load(open("profiles/ncu_traffic.json"))
except Exception:
    prev = {}
if os.path.exists(extra_rep):
    o2 = subprocess.run(["ncu", "-i", extra_rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    r2 = list(csv.reader(io.StringIO(o2)))
    h2, u2, d2 = r2[0], r2[1], r2[2:]
    c2 = {h: i for i, h in enumerate(h2)}
    for r in d2:
        if "fps_bucket_kernel" in r[c2["Kernel Name"]]:
            tot = 0.0
            for nm in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[c2[nm]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u2[c2[nm]]]
            traffic["fps_sa1:bucketed"] = int(tot)
elif "fps_sa1:bucketed" in prev:
    traffic["fps_sa1:bucketed"] = prev["fps_sa1:bucketed"]
json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1, sort_keys=True)
with open(out_md, "w") as f:
    f.write("# One bench step (B=8, 40k points, bf16 arm, lanes=1) under `ncu --set full --clock-control none`\n\n"
            "Per-launch numbers are cold-cache and serialised (compare shares, not absolutes).  `traffic` = DRAM read + write.\n\n"
            "| bench name | kernel | time | DRAM rd MB | DRAM wr MB | DRAM % | tensor % | regs | grid | block | issue % |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for t in table:
        f.write("| " + " | ".join(t) + " |\n")
print(open(out_md).read())
