"""Condense an .ncu-rep (ncu --set full) into the per-launch table committed under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rN_x.md

Runs in the CPU-only container (ncu -i reads reports without a GPU).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block", "smem/CTA"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = load(path)
        col = {h: i for i, h in enumerate(hdr)}
        print("## %s\n" % path.split("/")[-1])
        for r in rows:
            name = r[col["Kernel Name"]]
            print("### `%s`  (launch id %s)\n" % (name.split("(")[0].replace("void ", ""), r[col["ID"]]))
            print("| metric | value |\n|---|---|")
            for m, label in METRICS:
                if m in col and r[col[m]] != "":
                    print("| %s (`%s`) | %s %s |" % (label, m, r[col[m]], units[col[m]]))
            # top stall reasons (warp-state sampling)
            stalls = []
            for h, i in col.items():
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]:
                    try:
                        stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            if stalls:
                print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))
            print()


if __name__ == "__main__":
    main()
