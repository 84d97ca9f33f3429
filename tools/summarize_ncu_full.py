"""Text summary of an `ncu --set full` report: one row per captured launch with the counters the roofline claims rest on.
Usage: python tools/summarize_ncu_full.py gpurun_out/x.ncu-rep profiles/x_summary.txt ["command that produced it"]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_pipe%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("launch__cluster_dim_x", "cluster_x")]
idx = [(hdr.index(m), n) for m, n in COLS if m in hdr]
ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
lines = [f"# {rep.split('/')[-1]}: ncu --set full --clock-control none (per-launch values; cold caches, ~40 replays per launch)"]
if cmd:
    lines.append(f"# command: {cmd}")
lines.append("# tensor_pipe% = sm__pipe_tensor_cycles_active (sees the tcgen05 / UTCHMMA pipe in this ncu build: 24 % for a kernel measured at "
             "22 % of the nominal 2.25 PFLOP/s); L2->SM = l1tex__m_xbar2l1tex_read_bytes")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("ob::", "")
    vals = []
    for i, n in idx:
        v = r[i]
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        vals.append(f"{n}={v}{units[i] if units[i] not in ('%', '') else ''}")
    lines.append(f"{name:45s} grid={r[gi]:14s} block={r[bi]:12s} " + " ".join(vals))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
