"""Summarises an `ncu --set full` capture of one forward (scripts/gpu_round.sh) into a CSV with one
row per tcgen05 launch and the DRAM traffic JSON that bench.py quotes as roofline.traffic.
Usage: python scripts/ncu_summary.py <report.ncu-rep> <batch in the capture> <tag>"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, batch, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
names = ["stem+downs.0.net.3+pool", "downs.1.net.0", "downs.1.net.3+pool", "downs.2.net.0",
         "downs.2.net.3+pool", "downs.3.net.0", "downs.3.net.3+pool", "bottleneck.net.0",
         "bottleneck.net.3", "ups.0(convT)+ups.1.net.0(cat)", "ups.1.net.3",
         "ups.2(convT)+ups.3.net.0(cat)", "ups.3.net.3", "ups.4(convT)+ups.5.net.0(cat)", "ups.5.net.3",
         "ups.6(convT)+ups.7.net.0(cat)", "ups.7.net.3+head"]   # the 17 launches of the composed decoder
want = ["Kernel Name", "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "lts__t_bytes.sum",
        "sm__inst_executed.sum"]
idx = []
for k in want:
    hit = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
    if hit:
        idx.append(hit[0])
assert len(data) == len(names), (len(data), len(names))
out = ROOT / "profiles" / f"ncu_full_tc_{tag}_batch{batch}.csv"
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["layer"] + [hdr[i] for i in idx])
    w.writerow([""] + [units[i] for i in idx])
    for n, r in zip(names, data):
        w.writerow([n] + [r[i] for i in idx])
col = {h: i for i, h in enumerate(hdr)}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
tot = 0.0
per = {}
for n, r in zip(names, data):
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[k]]) * scale[units[col[k]]]
    per[n] = b
    tot += b
traffic = {
    "source": f"ncu --set full, one forward at batch {batch} (profiles/{out.name}); DRAM bytes scale "
              "linearly with the batch, so the bench's batch-512 figure is this x 512 / batch",
    "batch": batch,
    "dram_bytes_per_launch": per,
    "launches": len(names),
    "dram_bytes_per_launch_avg_batch512": tot / len(names) * 512 / batch,
    "dram_bytes_per_frame_tc_launches": tot / batch,
}
(ROOT / "profiles" / f"traffic_{tag}.json").write_text(json.dumps(traffic, indent=1))
print(out, "DRAM MB/frame over the tcgen05 launches:", round(tot / batch / 1e6, 2))
