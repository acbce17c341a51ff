"""Extract the per-launch key metrics of an `ncu --set full` report into a small CSV + JSON under profiles/.
usage: python scripts/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r1_conv_tc_ncu"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_uniform.sum"]
keys = [k for k in keys if k in idx]


def tobytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


recs = []
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keys)
    w.writerow([units[idx[k]] for k in keys])
    for r in rows[2:]:
        w.writerow([r[idx[k]] for k in keys])
        rd = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        recs.append({"kernel": r[idx["Kernel Name"]].split("(")[0], "grid": r[idx["Grid Size"]], "time": r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]],
                     "dram_bytes": rd + wr,
                     "tensor_pipe_pct": float(r[idx["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]])})
summary = {"report": rep, "launches": recs, "mean_dram_bytes_per_launch": sum(x["dram_bytes"] for x in recs) / max(1, len(recs)),
           "mean_tensor_pipe_pct": sum(x["tensor_pipe_pct"] for x in recs) / max(1, len(recs))}
json.dump(summary, open(out + ".json", "w"), indent=1)
print(json.dumps(summary)[:600])
