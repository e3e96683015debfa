#!/usr/bin/env python
"""Key counters of every launch in an ncu report: python tools/ncu_summary.py report.ncu-rep [out.json]
Prints one block per launch and (optionally) writes the mean DRAM bytes per tensor-core GEMM launch as JSON for bench.py
(profiles/gemm_traffic.json, the roofline.traffic figure)."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]
col = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
tot, n = 0.0, 0  # DRAM bytes of the tensor-core GEMM launches only (the kernels bench.py's roofline line is about)
for r in body:
    print("---")
    for w in WANT:
        if w in col:
            print(f"  {w} [{units[col[w]]}] = {r[col[w]][:90]}")
    try:
        name = r[col["Kernel Name"]]
        if "gemm_bf16_2sm" not in name and "gemm_ln" not in name:
            continue
        tot += to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        n += 1
    except Exception:
        pass
if len(sys.argv) > 2 and n:
    json.dump({"source": "ncu --set full of one default step (tools/round_evidence.sh): dram__bytes_read.sum + dram__bytes_write.sum, mean over the "
                         "gemm_bf16_2sm_kernel and gemm_ln_2sm_kernel launches of one 8-environment x 1024-candidate plan",
               "report": rep, "launches": n, "dram_bytes_per_launch": tot / n}, open(sys.argv[2], "w"))
    print("mean dram bytes per tensor-core GEMM launch", tot / n)
