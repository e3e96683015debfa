#!/usr/bin/env python
"""One line per launch from the output of tools/ncu_summary.py: python tools/ncu_table.py raw_summary.txt > table.txt
HBM fraction = (dram read + write) / duration against MEASURED_PEAKS.json hbm_gbs; tensor % = sm__pipe_tensor_cycles_active."""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hbm = 6548.2
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
blocks = open(sys.argv[1]).read().split("---\n")[1:]
def val(b, key):
    m = re.search(re.escape(key) + r" \[([^\]]*)\] = ([^\n]*)", b)
    if not m: return None, None
    return m.group(1), m.group(2).strip()
def mb(b, key):
    u, v = val(b, key)
    if v is None: return 0.0
    return float(v.replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
def us(b):
    u, v = val(b, "gpu__time_duration.sum")
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
print(f"# HBM GB/s = (dram read + write) / duration, against MEASURED_PEAKS.json hbm_gbs {hbm:.0f}; tensor % = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
print(f"{'#':>2} {'kernel':<46} {'grid':>14} {'us':>8} {'tensor%':>7} {'dramR MB':>9} {'dramW MB':>9} {'GB/s':>6} {'of HBM':>6} {'L2hit%':>6}")
tot = tt = tw = 0.0; nt = 0; traffic_t = 0.0
per = []  # (index, name, us, tensor %)
for i, b in enumerate(blocks):
    name = val(b, "Kernel Name")[1]
    name = re.sub(r"^void ", "", name); name = re.sub(r"unnamed>::", "", name); name = re.sub(r"\(.*", "", name)
    grid = val(b, "Grid Size")[1]
    t = us(b); r, w = mb(b, "dram__bytes_read.sum"), mb(b, "dram__bytes_write.sum")
    tp = float(val(b, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[1] or 0)
    l2 = float(val(b, "lts__t_sector_hit_rate.pct")[1] or 0)
    gbs = (r + w) / t * 1e3 if t > 0 else 0.0
    print(f"{i:>2} {name[:46]:<46} {grid:>14} {t:8.1f} {tp:7.1f} {r:9.1f} {w:9.1f} {gbs:6.0f} {gbs / hbm:6.2f} {l2:6.1f}")
    tot += t
    per.append((i, name, t, tp))
    if name.startswith("gemm_bf16_2sm") or name.startswith("gemm_ln"):
        tt += t; tw += t * tp; nt += 1; traffic_t += (r + w)
print(f"# total {tot:.1f} us serialised; {nt} tensor-core GEMM launches (gemm_bf16_2sm + gemm_ln_2sm): {tt:.1f} us, mean DRAM bytes per launch {traffic_t / max(nt, 1):.1f} MB, time-weighted tensor pipe active {tw / max(tt, 1e-9):.1f} %")

# the encoder / decoder GEMMs of pass 2 (between the candidate kernel and the critic / scoring tail): the north-star's ">= 50 % tensor pipe" set
first = next((i for i, n, _, _ in per if n.startswith("candidates_kernel")), -1)
last = next((i for i, n, _, _ in per if i > first and (n.startswith("critic_input") or n.startswith("score_kernel"))), len(per))
def tw_of(sel):
    t = sum(u for i, n, u, p in per if sel(i, n)); w = sum(u * p for i, n, u, p in per if sel(i, n))
    return t, (w / t if t > 0 else 0.0)
in2 = lambda i: first < i < last
for label, sel in (("pass-2 encoder / decoder GEMMs", lambda i, n: in2(i) and (n.startswith("gemm_bf16_2sm") or n.startswith("gemm_ln"))),
                   ("  of which plain gemm_bf16_2sm", lambda i, n: in2(i) and n.startswith("gemm_bf16_2sm")),
                   ("  of which gemm_ln_2sm (residual + LayerNorm epilogue)", lambda i, n: in2(i) and n.startswith("gemm_ln")),
                   ("pass-1 GEMMs (B = number of environments)", lambda i, n: i < first and (n.startswith("gemm_bf16") or n.startswith("gemm_ln"))),
                   ("critic GEMMs", lambda i, n: i >= last and n.startswith("gemm_bf16"))):
    t, p = tw_of(sel)
    print(f"# {label}: {t:.1f} us, time-weighted tensor pipe active {p:.1f} %")
