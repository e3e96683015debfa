#!/usr/bin/env python
"""Fused residual GEMM + LayerNorm (m3pc_gemm_ln_bf16) vs the unfused pair (m3pc_gemm_bf16 residual + m3pc_layernorm), warm, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from m3pc_b200 import _native as nat
L = nat.lib()

def timed(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

for M in (13312, 106496):
    for K in (512, 2048):
        A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(512, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(512, device="cuda"); g = torch.ones(512, device="cuda"); be = torch.zeros(512, device="cuda")
        X = torch.randn(M, 512, device="cuda"); Y = torch.empty(M, 512, device="cuda", dtype=torch.bfloat16)
        fused = timed(lambda: nat.check(L.m3pc_gemm_ln_bf16(A.data_ptr(), W.data_ptr(), b.data_ptr(), X.data_ptr(), Y.data_ptr(), g.data_ptr(), be.data_ptr(), None, 1, M, K, None)))
        def unf():
            nat.check(L.m3pc_gemm_bf16(A.data_ptr(), W.data_ptr(), b.data_ptr(), X.data_ptr(), M, 512, K, 2, None))
            nat.check(L.m3pc_layernorm(X.data_ptr(), g.data_ptr(), be.data_ptr(), Y.data_ptr(), M, 512, 1, None))
        unfused = timed(unf)
        gemm_only = timed(lambda: nat.check(L.m3pc_gemm_bf16(A.data_ptr(), W.data_ptr(), b.data_ptr(), X.data_ptr(), M, 512, K, 2, None)))
        fl = 2.0 * M * 512 * K
        print(f"M={M} K={K}: fused {fused:.1f} us ({fl / fused / 1e6:.0f} TF/s) | unfused GEMM+LN {unfused:.1f} us (GEMM alone {gemm_only:.1f} us) | rounds {-(-M // 256) / 74:.2f}", flush=True)
