#!/usr/bin/env python
"""Launch the tensor-core GEMM a few times on one shape (for ncu): python tools/gemm_one.py M N K flags [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3pc_b200 import _native as nat
M, N, K, flags = [int(x) for x in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
L = nat.lib()
A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); b = torch.randn(N, device="cuda")
C = torch.zeros(M, N, device="cuda") if flags & 2 else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(reps):
    nat.check(L.m3pc_gemm_bf16(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags, None))
torch.cuda.synchronize()
