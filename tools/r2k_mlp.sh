#!/bin/bash
# round 2k: fused MLP kernel: parity, then timing against the two launches it replaces (+ tuning-build experiments)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fused_mlp" 2>&1 | tail -8
cat > /tmp/mlp_time.py <<'PY'
import torch, sys, os
sys.path.insert(0, ".")
from m3pc_b200 import _native as nat
L = nat.lib()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for i in range(reps):
        flush.fill_(float(i)); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for M in [int(a) for a in sys.argv[1:]]:
    Y = torch.randn(M, 512, device="cuda").bfloat16(); W1 = (torch.randn(2048, 512, device="cuda") / 22).bfloat16(); W2 = (torch.randn(512, 2048, device="cuda") / 45).bfloat16()
    b1, b2 = torch.randn(2048, device="cuda"), torch.randn(512, device="cuda"); X = torch.randn(M, 512, device="cuda"); hid = torch.empty(M, 2048, device="cuda", dtype=torch.bfloat16)
    def two():
        nat.check(L.m3pc_gemm_bf16(Y.data_ptr(), W1.data_ptr(), b1.data_ptr(), hid.data_ptr(), M, 2048, 512, 1, None))
        nat.check(L.m3pc_gemm_bf16(hid.data_ptr(), W2.data_ptr(), b2.data_ptr(), X.data_ptr(), M, 512, 2048, 2, None))
    fused = lambda: nat.check(L.m3pc_mlp_fused_bf16(Y.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), X.data_ptr(), M, None))
    t2, tf = timed(two), timed(fused); fl = 2 * 2.0 * M * 512 * 2048
    print(f"tune={os.environ.get('M3PC_TUNE_MLP','-')} M={M}: two launches {t2:.1f} us ({fl/t2/1e6:.0f} TF/s) | fused {tf:.1f} us ({fl/tf/1e6:.0f} TF/s)", flush=True)
PY
timeout 120 python /tmp/mlp_time.py 26624 57344 106496
for t in 1 2 3; do M3PC_LIB=tuning M3PC_TUNE_MLP=$t timeout 120 python /tmp/mlp_time.py 106496; done
