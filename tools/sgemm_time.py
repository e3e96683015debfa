"""fp32-mode GEMM (sgemm.cu) timing on the shapes of one pass-2 step: TFLOP/s against the FFMA peak of the device."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from m3pc_b200 import _native as nat

L = nat.lib()
sms = torch.cuda.get_device_properties(0).multi_processor_count
for M, N, K, flags in [(106496, 1536, 512, 0), (106496, 512, 512, 2), (106496, 2048, 512, 1), (106496, 512, 2048, 2), (13312, 1536, 512, 0), (13312, 512, 2048, 2)]:
    A, W, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    for _ in range(2):
        nat.check(L.m3pc_gemm_fp32(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags, None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        nat.check(L.m3pc_gemm_fp32(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags, None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print(f"sgemm M={M} N={N} K={K} flags={flags}: {ms:.3f} ms  {tf:.1f} TFLOP/s  ({tf / (sms * 128 * 2 * 1.965e9 / 1e12):.2f} of the FFMA peak at 1965 MHz)")
