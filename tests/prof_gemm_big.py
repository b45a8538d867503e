"""Throughput of the large-shape GEMMs (prefill / vocoder) through the C ABI.  Usage: python tests/prof_gemm_big.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_util import gemm

shapes = [(4096, 2304, 768, 256), (4096, 768, 768, 256), (4096, 768, 768, 128), (4096, 6144, 768, 256), (4096, 768, 3072, 256), (4096, 768, 3072, 128),
          (131072, 1536, 512, 256), (131072, 512, 1536, 256), (131072, 512, 1536, 128), (131072, 1026, 512, 256), (16384, 4096, 4096, 256)]
for (M, N, K, bn) in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.3).half()
    B = (torch.randn(N, K, device="cuda") * 0.05).half()
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        gemm(A, B, out_f16=True, block_n=bn, out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    n = 20
    for _ in range(n):
        gemm(A, B, out_f16=True, block_n=bn, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    print(f"M{M} N{N} K{K} bn{bn}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
