"""One large persistent-GEMM shape, a few launches (run under ncu).  Usage: python tests/prof_gemm_one.py M N K bn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_util import gemm

M, N, K, bn = (int(v) for v in sys.argv[1:5])
A = (torch.randn(M, K, device="cuda") * 0.3).half()
B = (torch.randn(N, K, device="cuda") * 0.05).half()
out = torch.zeros(M, N, device="cuda", dtype=torch.float16)
for _ in range(4):
    gemm(A, B, out_f16=True, block_n=bn, out=out)
torch.cuda.synchronize()
