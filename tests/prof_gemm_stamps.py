"""Bring-up: where does a decode-shaped GEMM CTA spend its time?  (clock64 stamps per role)"""
import ctypes, os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(__file__))
from chatttsplus_b200 import _lib
from gpu_util import gemm
lib = _lib.lib()
lib.ctp_debug_gemm_stamps.argtypes = [ctypes.c_void_p]
for (F, T, K, split) in [(2304, 32, 768, 8), (768, 32, 768, 12), (6144, 32, 768, 3), (768, 32, 3072, 24)]:
    W = (torch.randn(F, K, device="cuda") * 0.05).half()
    X = torch.zeros(64, K, device="cuda", dtype=torch.float16); X[:T] = torch.randn(T, K, device="cuda").half()
    out = torch.zeros(T, F, device="cuda")
    n_cta = ((F + 127) // 128) * split
    for rep in range(3):
        dbg = torch.zeros(n_cta, 8, dtype=torch.int64, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").fill_(rep)  # flush L2
        torch.cuda.synchronize()
        lib.ctp_debug_gemm_stamps(ctypes.c_void_p(dbg.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gemm(W, X[:T], atomic=True, swap=True, block_n=32, split_k=split, out=out)
        e1.record()
        torch.cuda.synchronize()
        lib.ctp_debug_gemm_stamps(None)
        d = dbg.cpu().double()
        t0 = d[:, 0:1]
        rel = (d - t0)
        names = ["start", "setup_done", "first_full", "mma_committed", "accum_seen", "epi_done", "dealloc_done"]
        print(f"F{F} K{K} split{split} rep{rep}: event {e0.elapsed_time(e1)*1e3:.1f}us; mean cycles since CTA start:",
              {n: int(rel[:, i].mean()) for i, n in enumerate(names)}, "max end", int(rel[:, 6].max()),
              "cta start spread", int((d[:, 0].max() - d[:, 0].min())))
