"""First-contact diagnostics for a fresh GPU box: prints rather than asserts, so one run tells as much as possible.
    python tests/gpu_diag.py [gemm|gpt|all]
"""
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def gemm_probe(tag=""):
    from gpu_util import gemm
    ok_all = True
    for (M, N, K, bn) in [(128, 128, 64, 128), (128, 128, 128, 128), (128, 32, 64, 32), (256, 256, 768, 256), (128, 64, 192, 64)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).half()
        B = torch.randn(N, K, device="cuda", generator=g).half()
        t0 = time.time()
        out = gemm(A, B, block_n=bn)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        err = (out - ref).abs()
        ok = err.max().item() < 2e-2 * ref.abs().max().item()
        ok_all &= ok
        print(f"[gemm{tag}] {M}x{N}x{K} bn{bn}: max err {err.max().item():.4g} ref max {ref.abs().max().item():.4g} "
              f"{'OK' if ok else 'BAD'} ({time.time() - t0:.2f}s)")
        if not ok:
            bad = err > 2e-2 * ref.abs().max()
            print("   bad fraction", bad.float().mean().item(), "bad rows%8 hist", torch.bincount(bad.nonzero()[:, 0] % 8, minlength=8).tolist(),
                  "bad cols%8 hist", torch.bincount(bad.nonzero()[:, 1] % 8, minlength=8).tolist())
            # is it a permutation of K chunks? compare against per-16-chunk partial products
            if K == 64:
                for perm_name, idx in [("k16 chunks reversed", [3, 2, 1, 0]), ("only chunk0 x4", [0, 0, 0, 0])]:
                    Ap = torch.cat([A[:, 16 * i:16 * i + 16] for i in idx], 1)
                    r2 = Ap.float() @ torch.cat([B[:, 16 * i:16 * i + 16] for i in idx], 1).float().t()
                    print("   hypothesis", perm_name, "err", (out - r2).abs().max().item())
            print("   out[0,:4]", out[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())
    return ok_all


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    if what in ("gemm", "all"):
        ok = gemm_probe()
        if not ok and "CTP_DESC" not in os.environ:
            # sweep descriptor variants in subprocesses (the library reads CTP_DESC once at init)
            for desc in ["0,64,2,2", "1,64,2,4", "64,1,2,2", "1,64,1,2", "1,64,4,2", "1,64,6,2", "8,64,0,16", "1,8,0,2"]:
                env = dict(os.environ, CTP_DESC=desc)
                r = subprocess.run([sys.executable, __file__, "gemm_sub"], env=env, capture_output=True, text=True, timeout=300)
                print(f"--- CTP_DESC={desc} rc={r.returncode}")
                print("\n".join(l for l in r.stdout.splitlines() if l.startswith("[gemm")))
                if r.returncode:
                    print(r.stderr[-400:])
    if what == "gemm_sub":
        gemm_probe(tag=" " + os.environ.get("CTP_DESC", ""))


if __name__ == "__main__":
    main()
