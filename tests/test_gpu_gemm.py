"""tcgen05/TMA GEMM parity (through the C ABI) against a plain fp32 torch matmul of the same fp16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B):
    return A.float() @ B.float().t()


def _check(out, ref, tol_rel=2e-3, what=""):
    out = out.float()
    err = (out - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    bad = (err > tol_rel * scale).nonzero()
    if bad.numel():
        r, c = bad[0].tolist()
        rows = sorted(set(bad[:, 0].tolist()))[:16]
        cols = sorted(set(bad[:, 1].tolist()))[:16]
        raise AssertionError(f"{what}: {bad.shape[0]}/{out.numel()} elements off; max err {err.max().item():.4g} (scale {scale:.4g}); "
                             f"first bad ({r},{c}) got {out[r, c].item():.5g} want {ref[r, c].item():.5g}; bad rows {rows} cols {cols}")


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128),      # one tile, one k-block
    (128, 128, 256, 128),     # k loop
    (256, 384, 768, 128),     # multi-tile
    (512, 512, 768, 256),     # wide N tile
    (200, 100, 728, 128),     # ragged M, N and K tail (vocos embed-like)
    (4096, 2304, 768, 256),   # prefill QKV shape
    (1024, 1026, 512, 128),   # vocos head shape
    (20000, 520, 256, 256),   # persistent kernel: several tiles per CTA, ragged M and N
    (19000, 1000, 768, 128),  # persistent kernel, BN=128, both TMEM accumulators cycle many times
])
def test_gemm_normal(M, N, K, bn):
    from gpu_util import gemm
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).half()
    out = gemm(A, B, block_n=bn)
    _check(out, _ref(A, B), what=f"gemm {M}x{N}x{K} bn{bn}")


@pytest.mark.parametrize("F,T,K,bn,split", [
    (2304, 32, 768, 32, 8),    # decode QKV
    (768, 32, 768, 32, 12),    # decode o_proj
    (6144, 32, 768, 32, 3),    # decode gate/up
    (768, 32, 3072, 32, 24),   # decode down
    (2504, 32, 768, 32, 6),    # heads (ragged M)
    (2304, 5, 768, 32, 4),     # small batch
    (768, 48, 768, 64, 2),     # batch > 32
])
def test_gemm_swap_splitk(F, T, K, bn, split):
    from gpu_util import gemm
    g = torch.Generator(device="cuda").manual_seed(F + T)
    W = (torch.randn(F, K, device="cuda", generator=g) * 0.05).half()
    X = torch.zeros(64, K, device="cuda", dtype=torch.float16)
    X[:T] = (torch.randn(T, K, device="cuda", generator=g)).half()
    out = torch.zeros(T, F, device="cuda")
    gemm(W, X[:T], atomic=True, swap=True, block_n=bn, split_k=split, out=out)
    _check(out, _ref(X[:T], W), what=f"swap gemm F{F} T{T} K{K} split{split}")


def test_gemm_epilogue_bias_gelu_f16():
    from gpu_util import gemm
    g = torch.Generator(device="cuda").manual_seed(5)
    A = (torch.randn(300, 512, device="cuda", generator=g) * 0.3).half()
    B = (torch.randn(2048, 512, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(2048, device="cuda", generator=g) * 0.1
    out = gemm(A, B, out_f16=True, gelu=True, bias=bias, block_n=256)
    ref = torch.nn.functional.gelu(_ref(A, B) + bias)
    _check(out, ref, tol_rel=3e-3, what="bias+gelu f16")


def test_gemm_overlapping_rows_im2col():
    """k3 convolution as a GEMM over overlapping row windows (row pitch C < K = 3C)."""
    from chatttsplus_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(9)
    T, Cc, O = 300, 128, 256
    x = torch.zeros(T + 2, Cc, device="cuda", dtype=torch.float16)
    x[1:T + 1] = (torch.randn(T, Cc, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(O, 3 * Cc, device="cuda", generator=g) * 0.05).half()
    out = torch.zeros(T, O, device="cuda")
    st = _lib.lib().ctp_gemm_f16(T, O, 3 * Cc, _lib.ptr(x), Cc, _lib.ptr(W), 3 * Cc, _lib.ptr(out), O, None, 0, 128, 1,
                                 _lib.stream_ptr())
    _lib.check(st, "gemm im2col")
    win = torch.cat([x[0:T], x[1:T + 1], x[2:T + 2]], dim=1)
    _check(out, win.float() @ W.float().t(), what="im2col")


@pytest.mark.parametrize("M,N,K,bn,f16", [(9000, 2048, 512, 256, True), (9000, 1027, 512, 128, True), (5000, 771, 320, 256, False)])
def test_gemm_persistent_epilogue(M, N, K, bn, f16):
    """Large-M shapes take the persistent double-buffered kernel; bias + GELU, aligned and unaligned output rows."""
    from gpu_util import gemm
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.3).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    out = gemm(A, B, out_f16=f16, gelu=True, bias=bias, block_n=bn)
    ref = torch.nn.functional.gelu(_ref(A, B) + bias)
    _check(out, ref, tol_rel=3e-3, what=f"persistent bias+gelu {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K,f16", [(9000, 2048, 512, True), (700, 1027, 512, True), (5000, 771, 320, False), (1281, 256, 1536, False)])
def test_gemm_cta_pair_kernel(M, N, K, f16, monkeypatch):
    """cta_group::2 kernel (two CTAs of a cluster on one 256 x 256 tile, forced for every eligible shape): odd numbers of 128-row tiles (the
    second CTA of the last pair is all padding), partial and unaligned N tiles, bias + GELU epilogue; fp16 and fp32 outputs."""
    import subprocess, sys, os, textwrap
    # the switch is read once per process: run the forced mode in a child
    code = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        from gpu_util import gemm
        g = torch.Generator(device="cuda").manual_seed({M + N})
        A = (torch.randn({M}, {K}, device="cuda", generator=g) * 0.3).half()
        B = (torch.randn({N}, {K}, device="cuda", generator=g) * 0.05).half()
        bias = torch.randn({N}, device="cuda", generator=g) * 0.1
        out = gemm(A, B, out_f16={f16}, gelu=True, bias=bias, block_n=256)
        ref = torch.nn.functional.gelu(A.float() @ B.float().t() + bias)
        err = float((out.float() - ref).abs().max()); scale = float(ref.abs().max())
        print("ERR", err, scale)
        assert err <= 3e-3 * max(1.0, scale), (err, scale)
    """)
    env = dict(os.environ, CTP_GEMM_2CTA="2")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
