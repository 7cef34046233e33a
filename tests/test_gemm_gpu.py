"""GPU parity of the tcgen05 GEMM (fp16 in, fp32 accumulate) against a torch fp32 reference.

Tolerances: inputs are fp16-representable, products are exact in fp32, so the only difference to
the fp32 torch reference is accumulation order: |err| <= 1e-3 * max|ref| for fp32 outputs;
fp16 outputs add one fp16 rounding (2^-11 relative).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


@pytest.mark.parametrize("b_mn_major", [False, True])
def test_umma_probe(b_mn_major):
    from foundpose_b200 import _native

    a = _rand((128, 64), 1)
    b = _rand((64, 64), 2)
    out = _native.umma_probe(a, b, b_mn_major)
    torch.cuda.synchronize()
    ref = a.float() @ (b.float() if b_mn_major else b.float().t())
    err = (out - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item(), err


@pytest.mark.parametrize(
    "m,n,k",
    [(128, 128, 64), (128, 256, 64), (256, 256, 128), (901, 3072, 1024), (3604, 1024, 4096),
     (1000, 384, 384), (57, 256, 1024), (20000, 1024, 1024)],
)
def test_gemm_bias_f32(m, n, k):
    from foundpose_b200 import _native

    a = _rand((m, k), 3)
    b = _rand((n, k), 4, 0.05)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(5)).cuda()
    out = torch.full((m, n), float("nan"), device="cuda")
    out16 = torch.empty((m, n), device="cuda", dtype=torch.float16)
    _native.gemm_tn_f16(a, b, _native.EPI_BIAS_F32, bias=bias, out_f32=out, out_f16=out16)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 1e-3 * scale
    assert (out16.float() - ref).abs().max().item() <= 2e-3 * scale


def test_gemm_epilogues():
    from foundpose_b200 import _native

    m, n, k = 1802, 1024, 1024
    a = _rand((m, k), 6)
    b = _rand((n, k), 7, 0.03)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(8)).cuda()
    gamma = torch.randn(n, generator=torch.Generator().manual_seed(9)).cuda()
    acc = a.float() @ b.float().t() + bias

    o16 = torch.empty((m, n), device="cuda", dtype=torch.float16)
    _native.gemm_tn_f16(a, b, _native.EPI_BIAS_F16, bias=bias, out_f16=o16)
    assert (o16.float() - acc).abs().max().item() <= 2e-3 * acc.abs().max().item()

    _native.gemm_tn_f16(a, b, _native.EPI_BIAS_GELU_F16, bias=bias, out_f16=o16)
    ref = torch.nn.functional.gelu(acc)
    assert (o16.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()

    x = torch.randn(m, n, generator=torch.Generator().manual_seed(10)).cuda()
    ref = x + gamma * acc
    _native.gemm_tn_f16(a, b, _native.EPI_RESID_F32, bias=bias, gamma=gamma, out_f32=x)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_gemm_throughput_report():
    """Not an assertion on speed: prints achieved TFLOP/s so the gpurun log shows it."""
    from foundpose_b200 import _native

    for (m, n, k) in [(57664, 3072, 1024), (57664, 1024, 1024), (57664, 4096, 1024), (57664, 1024, 4096)]:
        a = _rand((m, k), 11)
        b = _rand((n, k), 12, 0.03)
        o16 = torch.empty((m, n), device="cuda", dtype=torch.float16)
        for _ in range(3):
            _native.gemm_tn_f16(a, b, _native.EPI_BIAS_F16, out_f16=o16)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        iters = 10
        for _ in range(iters):
            _native.gemm_tn_f16(a, b, _native.EPI_BIAS_F16, out_f16=o16)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        print(f"GEMM {m}x{n}x{k}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s")
        ref = torch.matmul(a, b.t())
        for _ in range(3):
            torch.matmul(a, b.t(), out=ref)
        s.record()
        for _ in range(iters):
            torch.matmul(a, b.t(), out=ref)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        print(f"  cuBLAS fp16 same shape: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s")


def test_residual_epilogue_modes_agree():
    """The two residual epilogues (TMA reduce-add through the L2 vs TMA load + add + store) compute the same
    x += gamma * (A B^T + bias); they differ by at most the rounding of one fused multiply-add."""
    import ctypes

    from foundpose_b200 import _native

    lib = _native.load()
    m, n, k = 3000, 512, 256            # pair kernel (N % 256 == 0, M > 256), ragged last row tile
    a, b = _rand((m, k), 21), _rand((n, k), 22, 0.05)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(23)).cuda()
    gamma = torch.randn(n, generator=torch.Generator().manual_seed(24)).cuda()
    x0 = torch.randn(m, n, generator=torch.Generator().manual_seed(25)).cuda()
    outs = []
    try:
        for mode in (0, 1):
            lib.fp_gemm_force_1sm(ctypes.c_int(mode << 1))
            x = x0.clone()
            _native.gemm_tn_f16(a, b, _native.EPI_RESID_F32, bias=bias, gamma=gamma, out_f32=x)
            torch.cuda.synchronize()
            outs.append(x)
    finally:
        lib.fp_gemm_force_1sm(ctypes.c_int(0))
    ref = x0 + gamma * (a.float() @ b.float().t() + bias)
    assert (outs[0] - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    assert (outs[1] - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    assert (outs[0] - outs[1]).abs().max().item() <= 4e-6 * ref.abs().max().item()
