"""tcgen05 GEMM (dlimgedit_b200/csrc/kernels/gemm.cu) against a plain PyTorch fp32 reference of the same op.

Tolerances: operands are bf16 (or tf32) with fp32 accumulation, so the reference is computed in fp32 from the
same rounded operands; the 16-bit output adds one rounding (rel 2^-11 fp16 / 2^-8 bf16)."""
import pytest
import torch

from gpu_util import act_dtype, gemm

pytestmark = pytest.mark.gpu


def _ref(a, b, bias, residual, act):
    y = a.float() @ b.float().t()
    if bias is not None:
        y = y + bias
    if residual is not None:
        y = y + residual.float()
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    return y


SHAPES = [  # (M, N, K): the encoder's real shapes incl. K tails (160, 288, 320) and M tails
    (256, 64, 64), (1000, 256, 64), (4096, 64, 256), (777, 128, 64), (513, 160, 128), (4900, 480, 160),
    (4096, 160, 640), (640, 64, 288), (4900, 960, 320), (4096, 1280, 320), (4096, 320, 1280), (4096, 256, 2304),
    (17689, 384, 128), (100, 512, 128), (300, 16, 64),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_f16_gemm_plain(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(act_dtype())
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(act_dtype())
    bias = torch.randn(N, device="cuda", generator=g)
    ref = _ref(a, b, bias, None, 0)
    out = gemm(a, b, bias=bias, out_f32=True)
    assert torch.allclose(out, ref, atol=2e-3, rtol=2e-3), float((out - ref).abs().max())
    out16 = gemm(a, b, bias=bias)
    assert torch.allclose(out16.float(), ref, atol=2e-2, rtol=1e-2)
    simt = gemm(a, b, bias=bias, out_f32=True, simt=True)
    assert torch.allclose(simt, ref, atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("act", [0, 1, 2])
def test_f16_gemm_epilogues(act):
    g = torch.Generator(device="cuda").manual_seed(act)
    M, N, K = 1234, 320, 160
    a = torch.randn(M, K, device="cuda", generator=g).to(act_dtype())
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(act_dtype())
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(act_dtype())
    ref = _ref(a, b, bias, res, act)
    out = gemm(a, b, bias=bias, residual=res, act=act)
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=1e-2), float((out.float() - ref).abs().max())
    # no bias
    out = gemm(a, b, act=act, out_f32=True)
    assert torch.allclose(out, _ref(a, b, None, None, act), atol=2e-3, rtol=2e-3)


def test_f16_gemm_row_scatter_in_place_residual():
    """proj epilogue of a TinyViT block: windowed rows scatter to token rows, residual read in place, padding dropped."""
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K, rows = 900, 128, 128, 700
    a = torch.randn(M, K, device="cuda", generator=g).to(act_dtype())
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(act_dtype())
    bias = torch.randn(N, device="cuda", generator=g)
    perm = torch.randperm(M, device="cuda", generator=g)
    row_map = torch.full((M,), -1, device="cuda", dtype=torch.int32)
    row_map[perm[:rows]] = torch.arange(rows, device="cuda", dtype=torch.int32)
    x = torch.randn(rows, N, device="cuda", generator=g).to(act_dtype())
    ref = x.float().clone()
    full = _ref(a, b, bias, None, 0)
    ref[row_map[perm[:rows]].long()] += full[perm[:rows]]
    out = x.clone()
    gemm(a, b, bias=bias, residual=out, row_map=row_map, out=out)
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("M,N,K", [(4096, 128, 256), (4096, 256, 128), (16384, 128, 64), (8192, 256, 256), (333, 128, 256)])
def test_tf32_gemm(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.double() @ b.double().t() + bias.double()
    out = gemm(a, b, bias=bias, out_f32=True)
    # tf32 operands: 10-bit mantissa -> relative 2^-11 per product, accumulated in fp32
    err = float((out.double() - ref).abs().max())
    assert err < 5e-3, err
    out_g = gemm(a, b, bias=bias, act=1, out_f32=True)
    assert torch.allclose(out_g.double(), torch.nn.functional.gelu(ref), atol=5e-3)
    simt = gemm(a, b, bias=bias, out_f32=True, simt=True)
    assert torch.allclose(simt.double(), ref, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("M,N,K,ks", [(448, 256, 2048, 8), (7, 256, 2048, 8), (1792, 256, 2048, 8), (130, 64, 512, 4)])
def test_tf32_gemm_split_k(M, N, K, ks):
    """Split-K (the decoder token MLP's 2048 -> 256 Linear): `ks` partial results in row blocks of M rounded up to 128."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    mpad = (M + 127) // 128 * 128
    out = torch.full((ks * mpad, N), float("nan"), device="cuda")
    gemm(a, b, out_f32=True, simt=ks, out=out)
    parts = out.view(ks, mpad, N)
    assert torch.isfinite(parts[:, :M]).all()
    kk = K // ks
    for s in range(ks):  # every part is the product over its own k range
        ref = a[:, s * kk:(s + 1) * kk].double() @ b[:, s * kk:(s + 1) * kk].double().t()
        assert float((parts[s, :M].double() - ref).abs().max()) < 5e-3
    whole = gemm(a, b, out_f32=True)
    assert float((parts[:, :M].sum(0) - whole).abs().max()) < 1e-3


def test_gemm_back_to_back_is_deterministic():
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(65536, 64, device="cuda", generator=g).to(act_dtype())
    b = torch.randn(256, 64, device="cuda", generator=g).to(act_dtype())
    o1 = gemm(a, b)
    o2 = gemm(a, b)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("M,N,K,act", [(4096, 480, 160, 0), (1000, 512, 128, 1), (4096, 1280, 320, 1), (333, 384, 128, 0)])
def test_f16_gemm_folded_layernorm(M, N, K, act):
    """qkv / fc1 with the preceding LayerNorm folded in: A is the raw rows, gamma is in the weights, beta in the bias,
    the weight rows are centred, and the epilogue applies rstd * acc + bias.  Reference: LayerNorm then Linear in fp32."""
    import dlimgedit_b200 as dl
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = (torch.randn(M, K, device="cuda", generator=g) * 1.5 + 0.7).to(act_dtype())
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.2 * torch.randn(K, device="cuda", generator=g)
    beta = 0.3 * torch.randn(K, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ w.t() + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    wg = w * gamma
    w16 = (wg - wg.mean(1, keepdim=True)).to(act_dtype())  # centred rows: the mean of x cancels inside the GEMM
    bias = b + w @ beta
    stats = torch.zeros(M, 2, device="cuda")
    r = dl.debug().layernorm_stats(None, x.data_ptr(), M, K, 1e-5, stats.data_ptr())
    assert r == 0, dl.api().last_error()
    torch.cuda.synchronize()
    xf = x.float()
    assert torch.allclose(stats[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], (xf.var(1, unbiased=False) + 1e-5).rsqrt(), atol=1e-5, rtol=1e-4)
    out = gemm(x, w16, bias=bias, act=act, ln_stats=stats)
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2), float((out.float() - ref).abs().max())
    simt = gemm(x, w16, bias=bias, act=act, ln_stats=stats, simt=True, out_f32=True)
    assert torch.allclose(simt, ref, atol=1e-2, rtol=1e-2), float((simt - ref).abs().max())


@pytest.mark.parametrize("M,C,K", [(4096, 160, 640), (1000, 128, 512), (4096, 320, 1280)])
def test_f16_gemm_row_statistics_feed_the_next_folded_layernorm(M, C, K):
    """fc2 (+ residual) writes per-row partial (sum, sum of squares) of its output next to the output itself; the
    following qkv GEMM turns them into 1/std in its epilogue.  C = 320 runs as two N tiles -> two partial sums per row."""
    g = torch.Generator(device="cuda").manual_seed(M + C)
    h = torch.randn(M, K, device="cuda", generator=g).to(act_dtype())
    w2 = (torch.randn(C, K, device="cuda", generator=g) / K ** 0.5).to(act_dtype())
    b2 = torch.randn(C, device="cuda", generator=g)
    res = (torch.randn(M, C, device="cuda", generator=g) + 0.5).to(act_dtype())
    parts = 2 if C == 320 else 1
    stats = torch.zeros(M, parts, 2, device="cuda")
    x = gemm(h, w2, bias=b2, residual=res, stats_out=stats)
    ref_x = h.float() @ w2.float().t() + b2 + res.float()
    assert torch.allclose(x.float(), ref_x, atol=3e-2, rtol=1e-2)
    s = stats.sum(1)
    assert torch.allclose(s[:, 0], ref_x.sum(1), atol=2e-2, rtol=1e-3)
    assert torch.allclose(s[:, 1], (ref_x * ref_x).sum(1), atol=5e-2, rtol=2e-3)
    # consumer: LayerNorm folded into the next Linear, statistics taken from the partial sums
    N = 3 * C
    w = torch.randn(N, C, device="cuda", generator=g) / C ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    wg = w * gamma
    w16 = (wg - wg.mean(1, keepdim=True)).to(act_dtype())
    out = gemm(x, w16, bias=b + w @ beta, ln_stats=stats, ln_parts=parts)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ w.t() + b
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=1e-2), float((out.float() - ref).abs().max())


@pytest.mark.parametrize("M,C", [(4096, 160), (1000, 128), (40000, 160), (300, 128)])
def test_fused_mlp(M, C):
    """fc1 (+ folded LayerNorm, GELU) -> hidden activation in TMEM / shared memory -> fc2 + residual, one kernel.
    Reference: plain PyTorch fp32 LayerNorm / Linear / GELU / Linear on the same 16-bit input."""
    import dlimgedit_b200 as dl
    g = torch.Generator(device="cuda").manual_seed(M + C)
    H = 4 * C
    x = (torch.randn(M, C, device="cuda", generator=g) * 1.3 + 0.4).to(act_dtype())
    w1 = torch.randn(H, C, device="cuda", generator=g) / C ** 0.5
    b1 = 0.2 * torch.randn(H, device="cuda", generator=g)
    w2 = torch.randn(C, H, device="cuda", generator=g) / H ** 0.5
    b2 = 0.2 * torch.randn(C, device="cuda", generator=g)
    gamma = 1 + 0.2 * torch.randn(C, device="cuda", generator=g)
    beta = 0.3 * torch.randn(C, device="cuda", generator=g)
    xf = x.float()
    hid = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xf, (C,), gamma, beta, 1e-5) @ w1.t() + b1)
    ref = xf + hid @ w2.t() + b2
    wg = w1 * gamma
    w1c = (wg - wg.mean(1, keepdim=True)).to(act_dtype()).contiguous()
    b1f = (b1 + w1 @ beta).contiguous()
    sums = torch.stack([xf.sum(1), (xf * xf).sum(1)], 1).contiguous()
    out = torch.zeros(M, C, device="cuda", dtype=act_dtype())
    stats = torch.zeros(M, 2, device="cuda")
    r = dl.debug().mlp_fused(None, x.data_ptr(), M, C, w1c.data_ptr(), b1f.data_ptr(), sums.data_ptr(),
                             w2.to(act_dtype()).contiguous().data_ptr(), b2.data_ptr(), out.data_ptr(), stats.data_ptr())
    assert r == 0, dl.api().last_error()
    torch.cuda.synchronize()
    assert torch.allclose(out.float(), ref, atol=4e-2, rtol=1.5e-2), float((out.float() - ref).abs().max())
    assert torch.allclose(stats[:, 0], ref.sum(1), atol=5e-2, rtol=2e-3)
    assert torch.allclose(stats[:, 1], (ref * ref).sum(1), atol=0.2, rtol=4e-3)
    # in place (out aliases x), as the engine calls it
    xi = x.clone()
    r = dl.debug().mlp_fused(None, xi.data_ptr(), M, C, w1c.data_ptr(), b1f.data_ptr(), sums.data_ptr(),
                             w2.to(act_dtype()).contiguous().data_ptr(), b2.data_ptr(), xi.data_ptr(), None)
    assert r == 0, dl.api().last_error()
    torch.cuda.synchronize()
    assert torch.equal(xi, out)


@pytest.mark.parametrize("B,H,W,C,N", [(2, 64, 64, 256, 256), (3, 8, 32, 64, 48), (1, 6, 128, 128, 160)])
def test_conv3x3_implicit_gemm(B, H, W, C, N):
    """3x3 convolution with zero padding as an implicit GEMM (4D TMA boxes shifted by the tap, out-of-image elements
    zero-filled by TMA) against torch.nn.functional.conv2d in fp32 on the same 16-bit inputs."""
    import dlimgedit_b200 as dl
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(act_dtype())
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).to(act_dtype())
    b = 0.3 * torch.randn(N, device="cuda", generator=g)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    wk = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()  # K index = (ky * 3 + kx) * C + c
    out = torch.zeros(B * H * W, N, device="cuda", dtype=act_dtype())
    r = dl.debug().conv3x3(None, x.data_ptr(), B, H, W, C, wk.data_ptr(), b.data_ptr(), N, out.data_ptr())
    assert r == 0, dl.api().last_error()
    torch.cuda.synchronize()
    assert torch.allclose(out.float().view(B, H, W, N), ref, atol=2e-2, rtol=1e-2), float((out.float().view(B, H, W, N) - ref).abs().max())


@pytest.mark.parametrize("offset", [5.0, 60.0, 200.0])
def test_folded_layernorm_with_large_row_means(offset):
    """the folded LayerNorm forms var = E[x^2] - mean^2 from fp32 row sums and relies on the centred weight rows
    summing to zero AFTER their rounding to 16 bits.  Rows whose mean is 5..200 standard deviations away from zero are the
    case where both would break (real TinyViT tokens are not zero-mean).  Weights are prepared the way the engine loads
    them (csrc/model.cu load_linear16_ln: the rounding residual of each row is folded back into its smallest element)."""
    M, N, K = 2048, 480, 160
    g = torch.Generator(device="cuda").manual_seed(int(offset))
    x = (torch.randn(M, K, device="cuda", generator=g) + offset).to(act_dtype())
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.2 * torch.randn(K, device="cuda", generator=g)
    beta = 0.3 * torch.randn(K, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (K,), gamma.double(), beta.double(), 1e-5) @ w.double().t() + b.double()
    wg = (w * gamma).double()
    w16 = (wg - wg.mean(1, keepdim=True)).to(act_dtype())
    for _ in range(2):  # fold the rounding residual of every row into its smallest element
        r = w16.double().sum(1)
        j = w16.abs().argmin(1)
        rows = torch.arange(N, device="cuda")
        w16[rows, j] = (w16[rows, j].double() - r).to(act_dtype())
    assert float(w16.double().sum(1).abs().max()) < 1e-5
    bias = b + w @ beta
    # row sums as the producing GEMM's epilogue leaves them: fp32 (sum, sum of squares), one part
    xf = x.float()
    stats = torch.stack([xf.sum(1), (xf * xf).sum(1)], 1).view(M, 1, 2).contiguous()
    out = gemm(x, w16, bias=bias, ln_stats=stats, ln_parts=1)
    err = float((out.double() - ref).abs().max())
    print(f"row mean / std = {offset}: max abs error {err:.4f} on unit-scale outputs")
    assert err < 4e-2  # the 16-bit rounding of the output alone is ~4e-3 at |y| ~ 4
