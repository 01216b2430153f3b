"""GPU: the tcgen05 tensor-core back end (fp16 hi/lo split operands, three MMAs per product, fp32
accumulators in TMEM) against fp64 torch restatements.  The bar is fp32-grade accuracy: the error
must be of the order of fp32 summation noise, not of fp16/TF32 rounding (which would be ~1e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
RTOL = 4e-6  # (scaled up for K > 2048) relative to max |reference|; single-pass fp16/TF32 would sit near 5e-4


def _ops():
    from mage_b200 import ops
    return ops


@pytest.fixture(params=[("auto", 0, -1, True), ("pair256", 256, 1, True), ("pair192", 192, 1, True), ("pair128", 128, 1, True), ("pair64", 64, 1, True),
                        ("single", 0, 0, True), ("auto-nohalo", 0, -1, False), ("pair256-nohalo", 256, 1, False),
                        ("single-nohalo", 0, 0, False), ("nsplit", 0, -1, True), ("auto-no-nsplit", 0, -1, True)],
                ids=lambda p: p[0], autouse=True)
def tile_cfg(request):
    """Every test runs under each tile selection: automatic, CTA-pair (cta_group::2) with 256/128/64-wide N tiles where the
    shape allows it, single-CTA tiles only -- and the convolutions with and without halo reuse."""
    ops = _ops()
    ops.tc_tuning(request.param[1], request.param[2])
    ops.tc_conv_halo(request.param[3])
    ops.tc_nsplit(2 if request.param[0] == "nsplit" else 0 if request.param[0] == "auto-no-nsplit" else 1)
    yield request.param[0]
    ops.tc_tuning(0, -1)
    ops.tc_conv_halo(True)
    ops.tc_nsplit(1)


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _unsplit(s):
    s = s.detach().cpu().double()
    return s[0] + s[1] / 2048.0


def _close(got, want, rtol=RTOL):
    got = got.detach().cpu().double()
    want = want.double()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= rtol * ref, f"max abs err {err:.3e} vs ref magnitude {ref:.3e} (ratio {err / ref:.2e})"


def _act(x, act):
    return [lambda t: t, F.relu, lambda t: t * torch.sigmoid(1.702 * t), F.gelu, torch.tanh][act](x)


def test_split_roundtrip_and_flag():
    ops = _ops()
    x = torch.cat([_rand(64, 512, seed=1), _rand(64, 512, seed=2, scale=1e-3), _rand(64, 512, seed=3, scale=300.0)])
    s = ops.split(x.to(DEV))
    assert s.shape == (2, 192, 512) and s.dtype == torch.float16
    err = (_unsplit(s) - x.double()).abs()
    assert (err <= 2.0 ** -23 * x.double().abs() + 1e-10).all(), f"split error {err.max():.3e}"
    r = ops.split(x.to(DEV), relu=True)
    assert (_unsplit(r) - F.relu(x).double()).abs().max() <= 2.0 ** -23 * x.abs().max()
    ops.check_flag(DEV)  # in range: no complaint
    ops.split(torch.full((4, 8), 1e5, device=DEV))
    with pytest.raises(RuntimeError):
        ops.check_flag(DEV)
    ops.check_flag(DEV)  # flag was cleared


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 512, 512), (4096, 1536, 512), (2048, 1536, 512), (2048, 2048, 512), (2048, 512, 2048),
                                   (1024, 1536, 512), (1000, 2048, 512), (513, 512, 2048),
                                   (1000, 64, 576), (20000, 128, 64), (40, 1024, 256), (16384, 512, 512), (2560, 256, 1024)])
def test_gemm_tc_shapes(M, N, K):
    ops = _ops()
    a, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    out, _, _ = ops.gemm_tc(ops.split(a.to(DEV)), ops.split(w.to(DEV)), b.to(DEV))
    torch.cuda.synchronize()
    _close(out, a.double() @ w.double().t() + b.double())
    ops.check_flag(DEV)


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_gemm_tc_epilogues(act):
    ops = _ops()
    M, N, K = 700, 512, 256
    a, w, b, r = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3), _rand(M, N, seed=4)
    asp, wsp = ops.split(a.to(DEV)), ops.split(w.to(DEV))
    want = _act(a.double() @ w.double().t() + b.double(), act) + r.double()
    out, sp, spr = ops.gemm_tc(asp, wsp, b.to(DEV), residual=r.to(DEV), act=act, want=("f32", "split", "split_relu"))
    _close(out, want)
    _close(_unsplit(sp), want)
    _close(_unsplit(spr), F.relu(want))
    # activation after the residual, residual read through a ReLU; split-only output
    want2 = _act(a.double() @ w.double().t() + b.double() + F.relu(r).double(), act)
    out2, sp2, _ = ops.gemm_tc(asp, wsp, b.to(DEV), residual=r.to(DEV), act=act | 0x100 | 0x200, want=("split",))
    assert out2 is None
    _close(_unsplit(sp2), want2)


def test_gemm_tc_inplace_residual_and_resmod():
    ops = _ops()
    M, N, K = 512, 512, 512
    a, w, x = _rand(M, K, seed=5), _rand(N, K, seed=6, scale=K ** -0.5), _rand(M, N, seed=7)
    asp, wsp = ops.split(a.to(DEV)), ops.split(w.to(DEV))
    xd = x.to(DEV).clone()
    ops.gemm_tc(asp, wsp, None, residual=xd, out=xd)  # x += a @ w.T in place
    _close(xd, a.double() @ w.double().t() + x.double())
    tab = _rand(256, N, seed=8)
    out, _, _ = ops.gemm_tc(asp, wsp, None, residual=tab.to(DEV), res_mod=256)
    _close(out, a.double() @ w.double().t() + tab.double().repeat(2, 1))


@pytest.mark.parametrize("M,K", [(128, 512), (2048, 512), (2048, 2048), (389, 512), (16384, 512), (4096 + 128, 2048), (16384, 2048)])
def test_gemm_tc_with_fused_layernorm_is_bit_identical_to_the_two_kernels(M, K):
    """mage_gemm_tc_ln (the residual-stream GEMM whose epilogue also LayerNorms the finished 128-row blocks, whichever CTA completes
    them) against mage_gemm_tc followed by mage_layernorm_f32: the fp32 result and the split LayerNorm output must be the SAME
    BITS under every tile selection -- including ragged M, an odd number of row tiles, several launches on the same counters
    (they must come back to zero) and in-place residual."""
    ops = _ops()
    a, w, x = _rand(M, K, seed=15), _rand(512, K, seed=16, scale=K ** -0.5), _rand(M, 512, seed=17)
    bias, gamma, beta = _rand(512, seed=18).to(DEV), (1 + 0.1 * _rand(512, seed=19)).to(DEV), (0.1 * _rand(512, seed=20)).to(DEV)
    asp, wsp = ops.split(a.to(DEV)), ops.split(w.to(DEV))
    want_x = x.to(DEV).clone()
    ops.gemm_tc(asp, wsp, bias, residual=want_x, out=want_x)
    want_u = torch.empty(2, M, 512, device=DEV, dtype=torch.float16)
    ops.layernorm(want_x, gamma, beta, out_split=want_u)
    counters = torch.zeros((M + 127) // 128, device=DEV, dtype=torch.int32)
    for it in range(3):
        got_x = x.to(DEV).clone()
        got_u = torch.full((2, M, 512), float("nan"), device=DEV, dtype=torch.float16)
        ops.gemm_tc_ln(asp, wsp, bias, residual=got_x, out=got_x, gamma=gamma, beta=beta, ln_out=got_u, counters=counters)
        torch.cuda.synchronize()
        assert torch.equal(got_x, want_x), it
        assert torch.equal(got_u.view(torch.int16), want_u.view(torch.int16)), (it, (got_u.float() - want_u.float()).abs().max().item())
        assert int(counters.abs().sum()) == 0, it
    ops.check_flag(DEV)
    _close(want_x, a.double() @ w.double().t() + bias.cpu().double() + x.double())


def test_gemm_tc_matches_simt_kernel_to_fp32_noise():
    ops = _ops()
    M, N, K = 2048, 512, 2048
    a, w = _rand(M, K, seed=11), _rand(N, K, seed=12, scale=K ** -0.5)
    want = a.double() @ w.double().t()
    simt = ops.gemm(a.to(DEV), w.to(DEV))
    tc, _, _ = ops.gemm_tc(ops.split(a.to(DEV)), ops.split(w.to(DEV)))
    e_simt = (simt.cpu().double() - want).abs().max().item()
    e_tc = (tc.cpu().double() - want).abs().max().item()
    assert e_tc <= 4 * e_simt + 1e-7, f"tensor-core error {e_tc:.3e} vs fp32 FFMA error {e_simt:.3e}"


def test_embedding_split():
    ops = _ops()
    table = _rand(512, 1024, seed=3)
    idx = torch.randint(0, 512, (3, 16, 16), generator=torch.Generator().manual_seed(4))
    tsp = ops.split(table.to(DEV))
    got = ops.embedding_split(idx.to(DEV), tsp)
    assert got.shape == (2, 3, 16, 16, 1024)
    assert torch.equal(got.cpu(), tsp.cpu()[:, idx.reshape(-1)].reshape(2, 3, 16, 16, 1024))


def _conv_ref(x, w, b, pad):
    """x [n,H,W,C] NHWC, w [Cout,KH,KW,Cin] -> NHWC fp64."""
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double() if b is not None else None, padding=pad)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("n,H,Cin,Cout,k", [(3, 16, 64, 64, 3), (2, 16, 128, 512, 3), (2, 32, 64, 256, 3), (1, 64, 64, 64, 3),
                                            (1, 128, 64, 64, 3), (2, 16, 512, 512, 3), (2, 32, 256, 128, 1)])
def test_conv2d_tc_plain(n, H, Cin, Cout, k):
    ops = _ops()
    x = _rand(n, H, H, Cin, seed=1)
    w = _rand(Cout, k, k, Cin, seed=2, scale=(k * k * Cin) ** -0.5)
    b = _rand(Cout, seed=3)
    out, sp, _ = ops.conv2d_tc(ops.split(x.to(DEV)), ops.split(w.to(DEV)), b.to(DEV), pad=(k // 2, k // 2), act=1,
                               want=("f32", "split"))
    want = F.relu(_conv_ref(x, w, b, k // 2))
    rtol = RTOL * max(1.0, k * k * Cin / 2048)  # the tensor core's fp32 accumulator rounds once per 16-deep MMA: error grows ~K
    _close(out, want, rtol)
    _close(_unsplit(sp), want, rtol)
    ops.check_flag(DEV)


def test_conv2d_tc_residual_modes():
    ops = _ops()
    n, H, Cin, Cout = 2, 32, 64, 128
    x, w, b = _rand(n, H, H, Cin, seed=1), _rand(Cout, 3, 3, Cin, seed=2, scale=(9 * Cin) ** -0.5), _rand(Cout, seed=3)
    xs, ws = ops.split(x.to(DEV)), ops.split(w.to(DEV))
    base = _conv_ref(x, w, b, 1)
    r1 = _rand(n, H, H, Cout, seed=4)
    out, _, _ = ops.conv2d_tc(xs, ws, b.to(DEV), pad=(1, 1), residual=r1.to(DEV), res_mode=1)
    _close(out, base + r1.double())
    r2 = _rand(n, H // 2, H // 2, Cout, seed=5)  # stored at half resolution, read through nearest x2
    out, _, _ = ops.conv2d_tc(xs, ws, b.to(DEV), pad=(1, 1), residual=r2.to(DEV), res_mode=2)
    up = r2.double().repeat_interleave(2, 1).repeat_interleave(2, 2)
    _close(out, base + up)
    r3 = _rand(H * H, Cout, seed=6)  # one map shared by all images (H/W positional embeddings)
    out, _, _ = ops.conv2d_tc(xs, ws, None, pad=(1, 1), residual=r3.to(DEV), res_mode=3)
    _close(out, _conv_ref(x, w, None, 1) + r3.double().view(1, H, H, Cout))


def test_conv2d_tc_transpose_phases():
    """ConvTranspose2d(4,2,1) as four 2x2 sub-pixel phase convolutions scattered into the full output."""
    ops = _ops()
    n, H, Cin, Cout = 2, 16, 256, 256
    x = _rand(n, H, H, Cin, seed=1)
    wt = _rand(Cin, Cout, 4, 4, seed=2, scale=(4 * Cin) ** -0.5)  # ConvTranspose2d layout
    b = _rand(Cout, seed=3)
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    xs = ops.split(x.to(DEV))
    out = torch.empty(n, 2 * H, 2 * H, Cout, device=DEV)
    taps = {0: (3, 1), 1: (2, 0)}
    for py in (0, 1):
        for px in (0, 1):
            sub = wt[:, :, list(taps[py]), :][:, :, :, list(taps[px])].permute(1, 2, 3, 0).contiguous()  # [Cout,2,2,Cin]
            ops.conv2d_tc(xs, ops.split(sub.to(DEV)), b.to(DEV), pad=(1 - py, 1 - px), out=out, out_hw=(H, H),
                          scatter=(2, 2, py, px), full_hw=(2 * H, 2 * H))
    _close(out, want)


def test_tc_rejects_unsupported_shapes():
    from mage_b200 import _lib
    ops = _ops()
    a, w = ops.split(_rand(64, 96, seed=1).to(DEV)), ops.split(_rand(64, 96, seed=2).to(DEV))
    with pytest.raises(_lib.MageCudaError):
        ops.gemm_tc(a, w)  # K % 64 != 0 -> MAGE_ENOTSUP, never a silent fallback


@pytest.mark.parametrize("n,H,res_mode", [(2, 128, 2), (3, 16, 1), (2, 32, 2)])
def test_conv2d_tc_pixel_head_matches_unfused(n, H, res_mode):
    """The fused last decoder layer (3x3 conv 64->256 + skip -> ReLU -> 1x1 to 3 channels -> tanh, vqvae_model.py:210-213)
    against the unfused kernels and an fp64 restatement."""
    ops = _ops()
    Cin, Cout = 64, 256
    x, w, b = _rand(n, H, H, Cin, seed=1), _rand(Cout, 3, 3, Cin, seed=2, scale=(9 * Cin) ** -0.5), _rand(Cout, seed=3)
    Hr = H // 2 if res_mode == 2 else H
    r = _rand(n, Hr, Hr, Cout, seed=4)
    hw, hb = _rand(3, Cout, seed=5, scale=Cout ** -0.5), _rand(3, seed=6, scale=0.1)
    xs, ws_ = ops.split(x.to(DEV)), ops.split(w.to(DEV))
    out = torch.full((n, 5, 3, H, H), 7.0, device=DEV)           # frame slot 2 of a [n, L=5, 3, H, W] clip
    ops.conv2d_tc_pixel_head(xs, ws_, b.to(DEV), pad=(1, 1), residual=r.to(DEV), res_mode=res_mode, head_w=hw.to(DEV),
                             head_b=hb.to(DEV), out=out[:, 2], out_img_stride=5 * 3 * H * H)
    torch.cuda.synchronize()
    conv = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), padding=1)
    res = r.permute(0, 3, 1, 2).double()
    if res_mode == 2:
        res = F.interpolate(res, scale_factor=2, mode="nearest")
    want = torch.tanh(F.conv2d(F.relu(conv + res), hw.double().view(3, Cout, 1, 1), hb.double()))
    _close(out[:, 2], want, rtol=2e-6)
    assert (out[:, 1] == 7.0).all() and (out[:, 3] == 7.0).all()   # neighbouring frames untouched
    full, _, _ = ops.conv2d_tc(xs, ws_, b.to(DEV), pad=(1, 1), residual=r.to(DEV), res_mode=res_mode)
    ref = torch.empty(n, 3, H, H, device=DEV)
    ops.conv1x1_tanh_nchw(full, hw.to(DEV), hb.to(DEV), ref, 3 * H * H)
    _close(out[:, 2], ref.cpu(), rtol=2e-6)


@pytest.mark.parametrize("n,H,W,Cin,Cout,k,pad", [(3, 16, 8, 64, 64, 3, (1, 1)), (1, 32, 24, 128, 192, 3, (1, 1)), (2, 16, 16, 64, 256, 2, (1, 0)),
                                                 (2, 48, 40, 64, 64, 2, (0, 1)), (5, 16, 16, 512, 512, 3, (1, 1))])
def test_conv2d_tc_halo_geometries(n, H, W, Cin, Cout, k, pad, tile_cfg):
    """Non-square maps, odd tile counts (no CTA pairing possible), 2x2 phase taps with one-sided padding, many channel blocks."""
    ops = _ops()
    if "nohalo" in tile_cfg and 128 % W != 0:
        from mage_b200 import _lib
        with pytest.raises(_lib.MageCudaError):   # the per-tap-box kernel tiles rows of 128 / W pixels: loud refusal, no fallback
            ops.conv2d_tc(ops.split(_rand(n, H, W, Cin).to(DEV)), ops.split(_rand(Cout, k, k, Cin).to(DEV)), None, pad=pad)
        return
    x = _rand(n, H, W, Cin, seed=1)
    w = _rand(Cout, k, k, Cin, seed=2, scale=(k * k * Cin) ** -0.5)
    b = _rand(Cout, seed=3)
    Ho, Wo = (H, W)
    out, _, _ = ops.conv2d_tc(ops.split(x.to(DEV)), ops.split(w.to(DEV)), b.to(DEV), pad=pad, act=1, out_hw=(Ho, Wo))
    xp = x.permute(0, 3, 1, 2).double()
    if k == 2:   # out[y,x] = sum in[y - pad_y + ky, x - pad_x + kx] w[ky,kx]: pad before = pad, after = 1 - pad
        xp = F.pad(xp, (pad[1], 1 - pad[1], pad[0], 1 - pad[0]))
        want = F.conv2d(xp, w.permute(0, 3, 1, 2).double(), b.double())
    else:
        want = F.conv2d(xp, w.permute(0, 3, 1, 2).double(), b.double(), padding=pad)
    want = F.relu(want).permute(0, 2, 3, 1)
    _close(out, want, RTOL * max(1.0, k * k * Cin / 2048))
    ops.check_flag(DEV)


@pytest.mark.parametrize("n,C,H,W,k", [(2, 3, 128, 128, 7), (3, 1, 32, 16, 5), (1, 3, 16, 8, 3)])
def test_first_conv_as_im2row_plus_column_conv(n, C, H, W, k, tile_cfg):
    """encoder.0 (vqvae_model.py:193: Conv2d(C, dim, 7, padding=3) on a planar image) as mage_patch_rows_split_f32 followed by a
    kx1 tensor-core convolution over 64 im2row channels."""
    ops = _ops()
    if "nohalo" in tile_cfg and 128 % W != 0:
        pytest.skip("per-tap-box kernel tiles rows of 128 / W pixels")
    Cout = 256
    x = _rand(n, C, H, W, seed=1)
    w = _rand(Cout, C, k, k, seed=2, scale=(k * k * C) ** -0.5)
    b = _rand(Cout, seed=3)
    rows = ops.patch_rows_split(x.to(DEV), k, k // 2)
    want_rows = torch.zeros(n, H, W, 64, dtype=torch.float64)
    xp = F.pad(x.double(), (k // 2, k // 2))
    for kx in range(k):
        for c in range(C):
            want_rows[..., kx * C + c] = xp[:, c, :, kx:kx + W]
    _close(_unsplit(rows), want_rows, rtol=2.0 ** -22)
    w2 = torch.zeros(Cout, k, 1, 64)
    w2[:, :, 0, : k * C] = w.permute(0, 2, 3, 1).reshape(Cout, k, k * C)
    out, _, spr = ops.conv2d_tc(rows, ops.split(w2.to(DEV)), b.to(DEV), pad=(k // 2, 0), want=("f32", "split_relu"))
    want = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2).permute(0, 2, 3, 1)
    _close(out, want)
    _close(_unsplit(spr), F.relu(want))


@pytest.mark.parametrize("n,H,C,Cout", [(2, 32, 256, 256), (3, 32, 64, 128)])
def test_strided_conv_as_s2d_valid_conv(n, H, C, Cout, tile_cfg):
    """vqvae_model.py:175 (Conv2d 4x4, stride 2, pad 1) on the tensor cores: padded space-to-depth (mage_s2d_pad_split_f32) followed
    by a 2x2 stride-1 VALID convolution with the re-indexed weights -- the same products, each weight used once."""
    ops = _ops()
    x = F.relu(_rand(n, C, H, H, seed=1))
    w = _rand(Cout, C, 4, 4, seed=2, scale=(16 * C) ** -0.5)
    b = _rand(Cout, seed=3)
    want = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    s = ops.s2d_pad_split(x.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert tuple(s.shape) == (2, n, H // 2 + 1, H // 2 + 1, 4 * C)
    # s2d[n, Y, X, (py*2+px)*C + c] == x[n, c, 2Y+py-1, 2X+px-1]
    xp = F.pad(x, (1, 1, 1, 1)).double()
    sd = _unsplit(s).view(n, H // 2 + 1, H // 2 + 1, 2, 2, C)
    for py in (0, 1):
        for px in (0, 1):
            ref = xp[:, :, py::2, px::2][:, :, : H // 2 + 1, : H // 2 + 1].permute(0, 2, 3, 1)
            assert (sd[:, :, :, py, px] - ref).abs().max() <= 2.0 ** -22 * x.abs().max()
    w2 = w.view(Cout, C, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(Cout, 2, 2, 4 * C).contiguous()
    out, _, _ = ops.conv2d_tc(s, ops.split(w2.to(DEV)), b.to(DEV), pad=(0, 0))
    torch.cuda.synchronize()
    ops.check_flag(DEV)
    assert tuple(out.shape) == (n, H // 2, H // 2, Cout)
    _close(out, want)


@pytest.mark.parametrize("n,H,Cin,Cout", [(4, 32, 64, 64), (2, 64, 64, 256)])
def test_conv2d_tc_single_pass_is_fp16_grade_not_fp32_grade(n, H, Cin, Cout, tile_cfg):
    """passes=1 (decoder precision budget): hi*hi products only.  The result must equal the convolution of the fp16-ROUNDED
    operands to fp32-accumulation noise (i.e. the mode really drops the two correction MMAs and nothing else), and sit at fp16
    distance (~1e-3 .. 1e-4) from the exact product -- while passes=3 stays fp32-grade."""
    ops = _ops()
    x, w, b = _rand(n, H, H, Cin, seed=1), _rand(Cout, 3, 3, Cin, seed=2, scale=(9 * Cin) ** -0.5), _rand(Cout, seed=3)
    xs, ws = ops.split(x.to(DEV)), ops.split(w.to(DEV))
    exact = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), padding=1).permute(0, 2, 3, 1)
    rounded = F.conv2d(x.half().double().permute(0, 3, 1, 2), w.half().double().permute(0, 3, 1, 2), b.double(), padding=1).permute(0, 2, 3, 1)
    o3, _, _ = ops.conv2d_tc(xs, ws, b.to(DEV), pad=(1, 1))
    o1, _, _ = ops.conv2d_tc(xs, ws, b.to(DEV), pad=(1, 1), passes=1)
    torch.cuda.synchronize()
    _close(o3, exact)
    if "nohalo" in tile_cfg:
        _close(o1, exact)            # outside the halo kernel the request falls back to the fp32-grade product
        return
    _close(o1, rounded)
    err = (o1.cpu().double() - exact).abs().max().item() / exact.abs().max().item()
    assert 2e-5 < err < 3e-3, err


@pytest.mark.parametrize("n_img,axis", [(2, 1), (2, 2), (5, 1), (9, 2), (64, 2)])
def test_fused_qkv_axial_attention(n_img, axis, tile_cfg):
    """mage_qkv_axial_attn_tc (AxialAttentionBlock.attention of the H / W blocks, mage_model.py:31-33,36-47): one tcgen05 GEMM over
    the head-permuted packed in-projection whose epilogue runs softmax(q k^T / sqrt(32)) v over the 16 positions of every line.
    Against an fp64 restatement (nn.MultiheadAttention semantics on the permuted axis) and against the two-kernel form."""
    ops = _ops()
    R, H, C = 16, 16, 512
    M = n_img * R * R
    u = _rand(M, C, seed=1)
    w_in = _rand(3 * C, C, seed=2, scale=C ** -0.5)
    b_in = _rand(3 * C, seed=3, scale=0.1)
    scale = 32 ** -0.5
    qkv = (u.double() @ w_in.double().t() + b_in.double()).view(n_img, R, R, 3, H, 32)      # [img, h, w, part, head, d]
    q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]                             # [img, h, w, head, d]
    ax = 1 if axis == 1 else 2                                                                  # attended spatial dim
    mv = lambda t: t.movedim(ax, -2)                                                            # [img, other, head, S, d]
    att = torch.softmax(mv(q) @ mv(k).transpose(-1, -2) * scale, -1) @ mv(v)
    want = att.movedim(-2, ax).reshape(M, C)
    us = ops.split(u.to(DEV))
    wp, bp = ops.permute_qkv_for_axial(w_in.to(DEV), b_in.to(DEV), H)
    out = torch.empty(2, M, C, device=DEV, dtype=torch.float16)
    ops.qkv_axial_attn_tc(us, ops.split(wp), bp, out, n_img=n_img, R=R, n_head=H, axis=axis, scale=scale)
    torch.cuda.synchronize()
    ops.check_flag(DEV)
    _close(_unsplit(out), want, rtol=6e-6)
    # the two-kernel form (QKV GEMM -> fp32 [M, 3C] -> axial_attn_kernel) agrees to fp32 noise
    qkv2, _, _ = ops.gemm_tc(us, ops.split(w_in.to(DEV)), b_in.to(DEV))
    out2 = torch.empty(2, M, C, device=DEV, dtype=torch.float16)
    ops.axial_attn(qkv2, None, B=n_img, R=R, n_head=H, axis=axis, scale=scale, out_split=out2)
    assert (_unsplit(out) - _unsplit(out2)).abs().max() <= 4e-6 * want.abs().max()
