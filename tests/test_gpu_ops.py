"""GPU: every C-ABI kernel against a plain torch fp64/fp32 CPU restatement of the same op
(tolerances are fp32 summation-order noise; index outputs must match exactly away from ties)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from mage_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _close(got, want, rtol=2e-5, atol=2e-5):
    got = got.detach().cpu().double()
    want = want.double()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= atol + rtol * ref, f"max abs err {err:.3e} vs ref magnitude {ref:.3e}"


def _act(x, act):
    return [lambda t: t, F.relu, lambda t: t * torch.sigmoid(1.702 * t), F.gelu, torch.tanh][act](x)


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (4096, 1536, 512), (1000, 2048, 512), (513, 512, 2048),
                                   (77, 3, 256), (1000, 64, 576), (256, 1, 1024), (20000, 128, 64), (40, 1024, 256)])
def test_gemm_shapes(M, N, K):
    ops = _ops()
    a, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    out = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV))
    _close(out, a.double() @ w.double().t() + b.double())


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_gemm_epilogues(act):
    ops = _ops()
    M, N, K = 700, 512, 256
    a, w, b, r = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3), _rand(M, N, seed=4)
    want = _act(F.relu(a).double() @ w.double().t() + b.double(), act) + r.double()
    out = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV), residual=r.to(DEV), act=act, relu_a=True)
    _close(out, want)
    # activation after the residual, residual read through a ReLU
    want2 = _act(a.double() @ w.double().t() + b.double() + F.relu(r).double(), act)
    out2 = ops.gemm(a.to(DEV), w.to(DEV), b.to(DEV), residual=r.to(DEV), act=act | 0x100 | 0x200)
    _close(out2, want2)


def test_gemm_inplace_residual_resmod_and_strided_a():
    ops = _ops()
    M, N, K = 512, 512, 512
    big = _rand(M, 3 * K, seed=5).to(DEV)
    w, x = _rand(N, K, seed=6, scale=K ** -0.5), _rand(M, N, seed=7)
    xd = x.to(DEV).clone()
    ops.gemm(big[:, K:2 * K], w.to(DEV), None, residual=xd, out=xd)  # x += A_view @ W^T
    _close(xd, big[:, K:2 * K].cpu().double() @ w.double().t() + x.double())
    tab = _rand(256, N, seed=8)
    out = ops.gemm(big[:, :K], w.to(DEV), None, residual=tab.to(DEV), res_mod=256)
    _close(out, big[:, :K].cpu().double() @ w.double().t() + tab.double().repeat(2, 1))


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n,H,Cin,Cout,k,stride", [(3, 20, 64, 64, 3, 1), (2, 16, 128, 512, 3, 1), (2, 32, 256, 256, 4, 2),
                                                    (1, 16, 512, 512, 3, 1), (5, 12, 64, 256, 3, 1), (2, 9, 4, 8, 3, 1)])
def test_conv2d_plain(n, H, Cin, Cout, k, stride):
    ops = _ops()
    x, w, b = _rand(n, Cin, H, H, seed=1), _rand(Cout, Cin, k, k, seed=2, scale=(Cin * k * k) ** -0.5), _rand(Cout, seed=3)
    pad = 1
    want = F.conv2d(F.relu(x).double(), w.double(), b.double(), stride=stride, padding=pad)
    out = ops.conv2d(_nhwc(x).to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), stride=stride, pad=(pad, pad),
                     relu_in=True)
    _close(out.permute(0, 3, 1, 2), want)


def test_conv2d_upsampled_input_and_residual_modes():
    ops = _ops()
    n, H, Cin, Cout = 2, 8, 64, 256
    x, w, b = _rand(n, Cin, H, H, seed=1), _rand(Cout, Cin, 3, 3, seed=2, scale=(Cin * 9) ** -0.5), _rand(Cout, seed=3)
    xu = F.interpolate(x, scale_factor=2, mode="nearest")
    res_lo = _rand(n, Cout, H, H, seed=4)
    res_hi = _rand(n, Cout, 2 * H, 2 * H, seed=5)
    res_map = _rand(1, Cout, 2 * H, 2 * H, seed=6)
    base = F.relu(F.conv2d(xu.double(), w.double(), b.double(), padding=1))
    wd, bd, xd = w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), _nhwc(x).to(DEV)
    out = ops.conv2d(xd, wd, bd, pad=(1, 1), in_up=True, act=1)
    _close(out.permute(0, 3, 1, 2), base)
    out = ops.conv2d(xd, wd, bd, pad=(1, 1), in_up=True, act=1, residual=_nhwc(res_hi).to(DEV), res_mode=1)
    _close(out.permute(0, 3, 1, 2), base + res_hi.double())
    out = ops.conv2d(xd, wd, bd, pad=(1, 1), in_up=True, act=1, residual=_nhwc(res_lo).to(DEV), res_mode=2)
    _close(out.permute(0, 3, 1, 2), base + F.interpolate(res_lo, scale_factor=2, mode="nearest").double())
    out = ops.conv2d(xd, wd, bd, pad=(1, 1), in_up=True, act=1, residual=_nhwc(res_map)[0].contiguous().to(DEV), res_mode=3)
    _close(out.permute(0, 3, 1, 2), base + res_map.double())


@pytest.mark.parametrize("Cout", [256, 1])
def test_conv_transpose_phases(Cout):
    """ConvTranspose2d(4,2,1) as four 2x2 sub-pixel convolutions scattered into the output."""
    ops = _ops()
    n, H, Cin = 2, 16, 256
    x = _rand(n, Cin, H, H, seed=1)
    wt = _rand(Cin, Cout, 4, 4, seed=2, scale=(Cin * 4) ** -0.5)
    b = _rand(Cout, seed=3)
    want = torch.tanh(F.conv_transpose2d(x.double(), wt.double(), b.double(), stride=2, padding=1))
    out = torch.zeros(n, 2 * H, 2 * H, Cout, device=DEV)
    taps = {0: (3, 1), 1: (2, 0)}
    for py in (0, 1):
        for px in (0, 1):
            sub = wt[:, :, list(taps[py]), :][:, :, :, list(taps[px])].permute(1, 2, 3, 0).contiguous()
            ops.conv2d(_nhwc(x).to(DEV), sub.to(DEV), b.to(DEV), pad=(1 - py, 1 - px), act=4, out=out, out_hw=(H, H),
                       scatter=(2, 2, py, px), full_hw=(2 * H, 2 * H))
    _close(out.permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("Cin,k,stride,pad,H", [(3, 7, 1, 3, 40), (1, 4, 2, 1, 64)])
def test_conv2d_first(Cin, k, stride, pad, H):
    ops = _ops()
    n, Cout = 2, 256
    x, w, b = _rand(n, Cin, H, H, seed=1), _rand(Cout, Cin, k, k, seed=2, scale=(Cin * k * k) ** -0.5), _rand(Cout, seed=3)
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad))
    wt = w.permute(1, 2, 3, 0).reshape(-1, Cout).contiguous()
    out = ops.conv2d_first(x.to(DEV), wt.to(DEV), b.to(DEV), cout=Cout, kh=k, kw=k, stride=stride, pad=pad, act=1)
    _close(out.permute(0, 3, 1, 2), want)


def test_conv1x1_tanh_planar_with_image_stride():
    ops = _ops()
    n, H, Cin, Cout, L = 3, 16, 256, 3, 4
    x, w, b = _rand(n, H, H, Cin, seed=1), _rand(Cout, Cin, seed=2, scale=Cin ** -0.5), _rand(Cout, seed=3)
    video = torch.zeros(n, L, Cout, H, H, device=DEV)
    ops.conv1x1_tanh_nchw(x.to(DEV), w.to(DEV), b.to(DEV), video[:, 2], L * Cout * H * H)
    want = torch.tanh(torch.einsum("nhwk,ck->nchw", F.relu(x).double(), w.double()) + b.double().view(1, -1, 1, 1))
    _close(video[:, 2], want)
    assert video[:, 0].abs().max().item() == 0 and video[:, 3].abs().max().item() == 0


def test_maxpool_layernorm_embedding_argmax_misc():
    ops = _ops()
    x = _rand(2, 12, 12, 64, seed=1)
    _close(ops.maxpool2x2(x.to(DEV)).permute(0, 3, 1, 2), F.max_pool2d(x.permute(0, 3, 1, 2), 2))
    for C, eps in ((512, 1e-5), (512, 1e-8), (1024, 1e-5), (128, 1e-5)):
        r, g, b = _rand(333, C, seed=2, scale=3.0) + 0.5, _rand(C, seed=3), _rand(C, seed=4)
        _close(ops.layernorm(r.to(DEV), g.to(DEV), b.to(DEV), eps), F.layer_norm(r.double(), (C,), g.double(), b.double(), eps), 1e-5, 1e-5)
    table = _rand(512, 1024, seed=5)
    idx = torch.randint(0, 512, (1000,), generator=torch.Generator().manual_seed(6))
    assert torch.equal(ops.embedding(idx.to(DEV), table.to(DEV)).cpu(), table[idx])
    lg = _rand(3000, 512, seed=7)
    lg[5, 100] = lg[5, 300] = 50.0  # exact tie -> lowest index, like torch.max
    assert torch.equal(ops.argmax_rows(lg.to(DEV)).cpu(), torch.max(lg, -1)[1])
    a = _rand(3, 16, 16, 512, seed=8)
    s, v = torch.rand(3), _rand(512, seed=9)
    ad = a.to(DEV).clone()
    ops.add_scaled_vec(ad, s.to(DEV), v.to(DEV))
    _close(ad, a.double() + (s.view(3, 1) @ v.view(1, -1)).double().view(3, 1, 1, 512), 1e-6, 1e-6)
    nz = _rand(3, 64, 16, 16, seed=10)
    assert torch.equal(ops.nchw_to_nhwc(nz.to(DEV)).cpu(), nz.permute(0, 2, 3, 1).contiguous())


def test_adain():
    ops = _ops()
    x, g, b = _rand(3, 512, 16, 16, seed=1, scale=2.0) + 1.0, _rand(3, 512, 16, 16, seed=2), _rand(3, 512, 16, 16, seed=3)
    want = g.double() * F.instance_norm(x.double(), eps=1e-5) + b.double()
    out = ops.adain(_nhwc(x).to(DEV), _nhwc(g).to(DEV), _nhwc(b).to(DEV), 1e-5)
    _close(out.permute(0, 3, 1, 2), want, 1e-5, 1e-5)


def test_text_embed_matches_reference_front_end():
    ops = _ops()
    B, T, C, V = 4, 14, 512, 50
    text = torch.randint(3, V, (B, T), generator=torch.Generator().manual_seed(1))
    text[1, 9:] = 0
    text[3, 12:] = 0
    tok, pos = _rand(V, C, seed=2, scale=0.02), _rand(38, C, seed=3, scale=0.02)
    tok[0] = 0
    g, b = 1 + 0.1 * _rand(C, seed=4), 0.05 * _rand(C, seed=5)
    x, klen = ops.text_embed(text.to(DEV), tok.to(DEV), pos.to(DEV), g.to(DEV), b.to(DEV), 0, 1e-8)
    want = F.layer_norm((tok[text] + pos[:T].unsqueeze(0)).double(), (C,), g.double(), b.double(), 1e-8)
    want = want * (text != 0).unsqueeze(-1)
    _close(x, want, 2e-5, 2e-5)
    assert klen.cpu().tolist() == (text != 0).sum(-1).tolist()


def _sdpa(q, k, v, mask=None):
    s = q.double() @ k.double().transpose(-1, -2) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return torch.softmax(s, -1) @ v.double()


def test_mha_text_self_attention_with_key_padding():
    ops = _ops()
    B, T, H, W = 3, 20, 16, 512
    qkv = _rand(B * T, 3 * W, seed=1)
    klen = torch.tensor([20, 13, 8], dtype=torch.int32)
    out = torch.empty(B * T, W, device=DEV)
    d = qkv.to(DEV)
    ops.mha(d, d[:, W:], d[:, 2 * W:], out, n_outer=B, n_inner=1, n_head=H, Sq=T, Sk=T, q_strides=(T * 3 * W, 0, 3 * W),
            k_strides=(T * 3 * W, 0, 3 * W), v_strides=(T * 3 * W, 0, 3 * W), o_strides=(T * W, 0, W), key_len=klen.to(DEV),
            scale=1 / math.sqrt(32))
    q, k, v = [t.view(B, T, H, 32).transpose(1, 2) for t in qkv.split(W, dim=1)]
    mask = (torch.arange(T).view(1, 1, 1, T) >= klen.view(B, 1, 1, 1))
    _close(out.view(B, T, H, 32).transpose(1, 2), _sdpa(q, k, v, mask), 1e-5, 1e-5)


@pytest.mark.parametrize("kind", [1, 2])
def test_mha_axial_spatial(kind):
    """H- and W-axial attention through strides over rows ordered (b, h, w)."""
    ops = _ops()
    B, R, C, H = 3, 16, 512, 16
    qkv = _rand(B * R * R, 3 * C, seed=kind)
    d = qkv.to(DEV)
    out = torch.empty(B * R * R, C, device=DEV)
    inner, seq = (1, R) if kind == 1 else (R, 1)
    st = (R * R * 3 * C, inner * 3 * C, seq * 3 * C)
    ops.mha(d, d[:, C:], d[:, 2 * C:], out, n_outer=B, n_inner=R, n_head=H, Sq=R, Sk=R, q_strides=st, k_strides=st, v_strides=st,
            o_strides=(R * R * C, inner * C, seq * C), key_len=None, scale=1 / math.sqrt(32))
    t = qkv.view(B, R, R, 3, H, 32)  # [b,h,w,3,head,32]
    ax = 1 if kind == 1 else 2
    q, k, v = [t[:, :, :, i].movedim(ax, -2) for i in range(3)]  # [b,other,head,S,32]
    want = _sdpa(q, k, v).movedim(-2, ax).reshape(B * R * R, C)
    _close(out, want, 1e-5, 1e-5)


@pytest.mark.parametrize("kind", [1, 2])
def test_axial_attention_tile_kernel(kind):
    """The specialised one-warp-per-(line, head) axial kernel: fp32 and split outputs, vs torch SDPA in fp64."""
    ops = _ops()
    B, R, C, H = 5, 16, 512, 16
    qkv = _rand(B * R * R, 3 * C, seed=10 + kind)
    d = qkv.to(DEV)
    out = torch.empty(B * R * R, C, device=DEV)
    sp = torch.empty(2, B * R * R, C, device=DEV, dtype=torch.float16)
    ops.axial_attn(d, out, B=B, R=R, n_head=H, axis=kind, scale=1 / math.sqrt(32), out_split=sp)
    t = qkv.view(B, R, R, 3, H, 32)
    ax = 1 if kind == 1 else 2
    q, k, v = [t[:, :, :, i].movedim(ax, -2) for i in range(3)]
    want = _sdpa(q, k, v).movedim(-2, ax).reshape(B * R * R, C)
    _close(out, want, 1e-5, 1e-5)
    rec = sp.cpu().double()
    _close(rec[0] + rec[1] / 2048.0, want, 1e-5, 1e-5)
    only_split = torch.empty_like(sp)
    ops.axial_attn(d, None, B=B, R=R, n_head=H, axis=kind, scale=1 / math.sqrt(32), out_split=only_split)
    assert torch.equal(only_split, sp)


@pytest.mark.parametrize("impl", ["generic", "tma"])
def test_temporal_attention_with_kv_cache(impl):
    ops = _ops()
    M, L, C, H = 700, 12, 512, 16
    steps = [_rand(M, 3 * C, seed=10 + p) for p in range(L)]
    kc = torch.zeros(M, L, C, device=DEV)
    vc = torch.zeros(M, L, C, device=DEV)
    for p in range(L):
        d = steps[p].to(DEV)
        out = torch.empty(M, C, device=DEV)
        if impl == "tma":
            ops.temporal_attn_step(d, kc, vc, out, p, 1 / math.sqrt(32))
        else:
            ops.kv_append(d, kc, vc, p)
            ops.mha(d, kc, vc, out, n_outer=M, n_inner=1, n_head=H, Sq=1, Sk=p + 1, q_strides=(3 * C, 0, 0),
                    k_strides=(L * C, 0, C), v_strides=(L * C, 0, C), o_strides=(C, 0, 0), key_len=None, scale=1 / math.sqrt(32))
        q = steps[p][:, :C].view(M, H, 1, 32)
        K = torch.stack([s[:, C:2 * C] for s in steps[:p + 1]], 1).view(M, p + 1, H, 32).transpose(1, 2)
        V = torch.stack([s[:, 2 * C:] for s in steps[:p + 1]], 1).view(M, p + 1, H, 32).transpose(1, 2)
        _close(out.view(M, H, 1, 32), _sdpa(q, K, V), 1e-5, 1e-5)
    if impl == "tma":   # cache layout of the TMA-staged kernel: [M][2 head-halves][L][256]
        k4, v4 = kc.view(M, 2, L, C // 2).cpu(), vc.view(M, 2, L, C // 2).cpu()
        assert torch.equal(torch.cat([k4[:, 0, 3], k4[:, 1, 3]], 1), steps[3][:, C:2 * C])
        assert torch.equal(torch.cat([v4[:, 0, L - 1], v4[:, 1, L - 1]], 1), steps[L - 1][:, 2 * C:])
    else:
        assert torch.equal(kc[:, 3].cpu(), steps[3][:, C:2 * C])
        assert torch.equal(vc[:, L - 1].cpu(), steps[L - 1][:, 2 * C:])


def test_mha_cross_attention_long_query():
    ops = _ops()
    B, HW, T, C, H = 2, 256, 20, 512, 16
    q, kv = _rand(B * HW, C, seed=1), _rand(B * T, 2 * C, seed=2)
    out = torch.empty(B * HW, C, device=DEV)
    qd, kvd = q.to(DEV), kv.to(DEV)
    ops.mha(qd, kvd, kvd[:, C:], out, n_outer=B, n_inner=1, n_head=H, Sq=HW, Sk=T, q_strides=(HW * C, 0, C),
            k_strides=(T * 2 * C, 0, 2 * C), v_strides=(T * 2 * C, 0, 2 * C), o_strides=(HW * C, 0, C), key_len=None,
            scale=1 / math.sqrt(32))
    Q = q.view(B, HW, H, 32).transpose(1, 2)
    K = kv[:, :C].reshape(B, T, H, 32).transpose(1, 2)
    V = kv[:, C:].reshape(B, T, H, 32).transpose(1, 2)
    _close(out.view(B, HW, H, 32).transpose(1, 2), _sdpa(Q, K, V), 1e-5, 1e-5)


@pytest.mark.parametrize("D", [256, 1024])
def test_vq_argmin_matches_reference_formula(D):
    from oracle import mage_oracle as orc
    ops = _ops()
    N, K = 3000, 512
    cb = _rand(K, D, seed=1)
    z = cb[torch.randint(0, K, (N,), generator=torch.Generator().manual_seed(2))] + 0.3 * _rand(N, D, seed=3)
    dist = orc.vq_distances(z, cb)
    want = torch.min(dist, 1)[1]
    got = ops.vq_argmin(z.to(DEV), cb.to(DEV)).cpu()
    top2 = torch.topk(dist, 2, dim=1, largest=False)[0]
    gap = top2[:, 1] - top2[:, 0]
    bad = (got != want) & (gap > 1e-3)
    assert not bad.any(), f"{int(bad.sum())} wrong codes away from ties"
    assert (got != want).sum().item() <= 2
    # exact tie between two identical codes -> lowest index (torch.min semantics)
    cb2 = cb.clone()
    cb2[400] = cb2[17]
    z2 = cb2[17:18].repeat(64, 1)
    assert ops.vq_argmin(z2.to(DEV), cb2.to(DEV)).cpu().tolist() == [17] * 64


def test_bad_arguments_are_rejected_not_run():
    from mage_b200._lib import MageCudaError
    ops = _ops()
    with pytest.raises(MageCudaError):
        ops.gemm(torch.zeros(4, 6, device=DEV), torch.zeros(8, 6, device=DEV))  # K % 4 != 0
    with pytest.raises(MageCudaError):
        ops.layernorm(torch.zeros(4, 100, device=DEV), torch.zeros(100, device=DEV), torch.zeros(100, device=DEV))
    with pytest.raises(AssertionError):
        ops.gemm(torch.zeros(4, 8), torch.zeros(8, 8))  # CPU tensors: no fallback


def test_token_taps_equals_conv_plus_linear():
    """in_linear(conv3x3(E[tok]) + pos) (mage_model.py:674-676,375) as nine table lookups per pixel."""
    ops = _ops()
    import torch.nn.functional as F
    B, R, K, C = 3, 16, 512, 512
    g = torch.Generator().manual_seed(5)
    E = torch.randn(K, C, generator=g) * 0.02
    Wc = torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5
    Win = torch.randn(C, C, generator=g) * C ** -0.5
    b_in = torch.randn(C, generator=g) * 0.1
    pos = torch.randn(R * R, C, generator=g) * C ** -0.5
    tok = torch.randint(0, K, (B, R, R), generator=g)
    emb = E[tok].permute(0, 3, 1, 2).double()                                  # [B,C,R,R]
    f = F.conv2d(emb, Wc.double(), padding=1).permute(0, 2, 3, 1) + pos.double().view(1, R, R, C)
    want = f @ Win.double().t() + b_in.double()
    comp = torch.einsum("oc,cikl->klio", Win.double(), Wc.double())
    table = torch.einsum("ei,klio->kleo", E.double(), comp).reshape(9, K, C).float().contiguous()
    posW = (pos.double() @ Win.double().t()).float()
    out = torch.empty(B * R * R, C, device=DEV)
    ops.token_taps(tok.to(DEV), table.to(DEV), posW.to(DEV), b_in.to(DEV), out)
    _close(out.view(B, R, R, C), want, 2e-6, 2e-6)
    # the variant that also applies the first block's ln_1 to each finished row: same fp32 rows, and the split LayerNorm output is
    # bit-identical to mage_layernorm_f32 run on them
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    out2 = torch.empty_like(out)
    u2 = torch.full((2, B * R * R, C), float("nan"), device=DEV, dtype=torch.float16)
    ops.token_taps_ln(tok.to(DEV), table.to(DEV), posW.to(DEV), b_in.to(DEV), out2, gamma, beta, u2)
    u = torch.empty_like(u2)
    ops.layernorm(out, gamma, beta, out_split=u)
    assert torch.equal(out2, out) and torch.equal(u2.view(torch.int16), u.view(torch.int16))
    ops.check_flag(DEV)


def test_objective_kernels_group_norm_cross_entropy_reparam_kl_sum():
    """The bandwidth-bound pieces of the stage-2 objective's forward half (mage_b200.h, last section) against torch fp64."""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    # GroupNorm(16) over (32 channels x T frames x HW) per sample, frame-major rows, + residual + ReLU
    T, B, HW, C = 3, 2, 64, 512
    x = torch.randn(T, B, HW, C, generator=g) * 2 + 0.3
    res = torch.randn(T, B, HW, C, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xd = x.to(DEV).view(-1, C)
    part = torch.empty(T, B, 16, 2, device=DEV, dtype=torch.float64)
    ops.gn_partial(xd, part, B, HW, groups=16)
    out = torch.empty_like(xd)
    sp = torch.empty(2, T * B * HW, C, device=DEV, dtype=torch.float16)
    ops.gn_apply(xd, part, gamma.to(DEV), beta.to(DEV), B, HW, relu=True, residual=res.to(DEV).view(-1, C), out=out, out_split=sp)
    ncthw = x.permute(1, 3, 0, 2).double()                                     # [B, C, T, HW]
    want = F.relu(F.group_norm(ncthw, 16, gamma.double(), beta.double(), 1e-5) + res.permute(1, 3, 0, 2).double())
    _close(out.view(T, B, HW, C).permute(1, 3, 0, 2), want, 1e-5, 1e-5)
    _close(sp[0].float() + sp[1].float() / 2048, out.cpu(), 1e-6, 1e-6)
    ops.gn_apply(xd, part, gamma.to(DEV), beta.to(DEV), B, HW, relu=False, out=out)
    _close(out.view(T, B, HW, C).permute(1, 3, 0, 2), F.group_norm(ncthw, 16, gamma.double(), beta.double(), 1e-5), 1e-5, 1e-5)
    # cross-entropy rows (strided logits) and the fixed-order mean
    rows, K = 1000, 512
    lg = torch.randn(rows, K + 8, generator=g) * 3
    tgt = torch.randint(0, K, (rows,), generator=g)
    loss = torch.empty(rows, device=DEV)
    ops.cross_entropy_rows(lg.to(DEV)[:, :K], tgt.to(DEV), loss)
    want = F.cross_entropy(lg[:, :K].double(), tgt, reduction="none")
    _close(loss, want, 2e-6, 2e-6)
    mean = ops.scaled_sum(loss, 1.0 / rows)
    assert abs(mean.item() - want.mean().item()) <= 2e-6 * want.mean().item()
    ops.check_flag(DEV)
    bad = tgt.clone()
    bad[17] = K
    ops.cross_entropy_rows(lg.to(DEV)[:, :K], bad.to(DEV), loss)
    with pytest.raises(IndexError):
        ops.check_flag(DEV)
    # reparameterisation + KL integrand
    B, HW, Cz = 3, 256, 64
    ml = torch.randn(B * HW, 2 * Cz, generator=g)
    eps = torch.randn(B, Cz, 16, 16, generator=g)
    z, kl_rows = ops.reparam_kl(ml.to(DEV), eps.to(DEV), B, HW)
    mu = ml[:, :Cz].view(B, HW, Cz).permute(0, 2, 1).double()
    lv = ml[:, Cz:].view(B, HW, Cz).permute(0, 2, 1).double()
    _close(z, eps.view(B, Cz, HW).double() * torch.exp(0.5 * lv) + mu, 2e-6, 2e-6)
    _close(kl_rows, (1 + lv - mu ** 2 - lv.exp()).sum((1, 2)), 2e-6, 1e-3)


def test_conv3d_as_one_implicit_gemm_over_frame_triples():
    """BasicBlock's Conv3d 3x3x3 (mage_model.py:267,270,273: padding 1, temporal stride 2 or 1, no bias) = ONE tensor-core
    convolution over 3*Cin channels: the three temporal taps sit side by side on the channel axis of the gathered frame triples
    (engine._frame_triples), the kernel is reordered to [Cout, ky, kx, kt*Cin + ci]."""
    from mage_b200.engine import SamplerEngine
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    B, R, Cin, Cout = 2, 16, 128, 128
    for T, stride_t in ((5, 2), (4, 2), (1, 2), (3, 1)):
        x = torch.randn(T, B, R, R, Cin, generator=g)
        w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (27 * Cin) ** -0.5
        want = F.conv3d(x.permute(1, 4, 0, 2, 3).double(), w.double(), None, stride=(stride_t, 1, 1), padding=1)   # [B,Cout,T',R,R]
        xt = SamplerEngine._frame_triples(x.to(DEV), stride_t)
        To = xt.shape[0]
        assert To == want.shape[2] and xt.shape[-1] == 3 * Cin
        wf = w.permute(0, 3, 4, 2, 1).reshape(Cout, 3, 3, 3 * Cin).contiguous()
        y, _, _ = ops.conv2d_tc(ops.split(xt.view(To * B, R, R, 3 * Cin)), ops.split(wf.to(DEV)), None, pad=(1, 1))
        _close(y.view(To, B, R, R, Cout).permute(1, 4, 0, 2, 3), want, 1e-5, 1e-5)   # fp32-grade for K = 3456 (one fp16 pass: ~5e-4)
    ops.check_flag(DEV)


@pytest.mark.parametrize("M,pos,Lmax", [(3000, 0, 32), (3000, 1, 32), (2500, 5, 32), (700, 15, 32), (700, 31, 32), (64, 3, 8), (300, 23, 24),
                                        (37, 0, 16), (40, 60, 64)])
def test_temporal_attn_ring_kernel_equals_one_cta_per_unit(M, pos, Lmax):
    """The temporal attention step as a persistent kernel with a ring of staging slots (optional) against the one-CTA-per-unit form:
    same arithmetic per (location, head half), so attention output (fp32 and split) and the appended K/V rows are bit-identical --
    with more units than resident CTAs x slots (the ring wraps), fewer units than CTAs, every slot count, and the clip length at
    which two slots no longer fit (one-shot form takes over)."""
    ops = _ops()
    C = 512
    g = torch.Generator().manual_seed(21)
    qkv = torch.randn(M, 3 * C, generator=g).to(DEV)
    kc0 = torch.randn(M, Lmax, C, generator=g).to(DEV)
    vc0 = torch.randn(M, Lmax, C, generator=g).to(DEV)
    res = []
    try:
        for ring in (True, False):
            ops.temporal_attn_ring(ring)
            kc, vc = kc0.clone(), vc0.clone()
            out = torch.full((M, C), float("nan"), device=DEV)
            sp = torch.zeros(2, M, C, device=DEV, dtype=torch.float16)
            ops.temporal_attn_step(qkv, kc, vc, out, pos, 32 ** -0.5, out_split=sp)
            torch.cuda.synchronize()
            res.append((out, sp, kc, vc))
    finally:
        ops.temporal_attn_ring(False)   # the default
    for a, b in zip(*res):
        assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a.view(torch.int16),
                           b.view(torch.int32) if b.dtype == torch.float32 else b.view(torch.int16))
    assert torch.isfinite(res[0][0]).all()
    ops.check_flag(DEV)


def test_handles_are_independent():
    """all library state lives in the opaque handle.  A second handle on the same device has its own launch
    counter and tuning switches; using it does not disturb the handle the package works through."""
    import ctypes

    from mage_b200 import _lib, ops
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.mage_ctx_create(0, ctypes.byref(h)) == 0 and h.value
    assert L.mage_ctx_device(h) == 0 and L.mage_launch_count(h) == 0
    n0 = ops.launch_count()
    x = torch.randn(64, 512, device="cuda")
    g, b = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda")
    want = ops.layernorm(x, g, b)
    assert ops.launch_count() == n0 + 1 and L.mage_launch_count(h) == 0
    assert L.mage_tc_tuning(h, 64, 0) == 0 and L.mage_pdl(h, 1) == 0     # switches of the second handle only
    out = torch.empty_like(x)
    assert L.mage_layernorm_f32(h, x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), None, 0, None, 64, 512, 1e-5,
                                torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert torch.equal(out, want) and L.mage_launch_count(h) == 1 and ops.launch_count() == n0 + 1
    assert L.mage_layernorm_f32(None, x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), None, 0, None, 64, 512, 1e-5, None) == -1
    assert L.mage_ctx_destroy(h) == 0
    h2 = ctypes.c_void_p()
    assert L.mage_ctx_create(99, ctypes.byref(h2)) != 0 and not h2.value   # no such device


@pytest.mark.parametrize("M,pos0,n_pos,Lmax", [(64, 0, 5, 8), (33, 3, 4, 10), (16, 9, 23, 32), (8, 31, 1, 32), (8, 0, 16, 16), (8, 0, 24, 24),
                                               (8, 20, 4, 24)])  # 16 / 24: dynamic shared memory of exactly 48 KB next to the static part
def test_temporal_attn_seq_equals_position_by_position(M, pos0, n_pos, Lmax):
    """mage_temporal_attn_seq_f32 (n_pos consecutive positions in one launch, the K/V prefix staged once) against n_pos calls of
    mage_temporal_attn_step_f32: same math order -> bit-identical attention output and cache contents."""
    from mage_b200 import ops
    C = 512
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(n_pos * M, 3 * C, generator=g).cuda()
    pre = torch.randn(2, M, Lmax, C, generator=g).cuda()
    k1, v1 = pre[0].clone(), pre[1].clone()
    k2, v2 = pre[0].clone(), pre[1].clone()
    scale = 32 ** -0.5
    u1 = torch.zeros(2, n_pos * M, C, device="cuda", dtype=torch.float16)
    u2 = torch.zeros_like(u1)
    for s in range(n_pos):
        ops.temporal_attn_step(qkv[s * M:(s + 1) * M], k1, v1, None, pos0 + s, scale, out_split=u1[:, s * M:(s + 1) * M])
    ops.temporal_attn_seq(qkv, k2, v2, pos0, n_pos, scale, out_split=u2)
    torch.cuda.synchronize()
    assert torch.equal(u1, u2)
    assert torch.equal(k1, k2) and torch.equal(v1, v2)
