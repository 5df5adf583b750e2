"""GPU parity tests of every C-ABI kernel against a CPU fp32 restatement of the same op (oracle functions where the
op exists in the reference, plain torch otherwise).  Integer / index outputs are compared bit-exactly."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import styler_oracle as so
from oracle import stft_oracle

pytestmark = pytest.mark.gpu

ACTS = {0: lambda v: v, 1: torch.relu, 2: torch.tanh}


def _ops():
    from styler_b200 import ops
    return ops


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def conv_ref(x, w, bias, pad, act=0, residual=None, ln=None, act2=0, lens=None, dot=None):
    """fp32 CPU statement of the styler_conv1d_fwd contract; x [B,T,Cin], w [KS,N,Cin]."""
    y = F.conv1d(x.transpose(1, 2), w.permute(1, 2, 0).contiguous(), bias, padding=pad).transpose(1, 2)
    y = ACTS[act](y)
    if residual is not None:
        y = y + residual
    if ln is not None:
        y = F.layer_norm(y, (y.shape[-1],), ln[0], ln[1], 1e-5)
    y = ACTS[act2](y)
    d = None
    if dot is not None:
        d = y @ dot[0] + dot[1]
    if lens is not None:
        m = so.mask_from_lengths(lens, y.shape[1])
        y = y.masked_fill(m.unsqueeze(-1), 0)
        if d is not None:
            d = d.masked_fill(m, 0)
    return y, d


CONV_CASES = [
    # name,               B, T,   Cin,  N,    KS, kw
    ("linear_bias",       2, 200, 256,  256,  1, dict()),
    ("ffn1_k9_relu",      2, 300, 256,  1024, 9, dict(act=1)),
    ("ffn2_res_ln_mask",  3, 130, 1024, 256,  1, dict(res=True, ln=True, lens=True)),
    ("pred_k3_relu_ln_dot", 2, 257, 256, 256, 3, dict(act=1, ln=True, lens=True, dot=True)),
    ("postnet_in_k5_tanh", 2, 140, 80,  512,  5, dict(act=2)),
    ("postnet_out_k5_res_f32", 2, 140, 512, 80, 5, dict(res=True, f32=True)),
    ("qkv_vt_split",      2, 150, 256,  768,  1, dict(vt=True)),
    ("lstm_proj_k160",    2, 128, 160,  640,  1, dict(f32only=True)),
    ("cls_ln_relu",       2, 64,  128,  256,  1, dict(ln=True, act2=1)),
    ("audio_k5_320",      2, 200, 320,  320,  5, dict()),
    ("row_broadcast_res", 2, 96,  128,  256,  1, dict(act=1, res_row=True)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_conv1d(cuda, case, dtype, impl):
    ops = _ops()
    name, B, T, Cin, N, KS, kw = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(B, T, Cin, generator=g)
    w = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    bias = torch.randn(N, generator=g) * 0.1
    xq, wq = x.to(dtype).float(), w.to(dtype).float()          # operands exactly as the kernel sees them
    res = torch.randn(B, T, N, generator=g) if kw.get("res") else None
    res_row = torch.randn(B, N, generator=g) if kw.get("res_row") else None
    ln = (1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)) if kw.get("ln") else None
    lens = torch.tensor([T, max(1, T // 2), max(1, T - 7)][:B], dtype=torch.int64) if kw.get("lens") else None
    dot = ((torch.rand(N, generator=g) - 0.5) * 0.2, 0.3) if kw.get("dot") else None
    resq = res.to(dtype).float() if res is not None else (res_row.to(dtype).float().unsqueeze(1) if res_row is not None else None)
    pad = (KS - 1) // 2
    ref, ref_dot = conv_ref(xq, wq, bias, pad, kw.get("act", 0), resq, ln, kw.get("act2", 0), lens, dot)

    dev = cuda
    xd, wd = x.to(dev, dtype), w.to(dev, dtype)
    args = dict(pad=pad, act=kw.get("act", 0), act2=kw.get("act2", 0),
                impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_TC)
    if res is not None:
        args["residual"] = res.to(dev, dtype)
    if res_row is not None:
        args["residual_row"] = res_row.to(dev, dtype)
    if ln is not None:
        args["ln"] = (ln[0].to(dev), ln[1].to(dev))
    if lens is not None:
        args["lens"] = lens.to(dev)
    if dot is not None:
        args["dot"] = (dot[0].to(dev), dot[1])
    out_f32 = torch.zeros(B, T, N, device=dev) if (kw.get("f32") or kw.get("f32only")) else None
    if out_f32 is not None:
        args["out_f32"] = out_f32
    vt = None
    if kw.get("vt"):
        Tp = (T + 63) // 64 * 64
        vt = torch.zeros(B, 256, Tp, device=dev, dtype=dtype)
        args.update(vt=vt, vt_col0=512)
    if kw.get("f32only"):
        args["want_out"] = False
    got = ops.conv1d(xd, wd, bias.to(dev), **args)
    torch.cuda.synchronize()
    got_dot = None
    if dot is not None:
        got, got_dot = got
    # tolerance: operands are identical, so only accumulation order (and tf32 operand rounding) differs
    tol_acc = 2e-3 if (dtype == torch.float32 and impl == "tc") else 2e-5
    tol_out = tol_acc + {torch.bfloat16: 8e-3, torch.float16: 1e-3}.get(dtype, 0.0)    # 16-bit storage of the result
    if kw.get("ln") and dtype != torch.float32 and impl == "simt":
        tol_out += 2e-2 if dtype == torch.bfloat16 else 2.5e-3         # simt stages pre-LN rows in the storage type
    if vt is not None:
        assert rel_err(got, ref[..., :512]) < tol_out
        assert rel_err(vt[:, :, :T].transpose(1, 2), ref[..., 512:]) < tol_out
    elif not kw.get("f32only"):
        assert rel_err(got, ref) < tol_out, name
    if out_f32 is not None:
        assert rel_err(out_f32, ref) < (tol_out if (kw.get("ln") and impl == "simt") else tol_acc)
    if got_dot is not None:
        assert rel_err(got_dot, ref_dot) < tol_out + 1e-3


PERSIST_CASES = [
    # more output tiles than CTA slots on a 148-SM part -> the persistent tile loop (two TMEM accumulators, operand ring and
    # residual prefetch running across tile boundaries); ragged T so the last tile of every utterance is partial
    # name,                B,  T,    Cin,  N,   KS, kw
    ("qkv_like",           10, 2100, 256,  768, 1, dict()),
    ("outproj_res_ln",     10, 2100, 256,  256, 1, dict(res=True, ln=True, lens=True)),
    ("ffn2_res_ln",        10, 2100, 1024, 256, 1, dict(res=True, ln=True, lens=True)),
    ("pred_k3_relu_ln",    10, 2100, 256,  256, 3, dict(act=1, ln=True)),
    ("mel_linear_n64_res", 40, 1000, 256,  64,  1, dict(res=True)),
    ("postnet_in_tanh",    10, 2100, 80,   512, 5, dict(act=2)),
    # 320-channel audio-encoder conv: full-width N tile (one A stage, two MMAs of 160 columns), >= 148 M tiles
    ("audio_k5_320_wide",  10, 2100, 320,  320, 5, dict()),
    ("audio_k5_320_wide_relu_res", 9, 2300, 320, 320, 5, dict(act=1, res=True)),
]


@pytest.mark.parametrize("case", PERSIST_CASES, ids=[c[0] for c in PERSIST_CASES])
def test_conv1d_tc_persistent(cuda, case):
    ops = _ops()
    name, B, T, Cin, N, KS, kw = case
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(B, T, Cin, generator=g)
    w = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    bias = torch.randn(N, generator=g) * 0.1
    xq, wq = x.to(dtype).float(), w.to(dtype).float()
    res = torch.randn(B, T, N, generator=g) if kw.get("res") else None
    ln = (1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)) if kw.get("ln") else None
    lens = (torch.randint(1, T + 1, (B,), generator=g).to(torch.int64)) if kw.get("lens") else None
    resq = res.to(dtype).float() if res is not None else None
    pad = (KS - 1) // 2
    ref, _ = conv_ref(xq, wq, bias, pad, kw.get("act", 0), resq, ln, 0, lens, None)
    args = dict(pad=pad, act=kw.get("act", 0), impl=ops.IMPL_TC)
    if res is not None:
        args["residual"] = res.to(cuda, dtype)
    if ln is not None:
        args["ln"] = (ln[0].to(cuda), ln[1].to(cuda))
    if lens is not None:
        args["lens"] = lens.to(cuda)
    got = ops.conv1d(x.to(cuda, dtype), w.to(cuda, dtype), bias.to(cuda), **args)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    # per-utterance check so that one bad tile cannot hide behind the global max
    for b in range(B):
        assert rel_err(got[b], ref[b]) < 1e-2, (name, b)
    # a second launch into the same buffers gives bitwise the same result (no cross-tile race)
    got2 = ops.conv1d(x.to(cuda, dtype), w.to(cuda, dtype), bias.to(cuda), **args)
    torch.cuda.synchronize()
    assert torch.equal(got, got2)


PAIR_CASES = [
    # CTA pairs (tcgen05 cta_group::2, M = 256 over two SMs): 256-wide N tiles, no residual.  Ragged T (partial last tile per
    # utterance), odd tiles-per-utterance so a pair straddles two utterances, dilation, every activation.
    # name,             B,  T,    Cin, N,    KS, kw
    ("ffn1_k9_relu",    4,  1024, 256, 1024, 9, dict(act=1)),
    ("postnet_k5_tanh", 6,  300,  512, 512,  5, dict(act=2)),
    ("linear_256",      2,  700,  256, 256,  1, dict()),
    ("k3_dil3_n512",    2,  1000, 128, 512,  3, dict(dil=3)),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_conv1d_tc_cta_pairs(cuda, case):
    from styler_b200 import _lib
    ops = _ops()
    name, B, T, Cin, N, KS, kw = case
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(B, T, Cin, generator=g)
    w = (torch.rand(KS, N, Cin, generator=g) * 2 - 1) / math.sqrt(Cin * KS)
    bias = torch.randn(N, generator=g) * 0.1
    dil = kw.get("dil", 1)
    pad = dil * (KS - 1) // 2
    y = F.conv1d(x.to(dtype).float().transpose(1, 2), w.to(dtype).float().permute(1, 2, 0).contiguous(), bias, padding=pad,
                 dilation=dil).transpose(1, 2)
    ref = ACTS[kw.get("act", 0)](y)
    xd, wd, bd = x.to(cuda, dtype), w.to(cuda, dtype), bias.to(cuda)
    args = dict(pad=pad, act=kw.get("act", 0), dilation=dil, impl=ops.IMPL_TC)
    try:
        _lib.set_tuning("TC_PERSIST", 0)
        _lib.set_tuning("TC_BN", 256)                 # small problems would otherwise pick a narrower N tile
        _lib.set_tuning("TC_2CTA", 0)
        single = ops.conv1d(xd, wd, bd, **args)
        _lib.set_tuning("TC_2CTA", 2)                 # wherever legal
        n0 = _lib.launch_count()
        pair = ops.conv1d(xd, wd, bd, **args)
        pair2 = ops.conv1d(xd, wd, bd, **args)
        torch.cuda.synchronize()
        assert _lib.launch_count() == n0 + 2
    finally:
        _lib.set_tuning("TC_2CTA", -1)
        _lib.set_tuning("TC_BN", -1)
        _lib.set_tuning("TC_PERSIST", -1)
    assert torch.isfinite(pair.float()).all()
    for b in range(B):
        assert rel_err(pair[b], ref[b]) < 1e-2, (name, b)
    assert torch.equal(pair, pair2), "two launches must agree bitwise"
    assert torch.equal(pair, single), "the pair form accumulates in the same order as the single-CTA form"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("T", [128, 200, 520])
@pytest.mark.parametrize("v_layout", ["transposed", "rowmajor"])
def test_attention(cuda, dtype, impl, T, v_layout):
    ops = _ops()
    B, H = 3, 4
    g = torch.Generator().manual_seed(T)
    qk = torch.randn(B, T, 512, generator=g)
    v = torch.randn(B, T, 256, generator=g)
    lens = torch.tensor([T, max(1, T // 3), max(1, T - 5)], dtype=torch.int64)
    qkq, vq = qk.to(dtype).float(), v.to(dtype).float()
    q = qkq[..., :256].view(B, T, H, 64).permute(0, 2, 1, 3)
    k = qkq[..., 256:].view(B, T, H, 64).permute(0, 2, 1, 3)
    vv = vq.view(B, T, H, 64).permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2)
    s = s.masked_fill(so.mask_from_lengths(lens, T)[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vv).permute(0, 2, 1, 3).reshape(B, T, 256)
    Tp = (T + 63) // 64 * 64
    vt = torch.zeros(B, 256, Tp, dtype=dtype)
    vt[:, :, :T] = v.to(dtype).transpose(1, 2)
    if v_layout == "rowmajor" and impl == "tc" and dtype == torch.float32:
        pytest.skip("row-major V on the tensor-core path is bf16 only (tf32 keeps the transposed-V layout)")
    if v_layout == "rowmajor":
        got = ops.attention(torch.cat([qk, v], dim=-1).to(cuda, dtype), None, lens.to(cuda), H,
                            impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_TC)
    else:
        got = ops.attention(qk.to(cuda, dtype), vt.to(cuda), lens.to(cuda), H,
                            impl=ops.IMPL_SIMT if impl == "simt" else ops.IMPL_TC)
    torch.cuda.synchronize()
    tol = {("simt", torch.float32): 2e-5, ("simt", torch.bfloat16): 8e-3, ("tc", torch.float32): 3e-3,
           ("tc", torch.bfloat16): 1.5e-2, ("simt", torch.float16): 1e-3, ("tc", torch.float16): 2e-3}[(impl, dtype)]
    assert rel_err(got, ref) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
def test_attention_tc_edge_lengths_and_rising_max(cuda, dtype):
    """Key-padding edge cases of the tensor-core kernel (a key half of a tile with no valid key, one valid key,
    exact tile boundaries) and scores whose row maximum keeps growing from key tile to key tile, so the lazily
    raised running max / in-TMEM rescale of the accumulator is exercised (Modules.py:16-23 semantics)."""
    ops = _ops()
    H, T = 4, 400
    lens = torch.tensor([1, 40, 64, 65, 128, 129, 200, 400], dtype=torch.int64)
    B = lens.numel()
    g = torch.Generator().manual_seed(11)
    qk = torch.randn(B, T, 512, generator=g)
    qk[..., 256:] *= torch.linspace(0.1, 1.6, T).view(1, T, 1)          # keys grow with position -> maxima keep rising
    v = torch.randn(B, T, 256, generator=g)
    qkq, vq = qk.to(dtype).float(), v.to(dtype).float()
    q = qkq[..., :256].view(B, T, H, 64).permute(0, 2, 1, 3)
    k = qkq[..., 256:].view(B, T, H, 64).permute(0, 2, 1, 3)
    vv = vq.view(B, T, H, 64).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)).masked_fill(so.mask_from_lengths(lens, T)[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vv).permute(0, 2, 1, 3).reshape(B, T, 256)
    if dtype != torch.float32:
        got = ops.attention(torch.cat([qk, v], dim=-1).to(cuda, dtype), None, lens.to(cuda), H, impl=ops.IMPL_TC)
    else:
        Tp = (T + 63) // 64 * 64
        vt = torch.zeros(B, 256, Tp, dtype=dtype)
        vt[:, :, :T] = v.transpose(1, 2)
        got = ops.attention(qk.to(cuda, dtype), vt.to(cuda), lens.to(cuda), H, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    # f32 runs as tf32 on the tensor pipe: operand rounding (2^-11) times scores of magnitude ~40 -> percent-level p error
    tol = {torch.bfloat16: 1.5e-2, torch.float16: 3e-3}.get(dtype, 2e-2)
    for b in range(B):
        assert rel_err(got[b], ref[b]) < tol, (b, int(lens[b]))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_attention_tc_persistent_items(cuda, dtype):
    """More (utterance, head, query-tile) work items than resident CTA slots: every CTA of the persistent attention kernel walks
    several items, so the K/V ring, the S / P / PV hand-shakes, the Q refill and the O accumulators run across item
    boundaries.  Ragged lengths give items with 1..8 key tiles next to each other (and a different tile count on either
    side of most boundaries).  Checked per utterance against the CPU statement, and bitwise against the one-item-per-CTA
    form of the same kernel (ATTN_PERSIST=0)."""
    from styler_b200 import _lib
    ops = _ops()
    H, T = 4, 1024
    B = 12 if dtype == torch.bfloat16 else 6            # 384 / 192 items > 296 / 148 slots
    g = torch.Generator().manual_seed(23)
    lens = torch.randint(1, T + 1, (B,), generator=g).to(torch.int64)
    lens[0], lens[1], lens[2] = T, 1, 129
    qk = torch.randn(B, T, 512, generator=g) * 0.7
    v = torch.randn(B, T, 256, generator=g)
    qkq, vq = qk.to(dtype).float(), v.to(dtype).float()
    q = qkq[..., :256].view(B, T, H, 64).permute(0, 2, 1, 3)
    k = qkq[..., 256:].view(B, T, H, 64).permute(0, 2, 1, 3)
    vv = vq.view(B, T, H, 64).permute(0, 2, 1, 3)
    sc = (q @ k.transpose(-1, -2)).masked_fill(so.mask_from_lengths(lens, T)[:, None, None, :], float("-inf"))
    ref = (torch.softmax(sc, -1) @ vv).permute(0, 2, 1, 3).reshape(B, T, 256)

    def run():
        if dtype == torch.bfloat16:
            return ops.attention(torch.cat([qk, v], dim=-1).to(cuda, dtype), None, lens.to(cuda), H, impl=ops.IMPL_TC)
        vt = torch.zeros(B, 256, T, dtype=dtype)
        vt[:, :, :T] = v.transpose(1, 2)
        return ops.attention(qk.to(cuda, dtype), vt.to(cuda), lens.to(cuda), H, impl=ops.IMPL_TC)

    try:
        _lib.set_tuning("ATTN_PERSIST", 1)
        got = run()
        got2 = run()
        _lib.set_tuning("ATTN_PERSIST", 0)
        one = run()
        torch.cuda.synchronize()
    finally:
        _lib.set_tuning("ATTN_PERSIST", -1)
    assert torch.isfinite(got.float()).all()
    tol = 1.5e-2 if dtype == torch.bfloat16 else 5e-3
    for b in range(B):
        assert rel_err(got[b], ref[b]) < tol, (b, int(lens[b]))
    assert torch.equal(got, got2) and torch.equal(got, one)


def test_embed_add_cast(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    seq = torch.randint(0, 152, (3, 37), generator=g)
    emb, pos = torch.randn(152, 256, generator=g), torch.randn(37, 256, generator=g)
    got = ops.embed_pos(seq.to(cuda), emb.to(cuda), pos.to(cuda), torch.float32)
    assert torch.equal(got.cpu(), emb[seq] + pos.unsqueeze(0))
    a, a2, rv = torch.randn(3, 37, 256, generator=g), torch.randn(3, 37, 256, generator=g), torch.randn(3, 256, generator=g)
    got = ops.add(a.to(cuda), a2.to(cuda), rv.to(cuda), pos.to(cuda))
    assert torch.allclose(got.cpu(), ((a + rv.unsqueeze(1)) + pos.unsqueeze(0)) + a2, atol=1e-6)
    assert torch.equal(ops.cast(a.to(cuda), torch.bfloat16).cpu(), a.to(torch.bfloat16))


def test_quantize_index_bit_exact(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(4, 1000, generator=g)
    x[0, :300] = 0.0
    x[1, :255] = (torch.arange(255) + 0.5) / 255.0          # exact .5 ties -> half-to-even
    x[2, 0], x[2, 1] = 1.0, -0.25
    got = ops.quantize_index(x.to(cuda)).cpu().long()
    ref = so.quantize_index(x.clamp(max=1.0))
    bad = (got != ref).nonzero()
    assert bad.numel() == 0, (bad[:8], x[got != ref][:8], got[got != ref][:8], ref[got != ref][:8])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_onehot_conv_groupnorm(cuda, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    B, T, C = 2, 77, 320
    idx = torch.randint(0, 257, (B, T), generator=g)
    w = torch.randn(C, 257, 5, generator=g) * 0.1
    bias = torch.randn(C, generator=g) * 0.1
    ref = F.conv1d(F.one_hot(idx, 257).float().transpose(1, 2), w, bias, padding=2)
    wg = w.permute(2, 1, 0).contiguous()
    got = ops.onehot_conv(idx.int().to(cuda), wg.to(cuda), bias.to(cuda), dtype)
    assert rel_err(got.transpose(1, 2), ref) < (1e-5 if dtype == torch.float32 else 8e-3)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xin = got.clone()
    ref2 = F.relu(F.group_norm(xin.float().cpu().transpose(1, 2), C // 16, gamma, beta, 1e-5)).transpose(1, 2)
    ops.groupnorm_relu_(xin, gamma.to(cuda), beta.to(cuda))
    assert rel_err(xin, ref2) < (2e-5 if dtype == torch.float32 else 8e-3)


def test_mel_calibrator(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    B, Tr, C, Lm = 5, 97, 64, 40
    x = torch.randn(B, Tr, C, generator=g)
    mel_len = torch.tensor([97, 40, 13, 96, 41])
    src_len = torch.tensor([40, 40, 40, 7, 39])            # compress, equal, expand, compress, compress(1.05x)
    ref = so.mel_calibrator(x, mel_len, src_len)
    got = ops.mel_calibrator(x.to(cuda), mel_len.to(cuda), src_len.to(cuda), Lm).cpu()
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
def test_gn_calibrator_fused(cuda, dtype):
    """GroupNorm + ReLU applied inside the Mel Calibrator (raw conv output in, normalised tensor never written) against the
    CPU statement F.group_norm -> relu -> mel_calibrator, and against the two-pass kernels; compress / equal / expand lengths."""
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    B, Tr, Cin, C, Lm = 5, 200, 64, 320, 40
    x = torch.randn(B, Tr, Cin, generator=g)
    w = (torch.rand(5, C, Cin, generator=g) * 2 - 1) / math.sqrt(5 * Cin)
    bias = torch.randn(C, generator=g) * 0.1
    gamma, beta = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    mel_len = torch.tensor([200, 40, 13, 193, 41])
    src_len = torch.tensor([40, 40, 40, 7, 39])
    xd, wd = x.to(cuda, dtype), w.to(cuda, dtype)
    part = torch.empty(B, (Tr + 127) // 128, C // 16, 2, device=cuda, dtype=torch.float32)
    y = ops.conv1d(xd, wd, bias.to(cuda), pad=2, impl=ops.IMPL_TC, gn_partial=part)
    raw = y.clone()
    fused = ops.gn_calibrator(y, gamma.to(cuda), beta.to(cuda), part, mel_len.to(cuda), src_len.to(cuda), Lm)
    assert torch.equal(y, raw), "the raw conv output must not be modified"
    two = ops.mel_calibrator(ops.groupnorm_relu_partial_(y, gamma.to(cuda), beta.to(cuda), part), mel_len.to(cuda), src_len.to(cuda), Lm)
    torch.cuda.synchronize()
    yr = raw.float().cpu()                                           # GroupNorm over the padded Tr grid, 16 channels per group
    ref = so.mel_calibrator(F.relu(F.group_norm(yr.transpose(1, 2), C // 16, gamma, beta, 1e-5)).transpose(1, 2), mel_len, src_len)
    tol = {torch.float32: 2e-5, torch.bfloat16: 8e-3, torch.float16: 1e-3}[dtype]
    assert rel_err(fused, ref) < tol, rel_err(fused, ref)
    assert rel_err(two, ref) < 2 * tol


@pytest.mark.parametrize("H,Cin", [(80, 256), (64, 320)])
def test_bilstm(cuda, H, Cin):
    ops = _ops()
    sd = {k: v for k, v in so.make_state_dict(0).items() if "lstm_1." in k or "lstm_2." in k}
    p = "style_modeling.style_encoder.audio_encoder.lstm_%d." % (1 if H == 80 else 2)
    g = torch.Generator().manual_seed(11)
    B, Ln = 3, 50
    x = torch.randn(B, Ln, Cin, generator=g)
    ref = so.bilstm2(sd, p, x)
    cur = x.to(cuda)
    for layer in range(2):
        wih = torch.cat([sd[p + "weight_ih_l%d" % layer], sd[p + "weight_ih_l%d_reverse" % layer]], 0)
        b = torch.cat([sd[p + "bias_ih_l%d" % layer] + sd[p + "bias_hh_l%d" % layer],
                       sd[p + "bias_ih_l%d_reverse" % layer] + sd[p + "bias_hh_l%d_reverse" % layer]], 0)
        whh = torch.stack([sd[p + "weight_hh_l%d" % layer], sd[p + "weight_hh_l%d_reverse" % layer]], 0)
        perm = ops.lstm_quad_order(H)                      # the kernel reads gx as [dir][unit][gate]
        wih, b = wih[perm], b[perm]
        gx = torch.empty(B, Ln, 8 * H, device=cuda)
        ops.conv1d(cur, wih.unsqueeze(0).contiguous().to(cuda), b.to(cuda), out_f32=gx, want_out=False, impl=ops.IMPL_SIMT)
        cur = ops.bilstm_layer(gx, whh.contiguous().to(cuda), torch.float32)
    assert rel_err(cur, ref) < 2e-5


@pytest.mark.parametrize("H", [80, 64])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
def test_bilstm_multi_utterance_form_is_bitwise_the_single_form(cuda, H, dtype):
    """Four utterances per CTA (batches >= 16, LSTM_MULTI) run the same arithmetic in the same order as one utterance per CTA;
    B = 18 leaves a partial last CTA (two absent utterances)."""
    from styler_b200 import _lib
    ops = _ops()
    g = torch.Generator().manual_seed(17)
    B, Ln = 18, 37
    gx = torch.randn(B, Ln, 8 * H, generator=g).to(cuda)
    whh = (torch.randn(2, 4 * H, H, generator=g) * 0.2).to(cuda)
    try:
        _lib.set_tuning("LSTM_MMA", 0)
        _lib.set_tuning("LSTM_MULTI", 0)
        single = ops.bilstm_layer(gx, whh, dtype)
        _lib.set_tuning("LSTM_MULTI", 1)
        multi = ops.bilstm_layer(gx, whh, dtype)
        torch.cuda.synchronize()
    finally:
        _lib.set_tuning("LSTM_MULTI", -1)
        _lib.set_tuning("LSTM_MMA", -1)
    assert torch.isfinite(multi.float()).all()
    assert torch.equal(single, multi)


@pytest.mark.parametrize("H", [80, 64])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("B", [18, 150])        # 8 utterances per CTA / 16 per CTA (more CTAs than SM slots otherwise)
def test_bilstm_tensor_core_recurrence(cuda, H, dtype, B):
    """mma.sync form of the recurrence (16-bit activations, sixteen utterances per CTA; B = 18 leaves a CTA with two utterances)
    against the fp32 recurrence written out in torch on the same gx / W_hh (gate order i, f, g, o; gx in [dir][unit][gate] order),
    and against the CUDA-core kernel.  h enters the product rounded to the activation type -> storage-type tolerance."""
    from styler_b200 import _lib
    ops = _ops()
    g = torch.Generator().manual_seed(19)
    Ln = 41
    gx = torch.randn(B, Ln, 8 * H, generator=g)
    whh = torch.randn(2, 4 * H, H, generator=g) * 0.2
    ref = torch.zeros(B, Ln, 2 * H)
    for d in range(2):
        h, c = torch.zeros(B, H), torch.zeros(B, H)
        steps = range(Ln - 1, -1, -1) if d else range(Ln)
        for t in steps:
            pre = gx[:, t, d * 4 * H:(d + 1) * 4 * H].view(B, H, 4) + (h @ whh[d].t()).view(B, 4, H).transpose(1, 2)
            i_, f_, g_, o_ = torch.sigmoid(pre[..., 0]), torch.sigmoid(pre[..., 1]), torch.tanh(pre[..., 2]), torch.sigmoid(pre[..., 3])
            c = f_ * c + i_ * g_
            h = o_ * torch.tanh(c)
            ref[:, t, d * H:(d + 1) * H] = h
    try:
        _lib.set_tuning("LSTM_MMA", 1)
        got = ops.bilstm_layer(gx.to(cuda), whh.to(cuda), dtype)
        _lib.set_tuning("LSTM_MMA", 0)
        simt = ops.bilstm_layer(gx.to(cuda), whh.to(cuda), dtype)
        torch.cuda.synchronize()
    finally:
        _lib.set_tuning("LSTM_MMA", -1)
    tol = 1.5e-2 if dtype == torch.bfloat16 else 2e-3
    assert torch.isfinite(got.float()).all()
    assert rel_err(simt, ref) < tol
    assert rel_err(got, ref) < tol, rel_err(got, ref)


def test_classifier_tail_duration(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    h = torch.randn(3, 21, 256, generator=g)
    w, b = torch.randn(2, 256, generator=g) * 0.1, torch.randn(2, generator=g)
    ref = F.log_softmax(F.linear(h, w, b), -1).mean(1)
    got = ops.classifier_tail(h.to(cuda), w.to(cuda), b.to(cuda)).cpu()
    assert torch.allclose(got, ref, atol=1e-5)


def test_duration_round_exact(cuda):
    """modules.py:357-358 is an fp -> int boundary: clamp(round(exp(log_d) - log_offset) * d_control, min=0) must be EQUAL to
    the reference arithmetic, not close.  expf on the device and torch.exp on the host may differ in the last ulp, which only
    matters within an ulp of a .5 tie: (1) inputs are drawn with exp(log_d) at least 1e-3 away from every tie -> require
    equality on all of them; (2) the tie rule itself (half to even, torch.round) is pinned with exactly representable ties."""
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    log_d = torch.randn(64, 257, generator=g) * 1.5
    frac = (torch.exp(log_d.double()) - 1.0) % 1.0
    keep = (frac - 0.5).abs() > 1e-3
    log_d = torch.where(keep, log_d, torch.zeros_like(log_d))            # exp(0) - 1 = 0: far from a tie
    assert keep.float().mean() > 0.99
    for ctl in (1.0, 1.3, 0.5):
        got = ops.duration_round(log_d.to(cuda), 1.0, ctl).cpu()
        ref = so.duration_from_log(log_d, ctl)
        assert torch.equal(got, ref), (ctl, (got != ref).sum().item())
        assert (got >= 0).all()
    # exact ties: exp(0) == 1 exactly, so rint(1 - off) sees k + 0.5 with no rounding error; half-to-even, then * d_control
    zeros = torch.zeros(1, 8, device=cuda)
    for off, want in ((0.5, 0.0), (-0.5, 2.0), (-1.5, 2.0), (-2.5, 4.0), (-3.5, 4.0), (1.5, 0.0), (2.5, 0.0)):
        got = ops.duration_round(zeros, off, 1.0).cpu()
        assert (got == want).all(), (off, got[0, 0].item(), want)
        assert (got == torch.clamp(torch.round(torch.tensor(1.0 - off)), min=0)).all()
    assert (ops.duration_round(zeros, -1.5, 1.5).cpu() == 3.0).all()      # round BEFORE scaling by d_control (2 * 1.5)
    # very negative predictions clamp at zero; the LengthRegulator then truncates toward zero (int(x.item()))
    assert (ops.duration_round(torch.full((1, 4), -30.0, device=cuda), 1.0, 1.7).cpu() == 0).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dur_kind", ["int64", "float32"])
def test_length_regulator_bit_exact(cuda, dtype, dur_kind):
    ops = _ops()
    g = torch.Generator().manual_seed(17)
    B, Ln, C = 4, 33, 1280
    x = torch.randn(B, Ln, C, generator=g).to(dtype)
    if dur_kind == "int64":
        dur = torch.randint(0, 9, (B, Ln), generator=g)
        dur[1] = 0                                            # an utterance that expands to nothing
    else:
        dur = torch.rand(B, Ln, generator=g) * 6.0            # 2.6 -> 2 (truncation, modules.py:415-416)
    for Tmax in (None, 50, 400):
        ref, ref_len = so.length_regulator(x.float(), dur, Tmax)
        T = ref.shape[1]
        got, mel_len, cum = ops.length_regulator(x.to(cuda), dur.to(cuda), T)
        assert torch.equal(mel_len.cpu(), ref_len)
        reps = dur.long() if dur_kind == "int64" else dur.double().trunc().long()
        assert torch.equal(cum.cpu().long(), reps.cumsum(1))
        assert torch.equal(got.float().cpu(), ref)            # pure copy: bit exact in any dtype


def test_bucket_embed_sum(cuda):
    ops = _ops()
    sd = so.make_state_dict(0)
    P = "style_modeling."
    g = torch.Generator().manual_seed(19)
    B, T, C = 2, 45, 256
    enc = torch.randn(B, T, 1280, generator=g)
    p = torch.rand(B, T, generator=g) * 900
    e = torch.rand(B, T, generator=g) * 600
    p[0, :5] = sd[P + "pitch_bins"][[0, 1, 100, 253, 254]]     # exactly on a boundary: bins[i-1] < x <= bins[i]
    e[0, :3] = torch.tensor([0.0, 0.1, 1e4])
    pi, ei = torch.bucketize(p, sd[P + "pitch_bins"]), torch.bucketize(e, sd[P + "energy_bins"])
    text, spk, noise = enc[..., :256], enc[..., 512:768], enc[..., 1024:]
    ref = text + F.embedding(pi, sd[P + "pitch_embedding.weight"]) + spk + F.embedding(ei, sd[P + "energy_embedding.weight"])
    encd = enc.to(cuda)
    out, out_n, gpi, gei = ops.bucket_embed_sum(encd[..., :256], encd[..., 512:768], encd[..., 1024:], p.to(cuda), e.to(cuda),
                                                1.0, 1.0, sd[P + "pitch_bins"].to(cuda), sd[P + "energy_bins"].to(cuda),
                                                sd[P + "pitch_embedding.weight"].to(cuda),
                                                sd[P + "energy_embedding.weight"].to(cuda), want_idx=True)
    assert torch.equal(gpi.cpu().long(), pi) and torch.equal(gei.cpu().long(), ei)
    assert torch.equal(out.cpu(), ref)
    assert torch.allclose(out_n.cpu(), ref + noise, atol=1e-6)
    # control factors (modules.py:370,380): scaled predictions come back as separate outputs, the inputs stay untouched, and
    # the embedding rows on their own (predict_inference, modules.py:299-309) are exact copies of the table rows
    pd, ed = p.to(cuda), e.to(cuda)
    res = ops.bucket_embed_sum(None, None, None, pd, ed, 1.3, 0.7, sd[P + "pitch_bins"].to(cuda), sd[P + "energy_bins"].to(cuda),
                               sd[P + "pitch_embedding.weight"].to(cuda), sd[P + "energy_embedding.weight"].to(cuda),
                               want_idx=True, want_scaled=True, want_emb=True, want_sum=False)
    _, _, gpi, gei, psc, esc, pemb, eemb = res
    assert torch.equal(pd.cpu(), p) and torch.equal(ed.cpu(), e), "inputs must not be modified"
    ps, es = p * 1.3, e * 0.7
    assert torch.equal(psc.cpu(), ps) and torch.equal(esc.cpu(), es)
    pi2, ei2 = torch.bucketize(ps, sd[P + "pitch_bins"]), torch.bucketize(es, sd[P + "energy_bins"])
    assert torch.equal(gpi.cpu().long(), pi2) and torch.equal(gei.cpu().long(), ei2)
    assert torch.equal(pemb.cpu(), F.embedding(pi2, sd[P + "pitch_embedding.weight"]))
    assert torch.equal(eemb.cpu(), F.embedding(ei2, sd[P + "energy_embedding.weight"]))


def test_stft_mel(cuda):
    ops = _ops()
    gold = torch.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "stft_b3_n6000.pt"))
    g = torch.Generator().manual_seed(gold["seed"])
    y = (torch.rand(*gold["shape"], generator=g) * 2 - 1) * 0.5
    basis = torch.from_numpy(stft_oracle.slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0))
    mel, energy = ops.stft_mel(y.to(cuda), basis.to(cuda))
    assert (mel.cpu() - gold["mel"]).abs().max() < 1e-3
    assert rel_err(energy, gold["energy"]) < 1e-4
    # a longer, config-4 shaped utterance against the oracle
    y2 = (torch.rand(2, 88200, generator=g) * 2 - 1) * 0.5
    m_ref, e_ref = stft_oracle.mel_spectrogram(y2)
    m2, e2 = ops.stft_mel(y2.to(cuda), basis.to(cuda))
    assert m2.shape == (2, 80, 345)
    assert (m2.cpu() - m_ref).abs().max() < 1e-3 and rel_err(e2, e_ref) < 1e-4


@pytest.mark.parametrize("occ", [1, 2, 3], ids=["three_ctas", "twelve_warps", "direct_loads"])
def test_stft_occupancy_shapes_bitwise(cuda, occ):
    """The STFT kernel's higher-occupancy shapes (STFT_OCC: 16 frames x 3 CTAs per SM / 24 frames x 12 warps, magnitudes aliased onto
    the FFT exchange buffer; 3: samples read from global memory by the frame's own warp instead of a staged block window) run the
    same arithmetic per frame as the 32-frame shape: every output must be bitwise equal, on the
    plain entry and on the extended one (ragged per-utterance lengths, frame-major mel, clamp + clip flag, energy rescaling)."""
    from styler_b200 import _lib
    ops = _ops()
    g = torch.Generator().manual_seed(41)
    basis = torch.from_numpy(stft_oracle.slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0)).to(cuda)
    y = ((torch.rand(5, 30000, generator=g) * 2 - 1) * 0.8).to(cuda)
    ns = torch.tensor([30000, 12345, 700, 29999, 8192], dtype=torch.int64, device=cuda)

    def run():
        a = ops.stft_mel(y, basis)
        b = ops.stft_mel_ex(y * 1.5, basis, in_scale=1.0, clamp=True, frame_major=True, energy_range=(0.1, 525.43), n_samples=ns)
        c = ops.stft_mel_ex(y, basis, n_samples=ns)
        torch.cuda.synchronize()
        return list(a) + list(b) + [c[0], c[1]]

    try:
        _lib.set_tuning("STFT_OCC", 0)
        base = run()
        _lib.set_tuning("STFT_OCC", occ)
        n0 = _lib.launch_count()
        got = run()
        assert _lib.launch_count() == n0 + 6        # band kernel + transform kernel per call
    finally:
        _lib.set_tuning("STFT_OCC", -1)
    assert len(base) == len(got) == 8
    for i, (p, q) in enumerate(zip(base, got)):
        assert torch.isfinite(q.float()).all(), i
        assert torch.equal(p, q), i
    assert base[4].any()                             # the clip flag saw the scaled samples below -1
