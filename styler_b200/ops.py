"""Thin tensor-level wrappers over the C ABI (one Python function per entry point of include/styler_b200.h).

Tensors are torch CUDA tensors used purely as device buffers; every function enqueues hand-written CUDA kernels
on the current stream and returns its output tensor(s).  Activations are [B, T, C] views whose last stride is 1.
"""
import ctypes
import functools

import torch

from . import _lib as L
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, IMPL_AUTO, IMPL_SIMT, IMPL_TC  # noqa: F401


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first tensor argument current: the library launches on the CURRENT device's
    stream (`_lib.stream_ptr()`), so a model living on cuda:1 while cuda:0 is current (nn.DataParallel replicas,
    `model.to('cuda:1')`) must switch first -- otherwise device-1 pointers would be launched on device 0."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in args:
            if torch.is_tensor(a):
                dev = a.device
                break
        if dev is None:
            for a in kwargs.values():
                if torch.is_tensor(a):
                    dev = a.device
                    break
        if dev is None or dev.type != "cuda":
            raise RuntimeError("styler_b200.ops.%s: expected CUDA tensors (the product path has no CPU implementation)" % fn.__name__)
        if dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _v3(t, name="tensor"):
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError("%s must be a [B,T,C] view with unit channel stride, got %s / %s" % (name, tuple(t.shape), t.stride()))
    return t, int(t.stride(0)), int(t.stride(1))


@_on_tensor_device
def conv1d(x, w, bias=None, *, pad=0, act=ACT_NONE, residual=None, residual_row=None, residual_f32=None, ln=None, ln_eps=1e-5,
           act2=ACT_NONE, lens=None, dot=None, out=None, want_out=True, out_f32=None, vt=None, vt_col0=0,
           impl=IMPL_AUTO, dilation=1, act_slope=0.0, residual_inv_lrelu=False, out2_f32=None, gn_partial=None):
    """y = epilogue(conv1d(x, w)) -- see styler_conv1d_fwd.  x [B,T,Cin]; w packed [KS,N,Cin] (same dtype as x).

    residual: [B,T,N] tensor added after `act`; residual_row: [B,N] row broadcast over t instead.
    ln: (gamma, beta) fp32 -> LayerNorm over N;  dot: (w[N] fp32, bias float) -> also returns the [B,T] fp32 row dot.
    vt: preallocated [B, N - vt_col0, Tpad] tensor receiving columns >= vt_col0 transposed.
    dilation: tap spacing; act_slope: negative-side slope of ACT_LRELU; residual_inv_lrelu: `residual` holds lrelu(r).
    out2_f32: second fp32 destination with the strides of out_f32 (e.g. a peer-mapped slice of rank 0's gather buffer).
    gn_partial: fp32 [B, ceil(T/128), N/16, 2] receiving the GroupNorm partial sums of the output (tensor-core path only).
    Returns out (dtype of x) unless want_out=False; with `dot`, returns (out_or_None, dot_out).
    """
    x, x_bs, x_ld = _v3(x, "x")
    L.require_cuda(x, w)
    B, T, Cin = x.shape
    KS, N, Cw = w.shape
    assert Cw == Cin and w.is_contiguous() and w.dtype == x.dtype, (w.shape, x.shape, w.dtype, x.dtype)
    a = L.Conv1dArgs()
    a.x, a.x_bstride, a.x_ld = x.data_ptr(), x_bs, x_ld
    a.B, a.T, a.Cin = B, T, Cin
    a.w, a.N, a.KS, a.pad = w.data_ptr(), N, KS, pad
    a.bias = bias.data_ptr() if bias is not None else None
    a.act, a.act2 = act, act2
    a.dilation, a.act_slope, a.residual_inv_lrelu = int(dilation), float(act_slope), 1 if residual_inv_lrelu else 0
    keep = [x, w, bias]
    if residual is not None:
        r, r_bs, r_ld = _v3(residual, "residual")
        assert r.dtype == x.dtype and r.shape[0] == B and r.shape[2] >= N
        a.residual, a.r_bstride, a.r_ld = r.data_ptr(), r_bs, r_ld
        keep.append(r)
    elif residual_f32 is not None:
        r, r_bs, r_ld = _v3(residual_f32, "residual_f32")
        assert r.dtype == torch.float32 and r.shape[0] == B and r.shape[2] >= N
        a.residual, a.r_bstride, a.r_ld, a.residual_is_f32 = r.data_ptr(), r_bs, r_ld, 1
        keep.append(r)
    elif residual_row is not None:
        assert residual_row.dtype == x.dtype and residual_row.dim() == 2 and residual_row.stride(1) == 1
        a.residual, a.r_bstride, a.r_ld = residual_row.data_ptr(), int(residual_row.stride(0)), 0
        keep.append(residual_row)
    if ln is not None:
        a.ln_gamma, a.ln_beta, a.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), ln_eps
    if lens is not None:
        assert lens.dtype == torch.int64
        a.lens = lens.data_ptr()
    dot_out = None
    if dot is not None:
        dot_out = torch.empty(B, T, device=x.device, dtype=torch.float32)
        a.dot_w, a.dot_b, a.dot_out = dot[0].data_ptr(), float(dot[1]), dot_out.data_ptr()
    if out is None and not want_out and (dot is not None or ln is not None) and out_f32 is None and \
            (impl == IMPL_SIMT or (impl == IMPL_AUTO and B * T < 64)):
        want_out = True      # the CUDA-core path stages pre-LayerNorm rows in the output buffer
    if out is None and want_out:
        ncols = N if vt is None else vt_col0
        out = torch.empty(B, T, ncols, device=x.device, dtype=x.dtype)
    if out is not None:
        o, o_bs, o_ld = _v3(out, "out")
        assert o.dtype == x.dtype
        a.out, a.o_bstride, a.o_ld = o.data_ptr(), o_bs, o_ld
    if out_f32 is not None:
        f, f_bs, f_ld = _v3(out_f32, "out_f32")
        assert f.dtype == torch.float32
        a.out_f32, a.of_bstride, a.of_ld = f.data_ptr(), f_bs, f_ld
    if out2_f32 is not None:
        f2, f2_bs, f2_ld = _v3(out2_f32, "out2_f32")
        assert out_f32 is not None and f2.dtype == torch.float32 and (f2_bs, f2_ld) == (f_bs, f_ld) and f2.shape == f.shape
        a.out2_f32 = f2.data_ptr()
    if vt is not None:
        assert vt.dtype == x.dtype and vt.dim() == 3 and vt.stride(2) == 1
        a.vt, a.vt_col0, a.vt_bstride, a.vt_ld = vt.data_ptr(), vt_col0, int(vt.stride(0)), int(vt.stride(1))
    if gn_partial is not None:
        assert gn_partial.dtype == torch.float32 and gn_partial.is_contiguous() and gn_partial.shape == (B, (T + 127) // 128, N // 16, 2)
        a.gn_partial = gn_partial.data_ptr()
    a.dtype, a.impl = L.dtype_code(x.dtype), impl
    L.check(L.lib().styler_conv1d_fwd(ctypes.byref(a), L.stream_ptr()), "conv1d")
    if dot is not None:
        return out, dot_out
    return out


def make_fft_weights(wqkv, bqkv, wfc, bfc, ln1, w1, b1, w2, b2, ln2, n_head=4, ln_eps=1e-5):
    """Pack the pointers of one FFT block's (already packed, device-resident) weights into the C struct once; the
    returned object keeps the tensors alive."""
    fw = L.FftWeights()
    fw.d_model, fw.d_inner, fw.n_head = wfc.shape[1], w1.shape[1], n_head
    fw.wqkv, fw.bqkv, fw.wfc, fw.bfc = wqkv.data_ptr(), bqkv.data_ptr(), wfc.data_ptr(), bfc.data_ptr()
    fw.ln1_gamma, fw.ln1_beta = ln1[0].data_ptr(), ln1[1].data_ptr()
    fw.w1, fw.b1, fw.ks1 = w1.data_ptr(), b1.data_ptr(), w1.shape[0]
    fw.w2, fw.b2, fw.ks2 = w2.data_ptr(), b2.data_ptr(), w2.shape[0]
    fw.ln2_gamma, fw.ln2_beta, fw.ln_eps = ln2[0].data_ptr(), ln2[1].data_ptr(), ln_eps
    fw._keep = (wqkv, bqkv, wfc, bfc, ln1, w1, b1, w2, b2, ln2)
    return fw


@_on_tensor_device
def fftblock(x, fw, lens, *, out=None, impl=IMPL_AUTO):
    """One FFT block (styler_fftblock_fwd): x [B,T,d_model] -> y [B,T,d_model]; `out` may be a channel slice of a wider buffer."""
    x, x_bs, x_ld = _v3(x, "x")
    B, T, D = x.shape
    assert D == fw.d_model
    if out is None:
        out = torch.empty(B, T, D, device=x.device, dtype=x.dtype)
    o, o_bs, o_ld = _v3(out, "out")
    code = L.dtype_code(x.dtype)
    nbytes = int(L.lib().styler_fftblock_workspace_bytes(B, T, fw.d_model, fw.d_inner, code))
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    L.check(L.lib().styler_fftblock_fwd(ctypes.byref(fw), L.ptr(x), x_bs, x_ld, L.ptr(o), o_bs, o_ld, L.ptr(lens), B, T, code,
                                        impl, L.ptr(ws), nbytes, L.stream_ptr()), "fftblock")
    return out


def make_predictor_weights(c1, b1, ln1, c2, b2, ln2, lw, lb, ln_eps=1e-5):
    """styler_predictor_weights over already packed, device-resident tensors (kept alive by the returned object)."""
    pw = L.PredictorWeights()
    pw.ks, pw.channels, pw.c_in = c1.shape[0], c1.shape[1], c1.shape[2]
    pw.w1, pw.b1, pw.ln1_gamma, pw.ln1_beta = c1.data_ptr(), b1.data_ptr(), ln1[0].data_ptr(), ln1[1].data_ptr()
    pw.w2, pw.b2, pw.ln2_gamma, pw.ln2_beta = c2.data_ptr(), b2.data_ptr(), ln2[0].data_ptr(), ln2[1].data_ptr()
    pw.lin_w, pw.lin_b, pw.ln_eps = lw.data_ptr(), float(lb), ln_eps
    pw._keep = (c1, b1, ln1, c2, b2, ln2, lw)
    return pw


def make_decoder_weights(fft_structs, mel_w, mel_b, postnet):
    """styler_decoder_weights: `fft_structs` = list of FftWeights, `postnet` = list of (w [KS,N,Cin], b) or empty."""
    dw = L.DecoderWeights()
    arr = (L.FftWeights * len(fft_structs))(*fft_structs)
    dw.n_layers, dw.layers = len(fft_structs), arr
    dw.mel_w, dw.mel_b, dw.n_mel = mel_w.data_ptr(), mel_b.data_ptr(), mel_w.shape[1]
    pn = None
    if postnet:
        pn = L.PostnetWeights()
        pn.n_layers, pn.n_mel, pn.channels, pn.ks = len(postnet), mel_w.shape[1], postnet[0][0].shape[1], postnet[0][0].shape[0]
        for j, (w, b) in enumerate(postnet):
            pn.w[j], pn.b[j] = w.data_ptr(), b.data_ptr()
        dw.postnet = ctypes.pointer(pn)
    dw._keep = (arr, fft_structs, mel_w, mel_b, postnet, pn)
    return dw


@_on_tensor_device
def predictor(x, pw, lens, *, impl=IMPL_AUTO):
    """StylePredictor.forward in one native call (styler_predictor_fwd): x [B,T,C] view -> fp32 [B,T]."""
    x, x_bs, x_ld = _v3(x, "x")
    B, T, _ = x.shape
    code = L.dtype_code(x.dtype)
    out = torch.empty(B, T, device=x.device, dtype=torch.float32)
    nbytes = int(L.lib().styler_predictor_workspace_bytes(B, T, pw.channels, code))
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    L.check(L.lib().styler_predictor_fwd(ctypes.byref(pw), L.ptr(x), x_bs, x_ld, L.ptr(lens), L.ptr(out), B, T, code, impl,
                                         L.ptr(ws), nbytes, L.stream_ptr()), "predictor")
    return out


@_on_tensor_device
def decoder(x, dw, pos, lens, *, mel_out=None, post_out=None, mel_out2=None, post_out2=None, impl=IMPL_AUTO):
    """STYLER.decode in one native call (styler_decoder_fwd): x [B,T,D] contiguous -> (mel fp32 [B,T,n_mel], post fp32 or None).
    pos=None: x already carries the position rows (bucket_embed_sum(..., pos=...)) and is consumed in place."""
    assert x.is_contiguous() and (pos is None or (pos.is_contiguous() and pos.dtype == torch.float32 and pos.shape[0] >= x.shape[1]))
    B, T, _ = x.shape
    code = L.dtype_code(x.dtype)
    nm = dw.n_mel
    has_post = bool(dw.postnet)
    if mel_out is None:
        mel_out = torch.empty(B, T, nm, device=x.device, dtype=torch.float32)
    if post_out is None and has_post:
        post_out = torch.empty(B, T, nm, device=x.device, dtype=torch.float32)
    for t in (mel_out, post_out, mel_out2, post_out2):
        assert t is None or (t.is_contiguous() and t.dtype == torch.float32 and t.shape == (B, T, nm))
    nbytes = int(L.lib().styler_decoder_workspace_bytes(ctypes.byref(dw), B, T, code))
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    L.check(L.lib().styler_decoder_fwd(ctypes.byref(dw), L.ptr(x), L.ptr(pos), L.ptr(lens), L.ptr(mel_out), L.ptr(post_out),
                                       L.ptr(mel_out2), L.ptr(post_out2) if has_post else None, B, T, code, impl, L.ptr(ws), nbytes,
                                       L.stream_ptr()), "decoder")
    return mel_out, (post_out if has_post else None)


@_on_tensor_device
def attention(qk, vt, lens, n_head=4, *, out=None, impl=IMPL_AUTO):
    """ctx[B,T,H*64] = softmax(mask(Q K^T)) V ; qk [B,T,2*H*64] (Q pre-scaled) + vt [B,H*64,Tpad], or the fused
    qkv [B,T,3*H*64] with vt=None (V read row-major)."""
    qk, qk_bs, qk_ld = _v3(qk, "qk")
    B, T, _ = qk.shape
    D = n_head * 64
    if out is None:
        out = torch.empty(B, T, D, device=qk.device, dtype=qk.dtype)
    o, o_bs, o_ld = _v3(out, "ctx")
    L.check(L.lib().styler_attention_fwd(L.ptr(qk), qk_bs, qk_ld, L.ptr(vt), int(vt.stride(0)) if vt is not None else 0,
                                         int(vt.stride(1)) if vt is not None else 0,
                                         L.ptr(lens), L.ptr(o), o_bs, o_ld, B, T, n_head, L.dtype_code(qk.dtype), impl,
                                         L.stream_ptr()), "attention")
    return out


@_on_tensor_device
def lrelu_mean(a, b=None, c=None, *, slope_in, slope_out, out=None):
    """out = lrelu(mean_k x_k, slope_out) where the inputs hold y_k = lrelu(x_k, slope_in) (styler_lrelu_mean_fwd)."""
    L.require_cuda(a)
    ts = [t for t in (a, b, c) if t is not None]
    assert all(t.is_contiguous() and t.shape == a.shape and t.dtype == a.dtype for t in ts)
    if out is None:
        out = torch.empty_like(a)
    L.check(L.lib().styler_lrelu_mean_fwd(L.ptr(a), L.ptr(b), L.ptr(c), float(slope_in), float(slope_out), L.ptr(out),
                                          a.numel(), L.dtype_code(a.dtype), L.stream_ptr()), "lrelu_mean")
    return out


@_on_tensor_device
def embed_pos(src_seq, emb, pos, dtype):
    B, Ln = src_seq.shape
    D = emb.shape[1]
    out = torch.empty(B, Ln, D, device=src_seq.device, dtype=dtype)
    L.check(L.lib().styler_embed_pos_fwd(L.ptr(src_seq), L.ptr(emb), emb.shape[0], L.ptr(pos), L.ptr(out), B, Ln, D,
                                         L.dtype_code(dtype), L.stream_ptr()), "embed_pos")
    return out


@_on_tensor_device
def add(a, a2=None, rowvec=None, pos=None, out=None):
    """out = (a or 0) (+ a2) (+ rowvec[b] broadcast over t) (+ pos[t] fp32 broadcast over b)."""
    if a is None:
        o, o_bs, o_ld = _v3(out, "out")
        B, T, C = o.shape
        a_bs = a_ld = 0
    else:
        a, a_bs, a_ld = _v3(a, "a")
        B, T, C = a.shape
        if out is None:
            out = torch.empty(B, T, C, device=a.device, dtype=a.dtype)
        o, o_bs, o_ld = _v3(out, "out")
    a2p, a2_bs, a2_ld = None, 0, 0
    if a2 is not None:
        a2, a2_bs, a2_ld = _v3(a2, "a2")
        a2p = L.ptr(a2)
    L.check(L.lib().styler_add_fwd(L.ptr(a), a_bs, a_ld, a2p, a2_bs, a2_ld, L.ptr(rowvec),
                                   int(rowvec.stride(0)) if rowvec is not None else 0, L.ptr(pos), L.ptr(o), o_bs, o_ld,
                                   B, T, C, L.dtype_code(o.dtype), L.stream_ptr()), "add")
    return out


@_on_tensor_device
def cast(x, dtype):
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    L.check(L.lib().styler_cast_fwd(L.ptr(x), L.ptr(out), x.numel(), L.dtype_code(dtype), L.stream_ptr()), "cast")
    return out


@_on_tensor_device
def quantize_index(x):
    x = x.contiguous()
    idx = torch.empty(x.shape, device=x.device, dtype=torch.int32)
    L.check(L.lib().styler_quantize_index_fwd(L.ptr(x), L.ptr(idx), x.numel(), L.stream_ptr()), "quantize_index")
    return idx


@_on_tensor_device
def onehot_conv(idx, wg, bias, dtype):
    """idx int32 [B,T]; wg fp32 [KS, nidx, C]; -> [B,T,C]."""
    B, T = idx.shape
    KS, nidx, C = wg.shape
    out = torch.empty(B, T, C, device=idx.device, dtype=dtype)
    L.check(L.lib().styler_onehot_conv_fwd(L.ptr(idx), L.ptr(wg), L.ptr(bias), L.ptr(out), B, T, C, nidx, KS,
                                           L.dtype_code(dtype), L.stream_ptr()), "onehot_conv")
    return out


@_on_tensor_device
def groupnorm_relu_(x, gamma, beta, ch_per_group=16, eps=1e-5):
    x3, bs, ld = _v3(x, "x")
    B, T, C = x3.shape
    ws = torch.empty(B * (C // ch_per_group) * 2, device=x.device, dtype=torch.float32)
    L.check(L.lib().styler_groupnorm_relu_fwd(L.ptr(x3), bs, ld, L.ptr(gamma), L.ptr(beta), L.ptr(ws), B, T, C,
                                              ch_per_group, eps, L.dtype_code(x.dtype), L.stream_ptr()), "groupnorm_relu")
    return x


@_on_tensor_device
def groupnorm_relu_partial_(x, gamma, beta, partial, eps=1e-5):
    """GroupNorm(16 channels per group) + ReLU in place from the partial statistics the producing conv1d left in `partial`."""
    x3, bs, ld = _v3(x, "x")
    B, T, C = x3.shape
    ws = torch.empty(B * (C // 16) * 2, device=x.device, dtype=torch.float32)
    L.check(L.lib().styler_groupnorm_relu_partial_fwd(L.ptr(x3), bs, ld, L.ptr(gamma), L.ptr(beta), L.ptr(partial), partial.shape[1],
                                                      L.ptr(ws), B, T, C, eps, L.dtype_code(x.dtype), L.stream_ptr()),
            "groupnorm_relu_partial")
    return x


@_on_tensor_device
def mel_calibrator(x, mel_len, src_len, Lmax, out=None):
    x, x_bs, x_ld = _v3(x, "x")
    B, Tr, C = x.shape
    if out is None:
        out = torch.empty(B, Lmax, C, device=x.device, dtype=x.dtype)
    o, o_bs, o_ld = _v3(out, "out")
    L.check(L.lib().styler_mel_calibrator_fwd(L.ptr(x), x_bs, x_ld, L.ptr(mel_len), L.ptr(src_len), L.ptr(o), o_bs, o_ld,
                                              B, Tr, Lmax, C, L.dtype_code(x.dtype), L.stream_ptr()), "mel_calibrator")
    return out


@_on_tensor_device
def gn_calibrator(x, gamma, beta, partial, mel_len, src_len, Lmax, eps=1e-5, out=None):
    """GroupNorm(16 channels per group, statistics from the producing conv's `partial`) + ReLU + Mel Calibrator in one pass over
    the RAW conv output x (not modified): styler_gn_calibrator_fwd."""
    x, x_bs, x_ld = _v3(x, "x")
    B, Tr, C = x.shape
    if out is None:
        out = torch.empty(B, Lmax, C, device=x.device, dtype=x.dtype)
    o, o_bs, o_ld = _v3(out, "out")
    ws = torch.empty(B * (C // 16) * 2, device=x.device, dtype=torch.float32)
    L.check(L.lib().styler_gn_calibrator_fwd(L.ptr(x), x_bs, x_ld, L.ptr(gamma), L.ptr(beta), L.ptr(partial), partial.shape[1], L.ptr(ws),
                                             L.ptr(mel_len), L.ptr(src_len), L.ptr(o), o_bs, o_ld, B, Tr, Lmax, C, eps,
                                             L.dtype_code(x.dtype), L.stream_ptr()), "gn_calibrator")
    return out


def lstm_quad_order(H):
    """Row permutation that takes the stacked input projection [W_ih_fwd ; W_ih_rev] (PyTorch order [dir][gate i,f,g,o][unit])
    to the order the BiLSTM kernel reads gx in: [dir][unit][gate] -- the four gates of a unit are adjacent, element t of a
    direction's row belongs to thread t."""
    idx = torch.arange(8 * H).view(2, 4, H)            # value = original row
    return idx.permute(0, 2, 1).reshape(-1)


@_on_tensor_device
def bilstm_layer(gx, whh, dtype, out=None):
    """gx fp32 [B,L,8H] in quad order (see lstm_quad_order); whh fp32 [2,4H,H] (PyTorch row order) -> [B,L,2H]."""
    B, Ln, G8 = gx.shape
    H = G8 // 8
    assert gx.is_contiguous() and gx.dtype == torch.float32 and whh.is_contiguous()
    if out is None:
        out = torch.empty(B, Ln, 2 * H, device=gx.device, dtype=dtype)
    o, o_bs, o_ld = _v3(out, "out")
    L.check(L.lib().styler_bilstm_layer_fwd(L.ptr(gx), L.ptr(whh), L.ptr(o), o_bs, o_ld, B, Ln, H, L.dtype_code(dtype),
                                            L.stream_ptr()), "bilstm_layer")
    return out


@_on_tensor_device
def classifier_tail(h, w, b):
    h, h_bs, h_ld = _v3(h, "h")
    B, Ln, C = h.shape
    out = torch.empty(B, 2, device=h.device, dtype=torch.float32)
    L.check(L.lib().styler_classifier_tail_fwd(L.ptr(h), h_bs, h_ld, L.ptr(w), L.ptr(b), L.ptr(out), B, Ln, C,
                                               L.dtype_code(h.dtype), L.stream_ptr()), "classifier_tail")
    return out


@_on_tensor_device
def duration_round(log_d, log_offset=1.0, d_control=1.0):
    log_d = log_d.contiguous()
    out = torch.empty_like(log_d)
    L.check(L.lib().styler_duration_round_fwd(L.ptr(log_d), L.ptr(out), log_d.numel(), log_offset, d_control,
                                              L.stream_ptr()), "duration_round")
    return out


@_on_tensor_device
def length_regulator(x, duration, Tmax, out=None):
    """x [B,L,C]; duration int64 or float32 [B,L]; -> (out [B,Tmax,C], mel_len int64 [B], cum int32 [B,L])."""
    x, x_bs, x_ld = _v3(x, "x")
    B, Ln, C = x.shape
    duration = duration.contiguous()
    if out is None:
        out = torch.empty(B, Tmax, C, device=x.device, dtype=x.dtype)
    o, o_bs, o_ld = _v3(out, "out") if Tmax > 0 else (out, 0, C)
    mel_len = torch.empty(B, device=x.device, dtype=torch.int64)
    cum = torch.empty(B, Ln, device=x.device, dtype=torch.int32)
    d64 = duration if duration.dtype == torch.int64 else None
    d32 = duration if duration.dtype == torch.float32 else None
    if d64 is None and d32 is None:
        raise TypeError("duration must be int64 or float32")
    L.check(L.lib().styler_length_regulator_fwd(L.ptr(x), x_bs, x_ld, L.ptr(d64), L.ptr(d32), L.ptr(o), o_bs, o_ld,
                                                L.ptr(mel_len), L.ptr(cum), B, Ln, Tmax, C, L.dtype_code(x.dtype),
                                                L.stream_ptr()), "length_regulator")
    return out, mel_len, cum


def length_regulator_scan(duration):
    """Integer part only (scan + totals), used to size the output before the expand: -> (mel_len, cum)."""
    B, Ln = duration.shape
    dummy = torch.empty(B, Ln, 8, device=duration.device, dtype=torch.float32)
    _, mel_len, cum = length_regulator(dummy, duration, 0, out=dummy)
    return mel_len, cum


@_on_tensor_device
def bucket_embed_sum(text, spk, noise, p_val, e_val, p_scale, e_scale, pitch_bins, energy_bins, pitch_emb, energy_emb,
                     want_noisy=True, want_idx=False, out=None, out_noisy=None, want_scaled=False, want_emb=False,
                     want_sum=True, emb_dtype=None, pos=None):
    """bucketize + embedding (+ 4-way sum), styler_bucket_embed_sum_fwd.  Inputs are never modified.
    pos: fp32 [>=T, C] position rows added to out / out_noisy (the decoder is then run with pos=None).
    Returns (out, out_noisy, p_idx, e_idx) and, appended when asked for, (p_scaled, e_scaled) = the predictions times their
    control factors (fp32) and (pitch_embedding, energy_embedding) = the two embedding rows on their own (dtype of text)."""
    p_val, e_val = p_val.contiguous(), e_val.contiguous()
    assert p_val.dtype == torch.float32 and e_val.dtype == torch.float32 and p_val.shape == e_val.shape
    L.require_cuda(p_val, e_val)
    B, T = p_val.shape
    C = pitch_emb.shape[1]
    dev = p_val.device
    in_bs = in_ld = 0
    dt = None
    if want_sum:
        text, in_bs, in_ld = _v3(text, "text")
        spk3, s_bs, s_ld = _v3(spk, "spk")
        assert (s_bs, s_ld) == (in_bs, in_ld), "text/spk/noise must share strides (slices of one expanded buffer)"
        assert text.shape == (B, T, C)
        dt = text.dtype
        if out is None:
            out = torch.empty(B, T, C, device=dev, dtype=dt)
        out_n = (out_noisy if out_noisy is not None else torch.empty_like(out)) if want_noisy else None
        assert out.is_contiguous() and (out_n is None or out_n.is_contiguous())
    else:
        text = spk3 = noise = out = out_n = None
        want_noisy = False
    assert pos is None or (pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape[0] >= T and pos.shape[1] == C and want_sum)
    p_idx = torch.empty(B, T, device=dev, dtype=torch.int32) if want_idx else None
    e_idx = torch.empty(B, T, device=dev, dtype=torch.int32) if want_idx else None
    p_sc = torch.empty(B, T, device=dev, dtype=torch.float32) if want_scaled else None
    e_sc = torch.empty(B, T, device=dev, dtype=torch.float32) if want_scaled else None
    emb_dt = dt if dt is not None else (emb_dtype or torch.float32)
    p_emb = torch.empty(B, T, C, device=dev, dtype=emb_dt) if want_emb else None
    e_emb = torch.empty(B, T, C, device=dev, dtype=emb_dt) if want_emb else None
    L.check(L.lib().styler_bucket_embed_sum_fwd(L.ptr(text), L.ptr(spk3), L.ptr(noise) if want_noisy else None, in_bs,
                                                in_ld, L.ptr(p_val), L.ptr(e_val), p_scale, e_scale, L.ptr(pitch_bins),
                                                L.ptr(energy_bins), pitch_bins.numel(), L.ptr(pitch_emb),
                                                L.ptr(energy_emb), L.ptr(out), L.ptr(out_n),
                                                int(out.stride(0)) if out is not None else 0,
                                                int(out.stride(1)) if out is not None else 0, L.ptr(p_idx), L.ptr(e_idx),
                                                L.ptr(p_sc), L.ptr(e_sc), L.ptr(p_emb), L.ptr(e_emb), L.ptr(pos), B, T, C,
                                                L.dtype_code(emb_dt), L.stream_ptr()), "bucket_embed_sum")
    res = (out, out_n, p_idx, e_idx)
    if want_scaled:
        res = res + (p_sc, e_sc)
    if want_emb:
        res = res + (p_emb, e_emb)
    return res


@_on_tensor_device
def stft_mel(y, mel_basis):
    """y fp32 [B,N] -> (mel fp32 [B,n_mels,F], energy fp32 [B,F])."""
    y = y.contiguous()
    B, N = y.shape
    n_mels = mel_basis.shape[0]
    F = 1 + N // 256
    mel = torch.empty(B, n_mels, F, device=y.device, dtype=torch.float32)
    energy = torch.empty(B, F, device=y.device, dtype=torch.float32)
    band = torch.empty(2 * n_mels, device=y.device, dtype=torch.int32)
    L.check(L.lib().styler_stft_mel_fwd(L.ptr(y), B, N, L.ptr(mel_basis), n_mels, L.ptr(band), L.ptr(mel), L.ptr(energy),
                                        L.stream_ptr()), "stft_mel")
    return mel, energy


@_on_tensor_device
def stft_mel_ex(y, mel_basis, *, in_scale=1.0, clamp=False, frame_major=False, energy_range=None, n_samples=None):
    """Fused front end (styler_stft_mel_ex_fwd): returns (mel, energy, clip_flag int32 [B] or None, e_input or None).
    n_samples (int64 [B], device): true length of every zero-padded row -- each utterance is reflected around its own end
    and frames beyond 1 + n_samples[b] // 256 come back as zeros."""
    y = y.contiguous()
    B, N = y.shape
    n_mels = mel_basis.shape[0]
    F = 1 + N // 256
    mel = torch.empty((B, F, n_mels) if frame_major else (B, n_mels, F), device=y.device, dtype=torch.float32)
    energy = torch.empty(B, F, device=y.device, dtype=torch.float32)
    band = torch.empty(2 * n_mels, device=y.device, dtype=torch.int32)
    flag = torch.empty(B, device=y.device, dtype=torch.int32) if clamp else None
    e_in = torch.empty(B, F, device=y.device, dtype=torch.float32) if energy_range is not None else None
    e_min, e_max = (float(energy_range[0]), float(energy_range[1])) if energy_range is not None else (0.0, 1.0)
    if n_samples is not None:
        assert n_samples.dtype == torch.int64 and n_samples.is_cuda and n_samples.numel() == B
        n_samples = n_samples.contiguous()
    L.check(L.lib().styler_stft_mel_ex_fwd(L.ptr(y), B, N, L.ptr(mel_basis), n_mels, L.ptr(band), L.ptr(mel), L.ptr(energy),
                                           float(in_scale), 1 if clamp else 0, L.ptr(flag), 1 if frame_major else 0,
                                           L.ptr(e_in), e_min, e_max, L.ptr(n_samples), L.stream_ptr()), "stft_mel_ex")
    return mel, energy, flag, e_in


@_on_tensor_device
def f0_norm(f0, lens=None):
    """f0_normalization (utils.py:387-409) over a padded [B,T] fp32 batch of log-f0 contours."""
    f0 = f0.contiguous()
    L.require_cuda(f0)
    assert f0.dtype == torch.float32 and f0.dim() == 2
    out = torch.empty_like(f0)
    L.check(L.lib().styler_f0_norm_fwd(L.ptr(f0), L.ptr(lens), L.ptr(out), f0.shape[0], f0.shape[1], L.stream_ptr()), "f0_norm")
    return out
