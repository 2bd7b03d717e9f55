"""PyTorch custom ops (namespace ``sdod``) over the C ABI of libsdod_b200.so.

Each op is a thin marshalling layer: tensors -> device pointers + sizes -> one ``sdod_*`` C call on the
current CUDA stream.  CUDA tensors only — there is no CPU path (the product fails loudly instead).

Op surface mirrored from the reference: ``sdod::GroupNorm(input, weight, bias; num_groups, eps)`` and
``sdod::ParameterlessGroupNorm`` (sdod/efficient_gn.py:14-26, csrc/sdod_ops/config/group_norm.xml).
"""
from typing import Optional

import torch

from . import _cabi as C

_ws_cache = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise C.SdodError("sdod ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def _dt(t):
    if t.dtype == torch.float32:
        return C.F32
    if t.dtype == torch.bfloat16:
        return C.BF16
    raise C.SdodError("unsupported dtype %s (float32 / bfloat16 only)" % t.dtype)


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _gn_workspace(device, nbytes):
    key = (device.index, _stream())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)   # zero-filled once; kernels keep it clean
        _ws_cache[key] = ws
    return ws


class SdodTypeError(TypeError):
    pass


@torch.library.custom_op("sdod::group_norm", mutates_args=(), device_types="cuda")
def group_norm(x: torch.Tensor, num_groups: int, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float,
               silu: bool = False, add_nc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm(+SiLU)(+x+add_nc[n,c]) on [N,C,*spatial].  Kernels: fp32 / bf16, NCHW-contiguous (any shape) or channels_last (C % 8 == 0,
    num_groups <= 64, N <= 256).  Like the reference's F.group_norm every other floating input is still served: fp16 / fp64 are computed in
    fp32 and cast back, channels_last tensors outside the NHWC kernels' constraints go through the NCHW kernel on a contiguous copy."""
    _need_cuda(x, weight, bias, add_nc)
    n, c = x.shape[0], x.shape[1]
    hw = 1
    for s in x.shape[2:]:
        hw *= s
    if c % num_groups != 0:
        raise ValueError("num_channels must be divisible by num_groups")
    if x.dtype not in (torch.float32, torch.bfloat16):
        if not x.is_floating_point():
            raise SdodTypeError("group_norm: floating-point input required, got %s" % x.dtype)
        return group_norm(x.float(), num_groups, weight, bias, eps, silu, add_nc).to(x.dtype)
    nhwc = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
    if nhwc and (c % 8 != 0 or num_groups > 64 or n > 256):
        nhwc = False
    if not nhwc:
        x = x.contiguous()
    y = torch.empty_like(x)
    w, b, a = _f32(weight), _f32(bias), _f32(add_nc)
    layout = C.NHWC if nhwc else C.NCHW
    ws, ws_n = None, 0
    if nhwc:
        ws_n = C.lib().sdod_group_norm_workspace(n, c, hw, num_groups, layout)
        ws = _gn_workspace(x.device, ws_n)
    C.check(C.lib().sdod_group_norm(_stream(), _p(x), _p(y), _p(w), _p(b), _p(a), n, c, hw, num_groups, eps, _dt(x), layout,
                                    int(silu), _p(ws), ws.numel() if ws is not None else 0), "sdod_group_norm")
    return y


@group_norm.register_fake
def _(x, num_groups, weight, bias, eps, silu=False, add_nc=None):
    return torch.empty_like(x)


def group_norm_nhwc(x, num_groups, weight, bias, eps, silu=False, add_nc=None, out_dtype=None):
    """x: [N, HW, C] (tokens x channels) bf16/f32 contiguous — the layout used inside the UNet/VAE.
    out_dtype may differ from x.dtype (fp32 residual stream in, bf16 GEMM operand out)."""
    _need_cuda(x)
    n, hw, c = x.shape
    y = torch.empty_like(x, dtype=out_dtype or x.dtype)
    w, b, a = _f32(weight), _f32(bias), _f32(add_nc)
    ws_n = C.lib().sdod_group_norm_workspace(n, c, hw, num_groups, C.NHWC)
    ws = _gn_workspace(x.device, ws_n)
    C.check(C.lib().sdod_group_norm_nhwc(_stream(), _p(x), _dt(x), _p(y), _dt(y), _p(w), _p(b), _p(a), n, c, hw, num_groups, eps,
                                         int(silu), _p(ws), ws.numel()), "sdod_group_norm_nhwc")
    return y


def group_norm_nhwc2(xa, xb, num_groups, weight, bias, eps, silu=False, want_raw=False, out_dtype=None):
    """Single-launch GroupNorm over the channel concatenation [xa | xb] (xb may be None); xa/xb: [N, HW, C*] contiguous.
    Returns y, or (y, raw) with raw = bf16 copy of the un-normalised concatenation.  L2-resident tensors only."""
    _need_cuda(xa)
    n, hw, ca = xa.shape
    cb = 0 if xb is None else xb.shape[2]
    y = torch.empty(n, hw, ca + cb, dtype=out_dtype or xa.dtype, device=xa.device)
    raw = torch.empty(n, hw, ca + cb, dtype=torch.bfloat16, device=xa.device) if want_raw else None
    w, b = _f32(weight), _f32(bias)
    ws = _gn_workspace(xa.device, C.lib().sdod_group_norm_workspace(n, ca + cb, hw, num_groups, C.NHWC))
    C.check(C.lib().sdod_group_norm_nhwc2(_stream(), _p(xa), ca, _p(xb), cb, _dt(xa), _p(y), _dt(y), _p(raw), _p(w), _p(b), n, hw, num_groups, eps,
                                          int(silu), _p(ws), ws.numel()), "sdod_group_norm_nhwc2")
    return (y, raw) if want_raw else y


@torch.library.custom_op("sdod::layer_norm", mutates_args=(), device_types="cuda")
def layer_norm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x, dtype=torch.bfloat16)
    w, b = _f32(weight), _f32(bias)
    rows = x.numel() // x.shape[-1]
    C.check(C.lib().sdod_layer_norm(_stream(), _p(x), _dt(x), _p(y), _p(w), _p(b), rows, x.shape[-1], eps), "sdod_layer_norm")
    return y


@layer_norm.register_fake
def _(x, weight, bias, eps):
    return torch.empty_like(x, dtype=torch.bfloat16)


def dpm_schedule(steps=20, timesteps=1000, lin_start=0.00085, lin_end=0.0120):
    """Host tables, bit-identical to the reference DPMSolver (dpm_solver.cpp:84-131)."""
    import ctypes
    import numpy as np
    names = ["ts", "log_alphas", "lambdas", "sigmas", "alphas", "phis", "i2rs", "model_ts"]
    arrs = {k: np.zeros(steps + 1, dtype=np.float32) for k in names}
    C.check(C.lib().sdod_dpm_schedule(timesteps, lin_start, lin_end, steps, *[arrs[k].ctypes.data_as(ctypes.c_void_p) for k in names]),
            "sdod_dpm_schedule")
    return arrs


def dpm_coeffs(step, steps=20, timesteps=1000, lin_start=0.00085, lin_end=0.0120):
    import ctypes
    f = [ctypes.c_float() for _ in range(5)]
    order = ctypes.c_int()
    C.check(C.lib().sdod_dpm_coeffs(timesteps, lin_start, lin_end, steps, step, *[ctypes.cast(ctypes.byref(v), ctypes.c_void_p) for v in f],
                                    ctypes.cast(ctypes.byref(order), ctypes.c_void_p)), "sdod_dpm_coeffs")
    return dict(sigma_s=f[0].value, alpha_s=f[1].value, c_x=f[2].value, c_prev=f[3].value, c_y0=f[4].value, order=order.value)


def ddim_schedule(steps=20, timesteps=1000, lin_start=0.00085, lin_end=0.0120):
    """DDIM (eta 0) tables for sdod::cfg_dpm_step: model_ts [steps] and per-step dicts (order 1).  Public CompVis ddim.py, parity unpinned."""
    import ctypes
    import numpy as np
    ts = np.zeros(steps, dtype=np.float32)
    co = np.zeros((steps, 5), dtype=np.float32)
    C.check(C.lib().sdod_ddim_schedule(timesteps, lin_start, lin_end, steps, ts.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p)),
            "sdod_ddim_schedule")
    keys = ("sigma_s", "alpha_s", "c_x", "c_prev", "c_y0")
    return ts, [dict(zip(keys, (float(v) for v in row)), order=1) for row in co]


@torch.library.custom_op("sdod::cfg_dpm_step", mutates_args=("x", "y_prev"), device_types="cuda")
def cfg_dpm_step(x: torch.Tensor, y_prev: torch.Tensor, eps_c: torch.Tensor, eps_u: Optional[torch.Tensor], guidance: float,
                 sigma_s: float, alpha_s: float, c_x: float, c_prev: float, c_y0: float, order: int) -> None:
    """Fused CFG combine + eps->x0 + DPM-Solver++(2M) update, in place on fp32 x / y_prev."""
    _need_cuda(x, y_prev, eps_c, eps_u)
    assert x.dtype == torch.float32 and y_prev.dtype == torch.float32 and x.is_contiguous() and y_prev.is_contiguous()
    eps_c = eps_c.contiguous()
    eps_u = None if eps_u is None else eps_u.contiguous()
    C.check(C.lib().sdod_cfg_dpm_step(_stream(), _p(x), _p(y_prev), _p(eps_c), _p(eps_u), _dt(eps_c), x.numel(), guidance, sigma_s,
                                      alpha_s, c_x, c_prev, c_y0, order, None), "sdod_cfg_dpm_step")


def timestep_sinusoid(t, dim=320, max_period=10000.0):
    _need_cuda(t)
    t = t.to(torch.float32).contiguous()
    out = torch.empty(t.numel(), dim, dtype=torch.float32, device=t.device)
    C.check(C.lib().sdod_timestep_sinusoid(_stream(), _p(t), t.numel(), dim, max_period, _p(out)), "sdod_timestep_sinusoid")
    return out


def randn(n, seed, offset=0, device="cuda"):
    out = torch.empty(n, dtype=torch.float32, device=device)
    C.check(C.lib().sdod_randn(_stream(), _p(out), n, seed, offset), "sdod_randn")
    return out


def image_to_u8(img):
    _need_cuda(img)
    img = img.contiguous()
    out = torch.empty(img.shape, dtype=torch.uint8, device=img.device)
    C.check(C.lib().sdod_image_to_u8(_stream(), _p(img), _dt(img), _p(out), img.numel()), "sdod_image_to_u8")
    return out


def _epilogue(out, bias=None, row_bias=None, rows_per_group=0, residual=None, alpha=1.0, act=C.ACT_NONE, out_mode=None,
              c2=None, c3=None, heads=0, head_dim=0, tokens=0, dpad=0, tok_pad=0, vt_rows=0):
    e = C.Epilogue()
    e.C, e.C2, e.C3 = _p(out), _p(c2), _p(c3)
    e.ldc = out.stride(-2) if out.dim() >= 2 else 0
    e.strideC = out.stride(0) if out.dim() == 3 else 0
    e.bias, e.row_bias, e.rows_per_group = _p(bias), _p(row_bias), rows_per_group
    e.residual = _p(residual)
    e.residual_f32 = int(residual is not None and residual.dtype == torch.float32)
    if residual is not None:
        e.ldr = residual.stride(-2)
        e.strideR = residual.stride(0) if residual.dim() == 3 else 0
    e.alpha, e.act = alpha, act
    e.out_mode = out_mode if out_mode is not None else (C.OUT_F32 if out.dtype == torch.float32 else C.OUT_BF16)
    e.heads, e.head_dim, e.tokens, e.dpad, e.tok_pad, e.vt_rows = heads, head_dim, tokens, dpad, tok_pad, vt_rows
    return e


@torch.library.custom_op("sdod::linear", mutates_args=(), device_types="cuda")
def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
           act: int = 0, alpha: float = 1.0, out_f32: bool = False, row_bias: Optional[torch.Tensor] = None, rows_per_group: int = 0,
           block_n: int = 0, a2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(alpha * [a | a2] @ w^T + bias + row_bias[row // rows_per_group]) + residual on the tcgen05 GEMM.
    a [M,K] or [B,M,K] bf16; w [N,K (+K2)] or [B,N,K] bf16 (K % 64 == 0); a2 [M,K2] optional second operand (K-concatenated).
    act=GEGLU expects w rows packed per 256-row tile."""
    _need_cuda(a, w, bias, residual, row_bias, a2)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    a, w = a.contiguous(), w.contiguous()
    batch = a.shape[0] if a.dim() == 3 else 1
    M, K = a.shape[-2], a.shape[-1]
    N = w.shape[-2]
    n_out = N // 2 if act == C.ACT_GEGLU else N
    shape = (batch, M, n_out) if a.dim() == 3 else (M, n_out)
    out = torch.empty(shape, dtype=torch.float32 if out_f32 else torch.bfloat16, device=a.device)
    bias, row_bias = _f32(bias), _f32(row_bias)
    residual = None if residual is None else residual.contiguous()
    d = C.GemmDesc()
    d.A, d.lda, d.strideA = _p(a), K, M * K
    d.W, d.ldw, d.strideW = _p(w), w.shape[-1], (N * K if w.dim() == 3 else 0)
    d.M, d.N, d.K, d.batch, d.block_n = M, N, K, batch, block_n
    if a2 is not None:
        a2 = a2.contiguous()
        assert a2.dtype == torch.bfloat16 and a2.shape[0] == M and w.shape[-1] == K + a2.shape[1]
        d.A2, d.lda2, d.K2 = _p(a2), a2.shape[1], a2.shape[1]
    d.epi = _epilogue(out, bias, row_bias, rows_per_group, residual, alpha, act)
    C.check(C.lib().sdod_gemm_bf16(_stream(), d), "sdod_gemm_bf16")
    return out


def linear_accumulate_(x: torch.Tensor, a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """In place on the fp32 stream: x += alpha * a @ w^T + bias (x fp32 [M,N] contiguous, a [M,K] / w [N,K] bf16).  The residual IS the
    output tensor, so eligible shapes take the TMA reduce-add epilogue (no residual load in the kernel); others add and store as usual."""
    _need_cuda(x, a, w, bias)
    assert x.dtype == torch.float32 and x.is_contiguous() and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    a, w = a.contiguous(), w.contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert x.shape == (M, N)
    d = C.GemmDesc()
    d.A, d.lda, d.strideA = _p(a), K, M * K
    d.W, d.ldw, d.strideW = _p(w), K, 0
    d.M, d.N, d.K, d.batch, d.block_n = M, N, K, 1, 0
    d.epi = _epilogue(x, _f32(bias), None, 0, x, alpha, C.ACT_NONE)
    C.check(C.lib().sdod_gemm_bf16(_stream(), d), "sdod_gemm_bf16 (in place)")
    return x


@linear.register_fake
def _(a, w, bias=None, residual=None, act=0, alpha=1.0, out_f32=False, row_bias=None, rows_per_group=0, block_n=0, a2=None):
    n = w.shape[-2] // 2 if act == C.ACT_GEGLU else w.shape[-2]
    return a.new_empty(a.shape[:-1] + (n,), dtype=torch.float32 if out_f32 else torch.bfloat16)


def linear_ln(a, w, bias, residual, ln_weight, ln_bias, eps=1e-5):
    """(y fp32, LayerNorm(y) bf16) with y = a @ w^T + bias + residual: the LayerNorm runs in the GEMM's epilogue (N tiles of a row
    block = one thread-block cluster).  Raises SdodError when the shape does not allow the fusion."""
    _need_cuda(a, w, bias, residual, ln_weight, ln_bias)
    a, w = a.contiguous(), w.contiguous()
    M, K = a.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=a.device)
    ln = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    bias, ln_weight, ln_bias = _f32(bias), _f32(ln_weight), _f32(ln_bias)
    residual = None if residual is None else residual.contiguous()
    d = C.GemmDesc()
    d.A, d.lda, d.strideA = _p(a), K, M * K
    d.W, d.ldw, d.strideW = _p(w), K, 0
    d.M, d.N, d.K, d.batch, d.block_n = M, N, K, 1, 0
    e = _epilogue(y, bias, None, 0, residual, 1.0, 0)
    e.ln_out, e.ld_ln, e.ln_weight, e.ln_bias, e.ln_eps = _p(ln), N, _p(ln_weight), _p(ln_bias), eps
    d.epi = e
    C.check(C.lib().sdod_gemm_bf16(_stream(), d), "sdod_gemm_bf16 (+LayerNorm)")
    return y, ln


@torch.library.custom_op("sdod::conv3x3", mutates_args=(), device_types="cuda")
def conv3x3(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            row_bias: Optional[torch.Tensor] = None, act: int = 0, block_n: int = 0, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Implicit-GEMM conv3x3 (stride 1, pad 1).  x [B,H,W,Cin] bf16 NHWC, wt [Cout, 9*Cin] bf16 (k=(ky*3+kx)*Cin+c),
    row_bias [B,Cout] (timestep-embedding add), residual [B,H,W,Cout].  x2 [B,H,W,Cin2]: a fused 1x1 convolution whose
    weights are wt[:, 9*Cin:].  Returns [B,H,W,Cout] bf16."""
    _need_cuda(x, wt, bias, residual, row_bias, x2)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and x.dim() == 4
    x, wt = x.contiguous(), wt.contiguous()
    B, H, W, Cin = x.shape
    Cout = wt.shape[0]
    out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=x.device)
    bias, row_bias = _f32(bias), _f32(row_bias)
    residual = None if residual is None else residual.contiguous()
    d = C.ConvDesc()
    d.X, d.Wt, d.B, d.H, d.W, d.Cin, d.Cout, d.block_n = _p(x), _p(wt), B, H, W, Cin, Cout, block_n
    if x2 is not None:
        x2 = x2.contiguous()
        assert x2.dtype == torch.bfloat16 and x2.shape[:3] == x.shape[:3] and wt.shape[1] == 9 * Cin + x2.shape[3]
        d.X2, d.ldx2, d.Cin2 = _p(x2), x2.shape[3], x2.shape[3]
    e = _epilogue(out.view(B * H * W, Cout), bias, row_bias, H * W, None, 1.0, act)
    if residual is not None:
        e.residual, e.ldr, e.strideR, e.residual_f32 = _p(residual), Cout, 0, int(residual.dtype == torch.float32)
    d.epi = e
    C.check(C.lib().sdod_conv3x3_bf16(_stream(), d), "sdod_conv3x3_bf16")
    return out


@conv3x3.register_fake
def _(x, wt, bias=None, residual=None, row_bias=None, act=0, block_n=0, x2=None):
    return x.new_empty(x.shape[:3] + (wt.shape[0],))


def conv3x3_s2(x, wt, bias=None, out_f32=False, block_n=0):
    """Stride-2 conv3x3 with padding 1 (UNet Downsample) as an implicit GEMM on TMA boxes with element strides 2.
    x [B,H,W,Cin] bf16 NHWC (H, W even), wt [Cout, 9*Cin] bf16 as pack_conv3x3_weight lays it out -> [B,H/2,W/2,Cout]."""
    _need_cuda(x, wt, bias)
    assert x.dtype == torch.bfloat16 and wt.dtype == torch.bfloat16 and x.dim() == 4
    x, wt = x.contiguous(), wt.contiguous()
    B, H, W, Cin = x.shape
    Cout = wt.shape[0]
    out = torch.empty(B, H // 2, W // 2, Cout, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    d = C.ConvDesc()
    d.X, d.Wt, d.B, d.H, d.W, d.Cin, d.Cout, d.block_n, d.stride = _p(x), _p(wt), B, H, W, Cin, Cout, block_n, 2
    d.epi = _epilogue(out.view(-1, Cout), _f32(bias), None, 0, None, 1.0, 0)
    C.check(C.lib().sdod_conv3x3_bf16(_stream(), d), "sdod_conv3x3_bf16")
    return out


def conv3x3_up2(x, w_oihw, bias=None, out_f32=False, block_n=0):
    """conv3x3(nearest_upsample_2x(x)) without the upsampled tensor (sub-pixel form, 4/9 of the multiply-adds).
    x [B,H,W,Cin] bf16 NHWC, w_oihw [Cout,Cin,3,3] fp32 -> [B,2H,2W,Cout] bf16 (or fp32)."""
    _need_cuda(x, w_oihw, bias)
    assert x.dtype == torch.bfloat16 and x.dim() == 4
    x = x.contiguous()
    w = w_oihw.detach().to(torch.float32).contiguous()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    wt = torch.empty(4, Cout, 4 * Cin, dtype=torch.bfloat16, device=x.device)
    C.check(C.lib().sdod_pack_conv3x3_up2_weight(_stream(), _p(w), _p(wt), Cout, Cin), "sdod_pack_conv3x3_up2_weight")
    out = torch.empty(B, 2 * H, 2 * W, Cout, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    d = C.ConvDesc()
    d.X, d.Wt, d.B, d.H, d.W, d.Cin, d.Cout, d.block_n, d.upsample2x = _p(x), _p(wt), B, H, W, Cin, Cout, block_n, 1
    d.epi = _epilogue(out.view(B * 4 * H * W, Cout), _f32(bias), None, 0, None, 1.0, 0)
    C.check(C.lib().sdod_conv3x3_bf16(_stream(), d), "sdod_conv3x3_bf16")
    return out


def pack_conv3x3_weight(w_oihw, kpad=None):
    """[Cout,Cin,3,3] fp32 -> [Cout, 9*Cin (padded to kpad)] bf16 with k = (ky*3+kx)*Cin + c."""
    _need_cuda(w_oihw)
    w = w_oihw.detach().to(torch.float32).contiguous()
    cout, cin = w.shape[0], w.shape[1]
    kpad = kpad or 9 * cin
    out = torch.empty(cout, kpad, dtype=torch.bfloat16, device=w.device)
    C.check(C.lib().sdod_pack_conv3x3_weight(_stream(), _p(w), _p(out), cout, cin, kpad), "sdod_pack_conv3x3_weight")
    return out


def pack_geglu_weight(w, bias, block_n=128):
    """Interleave GEGLU projection rows so each block_n-row tile holds [value half | gate half]."""
    n2 = w.shape[0]
    n = n2 // 2
    half = block_n // 2
    assert n % half == 0
    idx = []
    for t in range(n // half):
        idx += list(range(t * half, (t + 1) * half)) + list(range(n + t * half, n + (t + 1) * half))
    idx = torch.tensor(idx, device=w.device)
    return w[idx].contiguous(), (None if bias is None else bias[idx].contiguous())


def im2col3x3(x, stride=1, kpad=None):
    _need_cuda(x)
    x = x.contiguous()
    B, H, W, Cc = x.shape
    kpad = kpad or 9 * Cc
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    out = torch.empty(B * Ho * Wo, kpad, dtype=torch.bfloat16, device=x.device)
    C.check(C.lib().sdod_im2col3x3(_stream(), _p(x), _dt(x), _p(out), B, H, W, Cc, stride, kpad), "sdod_im2col3x3")
    return out


def upsample2x(x):
    _need_cuda(x)
    x = x.contiguous()
    B, H, W, Cc = x.shape
    out = torch.empty(B, 2 * H, 2 * W, Cc, dtype=torch.bfloat16, device=x.device)
    C.check(C.lib().sdod_upsample2x_nhwc(_stream(), _p(x), _dt(x), _p(out), B, H, W, Cc), "sdod_upsample2x_nhwc")
    return out


def concat_channels(a, b):
    _need_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    rows = a.numel() // a.shape[-1]
    out = torch.empty(a.shape[:-1] + (a.shape[-1] + b.shape[-1],), dtype=a.dtype, device=a.device)
    assert a.dtype == b.dtype
    C.check(C.lib().sdod_concat_channels(_stream(), _p(a), a.shape[-1], _p(b), b.shape[-1], _p(out), rows, _dt(a)), "sdod_concat_channels")
    return out


def softmax_rows(x, scale=1.0):
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    rows = x.numel() // x.shape[-1]
    C.check(C.lib().sdod_softmax_rows(_stream(), _p(x), _p(y), rows, x.shape[-1], x.shape[-1], scale), "sdod_softmax_rows")
    return y


def nchw_f32_to_nhwc_bf16(x):
    _need_cuda(x)
    x = x.to(torch.float32).contiguous()
    n, c = x.shape[:2]
    hw = x.numel() // (n * c)
    out = torch.empty((n,) + tuple(x.shape[2:]) + (c,), dtype=torch.bfloat16, device=x.device)
    C.check(C.lib().sdod_nchw_f32_to_nhwc_bf16(_stream(), _p(x), _p(out), n, c, hw), "sdod_nchw_f32_to_nhwc_bf16")
    return out


def nhwc_to_nchw_f32(x):
    _need_cuda(x)
    x = x.contiguous()
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    out = torch.empty((n, c) + tuple(x.shape[1:-1]), dtype=torch.float32, device=x.device)
    C.check(C.lib().sdod_nhwc_to_nchw_f32(_stream(), _p(x), _dt(x), _p(out), n, c, hw), "sdod_nhwc_to_nchw_f32")
    return out


_splitk = {}


def enable_splitk(enable=True, device="cuda", mbytes=64):
    """Give this thread's GEMM / conv calls a split-K scratch (small-M layers); enable=False turns it off."""
    if not enable:
        C.check(C.lib().sdod_set_splitk_workspace(None, 0, None, 0), "sdod_set_splitk_workspace")
        return
    key = torch.device(device).index or 0
    if key not in _splitk:
        _splitk[key] = (torch.empty(mbytes << 18, dtype=torch.float32, device=device), torch.zeros(4096, dtype=torch.int32, device=device))
    ws, cnt = _splitk[key]
    C.check(C.lib().sdod_set_splitk_workspace(ws.data_ptr(), ws.numel() * 4, cnt.data_ptr(), cnt.numel()), "sdod_set_splitk_workspace")


def launch_count():
    return int(C.lib().sdod_launch_count())


# ------------------------------------------------------------------------------------------ attention
def head_geometry(head_dim, n_kv):
    """(dpad, vt_rows, kv_pad) of the attention operand layouts (include/sdod_kernels.h)."""
    return 64 * ((head_dim + 63) // 64), 16 * ((head_dim + 15) // 16), 8 * ((n_kv + 7) // 8)


def qkv_project(x, w_qkv, heads, head_dim, tokens):
    """x [B*tokens, Cin] bf16, w_qkv [3*heads*head_dim, Cin] -> (Qh, Kh, Vt) written directly in the attention
    operand layouts by the GEMM epilogue (SDOD_OUT_QKV)."""
    _need_cuda(x, w_qkv)
    x, w_qkv = x.contiguous(), w_qkv.contiguous()
    M, K = x.shape
    B = M // tokens
    dpad, vt_rows, kv_pad = head_geometry(head_dim, tokens)
    qh = torch.zeros(B * heads, tokens, dpad, dtype=torch.bfloat16, device=x.device)
    kh = torch.zeros(B * heads, tokens, dpad, dtype=torch.bfloat16, device=x.device)
    vt = torch.zeros(B * heads, vt_rows, kv_pad, dtype=torch.bfloat16, device=x.device)
    d = C.GemmDesc()
    d.A, d.lda, d.strideA = _p(x), K, M * K
    d.W, d.ldw, d.strideW = _p(w_qkv), K, 0
    d.M, d.N, d.K, d.batch, d.block_n = M, w_qkv.shape[0], K, 1, 0
    d.epi = _epilogue(qh, out_mode=C.OUT_QKV, c2=kh, c3=vt, heads=heads, head_dim=head_dim, tokens=tokens, dpad=dpad,
                      tok_pad=kv_pad, vt_rows=vt_rows)
    C.check(C.lib().sdod_gemm_bf16(_stream(), d), "sdod_gemm_bf16")
    return qh, kh, vt


def pack_heads(t, heads, head_dim, transpose=False):
    """[B, N, heads*head_dim] -> HEADS [B*heads, N, dpad] or HEADS_T [B*heads, vt_rows, kv_pad] (torch-side packing)."""
    B, N, _ = t.shape
    dpad, vt_rows, kv_pad = head_geometry(head_dim, N)
    th = t.reshape(B, N, heads, head_dim).permute(0, 2, 1, 3).reshape(B * heads, N, head_dim)
    if not transpose:
        out = torch.zeros(B * heads, N, dpad, dtype=t.dtype, device=t.device)
        out[:, :, :head_dim] = th
    else:
        out = torch.zeros(B * heads, vt_rows, kv_pad, dtype=t.dtype, device=t.device)
        out[:, :head_dim, :N] = th.transpose(1, 2)
    return out


@torch.library.custom_op("sdod::attention", mutates_args=(), device_types="cuda")
def attention(qh: torch.Tensor, kh: torch.Tensor, vt: torch.Tensor, batch: int, heads: int, head_dim: int, n_kv: int,
              scale: float) -> torch.Tensor:
    """Fused flash-style attention.  qh [B*heads,Nq,dpad], kh [B*heads,Nkv,dpad], vt [B*heads,vt_rows,kv_pad]
    -> [B, Nq, heads*head_dim] bf16."""
    _need_cuda(qh, kh, vt)
    qh, kh, vt = qh.contiguous(), kh.contiguous(), vt.contiguous()
    nq, dpad = qh.shape[1], qh.shape[2]
    out = torch.empty(batch, nq, heads * head_dim, dtype=torch.bfloat16, device=qh.device)
    C.check(C.lib().sdod_attention_bf16(_stream(), _p(qh), _p(kh), _p(vt), _p(out), batch, heads, nq, n_kv, head_dim, dpad,
                                        vt.shape[2], scale), "sdod_attention_bf16")
    return out


def attention_causal(qh, kh, vt, batch, heads, head_dim, n_kv, scale):
    """attention() with a causal mask (key k visible to query q iff k <= q): the CLIP text encoder's self-attention."""
    _need_cuda(qh, kh, vt)
    qh, kh, vt = qh.contiguous(), kh.contiguous(), vt.contiguous()
    nq, dpad = qh.shape[1], qh.shape[2]
    out = torch.empty(batch, nq, heads * head_dim, dtype=torch.bfloat16, device=qh.device)
    C.check(C.lib().sdod_attention_causal_bf16(_stream(), _p(qh), _p(kh), _p(vt), _p(out), batch, heads, nq, n_kv, head_dim, dpad,
                                               vt.shape[2], scale), "sdod_attention_causal_bf16")
    return out


@attention.register_fake
def _(qh, kh, vt, batch, heads, head_dim, n_kv, scale):
    return qh.new_empty(batch, qh.shape[1], heads * head_dim)
