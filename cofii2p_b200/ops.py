"""Tensor-level wrappers over the C ABI (include/cofi_b200.h).

PyTorch is used here for device memory and streams only: every wrapper allocates its output with
`torch.empty`, passes raw device pointers + sizes + the current CUDA stream to libcofi_b200.so and returns
the output tensor.  There is no fallback of any kind: CPU tensors raise, a missing library raises at import.
All launches are CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import lib as _libmod



class _Lib:
    """libcofi_b200.so, opened on first use (not at import: bench.py's CPU reference arm constructs the model class only to
    mint its state_dict and must not map the product library into its process).  A missing or ABI-mismatched library
    raises ImportError / AttributeError at the first call -- there is no fallback of any kind."""

    def __getattr__(self, name):
        return getattr(_libmod.load(), name)


_lib = _Lib()

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3
ENGINE_FP32, ENGINE_TF32, ENGINE_TF32X3 = 0, 1, 2
ENGINE_TF32X3S = 3   # C-ABI only: 3xTF32 with pre-split weights (selected automatically, see _x3_weights)
_ENGINES = {"fp32": ENGINE_FP32, "tf32": ENGINE_TF32, "tf32x3": ENGINE_TF32X3}
_engine = ENGINE_FP32


# Named presets = global engine + per-group policy (groups are tagged in the model code, see `group`).
# "parity": the engine bench.py reports -- everything in 3xTF32 (fp32-grade on the tensor cores) except the two stages whose
# reduced-precision variants were measured harmless (tools/precision_sweep.py, profiles/r2_precision_sweep.md): the tcgen05
# flash-attention kernel (tf32 operands: 8.6e-5 on the golden outputs) and the KPConv aggregate / weight-apply pair (fp16
# operands: 5.1e-4 on the fine point features, <= 1.7e-4 elsewhere); correspondences stay bit-identical to the reference.
PRESETS = {"parity": ("tf32x3", {"tr_attn": "tf32", "kpconv": "tf32"})}
_preset = None


def set_engine(name: str) -> None:
    """Select the contraction engine for GEMM/conv/attention: 'fp32' (SIMT, exact-order parity engine), 'tf32' (tcgen05
    kind::tf32), 'tf32x3' (tcgen05 3xTF32 split, fp32-grade), or a preset of PRESETS (global engine + per-group policy)."""
    global _engine, _preset
    if name in PRESETS:
        base, pol = PRESETS[name]
        _engine, _preset = _ENGINES[base], name
        set_policy(pol)
        return
    _engine, _preset = _ENGINES[name], None
    set_policy(None)


def get_engine() -> str:
    return _preset if _preset is not None else {v: k for k, v in _ENGINES.items()}[_engine]


# Weights epoch: every cache derived from parameter VALUES (K-major weight packs, fp16 copies, folded BatchNorm, captured
# CUDA graphs) is keyed on it.  Raw-pointer updates (the fused Adam kernel, graph replays that update BatchNorm running
# statistics) do not bump tensor._version, so whoever changes parameter values outside autograd's view calls
# bump_weights_epoch() (TrainStep.step, CoFiI2P.load_state_dict, TrainStep._flatten).
_weights_epoch = 0


def bump_weights_epoch() -> int:
    global _weights_epoch
    _weights_epoch += 1
    return _weights_epoch


def weights_epoch() -> int:
    return _weights_epoch


# Per-group precision policy.  The model tags its stages (`with ops.group("kpconv"): ...`); a policy maps a group name to
# an engine and overrides the global engine for the launches issued inside that group (innermost tagged group with a policy
# entry wins).  `set_engine("parity")` is such a preset (see PRESETS): 3xTF32 everywhere except the stages measured
# harmless at lower precision.  Host-side only: a captured CUDA graph bakes the choice in.
_policy: dict = {}
_groups: list = []


def set_policy(policy: Optional[dict]) -> None:
    """policy: {group name: engine name} (None / {} clears it)."""
    global _policy
    _policy = {k: _ENGINES[v] for k, v in (policy or {}).items()}


def get_policy() -> dict:
    inv = {v: k for k, v in _ENGINES.items()}
    return {k: inv[v] for k, v in _policy.items()}


class group:
    """Context manager tagging the launches of a model stage (see set_policy)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        _groups.append(self.name)
        return self

    def __exit__(self, *exc):
        _groups.pop()
        return False


def _eng() -> int:
    if _policy:
        for g in reversed(_groups):
            e = _policy.get(g)
            if e is not None:
                return e
    return _engine


def _chk(rc: int, name: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {_libmod.last_error()}")


# Optional per-op profiling (bench.py's roofline leg): when enabled every C-ABI call is bracketed by CUDA
# events on the launching stream and annotated with the algorithmic flops / bytes of that launch.
_prof = None
_prof_meta = (0.0, 0.0)
_prof_tag = ""
_prof_calls = None      # raw per-call records of the last profile_stop: [(name, ms, launches, flops, bytes, tag)]


def profile_start() -> None:
    global _prof
    _prof = []


def profile_stop(ridge: Optional[float] = None, raw: bool = False):
    """-> {op: dict(calls, ms, flops, bytes)} (synchronises).  With `ridge` (flop per byte at which the tensor roof meets the
    HBM roof) the contraction entry points are split per call into "<op>|tensor" (arithmetic intensity above the ridge) and
    "<op>|hbm": one entry point serves both the K <= 128 streaming contractions and the large-K ones, and a single
    roofline for the mix would describe neither."""
    global _prof, _prof_calls
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    calls = []
    _prof_calls = [(name, e0.elapsed_time(e1), nl, fl, by, tag) for name, e0, e1, nl, fl, by, tag in rec]
    for name, e0, e1, nl, fl, by, _tag in rec:
        if ridge is not None and (name.startswith("cofi_gemm") or name.startswith("cofi_conv2d")) and by > 0:
            name = name + ("|tensor" if fl / by > ridge else "|hbm")
        calls.append((name, nl))
        d = out.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return (out, calls) if raw else out   # calls: [(op name, kernels launched)] in launch order (tools/ncu_traffic.py)


def _meta(flops: float = 0.0, nbytes: float = 0.0, tag: str = "") -> None:
    """Algorithmic work (and a shape tag) of the NEXT _call (consumed by the profiler only)."""
    global _prof_meta, _prof_tag
    _prof_meta = (float(flops), float(nbytes))
    _prof_tag = tag


def _call(name: str, *args) -> None:
    global _prof_meta, _prof_tag
    if _prof is None:
        rc = getattr(_lib, name)(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _libmod.launch_count()
        e0.record()
        rc = getattr(_lib, name)(*args)
        e1.record()
        _prof.append((name, e0, e1, _libmod.launch_count() - n0) + _prof_meta + (_prof_tag,))
    _prof_meta = (0.0, 0.0)
    _prof_tag = ""
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {_libmod.last_error()}")


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (cofii2p_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    return t


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (cofii2p_b200 has no CPU path)")
    if t.dtype != torch.int64:
        raise RuntimeError(f"{name}: expected int64, got {t.dtype}")
    if not t.is_contiguous():
        t = t.contiguous()
    return t


def _rows(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """2-D fp32 row-major view with unit inner stride; returns (tensor, leading dimension)."""
    _f32(t, name)
    if t.dim() != 2:
        raise RuntimeError(f"{name}: expected a 2-D tensor, got {tuple(t.shape)}")
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    return t, ld


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------ point stream
def pack_points(points: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
    points = _f32(points, "points").contiguous()
    feats, ldf = _rows(feats, "feats")
    rows = points.shape[0]
    out = torch.empty((rows, 4), dtype=torch.float32, device=points.device)
    _call("cofi_pack_points", _p(points), _p(feats), ldf, feats.shape[1], rows, _p(out), _st())
    return out


def kpconv_aggregate(feats, s_packed, q_points, nbr, kernel_points, sigma: float, frames: int = 1,
                     kp_reach: float = 0.0):
    feats, ldf = _rows(feats, "feats")
    q_points = _f32(q_points, "q_points").contiguous()
    nbr = _i64(nbr, "nbr")
    kernel_points = _f32(kernel_points, "kernel_points").contiguous()
    total_q, H = nbr.shape
    Mq = total_q // frames
    Ns = s_packed.shape[0] // frames
    C, K = feats.shape[1], kernel_points.shape[0]
    agg = torch.empty((total_q, K * C), dtype=torch.float32, device=feats.device)
    cnt = torch.empty((total_q,), dtype=torch.float32, device=feats.device)
    # compulsory HBM bytes (aggregate stage): indices + features + packed coords + queries + output
    _meta(2.0 * total_q * K * H * C,
          8.0 * total_q * H + 4.0 * feats.shape[0] * C + 16.0 * s_packed.shape[0] + 12.0 * total_q
          + 4.0 * total_q * K * C + 4.0 * total_q)
    _call("cofi_kpconv_aggregate", _p(feats), ldf, C, _p(s_packed), _p(q_points), _p(nbr), H, Mq, Ns, frames,
                                    _p(kernel_points), K, float(sigma), float(kp_reach), _p(agg), _p(cnt), _st())
    return agg, cnt


def kpconv_aggregate_f16(feats, s_packed, q_points, nbr, kernel_points, sigma: float, frames: int = 1,
                         kp_reach: float = 0.0):
    """fp16 aggregate [M, K*C] (tf32 engine); see cofi_kpconv_aggregate_f16."""
    feats, ldf = _rows(feats, "feats")
    q_points = _f32(q_points, "q_points").contiguous()
    nbr = _i64(nbr, "nbr")
    kernel_points = _f32(kernel_points, "kernel_points").contiguous()
    total_q, H = nbr.shape
    Mq, Ns = total_q // frames, s_packed.shape[0] // frames
    C, K = feats.shape[1], kernel_points.shape[0]
    agg = torch.empty((total_q, K * C), dtype=torch.float16, device=feats.device)
    cnt = torch.empty((total_q,), dtype=torch.float32, device=feats.device)
    _meta(2.0 * total_q * K * H * C,
          8.0 * total_q * H + 4.0 * feats.shape[0] * C + 16.0 * s_packed.shape[0] + 12.0 * total_q
          + 2.0 * total_q * K * C + 4.0 * total_q)
    _call("cofi_kpconv_aggregate_f16", _p(feats), ldf, C, _p(s_packed), _p(q_points), _p(nbr), H, Mq, Ns, frames,
          _p(kernel_points), K, float(sigma), float(kp_reach), _p(agg), _p(cnt), _st())
    return agg, cnt


def gemm_f16(a_half, w_half, bias=None, rowdiv=None, act: int = ACT_NONE):
    """fp16 operands, fp32 output (tcgen05 kind::f16)."""
    if a_half.dtype != torch.float16 or w_half.dtype != torch.float16 or not a_half.is_cuda:
        raise RuntimeError("gemm_f16: CUDA fp16 operands expected")
    a_half, w_half = a_half.contiguous(), w_half.contiguous()
    M, K = a_half.shape
    N = w_half.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a_half.device)
    _meta(2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N, f"{M}x{N}x{K} f16")
    _call("cofi_gemm_f16", _p(a_half), K, _p(w_half), K, _p(out), N, M, N, K, _p(bias), _p(rowdiv), act, _st())
    return out


def maxpool_rows(x, nbr, frames: int = 1):
    x, ldx = _rows(x, "x")
    nbr = _i64(nbr, "nbr")
    total_q, H = nbr.shape
    out = torch.empty((total_q, x.shape[1]), dtype=torch.float32, device=x.device)
    _call("cofi_maxpool_rows", _p(x), ldx, x.shape[1], _p(nbr), H, total_q // frames, x.shape[0] // frames, frames,
                                _p(out), out.stride(0), _st())
    return out


def maxpool_rows_f16(x_half, nbr, frames: int = 1):
    """max over neighbours on an fp16 copy of the features (tf32 engine); fp32 result."""
    if x_half.dtype != torch.float16 or not x_half.is_cuda:
        raise RuntimeError("maxpool_rows_f16: CUDA fp16 features expected")
    x_half = x_half.contiguous()
    nbr = _i64(nbr, "nbr")
    total_q, H = nbr.shape
    C = x_half.shape[1]
    out = torch.empty((total_q, C), dtype=torch.float32, device=x_half.device)
    _meta(1.0 * total_q * H * C, 8.0 * total_q * H + 2.0 * x_half.numel() + 4.0 * total_q * C)
    _call("cofi_maxpool_rows_f16", _p(x_half), C, C, _p(nbr), H, total_q // frames, x_half.shape[0] // frames, frames,
          _p(out), C, _st())
    return out


def gather_rows(x, idx: Optional[torch.Tensor], idx_stride: int = 1, frames: int = 1, out: Optional[torch.Tensor] = None,
                rows_out: Optional[int] = None):
    """out[i,:C] = x[idx[i*idx_stride]] per frame (idx None -> identity copy). `out` may be a column slice of a
    wider buffer."""
    x, ldx = _rows(x, "x")
    C = x.shape[1]
    if idx is not None:
        if not idx.is_cuda or idx.dtype != torch.int64:
            raise RuntimeError("gather_rows: idx must be a CUDA int64 tensor")
        total_q = rows_out if rows_out is not None else (idx.numel() // idx_stride)
    else:
        total_q = x.shape[0]
    if out is None:
        out = torch.empty((total_q, C), dtype=torch.float32, device=x.device)
    if out.stride(1) != 1:
        raise RuntimeError("gather_rows: out must have unit inner stride")
    _call("cofi_gather_rows", _p(x), ldx, C, _p(idx), idx_stride, total_q // frames, x.shape[0] // frames, frames,
                               _p(out), out.stride(0), _st())
    return out


# ------------------------------------------------------------------------------------------ contractions
# 3xTF32 with pre-split weights (csrc/gemm_x3.cu): the weight operand of a contraction is a constant of the forward pass, so
# its two tf32-exact planes [2, N, K] are computed once (cofi_split_tf32) and cached per (storage, view, weights epoch).
_split_cache: dict = {}
_split_epoch = -1


def _is_weight(w: torch.Tensor) -> bool:
    """True for nn.Parameters and views of them (w1[:, :C]); activations used as the W operand are split in the kernel."""
    base = w._base if w._base is not None else w
    return isinstance(w, torch.nn.Parameter) or isinstance(base, torch.nn.Parameter)


def split_tf32(w: torch.Tensor) -> torch.Tensor:
    """[2, N, K] fp32: plane 0 = w rounded to tf32, plane 1 = (w - plane 0) rounded to tf32 (cached, see above)."""
    global _split_epoch
    if _split_epoch != _weights_epoch:
        _split_cache.clear()
        _split_epoch = _weights_epoch
    w, ldw = _rows(w, "w")
    key = (w.data_ptr(), tuple(w.shape), ldw, getattr(w, "_version", 0))
    hit = _split_cache.get(key)
    if hit is not None:
        return hit[0]
    N, K = w.shape
    out = torch.empty((2, N, K), dtype=torch.float32, device=w.device)
    _call("cofi_split_tf32", _p(w), ldw, N, K, _p(out), _st())
    _split_cache[key] = (out, w)   # keep the source alive: the key holds its address
    return out


def _x3_weights(w: torch.Tensor, eng: int, const_w: Optional[bool]):
    """(W pointer tensor, ldw, engine) for a contraction: under 3xTF32 a constant weight operand is replaced by its cached
    split planes and the persistent TF32X3S kernel is selected."""
    if eng == ENGINE_TF32X3 and (const_w if const_w is not None else _is_weight(w)) and w.shape[0] >= 16 and w.shape[1] % 4 == 0:
        return split_tf32(w), w.shape[1], ENGINE_TF32X3S
    w, ldw = _rows(w, "w")
    return w, ldw, eng


def gemm(a, w, bias=None, rowdiv=None, act: int = ACT_NONE, out: Optional[torch.Tensor] = None,
         accumulate: bool = False, engine: Optional[int] = None, const_w: Optional[bool] = None):
    """act((a @ w.T) / rowdiv[:,None] + bias (+ out if accumulate)); w is [N,K] (nn.Linear layout).  const_w: w is a
    constant of the forward pass (default: decided by _is_weight)."""
    a, lda = _rows(a, "a")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise RuntimeError(f"gemm: K mismatch {a.shape} x {w.shape}")
    w, ldw, eng = _x3_weights(w, _eng() if engine is None else engine, const_w)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ldc = out.stride(0) if M > 1 else max(N, out.stride(0))
    _meta(2.0 * M * N * K, 4.0 * (M * K + N * K + M * N), f"{M}x{N}x{K}")
    _call("cofi_gemm", _p(a), lda, _p(w), ldw, _p(out), ldc, M, N, K, _p(bias), _p(rowdiv), int(accumulate), act, eng, _st())
    return out


def gemm_colstats(a, w, bias=None, rowdiv=None, const_w: Optional[bool] = None, out: Optional[torch.Tensor] = None,
                  accumulate: bool = False):
    """tensor-core GEMM that also returns the per-128-row-tile column statistics of its output (for norm_rows_pre).
    accumulate: `out` (contiguous [M, N]) is read and rewritten, out += a @ w.T + bias; the statistics describe the sum."""
    a, lda = _rows(a, "a")
    M, K = a.shape
    N = w.shape[0]
    w, ldw, eng = _x3_weights(w, _eng(), const_w)
    if accumulate:
        if out is None or rowdiv is not None or tuple(out.shape) != (M, N) or not out.is_contiguous() or out.dtype != torch.float32:
            raise RuntimeError("gemm_colstats(accumulate): needs a contiguous float32 [M, N] `out` and no rowdiv")
    else:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    stats = torch.empty((M // 128, N, 2), dtype=torch.float32, device=a.device)
    _meta(2.0 * M * N * K, 4.0 * (M * K + N * K + M * N * (2 if accumulate else 1)), f"{M}x{N}x{K}" + (" acc" if accumulate else ""))
    if accumulate:
        _call("cofi_gemm_colstats_acc", _p(a), lda, _p(w), ldw, _p(out), N, M, N, K, _p(bias), eng, _p(stats), _st())
    else:
        _call("cofi_gemm_colstats", _p(a), lda, _p(w), ldw, _p(out), N, M, N, K, _p(bias), _p(rowdiv), eng, _p(stats), _st())
    return out, stats


def gemm_f16_colstats(a_half, w_half, bias=None, rowdiv=None):
    a_half, w_half = a_half.contiguous(), w_half.contiguous()
    M, K = a_half.shape
    N = w_half.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a_half.device)
    stats = torch.empty((M // 128, N, 2), dtype=torch.float32, device=a_half.device)
    _meta(2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N, f"{M}x{N}x{K} f16")
    _call("cofi_gemm_f16_colstats", _p(a_half), K, _p(w_half), K, _p(out), N, M, N, K, _p(bias), _p(rowdiv), _p(stats), _st())
    return out, stats


def colstats_ok(rows: int, frames: int, n: int) -> bool:
    """GEMM-epilogue statistics apply on the tensor-core engines when every frame is a whole number of 128-row tiles."""
    return _eng() in (ENGINE_TF32, ENGINE_TF32X3) and rows % frames == 0 and (rows // frames) % 128 == 0 and n >= 16 and n % 4 == 0


def norm_rows_pre(x, stats, frames: int, groups: int, gamma, beta, eps: float = 1e-5, residual=None, act: int = ACT_NONE,
                  return_mr: bool = False):
    """GroupNorm whose statistics come from the producing GEMM's epilogue (`stats` of gemm_colstats).  return_mr: also the
    (mean, rstd) per (frame, group) [frames*groups, 2] the finalize kernel derived -- what cofi_norm_rows_bwd consumes."""
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ws = torch.empty((frames * groups * 2,), dtype=torch.float32, device=x.device)
    ldr = 0
    if residual is not None:
        residual, ldr = _rows(residual, "residual")
    _meta(4.0 * rows * C, 4.0 * rows * C * (2 + (residual is not None)))
    _call("cofi_norm_rows_pre", _p(x), ldx, rows // frames, C, frames, groups, _p(gamma), _p(beta), float(eps), _p(residual), ldr,
          act, _p(y), C, _p(stats), _p(ws), _st())
    return (y, ws.view(frames * groups, 2)) if return_mr else y


def gemm_ln(a, w, gamma, beta, eps: float = 1e-5, bias=None, act: int = ACT_NONE, residual=None,
            engine: Optional[int] = None, out: Optional[torch.Tensor] = None):
    """act(LayerNorm(a @ w.T + bias)) + residual, one kernel on the tensor-core engines when N <= 128.
    `out`: optional contiguous [M, N] destination (e.g. a row slice of a larger buffer)."""
    a, lda = _rows(a, "a")
    M, K = a.shape
    N = w.shape[0]
    fused = N <= 128 and N % 32 == 0   # unfused shapes (GEMM, then LayerNorm) keep the in-kernel split path
    w, ldw, eng = _x3_weights(w, _eng() if engine is None else engine, None if fused else False)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    elif tuple(out.shape) != (M, N) or not out.is_contiguous() or out.dtype != torch.float32:
        raise RuntimeError(f"gemm_ln: out must be a contiguous float32 [{M}, {N}] tensor")
    ldr = 0
    if residual is not None:
        residual, ldr = _rows(residual, "residual")
    _meta(2.0 * M * N * K, 4.0 * (M * K + N * K + M * N * (2 if residual is not None else 1)), f"{M}x{N}x{K} ln")
    _call("cofi_gemm_ln", _p(a), lda, _p(w), ldw, _p(out), N, M, N, K, _p(bias), _p(gamma), _p(beta), float(eps), act,
          _p(residual), ldr, eng, _st())
    return out


def conv2d_nhwc(x, w_packed, kh: int, kw: int, stride: int, pad: int, scale=None, shift=None, residual=None,
                act: int = ACT_NONE, engine: Optional[int] = None):
    """x [B,H,W,Cin] fp32 contiguous, w_packed [Cout, kh*kw*Cin]."""
    _f32(x, "x")
    x = x.contiguous()
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    if residual is not None:
        residual = residual.contiguous()
    eng = _eng() if engine is None else engine
    if eng == ENGINE_TF32X3 and Cout >= 16 and stride in (1, 2) and Wo % 64 == 0 and Ho % 2 == 0:
        w_packed, eng = split_tf32(w_packed), ENGINE_TF32X3S   # packed conv weights are constants (imagenet.py caches them)
    _meta(2.0 * B * Ho * Wo * Cout * kh * kw * Cin, 4.0 * (x.numel() + w_packed.numel() + y.numel()),
          f"{B}x{H}x{W}x{Cin}->{Cout} k{kh}s{stride}")
    _call("cofi_conv2d_nhwc", _p(x), B, H, W, Cin, _p(w_packed), Cout, kh, kw, stride, pad, _p(scale), _p(shift),
                               _p(residual), act, _p(y), eng, _st())
    return y


# ------------------------------------------------------------------------------------------ normalisations
def norm_rows(x, frames: int, groups: int, gamma=None, beta=None, eps: float = 1e-5, residual=None,
              act: int = ACT_NONE, want_stats: bool = False):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    R = rows // frames
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ws = _ws(_lib.cofi_norm_rows_workspace(frames, C), x.device)
    mean = var = None
    if want_stats:
        mean = torch.empty((frames, groups), dtype=torch.float32, device=x.device)
        var = torch.empty((frames, groups), dtype=torch.float32, device=x.device)
    ldr = 0
    if residual is not None:
        residual, ldr = _rows(residual, "residual")
    _meta(8.0 * rows * C, 4.0 * rows * C * (3 + (residual is not None)))
    _call("cofi_norm_rows", _p(x), ldx, R, C, frames, groups, _p(gamma), _p(beta), float(eps), _p(residual), ldr, act,
                             _p(y), C, _p(ws), _p(mean), _p(var), _st())
    if want_stats:
        return y, mean, var
    return y


def affine_rows(x, scale=None, shift=None, residual=None, act: int = ACT_NONE):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ldr = 0
    if residual is not None:
        residual, ldr = _rows(residual, "residual")
    _call("cofi_affine_rows", _p(x), ldx, rows, C, _p(scale), _p(shift), _p(residual), ldr, act, _p(y), C, _st())
    return y


def layer_norm_rows(x, gamma, beta, eps: float = 1e-5, act: int = ACT_NONE, residual=None):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ldr = 0
    if residual is not None:
        residual, ldr = _rows(residual, "residual")
    _call("cofi_layer_norm_rows", _p(x), ldx, rows, C, _p(gamma), _p(beta), float(eps), act, _p(residual), ldr,
                                   _p(y), C, _st())
    return y


def l2norm_rows(x, add=None, out: Optional[torch.Tensor] = None):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    if out is None:
        out = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ldadd = 0
    if add is not None:
        add, ldadd = _rows(add, "add")
    _call("cofi_l2norm_rows", _p(x), ldx, rows, C, _p(add), ldadd, _p(out), out.stride(0), _st())
    return out


def colnorm_rows(x, frames: int = 1):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    ws = _ws(_lib.cofi_colnorm_workspace(frames, C), x.device)
    _call("cofi_colnorm_rows", _p(x), ldx, rows // frames, C, frames, _p(ws), _p(y), C, _st())
    return y


# ------------------------------------------------------------------------------------------ image helpers
def nchw_to_nhwc(x, cpad: Optional[int] = None):
    _f32(x, "x")
    x = x.contiguous()
    B, C, H, W = x.shape
    cpad = C if cpad is None else cpad
    y = torch.empty((B, H, W, cpad), dtype=torch.float32, device=x.device)
    _call("cofi_nchw_to_nhwc", _p(x), B, C, H, W, cpad, _p(y), _st())
    return y


def nhwc_to_nchw(x):
    _f32(x, "x")
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    _call("cofi_nhwc_to_nchw", _p(x), B, H, W, C, _p(y), _st())
    return y


def maxpool2d_3x3s2_nhwc(x):
    _f32(x, "x")
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=torch.float32, device=x.device)
    _call("cofi_maxpool2d_3x3s2_nhwc", _p(x), B, H, W, C, _p(y), _st())
    return y


def upsample2x_cat_nhwc(x1, x2):
    _f32(x1, "x1")
    _f32(x2, "x2")
    x1, x2 = x1.contiguous(), x2.contiguous()
    B, H, W, C1 = x1.shape
    C2 = x2.shape[3]
    if tuple(x2.shape[:3]) != (B, 2 * H, 2 * W):
        raise RuntimeError(f"upsample2x_cat_nhwc: shape mismatch {tuple(x1.shape)} vs {tuple(x2.shape)}")
    y = torch.empty((B, 2 * H, 2 * W, C1 + C2), dtype=torch.float32, device=x1.device)
    _call("cofi_upsample2x_cat_nhwc", _p(x1), B, H, W, C1, _p(x2), C2, _p(y), _st())
    return y


# ------------------------------------------------------------------------------------------ transformer
def posenc_sine(coords, d_model: int, dim_t: torch.Tensor):
    coords = _f32(coords, "coords").contiguous()
    rows, n_dim = coords.shape
    out = torch.empty((rows, d_model), dtype=torch.float32, device=coords.device)
    _call("cofi_posenc_sine", _p(coords), rows, n_dim, d_model, _p(dim_t), _p(out), _st())
    return out


def attention(q, k, v, frames: int, heads: int, scale: float, engine: Optional[int] = None):
    q, k, v = _f32(q, "q").contiguous(), _f32(k, "k").contiguous(), _f32(v, "v").contiguous()
    L, S = q.shape[0] // frames, k.shape[0] // frames
    D = q.shape[1] // heads
    out = torch.empty_like(q)
    _meta(4.0 * frames * L * S * heads * D, 4.0 * (2 * q.numel() + 2 * k.numel()))
    _call("cofi_attention", _p(q), _p(k), _p(v), L, S, frames, heads, D, float(scale), _p(out),
                             _eng() if engine is None else engine, _st())
    return out


def attention_vt(q, k, vt, frames: int, heads: int, scale: float):
    """tcgen05 flash attention; vt = V^T [heads*D, frames*S] (from gemm(W_v, source))."""
    q, k, vt = _f32(q, "q").contiguous(), _f32(k, "k").contiguous(), _f32(vt, "vt").contiguous()
    L, S = q.shape[0] // frames, k.shape[0] // frames
    D = q.shape[1] // heads
    out = torch.empty_like(q)
    _meta(4.0 * frames * L * S * heads * D, 4.0 * (2 * q.numel() + 2 * k.numel()))
    _call("cofi_attention_vt", _p(q), _p(k), _p(vt), L, S, frames, heads, D, float(scale), _p(out), _st())
    return out


def engine_id() -> int:
    """The engine in force for the current group (policy override or the global engine)."""
    return _eng()


# ------------------------------------------------------------------------------------------ matching
def sim_argmin(pt, px, frames: int = 1, engine: Optional[int] = None, pt_h=None, px_h=None, stats=None):
    """Fused similarity + arg-min (reference model/network.py:174-179): (argmin pixel per point row, min distance).
    engine None (every product path): the EXACT result -- on the tensor cores when C is 64 or 128 (tcgen05 fp16 candidate
    pass + exact fp32 re-rank, bit-identical to the fp32 engine), on the fp32 SIMT kernel otherwise.  engine=ENGINE_FP32
    forces the SIMT kernel; ENGINE_TF32 / TF32X3 select the approximate tcgen05 tf32 kernel (tests, tools)."""
    pt, ldpt = _rows(pt, "pt")
    px, ldpx = _rows(px, "px")
    C = pt.shape[1]
    if engine is None:
        if C in (64, 128) and ldpx % 4 == 0 and px.data_ptr() % 16 == 0:
            return sim_argmin_exact(pt, px, frames, pt_h, px_h, stats=stats)
        engine = ENGINE_FP32
    Npt, Npx = pt.shape[0] // frames, px.shape[0] // frames
    idx = torch.empty((pt.shape[0],), dtype=torch.int64, device=pt.device)
    val = torch.empty((pt.shape[0],), dtype=torch.float32, device=pt.device)
    _meta(2.0 * frames * Npt * Npx * pt.shape[1], 4.0 * (pt.numel() + px.numel()) + 12.0 * pt.shape[0])
    _call("cofi_sim_argmin", _p(pt), ldpt, _p(px), ldpx, Npt, Npx, pt.shape[1], frames, _p(idx), _p(val), engine, _st())
    return idx, val


def cast_f16_bound(x, bound2_slot: torch.Tensor):
    """fp16 copy of the rows of x; bound2_slot (1-element fp32 view, zeroed by the caller) receives max |row|^2."""
    x, ldx = _rows(x, "x")
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _call("cofi_cast_f16_bound", _p(x), ldx, x.shape[0], x.shape[1], _p(y), x.shape[1], _p(bound2_slot), _st())
    return y


def sim_argmin_exact(pt, px, frames: int = 1, pt_h=None, px_h=None, stats=None):
    """tcgen05 candidate pass over fp16 copies + exact fp32 re-rank (cofi_sim_argmin_exact).  pt_h / px_h: fp16 copies of
    the rows when the producer already emitted them (l2norm_rows_f16: unit-norm rows, fixed margin); otherwise they are
    made here together with the norm bound that scales the margin.  stats: optional int32[2] device tensor (accumulates
    the number of re-ranked candidates and of rows that fell back to the full exact scan)."""
    pt, ldpt = _rows(pt, "pt")
    px, ldpx = _rows(px, "px")
    Npt, Npx = pt.shape[0] // frames, px.shape[0] // frames
    C = pt.shape[1]
    bound2 = None
    if pt_h is None or px_h is None:
        bound2 = torch.zeros((2,), dtype=torch.float32, device=pt.device)
        pt_h = cast_f16_bound(pt, bound2[0:1])
        px_h = cast_f16_bound(px, bound2[1:2])
    if pt_h.dtype != torch.float16 or px_h.dtype != torch.float16 or not pt_h.is_contiguous() or not px_h.is_contiguous():
        raise RuntimeError("sim_argmin_exact: contiguous CUDA fp16 copies expected")
    idx = torch.empty((pt.shape[0],), dtype=torch.int64, device=pt.device)
    val = torch.empty((pt.shape[0],), dtype=torch.float32, device=pt.device)
    ws = _ws(_lib.cofi_sim_argmin_exact_workspace(Npt, Npx, frames), pt.device)
    _meta(2.0 * frames * Npt * Npx * C, 2.0 * (pt.numel() + px.numel()) + 12.0 * pt.shape[0])
    _call("cofi_sim_argmin_exact", _p(pt), ldpt, _p(px), ldpx, _p(pt_h), C, _p(px_h), C, Npt, Npx, C, frames, _p(bound2),
          _p(idx), _p(val), _p(ws), _p(stats), _st())
    return idx, val


def l2norm_rows_f16(x):
    """-> (F.normalize(x) fp32, its fp16 copy): the producer of cofi_sim_argmin_exact's operands."""
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    y = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    yh = torch.empty((rows, C), dtype=torch.float16, device=x.device)
    _call("cofi_l2norm_rows_f16", _p(x), ldx, rows, C, _p(y), C, _p(yh), C, _st())
    return y, yh


def cast_f16(x):
    x, ldx = _rows(x, "x")
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _call("cofi_cast_f16", _p(x), ldx, x.shape[0], x.shape[1], _p(y), x.shape[1], _st())
    return y


def sim_argmin_f16(pt_h, px_h, frames: int = 1):
    """fp16 feature rows (cast_f16 of the L2-normalised features) -> (argmin index, min distance)."""
    if pt_h.dtype != torch.float16 or px_h.dtype != torch.float16 or not pt_h.is_cuda:
        raise RuntimeError("sim_argmin_f16: CUDA fp16 operands expected")
    pt_h, px_h = pt_h.contiguous(), px_h.contiguous()
    Npt, Npx = pt_h.shape[0] // frames, px_h.shape[0] // frames
    C = pt_h.shape[1]
    idx = torch.empty((pt_h.shape[0],), dtype=torch.int64, device=pt_h.device)
    val = torch.empty((pt_h.shape[0],), dtype=torch.float32, device=pt_h.device)
    ctas = ((Npt + 255) // 256) * frames
    nsplit = max(1, min((Npx + 127) // 128, (2 * 148 + ctas - 1) // ctas)) if ctas < 148 else 1
    ws_i = ws_v = None
    if nsplit > 1:
        ws_i = torch.empty((nsplit, pt_h.shape[0]), dtype=torch.int64, device=pt_h.device)
        ws_v = torch.empty((nsplit, pt_h.shape[0]), dtype=torch.float32, device=pt_h.device)
    _meta(2.0 * frames * Npt * Npx * C, 2.0 * (pt_h.numel() + px_h.numel()) + 12.0 * pt_h.shape[0])
    _call("cofi_sim_argmin_f16", _p(pt_h), C, _p(px_h), C, Npt, Npx, C, frames, _p(idx), _p(val), nsplit, _p(ws_i), _p(ws_v),
          _st())
    return idx, val


def select_matches(score, best_idx, frames: int, grid_h: int, grid_w: int, thresholds: torch.Tensor, min_count: int = 4,
                   xy_scale: float = 1.0, xmax: int = 62, ymax: int = 18):
    """xmax / ymax default to the reference's literals (model/network.py:184)."""
    score = _f32(score, "score").contiguous().view(-1)
    best_idx = _i64(best_idx, "best_idx")
    Npt = score.numel() // frames
    cnt = torch.empty((frames, 2), dtype=torch.int32, device=score.device)
    oidx = torch.empty((frames, Npt), dtype=torch.int64, device=score.device)
    oxy = torch.empty((frames, 2, Npt), dtype=torch.float32, device=score.device)
    _call("cofi_select_matches", _p(score), _p(best_idx), Npt, frames, grid_h, grid_w, int(xmax), int(ymax), _p(thresholds),
                                  thresholds.numel(), min_count, float(xy_scale), _p(cnt), _p(oidx), _p(oxy), _st())
    return cnt, oidx, oxy


KNN_DIRECT, KNN_EXPANDED, KNN_NOCULL = 0, 1, 0x100
KNN_NOFAST = 0x200   # test/debug: iterative sort/merge selection for every query (no threshold-selection fast path)


def knn_pyramid(points, frames: int = 1, k: int = 128, mode: int = KNN_DIRECT, want=("neighbors", "subsampling", "upsampling"),
                workspace: Optional[torch.Tensor] = None, k_up: Optional[int] = None, out: Optional[dict] = None):
    """All KNN-k tables of a point pyramid in four launches (two sorts, two query passes; reference model/kpconv/preprocess_data.py:75-99,172-190).
    points: list of [frames*n_l, 3] fp32 CUDA tensors -> dict(neighbors, subsampling, upsampling) of int64 tables with
    frame-local indices, rows ascending in (distance, index).  k_up: columns of the upsampling tables (default k; the
    model only reads column 0, reference model/kpconv/functional.py:20, so the engine asks for 1)."""
    k_up = k if k_up is None else k_up
    pts = [_f32(p, "points").contiguous() for p in points]
    L, dev = len(pts), pts[0].device
    n = [p.shape[0] // frames for p in pts]
    for p, nl in zip(pts, n):
        if p.dim() != 2 or p.shape[1] != 3 or nl * frames != p.shape[0] or nl < 1:
            raise RuntimeError(f"knn_pyramid: every level must be [frames*n, 3], got {tuple(p.shape)} for frames={frames}")
    n_arr = (ctypes.c_int64 * L)(*n)
    if workspace is None:
        workspace = _ws(_lib.cofi_knn_pyramid_workspace(n_arr, L, frames), dev)
    if out is None:  # `out`: tables of a previous call with the same arguments, overwritten in place (static engine buffers)
        out = {"neighbors": [], "subsampling": [], "upsampling": []}
        if "neighbors" in want:
            out["neighbors"] = [torch.empty((frames * n[l], k), dtype=torch.int64, device=dev) for l in range(L)]
        if "subsampling" in want:
            out["subsampling"] = [torch.empty((frames * n[l + 1], k), dtype=torch.int64, device=dev) for l in range(L - 1)]
        if "upsampling" in want:
            out["upsampling"] = [torch.empty((frames * n[l], k_up), dtype=torch.int64, device=dev) for l in range(L - 1)]

    def arr(ts):
        return (ctypes.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts]) if ts else None
    pairs = sum(n[l] * n[l] for l in range(L)) * ("neighbors" in want) + \
        sum(2 * n[l] * n[l + 1] for l in range(L - 1)) * (("subsampling" in want) + ("upsampling" in want)) / 2
    nb = sum(t.numel() * 8 for v in out.values() for t in v) + sum(p.numel() * 4 for p in pts)
    _meta(8.0 * pairs * frames, nb)
    _call("cofi_knn_pyramid", arr(pts), n_arr, L, frames, k, k_up, mode, arr(out["neighbors"]), arr(out["subsampling"]),
          arr(out["upsampling"]), _p(workspace), _st())
    return out


def half_sample_pyramid(points0, frames: int = 1, levels: int = 5, seed: int = 0, want_index: bool = False):
    """Device-side random half-sampling (reference model/kpconv/preprocess_data.py:52-68, WITH replacement; counter-based
    Philox draw, see cofi_half_sample_pyramid).  points0 [frames*n0, 3] -> list of `levels` clouds [frames*(n0>>l), 3]
    (entry 0 is points0 itself) and, with want_index, the frame-local level-0 row every sampled row was copied from."""
    points0 = _f32(points0, "points0").contiguous()
    n0 = points0.shape[0] // frames
    if points0.dim() != 2 or points0.shape[1] != 3 or n0 * frames != points0.shape[0] or (n0 >> (levels - 1)) < 1:
        raise RuntimeError(f"half_sample_pyramid: expected [frames*n0, 3] with n0 >= 2^(levels-1), got {tuple(points0.shape)}")
    dev = points0.device
    outs = [points0] + [torch.empty((frames * (n0 >> l), 3), dtype=torch.float32, device=dev) for l in range(1, levels)]
    idxs = [None] + [torch.empty((frames * (n0 >> l),), dtype=torch.int64, device=dev) for l in range(1, levels)] \
        if want_index else None
    po = (ctypes.c_void_p * levels)(*[t.data_ptr() for t in outs])
    pi = (ctypes.c_void_p * levels)(*[0 if t is None else t.data_ptr() for t in idxs]) if want_index else None
    _call("cofi_half_sample_pyramid", _p(points0), n0, frames, levels, int(seed) & 0xFFFFFFFFFFFFFFFF, po, pi, _st())
    return (outs, idxs) if want_index else outs


def knn_table(src, qry, frames: int = 1, k: int = 128, mode: int = KNN_DIRECT):
    """out[frames*nq, k]: the k nearest rows of src for every row of qry (`knn(nodes, points, k)`,
    reference model/kpconv/preprocess_data.py:131-143)."""
    src, qry = _f32(src, "src").contiguous(), _f32(qry, "qry").contiguous()
    ns, nq = src.shape[0] // frames, qry.shape[0] // frames
    if src.dim() != 2 or qry.dim() != 2 or src.shape[1] != 3 or qry.shape[1] != 3 or ns < 1 or nq < 1:
        raise RuntimeError(f"knn_table: expected [frames*n, 3] clouds, got {tuple(src.shape)} and {tuple(qry.shape)}")
    out = torch.empty((frames * nq, k), dtype=torch.int64, device=src.device)
    ws = _ws(_lib.cofi_knn_table_workspace(ns, nq, frames), src.device)
    _meta(8.0 * ns * nq * frames, out.numel() * 8 + (src.numel() + qry.numel()) * 4)
    _call("cofi_knn_table", _p(src), ns, _p(qry), nq, frames, k, mode, _p(out), _p(ws), _st())
    return out


def nn_argmin(points, nodes):
    points, nodes = _f32(points, "points").contiguous(), _f32(nodes, "nodes").contiguous()
    n = points.shape[0]
    idx = torch.empty((n,), dtype=torch.int64, device=points.device)
    _call("cofi_nn_argmin", _p(points), n, _p(nodes), nodes.shape[0], _p(idx), _st())
    return idx


def nn_argmin_batched(points, nodes, frames: int):
    """frame f: nearest row of nodes[f] for every row of points[f] (frame-local indices)."""
    points, nodes = _f32(points, "points").contiguous(), _f32(nodes, "nodes").contiguous()
    n, M = points.shape[0] // frames, nodes.shape[0] // frames
    idx = torch.empty((frames * n,), dtype=torch.int64, device=points.device)
    _call("cofi_nn_argmin_batched", _p(points), n, _p(nodes), M, frames, _p(idx), _st())
    return idx


def extract_patch_batched(map_nhwc, centers, err_flag: Optional[torch.Tensor] = None):
    """map [B,H,W,C]; centers [B,2,n] fp32 -> [B,n,C,4,4]."""
    _f32(map_nhwc, "map")
    map_nhwc = map_nhwc.contiguous()
    centers = _f32(centers, "centers").contiguous()
    B, H, W, C = map_nhwc.shape
    n = centers.shape[2]
    out = torch.empty((B, n, C, 4, 4), dtype=torch.float32, device=map_nhwc.device)
    _call("cofi_extract_patch_batched", _p(map_nhwc), H, W, C, B, _p(centers), n, _p(out), _p(err_flag), _st())
    return out


def extract_patch(map_nhwc, b: int, centers, err_flag: Optional[torch.Tensor] = None):
    """map [B,H,W,C]; centers [2,n] fp32 (x row 0, y row 1) -> [n,C,4,4]."""
    _f32(map_nhwc, "map")
    map_nhwc = map_nhwc.contiguous()
    centers = _f32(centers, "centers").contiguous()
    _, H, W, C = map_nhwc.shape
    n = centers.shape[1]
    out = torch.empty((n, C, 4, 4), dtype=torch.float32, device=map_nhwc.device)
    _call("cofi_extract_patch", _p(map_nhwc), H, W, C, b, _p(centers), n, _p(out), _p(err_flag), _st())
    return out


def fine_match(patch, pc):
    """patch [n,C,16] (or [n,C,4,4]), pc [n,C] -> argmax index [n] int64 (evaluation/eval_all.py:99-102)."""
    patch = _f32(patch, "patch").contiguous()
    pc = _f32(pc, "pc").contiguous()
    n, C = pc.shape
    idx = torch.empty((n,), dtype=torch.int64, device=pc.device)
    _call("cofi_fine_match", _p(patch), _p(pc), n, C, _p(idx), _st())
    return idx


# ============================================================================================ training (backward.cu)
def transpose2d(x):
    """[R, C] -> [C, R] (the tiled NHWC->NCHW transpose with B = H = 1)."""
    x, _ = _rows(x.contiguous(), "x")
    R, C = x.shape
    y = torch.empty((C, R), dtype=torch.float32, device=x.device)
    _call("cofi_nhwc_to_nchw", _p(x), 1, 1, R, C, _p(y), _st())
    return y


def act_bwd(dy, y, act: int):
    dy, y = dy.contiguous(), y.contiguous()
    dx = torch.empty_like(dy)
    _call("cofi_act_bwd", _p(dy), _p(y), dy.numel(), act, _p(dx), _st())
    return dx


def rowscale(x, rowdiv):
    x = x.contiguous()
    y = torch.empty_like(x)
    _call("cofi_rowscale", _p(x), x.shape[0], x.shape[1], _p(rowdiv), _p(y), _st())
    return y


def colsum(x):
    x, ldx = _rows(x, "x")
    out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    ws = _ws(_lib.cofi_colsum_workspace(x.shape[1]), x.device)
    _call("cofi_colsum", _p(x), ldx, x.shape[0], x.shape[1], _p(out), 0, _p(ws), _st())
    return out


def gemm_tn(a, b):
    """a [R, Mo], b [R, No] -> a^T b [Mo, No] (tcgen05 MN-major under the tf32 engine, SIMT fp32 otherwise; deterministic
    split over R)."""
    a, lda = _rows(a, "a")
    b, ldb = _rows(b, "b")
    R, Mo = a.shape
    No = b.shape[1]
    out = torch.empty((Mo, No), dtype=torch.float32, device=a.device)
    ws = _ws(_lib.cofi_gemm_tn_workspace(R, Mo, No), a.device)
    _meta(2.0 * R * Mo * No, 4.0 * (R * Mo + R * No + Mo * No))
    _call("cofi_gemm_tn", _p(a), lda, _p(b), ldb, _p(out), R, Mo, No, 0, _eng(), _p(ws), _st())
    return out


def norm_rows_stats(x, frames: int, groups: int, eps: float):
    x, ldx = _rows(x, "x")
    rows, C = x.shape
    ws = _ws(_lib.cofi_norm_rows_workspace(frames, C), x.device)
    mr = torch.empty((frames * groups, 2), dtype=torch.float32, device=x.device)
    _call("cofi_norm_rows_stats", _p(x), ldx, rows // frames, C, frames, groups, float(eps), _p(ws), _p(mr), _st())
    return mr


def norm_rows_bwd(x, dy, y, mean_rstd, frames: int, groups: int, gamma, act: int, want_res: bool):
    x, dy, y = x.contiguous(), dy.contiguous(), y.contiguous()
    rows, C = x.shape
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_res else None
    dgamma = torch.empty((C,), dtype=torch.float32, device=x.device) if gamma is not None else None
    dbeta = torch.empty((C,), dtype=torch.float32, device=x.device) if gamma is not None else None
    ws = _ws(_lib.cofi_norm_rows_bwd_workspace(frames, C), x.device)
    _call("cofi_norm_rows_bwd", _p(x), _p(dy), _p(y), rows // frames, C, frames, groups, _p(mean_rstd), _p(gamma), act, _p(dx),
          _p(dres), _p(dgamma), _p(dbeta), 0, _p(ws), _st())
    return dx, dres, dgamma, dbeta


def layer_norm_bwd(x, dy, gamma, beta, eps: float, act: int):
    x, dy = x.contiguous(), dy.contiguous()
    dx, t1, t2 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    _call("cofi_layer_norm_bwd", _p(x), _p(dy), x.shape[0], x.shape[1], _p(gamma), _p(beta), float(eps), act, _p(dx), _p(t1),
          _p(t2), _st())
    return dx, colsum(t1), colsum(t2)


def l2norm_bwd(x, dy):
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    _call("cofi_l2norm_bwd", _p(x), _p(dy), x.shape[0], x.shape[1], _p(dx), _st())
    return dx


def colnorm_bwd(x, dy, frames: int):
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    ws = _ws(_lib.cofi_colnorm_bwd_workspace(frames, x.shape[1]), x.device)
    _call("cofi_colnorm_bwd", _p(x), _p(dy), x.shape[0] // frames, x.shape[1], frames, _p(dx), _p(ws), _st())
    return dx


def scatter_add_rows(dy, idx, idx_stride: int, frames: int, rows_src: int):
    dy, ldy = _rows(dy, "dy")
    C = dy.shape[1]
    dx = torch.zeros((rows_src, C), dtype=torch.float32, device=dy.device)
    _call("cofi_scatter_add_rows", _p(dy), ldy, C, _p(idx), idx_stride, dy.shape[0] // frames, rows_src // frames, frames, _p(dx),
          _st())
    return dx


def maxpool_rows_bwd(x, nbr, dy, frames: int):
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.zeros_like(x)
    _call("cofi_maxpool_rows_bwd", _p(x), x.shape[1], _p(nbr), nbr.shape[1], nbr.shape[0] // frames, x.shape[0] // frames, frames,
          _p(dy), _p(dx), _st())
    return dx


def kpconv_aggregate_bwd(dagg, C: int, s_packed, q_points, nbr, kernel_points, sigma: float, frames: int, kp_reach: float):
    dagg = dagg.contiguous()
    rows_src = s_packed.shape[0]
    dfeats = torch.zeros((rows_src, C), dtype=torch.float32, device=dagg.device)
    _call("cofi_kpconv_aggregate_bwd", _p(dagg), C, _p(s_packed), _p(q_points), _p(nbr), nbr.shape[1], nbr.shape[0] // frames,
          rows_src // frames, frames, _p(kernel_points), kernel_points.shape[0], float(sigma), float(kp_reach), _p(dfeats), _st())
    return dfeats


def upsample2x_cat_bwd(dy, C1: int):
    dy = dy.contiguous()
    B, Ho, Wo, Ct = dy.shape
    dx1 = torch.zeros((B, Ho // 2, Wo // 2, C1), dtype=torch.float32, device=dy.device)
    dx2 = torch.empty((B, Ho, Wo, Ct - C1), dtype=torch.float32, device=dy.device)
    _call("cofi_upsample2x_cat_bwd", _p(dy), B, Ho // 2, Wo // 2, C1, Ct - C1, _p(dx1), _p(dx2), _st())
    return dx1, dx2


def maxpool2d_3x3s2_bwd(x, dy):
    x, dy = x.contiguous(), dy.contiguous()
    B, H, W, C = x.shape
    dx = torch.zeros_like(x)
    _call("cofi_maxpool2d_3x3s2_bwd", _p(x), _p(dy), B, H, W, C, _p(dx), _st())
    return dx


def dilate2_nhwc(x):
    x = x.contiguous()
    B, H, W, C = x.shape
    y = torch.empty((B, 2 * H, 2 * W, C), dtype=torch.float32, device=x.device)
    _call("cofi_dilate2_nhwc", _p(x), B, H, W, C, _p(y), _st())
    return y


def extract_patch_bwd(dpatch, map_shape, b: int, centers):
    dpatch = dpatch.contiguous()
    B, H, W, C = map_shape
    dmap = torch.zeros(map_shape, dtype=torch.float32, device=dpatch.device)
    _call("cofi_extract_patch_bwd", _p(dpatch), H, W, C, b, _p(centers), centers.shape[1], _p(dmap), _st())
    return dmap


def extract_patch_batched_bwd(dpatch, map_shape, centers):
    dpatch, centers = dpatch.contiguous(), centers.contiguous()
    B, H, W, C = map_shape
    dmap = torch.zeros(map_shape, dtype=torch.float32, device=dpatch.device)
    _call("cofi_extract_patch_batched_bwd", _p(dpatch), H, W, C, B, _p(centers), centers.shape[2], _p(dmap), _st())
    return dmap


def conv2d_wgrad_nhwc(x, dy, kh: int, kw: int, stride: int, pad: int):
    x, dy = x.contiguous(), dy.contiguous()
    B, H, W, Cin = x.shape
    _, Ho, Wo, Cout = dy.shape
    dw = torch.empty((Cout, kh * kw * Cin), dtype=torch.float32, device=x.device)
    ws = _ws(_lib.cofi_conv2d_wgrad_workspace(B, Ho, Wo, Cout, kh, kw, Cin), x.device)
    _meta(2.0 * B * Ho * Wo * Cout * kh * kw * Cin, 4.0 * (x.numel() + dy.numel() + dw.numel()))
    _call("cofi_conv2d_wgrad_nhwc", _p(x), B, H, W, Cin, _p(dy), Cout, kh, kw, stride, pad, _p(dw), 0, _eng(), _p(ws), _st())
    return dw


def attention_fwd_lse(q, k, v, frames: int, heads: int, scale: float):
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    L, S = q.shape[0] // frames, k.shape[0] // frames
    out = torch.empty_like(q)
    lse = torch.empty((q.shape[0], heads), dtype=torch.float32, device=q.device)
    _meta(4.0 * frames * L * S * q.shape[1], 4.0 * (2 * q.numel() + 2 * k.numel()))
    if _eng() == ENGINE_TF32 and (frames * S) % 4 == 0:  # tcgen05 flash attention (K-major V^T operand)
        vt = transpose2d(v)
        _meta(4.0 * frames * L * S * q.shape[1], 4.0 * (2 * q.numel() + 2 * k.numel()))
        _call("cofi_attention_vt_lse", _p(q), _p(k), _p(vt), L, S, frames, heads, q.shape[1] // heads, float(scale), _p(out),
              _p(lse), _st())
        return out, lse
    _call("cofi_attention_fwd_lse", _p(q), _p(k), _p(v), L, S, frames, heads, q.shape[1] // heads, float(scale), _p(out), _p(lse),
          _st())
    return out, lse


def attention_bwd(q, k, v, out, dout, lse, frames: int, heads: int, scale: float):
    dout = dout.contiguous()
    L, S = q.shape[0] // frames, k.shape[0] // frames
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dsum = torch.empty((q.shape[0], heads), dtype=torch.float32, device=q.device)
    _meta(10.0 * frames * L * S * q.shape[1], 4.0 * (4 * q.numel() + 4 * k.numel()))
    D = q.shape[1] // heads
    if _eng() == ENGINE_TF32 and D == 32 and L % 4 == 0 and S % 4 == 0:
        # tcgen05 backward (tf32 operands, fp32 accumulate); Q^T, dO^T, K^T copies live in the workspace
        ws = _ws(_lib.cofi_attention_bwd_tc_workspace(L, S, frames, heads, D), q.device)
        _call("cofi_attention_bwd_tc", _p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), L, S, frames, heads, D, float(scale),
              _p(dq), _p(dk), _p(dv), _p(dsum), _p(ws), _st())
        return dq, dk, dv
    _call("cofi_attention_bwd", _p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), L, S, frames, heads, q.shape[1] // heads,
          float(scale), _p(dq), _p(dk), _p(dv), _p(dsum), _st())
    return dq, dk, dv


# ------------------------------------------------------------------------------------------ pose step
def pnp_ransac(image_points, object_points, cam, count: Optional[torch.Tensor] = None, iterations: int = 10000,
               threshold: float = 8.0, seed: int = 0):
    """Batched P3P-RANSAC (cofi_pnp_ransac).  image_points [B,n,2], object_points [B,n,3], cam [B,4] (fx, fy, cx, cy) fp32;
    count: optional int32 [B] or [B,s] (column 0 = real rows per frame).  -> (inlier count [B] i32, winning hypothesis [B]
    i32, pose [B,12] f64 (R row-major, t), inlier mask [B,n] u8)."""
    image_points, object_points = _f32(image_points, "image_points").contiguous(), _f32(object_points, "object_points").contiguous()
    cam = _f32(cam, "cam").contiguous()
    B, n = image_points.shape[0], image_points.shape[1]
    dev = image_points.device
    stride = 0
    if count is not None:
        if count.dtype != torch.int32 or not count.is_cuda or not count.is_contiguous():
            raise RuntimeError("pnp_ransac: count must be a contiguous CUDA int32 tensor")
        stride = 1 if count.dim() == 1 else count.shape[1]
    o_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    o_hyp = torch.empty((B,), dtype=torch.int32, device=dev)
    o_pose = torch.empty((B, 12), dtype=torch.float64, device=dev)
    o_inl = torch.empty((B, n), dtype=torch.uint8, device=dev)
    ws = torch.empty((B,), dtype=torch.int64, device=dev)
    _call("cofi_pnp_ransac", _p(image_points), _p(object_points), _p(count), stride, n, B, _p(cam), int(iterations), float(threshold),
          int(seed) & 0xFFFFFFFFFFFFFFFF, _p(o_cnt), _p(o_hyp), _p(o_pose), _p(o_inl), _p(ws), _st())
    return o_cnt, o_hyp, o_pose, o_inl


# ------------------------------------------------------------------------------------------ fused training losses
def desc_loss_fwd(img_tok, pix, pc_tok, kpt, mask, frames: int, pos_margin: float, neg_margin: float, log_scale: float = 10.0,
                  want_grad: bool = True, want_dists: bool = False):
    """cofi_desc_loss: per-frame loss [frames] (+ dists [frames,n,n]) and the compact gradients w.r.t. the gathered rows."""
    img_tok, pc_tok = _f32(img_tok, "img_tok").contiguous(), _f32(pc_tok, "pc_tok").contiguous()
    pix, kpt, mask = _i64(pix, "pix"), _i64(kpt, "kpt"), _f32(mask, "mask").contiguous()
    n, C = pix.numel() // frames, img_tok.shape[1]
    dev = img_tok.device
    loss = torch.empty((frames,), dtype=torch.float32, device=dev)
    dists = torch.empty((frames, n, n), dtype=torch.float32, device=dev) if want_dists else None
    d_img = torch.empty((frames, n, C), dtype=torch.float32, device=dev) if want_grad else None
    d_pc = torch.empty((frames, n, C), dtype=torch.float32, device=dev) if want_grad else None
    _call("cofi_desc_loss", _p(img_tok), _p(pix), img_tok.shape[0] // frames, _p(pc_tok), _p(kpt), pc_tok.shape[0] // frames,
          _p(mask), n, C, frames, float(pos_margin), float(neg_margin), float(log_scale), _p(loss), _p(dists), _p(d_img),
          _p(d_pc), _st())
    return loss, dists, d_img, d_pc


def overlap_loss_fwd(score_tok, idx, n_in: int, frames: int, want_grad: bool = True):
    score_tok, idx = _f32(score_tok, "score").contiguous(), _i64(idx, "idx")
    N = idx.numel() // frames
    loss = torch.empty((frames,), dtype=torch.float32, device=score_tok.device)
    d = torch.empty((frames, N), dtype=torch.float32, device=score_tok.device) if want_grad else None
    _call("cofi_overlap_loss", _p(score_tok), _p(idx), score_tok.numel() // frames, n_in, N - n_in, frames, _p(loss), _p(d), _st())
    return loss, d


def fine_circle_loss_fwd(patch, fpc, rel, frames: int, bad_flag: Optional[torch.Tensor] = None, m: float = 0.2,
                         gamma: float = 5.0, want_grad: bool = True):
    patch, fpc, rel = _f32(patch, "patch").contiguous(), _f32(fpc, "fpc").contiguous(), _i64(rel, "rel")
    rows, C = fpc.shape
    loss = torch.empty((frames,), dtype=torch.float32, device=fpc.device)
    d_patch = torch.empty_like(patch) if want_grad else None
    d_fpc = torch.empty_like(fpc) if want_grad else None
    _call("cofi_fine_circle_loss", _p(patch), _p(fpc), _p(rel), rows // frames, C, frames, float(m), float(gamma), _p(loss),
          _p(d_patch), _p(d_fpc), _p(bad_flag), _st())
    return loss, d_patch, d_fpc


def scatter_scaled_rows(src, idx, rows_dst: int, frames: int, scale_dev: Optional[torch.Tensor], scale_host: float = 1.0):
    """dst[f*rows_dst + idx[f, r]] += scale * src[f, r] into a fresh zero tensor (idx None: dst = scale * src)."""
    src = _f32(src, "src").contiguous()
    C = src.shape[-1]
    R = src.numel() // C // frames
    if idx is None:
        dst = torch.empty_like(src)
    else:
        dst = torch.zeros((frames * rows_dst, C), dtype=torch.float32, device=src.device)
    _call("cofi_scatter_scaled_rows", _p(src), _p(idx), R, rows_dst, frames, C, _p(scale_dev), float(scale_host), _p(dst), _st())
    return dst


def adam_step(p, g, m, v, lr: float, beta1: float, beta2: float, eps: float, step: int, grad_scale: float = 1.0):
    _call("cofi_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
          float(grad_scale), _st())
