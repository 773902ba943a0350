"""Differentiable host-side wrappers over the libnmb200 C ABI.

torch.autograd only *sequences* the calls: every forward and every backward is a
libnmb200 kernel launched on the current CUDA stream through ctypes.  Parameters
may carry an ``nm_grad`` attribute (a view into the flat gradient arena of
``neuralmonkey_b200.params.ParameterArena``); the backward passes then accumulate
weight gradients straight into that arena (GEMM epilogue ``beta = 1``) and return
``None`` to autograd, so no torch kernels run on the weight-gradient path.
"""
import os
from typing import Optional, Tuple

import torch

from neuralmonkey_b200 import lib
from neuralmonkey_b200.lib import call, ptr

_GEMM_BACKEND = lib.GEMM_AUTO


def set_gemm_backend(name: str) -> None:
    """'auto' (tcgen05 when TMA-addressable), 'simt' (exact fp32) or 'tc'."""
    global _GEMM_BACKEND
    _GEMM_BACKEND = {"auto": lib.GEMM_AUTO, "simt": lib.GEMM_SIMT, "tc": lib.GEMM_TC}[name]
    # the GRU recurrence follows: exact fp32 engine with 'simt', tensor cores otherwise
    call("nm_gru_set_mode", 1 if name == "simt" else 0)


def gemm_backend() -> int:
    return _GEMM_BACKEND


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError("expected float32, got {}".format(t.dtype))
    return t


def _rows(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """2-D view with unit inner stride; returns (tensor, leading dimension)."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.stride(1) != 1 and t.size(1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))
    if ld < t.size(1):
        t = t.contiguous()
        ld = t.size(1)
    return t, ld


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, trans_a: bool = False,
         trans_b: bool = False, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
         beta: float = 0.0, backend: Optional[int] = None) -> torch.Tensor:
    """out = act(op(a) @ op(b) + bias) + beta*out, all through nm_gemm (no torch math)."""
    a, lda = _rows(_f32(a))
    b, ldb = _rows(_f32(b))
    if out.dim() != 2 or (out.stride(1) != 1 and out.size(1) != 1):
        raise ValueError("gemm output must be a 2-D tensor with unit inner stride")
    ldc = out.stride(0) if out.size(0) > 1 else max(out.size(1), out.stride(0))
    m, k = (a.size(1), a.size(0)) if trans_a else (a.size(0), a.size(1))
    kb, n = (b.size(1), b.size(0)) if trans_b else (b.size(0), b.size(1))
    if k != kb or out.size(0) != m or out.size(1) != n:
        raise ValueError("gemm shape mismatch: op(a) [{},{}] op(b) [{},{}] out {}".format(
            m, k, kb, n, tuple(out.shape)))
    call("nm_gemm", int(trans_a), int(trans_b), m, n, k, ptr(a), lda, ptr(b), ldb, ptr(out), ldc,
         ptr(bias), lib.NM_ACT[act], float(beta),
         _GEMM_BACKEND if backend is None else backend, lib.stream())
    return out


def _sink(t: torch.Tensor) -> Optional[torch.Tensor]:
    return getattr(t, "nm_grad", None)


# -- weight gradients off the critical path ----------------------------------------------------------------
# The backward pass is a chain: every layer's input gradient feeds the next (older) layer, while its WEIGHT
# gradient feeds nothing until the optimizer runs.  A trainer may therefore open a window
# (`weight_grad_stream(True)` ... `join_weight_grads()`) in which the weight / bias gradient products of the dense
# layers and of the vocabulary projection are issued on a second stream: they fill the SMs the chain leaves idle
# (launch latencies, short grids, the 120-CTA recurrences) instead of lengthening it.  Both streams only ever ADD
# into disjoint parts of the flat gradient buffer: a variable's contributions all travel on the same stream.
# Inside a CUDA-graph capture the fork and the join become graph edges.  Off unless a trainer opens the window
# (NMB200_WGRAD_STREAM=0 keeps it shut).
_wg = {"stream": None, "keep": [], "open": False}


def weight_grad_stream(enable: bool) -> None:
    if enable and (os.environ.get("NMB200_WGRAD_STREAM", _WGRAD_STREAM_DEFAULT) != "1" or lib.profiling()):
        enable = False      # (the per-call profiler times calls one by one: no overlap while it runs)
    if enable and _wg["stream"] is None:
        _wg["stream"] = torch.cuda.Stream()
    _wg["open"] = bool(enable)


_WGRAD_STREAM_DEFAULT = "1"    # Transformer step 12.75 -> 12.25 ms, en-de unchanged; NMB200_WGRAD_STREAM=0 shuts it


def join_weight_grads() -> None:
    """The current stream waits for every weight gradient issued so far (and their operands may be freed)."""
    if _wg["keep"]:
        torch.cuda.current_stream().wait_stream(_wg["stream"])
        del _wg["keep"][:]


def _off_the_chain(fn, *operands) -> None:
    """Run `fn` (kernels that only add into the gradient buffer) on the weight-gradient stream when a trainer
    has opened the window, else right here.  `operands` are kept alive until the join."""
    if not _wg["open"]:
        fn()
        return
    side = _wg["stream"]
    side.wait_stream(torch.cuda.current_stream())      # the operands were produced on the chain's stream
    with torch.cuda.stream(side):
        fn()
    _wg["keep"].append(operands)


def _gru_weight_grads(x2, hp2, rh2, dzg, dzc, wg, wc, e, sinks):
    """Weight / bias gradients of one GRU direction (rows [:E] of the kernels from x, rows [E:] from the recurrent
    operand).  With all four variables in the gradient buffer the six launches leave the backward chain
    (`_off_the_chain`) and nothing is returned; otherwise (dWg, dbg, dWc, dbc) with None for buffered ones."""
    sg, sbg, sc, sbc = sinks

    def products(dwg, dwc, beta_g, beta_c):
        gemm(x2, dzg, dwg[:e], trans_a=True, beta=beta_g)
        gemm(hp2, dzg, dwg[e:], trans_a=True, beta=beta_g)
        gemm(x2, dzc, dwc[:e], trans_a=True, beta=beta_c)
        gemm(rh2, dzc, dwc[e:], trans_a=True, beta=beta_c)
    if sg is not None and sbg is not None and sc is not None and sbc is not None:
        def into_the_buffer():
            products(sg, sc, 1.0, 1.0)
            _bias_grad(dzg, sbg)
            _bias_grad(dzc, sbc)
        _off_the_chain(into_the_buffer, x2, hp2, rh2, dzg, dzc)
        return None, None, None, None
    dwg = sg if sg is not None else torch.empty_like(wg)
    dwc = sc if sc is not None else torch.empty_like(wc)
    products(dwg, dwc, 1.0 if sg is not None else 0.0, 1.0 if sc is not None else 0.0)
    return (None if sg is not None else dwg, _bias_grad(dzg, sbg), None if sc is not None else dwc,
            _bias_grad(dzc, sbc))


def _weight_grad(a: torch.Tensor, b: torch.Tensor, trans_a: bool, trans_b: bool,
                 sink: Optional[torch.Tensor], shape) -> Optional[torch.Tensor]:
    """op(a) @ op(b) accumulated into `sink` (returns None) or returned as a new tensor."""
    if sink is not None:
        gemm(a, b, sink, trans_a, trans_b, beta=1.0)
        return None
    out = torch.empty(shape, device=a.device, dtype=torch.float32)
    return gemm(a, b, out, trans_a, trans_b)


def _bias_grad(dy: torch.Tensor, sink: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    dy2, ld = _rows(dy)
    if sink is not None:
        call("nm_colsum", ptr(dy2), dy2.size(0), dy2.size(1), ld, ptr(sink), 1, lib.stream())
        return None
    out = torch.empty(dy2.size(1), device=dy.device, dtype=torch.float32)
    call("nm_colsum", ptr(dy2), dy2.size(0), dy2.size(1), ld, ptr(out), 0, lib.stream())
    return out


# ---------------------------------------------------------------------------
# K1 embedding
# ---------------------------------------------------------------------------
class _Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, mask):
        ids = ids.contiguous()
        n = ids.numel()
        v, e = table.shape
        out = torch.empty(tuple(ids.shape) + (e,), device=table.device, dtype=torch.float32)
        mask_c = mask.contiguous() if mask is not None else None
        call("nm_embed_fwd", ptr(ids), ptr(table), ptr(mask_c), ptr(out), n, e, v, lib.stream())
        ctx.save_for_backward(ids, mask_c)
        ctx.shape = (v, e)
        ctx.sink = _sink(table)
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, mask = ctx.saved_tensors
        v, e = ctx.shape
        dout = dout.contiguous()
        if ctx.sink is not None:
            dtable, ret = ctx.sink, None
        else:
            dtable = torch.zeros(v, e, device=dout.device, dtype=torch.float32)
            ret = dtable
        call("nm_embed_bwd", ptr(ids), ptr(dout), ptr(mask), ptr(dtable), ids.numel(), e, v,
             lib.stream())
        return None, ret, None


def embed(ids: torch.Tensor, table: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """table[ids] * mask[..., None]  (model/sequence.py:181-191; autoregressive.py:269-272)."""
    return _Embed.apply(ids, table, mask)


# ---------------------------------------------------------------------------
# dense projection
# ---------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        y = torch.empty(x2.size(0), w.size(1), device=x.device, dtype=torch.float32)
        gemm(x2, w, y, bias=b, act=act)
        ctx.save_for_backward(x2, w, y if act is not None else None)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.sinks = (_sink(w), _sink(b) if b is not None else None)
        ctx.in_shape = shape
        return y.view(tuple(shape[:-1]) + (w.size(1),))

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dy2, _ = _rows(dy2)
        if ctx.act is not None:
            dpre = torch.empty_like(y)
            dyc = dy2.contiguous()
            call("nm_act_bwd", ptr(y), ptr(dyc), ptr(dpre), y.numel(), lib.NM_ACT[ctx.act],
                 lib.stream())
        else:
            dpre = dy2
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x2.shape, device=dy.device, dtype=torch.float32)
            gemm(dpre, w, dx, trans_b=True)
            dx = dx.view(ctx.in_shape)
        w_sink, b_sink = ctx.sinks
        want_w, want_b = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        if (want_w and w_sink is not None) and (not want_b or b_sink is not None):
            def into_the_buffer():      # both accumulate into the gradient buffer: nothing to return
                _weight_grad(x2, dpre, True, False, w_sink, w.shape)
                if want_b:
                    _bias_grad(dpre, b_sink)
            _off_the_chain(into_the_buffer, x2, dpre)
            return dx, None, None, None
        if want_w:
            dw = _weight_grad(x2, dpre, True, False, w_sink, w.shape)
        if want_b:
            db = _bias_grad(dpre, b_sink)
        return dx, dw, db, None


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None,
           act: Optional[str] = None) -> torch.Tensor:
    """act(x @ w + b) over the last dim; w is [in, out] (tf.layers.dense layout)."""
    return _Linear.apply(x, w, b, act)


class _Dropout(torch.autograd.Function):
    """y = dropout(x) (+ residual), masks drawn inside the kernel (K15, csrc/dropout.cu); the backward pass
    re-draws the mask of the same (site, step) instead of reading one."""

    @staticmethod
    def forward(ctx, x, residual, keep_prob, site):
        from neuralmonkey_b200 import runtime
        xc = _f32(x).contiguous()
        rc = None if residual is None else _f32(residual).expand_as(x).contiguous()
        y = torch.empty_like(xc)
        state = runtime.dropout_state()
        call("nm_dropout_apply", ptr(xc), ptr(rc), ptr(y), xc.numel(), float(keep_prob), ptr(state), int(site),
             lib.stream())
        ctx.keep_prob, ctx.site, ctx.state = float(keep_prob), int(site), state
        ctx.has_residual = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        dyc = dy.contiguous()
        dx = torch.empty_like(dyc)
        call("nm_dropout_apply", ptr(dyc), None, ptr(dx), dyc.numel(), ctx.keep_prob, ptr(ctx.state), ctx.site,
             lib.stream())
        return dx, (dy if ctx.has_residual else None), None, None


def dropout(x: torch.Tensor, keep_prob: float, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x with each element kept with probability keep_prob and scaled by 1/keep_prob, plus `residual`."""
    from neuralmonkey_b200 import runtime
    return _Dropout.apply(x, residual, keep_prob, runtime.next_dropout_site())


def dropout_mask(shape, keep_prob: float, device=None) -> torch.Tensor:
    """A mask tensor (entries 0 or 1/keep_prob) for the kernels that take one as an operand."""
    from neuralmonkey_b200 import runtime
    mask = torch.empty(tuple(int(d) for d in shape), device=runtime.device(), dtype=torch.float32)
    call("nm_dropout_mask", ptr(mask), mask.numel(), float(keep_prob), ptr(runtime.dropout_state()),
         runtime.next_dropout_site(), lib.stream())
    return mask


class _Maxout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z2 = z.reshape(-1, z.shape[-1]).contiguous()
        m, two_o = z2.shape
        o = two_o // 2
        y = torch.empty(m, o, device=z.device, dtype=torch.float32)
        which = torch.empty(m, o, device=z.device, dtype=torch.uint8)
        call("nm_maxout_fwd", ptr(z2), ptr(y), ptr(which), m, o, lib.stream())
        ctx.save_for_backward(which)
        ctx.in_shape = z.shape
        return y.view(tuple(z.shape[:-1]) + (o,))

    @staticmethod
    def backward(ctx, dy):
        (which,) = ctx.saved_tensors
        m, o = which.shape
        dy2 = dy.reshape(m, o).contiguous()
        dz = torch.empty(m, 2 * o, device=dy.device, dtype=torch.float32)
        call("nm_maxout_bwd", ptr(dy2), ptr(which), ptr(dz), m, o, lib.stream())
        return dz.view(ctx.in_shape)


def maxout(z: torch.Tensor) -> torch.Tensor:
    """y[..., j] = max(z[..., j], z[..., O + j])  (nn/projection.py:7-35 as executed)."""
    return _Maxout.apply(z)


# ---------------------------------------------------------------------------
# gate arithmetic of the step-wise cell variants (N4): one launch per step and direction
# ---------------------------------------------------------------------------
class _NematusGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sg, gi, sc, ci, state):
        sg, gi, sc, ci, state = (t.contiguous() for t in (sg, gi, sc, ci, state))
        bsz, h = state.shape
        out = torch.empty_like(state)
        saved = torch.empty(bsz, 3 * h, device=state.device, dtype=torch.float32)
        call("nm_nematus_gate_fwd", ptr(sg), ptr(gi), ptr(sc), ptr(ci), ptr(state), ptr(out), ptr(saved), bsz, h,
             lib.stream())
        ctx.save_for_backward(saved, sc, state)
        return out

    @staticmethod
    def backward(ctx, dout):
        saved, sc, state = ctx.saved_tensors
        bsz, h = state.shape
        dout = dout.contiguous()
        dgates = torch.empty(bsz, 2 * h, device=state.device, dtype=torch.float32)
        dcpre, dsc, dstate = torch.empty_like(state), torch.empty_like(state), torch.empty_like(state)
        call("nm_nematus_gate_bwd", ptr(dout), ptr(saved), ptr(sc), ptr(state), ptr(dgates), ptr(dcpre), ptr(dsc),
             ptr(dstate), bsz, h, lib.stream())
        return dgates, dgates, dsc, dcpre, dstate


def nematus_gru_gate(state_gates: torch.Tensor, input_gates: torch.Tensor, state_cand: torch.Tensor,
                     input_cand: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """NematusGRUCell after its projections (nn/ortho_gru_cell.py:86-105): [r,u] = sigmoid(state_gates +
    input_gates); cand = tanh(state_cand * r + input_cand); u * state + (1 - u) * cand."""
    return _NematusGate.apply(state_gates, input_gates, state_cand, input_cand, state)


class _LSTMGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, c):
        z, c = z.contiguous(), c.contiguous()
        bsz, h = c.shape
        new_c, new_h = torch.empty_like(c), torch.empty_like(c)
        saved = torch.empty(bsz, 5 * h, device=c.device, dtype=torch.float32)
        call("nm_lstm_gate_fwd", ptr(z), ptr(c), ptr(new_c), ptr(new_h), ptr(saved), bsz, h, lib.stream())
        ctx.save_for_backward(saved, c)
        return new_c, new_h

    @staticmethod
    def backward(ctx, dnew_c, dnew_h):
        saved, c = ctx.saved_tensors
        bsz, h = c.shape
        dz = torch.empty(bsz, 4 * h, device=c.device, dtype=torch.float32)
        dc = torch.empty_like(c)
        call("nm_lstm_gate_bwd", ptr(dnew_c.contiguous() if dnew_c is not None else None),
             ptr(dnew_h.contiguous() if dnew_h is not None else None), ptr(saved), ptr(c), ptr(dz), ptr(dc), bsz, h,
             lib.stream())
        return dz, dc


def lstm_gate(z: torch.Tensor, c: torch.Tensor):
    """tf LSTMCell defaults after the [x, h] projection: z = (i, j, f, o); returns (c', h') with
    c' = sigmoid(f + 1) * c + sigmoid(i) * tanh(j), h' = sigmoid(o) * tanh(c')."""
    return _LSTMGate.apply(z, c)


# ---------------------------------------------------------------------------
# K7 layer norm
# ---------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        m, d = x2.shape
        y = torch.empty_like(x2)
        mean = torch.empty(m, device=x.device, dtype=torch.float32)
        rstd = torch.empty(m, device=x.device, dtype=torch.float32)
        call("nm_layernorm_fwd", ptr(x2), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), m, d,
             float(eps), lib.stream())
        ctx.save_for_backward(x2, gamma, mean, rstd)
        ctx.sinks = (_sink(gamma), _sink(beta))
        ctx.in_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma, mean, rstd = ctx.saved_tensors
        m, d = x2.shape
        dy2 = dy.reshape(m, d).contiguous()
        dx = torch.empty_like(x2)
        sg, sb = ctx.sinks
        if sg is not None and sb is not None:
            # the input gradient continues the chain; the parameter gradients only add into the gradient buffer
            call("nm_layernorm_bwd", ptr(x2), ptr(gamma), ptr(mean), ptr(rstd), ptr(dy2), ptr(dx), None, None, m, d,
                 lib.stream())
            _off_the_chain(lambda: call("nm_layernorm_bwd", ptr(x2), ptr(gamma), ptr(mean), ptr(rstd), ptr(dy2), None,
                                        ptr(sg), ptr(sb), m, d, lib.stream()), x2, dy2, mean, rstd)
            return dx.view(ctx.in_shape), None, None, None
        dg = sg if sg is not None else torch.zeros(d, device=dy.device, dtype=torch.float32)
        db = sb if sb is not None else torch.zeros(d, device=dy.device, dtype=torch.float32)
        call("nm_layernorm_bwd", ptr(x2), ptr(gamma), ptr(mean), ptr(rstd), ptr(dy2), ptr(dx), ptr(dg),
             ptr(db), m, d, lib.stream())
        return (dx.view(ctx.in_shape), None if sg is not None else dg,
                None if sb is not None else db, None)


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               eps: float = 1e-6) -> torch.Tensor:
    """tf_utils.layer_norm (tf_utils.py:189-219)."""
    return _LayerNorm.apply(x, gamma, beta, eps)


# ---------------------------------------------------------------------------
# K2 GRU layer (input projection hoisted onto the tensor cores + recurrent kernels)
# ---------------------------------------------------------------------------
class _GRULayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wg, bg, wc, bc, h0, lengths, reverse, drop_mask, sm_budget):
        # x [B,T,E]; wg [E+H,2H], bg [2H]; wc [E+H,H], bc [H]   (TF GRUCell kernels)
        bsz, t, e = x.shape
        h = wc.size(1)
        x2 = x.reshape(bsz * t, e)
        xproj = torch.empty(bsz * t, 3 * h, device=x.device, dtype=torch.float32)
        gemm(x2, wg[:e], xproj[:, :2 * h], bias=bg)
        gemm(x2, wc[:e], xproj[:, 2 * h:], bias=bc)
        states = torch.empty(bsz, t, h, device=x.device, dtype=torch.float32)
        raw = torch.empty_like(states) if drop_mask is not None else None
        final = torch.empty(bsz, h, device=x.device, dtype=torch.float32)
        gates = torch.empty(bsz, t, 3 * h, device=x.device, dtype=torch.float32)
        hprev = torch.empty(bsz, t, h, device=x.device, dtype=torch.float32)
        rh = torch.empty(bsz, t, h, device=x.device, dtype=torch.float32)
        h0c = h0.contiguous() if h0 is not None else None
        dm = drop_mask.contiguous() if drop_mask is not None else None
        call("nm_gru_seq_fwd", ptr(xproj), ptr(wg[e:]), ptr(wc[e:]), ptr(h0c), ptr(lengths), ptr(dm),
             int(reverse), ptr(states), ptr(raw), ptr(final), ptr(gates), ptr(hprev), ptr(rh), bsz, t,
             h, int(sm_budget), lib.stream())
        ctx.save_for_backward(x2, wg, wc, lengths, gates, hprev, rh, dm)
        ctx.dims = (bsz, t, e, h)
        ctx.reverse = reverse
        ctx.sm_budget = sm_budget
        ctx.has_h0 = h0 is not None
        ctx.sinks = (_sink(wg), _sink(bg), _sink(wc), _sink(bc))
        return states, final, (raw if raw is not None else states)

    @staticmethod
    def backward(ctx, dstates, dfinal, draw):
        x2, wg, wc, lengths, gates, hprev, rh, dm = ctx.saved_tensors
        bsz, t, e, h = ctx.dims
        dev = x2.device
        dstates = dstates.contiguous() if dstates is not None else None
        dfinal = dfinal.contiguous() if dfinal is not None else None
        if draw is not None:
            if dm is None:  # raw aliases states: autograd delivers the two gradients separately
                dstates = draw.contiguous() if dstates is None else dstates + draw
                draw = None
            else:
                draw = draw.contiguous()
        dxproj = torch.empty(bsz * t, 3 * h, device=dev, dtype=torch.float32)
        dh0 = torch.empty(bsz, h, device=dev, dtype=torch.float32) if ctx.has_h0 else None
        work = torch.empty(2 * bsz * h, device=dev, dtype=torch.float32)
        call("nm_gru_seq_bwd", ptr(wg[e:]), ptr(wc[e:]), ptr(lengths), ptr(dm), int(ctx.reverse),
             ptr(gates), ptr(hprev), ptr(dstates), ptr(draw), ptr(dfinal), ptr(dxproj), ptr(dh0),
             ptr(work), bsz, t, h, int(ctx.sm_budget), lib.stream())
        dzg, dzc = dxproj[:, :2 * h], dxproj[:, 2 * h:]
        dwg, dbg, dwc, dbc = _gru_weight_grads(x2, hprev.view(bsz * t, h), rh.view(bsz * t, h), dzg, dzc, wg, wc, e,
                                               ctx.sinks)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(bsz * t, e, device=dev, dtype=torch.float32)
            gemm(dzg, wg[:e], dx, trans_b=True)
            gemm(dzc, wc[:e], dx, trans_b=True, beta=1.0)
            dx = dx.view(bsz, t, e)
        return (dx, dwg, dbg, dwc, dbc, dh0, None, None, None, None)


def gru_layer(x: torch.Tensor, gates_kernel: torch.Tensor, gates_bias: torch.Tensor,
              cand_kernel: torch.Tensor, cand_bias: torch.Tensor,
              h0: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None,
              reverse: bool = False,
              drop_mask: Optional[torch.Tensor] = None, sm_budget: int = 0):
    """dynamic_rnn over a TF-1.12 GRUCell (encoders/recurrent.py:71-110).

    Returns (outputs [B,T,H], final state [B,H], raw outputs [B,T,H] = outputs before
    `drop_mask`); with `lengths` (int32) the outputs past
    each length are zero and the state is carried, with `reverse` the sequence is walked
    backwards inside its length (tf.reverse_sequence semantics)."""
    return _GRULayer.apply(x, gates_kernel, gates_bias, cand_kernel, cand_bias, h0, lengths, reverse,
                           drop_mask, sm_budget)


class _BiGRULayer(torch.autograd.Function):
    """Forward and backward direction of a bidirectional GRU layer over the same input, their recurrences
    in ONE launch each way (nm_gru_seq_fwd_pair / nm_gru_seq_bwd_pair): the two directions are
    independent, and one direction alone occupies only half of the SMs for T dependent steps."""

    @staticmethod
    def forward(ctx, x, lengths, wg_f, bg_f, wc_f, bc_f, wg_b, bg_b, wc_b, bc_b):
        bsz, t, e = x.shape
        h = wc_f.size(1)
        x2 = x.reshape(bsz * t, e)
        dev = x.device
        saved = []
        outs = []
        bufs = []
        for wg, bg, wc, bc in ((wg_f, bg_f, wc_f, bc_f), (wg_b, bg_b, wc_b, bc_b)):
            xproj = torch.empty(bsz * t, 3 * h, device=dev, dtype=torch.float32)
            gemm(x2, wg[:e], xproj[:, :2 * h], bias=bg)
            gemm(x2, wc[:e], xproj[:, 2 * h:], bias=bc)
            states = torch.empty(bsz, t, h, device=dev, dtype=torch.float32)
            final = torch.empty(bsz, h, device=dev, dtype=torch.float32)
            gates = torch.empty(bsz, t, 3 * h, device=dev, dtype=torch.float32)
            hprev = torch.empty(bsz, t, h, device=dev, dtype=torch.float32)
            rh = torch.empty(bsz, t, h, device=dev, dtype=torch.float32)
            bufs.append((xproj, states, final, gates, hprev, rh))
        (xp_f, st_f, fi_f, ga_f, hp_f, rh_f), (xp_b, st_b, fi_b, ga_b, hp_b, rh_b) = bufs
        call("nm_gru_seq_fwd_pair",
             ptr(xp_f), ptr(wg_f[e:]), ptr(wc_f[e:]), 0, ptr(st_f), ptr(fi_f), ptr(ga_f), ptr(hp_f), ptr(rh_f),
             ptr(xp_b), ptr(wg_b[e:]), ptr(wc_b[e:]), 1, ptr(st_b), ptr(fi_b), ptr(ga_b), ptr(hp_b), ptr(rh_b),
             ptr(lengths), bsz, t, h, lib.stream())
        ctx.save_for_backward(x2, lengths, wg_f, wc_f, wg_b, wc_b, ga_f, hp_f, rh_f, ga_b, hp_b, rh_b)
        ctx.dims = (bsz, t, e, h)
        ctx.sinks = tuple(_sink(w) for w in (wg_f, bg_f, wc_f, bc_f, wg_b, bg_b, wc_b, bc_b))
        return st_f, fi_f, st_b, fi_b

    @staticmethod
    def backward(ctx, dst_f, dfi_f, dst_b, dfi_b):
        x2, lengths, wg_f, wc_f, wg_b, wc_b, ga_f, hp_f, rh_f, ga_b, hp_b, rh_b = ctx.saved_tensors
        bsz, t, e, h = ctx.dims
        dev = x2.device

        def c(g):
            return g.contiguous() if g is not None else None

        dst_f, dfi_f, dst_b, dfi_b = c(dst_f), c(dfi_f), c(dst_b), c(dfi_b)
        dxp_f = torch.empty(bsz * t, 3 * h, device=dev, dtype=torch.float32)
        dxp_b = torch.empty(bsz * t, 3 * h, device=dev, dtype=torch.float32)
        work = torch.empty(2 * bsz * h, device=dev, dtype=torch.float32)
        call("nm_gru_seq_bwd_pair",
             ptr(wg_f[e:]), ptr(wc_f[e:]), 0, ptr(ga_f), ptr(hp_f), ptr(dst_f), ptr(dfi_f), ptr(dxp_f),
             ptr(wg_b[e:]), ptr(wc_b[e:]), 1, ptr(ga_b), ptr(hp_b), ptr(dst_b), ptr(dfi_b), ptr(dxp_b),
             ptr(lengths), ptr(work), bsz, t, h, lib.stream())
        grads = []
        dx = torch.empty(bsz * t, e, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        first = True
        for (wg, wc, hp, rh, dxp, sinks) in ((wg_f, wc_f, hp_f, rh_f, dxp_f, ctx.sinks[:4]),
                                             (wg_b, wc_b, hp_b, rh_b, dxp_b, ctx.sinks[4:])):
            dzg, dzc = dxp[:, :2 * h], dxp[:, 2 * h:]
            if dx is not None:      # the chain first: the input gradient is what the older layers wait for
                gemm(dzg, wg[:e], dx, trans_b=True, beta=0.0 if first else 1.0)
                gemm(dzc, wc[:e], dx, trans_b=True, beta=1.0)
                first = False
            grads += list(_gru_weight_grads(x2, hp.view(bsz * t, h), rh.view(bsz * t, h), dzg, dzc, wg, wc, e, sinks))
        return (dx.view(bsz, t, e) if dx is not None else None, None) + tuple(grads)


def gru_bilayer(x: torch.Tensor, lengths: torch.Tensor, cell_fw, cell_bw):
    """tf.nn.bidirectional_dynamic_rnn over two TF-1.12 GRUCells (encoders/recurrent.py:82-95).  cell_* =
    (gates kernel, gates bias, candidate kernel, candidate bias).  Returns (outputs fw [B,T,H], final fw [B,H],
    outputs bw, final bw); the backward direction walks each sentence from its last token
    (tf.reverse_sequence semantics), outputs at their original time index."""
    return _BiGRULayer.apply(x, lengths, *cell_fw, *cell_bw)


# ---------------------------------------------------------------------------
# K4 Bahdanau attention
# ---------------------------------------------------------------------------
class _Bahdanau(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keys, values, mask, qproj, v, bias):
        bsz, tx, a = keys.shape
        c = values.size(2)
        nq = qproj.size(1)
        keys, values, qproj = keys.contiguous(), values.contiguous(), qproj.contiguous()
        mask_c = mask.contiguous() if mask is not None else None
        dev = keys.device
        energies = torch.empty(bsz, nq, tx, device=dev, dtype=torch.float32)
        weights = torch.empty(bsz, nq, tx, device=dev, dtype=torch.float32)
        ctxv = torch.empty(bsz, nq, c, device=dev, dtype=torch.float32)
        call("nm_bahdanau_fwd", ptr(keys), ptr(values), ptr(mask_c), ptr(qproj), ptr(v), ptr(bias),
             ptr(energies), ptr(weights), ptr(ctxv), bsz, tx, nq, a, c, lib.stream())
        ctx.save_for_backward(keys, values, mask_c, qproj, v, energies, weights)
        ctx.sinks = (_sink(v), _sink(bias))
        ctx.mark_non_differentiable(weights)
        return ctxv, weights

    @staticmethod
    def backward(ctx, dctx, _dweights):
        keys, values, mask, qproj, v, energies, weights = ctx.saved_tensors
        bsz, tx, a = keys.shape
        c = values.size(2)
        nq = qproj.size(1)
        dev = keys.device
        dctx = dctx.contiguous()
        dkeys = torch.empty_like(keys)
        dvalues = torch.empty_like(values)
        dq = torch.empty_like(qproj)
        sv, sb = ctx.sinks
        dv = sv if sv is not None else torch.zeros(a, device=dev, dtype=torch.float32)
        db = sb if sb is not None else torch.zeros(1, device=dev, dtype=torch.float32)
        work = torch.empty(bsz * nq * tx, device=dev, dtype=torch.float32)
        call("nm_bahdanau_bwd", ptr(keys), ptr(values), ptr(mask), ptr(qproj), ptr(v), ptr(energies),
             ptr(weights), ptr(dctx), ptr(dkeys), ptr(dvalues), ptr(dq), ptr(dv), ptr(db), ptr(work),
             bsz, tx, nq, a, c, lib.stream())
        return (dkeys, dvalues, None, dq, None if sv is not None else dv,
                None if sb is not None else db)


def bahdanau_attention(keys: torch.Tensor, values: torch.Tensor, mask: Optional[torch.Tensor],
                       qproj: torch.Tensor, v: torch.Tensor,
                       bias: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Attention.attention (attention/feed_forward.py:125-166) for all query steps at once.

    keys [B,Tx,A], values [B,Tx,C], mask [B,Tx] or None, qproj [B,NQ,A], v [A], bias [1].
    Returns (contexts [B,NQ,C], weights [B,NQ,Tx])."""
    return _Bahdanau.apply(keys, values, mask, qproj, v, bias)


# ---------------------------------------------------------------------------
# K5/K6 vocabulary projection + cross-entropy
# ---------------------------------------------------------------------------
class _SmoothingTerm(torch.autograd.Function):
    """logit[m, target[m]] - mean_v logit[m, v] for logits = x @ W + b (+ -1e9 on the <unk> column) without
    materialising them: label smoothing adds `eps` times this to the plain cross-entropy
    (xent_smoothed = lse - (1 - eps) * logit_t - eps * mean_v logit = xent + eps * (logit_t - mean_v logit)).
    Host-level torch arithmetic on [M, K] tensors (a gather of M weight columns and a mean column); the
    weight / bias gradients are accumulated into the arena's gradient buffer like every other op's."""

    @staticmethod
    def forward(ctx, x, w, b, targets, unk_index, trans_w):
        vocab = w.size(0) if trans_w else w.size(1)
        cols = w.index_select(0, targets) if trans_w else w.index_select(1, targets).t()
        wmean = w.mean(0 if trans_w else 1)
        value = (x * cols).sum(1) - x @ wmean
        if b is not None:
            value = value + b.index_select(0, targets) - b.mean()
        if unk_index >= 0:
            value = value - 1e9 * (targets == unk_index).to(value.dtype) + 1e9 / vocab
        ctx.save_for_backward(x, cols, wmean, targets)
        ctx.cfg = (vocab, trans_w, b is not None, w.shape)
        ctx.sinks = (_sink(w), _sink(b) if b is not None else None)
        return value

    @staticmethod
    def backward(ctx, g):
        x, cols, wmean, targets = ctx.saved_tensors
        vocab, trans_w, has_bias, w_shape = ctx.cfg
        w_sink, b_sink = ctx.sinks
        dx = g.unsqueeze(1) * (cols - wmean) if ctx.needs_input_grad[0] else None
        gx = x * g.unsqueeze(1)                                    # [M, K]
        dw = w_sink if w_sink is not None else torch.zeros(w_shape, device=x.device, dtype=torch.float32)
        if trans_w:                                                 # W is [V, K]
            dw.index_add_(0, targets, gx)
            dw.sub_((gx.sum(0) / vocab).unsqueeze(0))
        else:                                                       # W is [K, V]
            dw.index_add_(1, targets, gx.t().contiguous())
            dw.sub_((gx.sum(0) / vocab).unsqueeze(1))
        db = None
        if has_bias:
            db = b_sink if b_sink is not None else torch.zeros(vocab, device=x.device, dtype=torch.float32)
            db.index_add_(0, targets, g)
            db.sub_(g.sum() / vocab)
        return (dx, None if w_sink is not None else dw, None if (b_sink is not None or not has_bias) else db,
                None, None, None)


def smoothing_term(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], targets: torch.Tensor,
                   unk_index: int = -1, trans_w: bool = False) -> torch.Tensor:
    return _SmoothingTerm.apply(x, w, b, targets, unk_index, trans_w)


def _xent16_enabled() -> bool:
    """The vocabulary projection with fp16 operands and fp16 dlogits (csrc/xent16.cu) is the default on
    the tensor-core engine; NMB200_XENT16=0 keeps every product in TF32 (the round-1 path)."""
    import os
    return os.environ.get("NMB200_XENT16", "1") != "0"


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _LogitsXent16(torch.autograd.Function):
    """The fp16-operand variant of _LogitsXent for W stored [K,V] with the bias right behind it in the
    gradient buffer (the layout the decoders declare).  Same results within TF32-class rounding (fp16
    has TF32's 10 mantissa bits; the operands here are O(1): activations after tanh, U(-0.5, 0.5)-scale
    weights, and (softmax - onehot) in [-1, 1]).

        logits = X16 [M,K] . WT16 [V,K]^T                  forward, and the recompute of the backward
        P16    = half((softmax - onehot) * mask) [M,V]      written ONCE, row-major (0.8 GB at the bench shape,
                                                            against 1.6 GB fp32 written once and read twice)
        dX     = P16 . W16 [K,V]^T * upstream[m]            K-major x K-major
        dW,db  = [X * u, u]16^T . P16 * max|upstream|       both operands MN-major (tcgen05 takes either
                                                            major for kind::f16): no transposed copy of P
    The upstream per-row gradient u = upstream / max|upstream| rides in the fp16 copy of X for dW and in
    the fp32 epilogue for dX, so P itself stays unnormalised."""

    @staticmethod
    def forward(ctx, x, w, b, targets, weights, unk_index, keep_logits):
        x2 = x.reshape(-1, x.shape[-1])
        x2, ldx = _rows(x2)
        m, k = x2.shape
        v = w.size(1)
        dev = x.device
        kpad = _pad8(k)
        targets = targets.reshape(-1).contiguous()
        weights = weights.reshape(-1).contiguous()
        w2, ldw = _rows(w)
        x16 = torch.empty(m, kpad, device=dev, dtype=torch.float16)
        call("nm_cast_f16", ptr(x2), ldx, ptr(x16), kpad, m, k, None, 0, 0, lib.stream())
        wt16 = torch.empty(v, kpad, device=dev, dtype=torch.float16)
        call("nm_cast_f16", ptr(w2), ldw, ptr(wt16), kpad, k, v, None, 1, 0, lib.stream())
        lse = torch.empty(m, device=dev, dtype=torch.float32)
        xent = torch.empty(m, device=dev, dtype=torch.float32)
        argmax = torch.empty(m, device=dev, dtype=torch.int64)
        logits = torch.empty(m, v, device=dev, dtype=torch.float32) if keep_logits else None
        part = torch.empty(lib.load().nm_logits_xent_scratch(m, v), device=dev, dtype=torch.float32)
        call("nm_logits_xent_fwd16", ptr(x16), kpad, ptr(wt16), kpad, ptr(b), unk_index, ptr(targets),
             ptr(weights), ptr(lse), ptr(xent), ptr(argmax), ptr(part), ptr(logits), v, m, v, k,
             lib.stream())
        ctx.save_for_backward(x2, w2, b, targets, weights, lse, x16, wt16)
        ctx.cfg = (unk_index, x.shape)
        ctx.sinks = (_sink(w), _sink(b) if b is not None else None)
        ctx.mark_non_differentiable(lse, argmax)
        if keep_logits:
            ctx.mark_non_differentiable(logits)
        return xent, lse, argmax, logits

    @staticmethod
    def backward(ctx, dxent, _dlse, _dargmax, _dlogits):
        x2, w2, b, targets, weights, lse, x16, wt16 = ctx.saved_tensors
        unk_index, in_shape = ctx.cfg
        m, k = x2.shape
        v = w2.size(1)
        dev = x2.device
        kpad, vpad = _pad8(k), _pad8(v)
        _, ldx = _rows(x2)
        _, ldw = _rows(w2)
        # (softmax - onehot) * mask in fp16, row-major; values in [-1, 1]
        dl16 = torch.empty(m, vpad, device=dev, dtype=torch.float16)
        call("nm_logits_xent_bwd16", ptr(x16), kpad, ptr(wt16), kpad, ptr(b), unk_index, ptr(targets),
             ptr(weights), ptr(lse), ptr(dl16), vpad, None, 0, m, v, k, lib.stream())
        upstream = dxent.reshape(-1).to(torch.float32).contiguous()    # per-row factor, applied in fp32
        dx = None
        if ctx.needs_input_grad[0]:
            w16 = torch.empty(k, vpad, device=dev, dtype=torch.float16)
            call("nm_cast_f16", ptr(w2), ldw, ptr(w16), vpad, k, v, None, 0, 0, lib.stream())
            dx = torch.empty(m, k, device=dev, dtype=torch.float32)
            call("nm_gemm_f16", m, k, v, ptr(dl16), vpad, ptr(w16), vpad, ptr(dx), k, None,
                 ptr(upstream), 0.0, 0, lib.stream())
            dx = dx.view(in_shape)
        w_sink, _b_sink = ctx.sinks
        # [dW; db] [K+1, V] += smax * [X * u, u]^T . P with u = upstream / smax: straight into the gradient
        # buffer (the weight rows, then the bias row)
        smax = upstream.abs().amax().clamp_min(1e-30).reshape(1)
        k1pad = _pad8(k + 1)
        xs16 = torch.empty(m, k1pad, device=dev, dtype=torch.float16)
        call("nm_cast_f16", ptr(x2), ldx, ptr(xs16), k1pad, m, k, ptr(upstream / smax), 0, 1, lib.stream())
        sink_aug = torch.as_strided(w_sink, (k + 1, v), (v, 1))
        _off_the_chain(lambda: call("nm_gemm_f16_tn", k + 1, v, m, ptr(xs16), k1pad, ptr(dl16), vpad, ptr(sink_aug),
                                    v, ptr(smax), 1.0, lib.stream()), xs16, dl16, smax, sink_aug)
        return dx, None, None, None, None, None, None


class _LogitsXent(torch.autograd.Function):
    """xent[m] = (logsumexp(x@W+b) - (x@W+b)[target]) * weights[m]; also lse and argmax."""

    @staticmethod
    def forward(ctx, x, w, b, targets, weights, unk_index, trans_w, keep_logits):
        x2 = x.reshape(-1, x.shape[-1])
        x2, ldx = _rows(x2)
        m, k = x2.shape
        v = w.size(0) if trans_w else w.size(1)
        dev = x.device
        targets = targets.reshape(-1).contiguous()
        weights = weights.reshape(-1).contiguous()
        lse = torch.empty(m, device=dev, dtype=torch.float32)
        xent = torch.empty(m, device=dev, dtype=torch.float32)
        argmax = torch.empty(m, device=dev, dtype=torch.int64)
        w2, ldw = _rows(w)
        logits = torch.empty(m, v, device=dev, dtype=torch.float32) if keep_logits else None
        fused = (_GEMM_BACKEND != lib.GEMM_SIMT and
                 lib.load().nm_gemm_uses_tc(0, int(trans_w), m, v, k, ldx, ldw, v) == 1 and
                 x2.data_ptr() % 16 == 0 and w2.data_ptr() % 16 == 0)
        if fused:
            part = torch.empty(lib.load().nm_logits_xent_scratch(m, v), device=dev,
                               dtype=torch.float32)
            call("nm_logits_xent_fwd", ptr(x2), ldx, ptr(w2), ldw, int(trans_w), ptr(b), unk_index,
                 ptr(targets), ptr(weights), ptr(lse), ptr(xent), ptr(argmax), ptr(part),
                 ptr(logits), v, m, v, k, lib.stream())
        else:
            if logits is None:
                logits = torch.empty(m, v, device=dev, dtype=torch.float32)
            bias_eff = b
            if unk_index >= 0:  # fold the -1e9 <unk> mask into the bias vector ([V], host plumbing)
                bias_eff = b.detach().clone() if b is not None else torch.zeros(
                    v, device=dev, dtype=torch.float32)
                bias_eff[unk_index] += -1e9
            gemm(x2, w2, logits, trans_b=trans_w, bias=bias_eff)
            call("nm_xent_fwd", ptr(logits), ptr(targets), ptr(weights), ptr(lse), ptr(xent),
                 ptr(argmax), m, v, v, lib.stream())
        ctx.save_for_backward(x2, w2, b, targets, weights, lse, None if fused else logits)
        ctx.cfg = (fused, unk_index, trans_w, x.shape)
        ctx.sinks = (_sink(w), _sink(b) if b is not None else None)
        ctx.mark_non_differentiable(lse, argmax)
        if keep_logits:
            ctx.mark_non_differentiable(logits)
            return xent, lse, argmax, logits
        return xent, lse, argmax, None

    @staticmethod
    def backward(ctx, dxent, _dlse, _dargmax, _dlogits):
        x2, w2, b, targets, weights, lse, logits = ctx.saved_tensors
        fused, unk_index, trans_w, in_shape = ctx.cfg
        m, k = x2.shape
        v = w2.size(0) if trans_w else w2.size(1)
        dev = x2.device
        # fold the upstream per-row gradient into the row weights (tiny [M] product)
        roww = torch.empty(m, device=dev, dtype=torch.float32)
        ones = torch.ones(1, device=dev, dtype=torch.float32)
        torch.mul(weights, dxent.reshape(-1), out=roww)
        if fused:
            dlogits = torch.empty(m, v, device=dev, dtype=torch.float32)
            _, ldx = _rows(x2)
            _, ldw = _rows(w2)
            call("nm_logits_xent_bwd", ptr(x2), ldx, ptr(w2), ldw, int(trans_w), ptr(b), unk_index,
                 ptr(targets), ptr(roww), ptr(lse), ptr(ones), ptr(dlogits), v, m, v, k, lib.stream())
        else:
            dlogits = logits  # in place
            call("nm_xent_bwd", ptr(logits), ptr(targets), ptr(roww), ptr(lse), ptr(ones),
                 ptr(dlogits), m, v, v, lib.stream())
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(m, k, device=dev, dtype=torch.float32)
            gemm(dlogits, w2, dx, trans_b=not trans_w)
            dx = dx.view(in_shape)
        w_sink, b_sink = ctx.sinks
        want_w, want_b = ctx.needs_input_grad[1], b is not None and ctx.needs_input_grad[2]
        if (want_w and want_b and not trans_w and w_sink is not None and b_sink is not None
                and b_sink.data_ptr() == w_sink.data_ptr() + 4 * k * v and w_sink.is_contiguous()):
            # The bias gradient is the column sum of dlogits = one more row of X^T . dlogits with a
            # column of ones appended to X, and the bias segment sits right behind the weight
            # segment in the flat gradient buffer: ONE GEMM writes both (the 301st row costs no
            # extra tile) instead of re-reading the [M,V] matrix for a column sum.
            kpad = (k + 1 + 3) // 4 * 4
            x_aug = torch.zeros(m, kpad, device=dev, dtype=torch.float32)
            x_aug[:, :k] = x2
            x_aug[:, k] = 1.0
            sink_aug = torch.as_strided(w_sink, (k + 1, v), (v, 1))
            gemm(x_aug[:, :k + 1], dlogits, sink_aug, trans_a=True, beta=1.0)
            return dx, None, None, None, None, None, None, None
        if want_w:
            if trans_w:   # w is [V,K]: dW = dlogits^T @ x
                dw = _weight_grad(dlogits, x2, True, False, w_sink, w2.shape)
            else:         # w is [K,V]: dW = x^T @ dlogits
                dw = _weight_grad(x2, dlogits, True, False, w_sink, w2.shape)
        if want_b:
            db = _bias_grad(dlogits, b_sink)
        return dx, dw, db, None, None, None, None, None


def logits_xent(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], targets: torch.Tensor,
                weights: torch.Tensor, unk_index: int = -1, trans_w: bool = False,
                keep_logits: bool = False):
    """Vocabulary projection + masked cross-entropy (decoders/autoregressive.py:288-316,450-459).

    Returns (xent [M], lse [M], argmax [M] int64, logits [M,V] or None)."""
    if (_xent16_enabled() and not trans_w and b is not None and _GEMM_BACKEND != lib.GEMM_SIMT
            and w.requires_grad and b.requires_grad):
        w_sink, b_sink = _sink(w), _sink(b)
        k, v = w.shape
        if (w_sink is not None and b_sink is not None and w_sink.is_contiguous()
                and b_sink.data_ptr() == w_sink.data_ptr() + 4 * k * v):
            return _LogitsXent16.apply(x, w, b, targets, weights, unk_index, keep_logits)
    return _LogitsXent.apply(x, w, b, targets, weights, unk_index, trans_w, keep_logits)


# ---------------------------------------------------------------------------
# K8 multi-head attention core
# ---------------------------------------------------------------------------
class _MHA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_mask, causal, heads):
        bsz, tq, d = q.shape
        tk = k.size(1)
        dh = d // heads
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        mask_c = key_mask.contiguous() if key_mask is not None else None
        out = torch.empty_like(q)
        probs = torch.empty(bsz, heads, tq, tk, device=q.device, dtype=torch.float32)
        call("nm_mha_fwd", ptr(q), ptr(k), ptr(v), ptr(mask_c), int(causal), ptr(out), ptr(probs), bsz,
             tq, tk, heads, dh, lib.stream())
        ctx.save_for_backward(q, k, v, mask_c, probs)
        ctx.cfg = (causal, heads)
        ctx.mark_non_differentiable(probs)
        return out, probs

    @staticmethod
    def backward(ctx, dout, _dprobs):
        q, k, v, mask, probs = ctx.saved_tensors
        causal, heads = ctx.cfg
        bsz, tq, d = q.shape
        tk = k.size(1)
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        work = torch.empty_like(probs)
        call("nm_mha_bwd", ptr(q), ptr(k), ptr(v), ptr(mask), int(causal), ptr(probs), ptr(dout),
             ptr(dq), ptr(dk), ptr(dv), ptr(work), bsz, tq, tk, heads, d // heads, lib.stream())
        return dq, dk, dv, None, None, None


class _MHADrop(torch.autograd.Function):
    """_MHA with attention-weight dropout: `drop_mask` [B, heads, Tq, Tk] holds 0 or 1/keep_prob."""

    @staticmethod
    def forward(ctx, q, k, v, key_mask, causal, heads, drop_mask):
        bsz, tq, d = q.shape
        tk = k.size(1)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        mask_c = key_mask.contiguous() if key_mask is not None else None
        drop_c = drop_mask.to(torch.float32).contiguous()
        out = torch.empty_like(q)
        probs = torch.empty(bsz, heads, tq, tk, device=q.device, dtype=torch.float32)
        call("nm_mha_fwd_drop", ptr(q), ptr(k), ptr(v), ptr(mask_c), int(causal), ptr(drop_c), ptr(out),
             ptr(probs), bsz, tq, tk, heads, d // heads, lib.stream())
        ctx.save_for_backward(q, k, v, mask_c, probs, drop_c)
        ctx.cfg = (causal, heads)
        ctx.mark_non_differentiable(probs)
        return out, probs

    @staticmethod
    def backward(ctx, dout, _dprobs):
        q, k, v, mask, probs, drop_c = ctx.saved_tensors
        causal, heads = ctx.cfg
        bsz, tq, d = q.shape
        tk = k.size(1)
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        work = torch.empty_like(probs)
        call("nm_mha_bwd_drop", ptr(q), ptr(k), ptr(v), ptr(mask), int(causal), ptr(drop_c), ptr(probs),
             ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(work), bsz, tq, tk, heads, d // heads, lib.stream())
        return dq, dk, dv, None, None, None, None


class _MHATensorCore(torch.autograd.Function):
    """The attention core as batched tcgen05 products (csrc/mha_tc.cu): softmax and its backward in the GEMM
    epilogues, all (sentence, head) pairs in one launch per product.  TF32 operands."""

    @staticmethod
    def forward(ctx, q, k, v, key_mask, causal, heads, drop_mask):
        bsz, tq, d = q.shape
        tk = k.size(1)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        mask_c = key_mask.contiguous() if key_mask is not None else None
        drop_c = drop_mask.to(torch.float32).contiguous() if drop_mask is not None else None
        tqp, tkp = (tq + 31) // 32 * 32, (tk + 31) // 32 * 32
        out = torch.empty_like(q)
        probs = torch.empty(bsz, heads, tqp, tkp, device=q.device, dtype=torch.float32)
        probs_drop = torch.empty_like(probs) if drop_c is not None else None
        call("nm_mha_tc_fwd", ptr(q), ptr(k), ptr(v), ptr(mask_c), int(causal), ptr(drop_c), ptr(out), ptr(probs),
             ptr(probs_drop), bsz, tq, tk, heads, d // heads, lib.stream())
        ctx.save_for_backward(q, k, v, mask_c, probs, probs_drop, drop_c)
        ctx.cfg = (causal, heads)
        weights = probs[:, :, :tq, :tk]
        ctx.mark_non_differentiable(weights)
        return out, weights

    @staticmethod
    def backward(ctx, dout, _dprobs):
        q, k, v, mask, probs, probs_drop, drop_c = ctx.saved_tensors
        causal, heads = ctx.cfg
        bsz, tq, d = q.shape
        tk = k.size(1)
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        work = torch.empty_like(probs)
        call("nm_mha_tc_bwd", ptr(q), ptr(k), ptr(v), ptr(mask), int(causal), ptr(drop_c), ptr(probs),
             ptr(probs_drop), ptr(dout), ptr(dq), ptr(dk), ptr(dv), ptr(work), bsz, tq, tk, heads, d // heads,
             lib.stream())
        return dq, dk, dv, None, None, None, None


_MHA_TC_DEFAULT = "1"     # verified in tests/test_gpu_mha_tc.py; NMB200_MHA_TC=0 keeps the CUDA-core kernels


def _mha_on_tensor_cores(bsz: int, tq: int, tk: int, heads: int, dh: int) -> bool:
    """Tensor-core attention follows the GEMM backend ('simt' = the exact fp32 kernels everywhere); whole
    sequences only - the single-query steps of the decoding loops stay on the row kernels.  NMB200_MHA_TC=0
    switches it off."""
    if _GEMM_BACKEND == lib.GEMM_SIMT or os.environ.get("NMB200_MHA_TC", _MHA_TC_DEFAULT) == "0" or tq < 8:
        return False
    return bool(lib.load().nm_mha_tc_supported(bsz, tq, tk, heads, dh))


def mha_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_mask: Optional[torch.Tensor],
             causal: bool, heads: int, drop_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax(mask(q/sqrt(dh) k^T)) v per head (attention/scaled_dot_product.py:184-214).  With
    `drop_mask` ([B, heads, Tq, Tk], 0 or 1/keep_prob) the context is (softmax * drop_mask) v; the
    returned weights are the undropped softmax."""
    if q.is_cuda and _mha_on_tensor_cores(q.shape[0], q.shape[1], k.shape[1], heads, q.shape[2] // heads):
        return _MHATensorCore.apply(q, k, v, key_mask, causal, heads, drop_mask)
    if drop_mask is not None:
        return _MHADrop.apply(q, k, v, key_mask, causal, heads, drop_mask)
    return _MHA.apply(q, k, v, key_mask, causal, heads)


# ---------------------------------------------------------------------------
# non-differentiable helpers
# ---------------------------------------------------------------------------
def xent_rows(logits: torch.Tensor, targets: Optional[torch.Tensor] = None,
              weights: Optional[torch.Tensor] = None, want_argmax: bool = False, first_col: int = 0):
    """Row statistics of materialised logits [M, V] (contiguous rows) over columns first_col..V-1:
    (lse [M], weighted xent [M] or None without targets, first-index argmax [M] int64 relative to
    first_col or None).  One `nm_xent_fwd` launch; the column offset is pointer arithmetic with ld = V."""
    m, v = logits.shape
    dev = logits.device
    lse = torch.empty(m, device=dev, dtype=torch.float32)
    xent = torch.empty(m, device=dev, dtype=torch.float32) if targets is not None else None
    arg = torch.empty(m, device=dev, dtype=torch.int64) if want_argmax else None
    call("nm_xent_fwd", ptr(logits) + 4 * first_col, ptr(targets), ptr(weights), ptr(lse), ptr(xent),
         ptr(arg), m, v - first_col, v, lib.stream())
    return lse, xent, arg


def log_softmax_from_lse(logits: torch.Tensor, lse: torch.Tensor) -> torch.Tensor:
    m, v = logits.shape
    out = torch.empty_like(logits)
    call("nm_log_softmax", ptr(logits), ptr(lse), ptr(out), m, v, logits.stride(0), lib.stream())
    return out


def beam_step(logprobs: torch.Tensor, logprob_sum: torch.Tensor, lengths: torch.Tensor,
              finished: torch.Tensor, alpha: float):
    """One BeamSearchDecoder step (beam_search_decoder.py:440-496).  Returns
    (scores, word_ids i64, beam_ids i32, logprob_sum', lengths' i32, finished' u8)."""
    bsz, k, v = logprobs.shape
    dev = logprobs.device
    scores = torch.empty(bsz, k, device=dev, dtype=torch.float32)
    words = torch.empty(bsz, k, device=dev, dtype=torch.int64)
    beams = torch.empty(bsz, k, device=dev, dtype=torch.int32)
    lsum = torch.empty(bsz, k, device=dev, dtype=torch.float32)
    lens = torch.empty(bsz, k, device=dev, dtype=torch.int32)
    fin = torch.empty(bsz, k, device=dev, dtype=torch.uint8)
    scratch = torch.empty(lib.load().nm_beam_scratch(bsz, k, v), device=dev, dtype=torch.int32)
    call("nm_beam_step", ptr(logprobs.contiguous()), ptr(logprob_sum.contiguous()),
         ptr(lengths.contiguous()), ptr(finished.contiguous()), float(alpha), ptr(scores), ptr(words),
         ptr(beams), ptr(lsum), ptr(lens), ptr(fin), ptr(scratch), bsz, k, v, lib.stream())
    return scores, words, beams, lsum, lens, fin


def beam_gather(x: torch.Tensor, beam_ids: torch.Tensor, bsz: int, k: int) -> torch.Tensor:
    """gather_flat (tf_utils.py:106-131) on a [B*k, ...] tensor."""
    x = x.contiguous()
    out = torch.empty_like(x)
    row_bytes = (x.numel() // (bsz * k)) * x.element_size()
    call("nm_beam_gather", ptr(x), ptr(beam_ids.contiguous()), ptr(out), bsz, k, row_bytes,
         lib.stream())
    return out


# ---------------------------------------------------------------------------
# K12 frozen VGG stack primitives (forward only)
# ---------------------------------------------------------------------------
def conv3x3_bias_relu(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """NHWC 3x3 SAME conv + bias + ReLU; w is HWIO (slim vgg_arg_scope)."""
    n, h, wd, cin = x.shape
    cout = w.shape[-1]
    x, w, b = _f32(x.detach()).contiguous(), _f32(w.detach()).contiguous(), _f32(b.detach())
    y = torch.empty(n, h, wd, cout, device=x.device, dtype=torch.float32)
    if _GEMM_BACKEND != lib.GEMM_SIMT and cout % 4 == 0:
        # tensor-core path: patch matrix (im2col) x filter matrix [9*Cin, Cout] through the tcgen05
        # GEMM with the bias + ReLU epilogue; the output rows ARE the NHWC pixels.  The patch matrix
        # is built for a few images at a time so it stays below ~768 MB.
        k = 9 * cin
        ld = (k + 3) // 4 * 4
        per_image = h * wd * ld * 4
        chunk = max(1, min(n, (768 << 20) // per_image))
        cols = torch.zeros(chunk * h * wd, ld, device=x.device, dtype=torch.float32)
        w2 = w.view(k, cout)
        y2 = y.view(n * h * wd, cout)
        for n0 in range(0, n, chunk):
            nb = min(chunk, n - n0)
            call("nm_im2col3x3", ptr(x[n0:n0 + nb]), ptr(cols), nb, h, wd, cin, ld, lib.stream())
            gemm(cols[:nb * h * wd, :k], w2, y2[n0 * h * wd:(n0 + nb) * h * wd], bias=b, act="relu")
        return y
    call("nm_conv3x3_bias_relu_fwd", ptr(x), ptr(w), ptr(b), ptr(y), n, h, wd, cin, cout, lib.stream())
    return y


def maxpool2x2(x: torch.Tensor) -> torch.Tensor:
    n, h, wd, c = x.shape
    x = _f32(x.detach()).contiguous()
    y = torch.empty(n, h // 2, wd // 2, c, device=x.device, dtype=torch.float32)
    call("nm_maxpool2x2_fwd", ptr(x), ptr(y), n, h, wd, c, lib.stream())
    return y
