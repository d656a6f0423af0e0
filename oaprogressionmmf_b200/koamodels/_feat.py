"""``FeaT`` / ``Transformer`` / ``Attention`` / ``FeedForward`` with the reference's module tree and
``state_dict`` keys (``koafusion/models/_core_trf.py:74-205``); ``FeaT.forward`` is one call into
``koa_feat_forward`` (tcgen05 GEMMs + fused attention/LayerNorm kernels), backward one call into
``koa_feat_backward``: inside ``FeaT`` the sub-modules only hold the parameters, the engine owns the compute. Used on
their own (the reference exports them, ``koafusion/models/__init__.py:1``), ``Transformer`` / ``Attention`` /
``FeedForward`` run operator by operator through ``_ops`` (same arithmetic, one launch per GEMM / LayerNorm / attention).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib, dataparallel
from . import _ops


class FeedForward(nn.Module):
    """Linear -> GELU -> Dropout -> Linear -> Dropout (``_core_trf.py:141-153``)."""

    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden_dim, dim),
                                 nn.Dropout(dropout))

    def forward(self, x):
        h = nn.functional.gelu(_ops.linear(x, self.net[0].weight, self.net[0].bias))
        return self.net[4](_ops.linear(self.net[2](h), self.net[3].weight, self.net[3].bias))


class Attention(nn.Module):
    """bias-free ``to_qkv`` + ``to_out`` with ``scale = dim ** -0.5`` (``_core_trf.py:156-182``)."""

    def __init__(self, dim, heads=8, dropout=0.0):
        super().__init__()
        self.heads = heads
        self.scale = dim ** -0.5
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(dropout))

    def forward(self, x, mask=None):
        """``(out (B, n, dim), attn (B, heads, n, n))`` as the reference returns them. ``mask`` is rejected: the reference's
        mask branch cannot run (``import torch.functional as F`` has no ``pad``, ``_core_trf.py:2,173``)."""
        if mask is not None:
            raise ValueError("mask is unsupported (the reference's mask branch is dead code, _core_trf.py:172-177)")
        out, attn = _ops.attention(_ops.linear(x, self.to_qkv.weight), self.heads, self.scale)
        return self.to_out[1](_ops.linear(out, self.to_out[0].weight, self.to_out[0].bias)), attn


class Transformer(nn.Module):
    """``depth`` pre-norm blocks named prenorm_0_d / attn_d / prenorm_1_d / ff_d (``_core_trf.py:185-205``)."""

    def __init__(self, dim, depth, heads, mlp_dim, dropout):
        super().__init__()
        self.depth = depth
        self.dim = dim
        self.heads = heads
        self.mlp_dim = mlp_dim
        self.dropout = dropout
        for d in range(depth):
            setattr(self, f"prenorm_0_{d}", nn.LayerNorm(dim))
            setattr(self, f"attn_{d}", Attention(dim, heads=heads, dropout=dropout))
            setattr(self, f"prenorm_1_{d}", nn.LayerNorm(dim))
            setattr(self, f"ff_{d}", FeedForward(dim, mlp_dim, dropout=dropout))

    def layer_params(self, d):
        ln0, attn = getattr(self, f"prenorm_0_{d}"), getattr(self, f"attn_{d}")
        ln1, ff = getattr(self, f"prenorm_1_{d}"), getattr(self, f"ff_{d}")
        return [ln0.weight, ln0.bias, attn.to_qkv.weight, attn.to_out[0].weight, attn.to_out[0].bias, ln1.weight, ln1.bias,
                ff.net[0].weight, ff.net[0].bias, ff.net[3].weight, ff.net[3].bias]

    def forward(self, x, mask=None):
        """``(x, attentions)``: pre-norm blocks without a final norm (``_core_trf.py:195-205``)."""
        attentions = []
        for d in range(self.depth):
            o, attn = getattr(self, f"attn_{d}")(_ops.layer_norm(x, getattr(self, f"prenorm_0_{d}")), mask)
            attentions.append(attn)
            x = o + x
            x = getattr(self, f"ff_{d}")(_ops.layer_norm(x, getattr(self, f"prenorm_1_{d}"))) + x
        return x, attentions


class _FeaTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "FeaT", tokens: torch.Tensor, compute_head: bool, need_bw: bool, seed: int, *params):
        lib = _lib.load()
        dev = _lib.require_same_device("FeaT", tokens, *params)
        b, n_p, dim = tokens.shape
        tr = mod.transformer
        desc = _lib.FeatDesc(batch=b, n_patches=n_p, dim=dim, depth=tr.depth, heads=tr.heads, mlp_dim=tr.mlp_dim,
                             num_classes=mod.num_classes, with_cls=1 if mod.with_cls else 0,
                             compute_head=1 if compute_head else 0, training=1 if mod.training else 0,
                             need_backward=1 if need_bw else 0, emb_dropout=float(mod._p_emb),
                             mlp_dropout=float(mod._p_mlp), seed=seed)
        nbytes = lib.koa_feat_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            _lib.check(-1, "koa_feat_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=tokens.device)
        n = n_p + (1 if mod.with_cls else 0)
        states = torch.empty((b, n, dim), dtype=torch.float32, device=tokens.device)
        logits = torch.empty((b, mod.num_classes), dtype=torch.float32, device=tokens.device) if compute_head else None
        table = _lib.ptr_table(mod._param_list())
        tokens = tokens.contiguous().float()
        with _lib.on_device(dev):
            _lib.check(lib.koa_feat_forward(C.byref(desc), table, tokens.data_ptr(), ws.data_ptr(), states.data_ptr(),
                                            None if logits is None else logits.data_ptr(), _lib.current_stream()),
                       "koa_feat_forward")
        ctx.mod, ctx.desc, ctx.ws, ctx.table = mod, desc, ws, table
        ctx.tokens_need_grad = ctx.needs_input_grad[1]
        ctx.token_shape = tokens.shape
        # attention maps [B, H, n, n] per layer (the reference returns them; every caller drops them)
        attns = []
        off, nb = C.c_size_t(), C.c_size_t()
        for layer in range(tr.depth):
            _lib.check(lib.koa_feat_probs_offset(C.byref(desc), layer, C.byref(off), C.byref(nb)), "koa_feat_probs_offset")
            attns.append(ws[off.value:off.value + nb.value].view(torch.float32).view(b, tr.heads, n, n))
        ctx.mark_non_differentiable(*attns)
        if logits is None:
            logits = torch.zeros((b, mod.num_classes), dtype=torch.float32, device=tokens.device)
            ctx.mark_non_differentiable(logits)
        return (logits, states, *attns)

    @staticmethod
    def backward(ctx, d_logits, d_states, *_):
        lib = _lib.load()
        mod = ctx.mod
        params = mod._param_list()
        # every table slot the engine accumulates into must exist (also for frozen parameters); the dead head of the
        # per-sequence transformers gets no slot at all: the reference leaves .grad = None there
        n_head = 6
        live = [None if p is None else p for p in params]
        if not ctx.desc.compute_head:
            live[-n_head:] = [None] * n_head
            d_logits = None
        gfull, flat = _lib.zeros_like_flat(live)
        grads = [g if (g is not None and p.requires_grad) else None for g, p in zip(gfull, params)]
        gtable = _lib.ptr_table(gfull)
        d_tokens = torch.empty(ctx.token_shape, dtype=torch.float32, device=ctx.ws.device) if ctx.tokens_need_grad else None
        ds = None if d_states is None else d_states.contiguous().float()
        dl = None if d_logits is None else d_logits.contiguous().float()
        dev = _lib.require_same_device("FeaT backward", ctx.ws, ds, dl)
        with _lib.on_device(dev):
            _lib.check(lib.koa_feat_backward(C.byref(ctx.desc), ctx.table, gtable, ctx.ws.data_ptr(),
                                             None if ds is None else ds.data_ptr(), None if dl is None else dl.data_ptr(),
                                             None if d_tokens is None else d_tokens.data_ptr(), _lib.current_stream()),
                       "koa_feat_backward")
        ctx.ws = None
        dataparallel.sync_flat(flat, [p for g, p in zip(grads, params) if g is not None])
        out_grads = [g for g, p in zip(grads, params) if p is not None]
        return (None, d_tokens, None, None, None, *out_grads)


class FeaT(nn.Module):
    """Token transformer used for slice aggregation and cross-modal fusion (``_core_trf.py:74-138``).
    ``forward(features, mask=None) -> (outputs (B, 1, classes), states (B, n, D), attentions)``."""

    def __init__(self, num_patches, patch_dim, emb_dim, depth, heads, mlp_dim, num_classes, emb_dropout=0.0,
                 with_cls=True, num_cls_tokens=1, mlp_dropout=0.0, num_outputs=1):
        super().__init__()
        if num_cls_tokens != 1 or num_outputs != 1:
            raise ValueError("the B200 path implements num_cls_tokens = num_outputs = 1 (all koafusion models)")
        if patch_dim != emb_dim:
            raise ValueError("the B200 path implements patch_dim == emb_dim (hard-wired in every koafusion model)")
        self.patch_dim = patch_dim
        self.num_outputs = num_outputs
        self.num_classes = num_classes
        self.with_cls = with_cls
        if self.with_cls:
            self.cls_token = nn.Parameter(torch.randn(1, num_cls_tokens, emb_dim))
        else:
            num_cls_tokens = 0
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + num_cls_tokens, emb_dim))
        self.patch_to_embedding = nn.Linear(self.patch_dim, emb_dim)
        self.emb_dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(emb_dim, depth, heads, mlp_dim, mlp_dropout)
        self.to_cls_token = nn.Identity()
        self.mlp_head0 = nn.Sequential(nn.LayerNorm(emb_dim), nn.Linear(emb_dim, mlp_dim), nn.GELU(),
                                       nn.Dropout(mlp_dropout), nn.Linear(mlp_dim, num_classes))
        self._p_emb = emb_dropout
        self._p_mlp = mlp_dropout

    def _param_list(self):
        """Engine table order (feat_engine.cu): cls, pos, embed w/b, 11 per layer, 6 head entries."""
        p = [self.cls_token if self.with_cls else None, self.pos_embedding, self.patch_to_embedding.weight,
             self.patch_to_embedding.bias]
        for d in range(self.transformer.depth):
            p += self.transformer.layer_params(d)
        h = self.mlp_head0
        p += [h[0].weight, h[0].bias, h[1].weight, h[1].bias, h[4].weight, h[4].bias]
        for t in p:
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda):
                raise _lib.KoaError("FeaT parameters must be contiguous fp32 CUDA tensors")
        return p

    def run(self, features: torch.Tensor, compute_head: bool = True):
        # nn.Dropout inside FeaT (emb_dropout, mlp_dropout) runs inside the engine with counter-based Philox masks;
        # the seed of this call comes from torch's CPU generator, so torch.manual_seed makes a run reproducible
        seed = 0
        if self.training and (self._p_emb or self._p_mlp):
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        self.last_dropout_seed = seed
        params = self._param_list()
        live = [p for p in params if p is not None]
        need_bw = torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in live))
        out = _FeaTFunction.apply(self, features, compute_head, need_bw, seed, *live)
        logits, states, attns = out[0], out[1], list(out[2:])
        return logits[:, None, :], states, attns

    def forward(self, features, mask=None):
        if mask is not None:
            raise ValueError("mask is unsupported (the reference's mask branch is dead code, _core_trf.py:172-177)")
        return self.run(features, compute_head=True)
