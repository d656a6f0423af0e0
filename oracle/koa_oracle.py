"""CPU oracle for the koafusion hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain, functional PyTorch-fp32 restatement of what the reference's ``koafusion.models`` classes
compute, driven by a flat ``state_dict`` with exactly the reference's keys. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module, and only as the checker / reported CPU baseline. The product path
(``oaprogressionmmf_b200``) never imports it and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so this restatement is
pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified
``koafusion.models`` from /root/reference (available only in the build container), runs every
model class on seeded weights/inputs, and commits logits / loss / per-parameter gradient norms /
BatchNorm running statistics under ``tests/golden/``. ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ------------------------------------------------------------------------------------------------
# Feature-extractor architecture tables (koafusion/models/_torchvision.py:265-330 and the torchvision
# twins selected by koafusion/models/_core_fes.py:6-15).
# ------------------------------------------------------------------------------------------------
ARCHS = {
    "resnet18": dict(block="basic", layers=(2, 2, 2, 2), groups=1, width_per_group=64),
    "resnet34": dict(block="basic", layers=(3, 4, 6, 3), groups=1, width_per_group=64),
    "resnet50": dict(block="bottleneck", layers=(3, 4, 6, 3), groups=1, width_per_group=64),
    "resnext50_32x4d": dict(block="bottleneck", layers=(3, 4, 6, 3), groups=32, width_per_group=4),
}
FE_OUT_CH = {"resnet18": 512, "resnet34": 512, "resnet50": 2048, "resnext50_32x4d": 2048}


def fe_block_plan(arch: str) -> List[dict]:
    """Per-block geometry in execution order (``ResNet._make_layer``, _torchvision.py:202-225)."""
    a = ARCHS[arch]
    expansion = 4 if a["block"] == "bottleneck" else 1
    inplanes = 64
    plan = []
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256, 512), a["layers"])):
        for bi in range(nblocks):
            stride = 2 if (li > 0 and bi == 0) else 1
            width = int(planes * (a["width_per_group"] / 64.0)) * a["groups"] if a["block"] == "bottleneck" else planes
            down = bi == 0 and (stride != 1 or inplanes != planes * expansion)
            plan.append(dict(layer=li + 4, index=bi, inplanes=inplanes, planes=planes, width=width,
                             out=planes * expansion, stride=stride, groups=a["groups"], downsample=down,
                             kind=a["block"]))
            inplanes = planes * expansion
    return plan


def fe_param_spec(arch: str, prefix: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys/shapes of ``nn.Sequential(*list(resnet.children())[:-1])`` under ``prefix``."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(name, c):
        spec.extend([(f"{name}.weight", (c,)), (f"{name}.bias", (c,)), (f"{name}.running_mean", (c,)),
                     (f"{name}.running_var", (c,)), (f"{name}.num_batches_tracked", ())])

    spec.append((f"{prefix}.0.weight", (64, 3, 7, 7)))
    bn(f"{prefix}.1", 64)
    for b in fe_block_plan(arch):
        p = f"{prefix}.{b['layer']}.{b['index']}"
        if b["kind"] == "bottleneck":
            spec.append((f"{p}.conv1.weight", (b["width"], b["inplanes"], 1, 1)))
            bn(f"{p}.bn1", b["width"])
            spec.append((f"{p}.conv2.weight", (b["width"], b["width"] // b["groups"], 3, 3)))
            bn(f"{p}.bn2", b["width"])
            spec.append((f"{p}.conv3.weight", (b["out"], b["width"], 1, 1)))
            bn(f"{p}.bn3", b["out"])
        else:
            spec.append((f"{p}.conv1.weight", (b["planes"], b["inplanes"], 3, 3)))
            bn(f"{p}.bn1", b["planes"])
            spec.append((f"{p}.conv2.weight", (b["planes"], b["planes"], 3, 3)))
            bn(f"{p}.bn2", b["planes"])
        if b["downsample"]:
            spec.append((f"{p}.downsample.0.weight", (b["out"], b["inplanes"], 1, 1)))
            bn(f"{p}.downsample.1", b["out"])
    return spec


def feat_param_spec(prefix: str, num_patches: int, dim: int, depth: int, mlp_dim: int, num_classes: int,
                    with_cls: bool) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict keys/shapes of ``FeaT`` (_core_trf.py:94-116,185-193) in registration order."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    n_tok = num_patches + (1 if with_cls else 0)
    if with_cls:
        spec.append((f"{prefix}.cls_token", (1, 1, dim)))
    spec.append((f"{prefix}.pos_embedding", (1, n_tok, dim)))
    spec += [(f"{prefix}.patch_to_embedding.weight", (dim, dim)), (f"{prefix}.patch_to_embedding.bias", (dim,))]
    t = f"{prefix}.transformer"
    for d in range(depth):
        spec += [(f"{t}.prenorm_0_{d}.weight", (dim,)), (f"{t}.prenorm_0_{d}.bias", (dim,)),
                 (f"{t}.attn_{d}.to_qkv.weight", (3 * dim, dim)),
                 (f"{t}.attn_{d}.to_out.0.weight", (dim, dim)), (f"{t}.attn_{d}.to_out.0.bias", (dim,)),
                 (f"{t}.prenorm_1_{d}.weight", (dim,)), (f"{t}.prenorm_1_{d}.bias", (dim,)),
                 (f"{t}.ff_{d}.net.0.weight", (mlp_dim, dim)), (f"{t}.ff_{d}.net.0.bias", (mlp_dim,)),
                 (f"{t}.ff_{d}.net.3.weight", (dim, mlp_dim)), (f"{t}.ff_{d}.net.3.bias", (dim,))]
    h = f"{prefix}.mlp_head0"
    spec += [(f"{h}.0.weight", (dim,)), (f"{h}.0.bias", (dim,)), (f"{h}.1.weight", (mlp_dim, dim)),
             (f"{h}.1.bias", (mlp_dim,)), (f"{h}.4.weight", (num_classes, mlp_dim)), (f"{h}.4.bias", (num_classes,))]
    return spec


# ------------------------------------------------------------------------------------------------
# Model descriptions: one plain dict per reference class (+ the 3-MRI pattern extensions).
# ------------------------------------------------------------------------------------------------
MODEL_NAMES = ("XR1Cnn", "MR1CnnTrf", "MR2CnnTrf", "XR1MR1CnnTrf", "XR1MR2CnnTrf", "XR1MR2C1CnnTrf",
               "MR3CnnTrf", "XR1MR3C1CnnTrf")


def model_param_spec(name: str, cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list equal to ``dict_models[name](cfg, None).state_dict()`` of the
    reference (koafusion/models/__init__.py:8-15). ``cfg`` is the model config as a plain dict."""
    nc = cfg["output_channels"]
    if name == "XR1Cnn":  # _xr1_cnn.py:15-39
        arch = cfg["fe"]["arch"]
        c = FE_OUT_CH[arch]
        hid = cfg["agg"]["hidden_size"]
        return fe_param_spec(arch, "_fe") + [("_agg.1.weight", (hid, c)), ("_agg.1.bias", (hid,)),
                                             ("_final.weight", (nc, hid)), ("_final.bias", (nc,))]
    agg = cfg["agg"]
    depth, mlp = agg["depth"], agg["mlp_dim"]
    if name == "MR1CnnTrf":  # _mrN_cnn_trf.py:20-89 (with_gap: one token per slice)
        arch = cfg["fe"]["arch"]
        c = FE_OUT_CH[arch]
        n = _mr1_num_tokens(cfg)
        return fe_param_spec(arch, "_fe") + feat_param_spec("_agg", n, c, depth, mlp, nc, True)
    if name == "MR2CnnTrf":  # _mrN_cnn_trf.py:150-213
        arch = cfg["fe"]["arch"]
        c = FE_OUT_CH[arch]
        n = agg["num_slices"][0] + agg["num_slices"][1]
        return (fe_param_spec(arch, "_fe0") + fe_param_spec(arch, "_fe1") +
                feat_param_spec("_agg", n, c, depth, mlp, nc, True))
    xr, mr = cfg["fe"]["xr"]["arch"], cfg["fe"]["mr"]["arch"]
    c = FE_OUT_CH[mr]
    ns = agg["num_slices"]
    p0, p1 = (1, 1)
    if name in ("XR1MR1CnnTrf", "XR1MR2CnnTrf", "XR1MR2C1CnnTrf"):
        p0, p1, _ = _xrmr_positions(cfg)
    if name == "XR1MR1CnnTrf":  # _xr1mrN.py:19-100
        return (fe_param_spec(xr, "_fe0") + fe_param_spec(mr, "_fe1") +
                feat_param_spec("_agg", p0 + ns[1] * p1, c, depth, mlp, nc, True))
    if name == "XR1MR2CnnTrf":  # _xr1mrN.py:169-296
        return (fe_param_spec(xr, "_fe0") + fe_param_spec(mr, "_fe1") + fe_param_spec(mr, "_fe2") +
                feat_param_spec("_agg_1", ns[1] * p1, c, depth, mlp, nc, False) +
                feat_param_spec("_agg_2", ns[2] * p1, c, depth, mlp, nc, False) +
                feat_param_spec("_agg_final", p0 + (ns[1] + ns[2]) * p1, c, depth, mlp, nc, True))
    if name == "XR1MR2C1CnnTrf":  # _xrNmrMcP.py:40-179
        clin = cfg["fe"]["clin"]
        return (fe_param_spec(xr, "_fe0") + fe_param_spec(mr, "_fe1") + fe_param_spec(mr, "_fe2") +
                [("_fe3._fe.0.weight", (clin["dim_out"], clin["dim_in"])), ("_fe3._fe.0.bias", (clin["dim_out"],))] +
                feat_param_spec("_agg_1", ns[1] * p1, c, depth, mlp, nc, False) +
                feat_param_spec("_agg_2", ns[2] * p1, c, depth, mlp, nc, False) +
                feat_param_spec("_agg_final", p0 + (ns[1] + ns[2]) * p1 + ns[3], c, depth, mlp, nc, True))
    # ---- pattern extensions (not in the reference; SURVEY.md §8 row a-ext) -----------------------
    if name == "MR3CnnTrf":  # three sequences, hierarchical like _xrNmrMcP.py without XR / clin
        return (fe_param_spec(mr, "_fe1") + fe_param_spec(mr, "_fe2") + fe_param_spec(mr, "_fe3") +
                feat_param_spec("_agg_1", ns[0], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_2", ns[1], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_3", ns[2], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_final", ns[0] + ns[1] + ns[2], c, depth, mlp, nc, True))
    if name == "XR1MR3C1CnnTrf":
        clin = cfg["fe"]["clin"]
        return (fe_param_spec(xr, "_fe0") + fe_param_spec(mr, "_fe1") + fe_param_spec(mr, "_fe2") +
                fe_param_spec(mr, "_fe3") +
                [("_fe4._fe.0.weight", (clin["dim_out"], clin["dim_in"])), ("_fe4._fe.0.bias", (clin["dim_out"],))] +
                feat_param_spec("_agg_1", ns[1], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_2", ns[2], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_3", ns[3], c, depth, mlp, nc, False) +
                feat_param_spec("_agg_final", 1 + ns[1] + ns[2] + ns[3] + ns[4], c, depth, mlp, nc, True))
    raise ValueError(f"unknown model {name}")


# spatial size of the extractor output without GAP, as the reference tabulates it (_mrN_cnn_trf.py:56, _xrNmrMcP.py:104-105)
_SPATIAL = {320: 10, 160: 5, 128: 4, 96: 3, 64: 2, 32: 1, 350: 11, 25: 1}


def _mr1_num_tokens(cfg: dict) -> int:
    """``vs['agg_in_len']`` of MR1CnnTrf (_mrN_cnn_trf.py:46-71)."""
    shape = list(cfg["input_size"][0])
    if cfg["downscale"]:
        shape = [round(s * d) for s, d in zip(shape, cfg["downscale"][0])]
    sp = (1, 1, 1) if cfg["fe"]["with_gap"] else tuple(_SPATIAL[e] for e in shape)
    return {"rc": shape[2] * sp[0] * sp[1], "cs": shape[0] * sp[1] * sp[2], "rs": shape[1] * sp[0] * sp[2]}[cfg["fe"]["dims_view"]]


def _xrmr_positions(cfg: dict) -> Tuple[int, int, bool]:
    """(positions per XR image, positions per MRI slice, GAP kept) of the XR + MRI classes: the extractors keep their GAP
    when EITHER modality asks for it, while the token counts follow each modality's own flag (_xrNmrMcP.py:47-55,112-122);
    only the two consistent settings (both with GAP, both without) are meaningful."""
    fe = cfg["fe"]
    gap = bool(fe["xr"]["with_gap"] or fe["mr"]["with_gap"])
    assert gap == bool(fe["xr"]["with_gap"] and fe["mr"]["with_gap"]), "mixed with_gap settings are inconsistent in the reference"
    shapes = [list(t) for t in cfg["input_size"]]
    if cfg["downscale"]:
        shapes = [[round(a * d) for a, d in zip(t, dd)] for t, dd in zip(shapes, cfg["downscale"])]
    p0 = 1 if gap else _SPATIAL[shapes[0][0]] * _SPATIAL[shapes[0][1]]
    p1 = 1 if gap else _SPATIAL[shapes[1][0]] * _SPATIAL[shapes[1][1]]
    return p0, p1, gap


def make_state_dict(spec: Sequence[Tuple[str, Tuple[int, ...]]], seed: int, pos_scale: float = 1.0,
                    device: str = "cpu", res_gain: float = 1.0) -> "OrderedDict[str, Tensor]":
    """Seeded, well-conditioned parameter values for a (key, shape) spec. Values are drawn on the CPU
    in spec order from one generator so that the build container (reference + oracle) and the GPU
    box (oracle + CUDA path) materialise bit-identical weights from (spec, seed).
    ``res_gain`` scales the weight of the last BatchNorm of every residual branch (bn3 of a bottleneck, bn2
    of a basic block): < 1 gives the near-identity blocks of a trained / zero-init-residual network instead of
    the perturbation-amplifying dynamics of a freshly initialised BatchNorm ResNet."""
    last_bn: Dict[str, str] = {}
    for key, _ in spec:  # name of the last bnK of each block
        if ".bn" in key and key.endswith(".weight"):
            blk, name = key.rsplit(".bn", 1)
            last_bn[blk] = max(last_bn.get(blk, ""), name)
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for key, shape in spec:
        leaf = key.rsplit(".", 1)[-1]
        is_norm = (".bn" in key or ".downsample.1." in key or "prenorm_" in key or key.endswith("mlp_head0.0.weight")
                   or key.endswith("mlp_head0.0.bias") or (len(shape) == 1 and _is_stem_bn(key)))
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf in ("pos_embedding", "cls_token"):
            t = torch.randn(shape, generator=g) * pos_scale
        elif len(shape) == 4:  # conv, Kaiming fan-out like _torchvision.py:187
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif is_norm and leaf == "weight":
            t = torch.rand(shape, generator=g) + 0.5
            if res_gain != 1.0 and ".bn" in key:
                blk, name = key.rsplit(".bn", 1)
                if last_bn.get(blk) == name:
                    t = t * res_gain
        elif is_norm and leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 2:  # nn.Linear weight
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:  # nn.Linear bias
            t = torch.randn(shape, generator=g) * 0.02
        sd[key] = t.to(device)
    return sd


def _is_stem_bn(key: str) -> bool:
    parts = key.split(".")
    return len(parts) >= 3 and parts[-2] == "1" and parts[-3].startswith("_fe")


# ------------------------------------------------------------------------------------------------
# 16-bit storage emulation
# ------------------------------------------------------------------------------------------------
# The CUDA path stores the CNN's forward activations and GEMM weight operands in fp16 (11-bit significand; every
# value is O(1) behind a BatchNorm), activation gradients in bf16 (fp32 range, no loss scaling needed), and
# accumulates in fp32. With ``emulate_16bit=True`` the oracle rounds at exactly those storage points (values stay
# fp32 tensors), so a comparison against it isolates kernel bugs from the precision choice; the fp32 oracle
# (default) is the reference semantics. A randomly initialised BatchNorm ResNet in train mode amplifies
# perturbations from block to block, so the two oracles themselves differ by more than one rounding at the feature
# level (measured in tests/test_gpu_parity.py).
class _RoundAct(torch.autograd.Function):
    """fp16 round of an activation in forward, bf16 round of its gradient in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundWeight(torch.autograd.Function):
    """fp16 round of a GEMM weight operand; the gradient reaches the fp32 master weight unrounded."""

    @staticmethod
    def forward(ctx, w):
        return w.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _ra(x: Tensor, emulate: bool) -> Tensor:
    return _RoundAct.apply(x) if emulate else x


def _rw(w: Tensor, emulate: bool) -> Tensor:
    return _RoundWeight.apply(w) if emulate else w


# ------------------------------------------------------------------------------------------------
# Functional forward passes
# ------------------------------------------------------------------------------------------------
def _bn(sd: StateDict, p: str, x: Tensor, training: bool) -> Tensor:
    """nn.BatchNorm2d (eps 1e-5, momentum 0.1; _torchvision.py:109-113,172). In training mode the
    running statistics in ``sd`` are updated in place, as the module does."""
    y = F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"],
                     training, 0.1, 1e-5)
    if training:
        sd[f"{p}.num_batches_tracked"] += 1
    return y


def fe_forward(sd: StateDict, prefix: str, arch: str, x: Tensor, training: bool, with_gap: bool = True,
               taps: Dict[str, Tensor] | None = None, emulate_16bit: bool = False) -> Tensor:
    """``ResNet._forward_impl`` without ``fc`` (_torchvision.py:227-239), blocks per :64-80 / :118-138.
    ``x`` is (N, 3, H, W). ``taps`` (optional) collects block outputs for intermediate parity checks."""
    e = emulate_16bit

    def tap(key, t):
        if taps is not None:
            taps[key] = t
        return t

    def conv(inp, key, **kw):  # conv output is stored in bf16 by the CUDA path
        return tap(key[:-len(".weight")] + ".y", _ra(F.conv2d(inp, _rw(sd[key], e), **kw), e))

    if e:
        # CUDA path: the three identical input channels are folded into the weights (sum over dim 1), image and
        # folded weights are rounded to bf16 (operands of the stem's tensor-core GEMM), the output is stored in bf16
        wf = _rw(sd[f"{prefix}.0.weight"].sum(dim=1, keepdim=True), e)
        x = tap(f"{prefix}.0.y", _ra(F.conv2d(_ra(x[:, :1], e), wf, stride=2, padding=3), e))
    else:
        x = tap(f"{prefix}.0.y", F.conv2d(x, sd[f"{prefix}.0.weight"], stride=2, padding=3))
    x = tap(f"{prefix}.stem", _ra(F.relu(_bn(sd, f"{prefix}.1", x, training)), e))
    x = tap(f"{prefix}.pool", F.max_pool2d(x, kernel_size=3, stride=2, padding=1))
    for b in fe_block_plan(arch):
        p = f"{prefix}.{b['layer']}.{b['index']}"
        identity = x
        if b["kind"] == "bottleneck":
            o = tap(f"{p}.a1", _ra(F.relu(_bn(sd, f"{p}.bn1", conv(x, f"{p}.conv1.weight"), training)), e))
            o = conv(o, f"{p}.conv2.weight", stride=b["stride"], padding=1, groups=b["groups"])
            o = tap(f"{p}.a2", _ra(F.relu(_bn(sd, f"{p}.bn2", o, training)), e))
            o = _bn(sd, f"{p}.bn3", conv(o, f"{p}.conv3.weight"), training)
        else:
            o = conv(x, f"{p}.conv1.weight", stride=b["stride"], padding=1)
            o = tap(f"{p}.a1", _ra(F.relu(_bn(sd, f"{p}.bn1", o, training)), e))
            o = _bn(sd, f"{p}.bn2", conv(o, f"{p}.conv2.weight", padding=1), training)
        if b["downsample"]:
            identity = _bn(sd, f"{p}.downsample.1", conv(x, f"{p}.downsample.0.weight", stride=b["stride"]), training)
        x = tap(p, _ra(F.relu(o + identity), e))
    if with_gap:
        x = x.mean(dim=(2, 3), keepdim=True)
    return x


def _layernorm(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], 1e-5)


def _dropout(x: Tensor, p: float, training: bool, mask: Tensor | None = None) -> Tensor:
    """nn.Dropout. ``mask`` (scale factors 0 or 1/(1-p), same shape as x) replaces torch's random stream so that a
    test can feed the oracle the very mask the CUDA path drew (koa_dropout_mask)."""
    if mask is not None:
        return x * mask.reshape(x.shape)
    return F.dropout(x, p, training) if p else x


def _mask(masks, key):
    return None if masks is None else masks.get(key)


def attention_forward(sd: StateDict, p: str, x: Tensor, heads: int, dropout: float, training: bool,
                      mask: Tensor | None = None) -> Tuple[Tensor, Tensor]:
    """``Attention.forward`` (_core_trf.py:167-182): bias-free qkv, feature index = (qkv, head, d),
    scale = model_dim ** -0.5 (NOT head_dim), softmax over keys, output projection with bias."""
    b, n, dim = x.shape
    d = dim // heads
    qkv = F.linear(x, sd[f"{p}.to_qkv.weight"]).reshape(b, n, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * dim ** -0.5, dim=-1)
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(b, n, dim)
    out = F.linear(out, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])
    return _dropout(out, dropout, training, mask), attn


def transformer_forward(sd: StateDict, p: str, x: Tensor, depth: int, heads: int, dropout: float,
                        training: bool, masks: Dict[str, Tensor] | None = None) -> Tensor:
    """``Transformer.forward`` (_core_trf.py:195-205): pre-norm residual blocks, no final norm;
    ``FeedForward`` (:141-153) = Linear, exact GELU, Dropout, Linear, Dropout."""
    for d in range(depth):
        o, _ = attention_forward(sd, f"{p}.attn_{d}", _layernorm(sd, f"{p}.prenorm_0_{d}", x), heads, dropout, training,
                                 _mask(masks, f"attn_out_{d}"))
        x = o + x
        h = _layernorm(sd, f"{p}.prenorm_1_{d}", x)
        h = F.gelu(F.linear(h, sd[f"{p}.ff_{d}.net.0.weight"], sd[f"{p}.ff_{d}.net.0.bias"]))
        h = _dropout(h, dropout, training, _mask(masks, f"ff_act_{d}"))
        h = F.linear(h, sd[f"{p}.ff_{d}.net.3.weight"], sd[f"{p}.ff_{d}.net.3.bias"])
        x = _dropout(h, dropout, training, _mask(masks, f"ff_out_{d}")) + x
    return x


def feat_forward(sd: StateDict, p: str, tokens: Tensor, depth: int, heads: int, emb_dropout: float,
                 mlp_dropout: float, training: bool, masks: Dict[str, Tensor] | None = None) -> Tuple[Tensor, Tensor]:
    """``FeaT.forward`` (_core_trf.py:118-138) → (head output (B,1,C), token states (B,n,D))."""
    x = F.linear(tokens, sd[f"{p}.patch_to_embedding.weight"], sd[f"{p}.patch_to_embedding.bias"])
    if f"{p}.cls_token" in sd:
        x = torch.cat((sd[f"{p}.cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    x = x + sd[f"{p}.pos_embedding"]
    x = _dropout(x, emb_dropout, training, _mask(masks, "emb"))
    states = transformer_forward(sd, f"{p}.transformer", x, depth, heads, mlp_dropout, training, masks)
    h = _layernorm(sd, f"{p}.mlp_head0.0", states[:, 0])
    h = F.gelu(F.linear(h, sd[f"{p}.mlp_head0.1.weight"], sd[f"{p}.mlp_head0.1.bias"]))
    h = _dropout(h, mlp_dropout, training, _mask(masks, "head"))
    out = F.linear(h, sd[f"{p}.mlp_head0.4.weight"], sd[f"{p}.mlp_head0.4.bias"])
    return out[:, None, :], states


def _slices_to_images(vol: Tensor) -> Tensor:
    """(B,1,R,C,S) → (B*S,3,R,C): einops rearrange + channel repeat (_xrNmrMcP.py:209-213)."""
    b, ch, r, c, s = vol.shape
    return vol.permute(0, 4, 1, 2, 3).reshape(b * s, ch, r, c).expand(-1, 3, -1, -1)


def _fe_tokens(sd, prefix, arch, images, training, batch, drop_p, taps=None, emulate_16bit=False, with_gap=True):
    """FE → Dropout2d → tokens "(b s) ch d0 d1 -> b (s d0 d1) ch" (_xrNmrMcP.py:226-232; d0 = d1 = 1 with GAP)."""
    f = fe_forward(sd, prefix, arch, images, training, with_gap, taps, emulate_16bit=emulate_16bit)
    if drop_p:
        f = F.dropout2d(f, drop_p, training)
    n, ch = f.shape[0], f.shape[1]
    return f.reshape(batch, n // batch, ch, -1).permute(0, 1, 3, 2).reshape(batch, -1, ch)


def model_forward(name: str, cfg: dict, sd: StateDict, inputs: Sequence[Tensor], training: bool,
                  taps: Dict[str, Tensor] | None = None, emulate_16bit: bool = False) -> Tensor:
    """Logits (B, output_channels) of ``dict_models[name]`` for positional ``inputs``.
    ``emulate_16bit`` rounds the feature extractors' stored tensors to bf16 (see ``fe_forward``); it exists to
    measure the precision floor of bf16 storage with no CUDA-path code involved."""
    agg = cfg["agg"]
    _tok = globals()["_fe_tokens"]

    def _fe_tokens(*a, **k):  # noqa: F811 - thread the emulation flag through every extractor call below
        return _tok(*a, emulate_16bit=emulate_16bit, **k)

    if name == "XR1Cnn":  # _xr1_cnn.py:48-81
        x = inputs[0]
        f = fe_forward(sd, "_fe", cfg["fe"]["arch"], x.expand(-1, 3, -1, -1), training, True, taps,
                       emulate_16bit=emulate_16bit).flatten(1)
        f = _dropout(f, agg["dropout"], training)
        f = F.relu(F.linear(f, sd["_agg.1.weight"], sd["_agg.1.bias"]))
        f = _dropout(f, agg["dropout"], training)
        return F.linear(f, sd["_final.weight"], sd["_final.bias"])

    depth, heads = agg["depth"], agg["heads"]
    ed, md = agg["emb_dropout"], agg["mlp_dropout"]

    def feat(p, tok):
        return feat_forward(sd, p, tok, depth, heads, ed, md, training)

    if name == "MR1CnnTrf":  # _mrN_cnn_trf.py:98-139
        vol = inputs[0]
        b = vol.shape[0]
        view = cfg["fe"]["dims_view"]
        v3 = vol.expand(-1, 3, -1, -1, -1)
        perm = {"rc": (0, 4, 1, 2, 3), "cs": (0, 2, 1, 3, 4), "rs": (0, 3, 1, 2, 4)}[view]
        imgs = v3.permute(*perm)
        imgs = imgs.reshape(-1, *imgs.shape[2:])
        tok = _fe_tokens(sd, "_fe", cfg["fe"]["arch"], imgs, training, b, cfg["fe"]["dropout"], taps,
                         with_gap=cfg["fe"]["with_gap"])
        if taps is not None:
            taps["tokens"] = tok
        out, states = feat("_agg", tok)
        if taps is not None:
            taps["states"] = states
        return out.flatten(1)
    if name == "MR2CnnTrf":  # _mrN_cnn_trf.py:222-272
        b = inputs[0].shape[0]
        arch, dp = cfg["fe"]["arch"], cfg["fe"]["dropout"]
        t0 = _fe_tokens(sd, "_fe0", arch, _slices_to_images(inputs[0]), training, b, dp, taps)
        t1 = _fe_tokens(sd, "_fe1", arch, _slices_to_images(inputs[1]), training, b, dp, taps)
        return feat("_agg", torch.cat([t0, t1], dim=1))[0].flatten(1)

    xr_cfg, mr_cfg = cfg["fe"]["xr"], cfg["fe"]["mr"]
    if name == "MR3CnnTrf":
        b = inputs[0].shape[0]
        toks = [_fe_tokens(sd, f"_fe{i + 1}", mr_cfg["arch"], _slices_to_images(inputs[i]), training, b,
                           mr_cfg["dropout"], taps) for i in range(3)]
        st = [feat(f"_agg_{i + 1}", toks[i])[1] for i in range(3)]
        return feat("_agg_final", torch.cat(st, dim=1))[0].flatten(1)

    b = inputs[0].shape[0]
    gap = bool(xr_cfg["with_gap"] or mr_cfg["with_gap"])  # one flag for every extractor (_xrNmrMcP.py:47)
    t0 = _fe_tokens(sd, "_fe0", xr_cfg["arch"], inputs[0].expand(-1, 3, -1, -1), training, b, xr_cfg["dropout"], taps,
                    with_gap=gap)
    if name == "XR1MR1CnnTrf":  # _xr1mrN.py:108-158
        t1 = _fe_tokens(sd, "_fe1", mr_cfg["arch"], _slices_to_images(inputs[1]), training, b, mr_cfg["dropout"], taps,
                        with_gap=gap)
        return feat("_agg", torch.cat([t0, t1], dim=1))[0].flatten(1)

    n_mr = 3 if name == "XR1MR3C1CnnTrf" else 2
    toks = [_fe_tokens(sd, f"_fe{i + 1}", mr_cfg["arch"], _slices_to_images(inputs[i + 1]), training, b,
                       mr_cfg["dropout"], taps, with_gap=gap) for i in range(n_mr)]
    # per-sequence transformers: all token states are forwarded, their heads are dead compute
    # (_xrNmrMcP.py:239-240, _xr1mrN.py:347-348)
    states = [feat(f"_agg_{i + 1}", toks[i])[1] for i in range(n_mr)]
    parts = [t0] + states
    if taps is not None:
        taps["tokens_xr"] = t0
        for i in range(n_mr):
            taps[f"tokens_mr{i + 1}"] = toks[i]
            taps[f"states_mr{i + 1}"] = states[i]
    if name in ("XR1MR2C1CnnTrf", "XR1MR3C1CnnTrf"):  # FeatC1: Linear → GELU → Dropout (_xrNmrMcP.py:15-29)
        cp = f"_fe{n_mr + 1}._fe.0"
        clin = F.gelu(F.linear(inputs[n_mr + 1], sd[f"{cp}.weight"], sd[f"{cp}.bias"]))
        parts.append(_dropout(clin, cfg["fe"]["clin"]["dropout"], training))
    elif name != "XR1MR2CnnTrf":
        raise ValueError(f"unknown model {name}")
    return feat("_agg_final", torch.cat(parts, dim=1))[0].flatten(1)


def focal_loss(logits: Tensor, target: Tensor, gamma: float = 2.0) -> Tensor:
    """``FocalLoss(gamma=2, reduction='mean')`` (koafusion/various/_losses.py:89-108)."""
    logpt = -F.cross_entropy(logits, target, reduction="none")
    return (-((1 - torch.exp(logpt)) ** gamma) * logpt).mean()


def train_step(name: str, cfg: dict, sd: StateDict, inputs: Sequence[Tensor], target: Tensor,
               taps: Dict[str, Tensor] | None = None, emulate_16bit: bool = False):
    """zero_grad → forward(train) → FocalLoss → backward (koafusion/run/train_prog_fus.py:133-165).
    Returns (logits, loss, {key: grad or None}). BN running stats in ``sd`` are updated in place."""
    params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    for v in params.values():
        v.requires_grad_(True)
        v.grad = None
    logits = model_forward(name, cfg, sd, inputs, True, taps, emulate_16bit=emulate_16bit)
    loss = focal_loss(logits, target)
    loss.backward()
    grads = {k: v.grad for k, v in params.items()}
    for v in params.values():
        v.requires_grad_(False)
    return logits.detach(), loss.detach(), grads


# ------------------------------------------------------------------------------------------------
# Canonical configs (shapes after down-scaling) used by tests, bench and the golden generator
# ------------------------------------------------------------------------------------------------
def make_config(name: str, *, xr_size=350, mr_size=160, slices=(64, 32, 25), depth=4, heads=8, mlp_dim=2048,
                xr_arch="resnext50_32x4d", mr_arch="resnet50", dropout=0.0, clin_dim=9, dim_out=2048,
                output_type="dict", with_gap=True, dims_view="rc") -> dict:
    """Model config dicts with the keys the reference constructors read. ``slices``
    are the per-sequence slice counts in the order the class takes its MRI inputs."""
    base = dict(name=name, debug=False, downscale=False, input_channels=1, output_channels=2,
                output_type=output_type, pretrained=False, path_pretrained=None, restore_weights=False)
    agg = dict(depth=depth, heads=heads, emb_dropout=dropout, mlp_dim=mlp_dim, mlp_dropout=dropout)
    if name == "XR1Cnn":
        return dict(base, input_size=[[xr_size, xr_size]],
                    fe=dict(arch=xr_arch, pretrained=False, with_gap=True, dropout=0.0),
                    agg=dict(hidden_size=512, dropout=0.5 if dropout else 0.0))
    if name == "MR1CnnTrf":
        return dict(base, input_size=[[mr_size, mr_size, slices[0]]],
                    fe=dict(arch=mr_arch, pretrained=False, with_gap=with_gap, dropout=dropout, dims_view=dims_view),
                    agg=dict(agg, num_slices=slices[0]))
    if name == "MR2CnnTrf":
        return dict(base, input_size=[[mr_size, mr_size, slices[0]], [mr_size, mr_size, slices[1]]],
                    fe=dict(arch=mr_arch, pretrained=False, with_gap=True, dropout=dropout),
                    agg=dict(agg, num_slices=[slices[0], slices[1]]))
    fe = dict(xr=dict(arch=xr_arch, pretrained=False, with_gap=with_gap, dropout=dropout),
              mr=dict(arch=mr_arch, pretrained=False, with_gap=with_gap, dropout=dropout))
    if name == "XR1MR1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], [mr_size, mr_size, slices[0]]], fe=fe,
                    agg=dict(agg, num_slices=[1, slices[0]]))
    if name == "XR1MR2CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], [mr_size, mr_size, slices[0]], [mr_size, mr_size, slices[1]]],
                    fe=fe, agg=dict(agg, num_slices=[1, slices[0], slices[1]]))
    clin = dict(dim_in=clin_dim, dim_out=dim_out, dropout=dropout)
    if name == "XR1MR2C1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size], [mr_size, mr_size, slices[0]], [mr_size, mr_size, slices[1]],
                                      [16]],
                    fe=dict(fe, clin=clin), agg=dict(agg, num_slices=[1, slices[0], slices[1], 1]))
    if name == "MR3CnnTrf":
        return dict(base, input_size=[[mr_size, mr_size, s] for s in slices[:3]], fe=fe,
                    agg=dict(agg, num_slices=list(slices[:3])))
    if name == "XR1MR3C1CnnTrf":
        return dict(base, input_size=[[xr_size, xr_size]] + [[mr_size, mr_size, s] for s in slices[:3]] + [[16]],
                    fe=dict(fe, clin=clin), agg=dict(agg, num_slices=[1] + list(slices[:3]) + [1]))
    raise ValueError(name)


def make_inputs(name: str, cfg: dict, batch: int, seed: int, device: str = "cpu"):
    """Synthetic N(0,1) images / volumes, clinical 9-vector (z-scores + one-hots,
    koafusion/datasets/oai/_dataset.py:254-266), Bernoulli(0.12) targets — all drawn on the CPU."""
    g = torch.Generator().manual_seed(seed)
    ins = []
    for shape in cfg["input_size"]:
        if len(shape) == 1:  # clinical
            z = torch.randn(batch, 3, generator=g)
            oh = [F.one_hot(torch.randint(0, 2, (batch,), generator=g), 2).float() for _ in range(3)]
            clin = torch.cat([z[:, 0:1], oh[0], z[:, 1:2], oh[1], oh[2], z[:, 2:3]], dim=1)
            ins.append(clin[:, None, :].to(device))
        else:
            ins.append(torch.randn(batch, 1, *shape, generator=g).to(device))
    target = (torch.rand(batch, generator=g) < 0.12).long().to(device)
    return ins, target
