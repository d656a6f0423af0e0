"""Feature extractors: parameter containers with the reference's ``state_dict`` layout + the call into
``koa_fe_forward`` / ``koa_fe_backward``.

The reference builds a torchvision ResNet (``koafusion/models/_core_fes.py:6-15``, vendored copy in
``_torchvision.py:141-330``), keeps ``list(children())[:-1]`` in an ``nn.Sequential`` and calls it on the
slice batch (``_xrNmrMcP.py:47-59,218-220``). Here the same module tree exists only to own the fp32 master
parameters under identical names (``_fe1.4.0.conv1.weight`` ...): the compute of the whole extractor is one
C call into libkoa_b200.so. There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
from torch import nn

from .. import _lib, dataparallel

_FP16_MAX = 65504.0


def poison_if_out_of_fp16_range(x: torch.Tensor, stem_weight: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """The raw output of the stem convolution is stored in fp16; every later tensor sits behind a BatchNorm. Its magnitude
    is at most ``max|x| * max_c sum|w_c|`` (the three input channels are folded into the 7x7 weights). Where that bound
    reaches the fp16 range the result cannot be trusted (measured: 1e7-scale inputs came back finite and wrong, because the
    ReLU kernels turn NaN into 0), so the features are replaced by NaN: the training loop sees a NaN loss, as it would after
    an overflow in PyTorch. Device-side (a min / max pass over the input, a few scalar kernels, one select): no host
    read-back. Inputs the reference's transform chain delivers (|x| < 5) or raw 8-bit data are far below the limit
    (about 2e4 with Kaiming-initialised weights)."""
    with torch.no_grad():
        lo, hi = torch.aminmax(x)
        bound = torch.maximum(hi, -lo) * stem_weight.abs().sum(dim=(1, 2, 3)).amax()
        ok = bound < _FP16_MAX  # False for NaN / inf inputs too
    return torch.where(ok, feat, feat.new_full((), float("nan")))


_BLOCKS = {
    "resnet18": ("basic", (2, 2, 2, 2), 1, 64),
    "resnet34": ("basic", (3, 4, 6, 3), 1, 64),
    "resnet50": ("bottleneck", (3, 4, 6, 3), 1, 64),
    "resnext50_32x4d": ("bottleneck", (3, 4, 6, 3), 32, 4),
}
FE_OUT_CH = {"resnet18": 512, "resnet34": 512, "resnet50": 2048, "resnext50_32x4d": 2048}


class _ResidualBlock(nn.Module):
    """conv/bn holders of one BasicBlock / Bottleneck (``_torchvision.py:34-138``); never called."""

    def __init__(self, kind: str, inplanes: int, planes: int, stride: int, groups: int, base_width: int):
        super().__init__()
        if kind == "bottleneck":
            width = int(planes * (base_width / 64.0)) * groups
            out = planes * 4
            self.conv1 = nn.Conv2d(inplanes, width, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(width)
            self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, groups=groups, bias=False)
            self.bn2 = nn.BatchNorm2d(width)
            self.conv3 = nn.Conv2d(width, out, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(out)
        else:
            out = planes
            self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or inplanes != out:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, out, 1, stride=stride, bias=False), nn.BatchNorm2d(out))
        self.stride = stride

    def units(self):
        u = [(self.conv1, self.bn1), (self.conv2, self.bn2)]
        if hasattr(self, "conv3"):
            u.append((self.conv3, self.bn3))
        if self.downsample is not None:
            u.append((self.downsample[0], self.downsample[1]))
        return u

    def forward(self, x):  # pragma: no cover - the engine runs the whole extractor
        raise RuntimeError("koafusion-b200 blocks are parameter holders; call the enclosing SliceEncoder")


class KoaResNet(nn.Module):
    """Same children, names, shapes and default initialisation as the reference ResNet
    (``_torchvision.py:170-190``): conv1, bn1, relu, maxpool, layer1..4, avgpool, fc."""

    def __init__(self, arch: str, num_classes: int = 1000):
        super().__init__()
        kind, layers, groups, wpg = _BLOCKS[arch]
        self.arch = arch
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        inplanes = 64
        expansion = 4 if kind == "bottleneck" else 1
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for bi in range(n):
                stride = 2 if (li > 0 and bi == 0) else 1
                blocks.append(_ResidualBlock(kind, inplanes, planes, stride, groups, wpg))
                inplanes = planes * expansion
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        """The classifier the reference's ``dict_fes[arch]()`` returns (``_torchvision.py:227-242``): extractor -> global
        average pool -> ``fc``. The koafusion model classes never call it (they keep ``children()[:-1]``); here it runs the
        CUDA extractor on this module's own children (shared, not copied) and the fp32 small-linear kernel for ``fc``."""
        from ._small import small_linear

        enc = SliceEncoder(self, with_gap=True)
        enc.train(self.training)
        return small_linear(enc(x).flatten(1), self.fc.weight, self.fc.bias)


def _make_fe(arch: str):
    def ctor(pretrained: bool = False, **_):
        if pretrained:
            raise RuntimeError(
                "pretrained=True needs the ImageNet checkpoint download of the reference "
                "(koafusion/models/_torchvision.py:258-261); load weights with load_state_dict instead")
        return KoaResNet(arch)

    return ctor


# ``dict_fes`` of the reference (koafusion/models/_core_fes.py:6-15). The four torchvision entries that no
# koafusion model class accepts (squeezenet/vgg/densenet/inception) are not provided.
dict_fes = {name: _make_fe(name) for name in _BLOCKS}


def _out_hw(x: int) -> int:
    """Spatial size after stem (7x7 s2 p3), max-pool (3x3 s2 p1) and three stride-2 stages."""
    x = (x + 6 - 7) // 2 + 1
    for _ in range(4):
        x = (x + 2 - 3) // 2 + 1
    return x


class _FEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc: "SliceEncoder", x: torch.Tensor, n_img: int, h: int, w: int, slices: int, need_bw: bool,
                *params):
        lib = _lib.load()
        dev = _lib.require_same_device("SliceEncoder", x, *params)
        desc = _lib.FeDesc(arch=_lib.ARCH_IDS[enc.arch], n_img=n_img, h=h, w=w, slices=slices,
                           with_gap=1 if enc.with_gap else 0, training=1 if enc.training else 0,
                           need_backward=1 if need_bw else 0,
                           input_for_backward=x.data_ptr() if slices == 0 else None)
        nbytes = lib.koa_fe_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            _lib.check(-1, "koa_fe_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        ch, oh, ow = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.koa_fe_out_shape(C.byref(desc), C.byref(ch), C.byref(oh), C.byref(ow)), "koa_fe_out_shape")
        feat = torch.empty((n_img, ch.value) if enc.with_gap else (n_img, oh.value * ow.value, ch.value),
                           dtype=torch.float32, device=x.device)
        table = enc._param_table()
        with _lib.on_device(dev):
            _lib.check(lib.koa_fe_forward(C.byref(desc), table, x.data_ptr(), ws.data_ptr(), feat.data_ptr(),
                                          _lib.current_stream()), "koa_fe_forward")
            if enc.training:
                torch._foreach_add_(enc._nbt(), 1)
        ctx.enc, ctx.desc, ctx.ws, ctx.table, ctx.x = enc, desc, ws, table, x
        ctx.n_params = len(params)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        lib = _lib.load()
        enc = ctx.enc
        params = enc._trainable()
        grads, flat = _lib.zeros_like_flat([p if p.requires_grad else None for p in params])
        gtable = _lib.ptr_table(grads)
        dev = _lib.require_same_device("SliceEncoder backward", dfeat, ctx.ws)
        dfeat = dfeat.contiguous().float()
        with _lib.on_device(dev):
            if not dataparallel.active() or flat is None:
                _lib.check(lib.koa_fe_backward(C.byref(ctx.desc), ctx.table, gtable, ctx.ws.data_ptr(), dfeat.data_ptr(),
                                               _lib.current_stream()), "koa_fe_backward")
            else:
                # Data-parallel run: the backward pass goes stage by stage (layer4 -> layer1 + stem), and the gradients
                # of a finished stage, a contiguous slice of the flat buffer (units are laid out in network order), go on
                # the wire while the earlier stages still compute. Only the last, smallest slice (stem + layer1, 0.9 MB
                # of the 94 MB of a ResNet-50) is reduced after this extractor's compute has ended.
                bounds = enc._stage_bounds(params, grads, flat)
                for begin, end, with_stem, lo, hi, stage_params in bounds:
                    _lib.check(lib.koa_fe_backward_range(C.byref(ctx.desc), ctx.table, gtable, ctx.ws.data_ptr(),
                                                         dfeat.data_ptr(), begin, end, with_stem, _lib.current_stream()),
                               "koa_fe_backward_range")
                    dataparallel.sync_flat(flat[lo:hi], stage_params)
        ctx.ws = None
        return (None, None, None, None, None, None, None, *grads)


class SliceEncoder(nn.Sequential):
    """``nn.Sequential(*list(resnet.children())[:-1 | -2])`` of the reference, executed by the CUDA engine.

    ``forward`` accepts what the reference feeds its extractor, an image batch (N, 3, H, W) whose three
    channels are the same grey image (``repeat(..., k=3)``), or the single-channel batch (N, 1, H, W).
    ``encode_volume`` takes the raw (B, 1, R, C, S) slice-innermost volume and fuses the einops
    rearrange + repeat into the stem (``koafusion/models/_xrNmrMcP.py:209-213``).
    """

    def __init__(self, resnet: KoaResNet, with_gap: bool = True):
        children = list(resnet.children())
        super().__init__(*(children[:-1] if with_gap else children[:-2]))
        self.arch = resnet.arch
        self.with_gap = with_gap
        self._channels_checked = False

    # -- parameter plumbing -----------------------------------------------------------------------
    def _units(self):
        units = [(self[0], self[1])]
        for li in range(4, 8):
            for blk in self[li]:
                units.extend(blk.units())
        return units

    def _trainable(self) -> List[torch.Tensor]:
        out = []
        for conv, bn in self._units():
            out += [conv.weight, bn.weight, bn.bias]
        return out

    def _stage_bounds(self, params, grads, flat):
        """(block_begin, block_end, with_stem, flat_lo, flat_hi, parameters) per backward stage, last stage first: layer4,
        layer3, layer2, then layer1 together with the stem. The flat gradient buffer holds the units in network order."""
        blocks = [len(self[li]) for li in range(4, 8)]
        units_per_layer = [sum(len(b.units()) for b in self[li]) for li in range(4, 8)]
        offs = {}
        base = flat.data_ptr()
        for i, g in enumerate(grads):
            if g is not None:
                offs[i] = ((g.data_ptr() - base) // 4, g.numel())
        live = sorted(offs)

        def span(u_lo, u_hi):  # flat range and parameters of units [u_lo, u_hi)
            idx = [i for i in live if 3 * u_lo <= i < 3 * u_hi]
            if not idx:
                return 0, 0, []
            lo = offs[idx[0]][0]
            hi = offs[idx[-1]][0] + offs[idx[-1]][1]
            return lo, hi, [params[i] for i in idx]

        out = []
        b_hi, u_hi = sum(blocks), 1 + sum(units_per_layer)
        for li in (3, 2, 1):
            b_lo, u_lo = b_hi - blocks[li], u_hi - units_per_layer[li]
            out.append((b_lo, b_hi, 0, *span(u_lo, u_hi)))
            b_hi, u_hi = b_lo, u_lo
        out.append((0, b_hi, 1, *span(0, u_hi)))
        return out

    def _nbt(self):
        return [bn.num_batches_tracked for _, bn in self._units()]

    def _param_table(self):
        t = []
        for conv, bn in self._units():
            t += [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        for p in t:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                raise _lib.KoaError("feature-extractor parameters must be contiguous fp32 CUDA tensors")
        return _lib.ptr_table(t)

    # -- compute ------------------------------------------------------------------------------------
    def _run(self, x, n_img, h, w, slices):
        # (n_img, C) with GAP, (n_img, h_out*w_out, C) without
        params = self._trainable()
        need_bw = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        feat = _FEFunction.apply(self, x, n_img, h, w, slices, need_bw, *params)
        return poison_if_out_of_fp16_range(x, self[0].weight, feat)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] not in (1, 3):
            raise ValueError(f"SliceEncoder expects (N, 1|3, H, W), got {tuple(x.shape)}")
        n, _, h, w = x.shape
        if x.shape[1] == 3 and not self._channels_checked:
            # The stem folds the three input channels into one (the reference always feeds `repeat(x, k=3)` of a grey
            # image, _xrNmrMcP.py:211-213): a genuine 3-channel image would silently lose two of them. Checked on the first
            # call of every encoder (one device read-back), and on every call with KOA_CHECK_CHANNELS=1.
            import os

            if x.stride(1) != 0 and not bool(((x[:, 0] == x[:, 1]) & (x[:, 0] == x[:, 2])).all()):
                raise ValueError("SliceEncoder folds the three input channels into one grey channel: the channels of this "
                                 "input differ (koafusion feeds repeat(x, k=3); colour images are not supported)")
            self._channels_checked = os.environ.get("KOA_CHECK_CHANNELS", "0") != "1"
        img = x[:, 0].contiguous().float()
        feat = self._run(img, n, h, w, 0)
        if self.with_gap:
            return feat.reshape(n, -1, 1, 1)  # (N, C, 1, 1) like AdaptiveAvgPool2d
        oh, ow = _out_hw(h), _out_hw(w)
        return feat.permute(0, 2, 1).reshape(n, feat.shape[-1], oh, ow)  # (N, C, h_out, w_out)

    def encode_volume(self, vol: torch.Tensor) -> torch.Tensor:
        """(B, 1, R, C, S) -> tokens (B, S * positions, C)."""
        if vol.dim() != 5 or vol.shape[1] != 1:
            raise ValueError(f"expected (B, 1, R, C, S), got {tuple(vol.shape)}")
        b, _, r, c, s = vol.shape
        feat = self._run(vol.contiguous().float(), b * s, r, c, s)
        return feat.reshape(b, -1, feat.shape[-1])

    def encode_image(self, img: torch.Tensor) -> torch.Tensor:
        """(B, 1, R, C) -> tokens (B, positions, C)."""
        if img.dim() != 4 or img.shape[1] != 1:
            raise ValueError(f"expected (B, 1, R, C), got {tuple(img.shape)}")
        b, _, r, c = img.shape
        feat = self._run(img.reshape(b, r, c).contiguous().float(), b, r, c, 0)
        return feat.reshape(b, -1, feat.shape[-1])
