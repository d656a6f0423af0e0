"""Drop-in versions of the six ``koafusion.models`` classes (+ the two 3-MRI pattern extensions).

Same constructor signature ``Model(config, path_weights)``, same config keys, same ``state_dict`` keys and
shapes, same forward signatures and return convention (``{"main": logits}`` or the bare tensor, chosen by
``config.output_type``) as the reference (``koafusion/models/__init__.py:8-15`` and the class files cited
per class). The compute is the CUDA path: one engine call per feature extractor / transformer.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
from torch import nn

from .. import _lib
from ._fe import FE_OUT_CH, SliceEncoder, dict_fes
from ._feat import FeaT
from ._small import FeatC1, small_linear

_SPATIAL = {320: 10, 160: 5, 128: 4, 96: 3, 64: 2, 32: 1, 350: 11, 25: 1}


def _cfg(config, key):
    return config[key]


def _output(config, logits):
    """``output_type`` switch shared by every class (e.g. ``_xrNmrMcP.py:259-264``)."""
    kind = config["output_type"] if not hasattr(config, "output_type") else config.output_type
    if kind == "main":
        return logits
    if kind == "dict":
        out = OrderedDict()
        out["main"] = logits
        return out
    raise ValueError(f"Unknown output_type: {kind}")


def _make_fe(arch, pretrained, with_gap):
    return SliceEncoder(dict_fes[arch](pretrained=pretrained), with_gap=with_gap)


def _drop2d(p):
    return nn.Dropout2d(p=p) if p else nn.Identity()


class _ChannelDropout(torch.autograd.Function):
    """``nn.Dropout2d`` of the reference on the extractor output ``(B*S, C, h, w)`` (``_xrNmrMcP.py:62-74,226-229``), on
    the engine's token layout ``(B, S*positions, C)``: one Philox draw per (slice, channel), shared by every spatial
    position of that slice; the backward pass regenerates the mask (``koa_channel_dropout``)."""

    @staticmethod
    def forward(ctx, tokens, n_img, positions, p, seed, site):
        lib = _lib.load()
        _lib.require_cuda(tokens, "channel dropout")
        x = tokens.contiguous().float()
        out = torch.empty_like(x)
        with _lib.on_device(x.device):
            _lib.check(lib.koa_channel_dropout(x.data_ptr(), out.data_ptr(), n_img, positions, x.shape[-1], seed, site,
                                               float(p), _lib.current_stream()), "koa_channel_dropout")
        ctx.args = (n_img, positions, float(p), seed, site)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n_img, positions, p, seed, site = ctx.args
        g = g.contiguous().float()
        out = torch.empty_like(g)
        with _lib.on_device(g.device):
            _lib.check(lib.koa_channel_dropout(g.data_ptr(), out.data_ptr(), n_img, positions, g.shape[-1], seed, site, p,
                                               _lib.current_stream()), "koa_channel_dropout")
        return out, None, None, None, None, None


def _apply_drop2d(drop, tokens, n_img, site=0):
    """Dropout2d on the (N, C, h, w) extractor output == per-(image, channel) dropout on the (B, S * positions, C) tokens.
    ``drop.last_seed`` keeps the Philox seed of the call (tests replay the mask with ``koa_dropout_mask``)."""
    if isinstance(drop, nn.Identity) or not drop.training or drop.p == 0:
        return tokens
    b, sp, c = tokens.shape
    positions = (b * sp) // n_img
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())  # torch's CPU generator: torch.manual_seed replays
    drop.last_seed, drop.last_site = seed, 0xD000 + site
    return _ChannelDropout.apply(tokens, n_img, positions, drop.p, seed, 0xD000 + site)


_BRANCH_STREAMS = {}
_BRANCH_STREAMS_ON = [None]  # None: follow the environment (KOA_BRANCH_STREAMS, default on)


def set_branch_streams(on):
    """Run the modality branches of the fusion models on their own CUDA streams (default) or one after the other on
    the current stream (``False``; what a per-kernel timing pass wants). ``None`` restores the environment default."""
    _BRANCH_STREAMS_ON[0] = on


def _branch_streams_on():
    import os

    if _BRANCH_STREAMS_ON[0] is not None:
        return bool(_BRANCH_STREAMS_ON[0])
    return os.environ.get("KOA_BRANCH_STREAMS", "1") != "0"


def _run_branches(fns):
    """Run the independent modality branches (extractor [+ per-sequence transformer]) of a fusion model, each on its
    own CUDA stream: their kernels are launched back to back by one host thread, but the small ones (transformer
    GEMMs at a few hundred rows, layer-4 convolutions, pooling) no longer leave SMs idle while the next branch waits.
    Autograd replays each branch's backward on the branch's stream. The reference computes the branches one after
    the other on one stream (``_xrNmrMcP.py:209-236``); the results are the same."""
    if len(fns) < 2 or not _branch_streams_on() or not torch.cuda.is_available():
        return [f() for f in fns]
    main = torch.cuda.current_stream()
    key = (main.device.index, len(fns))
    if key not in _BRANCH_STREAMS:
        _BRANCH_STREAMS[key] = [torch.cuda.Stream(device=main.device) for _ in range(len(fns))]
        # parameters shared by runs with and without branch streams keep their AccumulateGrad node: intentional
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    outs = []
    for f, st in zip(fns, _BRANCH_STREAMS[key]):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            outs.append(f())
    for o, st in zip(outs, _BRANCH_STREAMS[key]):
        main.wait_stream(st)
        o.record_stream(main)  # consumed on the main stream: keep the allocator from recycling it early
    return outs


def _scaled(shape, scale):
    return [round(s * d) for s, d in zip(shape, scale)]


class _Base(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if self.config["debug"]:
            print("Config at model init", self.config)
        self.vs = dict()

    def _finish(self, path_weights):
        if self.config["restore_weights"]:
            self.load_state_dict(torch.load(path_weights))

    def _replicate_for_data_parallel(self):
        """``nn.DataParallel`` over SEVERAL devices (the reference's wrapper, ``run/train_prog_fus.py:84``) replicates the
        module into one Python thread per GPU of one process. This path scales as one process per GPU instead
        (``oaprogressionmmf_b200.dataparallel.wrap`` after ``torch.distributed`` initialisation: same per-replica BatchNorm
        semantics, gradient all-reduce over NCCL overlapped with backward); a multi-device replication is refused here rather
        than left to fail inside a kernel (measured in round 2: misaligned-address fault). ``nn.DataParallel`` with ONE device
        never replicates and works as before."""
        raise TypeError("koafusion-b200 models do not support nn.DataParallel over several devices: launch one process per GPU "
                        "(torchrun) and wrap the model with oaprogressionmmf_b200.dataparallel.wrap(model) instead "
                        "(INTEGRATION.md section 1)")

    def _make_feat(self, num_patches, with_cls=True):
        agg = self.config["agg"]
        return FeaT(num_patches=num_patches, patch_dim=self.vs["agg_in_depth"], emb_dim=self.vs["agg_in_depth"],
                    depth=agg["depth"], heads=agg["heads"], mlp_dim=agg["mlp_dim"],
                    num_classes=self.config["output_channels"], emb_dropout=agg["emb_dropout"], with_cls=with_cls,
                    mlp_dropout=agg["mlp_dropout"])


class XR1Cnn(_Base):
    """XR-only classifier (``koafusion/models/_xr1_cnn.py:9-81``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        arch = config["fe"]["arch"]
        if arch not in FE_OUT_CH:
            raise ValueError("Unknown `num_elems` for `model.fe` output. Get via `model.debug=true`")
        self._fe = _make_fe(arch, config["fe"]["pretrained"], True)
        hid = config["agg"]["hidden_size"]
        self._agg = nn.Sequential(nn.Dropout(config["agg"]["dropout"]), nn.Linear(FE_OUT_CH[arch], hid), nn.ReLU(),
                                  nn.Dropout(config["agg"]["dropout"]))
        self._final = nn.Linear(hid, config["output_channels"])
        self._finish(path_weights)

    def forward(self, input):
        f = self._fe.encode_image(input).flatten(1)
        f = self._agg[0](f)
        f = small_linear(f, self._agg[1].weight, self._agg[1].bias, _lib.ACT_RELU)
        f = self._agg[3](f)
        return _output(self.config, small_linear(f, self._final.weight, self._final.bias))


class MR1CnnTrf(_Base):
    """One MRI sequence: per-slice CNN + slice-aggregation transformer (``_mrN_cnn_trf.py:12-139``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        fe = config["fe"]
        if fe["arch"] not in ("resnet18", "resnet34", "resnet50"):
            raise ValueError("Unsupported `model.fe.arch`")
        self._fe = _make_fe(fe["arch"], fe["pretrained"], fe["with_gap"])
        self._fe_drop = _drop2d(fe["dropout"])
        self.vs["fe_out_ch"] = FE_OUT_CH[fe["arch"]]
        t = list(config["input_size"][0])
        if config["downscale"]:
            t = _scaled(t, config["downscale"][0])
        self.vs["shape_in"] = t
        if fe["with_gap"]:
            self.vs["fe_out_spat"] = (1, 1, 1)
        else:
            try:
                self.vs["fe_out_spat"] = tuple({320: 10, 160: 5, 128: 4, 96: 3, 64: 2, 32: 1}[e] for e in t)
            except KeyError:
                raise ValueError("Unspecified `model.fe` output shape for given `model.input_size`")
        sp = self.vs["fe_out_spat"]
        if fe["dims_view"] == "rc":
            self.vs["agg_in_len"] = t[2] * sp[0] * sp[1]
        elif fe["dims_view"] == "cs":
            self.vs["agg_in_len"] = t[0] * sp[1] * sp[2]
        elif fe["dims_view"] == "rs":
            self.vs["agg_in_len"] = t[1] * sp[0] * sp[2]
        else:
            raise ValueError("Unsupported `model.fe.dims_view`")
        self.vs["agg_in_depth"] = self.vs["fe_out_ch"]
        self._agg = self._make_feat(self.vs["agg_in_len"])
        self._finish(path_weights)

    def forward(self, input):
        view = self.config["fe"]["dims_view"]
        vol = input
        if view == "cs":  # slice along rows: images are (c, s)
            vol = input.permute(0, 1, 3, 4, 2)
        elif view == "rs":  # slice along columns: images are (r, s)
            vol = input.permute(0, 1, 2, 4, 3)
        tok = _apply_drop2d(self._fe_drop, self._fe.encode_volume(vol), vol.shape[0] * vol.shape[-1])
        out, _, _ = self._agg.run(tok, compute_head=True)
        return _output(self.config, out.flatten(1))


class MR2CnnTrf(_Base):
    """Two MRI sequences, flat fusion in one transformer (``_mrN_cnn_trf.py:142-272``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        fe = config["fe"]
        if fe["arch"] not in ("resnet18", "resnet34", "resnet50"):
            raise ValueError("Unsupported `model.fe.arch`")
        self._fe0 = _make_fe(fe["arch"], fe["pretrained"], fe["with_gap"])
        self._fe1 = _make_fe(fe["arch"], fe["pretrained"], fe["with_gap"])
        self._fe0_drop = _drop2d(fe["dropout"])
        self._fe1_drop = _drop2d(fe["dropout"])
        self.vs["fe_out_ch"] = FE_OUT_CH[fe["arch"]]
        if fe["with_gap"]:
            self.vs["fe_out_spat"] = (1, 1)
        elif config["input_size"][0][0] == 320:
            self.vs["fe_out_spat"] = (5, 5)
        else:
            raise ValueError("Unspecified `model.fe` output shape for given `model.input_size`")
        ns = config["agg"]["num_slices"]
        self.vs["agg_in_len"] = (ns[0] + ns[1]) * math.prod(self.vs["fe_out_spat"])
        self.vs["agg_in_depth"] = self.vs["fe_out_ch"]
        self._agg = self._make_feat(self.vs["agg_in_len"])
        self._finish(path_weights)

    def forward(self, input0, input1):
        t0 = _apply_drop2d(self._fe0_drop, self._fe0.encode_volume(input0), input0.shape[0] * input0.shape[-1], 0)
        t1 = _apply_drop2d(self._fe1_drop, self._fe1.encode_volume(input1), input1.shape[0] * input1.shape[-1], 1)
        out, _, _ = self._agg.run(torch.cat([t0, t1], dim=1), compute_head=True)
        return _output(self.config, out.flatten(1))


class _XRMRBase(_Base):
    """Shared constructor bookkeeping of the XR + MRI classes (``_xr1mrN.py:19-85``, ``_xrNmrMcP.py:40-135``)."""

    def _build_fes(self, n_mr, first_mr_index=1):
        fe = self.config["fe"]
        assert fe["xr"]["arch"] in FE_OUT_CH
        assert fe["mr"]["arch"] in FE_OUT_CH
        gap = bool(fe["xr"]["with_gap"] or fe["mr"]["with_gap"])
        self._fe0 = _make_fe(fe["xr"]["arch"], fe["xr"]["pretrained"], gap)
        for i in range(n_mr):
            setattr(self, f"_fe{first_mr_index + i}", _make_fe(fe["mr"]["arch"], fe["mr"]["pretrained"], gap))
        self._fe0_drop = _drop2d(fe["xr"]["dropout"])
        for i in range(n_mr):
            setattr(self, f"_fe{first_mr_index + i}_drop", _drop2d(fe["mr"]["dropout"]))
        self.vs["fe0_out_ch"] = FE_OUT_CH[fe["xr"]["arch"]]
        self.vs["fe12_out_ch"] = self.vs["fe1_out_ch"] = FE_OUT_CH[fe["mr"]["arch"]]
        shapes = [list(s) for s in self.config["input_size"]]
        if self.config["downscale"]:
            shapes = [_scaled(s, d) for s, d in zip(shapes, self.config["downscale"])]
        for i, s in enumerate(shapes):
            self.vs[f"fe{i}_shape_in"] = s
        assert all(e in _SPATIAL for e in shapes[0])
        for i in range(1, 1 + n_mr):
            assert all(e in _SPATIAL for e in shapes[i][:2])  # slice dimension is not checked
        self.vs["fe0_out_spat"] = (1, 1) if fe["xr"]["with_gap"] else tuple(_SPATIAL[e] for e in shapes[0])
        for i in range(1, 1 + n_mr):
            self.vs[f"fe{i}_out_spat"] = (1, 1) if fe["mr"]["with_gap"] else tuple(_SPATIAL[e] for e in shapes[i][:2])
        ns = self.config["agg"]["num_slices"]
        self.vs["agg_in_len_0"] = math.prod(self.vs["fe0_out_spat"])
        for i in range(1, 1 + n_mr):
            self.vs[f"agg_in_len_{i}"] = ns[i] * math.prod(self.vs[f"fe{i}_out_spat"])
        self.vs["agg_in_depth"] = FE_OUT_CH[fe["mr"]["arch"]]

    def _mr_tokens(self, i, vol):
        return _apply_drop2d(getattr(self, f"_fe{i}_drop"), getattr(self, f"_fe{i}").encode_volume(vol),
                             vol.shape[0] * vol.shape[-1], i)

    def _xr_tokens(self, img):
        return _apply_drop2d(self._fe0_drop, self._fe0.encode_image(img), img.shape[0], 0)


class XR1MR1CnnTrf(_XRMRBase):
    """XR + one MRI sequence, flat (``_xr1mrN.py:11-158``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        self._build_fes(1)
        self._agg = self._make_feat(self.vs["agg_in_len_0"] + self.vs["agg_in_len_1"])
        self._finish(path_weights)

    def forward(self, input0, input1):
        t = torch.cat([self._xr_tokens(input0), self._mr_tokens(1, input1)], dim=1)
        out, _, _ = self._agg.run(t, compute_head=True)
        return _output(self.config, out.flatten(1))


class XR1MR2CnnTrf(_XRMRBase):
    """XR + two MRI sequences, hierarchical (``_xr1mrN.py:161-369``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        self._build_fes(2)
        self._agg_1 = self._make_feat(self.vs["agg_in_len_1"], with_cls=False)
        self._agg_2 = self._make_feat(self.vs["agg_in_len_2"], with_cls=False)
        self._agg_final = self._make_feat(self.vs["agg_in_len_0"] + self.vs["agg_in_len_1"] + self.vs["agg_in_len_2"])
        self._finish(path_weights)

    def forward(self, input0, input1, input2):
        t0 = self._xr_tokens(input0)
        # per-sequence transformers hand on all token states; their heads are dead compute in the reference
        # (_xr1mrN.py:347-348) and are skipped here
        _, s1, _ = self._agg_1.run(self._mr_tokens(1, input1), compute_head=False)
        _, s2, _ = self._agg_2.run(self._mr_tokens(2, input2), compute_head=False)
        out, _, _ = self._agg_final.run(torch.cat([t0, s1, s2], dim=1), compute_head=True)
        return _output(self.config, out.flatten(1))


class XR1MR2C1CnnTrf(_XRMRBase):
    """XR + two MRI sequences + clinical token: the reference's full model (``_xrNmrMcP.py:32-264``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        self._build_fes(2)
        self._fe3 = FeatC1(config=config["fe"]["clin"])
        self._fe3_drop = nn.Identity()
        self.vs["fe3_out_spat"] = (1,)
        self.vs["agg_in_len_3"] = config["agg"]["num_slices"][3] * 1
        self._agg_1 = self._make_feat(self.vs["agg_in_len_1"], with_cls=False)
        self._agg_2 = self._make_feat(self.vs["agg_in_len_2"], with_cls=False)
        self._agg_final = self._make_feat(self.vs["agg_in_len_0"] + self.vs["agg_in_len_1"] +
                                          self.vs["agg_in_len_2"] + self.vs["agg_in_len_3"])
        self._finish(path_weights)

    def forward(self, input0, input1, input2, input3):
        s1, s2, t0 = _run_branches([lambda: self._agg_1.run(self._mr_tokens(1, input1), compute_head=False)[1],
                                    lambda: self._agg_2.run(self._mr_tokens(2, input2), compute_head=False)[1],
                                    lambda: self._xr_tokens(input0)])
        t3 = self._fe3(input3)
        out, _, _ = self._agg_final.run(torch.cat([t0, s1, s2, t3], dim=1), compute_head=True)
        return _output(self.config, out.flatten(1))


# ---- pattern extensions required by BASELINE.json ("XR+3MRI+clin"); no such class in the reference -----------
class MR3CnnTrf(_Base):
    """Three MRI sequences, hierarchical like ``XR1MR2C1CnnTrf`` without XR / clinical (SURVEY.md §8 a-ext)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        mr = config["fe"]["mr"]
        for i in (1, 2, 3):
            setattr(self, f"_fe{i}", _make_fe(mr["arch"], mr["pretrained"], True))
            setattr(self, f"_fe{i}_drop", _drop2d(mr["dropout"]))
        self.vs["agg_in_depth"] = FE_OUT_CH[mr["arch"]]
        ns = config["agg"]["num_slices"]
        for i in (1, 2, 3):
            setattr(self, f"_agg_{i}", self._make_feat(ns[i - 1], with_cls=False))
        self._agg_final = self._make_feat(sum(ns[:3]))
        self._finish(path_weights)

    def forward(self, input0, input1, input2):
        states = []
        for i, vol in enumerate((input0, input1, input2), start=1):
            tok = _apply_drop2d(getattr(self, f"_fe{i}_drop"), getattr(self, f"_fe{i}").encode_volume(vol),
                                vol.shape[0] * vol.shape[-1], i)
            states.append(getattr(self, f"_agg_{i}").run(tok, compute_head=False)[1])
        out, _, _ = self._agg_final.run(torch.cat(states, dim=1), compute_head=True)
        return _output(self.config, out.flatten(1))


class XR1MR3C1CnnTrf(_XRMRBase):
    """XR + DESS + TSE + T2 map + clinical token (BASELINE.json configs 3-4; extension of ``_xrNmrMcP.py``)."""

    def __init__(self, config, path_weights):
        super().__init__(config)
        self._build_fes(3)
        self._fe4 = FeatC1(config=config["fe"]["clin"])
        ns = config["agg"]["num_slices"]
        for i in (1, 2, 3):
            setattr(self, f"_agg_{i}", self._make_feat(self.vs[f"agg_in_len_{i}"], with_cls=False))
        self._agg_final = self._make_feat(self.vs["agg_in_len_0"] + sum(self.vs[f"agg_in_len_{i}"] for i in (1, 2, 3)) +
                                          ns[4])
        self._finish(path_weights)

    def forward(self, input0, input1, input2, input3, input4):
        def mr(i, vol):
            return lambda: getattr(self, f"_agg_{i}").run(self._mr_tokens(i, vol), compute_head=False)[1]

        # the largest branch (DESS, 64 slices) first (the launch order was measured not to matter: 177-178 knees/s either way)
        s1, s2, s3, t0 = _run_branches([mr(1, input1), mr(2, input2), mr(3, input3), lambda: self._xr_tokens(input0)])
        parts = [t0, s1, s2, s3, self._fe4(input4)]
        out, _, _ = self._agg_final.run(torch.cat(parts, dim=1), compute_head=True)
        return _output(self.config, out.flatten(1))


dict_models = {
    "XR1Cnn": XR1Cnn,
    "MR1CnnTrf": MR1CnnTrf,
    "MR2CnnTrf": MR2CnnTrf,
    "XR1MR1CnnTrf": XR1MR1CnnTrf,
    "XR1MR2CnnTrf": XR1MR2CnnTrf,
    "XR1MR2C1CnnTrf": XR1MR2C1CnnTrf,
    # extensions
    "MR3CnnTrf": MR3CnnTrf,
    "XR1MR3C1CnnTrf": XR1MR3C1CnnTrf,
}
