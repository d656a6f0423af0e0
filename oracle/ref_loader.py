"""Loads the UNMODIFIED reference (``koafusion.models`` + ``FocalLoss``) from ``/root/reference`` (build container) or from
``oracle/_ref`` (the verbatim copy ``oracle/build_ref.py`` makes, which travels to the GPU box). Test infrastructure: used
by the golden-vector generators and by the CPU arm of ``bench.py``; the product never imports it."""
from __future__ import annotations

import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference", os.path.join(ROOT, "oracle", "_ref"))


def find_ref():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "koafusion", "models", "__init__.py")):
            return c
    return None


def _ensure_path():
    ref = find_ref()
    if ref is None:
        raise ImportError("the reference is neither at /root/reference nor at oracle/_ref (python oracle/build_ref.py)")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    return ref


class AttrDict(dict):
    """item + attribute access, as the reference reads its OmegaConf config both ways."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return [to_attr(v) for v in d]
    return d


def focal_loss(gamma=2):
    ref = _ensure_path()
    spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(ref, "koafusion/various/_losses.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FocalLoss(gamma=gamma)


def build_extension_model(name, cfg):
    """3-MRI pattern extension assembled from the reference's own blocks, the way _xrNmrMcP.py:40-179,209-255 assembles two
    sequences (the reference has no 3-MRI class; SURVEY.md 8 row a-ext)."""
    _ensure_path()
    import torch
    from einops import rearrange, repeat
    from koafusion.models._core_fes import dict_fes
    from koafusion.models._core_trf import FeaT
    from koafusion.models._xrNmrMcP import FeatC1
    from torch import nn

    agg = cfg["agg"]
    ns = agg["num_slices"]

    def fe(arch):
        return nn.Sequential(*list(dict_fes[arch](pretrained=False).children())[:-1])

    def feat(n, with_cls):
        return FeaT(num_patches=n, patch_dim=2048, emb_dim=2048, depth=agg["depth"], heads=agg["heads"],
                    mlp_dim=agg["mlp_dim"], num_classes=cfg["output_channels"], emb_dropout=agg["emb_dropout"],
                    with_cls=with_cls, mlp_dropout=agg["mlp_dropout"])

    class Ext(nn.Module):
        def __init__(self):
            super().__init__()
            mr = cfg["fe"]["mr"]["arch"]
            if name == "XR1MR3C1CnnTrf":
                self._fe0 = fe(cfg["fe"]["xr"]["arch"])
                self._fe1, self._fe2, self._fe3 = fe(mr), fe(mr), fe(mr)
                self._fe4 = FeatC1(config=cfg["fe"]["clin"])
                self._agg_1, self._agg_2, self._agg_3 = feat(ns[1], False), feat(ns[2], False), feat(ns[3], False)
                self._agg_final = feat(1 + ns[1] + ns[2] + ns[3] + ns[4], True)
            else:
                self._fe1, self._fe2, self._fe3 = fe(mr), fe(mr), fe(mr)
                self._agg_1, self._agg_2, self._agg_3 = feat(ns[0], False), feat(ns[1], False), feat(ns[2], False)
                self._agg_final = feat(ns[0] + ns[1] + ns[2], True)

        def forward(self, *ins):
            def mri(fe_, agg_, vol):
                b = vol.shape[0]
                t = rearrange(vol, "b ch r c s -> (b s) ch r c")
                t = repeat(t, "bs ch r c -> bs (k ch) r c", k=3)
                t = rearrange(fe_(t), "(b s) ch d0 d1 -> b (s d0 d1) ch", b=b)
                return agg_(t)[1]

            parts = []
            vols = ins
            if name == "XR1MR3C1CnnTrf":
                x = repeat(ins[0], "b ch r c -> b (k ch) r c", k=3)
                parts.append(rearrange(self._fe0(x), "b ch d0 d1 -> b (d0 d1) ch"))
                vols = ins[1:4]
            parts.append(mri(self._fe1, self._agg_1, vols[0]))
            parts.append(mri(self._fe2, self._agg_2, vols[1]))
            parts.append(mri(self._fe3, self._agg_3, vols[2]))
            if name == "XR1MR3C1CnnTrf":
                parts.append(self._fe4(ins[4]))
            out, _, _ = self._agg_final(torch.cat(parts, dim=1))
            return {"main": rearrange(out, "b head cls -> b (head cls)")}

    return Ext()


def build_model(name, cfg):
    """``koafusion.models.dict_models[name](config, None)`` of the unmodified reference (or the 3-MRI composition of its blocks)."""
    _ensure_path()
    from koafusion.models import dict_models

    acfg = to_attr(cfg)
    if name in dict_models:
        return dict_models[name](config=acfg, path_weights=None)
    return build_extension_model(name, acfg)
