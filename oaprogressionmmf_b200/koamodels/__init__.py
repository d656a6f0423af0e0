"""Host-side mirror of ``koafusion.models`` (reference: koafusion/models/__init__.py:1-15): same exported
names, same registries; the compute behind every class is libkoa_b200.so."""
from ._feat import Attention, FeaT, FeedForward, Transformer
from ._fe import KoaResNet, SliceEncoder, dict_fes
from ._models import (MR1CnnTrf, MR2CnnTrf, MR3CnnTrf, XR1Cnn, XR1MR1CnnTrf, XR1MR2C1CnnTrf, XR1MR2CnnTrf,
                      XR1MR3C1CnnTrf, dict_models, set_branch_streams)
from ._small import FeatC1

__all__ = ["Transformer", "FeaT", "FeedForward", "Attention", "XR1Cnn", "MR1CnnTrf", "MR2CnnTrf", "XR1MR1CnnTrf",
           "XR1MR2CnnTrf", "XR1MR2C1CnnTrf", "MR3CnnTrf", "XR1MR3C1CnnTrf", "FeatC1", "dict_models", "dict_fes",
           "KoaResNet", "SliceEncoder", "set_branch_streams"]
