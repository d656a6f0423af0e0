"""Checkpoint interoperability with the unmodified reference: a checkpoint written from this
package's model by the reference's own ``CheckpointHandler`` (``koafusion/various/_checkpoint.py:51-62``) loads, strict,
into the reference's model class, and a checkpoint written from the reference model restores this package's model through
the ``restore_weights`` / ``path_weights`` constructor path (``koafusion/models/_xrNmrMcP.py:181-182``).

CPU only and only in this container: the reference sources are imported from ``/root/reference`` (by file path where the
package cannot be imported); the test is skipped where that tree does not exist (the GPU box)."""
import importlib.util
import os
import sys

import pytest
import torch

from oaprogressionmmf_b200 import koamodels
from oracle import koa_oracle as ko
from tests.util import to_attr

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "koafusion")), reason="reference tree not present")


def _ref_checkpoint_module():
    spec = importlib.util.spec_from_file_location("ref_checkpoint", os.path.join(REF, "koafusion/various/_checkpoint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_models():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from koafusion.models import dict_models  # noqa: E402  (imports torch / torchvision / einops only)

    return dict_models


@pytest.mark.parametrize("name", ["XR1Cnn", "XR1MR2C1CnnTrf"])
def test_checkpoint_roundtrip_with_the_reference_handler(name, tmp_path):
    cfg = ko.make_config(name, xr_size=64, mr_size=64, slices=(3, 2, 2), depth=1)
    torch.manual_seed(778)
    mine = koamodels.dict_models[name](to_attr(cfg), None)
    with torch.no_grad():                      # BatchNorm buffers away from their initial values
        for k, v in mine.state_dict().items():
            if k.endswith("running_mean"):
                v.normal_()
            elif k.endswith("running_var"):
                v.uniform_(0.5, 2.0)
            elif k.endswith("num_batches_tracked"):
                v.fill_(7)
    handler = _ref_checkpoint_module().CheckpointHandler(path_root=str(tmp_path))
    handler.save_new_ckpt(mine, model_name=name, fold_idx=0, epoch_idx=3)      # unwrapped module: AttributeError branch
    path = handler.get_last_ckpt()
    assert os.path.basename(path) == f"{name}__fold_0__epoch_003.pth"

    ref = _ref_models()[name](to_attr(cfg), None)
    ref.load_state_dict(torch.load(path), strict=True)
    for (ka, va), (kb, vb) in zip(mine.state_dict().items(), ref.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka

    # the other direction, through the constructor: config.restore_weights + path_weights
    torch.manual_seed(5)
    ref2 = _ref_models()[name](to_attr(cfg), None)
    path2 = os.path.join(tmp_path, "from_reference.pth")
    torch.save(ref2.state_dict(), path2)
    restored = koamodels.dict_models[name](to_attr(dict(cfg, restore_weights=True)), path2)
    for (ka, va), (kb, vb) in zip(restored.state_dict().items(), ref2.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka

    # wrapped model (``model.module.state_dict()`` branch of save_new_ckpt): keys without the "module." prefix
    class Wrapped(torch.nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

    handler.save_new_ckpt(Wrapped(mine), model_name=name, fold_idx=1, epoch_idx=4)
    sd = torch.load(handler.get_last_ckpt())
    assert list(sd) == list(mine.state_dict()) and not any(k.startswith("module.") for k in sd)
