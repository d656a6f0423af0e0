"""Shared helpers of the parity tests."""
import torch


def rel(got, ref):
    """Relative L2 error ||got - ref|| / ||ref|| (the "relative error" of BASELINE.json's tolerance)."""
    got = got.detach().double().flatten()
    ref = ref.detach().double().flatten()
    return float((got - ref).norm() / (ref.norm() + 1e-30))


class AttrDict(dict):
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


def fresh_params(sd):
    """Leaf copies (requires_grad) of the float entries of an oracle state_dict, running stats excluded."""
    out = {}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            out[k] = v
            v.requires_grad_(True)
            v.grad = None
    return out


def bf16_ulp_tol():
    return 2.0 ** -8
