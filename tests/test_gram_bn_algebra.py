"""The algebra behind the y3-free bottleneck tail, checked against autograd in fp64 on the CPU.

The last 1x1 convolution of a bottleneck, its train-mode BatchNorm, the residual add and the ReLU
(``/root/reference/koafusion/models/_torchvision.py:130-136``) never materialise the wide conv output ``y3 = a2 @ W3^T``:

* forward statistics come from the Gram matrix of the narrow input, ``mean_c = W3[c] . colsum(a2) / N`` and
  ``E[y3_c^2] = W3[c] . (a2^T a2 / N) . W3[c]``; the GEMM epilogue then applies scale / shift / residual / ReLU;
* backward: with ``dy3 = k0 * G - k1 - k2 * y3`` (per-channel coefficients of the BatchNorm backward),
  ``sum(G * y3)_c = (G^T a2)[c] . W3[c]``,
  ``dW3 = k0 * (G^T a2) - k1 (x) colsum(a2) - k2 * (W3 @ a2^T a2)``,
  ``d(a2) = G @ (k0 * W3) - (k1 @ W3) - a2 @ (W3^T diag(k2) W3)``.

These are the formulas `csrc/fe_engine.cu` implements with tensor-core GEMMs; this test pins them.
"""
import torch


def _reference(a2, w3, gamma, beta, res, gout, eps):
    a2 = a2.clone().requires_grad_(True)
    w3 = w3.clone().requires_grad_(True)
    gamma = gamma.clone().requires_grad_(True)
    beta = beta.clone().requires_grad_(True)
    y3 = a2 @ w3.t()
    mean = y3.mean(0)
    var = y3.var(0, unbiased=False)
    z = (y3 - mean) / torch.sqrt(var + eps) * gamma + beta
    out = torch.relu(z + res)
    out.backward(gout)
    return out.detach(), a2.grad, w3.grad, gamma.grad, beta.grad, mean.detach(), var.detach()


def test_gram_form_matches_autograd():
    torch.manual_seed(3)
    n, w, c, eps = 257, 16, 64, 1e-5
    dt = torch.float64
    a2 = torch.relu(torch.randn(n, w, dtype=dt) + 0.3)
    w3 = torch.randn(c, w, dtype=dt) / w ** 0.5
    gamma = torch.rand(c, dtype=dt) + 0.5
    beta = torch.randn(c, dtype=dt) * 0.1
    res = torch.randn(n, c, dtype=dt)
    gout = torch.randn(n, c, dtype=dt)
    out_r, da2_r, dw3_r, dgamma_r, dbeta_r, mean_r, var_r = _reference(a2, w3, gamma, beta, res, gout, eps)

    # ---- forward without y3 ----
    sa2 = a2.sum(0)
    gram = a2.t() @ a2
    q = w3 @ (gram / n)                       # [c, w]
    mean = (w3 @ sa2) / n
    ey2 = (q * w3).sum(1)
    var = ey2 - mean * mean
    invstd = 1.0 / torch.sqrt(var + eps)
    scale, shift = gamma * invstd, beta - mean * gamma * invstd
    out = torch.relu((a2 @ w3.t()) * scale + shift + res)      # "epilogue" of the second GEMM pass
    assert torch.allclose(mean, mean_r, atol=1e-12) and torch.allclose(var, var_r, atol=1e-12)
    assert torch.allclose(out, out_r, atol=1e-10)

    # ---- backward without y3 ----
    g = gout * (out > 0)                       # gradient w.r.t. the pre-ReLU sum (what the producer epilogue hands over)
    sdz = g.sum(0)
    t = g.t() @ a2                             # [c, w]: the ordinary weight-gradient GEMM, on G instead of dy3
    sdzy = (t * w3).sum(1)
    dgamma = invstd * (sdzy - mean * sdz)
    dbeta = sdz
    k0 = gamma * invstd
    k2 = gamma * invstd * invstd * dgamma / n
    k1 = k0 * dbeta / n - k2 * mean
    dw3 = k0[:, None] * t - k1[:, None] * sa2[None, :] - k2[:, None] * (n * q)
    m = w3.t() @ (k2[:, None] * w3)            # [w, w]
    da2 = g @ (k0[:, None] * w3) - (k1 @ w3)[None, :] - a2 @ m
    assert torch.allclose(dgamma, dgamma_r, atol=1e-10) and torch.allclose(dbeta, dbeta_r, atol=1e-10)
    assert torch.allclose(dw3, dw3_r, atol=1e-9)
    assert torch.allclose(da2, da2_r, atol=1e-10)


def test_hi_lo_split_of_the_gram_matrix_keeps_fp32_accuracy():
    """Q = W3 @ (Gram / N) runs on the tensor cores with fp16 operands: the normalised Gram matrix is split into an
    fp16 head and an fp16 tail (22 significant bits together); W3 is the fp16 forward operand itself."""
    torch.manual_seed(4)
    n, w, c = 4096, 64, 256
    a2 = torch.relu(torch.randn(n, w) + 0.3).half().float()
    w3 = (torch.randn(c, w) / w ** 0.5).half().float()
    gram = (a2.double().t() @ a2.double() / n)
    hi = gram.float().half()
    lo = (gram.float() - hi.float()).half()
    q = w3.double() @ hi.double() + w3.double() @ lo.double()
    q_ref = w3.double() @ gram
    rel = float((q - q_ref).norm() / q_ref.norm())
    assert rel < 2e-6, rel
    ey2 = (q * w3.double()).sum(1)
    ey2_ref = ((a2.double() @ w3.double().t()) ** 2).mean(0)
    assert float(((ey2 - ey2_ref).abs() / ey2_ref).max()) < 1e-5
