"""Source-level guard of the programmatic-dependent-launch invariant (``KOA_PDL``, ``csrc/koa_common.cuh``): every kernel
that can be launched with the programmatic-stream-serialization attribute (through ``koa_launch_pdl``) must call
``griddep_wait()`` — a kernel that starts early and does not wait reads its predecessor's output while it is being
written — and must release its successors with ``griddep_launch_dependents()`` only after the wait (and, for the tcgen05
kernels, after the TMEM allocation). Static check on the sources; the behaviour itself is a GPU experiment."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oaprogressionmmf_b200", "csrc")
# launch sites that go through a function pointer variable: which kernel template the variable holds
ALIASES = {"gemm_api.cu": {"kern": ["gemm_conv_kernel", "gemm_wgrad_kernel"]}}


def _sources():
    return {f: open(os.path.join(CSRC, f)).read() for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))}


def _kernel_body(src_all, name):
    m = re.search(r"__global__[^;{]*?\b" + re.escape(name) + r"\s*\(", src_all, re.S)
    assert m, f"definition of {name} not found"
    i = src_all.index("{", m.end())
    depth, j = 1, i + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(src_all[j], 0)
        j += 1
    return src_all[i:j]


def test_every_pdl_launched_kernel_waits_before_it_releases():
    srcs = _sources()
    everything = "\n".join(srcs.values())
    launched = set()
    for fname, src in srcs.items():
        if fname == "koa_common.cuh":          # the helper's own definition and its comment
            continue
        for m in re.finditer(r"koa_launch_pdl\(\s*([^,]+),", src):
            expr = m.group(1).strip()
            if expr.startswith("void"):        # the template's own declaration in koa_common.cuh
                continue
            names = re.findall(r"\b([a-z_0-9]+_kernel)\b", expr)
            if not names:
                names = ALIASES.get(fname, {}).get(expr)
            assert names, f"{fname}: cannot tell which kernel `{expr}` is"
            launched.update(names)
    assert {"gemm_conv_kernel", "gemm_kmajor_kernel", "gemm_wgrad_kernel", "bn_act_fixed_kernel",
            "bn_bwd_apply_fixed_kernel", "bn_bwd_reduce_kernel", "layernorm_fwd_kernel", "layernorm_bwd_kernel",
            "attention_fwd_kernel", "attention_bwd_kernel"} == launched
    for name in sorted(launched):
        body = _kernel_body(everything, name)
        assert body.count("griddep_wait()") == 1 and body.count("griddep_launch_dependents()") == 1, name
        w, r = body.index("griddep_wait()"), body.index("griddep_launch_dependents()")
        assert w < r, f"{name}: releases its successors before it has waited for its predecessor"
        if "tmem_alloc" in body:
            assert body.index("tmem_alloc") < r, f"{name}: successors released before this CTA owns its TMEM columns"
        # nothing may touch global memory before the wait: no loads / TMA issue / atomics in the text before it
        head = body[:w]
        for token in ("__ldg(", "tma_load", "tma2_load", "atomicAdd(", "red_add", "ld_f("):
            assert token not in head, f"{name}: `{token}` before griddep_wait()"


def test_kernels_without_the_hook_are_never_launched_with_the_attribute():
    """The only place that sets the attribute is koa_launch_pdl."""
    for fname, src in _sources().items():
        if fname == "koa_common.cuh":
            assert src.count("cudaLaunchAttributeProgrammaticStreamSerialization") == 1
        else:
            assert "ProgrammaticStreamSerialization" not in src, fname
