"""The C-ABI boundary: every function declared in include/koa_b200.h is exported by libkoa_b200.so and bound in
oaprogressionmmf_b200/_lib.py; no torch types cross it; the product never routes through the oracle. CPU only (no
compute entry point is called)."""
import ctypes
import os
import re

from oaprogressionmmf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "koa_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(koa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("koa_fe_forward", "koa_fe_backward", "koa_feat_forward", "koa_feat_backward", "koa_gemm_bf16",
                 "koa_conv_fprop_bf16", "koa_conv_wgrad_bf16", "koa_gemm_wgrad_bf16", "koa_focal_loss",
                 "koa_layernorm_fwd", "koa_attention_fwd", "koa_maxpool_fwd", "koa_stem_pack", "koa_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in koa_b200.h but not exported"


def test_python_binding_covers_every_declared_symbol():
    assert set(_declared()) == set(_lib.SIGNATURES), (set(_declared()) ^ set(_lib.SIGNATURES))
    lib = _lib.load()
    assert lib.koa_version() >= 1


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # comments may mention PyTorch; declarations may not
    for bad in ("torch", "at::", "c10::", "Tensor", "std::"):
        assert bad not in code


def test_struct_mirrors_match_header_field_order():
    src = open(HEADER).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (struct, struct), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                out.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
        return out

    assert fields("koa_epilogue") == [f[0] for f in _lib.Epilogue._fields_]
    assert fields("koa_fe_desc") == [f[0] for f in _lib.FeDesc._fields_]
    assert fields("koa_feat_desc") == [f[0] for f in _lib.FeatDesc._fields_]
    assert fields("koa_adam_tensor") == [f[0] for f in _lib.AdamTensor._fields_]
    assert fields("koa_adam_hyper") == [f[0] for f in _lib.AdamHyper._fields_]
    assert ctypes.sizeof(_lib.AdamTensor) == 40 and ctypes.sizeof(_lib.AdamHyper) == 56
    assert fields("koa_augment") == [f[0] for f in _lib.Augment._fields_] and ctypes.sizeof(_lib.Augment) == 40


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oaprogressionmmf_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "koa_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn


def test_workspace_queries_work_without_a_gpu():
    """Planning entry points are pure host code: sizes for the full-size DESS extractor (B=16)."""
    lib = _lib.load()
    d = _lib.FeDesc(arch=_lib.ARCH_IDS["resnet50"], n_img=16 * 64, h=160, w=160, slices=64, with_gap=1, training=1,
                    need_backward=1)
    nbytes = lib.koa_fe_workspace_bytes(ctypes.byref(d))
    assert 10e9 < nbytes < 120e9
    c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.koa_fe_out_shape(ctypes.byref(d), ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)) == 0
    assert (c.value, h.value, w.value) == (2048, 5, 5)
    assert lib.koa_fe_num_units(ctypes.byref(d)) == 53
    bad = _lib.FeDesc(arch=7, n_img=1, h=160, w=160)
    assert lib.koa_fe_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"arch" in lib.koa_last_error()
    f = _lib.FeatDesc(batch=16, n_patches=123, dim=2048, depth=4, heads=8, mlp_dim=2048, num_classes=2, with_cls=1,
                      compute_head=1, training=1, need_backward=1)
    assert lib.koa_feat_workspace_bytes(ctypes.byref(f)) > 0
    assert lib.koa_feat_num_params(ctypes.byref(f)) == 4 + 44 + 6
    f.n_patches = 200
    assert lib.koa_feat_workspace_bytes(ctypes.byref(f)) == 0  # > 128 tokens is rejected, not silently wrong
