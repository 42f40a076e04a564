"""libapx.so loads without a GPU and exports every symbol include/apx.h declares; the ctypes table matches the header."""
import ctypes
import os
import re

import pytest

from ap_adapter_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "apx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    fns = _header_functions()
    assert "apx_ipattn_fwd" in fns and "apx_audiomae_fwd" in fns and "apx_ctx_kv" in fns and len(fns) >= 15


def test_library_exports_every_header_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [f for f in _header_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_ctypes_table_covers_header():
    assert sorted(_lib.SYMBOLS) == _header_functions()


def test_no_compute_queries_work_without_gpu():
    lib = _lib.load()
    assert lib.apx_version() >= 1
    assert lib.apx_get_impl() == _lib.APX_IMPL_TC
    assert lib.apx_set_impl(7) == 1 and b"unknown impl" in lib.apx_last_error()
    # sizes are pure host arithmetic
    assert lib.apx_ctx_kv_bytes(2, 8, 128, 256, 8) > 2 * 2 * 136 * 256 * 2
    assert lib.apx_ctx_kv_bytes(2, 8, 128, 250, 8) == 0  # C not divisible by H
    assert lib.apx_ipattn_workspace_bytes(2, 1000, 256) >= 2 * 2 * 1000 * 256 * 2
    assert lib.apx_audiomae_workspace_bytes(1) > 513 * 3072 * 2
    # packed-weight size: the all-heads kernel's sections (Wq as [C/64][8][H*d_pad][8] + Wo as [H*d_pad/8][C][8], one pair per
    # supported cluster size 1 / 2 / 4; C <= 256 only) + the one-head kernel's sections (Wq [H][C/64][8][d_pad][8] + Wo [H][Nq/8][d_pad][8])
    assert lib.apx_ipattn_pack_bytes(256, 8) == 3 * 2 * 256 * 256 * 2 + 2 * 256 * 256 * 2
    assert lib.apx_ipattn_pack_bytes(128, 8) == 2 * 2 * 128 * 128 * 2 + 2 * 128 * 128 * 2
    assert lib.apx_ipattn_pack_bytes(384, 8) == 384 * 384 * 2 * 2 and lib.apx_ipattn_pack_bytes(640, 8) == 640 * 640 * 2 * 2
    assert lib.apx_ipattn_pack_bytes(320, 8) == 384 * 320 * 2 + 384 * 384 * 2         # d = 40 -> d_pad = 48: Nq = 384
    assert lib.apx_ipattn_pack_bytes(512, 16) == 0                                    # 16 heads: neither fused kernel
    # coverage: contexts of up to 256 padded tokens fuse at every AudioLDM2 width; the un-pooled 8 + 512 context takes three kernels
    assert lib.apx_ipattn_fused_ok(256, 8, 8, 128) == 1 and lib.apx_ipattn_fused_ok(256, 8, 8, 8) == 1
    assert lib.apx_ipattn_fused_ok(384, 8, 8, 128) == 1 and lib.apx_ipattn_fused_ok(640, 8, 8, 128) == 1
    assert lib.apx_ipattn_fused_ok(320, 8, 77, 8) == 1
    assert lib.apx_ipattn_fused_ok(256, 8, 8, 512) == 0 and lib.apx_ipattn_fused_ok(512, 16, 8, 32) == 0
    # kernel choice (measured policy, apx_api.cu::pick_site_kernel): all-heads wherever it applies, one-head for small grids at
    # C <= 384 where it does not, three kernels at C = 640 and for large grids
    assert lib.apx_ipattn_site_kernel(2, 1000, 256, 8, 8, 128) == 1 and lib.apx_ipattn_site_kernel(16, 1000, 256, 8, 8, 128) == 1
    assert lib.apx_ipattn_site_kernel(2, 252, 384, 8, 8, 128) == 2 and lib.apx_ipattn_site_kernel(16, 252, 384, 8, 8, 128) == 0
    assert lib.apx_ipattn_site_kernel(1, 100, 192, 4, 8, 128) == 2   # narrow site on a small grid: the one-head cluster kernel
    assert lib.apx_ipattn_site_kernel(2, 64, 640, 8, 8, 128) == 0
    assert lib.apx_set_site_variant(_lib.APX_SITE_ONEHEAD) == 0
    try:
        assert lib.apx_ipattn_site_kernel(2, 64, 640, 8, 8, 128) == 2 and lib.apx_ipattn_site_kernel(2, 1000, 256, 8, 8, 128) == 2
    finally:
        assert lib.apx_set_site_variant(_lib.APX_SITE_AUTO) == 0
    assert lib.apx_set_site_variant(9) == 1


def test_null_pointers_are_rejected_not_dereferenced():
    lib = _lib.load()
    rc = lib.apx_ipattn_fwd(None, None, None, None, None, None, None, 2, 64, 256, 8, 8, 8, 0.5, 0, None, 0, None)
    assert rc == 1 and b"null pointer" in lib.apx_last_error()
    rc = lib.apx_linear(None, 0, None, 0, None, None, 0, None, 0, 1, 64, 64, 0, 0, 0, None)
    assert rc == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ap_adapter_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_cpu_tensors_fail_loudly():
    import torch
    from ap_adapter_b200.attention_processor import IPAttnProcessor2_0
    from oracle.ref_shims import StubAttention

    proc = IPAttnProcessor2_0(256, "t", 768, num_tokens=8, scale=0.5)
    attn = StubAttention(256, 768, 8)
    with torch.no_grad(), pytest.raises(_lib.ApxError, match="no CPU fallback"):
        proc(attn, torch.randn(2, 16, 256), encoder_hidden_states=torch.randn(2, 16, 768))


def test_site_kernel_policy_is_consistent_over_the_shape_space():
    """Host-side sweep (no launch): whatever `apx_ipattn_site_kernel` picks for (Bc, S, C, H, Lt, La) is a form that exists for that shape —
    a fused choice implies `apx_ipattn_fused_ok` and a packed-weight blob — and the size queries answer for every shape the UNet /
    the training collate can produce (C = 256 / 384 / 640, 8 heads, 8 text tokens, La = 8 ... 512: attention_processor.py:400-403,
    train_apadapter_v2.py:445-446)."""
    import itertools
    lib = _lib.load()
    seen = set()
    for c, h in itertools.product(range(64, 1281, 64), (1, 2, 4, 8, 16)):
        pack = lib.apx_ipattn_pack_bytes(c, h)
        for lt, la in itertools.product((0, 1, 8, 16, 77), (0, 1, 8, 32, 113, 128, 200, 248, 256, 512)):
            if lt + la == 0:
                continue
            fused = lib.apx_ipattn_fused_ok(c, h, lt, la)
            for bc, s in itertools.product((1, 2, 16), (1, 64, 252, 1000, 4096)):
                k = lib.apx_ipattn_site_kernel(bc, s, c, h, lt, la)
                assert k in (_lib.APX_SITE_AUTO, _lib.APX_SITE_ALLHEADS, _lib.APX_SITE_ONEHEAD), (bc, s, c, h, lt, la, k)
                assert k == 0 or (fused == 1 and pack > 0), (bc, s, c, h, lt, la, k, fused, pack)
                seen.add(k)
    assert seen == {0, 1, 2}
    for c, la in itertools.product((256, 384, 640), (8, 32, 128, 512)):
        assert lib.apx_ctx_kv_bytes(16, 8, la, c, 8) >= 16 * 2 * (8 + la) * c * 2
        assert lib.apx_ctx_kv_workspace_bytes(16, 8, la, c) > 0 and lib.apx_ipattn_workspace_bytes(16, 1000, c) > 0
        assert lib.apx_ipattn_bwd_workspace_bytes(16, 1000, c, la) > 0


def test_every_compute_entry_rejects_null_arguments_before_touching_the_device():
    """Error behaviour of the boundary: every compute entry point of include/apx.h validates its arguments on the host and returns
    APX_EINVAL (1) with a message naming itself — nothing is dereferenced, nothing is launched, nothing throws across the ABI
    (runs here without a GPU)."""
    lib = _lib.load()
    queries = {"apx_last_error", "apx_version", "apx_device_ok", "apx_set_impl", "apx_get_impl", "apx_set_site_variant", "apx_launch_count",
               "apx_ipattn_fused_ok", "apx_ipattn_site_kernel"}
    checked = 0
    for name, (_res, argtypes) in _lib.SYMBOLS.items():
        if name in queries or name.endswith("_bytes"):
            continue
        args = [None if t in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(t, ctypes._Pointer)
                else 1.0 if t in (ctypes.c_float, ctypes.c_double) else 0 for t in argtypes]
        rc = getattr(lib, name)(*args)
        msg = lib.apx_last_error().decode()
        assert rc == 1, (name, rc, msg)
        assert msg.startswith("apx_") and ("null pointer" in msg or "bad argument" in msg), (name, msg)
        checked += 1
    assert checked >= 20


def test_error_codes_by_class_for_bad_shapes():
    """include/apx.h: APX_EINVAL (1) for arguments that are wrong whatever the kernel, APX_EUNSUPPORTED (2) for shapes / dtypes outside
    what the sm_100a kernels implement, APX_EWORKSPACE (4) for a short caller workspace — all decided on the host, before any launch
    (non-null dummy pointers are never dereferenced on these paths; runs without a GPU)."""
    lib = _lib.load()
    p = ctypes.c_void_p(0x10000)
    EINVAL, EUNSUPPORTED, EWORKSPACE = 1, 2, 4

    def fwd(bc, s, c, h, lt, la, dtype=0, wsb=1 << 24):
        return lib.apx_ipattn_fwd(p, p, p, p, None, p, p, bc, s, c, h, lt, la, 0.5, dtype, p, wsb, None)

    assert fwd(2, 64, 250, 8, 8, 8) == EINVAL and b"bad shape" in lib.apx_last_error()          # C % H != 0
    assert fwd(0, 64, 256, 8, 8, 8) == EINVAL
    assert fwd(2, 64, 256, 8, 0, 0) == EINVAL                                                     # no context tokens at all
    assert fwd(2, 64, 256, 8, 8, 8, dtype=7) == EUNSUPPORTED and b"dtype" in lib.apx_last_error()
    assert fwd(2, 64, 256, 8, 8, 8, wsb=16) == EWORKSPACE and b"workspace" in lib.apx_last_error()
    ln = lambda a, rows, c: lib.apx_add_layernorm(a, None, p, p, 0, None, p, rows, c, 1e-5, 0, None)  # noqa: E731
    assert ln(p, 10, 100) == EUNSUPPORTED and ln(p, 10, 4096) == EUNSUPPORTED
    assert ln(ctypes.c_void_p(0x10004), 10, 256) == EINVAL and b"aligned" in lib.apx_last_error()
    gn = lambda c, groups: lib.apx_groupnorm_nhwc(p, p, p, None, None, 0, p, 2, 100, c, groups, 1e-5, 1, 0, None)  # noqa: E731
    assert gn(250, 32) == EINVAL and gn(64, 32) == EUNSUPPORTED
    assert lib.apx_flash_attention(p, 0, 192, p, 0, 192, p, 0, 192, p, 0, 192, 1, 64, 64, 192, 8, 0, None) == EUNSUPPORTED   # d = 24
    assert b"head dim 24" in lib.apx_last_error()
