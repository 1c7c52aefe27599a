"""CPU-only checks of the boundary: the C-ABI library builds, loads without a GPU, exports every symbol that
include/pio_b200.h declares, the ctypes structs match the header's field order, and the host mirror keeps the
reference's constructor signatures and state_dict layout."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from perceiverio_pytorch_b200 import _lib, build
    build.build()
    return _lib.load(build_if_missing=False)


def _header():
    return open(os.path.join(ROOT, "include", "pio_b200.h")).read()


def test_library_exports_every_declared_symbol(lib):
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(pio_\w+)\s*\(", _header(), flags=re.M))
    from perceiverio_pytorch_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.pio_abi_version() == 15


@pytest.mark.parametrize("struct,cname", [("LayerNormArgs", "pio_layernorm_args"), ("GemmArgs", "pio_gemm_args"),
                                          ("SoftmaxArgs", "pio_softmax_args"), ("AttentionArgs", "pio_attention_args"),
                                          ("CombineArgs", "pio_combine_args"), ("LinearF32Args", "pio_linear_f32_args"),
                                          ("LayerNormConcatArgs", "pio_layernorm_concat_args")])
def test_ctypes_structs_follow_the_header(struct, cname):
    from perceiverio_pytorch_b200 import _lib
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), _header(), flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.split(",")
        first = parts[0].split()[-1].lstrip("*")
        names.append(first)
        names.extend(p.strip().lstrip("*") for p in parts[1:])
    assert [f[0] for f in getattr(_lib, struct)._fields_] == names


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected with a status + message before any CUDA call (no GPU needed)."""
    from perceiverio_pytorch_b200 import _lib
    a = _lib.GemmArgs()
    rc = lib.pio_gemm_bf16(ctypes.byref(a), None)
    assert rc == -1
    assert b"null operand" in lib.pio_last_error()
    assert lib.pio_attention_supported(128, 128) == 0
    assert lib.pio_attention_supported(261, 261) == 0
    assert lib.pio_attention_supported(1024, 1024) == -2
    assert lib.pio_attention_key_tile(128, 128, 0) == 128
    assert lib.pio_attention_key_tile(322, 322, 1) == 64
    c = _lib.LayerNormConcatArgs()
    assert lib.pio_layernorm_concat_bf16(ctypes.byref(c), None) == -1 and b"null pointer" in lib.pio_last_error()
    buf = (ctypes.c_float * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    c = _lib.LayerNormConcatArgs(ptr, 0, 0, 0, ptr, ptr, 264, None, None, 1, 50177, 3, 258, 1e-5)
    assert lib.pio_layernorm_concat_bf16(ctypes.byref(c), None) == -1 and b"multiple of 4" in lib.pio_last_error()
    c = _lib.LayerNormConcatArgs(ptr, 0, 0, 0, ptr, ptr, 261, None, None, 1, 50176, 3, 258, 1e-5)
    assert lib.pio_layernorm_concat_bf16(ctypes.byref(c), None) == -1 and b"pad8" in lib.pio_last_error()


def test_layernorm_concat_supported_shapes():
    from perceiverio_pytorch_b200 import ops
    assert ops.layernorm_concat_supported(64, 50176, 3, 258)        # ImageNet pixels
    assert ops.layernorm_concat_supported(1, 182528, 64, 258)       # optical flow
    assert not ops.layernorm_concat_supported(1, 50177, 3, 258)     # positions not a multiple of 4
    assert not ops.layernorm_concat_supported(1, 1024, 3, 1300)     # too wide
    assert not ops.layernorm_concat_supported(4096, 1024, 64, 258)  # batch x features beyond one position group


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "perceiverio_pytorch_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


REF_SIGNATURES = {
    # transformer_primitives.py:34-45, :193-198, :233-242, :318-330; perceiver.py:34-50, :123-143
    "Attention": ["q_in_channels", "k_in_channels", "v_in_channels", "num_heads", "init_scale", "with_final_bias",
                  "final_init_scale_multiplier", "dropout_prob", "qk_out_channels", "v_out_channels",
                  "output_channels"],
    "MLP": ["in_channels", "out_channels", "widening_factor", "dropout_prob", "init_scale"],
    "SelfAttention": ["in_channels", "widening_factor", "dropout_prob", "dropout_attn_prob", "num_heads",
                      "att_init_scale", "dense_init_scale", "qk_channels", "v_channels"],
    "CrossAttention": ["q_in_channels", "kv_in_channels", "widening_factor", "dropout_prob", "dropout_attn_prob",
                       "num_heads", "attn_init_scale", "mlp_init_scale", "shape_for_attn", "use_query_residual",
                       "qk_channels", "v_channels"],
    "PerceiverEncoder": ["num_input_channels", "num_self_attends_per_block", "num_blocks", "num_latents",
                         "num_latent_channels", "qk_channels", "v_channels", "num_cross_attend_heads",
                         "num_self_attend_heads", "cross_attend_widening_factor", "self_attend_widening_factor",
                         "dropout_prob", "latent_pos_enc_init_scale", "cross_attention_shape_for_attn",
                         "use_query_residual"],
    "PerceiverDecoder": ["query_channels", "final_project_out_channels", "num_latent_channels", "qk_channels",
                         "v_channels", "use_query_residual", "output_w_init", "num_heads", "final_project"],
}


@pytest.mark.parametrize("cls", sorted(REF_SIGNATURES))
def test_constructor_signatures_match_reference(cls):
    import perceiverio_pytorch_b200 as pio
    got = [p for p in inspect.signature(getattr(pio, cls).__init__).parameters if p != "self"]
    assert got == REF_SIGNATURES[cls]


def test_constructor_signatures_match_live_reference():
    """Same check against the reference itself when it is mounted (build container)."""
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    if ref is None:
        pytest.skip("reference tree not mounted")
    import perceiverio_pytorch_b200 as pio
    for cls in REF_SIGNATURES:
        mod = ref.primitives if hasattr(ref.primitives, cls) else ref.perceiver
        want = inspect.signature(getattr(mod, cls).__init__)
        got = inspect.signature(getattr(pio, cls).__init__)
        assert [(p.name, p.default) for p in want.parameters.values()] == \
               [(p.name, p.default) for p in got.parameters.values()], cls


def test_state_dict_layout_matches_golden_fixture():
    """Our modules load the reference's state_dict strictly (names, shapes) and round-trip it unchanged."""
    import json
    import perceiverio_pytorch_b200 as pio
    from golden_util import golden_names, load_golden
    for name in golden_names():
        params, _, meta, _ = load_golden(name)
        ctor = json.loads(meta["ctor"])
        m = getattr(pio, ctor["cls"])(**ctor["kwargs"])
        m.load_state_dict(params, strict=True)
        sd = m.state_dict()
        assert list(sd.keys()) == list(params.keys()), name
        for k in params:
            assert torch.equal(sd[k], params[k]), (name, k)


def test_value_errors_match_reference():
    import perceiverio_pytorch_b200 as pio
    with pytest.raises(ValueError):
        pio.Attention(q_in_channels=30, k_in_channels=30, v_in_channels=30, num_heads=8)
    with pytest.raises(ValueError):
        pio.CrossAttention(q_in_channels=8, kv_in_channels=8, shape_for_attn="x")
    with pytest.raises(ValueError):
        pio.PerceiverEncoder(num_input_channels=8, num_latent_channels=30)
    with pytest.raises(ValueError):
        pio.PerceiverDecoder(query_channels=8, final_project_out_channels=8, output_w_init="ones")


def test_mask_helper_matches_reference_semantics():
    import perceiverio_pytorch_b200 as pio
    qm = torch.tensor([[True, False, True]])
    km = torch.tensor([[True, True, False, False]])
    m = pio.make_cross_attention_mask(qm, km)
    assert m.shape == (1, 3, 4)
    assert torch.equal(m, qm[:, :, None] & km[:, None, :])
    from perceiverio_pytorch_b200.primitives import _factor_mask
    rk, k2 = _factor_mask(m)
    assert torch.equal(k2, km) and torch.equal(rk, qm)
    dense = m.clone()  # no attached factors: must be recovered from the dense matrix
    rk, k2 = _factor_mask(dense)
    assert torch.equal(k2, km) and torch.equal(rk, qm)
    dense[0, 0, 2] = True  # no longer an outer product: not factorable, routed to the general (dense-mask) path
    assert _factor_mask(dense) is None
    from perceiverio_pytorch_b200.primitives import _route_mask
    rk, k2, general = _route_mask(dense, None, False, B=1, H=2, Nq=3, Nk=4, device="cpu")
    assert rk is None and k2 is None and general.dense_mask.dtype == torch.uint8
    assert torch.equal(general.dense_mask.bool(), dense) and general.bias is None and general.matrix is None
    # an additive bias is carried as a stride-0 broadcast view; return_matrix allocates [B, H, Nq, Nk]
    rk, k2, general = _route_mask(m, torch.zeros(1, 1, 3, 4), True, B=1, H=2, Nq=3, Nk=4, device="cpu")
    assert general.bias.shape == (1, 2, 3, 4) and general.bias.stride(1) == 0
    assert general.matrix.shape == (1, 2, 3, 4)
    rk, k2, general = _route_mask(m, None, False, B=1, H=2, Nq=3, Nk=4, device="cpu")
    assert general is None and torch.equal(k2, km)


def test_swap_hot_path_on_live_reference_keeps_state_dict():
    """install.swap_hot_path replaces _encoder/_decoder of a reference wrapper and keeps state_dict identical."""
    from oracle import ref_shim
    ns = ref_shim.load_wrappers()
    if ns is None:
        pytest.skip("reference tree not mounted")
    from perceiverio_pytorch_b200 import install
    import perceiverio_pytorch_b200 as pio
    torch.manual_seed(0)
    model = ns.language.LanguagePerceiver(num_self_attends_per_block=2, num_latents=32, num_latent_channels=64).eval()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    enc_ref, dec_ref = model.perceiver._encoder, model.perceiver._decoder
    install.swap_hot_path(model)
    assert isinstance(model.perceiver._encoder, pio.PerceiverEncoder)
    assert isinstance(model.perceiver._decoder, pio.PerceiverDecoder)
    after = model.state_dict()
    assert list(after.keys()) == list(before.keys())
    for k in before:
        assert torch.equal(before[k], after[k]), k
    # fuse_input patches PerceiverIO.forward on the instance only; the language model's embedding preprocessor is not an
    # image preprocessor, so the patched forward must fall back to the class's own forward (no recursion)
    install.swap_hot_path(model, fuse_input=True)
    assert "forward" in model.perceiver.__dict__ and list(model.state_dict().keys()) == list(before.keys())
    model.perceiver._encoder, model.perceiver._decoder = enc_ref, dec_ref      # reference modules back: runs on CPU
    with torch.inference_mode():
        ids = torch.randint(6, 262, (1, 2048))          # the recipe's fixed sequence length (language_perceiver.py:24-46)
        mask = torch.ones(1, 2048, dtype=torch.bool)
        out_patched = model(ids, mask)
        del model.perceiver.__dict__["forward"]
        out_plain = model(ids, mask)
    assert torch.equal(out_patched, out_plain)
