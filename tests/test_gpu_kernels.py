"""Per-kernel numerics on the B200: every C-ABI entry point against a plain PyTorch fp32 evaluation of the same op
(the floating-point kernels' own reference; the oracle covers the composed path in test_gpu_parity.py).
Tolerances: fp32 outputs of bf16-operand products 2e-3 of the output range, bf16 outputs 1e-2 (bf16 rounding)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    d = (got.double() - ref.double()).abs().max()
    return float(d / ref.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("rows,c", [(1000, 261), (513, 1024), (77, 322), (300, 1280), (64, 32), (50, 1026), (40003, 261),
                                    (30001, 322), (9001, 1026), (3, 261)])
def test_layernorm_cast(rows, c):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(rows, c, device="cuda") * 2 + 0.5
    g, b = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
    y = ops.layernorm_bf16(x, g, b)
    ref = torch.nn.functional.layer_norm(x, (c,), g, b, 1e-5)
    assert _rel(y[:, :c].float(), ref) < 8e-3
    if y.shape[1] > c:
        assert float(y[:, c:].float().abs().max()) == 0.0   # pad columns are exact zeros


def _gemm(M, N, K, *, batch=1, b_mn=False, bias_mode=0, act=0, alpha=1.0, residual=False, res_bcast=False,
          out="f32", kernel=0, cluster_m=None, tile_n=0):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(1)
    ldk = ops.pad8(K)
    A = torch.full((batch, M, ldk), 7.0, dtype=torch.bfloat16, device="cuda")   # pad = garbage that must not be read
    A[:, :, :K] = torch.randn(batch, M, K, device="cuda").to(torch.bfloat16)
    if not b_mn:
        Bm = torch.full((batch, N, ldk), 5.0, dtype=torch.bfloat16, device="cuda")
        Bm[:, :, :K] = torch.randn(batch, N, K, device="cuda").to(torch.bfloat16)
        ref = A[:, :, :K].float() @ Bm[:, :, :K].float().transpose(1, 2)
        ldb, strideB = ldk, N * ldk
    else:
        ldn = ops.pad8(N)
        Bm = torch.full((batch, K, ldn), 5.0, dtype=torch.bfloat16, device="cuda")
        Bm[:, :, :N] = torch.randn(batch, K, N, device="cuda").to(torch.bfloat16)
        ref = A[:, :, :K].float() @ Bm[:, :, :N].float()
        ldb, strideB = ldn, K * ldn
    ref = ref * alpha
    bias = None
    if bias_mode == 1:
        bias = torch.randn(N, device="cuda")
        ref = ref + bias
    elif bias_mode == 2:
        bias = torch.randn(M, device="cuda")
        ref = ref + bias[None, :, None]
    if act:
        ref = torch.nn.functional.gelu(ref)
    res = None
    if residual:
        res = torch.randn(1 if res_bcast else batch, M, N, device="cuda")
        ref = ref + res
    ldo = (N + 3) // 4 * 4 if out == "f32" else ops.pad8(N)
    o32 = torch.full((batch, M, ldo), float("nan"), device="cuda") if out == "f32" else None
    o16 = torch.zeros(batch, M, ldo, dtype=torch.bfloat16, device="cuda") if out == "bf16" else None
    ops.gemm(A, Bm, M=M, N=N, K=K, batch=batch, b_mn_major=b_mn, strideA=M * ldk, strideB=strideB, lda=ldk, ldb=ldb,
             bias=bias, bias_mode=bias_mode, act=act, alpha=alpha, residual=res, ldr=N,
             strideR=0 if res_bcast else M * N, out_f32=o32, ldo32=ldo, strideO32=M * ldo, out_bf16=o16, ldo16=ldo,
             strideO16=M * ldo, tile_n=tile_n, cluster_m=cluster_m, kernel=kernel)
    torch.cuda.synchronize()
    if out == "f32":
        got = o32[:, :, :N]
        assert not bool(torch.isnan(got).any())
        assert _rel(got, ref) < 2e-3
    else:
        assert _rel(o16[:, :, :N].float(), ref) < 1e-2


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("case", [
    dict(M=256, N=256, K=64), dict(M=256, N=256, K=64, out="bf16"), dict(M=1000, N=1000, K=1024, bias_mode=1, residual=True),
    dict(M=512, N=1024, K=1024, bias_mode=1, act=1, out="bf16"), dict(M=333, N=264, K=261),
    dict(M=300, N=328, K=200, bias_mode=2, alpha=0.25, out="bf16"),
    dict(M=512, N=1024, K=256, batch=5, bias_mode=1, residual=True, res_bcast=True),
    dict(M=200, N=304, K=72, batch=3, bias_mode=1, residual=True), dict(M=4096, N=3072, K=1024, out="bf16"),
])
def test_gemm_both_kernels(kernel, case):
    _gemm(kernel=kernel, **case)


@pytest.mark.parametrize("case", [
    dict(M=77, N=2, K=322), dict(M=333, N=261, K=261), dict(M=100, N=2, K=322, bias_mode=1),
    dict(M=1000, N=1024, K=512, b_mn=True), dict(M=300, N=322, K=2048, b_mn=True, batch=2),
    dict(M=640, N=64, K=200, cluster_m=2, tile_n=64), dict(M=700, N=704, K=1000, b_mn=True, cluster_m=2),
    dict(M=130, N=261, K=1000, residual=True, bias_mode=1),          # unaligned fp32 rows: single-CTA kernel only
])
def test_gemm_single_cta_shapes(case):
    _gemm(kernel=0, **case)


def test_gemm_pair_kernel_rejects_unaligned_output():
    from perceiverio_pytorch_b200 import ops
    A = torch.zeros(256, 64, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(261, 64, dtype=torch.bfloat16, device="cuda")
    o = torch.zeros(256, 261, device="cuda")
    with pytest.raises(RuntimeError, match="CTA-pair"):
        ops.gemm(A, W, M=256, N=261, K=64, out_f32=o, ldo32=261, kernel=2)


@pytest.mark.parametrize("b,r,c", [(2, 40, 512), (1, 7, 52097), (3, 130, 785), (2, 300, 2048), (1, 33, 256)])
def test_softmax(b, r, c):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    S = torch.randn(b, r, c, device="cuda") * 3
    km = (torch.rand(b, c, device="cuda") > 0.3).to(torch.uint8)
    rk = (torch.rand(b, r, device="cuda") > 0.2).to(torch.uint8)
    P = ops.softmax_bf16(S, c, 0.37, km, rk)
    ref = torch.softmax(torch.where(km[:, None, :].bool(), S * 0.37, torch.tensor(float("-inf"), device="cuda")), -1)
    ref = ref * rk[:, :, None]
    assert _rel(P[:, :, :c].float(), ref) < 1e-2
    # row pitch padded to a multiple of 4 floats (what the engine allocates): rows <= 2048 columns take the
    # warp-per-row register kernel; pad columns of P must come out as zeros
    lds = (c + 3) // 4 * 4
    Sp = torch.full((b, r, lds), 1e30, device="cuda")     # garbage in the pad must never be used
    Sp[:, :, :c] = S
    P2 = ops.softmax_bf16(Sp, c, 0.37, km, rk)
    assert _rel(P2[:, :, :c].float(), ref) < 1e-2
    assert float(P2[:, :, c:].float().abs().max()) == 0.0 if P2.shape[2] > c else True
    km0 = km.clone()
    km0[0] = 0                                            # a sample with no valid key at all: zero rows
    P3 = ops.softmax_bf16(Sp, c, 0.37, km0, None)
    assert float(P3[0].float().abs().max()) == 0.0


def _attn_ref(q, k, v, scale, km=None, rk=None):
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if km is not None:
        s = torch.where(km[:, None, None, :].bool(), s, torch.tensor(float("-inf"), device=s.device))
    p = torch.nan_to_num(torch.softmax(s, -1), nan=0.0)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v)
    if rk is not None:
        o = o * rk[:, :, None, None]
    return o.reshape(o.shape[0], o.shape[1], -1)


@pytest.mark.parametrize("case", [
    dict(B=1, H=1, Nq=128, Nk=1024, dqk=64, dv=64, qscale=4.0), dict(B=1, H=1, Nq=100, Nk=300, dqk=64, dv=64),
    dict(B=2, H=8, Nq=512, Nk=512, dqk=128, dv=128), dict(B=2, H=8, Nq=256, Nk=256, dqk=32, dv=160),
    dict(B=1, H=16, Nq=2048, Nk=2048, dqk=32, dv=32), dict(B=1, H=8, Nq=784, Nk=784, dqk=64, dv=64),
    dict(B=2, H=8, Nq=256, Nk=2048, dqk=32, dv=160, mask=True), dict(B=2, H=8, Nq=2048, Nk=256, dqk=32, dv=96, mask=True),
    dict(B=2, H=1, Nq=512, Nk=5000, dqk=261, dv=261, same_kv=True, q_bcast=True),
    dict(B=2, H=1, Nq=512, Nk=5000, dqk=261, dv=261, same_kv=True, splits=3, qscale=3.0),
    dict(B=1, H=1, Nq=2048, Nk=9000, dqk=322, dv=322, same_kv=True, splits=4, mask=True),
])
def test_streaming_attention(case):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(2)
    B, H, Nq, Nk, dqk, dv = (case[k] for k in ("B", "H", "Nq", "Nk", "dqk", "dv"))
    same_kv, q_bcast, splits = case.get("same_kv", False), case.get("q_bcast", False), case.get("splits", 1)
    dev = "cuda"
    ldq, ldk, ldv = ops.pad8(H * dqk), ops.pad8(H * dqk), ops.pad8(H * dv)
    Qb = 1 if q_bcast else B
    Q = torch.zeros(Qb, Nq, ldq, dtype=torch.bfloat16, device=dev)
    Q[:, :, :H * dqk] = (torch.randn(Qb, Nq, H * dqk, device=dev) * case.get("qscale", 1.0)).to(torch.bfloat16)
    K = torch.zeros(B, Nk, ldk, dtype=torch.bfloat16, device=dev)
    K[:, :, :H * dqk] = torch.randn(B, Nk, H * dqk, device=dev).to(torch.bfloat16)
    if same_kv:
        V = K
    else:
        V = torch.zeros(B, Nk, ldv, dtype=torch.bfloat16, device=dev)
        V[:, :, :H * dv] = torch.randn(B, Nk, H * dv, device=dev).to(torch.bfloat16)
    km = rk = None
    if case.get("mask"):
        km = (torch.rand(B, Nk, device=dev) > 0.3).to(torch.uint8)
        km[0, Nk // 2:] = 0
        rk = (torch.rand(B, Nq, device=dev) > 0.1).to(torch.uint8)
    scale = dqk ** -0.5
    qf = Q[:, :, :H * dqk].float().reshape(Qb, Nq, H, dqk).expand(B, Nq, H, dqk)
    ref = _attn_ref(qf, K[:, :, :H * dqk].float().reshape(B, Nk, H, dqk),
                    V[:, :, :H * dv].float().reshape(B, Nk, H, dv), scale, km, rk)
    O = ops.attention_fwd(Q, K, V, B=B, H=H, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv,
                          strideQ=0 if q_bcast else Nq * ldq, strideK=Nk * ldk, strideV=Nk * ldv,
                          ldq=ldq, ldk=ldk, ldv=ldv, scale=scale, key_mask=km, row_keep=rk, num_splits=splits)
    torch.cuda.synchronize()
    assert _rel(O[:, :, :H * dv].float(), ref) < 1.5e-2


def test_lse_combine():
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    parts, b, h, nq, dv = 3, 2, 4, 37, 40
    Op = torch.randn(parts, b, h, nq, dv, device="cuda")
    mp = torch.randn(parts, b, h, nq, device="cuda") * 3
    lp = torch.rand(parts, b, h, nq, device="cuda") + 0.1
    mp[1, 0] = float("-inf")
    lp[1, 0] = 0
    O = ops.attention_combine(Op, mp, lp)
    M = mp.max(0).values
    w = torch.exp(mp - M)
    ref = ((Op * w[..., None]).sum(0) / (lp * w).sum(0)[..., None]).permute(0, 2, 1, 3).reshape(b, nq, h * dv)
    assert _rel(O[:, :, :h * dv].float(), ref) < 1e-2


@pytest.mark.parametrize("m,n,k", [(1000, 2, 322), (77, 5, 1026), (4097, 16, 64), (3, 1, 7)])
def test_linear_f32(m, n, k):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    x, w, b = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda"), torch.randn(n, device="cuda")
    y = ops.linear_f32(x, w, b)
    ref = (x.double() @ w.double().t() + b.double()).float()
    assert _rel(y, ref) < 1e-5   # fp32 accumulation order only


def test_gemm_fused_layernorm_epilogues():
    """Producer side: fp32 output + raw bf16 copy + per-row (sum, sum of squares).  Consumer side: raw bf16 rows times
    W diag(gamma) with the normalisation applied per output row in the epilogue == LayerNorm(x) @ W^T + b."""
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(3)
    M, C, N = 2048, 1024, 768
    dev = "cuda"
    # ---- producer: x = h @ W2^T + b + residual
    h = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(C, C, device=dev) / 32).to(torch.bfloat16)
    b2 = torch.randn(C, device=dev)
    res = torch.randn(M, C, device=dev) * 2 + 0.3
    x = torch.empty(M, C, device=dev)
    xb = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    st = ops.empty_row_stats(M, C, dev).fill_(float("nan"))   # every slot must be written
    ops.gemm(h, w2, M=M, N=C, K=C, bias=b2, residual=res, ldr=C, out_f32=x, ldo32=C, out_bf16=xb, ldo16=C,
             row_stats_out=st)
    st_again = torch.empty_like(st)
    ops.gemm(h, w2, M=M, N=C, K=C, bias=b2, residual=res, ldr=C, out_f32=x, ldo32=C, out_bf16=xb, ldo16=C,
             row_stats_out=st_again, reverse_tiles=True)
    assert torch.equal(st, st_again)      # plain stores per half-tile: bit-reproducible, whatever the tile order
    assert st.shape == (M, ops.stats_parts(M, C), 2) and st.shape[1] == 4 * (C // 256)
    parts = st
    st = st.sum(1)
    x_ref = h.float() @ w2.float().t() + b2 + res
    assert _rel(x, x_ref) < 2e-3
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert _rel(st[:, 0], x.sum(1)) < 1e-4 and _rel(st[:, 1], (x * x).sum(1)) < 1e-4
    # ---- consumer: LN(x) @ W^T + b through the raw bf16 rows
    gamma, beta = 1 + 0.1 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)
    w = torch.randn(N, C, device=dev) / 32
    b = torch.randn(N, device=dev)
    wp = (w * gamma[None, :]).to(torch.bfloat16)
    bias = b + w @ beta
    colsum = wp.double().sum(1).float()
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ops.gemm(xb, wp, M=M, N=N, K=C, bias=bias, out_bf16=y, ldo16=N, row_stats_in=parts, ln_colsum=colsum,
             ln_channels=C, ln_eps=1e-5)
    ref = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + b
    assert _rel(y.float(), ref) < 1e-2
    # the epilogues exist in the CTA-pair kernel only: a problem it cannot take is rejected, not silently unfused
    with pytest.raises(RuntimeError, match="fused-LayerNorm"):
        ops.gemm(xb, wp, M=M // 2, N=N, K=C, batch=2, strideA=(M // 2) * C, strideB=0, bias=bias, out_bf16=y, ldo16=N,
                 strideO16=(M // 2) * N, row_stats_in=parts, ln_colsum=colsum, ln_channels=C)


@pytest.mark.parametrize("M,C,N,kernel", [(256, 1280, 1536, 1), (784, 512, 512, 1), (2048, 512, 1536, 0), (300, 320, 200, 1)])
def test_gemm_fused_layernorm_epilogues_small_latent_arrays(M, C, N, kernel):
    """The same producer / consumer epilogues in the single-CTA kernel (the batch-1 towers: 256 .. 2048 latent rows),
    whatever tile width the dispatch picks."""
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(M + C)
    dev = "cuda"
    h = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(C, C, device=dev) / C ** 0.5).to(torch.bfloat16)
    b2 = torch.randn(C, device=dev)
    res = torch.randn(M, C, device=dev) * 2 + 0.3
    x = torch.empty(M, C, device=dev)
    xb = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    # kernel = 1 forces the single-CTA kernel (tile width auto: >= 64 columns); two extra slots check the zero fill
    parts = (ops.stats_parts(M, C) if kernel == 0 else 2 * ((C + 63) // 64)) + 2
    st = ops.empty_row_stats(M, C, dev, parts=parts).fill_(float("nan"))
    ops.gemm(h, w2, M=M, N=C, K=C, bias=b2, residual=res, ldr=C, out_f32=x, ldo32=C, out_bf16=xb, ldo16=C,
             row_stats_out=st, kernel=kernel)
    x_ref = h.float() @ w2.float().t() + b2 + res
    assert _rel(x, x_ref) < 2e-3
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.isfinite(st).all() and float(st[:, -2:].abs().max()) == 0.0
    tot = st.sum(1)
    assert _rel(tot[:, 0], x.sum(1)) < 1e-4 and _rel(tot[:, 1], (x * x).sum(1)) < 1e-4
    gamma, beta = 1 + 0.1 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)
    w = torch.randn(N, C, device=dev) / C ** 0.5
    b = torch.randn(N, device=dev)
    wp = (w * gamma[None, :]).to(torch.bfloat16)
    colsum = wp.double().sum(1).float()
    y = torch.empty(M, ops.pad8(N), dtype=torch.bfloat16, device=dev)
    ops.gemm(xb, wp, M=M, N=N, K=C, bias=b + w @ beta, act=1, out_bf16=y, ldo16=y.stride(0), row_stats_in=st,
             ln_colsum=colsum, ln_channels=C, ln_eps=1e-5, kernel=kernel)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + b)
    assert _rel(y[:, :N].float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K,batch", [(1000, 520, 200, 1), (2048, 1024, 512, 1), (300, 256, 64, 3)])
def test_gemm_pair_kernel_reverse_tile_order_gives_identical_results(M, N, K, batch):
    """pio_gemm_args.reverse_tiles only changes the order in which the CTA pairs walk the output tiles."""
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(7)
    ldk = ops.pad8(K)
    A = torch.randn(batch, M, ldk, device="cuda").to(torch.bfloat16)
    W = torch.randn(batch, N, ldk, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(batch, M, N, device="cuda")
    outs = []
    for rev in (False, True):
        o = torch.zeros(batch, M, N, device="cuda")
        ops.gemm(A, W, M=M, N=N, K=K, batch=batch, strideA=M * ldk, strideB=N * ldk, lda=ldk, ldb=ldk, bias=bias,
                 residual=res, ldr=N, strideR=M * N, out_f32=o, ldo32=N, strideO32=M * N, kernel=2, reverse_tiles=rev)
        outs.append(o)
    ref = A[:, :, :K].float() @ W[:, :, :K].float().transpose(1, 2) + bias + res
    assert _rel(outs[0], ref) < 2e-3
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("B,Nq,Nk,dqk,dv,masks,fp16", [
    (1, 256, 64, 64, 64, False, False),          # one pair, one key tile
    (1, 300, 200, 51, 37, False, False),         # ragged everything, odd head sizes
    (2, 1000, 2048, 323, 322, False, False),     # the optical-flow decoder's folded head sizes
    (2, 700, 777, 100, 200, True, False),        # key mask + wiped rows + residual + bias
    (1, 5000, 2048, 323, 322, False, True),      # fp16 operands
    (3, 129, 65, 384, 384, True, True),          # largest head sizes
])
def test_decoder_attention_pair_kernel(B, Nq, Nk, dqk, dv, masks, fp16):
    """pio_decoder_attention_fwd (CTA-pair query-tiled kernel, distinct K / V) vs a plain fp32 evaluation."""
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(B * 1000 + Nq)
    dt = torch.float16 if fp16 else torch.bfloat16
    pad8 = ops.pad8
    q = torch.randn(B, Nq, dqk, device="cuda")
    k = torch.randn(B, Nk, dqk, device="cuda")
    v = torch.randn(B, Nk, dv, device="cuda")
    scale = 1.0 / dqk ** 0.5 * 2.0          # logits of a few units: the running max moves
    Q = torch.zeros(B, Nq, pad8(dqk), dtype=dt, device="cuda")
    K = torch.full((B, Nk, pad8(dqk) + 8), 7.0, dtype=dt, device="cuda")      # junk beyond dqk must never be read as data
    V = torch.full((B, Nk, pad8(dv) + 16), -3.0, dtype=dt, device="cuda")
    Q[:, :, :dqk], K[:, :, :dqk], V[:, :, :dv] = q.to(dt), k.to(dt), v.to(dt)
    km = rk = res = bias = None
    if masks:
        km = (torch.rand(B, Nk, device="cuda") > 0.3)
        km[0, : Nk // 2] = False
        if B > 1:
            km[1, :] = False                  # a sample without any valid key: every row is wiped
        rk = (torch.rand(B, Nq, device="cuda") > 0.2)
        res = torch.randn(B, Nq, dv + 3, device="cuda")
        bias = torch.randn(dv, device="cuda")
    out = ops.decoder_attention(Q, K, V, B=B, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv, ldq=Q.stride(1), ldk=K.stride(1),
                                ldv=V.stride(1), strideQ=Q.stride(0), strideK=K.stride(0), strideV=V.stride(0), scale=scale,
                                key_mask=km.to(torch.uint8) if km is not None else None,
                                row_keep=rk.to(torch.uint8) if rk is not None else None, bias=bias,
                                residual=res, ldr=res.stride(1) if res is not None else 0,
                                strideR=res.stride(0) if res is not None else 0)
    # the same launch with the fused LayerNorm of the output rows (the MLP's operand)
    g, bt = 1 + 0.1 * torch.randn(dv, device="cuda"), 0.1 * torch.randn(dv, device="cuda")
    out2, out_ln = ops.decoder_attention(Q, K, V, B=B, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv, ldq=Q.stride(1), ldk=K.stride(1),
                                         ldv=V.stride(1), strideQ=Q.stride(0), strideK=K.stride(0), strideV=V.stride(0),
                                         scale=scale, key_mask=km.to(torch.uint8) if km is not None else None,
                                         row_keep=rk.to(torch.uint8) if rk is not None else None, bias=bias,
                                         residual=res, ldr=res.stride(1) if res is not None else 0,
                                         strideR=res.stride(0) if res is not None else 0, ln=(g, bt, 1e-5))
    assert torch.equal(out2, out)
    ln_ref = torch.nn.functional.layer_norm(out.reshape(B * Nq, dv), (dv,), g, bt, 1e-5)
    assert out_ln.shape == (B * Nq, pad8(dv)) and out_ln.dtype == dt
    assert float((out_ln[:, :dv].float() - ln_ref).abs().max()) <= (4e-3 if fp16 else 3e-2) * float(ln_ref.abs().max())
    if out_ln.shape[1] > dv:
        assert float(out_ln[:, dv:].float().abs().max()) == 0.0
    out = out.reshape(B, Nq, dv)
    qf, kf, vf = Q[:, :, :dqk].float(), K[:, :, :dqk].float(), V[:, :, :dv].float()
    s = torch.einsum("bqd,bkd->bqk", qf, kf) * scale
    if km is not None:
        s = s.masked_fill(~km[:, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)          # rows without a valid key
    ref = torch.einsum("bqk,bkd->bqd", p, vf)
    if rk is not None:
        ref = ref * rk[:, :, None]
    if bias is not None:
        ref = ref + bias
    if res is not None:
        ref = ref + res[:, :, :dv]
    err = float((out - ref).abs().max() / ref.abs().max())
    assert torch.isfinite(out).all()
    assert err < (2e-3 if fp16 else 1e-2), err


@pytest.mark.parametrize("fp16", [False, True])
def test_linear_f32_with_second_16bit_operand(fp16):
    """pio_linear_f32 with x2 / w2: y = x W^T + x2 W2^T + b — the decoder tail (final_layer folded into fc2)."""
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(11)
    M, K, K2, N = 5000, 322, 322, 2
    dt = torch.float16 if fp16 else torch.bfloat16
    x = torch.randn(M, K + 2, device="cuda")[:, :K]                 # a row pitch that is not K
    x2 = torch.randn(M, ops.pad8(K2), device="cuda").to(dt)
    w, w2, b = torch.randn(N, K, device="cuda"), torch.randn(N, K2, device="cuda"), torch.randn(N, device="cuda")
    y = ops.linear_f32(x, w, b, x2=x2, w2=w2)
    ref = (x.double() @ w.double().t() + x2[:, :K2].double() @ w2.double().t() + b.double()).float()
    assert _rel(y, ref) < 1e-5


@pytest.mark.parametrize("f16", [False, True])
def test_gemm_pair_kernel_split_residual_stream(f16):
    """pio_gemm_args.out_lo16 / residual_hi16 / residual_lo16: the fp32 residual stream of the large towers as a pair of
    16-bit arrays (value = hi + lo).  The pair output equals the fp32 output to 2^-16 (bf16) / 2^-21 (fp16) per element,
    hi is exactly the raw 16-bit copy, the statistics are those of the fp32 values, and a pair residual input gives
    the same result as its fp32 sum."""
    from perceiverio_pytorch_b200 import engine, ops
    with engine.precision_scope("fp16" if f16 else "bf16"):
        d16 = ops.dtype16()
        M, N, K = 2048 + 96, 1024, 512          # a partial last row tile
        assert ops.gemm_uses_pair_kernel(M, N)
        torch.manual_seed(3)
        A = torch.randn(M, K, device="cuda").to(d16)
        W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(d16)
        bias = torch.randn(N, device="cuda")
        x = torch.randn(M, N, device="cuda") * 3.0
        parts = ops.stats_parts(M, N)
        y32 = torch.empty(M, N, device="cuda")
        yraw = torch.empty(M, N, device="cuda", dtype=d16)
        st32 = ops.empty_row_stats(M, N, "cuda", parts)
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual=x, ldr=N, out_f32=y32, ldo32=N, out_bf16=yraw, ldo16=N,
                 row_stats_out=st32)
        # pair output from an fp32 residual
        hi = torch.empty(M, N, device="cuda", dtype=d16)
        lo = torch.empty(M, N, device="cuda", dtype=d16)
        st = ops.empty_row_stats(M, N, "cuda", parts)
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual=x, ldr=N, out_bf16=hi, ldo16=N, out_lo16=lo, row_stats_out=st)
        assert torch.equal(hi, yraw)
        assert torch.equal(st, st32)
        tol = 2.0 ** (-20 if f16 else -15)
        err = ((hi.float() + lo.float()) - y32).abs().max() / y32.abs().max()
        assert err <= tol, err
        # pair residual in, fp32 out: the same as the fp32 residual hi + lo
        xs = hi.float() + lo.float()
        z_ref = torch.empty(M, N, device="cuda")
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual=xs, ldr=N, out_f32=z_ref, ldo32=N)
        z = torch.empty(M, N, device="cuda")
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual_hi16=hi, residual_lo16=lo, ldr16=N, out_f32=z, ldo32=N)
        assert (z - z_ref).abs().max() <= 1e-6 * z_ref.abs().max()
        # pair in, pair out, back to front
        hi2 = torch.empty(M, N, device="cuda", dtype=d16)
        lo2 = torch.empty(M, N, device="cuda", dtype=d16)
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual_hi16=hi, residual_lo16=lo, ldr16=N, out_bf16=hi2, ldo16=N,
                 out_lo16=lo2, reverse_tiles=True)
        assert ((hi2.float() + lo2.float()) - z_ref).abs().max() <= tol * z_ref.abs().max()
        # ... with the row statistics (pio_gemm2_stream_kernel: one slot per 64-column slice), twice, bit-identically
        st2 = ops.empty_row_stats(M, N, "cuda", parts).fill_(float("nan"))
        hi3 = torch.empty_like(hi2)
        lo3 = torch.empty_like(lo2)
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual_hi16=hi, residual_lo16=lo, ldr16=N, out_bf16=hi3, ldo16=N,
                 out_lo16=lo3, row_stats_out=st2)
        assert torch.equal(hi3, hi2) and torch.equal(lo3, lo2)
        assert torch.equal(hi3, z_ref.to(d16))
        ssum = st2.sum(1)
        assert not torch.isnan(ssum).any()
        assert _rel(ssum[:, 0], z_ref.sum(1)) < 1e-4 and _rel(ssum[:, 1], (z_ref * z_ref).sum(1)) < 1e-4
        st3 = torch.empty_like(st2)
        ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual_hi16=hi, residual_lo16=lo, ldr16=N, out_bf16=hi3, ldo16=N,
                 out_lo16=lo3, row_stats_out=st3, reverse_tiles=True)
        assert torch.equal(st2, st3)
        # a column tail (N = 984: the last 64-column slice holds 24 columns, the last chunk 24 of 32) and no bias
        Nt = 984
        hi4 = torch.zeros(M, N, device="cuda", dtype=d16)
        lo4 = torch.zeros(M, N, device="cuda", dtype=d16)
        st4 = ops.empty_row_stats(M, N, "cuda", parts).fill_(float("nan"))
        ops.gemm(A, W, M=M, N=Nt, K=K, residual_hi16=hi, residual_lo16=lo, ldr16=N, out_bf16=hi4, ldo16=N, out_lo16=lo4,
                 row_stats_out=st4)
        zt = A.float() @ W[:Nt].float().t() + xs[:, :Nt]
        assert ((hi4.float() + lo4.float())[:, :Nt] - zt).abs().max() <= 4 * tol * zt.abs().max()
        assert (hi4[:, Nt:] == 0).all() and (lo4[:, Nt:] == 0).all()      # the TMA stores clip the tail
        assert _rel(st4.sum(1)[:, 0], zt.sum(1)) < 1e-4
        # the single-CTA kernel does not implement the pair stream: it must say so
        with pytest.raises(RuntimeError, match="CTA-pair kernel"):
            ops.gemm(A, W, M=M, N=N, K=K, bias=bias, residual=x, ldr=N, out_bf16=hi, ldo16=N, out_lo16=lo, kernel=1)
