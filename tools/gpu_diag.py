"""On-GPU kernel diagnostics (run under gpurun).  Each group runs in its own subprocess with a timeout so a
trapping / hanging kernel cannot take the other groups down.   python tools/gpu_diag.py [group ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GROUPS = ["env", "ln", "gemm_k", "gemm_mn", "gemm_epi", "gemm_cluster", "gemm_pair", "softmax", "flash", "combine"]


def _err(got, ref):
    d = (got.double() - ref.double()).abs()
    return float(d.max()), float(d.max() / ref.double().abs().max().clamp_min(1e-30))


def _errmap(got, ref, rb=16, cb=16, maxr=8, maxc=16):
    import torch
    d = (got.double() - ref.double()).abs()
    m, n = d.shape[-2:]
    d = d.reshape(-1, m, n)[0]
    rows = []
    for i in range(0, min(m, rb * maxr), rb):
        rows.append(" ".join(f"{float(d[i:i + rb, j:j + cb].max()):8.2e}" for j in range(0, min(n, cb * maxc), cb)))
    return "\n".join(rows)


def group_env():
    import torch
    print("torch", torch.__version__, "cuda", torch.version.cuda, "dev", torch.cuda.get_device_name(0))
    print("cc", torch.cuda.get_device_capability(0), "sms", torch.cuda.get_device_properties(0).multi_processor_count)
    print("ref present:", os.path.isdir("/root/reference"), "cpus", os.cpu_count())
    from perceiverio_pytorch_b200 import _lib
    lib = _lib.load()
    print("abi", lib.pio_abi_version(), "check_device", lib.pio_check_device())


def group_ln():
    import torch
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    for rows, c in [(1000, 261), (513, 1024), (77, 322), (300, 1280), (64, 32), (50, 1026)]:
        x = torch.randn(rows, c, device="cuda") * 2 + 0.5
        g = torch.randn(c, device="cuda")
        b = torch.randn(c, device="cuda")
        y = ops.layernorm_bf16(x, g, b)
        ref = torch.nn.functional.layer_norm(x, (c,), g, b, 1e-5)
        torch.cuda.synchronize()
        e = _err(y[:, :c].float(), ref)
        padz = float(y[:, c:].float().abs().max()) if y.shape[1] > c else 0.0
        print(f"ln rows={rows} C={c}: abs {e[0]:.3e} rel {e[1]:.3e} pad {padz} {'OK' if e[1] < 8e-3 and padz == 0 else 'FAIL'}")
    x = torch.randn(40, 100, device="cuda")
    y = ops.layernorm_bf16(x, None, None, normalize=False)
    print("cast-only:", _err(y[:, :100].float(), x))


def _gemm_case(M, N, K, batch=1, b_mn=False, tile_n=0, bias_mode=0, act=0, alpha=1.0, residual=False, f32=True, bf16=False, verbose_map=False, cluster_m=None, kernel=None, res_bcast=False):
    import torch
    from perceiverio_pytorch_b200 import ops
    ld_k = ops.pad8(K)
    A = torch.zeros(batch, M, ld_k, dtype=torch.bfloat16, device="cuda")
    A[:, :, :K] = torch.randn(batch, M, K, device="cuda").to(torch.bfloat16)
    A[:, :, K:] = 7.0  # garbage in the pad: must never be read (TMA OOB zero-fill by logical K)
    if not b_mn:
        Bm = torch.zeros(batch, N, ld_k, dtype=torch.bfloat16, device="cuda")
        Bm[:, :, :K] = torch.randn(batch, N, K, device="cuda").to(torch.bfloat16)
        Bm[:, :, K:] = 5.0
        ref = A[:, :, :K].float() @ Bm[:, :, :K].float().transpose(1, 2)
        strideB = N * ld_k
        ldb = ld_k
    else:
        ld_n = ops.pad8(N)
        Bm = torch.zeros(batch, K, ld_n, dtype=torch.bfloat16, device="cuda")
        Bm[:, :, :N] = torch.randn(batch, K, N, device="cuda").to(torch.bfloat16)
        Bm[:, :, N:] = 5.0
        ref = A[:, :, :K].float() @ Bm[:, :, :N].float()
        strideB = K * ld_n
        ldb = ld_n
    ref = ref * alpha
    bias = None
    if bias_mode == 1:
        bias = torch.randn(N, device="cuda")
        ref = ref + bias
    elif bias_mode == 2:
        bias = torch.randn(M, device="cuda")
        ref = ref + bias[None, :, None]
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    res = None
    if residual:
        res = torch.randn(1 if res_bcast else batch, M, N, device="cuda")
        ref = ref + res
    o32 = torch.full((batch, M, N), float("nan"), device="cuda") if f32 else None
    ld16 = ops.pad8(N)
    o16 = torch.zeros(batch, M, ld16, dtype=torch.bfloat16, device="cuda") if bf16 else None
    ops.gemm(A, Bm, M=M, N=N, K=K, batch=batch, b_mn_major=b_mn, strideA=M * ld_k, strideB=strideB, lda=ld_k, ldb=ldb,
             bias=bias, bias_mode=bias_mode, act=act, alpha=alpha, residual=res, ldr=N, strideR=0 if res_bcast else M * N,
             out_f32=o32, ldo32=N, strideO32=M * N, out_bf16=o16, ldo16=ld16, strideO16=M * ld16, tile_n=tile_n,
             cluster_m=cluster_m, kernel=kernel)
    torch.cuda.synchronize()
    msgs = []
    ok = True
    if f32:
        e = _err(o32, ref)
        good = e[1] < 2e-3 and not bool(torch.isnan(o32).any())
        ok &= good
        msgs.append(f"f32 abs {e[0]:.3e} rel {e[1]:.3e}")
        if not good and verbose_map:
            print(_errmap(torch.nan_to_num(o32, nan=1e9), ref))
    if bf16:
        e = _err(o16[:, :, :N].float(), ref)
        good = e[1] < 1e-2
        ok &= good
        msgs.append(f"bf16 abs {e[0]:.3e} rel {e[1]:.3e}")
    print(f"gemm M={M} N={N} K={K} b={batch} mn={int(b_mn)} tn={tile_n} cl={cluster_m} kern={kernel} bias={bias_mode} act={act} res={int(residual)}: "
          + "; ".join(msgs) + (" OK" if ok else " FAIL"), flush=True)
    return ok


def group_gemm_k():
    _gemm_case(128, 64, 64, tile_n=64, verbose_map=True)
    _gemm_case(128, 128, 64, tile_n=128, verbose_map=True)
    _gemm_case(128, 256, 64, tile_n=256, verbose_map=True)
    _gemm_case(128, 256, 256, verbose_map=True)
    _gemm_case(256, 512, 1024)
    _gemm_case(1000, 1000, 1024)
    _gemm_case(333, 261, 261, verbose_map=True)
    _gemm_case(77, 2, 322)
    _gemm_case(4096, 3072, 1024)
    _gemm_case(512, 512, 128, batch=16)
    _gemm_case(200, 300, 72, batch=3)
    _gemm_case(32768, 1024, 1024)


def group_gemm_mn():
    _gemm_case(128, 64, 64, b_mn=True, tile_n=64, verbose_map=True)
    _gemm_case(128, 128, 64, b_mn=True, tile_n=128, verbose_map=True)
    _gemm_case(128, 256, 128, b_mn=True, tile_n=256, verbose_map=True)
    _gemm_case(1000, 1024, 512, b_mn=True)
    _gemm_case(300, 322, 2048, b_mn=True, batch=2)
    _gemm_case(130, 704, 1000, b_mn=True, verbose_map=True)


def group_gemm_epi():
    _gemm_case(512, 1024, 1024, bias_mode=1, act=1, f32=False, bf16=True)
    _gemm_case(512, 1024, 1024, bias_mode=1, residual=True, f32=True, bf16=True)
    _gemm_case(300, 261, 1024, bias_mode=1, residual=True, f32=True, bf16=True)
    _gemm_case(300, 322, 512, bias_mode=2, alpha=0.25, f32=True, bf16=True)
    _gemm_case(1000, 1000, 1024, bias_mode=1, f32=True)
    _gemm_case(100, 2, 322, bias_mode=1, f32=True)


def group_gemm_cluster():
    for cl in (1, 2, 4):
        _gemm_case(256, 256, 64, cluster_m=cl, verbose_map=True)
        _gemm_case(1000, 1000, 1024, cluster_m=cl, bias_mode=1, residual=True, bf16=True)
        _gemm_case(333, 261, 261, cluster_m=cl)
        _gemm_case(640, 64, 200, cluster_m=cl, tile_n=64)
        _gemm_case(640, 128, 200, cluster_m=cl, tile_n=128, batch=3)
        _gemm_case(700, 704, 1000, b_mn=True, cluster_m=cl)
        _gemm_case(300, 64, 300, b_mn=True, cluster_m=cl, tile_n=64, batch=2)
        _gemm_case(4096, 3072, 1024, cluster_m=cl)


def group_gemm_pair():
    """CTA-pair (cta_group::2) kernel, forced with kernel=2."""
    _gemm_case(256, 256, 64, kernel=2, verbose_map=True)
    _gemm_case(256, 256, 64, kernel=2, f32=False, bf16=True, verbose_map=True)
    _gemm_case(256, 256, 256, kernel=2, verbose_map=True)
    _gemm_case(512, 512, 1024, kernel=2, bias_mode=1, residual=True)
    _gemm_case(512, 1024, 1024, kernel=2, bias_mode=1, act=1, f32=False, bf16=True)
    _gemm_case(1000, 1000, 1024, kernel=2, bias_mode=1, residual=True)
    _gemm_case(1000, 1000, 1024, kernel=2, bias_mode=1, f32=False, bf16=True)
    _gemm_case(333, 264, 261, kernel=2, verbose_map=True)
    _gemm_case(300, 328, 200, kernel=2, bias_mode=2, alpha=0.25, f32=False, bf16=True)
    _gemm_case(512, 1024, 256, batch=5, kernel=2, bias_mode=1, residual=True, res_bcast=True)
    _gemm_case(200, 304, 72, batch=3, kernel=2, bias_mode=1, residual=True)
    _gemm_case(4096, 3072, 1024, kernel=2, f32=False, bf16=True)
    _gemm_case(32768, 1024, 1024, kernel=2, bias_mode=1, residual=True)
    _gemm_case(20000, 1024, 1024, kernel=2, bias_mode=1, act=1, f32=False, bf16=True)


def _time(fn, iters=20):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def group_perf():
    """Micro-benchmarks of the hot kernels at the ImageNet-recipe shapes (B=64)."""
    import torch
    from perceiverio_pytorch_b200 import ops
    dev = "cuda"
    M = 32768
    x = torch.randn(M, 1024, device=dev)
    g = torch.randn(1024, device=dev)
    t = _time(lambda: ops.layernorm_bf16(x, g, g))
    print(f"layernorm 32768x1024: {t * 1e3:.1f} us  {M * 1024 * 6 / t / 1e6:.0f} GB/s")
    xi = torch.randn(8 * 50176, 261, device=dev)
    gi = torch.randn(261, device=dev)
    t = _time(lambda: ops.layernorm_bf16(xi, gi, gi), 5)
    print(f"layernorm {xi.shape[0]}x261: {t * 1e3:.1f} us  {xi.shape[0] * (261 * 4 + 264 * 2) / t / 1e6:.0f} GB/s")
    A = torch.randn(M, 1024, device=dev).to(torch.bfloat16)
    res = torch.randn(M, 1024, device=dev)
    bias = torch.randn(3072, device=dev)
    for N in (1024, 3072):
        W = torch.randn(N, 1024, device=dev).to(torch.bfloat16)
        o16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o32 = torch.empty(M, N, device=dev)
        for cl in (1, 2, 4):
            for tn in (256, 128):
                t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, out_bf16=o16, ldo16=N, cluster_m=cl, tile_n=tn))
                line = f"gemm {M}x{N}x1024 cl={cl} tn={tn}: bf16-out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
                if N == 1024:
                    t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, residual=res, ldr=N, out_f32=o32, ldo32=N,
                                               cluster_m=cl, tile_n=tn))
                    line += f"; fp32 residual in/out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
                    t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, act=1, out_bf16=o16, ldo16=N,
                                               cluster_m=cl, tile_n=tn))
                    line += f"; gelu bf16-out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
                print(line, flush=True)
    for N in (1024, 3072):
        W = torch.randn(N, 1024, device=dev).to(torch.bfloat16)
        o16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        o32 = torch.empty(M, N, device=dev)
        t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, out_bf16=o16, ldo16=N, kernel=2))
        line = f"gemm {M}x{N}x1024 pair: bf16-out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
        if N == 1024:
            t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, residual=res, ldr=N, out_f32=o32, ldo32=N, kernel=2))
            line += f"; fp32 residual in/out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
            t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, out_f32=o32, ldo32=N, kernel=2))
            line += f"; fp32 out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
            t = _time(lambda: ops.gemm(A, W, M=M, N=N, K=1024, bias=bias, act=1, out_bf16=o16, ldo16=N, kernel=2))
            line += f"; gelu bf16-out {2 * M * N * 1024 / t / 1e9:.0f} TF/s"
        print(line, flush=True)
    # fused-LayerNorm epilogues
    W1 = torch.randn(1024, 1024, device=dev).to(torch.bfloat16)
    o16 = torch.empty(M, 1024, dtype=torch.bfloat16, device=dev)
    o32 = torch.empty(M, 1024, device=dev)
    st = ops.empty_row_stats(M, 1024, dev).zero_()
    cs = torch.randn(1024, device=dev)
    fl = 2 * M * 1024 * 1024
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, residual=res, ldr=1024, out_f32=o32, ldo32=1024, kernel=2))
    line = f"producer 32768x1024x1024: plain {fl / t / 1e9:.0f}"
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, residual=res, ldr=1024, out_f32=o32, ldo32=1024, out_bf16=o16, ldo16=1024, kernel=2))
    line += f"; +raw bf16 {fl / t / 1e9:.0f}"
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, residual=res, ldr=1024, out_f32=o32, ldo32=1024, row_stats_out=st, kernel=2))
    line += f"; +stats {fl / t / 1e9:.0f}"
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, residual=res, ldr=1024, out_f32=o32, ldo32=1024, out_bf16=o16, ldo16=1024, row_stats_out=st, kernel=2))
    line += f"; +both {fl / t / 1e9:.0f} TF/s"
    print(line, flush=True)
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, act=1, out_bf16=o16, ldo16=1024, kernel=2))
    line = f"consumer 32768x1024x1024 gelu: plain {fl / t / 1e9:.0f}"
    t = _time(lambda: ops.gemm(A, W1, M=M, N=1024, K=1024, bias=bias, act=1, out_bf16=o16, ldo16=1024, kernel=2, row_stats_in=st, ln_colsum=cs, ln_channels=1024))
    line += f"; +row affine {fl / t / 1e9:.0f} TF/s"
    print(line, flush=True)
    Ab = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    Wb = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    ob = torch.empty(8192, 8192, dtype=torch.bfloat16, device=dev)
    for kern in (1, 2):
        t = _time(lambda: ops.gemm(Ab, Wb, M=8192, N=8192, K=8192, out_bf16=ob, ldo16=8192, kernel=kern), 5)
        print(f"gemm 8192^3 kernel={kern}: {2 * 8192 ** 3 / t / 1e9:.0f} TF/s", flush=True)
    t = _time(lambda: torch.matmul(Ab, Wb.t()), 5)
    print(f"torch.matmul (cuBLAS) 8192^3: {2 * 8192 ** 3 / t / 1e9:.0f} TF/s")
    W = torch.randn(3072, 1024, device=dev).to(torch.bfloat16)
    t = _time(lambda: torch.matmul(A, W.t()))
    print(f"torch.matmul (cuBLAS) {M}x3072x1024: {2 * M * 3072 * 1024 / t / 1e9:.0f} TF/s")
    # tower attention 64 x 8 heads x 512 x 512 x 128
    qkv = torch.randn(64 * 512, 3072, device=dev).to(torch.bfloat16)
    f = lambda: ops.attention_fwd(qkv, qkv[:, 1024:], qkv[:, 2048:], B=64, H=8, Nq=512, Nk=512, dqk=128, dv=128,
                                  strideQ=512 * 3072, strideK=512 * 3072, strideV=512 * 3072, ldq=3072, ldk=3072, ldv=3072)
    qv, kv_, vv = qkv.view(-1), qkv.view(-1)[1024:], qkv.view(-1)[2048:]
    f = lambda: ops.attention_fwd(qv, kv_, vv, B=64, H=8, Nq=512, Nk=512, dqk=128, dv=128,
                                  strideQ=512 * 3072, strideK=512 * 3072, strideV=512 * 3072, ldq=3072, ldk=3072, ldv=3072)
    t = _time(f)
    print(f"tower attention 64x8x512x512x128: {t * 1e3:.1f} us  {4 * 64 * 8 * 512 * 512 * 128 / t / 1e9:.0f} TF/s")
    # encoder attention (folded): 64 x 512 x 50176 x 261
    B = 16
    kvn = torch.randn(B * 50176, 264, device=dev).to(torch.bfloat16)
    qf = torch.randn(512, 264, device=dev).to(torch.bfloat16)
    f = lambda: ops.attention_fwd(qf.view(-1), kvn.view(-1), kvn.view(-1), B=B, H=1, Nq=512, Nk=50176, dqk=261, dv=261,
                                  strideQ=0, strideK=50176 * 264, strideV=50176 * 264, ldq=264, ldk=264, ldv=264)
    t = _time(f, 5)
    print(f"encoder attention {B}x512x50176x261: {t * 1e3:.1f} us  {4 * B * 512 * 50176 * 261 / t / 1e9:.0f} TF/s "
          f"(K/V read once = {B * 50176 * 264 * 2 / t / 1e6:.0f} GB/s)")


def group_softmax():
    import torch
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(0)
    for b, r, c in [(2, 40, 512), (1, 7, 52097), (3, 130, 785)]:
        S = torch.randn(b, r, c, device="cuda") * 3
        km = (torch.rand(b, c, device="cuda") > 0.3).to(torch.uint8)
        rk = (torch.rand(b, r, device="cuda") > 0.2).to(torch.uint8)
        P = ops.softmax_bf16(S, c, 0.37, km, rk)
        ref = torch.softmax(torch.where(km[:, None, :].bool(), S * 0.37, torch.tensor(float("-inf"), device="cuda")), -1)
        ref = ref * rk[:, :, None]
        torch.cuda.synchronize()
        e = _err(P[:, :, :c].float(), ref)
        print(f"softmax {b}x{r}x{c}: abs {e[0]:.3e} rel {e[1]:.3e} {'OK' if e[1] < 1e-2 else 'FAIL'}")


def _attn_ref(q, k, v, scale, km=None, rk=None):
    import torch
    # q [B,Nq,H,dqk] k [B,Nk,H,dqk] v [B,Nk,H,dv] fp32
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if km is not None:
        s = torch.where(km[:, None, None, :].bool(), s, torch.tensor(float("-inf"), device=s.device))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v)
    if rk is not None:
        o = o * rk[:, :, None, None]
    return o.reshape(o.shape[0], o.shape[1], -1)


def _flash_case(B, H, Nq, Nk, dqk, dv, mask=False, splits=1, same_kv=False, qscale=1.0, q_bcast=False):
    import torch
    from perceiverio_pytorch_b200 import ops
    dev = "cuda"
    ldq, ldk, ldv = ops.pad8(H * dqk), ops.pad8(H * dqk), ops.pad8(H * dv)
    Qb = 1 if q_bcast else B
    Q = torch.zeros(Qb, Nq, ldq, dtype=torch.bfloat16, device=dev)
    Q[:, :, :H * dqk] = (torch.randn(Qb, Nq, H * dqk, device=dev) * qscale).to(torch.bfloat16)
    K = torch.zeros(B, Nk, ldk, dtype=torch.bfloat16, device=dev)
    K[:, :, :H * dqk] = torch.randn(B, Nk, H * dqk, device=dev).to(torch.bfloat16)
    if same_kv:
        assert dqk == dv
        V = K
    else:
        V = torch.zeros(B, Nk, ldv, dtype=torch.bfloat16, device=dev)
        V[:, :, :H * dv] = torch.randn(B, Nk, H * dv, device=dev).to(torch.bfloat16)
    km = rk = None
    if mask:
        km = (torch.rand(B, Nk, device=dev) > 0.3).to(torch.uint8)
        km[0, Nk // 2:] = 0
        rk = (torch.rand(B, Nq, device=dev) > 0.1).to(torch.uint8)
    scale = dqk ** -0.5
    qf = Q[:, :, :H * dqk].float().reshape(Qb, Nq, H, dqk).expand(B, Nq, H, dqk)
    ref = _attn_ref(qf, K[:, :, :H * dqk].float().reshape(B, Nk, H, dqk),
                    V[:, :, :H * dv].float().reshape(B, Nk, H, dv), scale, km, rk)
    t0 = time.time()
    O = ops.attention_fwd(Q, K, V, B=B, H=H, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv,
                          strideQ=0 if q_bcast else Nq * ldq, strideK=Nk * ldk, strideV=Nk * ldv,
                          ldq=ldq, ldk=ldk, ldv=ldv, scale=scale, key_mask=km, row_keep=rk, num_splits=splits)
    torch.cuda.synchronize()
    e = _err(O[:, :, :H * dv].float(), ref)
    ok = e[1] < 1.5e-2
    print(f"flash B={B} H={H} Nq={Nq} Nk={Nk} d={dqk}/{dv} mask={int(mask)} splits={splits} samekv={int(same_kv)} "
          f"qs={qscale}: abs {e[0]:.3e} rel {e[1]:.3e} {'OK' if ok else 'FAIL'} ({time.time() - t0:.2f}s)", flush=True)
    if not ok:
        print(_errmap(O[0, :, :H * dv].float(), ref[0]))
    return ok


def group_flash():
    _flash_case(1, 1, 128, 128, 64, 64)
    _flash_case(1, 1, 128, 256, 64, 64)
    _flash_case(1, 1, 128, 1024, 64, 64, qscale=4.0)
    _flash_case(1, 1, 100, 300, 64, 64)
    _flash_case(2, 8, 512, 512, 128, 128)
    _flash_case(2, 8, 256, 256, 32, 160)
    _flash_case(1, 16, 2048, 2048, 32, 32)
    _flash_case(1, 8, 784, 784, 64, 64)
    _flash_case(2, 8, 256, 2048, 32, 160, mask=True)
    _flash_case(2, 8, 2048, 256, 32, 96, mask=True)
    _flash_case(2, 1, 512, 5000, 261, 261, same_kv=True, q_bcast=True)
    _flash_case(2, 1, 512, 5000, 261, 261, same_kv=True, splits=3, qscale=3.0)
    _flash_case(1, 1, 2048, 9000, 322, 322, same_kv=True, splits=4, mask=True)


def group_combine():
    import torch
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
    ref = (Op * w[..., None]).sum(0) / (lp * w).sum(0)[..., None]
    ref = ref.permute(0, 2, 1, 3).reshape(b, nq, h * dv)
    torch.cuda.synchronize()
    e = _err(O[:, :, :h * dv].float(), ref)
    print(f"combine: abs {e[0]:.3e} rel {e[1]:.3e} {'OK' if e[1] < 1e-2 else 'FAIL'}")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        globals()["group_" + sys.argv[2]]()
        sys.exit(0)
    groups = sys.argv[1:] or GROUPS
    for g in groups:
        print(f"===== {g} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", g], timeout=300,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:])
            print(f"[{g}] exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"").decode("utf-8", "replace")[-6000:] if isinstance(e.stdout, bytes) else (e.stdout or "")[-6000:])
            print(f"[{g}] TIMEOUT", flush=True)
