"""Per-subsystem timing of the hot path at the FULL size of every BASELINE.json config (SURVEY.md section 8 shape
table): encoder cross-attend, latent tower, decoder (+ final layer), each timed alone with CUDA events, eagerly and
replayed from a CUDA graph; TFLOP/s use the reference-algorithm FLOPs of SURVEY.md section 8(d).

    python tools/bench_configs.py [--configs language classification flow multimodal] [--iters 10]

One JSON line per config.  Random-init weights (biases / LayerNorm affines perturbed), synthetic inputs."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import perceiverio_pytorch_b200 as pio  # noqa: E402
from perceiverio_pytorch_b200 import _lib  # noqa: E402
from perceiverio_pytorch_b200.graph import GraphedForward  # noqa: E402

CONFIGS = {
    # (language_perceiver.py:24-46)
    "language": dict(B=1, Nk=2048, Nq=2048,
                     enc=dict(num_input_channels=768, num_self_attends_per_block=26, num_blocks=1, num_latents=256,
                              num_latent_channels=1280, qk_channels=256, v_channels=1280, num_cross_attend_heads=8,
                              num_self_attend_heads=8),
                     dec=dict(query_channels=768, final_project_out_channels=768, num_latent_channels=1280,
                              qk_channels=256, v_channels=768, num_heads=8, use_query_residual=False,
                              final_project=False), masks=True),
    # (classification_perceiver.py:76-125, FOURIER_POS_PIXEL)
    "classification": dict(B=64, Nk=50176, Nq=1000,
                           enc=dict(num_input_channels=261, num_self_attends_per_block=6, num_blocks=8,
                                    num_latents=512, num_latent_channels=1024),
                           dec=dict(query_channels=1024, final_project_out_channels=1000, num_latent_channels=1024,
                                    use_query_residual=True), masks=False),
    # (flow_perceiver.py:47-97)
    "flow": dict(B=1, Nk=182528, Nq=182528,
                 enc=dict(num_input_channels=322, num_self_attends_per_block=24, num_blocks=1, num_latents=2048,
                          num_latent_channels=512, num_self_attend_heads=16),
                 dec=dict(query_channels=322, final_project_out_channels=2, num_latent_channels=512,
                          use_query_residual=False), masks=False),
    # (multimodal_perceiver.py:52-135), one chunk call
    "multimodal": dict(B=1, Nk=52097, Nq=6288,
                       enc=dict(num_input_channels=704, num_self_attends_per_block=8, num_blocks=1, num_latents=784,
                                num_latent_channels=512),
                       dec=dict(query_channels=1026, final_project_out_channels=512, num_latent_channels=512,
                                use_query_residual=False), masks=False),
}


def _blk(nq, nk, cq, ck, qk, v, o):
    return 2 * nq * cq * qk + 2 * nk * ck * (qk + v) + 2 * nq * nk * (qk + v) + 2 * nq * v * o + 4 * nq * o * o


def flops(cfg):
    e, d = cfg["enc"], cfg["dec"]
    nlat, c, cin = e["num_latents"], e["num_latent_channels"], e["num_input_channels"]
    qk_x = e.get("qk_channels") or cin
    v_x = e.get("v_channels") or qk_x
    enc = _blk(nlat, cfg["Nk"], c, cin, qk_x, v_x, c)
    qk_s = e.get("qk_channels") or c
    v_s = e.get("v_channels") or qk_s
    layers = e["num_self_attends_per_block"] * e["num_blocks"]
    tower = layers * _blk(nlat, nlat, c, c, qk_s, v_s, c)
    cq = d["query_channels"]
    qk_d = d.get("qk_channels") or c
    v_d = d.get("v_channels") or qk_d
    dec = _blk(cfg["Nq"], nlat, cq, c, qk_d, v_d, cq)
    if d.get("final_project", True):
        dec += 2 * cfg["Nq"] * cq * d["final_project_out_channels"]
    B = cfg["B"]
    return B * enc, B * tower, B * dec


def perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("bias"):
                prm.copy_((0.1 if "layer_norm" in name else 0.02) * torch.randn(prm.shape, generator=g))
            elif "layer_norm" in name:
                prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))


_L2_FLUSH = None


def timeit(fn, iters, flush=True):
    """Mean device time of fn() over `iters` calls (CUDA events).  With `flush`, a 256 MiB buffer (> the 126 MB L2) is
    overwritten before every call, so no call finds its operands or weights cached by the previous one; the flush's
    own time (measured the same way) is subtracted."""
    global _L2_FLUSH
    if flush and _L2_FLUSH is None:
        with torch.inference_mode(False):     # a normal tensor: it is overwritten in place inside and outside inference mode
            _L2_FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        if flush:
            _L2_FLUSH.zero_()
        fn()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / iters
    if flush:
        a.record()
        for _ in range(iters):
            _L2_FLUSH.zero_()
        b.record()
        torch.cuda.synchronize()
        t -= a.elapsed_time(b) / iters
    return t


def sustained_peak():
    peak = 1397.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("bf16_tflops_sustained", peak))
    return peak


def measure_config(name, iters=10, eager=True, precision=None):
    """One config at full size -> dict (see module docstring).  `eager=False` skips the eagerly launched timings and the
    per-family profile (bench.py's `other_configs` only wants the replayed-graph numbers)."""
    peak = sustained_peak()
    cfg = CONFIGS[name]
    from perceiverio_pytorch_b200 import engine
    # the operand format each model runs with after swap_hot_path(model) (install._auto_precision): fp16 for the
    # optical-flow regression head, bf16 otherwise; `precision` overrides
    precision = precision or ("fp16" if name == "flow" else "bf16")
    with engine.precision_scope(precision):
        res = _measure_config(name, cfg, peak, iters, eager)
    res["precision"] = precision
    return res


def _measure_config(name, cfg, peak, iters, eager):
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**cfg["enc"]).eval()
    dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
    perturb(enc, 1)
    perturb(dec, 2)
    enc, dec = enc.cuda(), dec.cuda()
    B, Nk, Nq = cfg["B"], cfg["Nk"], cfg["Nq"]
    x = torch.randn(B, Nk, cfg["enc"]["num_input_channels"], device="cuda")
    query = x if name == "flow" else torch.randn(B, Nq, cfg["dec"]["query_channels"], device="cuda")
    imask = qmask = None
    if cfg["masks"]:
        imask = torch.zeros(B, Nk, dtype=torch.bool, device="cuda")
        imask[:, :1500] = True
        qmask = imask[:, :Nq].clone()
    res = {"config": name, "B": B, "inputs": Nk, "latents": cfg["enc"]["num_latents"], "queries": Nq}
    with torch.inference_mode():
        lat = enc.latents(x)
        rk = imask.any(dim=1, keepdim=True).expand(B, lat.shape[1]) if imask is not None else None

        def f_enc():
            return enc.cross_attend._forward_factored(lat, x, key_mask=imask, row_keep=rk)[0]

        z0 = f_enc()

        def f_tower(z=z0):
            for _ in range(enc._num_blocks):
                for sa in enc.self_attends:
                    z = sa(z)
            return z

        z1 = f_tower()

        def f_dec():
            return dec(query, z1, query_mask=qmask)

        def f_all():
            return dec(query, enc(x, lat, input_mask=imask), query_mask=qmask)

        if eager:
            t_enc, t_tower, t_dec = timeit(f_enc, iters), timeit(f_tower, iters), timeit(f_dec, iters)
            t_all = timeit(f_all, iters)
            # per-kernel-family device time of one eager forward (library-side CUDA events around every launch)
            _lib.profile_read()
            _lib.profile_enable(True)
            f_all()
            torch.cuda.synchronize()
            _lib.profile_enable(False)
            res["kernel_families_one_forward"] = {k: {"ms": round(v["ms"], 4), "launches": v["launches"]}
                                                  for k, v in _lib.profile_read().items() if v["launches"] > 0}
            res["ms"] = {"encoder_xattn": round(t_enc, 4), "tower": round(t_tower, 4), "decoder": round(t_dec, 4),
                         "forward_eager": round(t_all, 4)}
    n0 = _lib.launch_count()
    g = GraphedForward(lambda xx: dec(xx if name == "flow" else query, enc(xx, enc.latents(xx), input_mask=imask),
                                      query_mask=qmask), [x], warmup=1)
    res["launches_per_forward"] = (_lib.launch_count() - n0) // 2      # one warm-up + the capture
    t_graph = timeit(lambda: g(g.inputs[0]), iters)
    # the same three subsystems replayed from CUDA graphs: at batch 1 the eager numbers above are bound by the
    # host's launch rate (7 launches per tower layer from Python), these are the device's
    g_enc = GraphedForward(lambda xx: enc.cross_attend._forward_factored(enc.latents(xx), xx, key_mask=imask,
                                                                         row_keep=rk)[0], [x], warmup=1)
    t_enc_g = timeit(lambda: g_enc(g_enc.inputs[0]), iters)
    g_tower = GraphedForward(lambda zz: f_tower(zz), [z0], warmup=1)
    t_tower_g = timeit(lambda: g_tower(g_tower.inputs[0]), iters)
    g_dec = GraphedForward(lambda zz: dec(query, zz, query_mask=qmask), [z1], warmup=1)
    t_dec_g = timeit(lambda: g_dec(g_dec.inputs[0]), iters)
    del g_enc, g_tower, g_dec
    fe, ft, fd = flops(cfg)
    tf = lambda fl, ms: round(fl / (ms * 1e-3) / 1e12, 1)   # noqa: E731
    res.update({
        "gflop_reference_algorithm": {"encoder": round(fe / 1e9, 1), "tower": round(ft / 1e9, 1),
                                      "decoder": round(fd / 1e9, 1)},
        "ms_graph": {"encoder_xattn": round(t_enc_g, 4), "tower": round(t_tower_g, 4), "decoder": round(t_dec_g, 4),
                     "forward": round(t_graph, 4)},
        "tflops": {"encoder_xattn": tf(fe, t_enc_g), "tower": tf(ft, t_tower_g), "decoder": tf(fd, t_dec_g),
                   "forward_graph": tf(fe + ft + fd, t_graph)},
        "frac_of_sustained_bf16_peak": {"encoder_xattn": round(tf(fe, t_enc_g) / peak, 3),
                                        "tower": round(tf(ft, t_tower_g) / peak, 3),
                                        "decoder": round(tf(fd, t_dec_g) / peak, 3),
                                        "forward_graph": round(tf(fe + ft + fd, t_graph) / peak, 3)},
        "samples_per_s_graph": round(B / (t_graph * 1e-3), 2),
        "l2": "256 MiB buffer overwritten before every timed call (its own time subtracted)"})
    del enc, dec, x, query, g
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=list(CONFIGS))
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    for name in args.configs:
        print(json.dumps(measure_config(name, args.iters)), flush=True)


if __name__ == "__main__":
    main()
