"""Generate tests/golden/full/*.npz: the LIVE reference's PerceiverEncoder + PerceiverDecoder forward on the four
BASELINE.json configurations at FULL size and depth (run in the build container only; ~1 minute of CPU time).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/make_golden_full.py [name ...]

Parameters and inputs are not stored: both are regenerated from seeds (oracle/full_configs.py) on whichever side
runs the comparison.  A fixture keeps a strided subsample of the reference's latents and output, their max-norms (the
denominator of the error metric) and the seeds.  The same script checks the CPU oracle against the complete reference
result (<= 1e-5), so the oracle is pinned at full size as well.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import full_configs as F  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "full")
PARAM_SEEDS = (101, 202)    # encoder, decoder
INPUT_SEED = 11


def main():
    ref = ref_shim.load_reference()
    assert ref is not None, "reference tree not mounted"
    R = ref.perceiver
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or sorted(F.FULL_CONFIGS)
    torch.set_num_threads(os.cpu_count() or 1)
    for name in names:
        cfg = F.FULL_CONFIGS[name]
        enc = F.seeded_fill(R.PerceiverEncoder(**cfg["enc"]).eval(), PARAM_SEEDS[0])
        dec = F.seeded_fill(R.PerceiverDecoder(**cfg["dec"]).eval(), PARAM_SEEDS[1])
        data = F.hot_path_inputs(name, INPUT_SEED)
        t0 = time.time()
        with torch.inference_mode():
            z = enc(data["inputs"], enc.latents(data["inputs"]), input_mask=data["input_mask"])
            out = dec(data["query"], z, query_mask=data["query_mask"])
        t_ref = time.time() - t0
        t0 = time.time()
        z_o, out_o = F.oracle_forward(name, dict(enc.state_dict()), dict(dec.state_dict()), data)
        t_orc = time.time() - t0
        ez = float((z_o - z).abs().max() / z.abs().max())
        eo = float((out_o - out).abs().max() / out.abs().max())
        assert ez < 1e-5 and eo < 1e-5, (name, ez, eo)
        sz, so = F.SUBSAMPLE[name]
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(
            path, latents=F.subsample(z, sz).numpy(), output=F.subsample(out, so).numpy(),
            latents_absmax=np.float64(z.abs().max()), output_absmax=np.float64(out.abs().max()),
            latents_shape=np.asarray(z.shape), output_shape=np.asarray(out.shape),
            latents_step=np.int64(sz), output_step=np.int64(so),
            param_seeds=np.asarray(PARAM_SEEDS), input_seed=np.int64(INPUT_SEED))
        print(f"{name}: reference {t_ref:.1f} s, oracle {t_orc:.1f} s (oracle vs reference: latents {ez:.1e}, output "
              f"{eo:.1e}); out {tuple(out.shape)} absmax {float(out.abs().max()):.3f}; "
              f"{os.path.getsize(path) / 1024:.0f} KiB", flush=True)


if __name__ == "__main__":
    main()
