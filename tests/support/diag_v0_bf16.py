"""bf16 error of the tiny mFormerV0 against its fp32 oracle over several data seeds (per head, max-abs relative)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import linnaeus_b200 as L
from oracle import mformer_v0_oracle as V0

dev = "cuda"
cfg0, nc0 = L.make_synthetic_config_v0("sm", 64, conv_embed=(16, 32), conv_out=(32, 64), conv_depths=(1, 2), conv_strides=((2,), (1, 1)),
                                       attn_dims=(64, 128), attn_depths=(2, 1), heads=(2, 4))
a0 = V0.arch_from_config(cfg0, nc0)
P0 = V0.synth_state_dict(a0, 0)
net = L.build_model(cfg0, nc0)
net.load_state_dict(P0)
net = net.to(dev).eval()
with torch.no_grad():
    for seed in range(6):
        for B in (2, 3):
            x0, m0 = V0.synth_batch(a0, B, seed)
            ref0 = V0.forward(P0, a0, x0, m0)
            row = []
            for dtype in (torch.float32, torch.bfloat16):
                got = net.set_compute_dtype(dtype)(x0.to(dev), m0.to(dev))
                errs = [float((got[k].float().cpu() - r).abs().max() / r.abs().max()) for k, r in ref0.items()]
                allr = torch.cat([r for r in ref0.values()], 1)
                allg = torch.cat([got[k].float().cpu() for k in ref0], 1)
                row.append((max(errs), float((allg - allr).abs().max() / allr.abs().max())))
            print(f"seed {seed} B {B}: fp32 per-head max {row[0][0]:.2e} | bf16 per-head max {row[1][0]:.4f}  over all heads {row[1][1]:.4f}  "
                  f"head maxabs {[round(float(r.abs().max()), 3) for r in ref0.values()]}")
