import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linnaeus_b200.functional as F
dev = "cuda"
shapes = {"s0a": (256 * 3136, 384, 96), "s0b": (256 * 3136, 96, 384), "s2a": (51200, 1536, 384), "s2b": (51200, 384, 1536), "s2c": (51200, 1152, 384), "s3": (13568, 3072, 768), "s1a": (200704, 768, 192), "s1b": (200704, 192, 768), "s2d": (51200, 384, 384), "s3b": (13568, 768, 3072), "s3c": (13568, 2304, 768), "hd": (256, 1576, 768), "s3d": (13568, 768, 768)}
which = sys.argv[1]
Mr, No, Ki = shapes[which]
dyw = torch.randn(Mr, No, device=dev).bfloat16()
xw = torch.randn(Mr, Ki, device=dev).bfloat16()
dww = torch.zeros(No, Ki, device=dev)
dbw = torch.zeros(No, device=dev)
for _ in range(3):
    F.wgrad(dyw, xw, out=dww, db_out=dbw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    F.wgrad(dyw, xw, out=dww, db_out=dbw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ref = dyw.float().t() @ xw.float()
err = float(((dww / 13) - ref).abs().max() / ref.abs().max())
print(f"{which} cfg={os.environ.get('LNX_WGRAD_CFG','default'):14s} {ms:.3f} ms {2*Mr*No*Ki/ms/1e9:.0f} TF/s err={err:.1e}")
