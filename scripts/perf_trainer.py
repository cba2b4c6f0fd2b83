"""Wall time of one Trainer.train_view iteration (host glue + AdamW) vs the device time of its fused step."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codenerf_b200 import synthetic as syn
from codenerf_b200.trainer import Trainer
HP = {"net_hyperparams": dict(syn.SRN_NET), "N_samples": 64, "near": 0.8, "far": 1.8, "loss_reg_coef": 1e-4,
      "lr_schedule": [{"type": "step", "lr": 1e-4, "interval": 250000}, {"type": "step", "lr": 1e-3, "interval": 250000}]}
for H in (64, 128):
    tr = Trainer(HP, n_objects=8, device="cuda", precision="bf16")
    focal = torch.tensor([131.25], dtype=torch.float64)
    imgs = torch.rand(1, H * H, 3, device="cuda")
    poses = torch.from_numpy(np.stack([syn.look_at_pose(3, 1.3)])).cuda()
    for _ in range(5): tr.train_view(focal, H, H, imgs, poses, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 30
    for _ in range(n): tr.train_view(focal, H, H, imgs, poses, 1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"view {H}x{H} ({H*H} rays): {dt*1e3:.2f} ms per iteration = {H*H/dt/1e6:.2f} Mrays/s")
