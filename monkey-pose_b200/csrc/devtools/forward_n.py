"""Development helper: a few pose forwards at a given batch / width, for captures under ncu.
    python monkey-pose_b200/csrc/devtools/forward_n.py [--batch 256] [--channels 25] [--timesteps 8] [--reps 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
import torch  # noqa: E402

import monkey_pose_b200 as mp  # noqa: E402
from monkey_pose_b200 import initialization as init  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--channels", type=int, default=25)
ap.add_argument("--timesteps", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--mode", default="bf16")
a = ap.parse_args()
P = init.pose_params(channels=a.channels, S=15, T=a.timesteps, hw=64, fc_hidden=1024, out=69, seed=3)
depth = torch.as_tensor(init.synthetic_depth(a.batch, seed=1234)).cuda()
m = mp.model()
m.channels, m.timesteps, m.compute_mode = a.channels, a.timesteps, a.mode
m.hidden_state = torch.as_tensor(init.hidden_init((a.batch, 64, 64, a.channels), seed=5)).cuda()
m.load_params(P)
for _ in range(a.reps):
    out = m.build(depth, 69)
torch.cuda.synchronize()
print("launches per forward:", m.gpu_launches, "finite:", bool(torch.isfinite(out).all()))
