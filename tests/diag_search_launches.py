"""Three search-stage train steps for an ncu launch list (profiles/r01/launches_search_fit.txt):
ncu --metrics gpu__time_duration.sum --clock-control none python tests/diag_search_launches.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npp_b200  # noqa: E402,F401
from npp_b200.plan import EncoderSpec, Plan, MODEL_LIGHT  # noqa: E402

rng = np.random.default_rng(0)
enc = EncoderSpec.from_proposals((512, 512), [[83.0, 172.5]], [[42.7, 38.4]], (rng.standard_normal(10) * 10).astype(np.float32),
                                 include_input=False)
plan = Plan(enc, depth=4, width=256, skip_layer=-1, max_rows=2048, model=MODEL_LIGHT)
plan.reset_parameters(0)
coords = torch.randint(0, 512, (2048, 2), device="cuda").float()
target = torch.rand(2048, 3, device="cuda")
loss = torch.zeros((), device="cuda")
for _ in range(3):
    plan.train_step(coords, target, None, 5e-4, loss)
torch.cuda.synchronize()
print("loss", loss.item())
