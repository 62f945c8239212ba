"""One north_star step for profilers (ncu): python tools/profile_step.py [samples] [steps]."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobevt_b200 as cb  # noqa: E402
from cobevt_b200 import configs, synth  # noqa: E402

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = configs.north_star()
model = cb.CorpBEVT(copy.deepcopy(cfg)).eval()
model.load_state_dict(synth.synth_state_dict(model, seed=0))
model = model.cuda()
batch = synth.to_device(synth.synth_batch(cfg, record_len=(5,) * samples, n_cams=6, seed=1), "cuda")
with torch.no_grad():
    for _ in range(steps):
        out = model(batch)
torch.cuda.synchronize()
print("done", out["dynamic_seg"].shape)
