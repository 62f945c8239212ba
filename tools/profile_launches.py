"""Per-launch CUDA-event timing of one north_star step (4 samples): python tools/profile_launches.py"""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobevt_b200 as cb
from cobevt_b200 import configs, synth, ops
samples = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = configs.north_star()
model = cb.CorpBEVT(copy.deepcopy(cfg)).eval()
model.load_state_dict(synth.synth_state_dict(model, seed=0))
model = model.cuda()
batch = synth.to_device(synth.synth_batch(cfg, record_len=(5,) * samples, n_cams=6, seed=1), "cuda")
with torch.no_grad():
    for _ in range(3):
        model(batch)
    torch.cuda.synchronize()
    ops.STATS.reset(); ops.STATS.profile = True
    model(batch)
    torch.cuda.synchronize()
tot = 0
for fam, label, fl, a, b in ops.STATS.detail:
    ms = a.elapsed_time(b); tot += ms
    if True:
        print(f"{ms*1e3:9.1f} us {fl/1e9:8.2f} GF {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s  {fam:12s} {label}")
print("total ms", tot)
