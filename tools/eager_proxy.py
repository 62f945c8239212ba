"""Reference-eager proxy on the GPU: the oracle's plain torch ops (pinned to the reference at 1e-5)
run on cuda in fp32 / TF32 / autocast-fp16, timed with CUDA events on the bench workload.  This is the
denominator of the north_star's '>= 6x over the reference PyTorch-eager path' (the reference itself
cannot travel to the GPU box).   python tools/eager_proxy.py [samples]"""
import copy, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cobevt_b200 as cb
from cobevt_b200 import configs, synth
from oracle import cobevt_oracle as orc

samples = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = configs.north_star()
model = cb.CorpBEVT(copy.deepcopy(cfg))
sd = {k: v.cuda() for k, v in synth.synth_state_dict(model, seed=0).items()}
batch = synth.to_device(synth.synth_batch(cfg, record_len=(5,) * samples, n_cams=6, seed=1), "cuda")

def run(tag, ctx):
    with torch.no_grad(), ctx:
        for _ in range(2):
            orc.corpbevt(dict(batch), sd, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            orc.corpbevt(dict(batch), sd, cfg)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{tag:28s} {ms:9.2f} ms/step  {samples / ms * 1e3:8.2f} samples/s   peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")

import contextlib
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
run("eager fp32 (no TF32)", contextlib.nullcontext())
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
run("eager fp32 + TF32", contextlib.nullcontext())
run("eager autocast fp16", torch.autocast("cuda", dtype=torch.float16))
