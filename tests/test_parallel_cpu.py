"""world_size-2 gloo test of the agent-row sharding + all-gather plumbing (parallel.py) on CPU.
The compute steps are the oracle's (test infrastructure); what is verified is that the sharded
forward is IDENTICAL to the single-process forward for aligned, straddling and ragged layouts."""
import copy
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, record_len, n_cams, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cobevt_b200 as cb
        from cobevt_b200 import configs, synth
        from cobevt_b200.parallel import sharded_forward, row_shard
        from oracle import cobevt_oracle as orc
        cfg = configs.small()
        model = cb.CorpBEVT(copy.deepcopy(cfg))
        sd = synth.synth_state_dict(model, seed=0)
        full = synth.synth_batch(cfg, record_len=record_len, n_cams=n_cams, seed=1)
        total = sum(record_len)
        r0, r1 = row_shard(total, world, rank)
        local = {k: full[k][r0:r1] for k in ("inputs", "intrinsic", "extrinsic")}

        def encode(lb):
            enc = cfg["encoder"]
            feats = orc.resnet_encoder(lb["inputs"], sd, "encoder.", enc["num_layers"], enc["id_pick"])
            return orc.fax_module(feats, lb["intrinsic"], lb["extrinsic"], sd, "fax.", cfg["fax"]).squeeze(1)

        def fuse(tiles, s0, s1):
            if s1 <= s0:
                return None
            st = cfg["sttf"]
            y, mask = orc.regroup(tiles, record_len[s0:s1], cfg["max_cav"])
            tm = full["transformation_matrix"][s0:s1]
            y = orc.sttf(y, tm, st["resolution"], st["downsample_rate"])
            cm = orc.roi_and_cav_mask(y.shape, mask, tm, st["resolution"], st["downsample_rate"])
            y = orc.swap_fusion_encoder(y.permute(0, 1, 4, 2, 3), cm, sd, "fusion_net.", cfg["fax_fusion"])
            y = orc.naive_decoder(y[:, None], sd, "decoder.", cfg["decoder"]["num_layer"])
            return orc.bev_seg_head(y.reshape(-1, *y.shape[2:]), sd, "seg_head.", cfg["target"], s1 - s0, 1)["dynamic_seg"]

        with torch.no_grad():
            out = sharded_forward(encode, fuse, local, list(record_len), rank, world)
            outs = [None] * world
            dist.all_gather_object(outs, out)
            if rank == 0:
                whole = orc.corpbevt(copy.deepcopy(full), sd, cfg)["dynamic_seg"]
                got = torch.cat([o for o in outs if o is not None], 0)
                ret["maxdiff"] = (got - whole).abs().max().item()
                ret["shape_ok"] = got.shape == whole.shape
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("record_len", [(2, 2), (3, 1), (1, 2, 3), (5,), (1, 1, 1)])
def test_sharded_equals_single_process(record_len):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() + sum(record_len) * 7 + len(record_len)) % 2000
    mp.spawn(_worker, args=(world, port, record_len, 2, ret), nprocs=world, join=True)
    assert ret["shape_ok"]
    # identical math on identical rows; row-batched BLAS may reorder sums -> allow fp32 noise
    assert ret["maxdiff"] < 1e-5, ret["maxdiff"]


def test_shard_math():
    from cobevt_b200.parallel import row_shard, needs_exchange
    assert [row_shard(20, 8, r) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 14), (14, 16), (16, 18), (18, 20)]
    assert not needs_exchange([5, 5, 5, 5], 2)
    assert not needs_exchange([5] * 8, 8)
    assert needs_exchange([3, 5, 5, 5, 2], 2)
    assert needs_exchange([5, 5, 5, 5], 8)        # C4: 20 rows over 8 ranks
    assert not needs_exchange([5], 1)
