"""world_size-2 gloo tests (CPU) of the data-parallel host logic: shard disjointness, mean all-reduce of the flat gradient
buffer, max-over-ranks timing reduction, lr schedule; plus the augmentation planner's RNG parity with the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from neuspeech1_b200.parallel import DataParallel, linear_warmup_decay


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dp = DataParallel(backend="gloo")
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)           # rank-dependent "gradient"
    dp.all_reduce_mean(flat)
    shard = dp.shard(103, epoch=3, seed=7)
    tmax = dp.max_over_ranks(10.0 + rank)
    dp.barrier()
    q.put((rank, flat[:5].tolist(), float(flat.sum()), shard, tmax))
    dp.close()


def test_two_rank_gloo_allreduce_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = torch.arange(1000, dtype=torch.float32) * 1.5                   # mean of x*1 and x*2
    for rank, head, total, shard, tmax in res:
        assert head == ref[:5].tolist() and abs(total - float(ref.sum())) < 1e-3
        assert tmax == 11.0
    s0, s1 = set(res[0][3]), set(res[1][3])
    assert not (s0 & s1) and len(s0) == len(s1) == 51 and len(s0 | s1) == 102      # drop_last: 103 -> 102


def test_single_process_is_identity():
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        os.environ.pop(k, None)
    dp = DataParallel()
    g = torch.ones(4)
    assert dp.all_reduce_mean(g) is g and dp.shard(10, shuffle=False) == list(range(10)) and dp.max_over_ranks(3.0) == 3.0


def test_linear_schedule_matches_hf():
    from transformers import get_linear_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1e-3)
    sch = get_linear_schedule_with_warmup(opt, 5, 20)
    for step in range(1, 20):
        opt.step(); sch.step()
        assert abs(sch.get_last_lr()[0] - linear_warmup_decay(step, 1e-3, 5, 20)) < 1e-12


def test_batch_augmenter_plan_matches_oracle_rng():
    """The product's planner consumes the RNG streams exactly like the oracle's draw_plan (= like utils/reader.py)."""
    from neuspeech1_b200.augment_eeg import BatchAugmenter, RandomShapeMasker
    from oracle import augment as A
    cfg = {"noise": {"prob": 0.5, "min_snr_dB": 20, "max_snr_dB": 50},
           "mask": {"prob": 0.7, "kwargs": {"unit": [1, 40], "mask_prob": 0.25, "random_type": 1}},
           "taylor": {"prob": 0.5}, "shift": {"prob": 0.5}}
    shapes = [(8, 700), (8, 1234), (8, 400), (8, 999)]
    torch.manual_seed(11); np.random.seed(11)
    ref = [A.draw_plan(s, cfg, max_length=6000) for s in shapes]
    torch.manual_seed(11); np.random.seed(11)
    kw = BatchAugmenter(cfg).plan(shapes, "cpu")
    assert kw["shift"].tolist() == [p.shift for p in ref]
    assert kw["e0"].tolist() == [p.edge0 for p in ref] and kw["e1"].tolist() == [p.edge1 for p in ref]
    assert kw["flags"].tolist() == [(1 if p.grid is not None else 0) | (2 if p.noise else 0) for p in ref]
    for b, p in enumerate(ref):
        if p.grid is not None:
            assert torch.equal(kw["grid"][b, :p.grid.numel()].float(), p.grid.reshape(-1))
    torch.manual_seed(5)
    m = RandomShapeMasker(unit=[1, 40], mask_prob=0.25, random_type=2)((16, 300))
    torch.manual_seed(5)
    assert torch.equal(m, A.shape_mask((16, 300), [1, 40], 0.25, 2))
