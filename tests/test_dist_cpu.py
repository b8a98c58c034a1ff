"""World-size-2 gloo test of the N > 1 plumbing: contiguous frame shards, bs_norm-independent
results, one all-gather in frame order."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from ihmr_b200 import dist as idist


def test_shard_ranges_cover_everything():
    for total, world in [(65536, 8), (10, 4), (7, 2), (3, 8)]:
        got = []
        for r in range(world):
            s, c = idist.shard_range(total, r, world)
            got += list(range(s, s + c))
        assert got == list(range(total))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = idist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    start, count = idist.shard_range(total, rank, world)
    from ihmr_b200 import synthetic
    raw = synthetic.make_raw_frames(start, count, seed=0)
    # stand-in for the refined rows of this shard: a deterministic function of the frame inputs
    params = torch.from_numpy(np.concatenate([raw["cam"], raw["init_trans"], raw["init_pose"], raw["init_shape"]], 1))
    local = idist.pack_results(params, params[:, 0] * 2, params[:, 1] * 3)
    full = idist.all_gather_results(local, total)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    total = 37                                   # uneven split: 19 + 18
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    from ihmr_b200 import synthetic
    raw = synthetic.make_raw_frames(0, total, seed=0)
    params = torch.from_numpy(np.concatenate([raw["cam"], raw["init_trans"], raw["init_pose"], raw["init_shape"]], 1))
    want = idist.pack_results(params, params[:, 0] * 2, params[:, 1] * 3)
    for r in range(2):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert got.shape == (total, 124) and torch.equal(got, want)
