"""world_size-2 gloo run of the multi-GPU host logic: clips are sharded by rank, each rank
accumulates its statistics vector, one all-reduce gives every rank the global totals."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import D

clips = D.clips
N_CLIPS = 6


def _fake_result(clip_id):
    g = torch.Generator().manual_seed(100 + clip_id)
    x = torch.rand(1, 3, 64, 64, generator=g)
    xh = (x + 0.03 * torch.randn(1, 3, 64, 64, generator=g)).clamp(0, 1)
    mask = (torch.rand(1, 1, 64, 64, generator=g) > 0.8).float()
    bpp_y = torch.tensor([1.0 + 0.1 * clip_id])
    bpp_z = torch.tensor([0.1])
    return {"dpb": {"frame": xh}, "bpp": bpp_y + bpp_z, "bpp_y": bpp_y, "bpp_z": bpp_z}, x, mask


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = clips.ClipStats("cpu")
    mine = clips.shard_clips(N_CLIPS, rank, world)
    for c in mine:
        res, x, mask = _fake_result(c)
        st.add_frame(res, x, mask)
    local = st.vec.clone()
    st.all_reduce()
    q.put((rank, mine, local.tolist(), st.vec.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stats_all_reduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = clips.ClipStats("cpu")
    for c in range(N_CLIPS):
        res, x, mask = _fake_result(c)
        ref.add_frame(res, x, mask)
    seen = []
    for rank, mine, local, total in got:
        seen += mine
        assert torch.allclose(torch.tensor(total, dtype=torch.float64), ref.vec, rtol=1e-12, atol=1e-9)
        assert local[6] == len(mine)
    assert sorted(seen) == list(range(N_CLIPS))
