"""CPU, gloo, world size 2: the host-side logic of the multi-GPU path -- length-balanced
sharding of independent utterances, weight broadcast at load, in-order gather.  No collective on the data path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsv_tts import _shard


def test_shard_by_length_is_a_balanced_partition():
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(40, 400, (257,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        shards = _shard.shard_by_length(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))                       # a partition
        counts = [len(s) for s in shards]
        assert max(counts) - min(counts) <= 1
        tot = [sum(lengths[i] for i in s) for s in shards]
        assert (max(tot) - min(tot)) <= max(lengths)                   # serpentine bound
    assert _shard.shard_by_length([], 4) == [[], [], [], []]
    assert _shard.shard_by_length([5, 5, 5], 2) == [[0], [1, 2]]      # deterministic ties


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weight broadcast: rank 0 holds the checkpoint, the others start from garbage
        g = torch.Generator().manual_seed(1)
        ref = {"a.weight": torch.randn(7, 5, generator=g), "b.bias": torch.randn(3, generator=g)}
        sd = ref if rank == 0 else {k: torch.full_like(v, float("nan")) for k, v in ref.items()}
        got = _shard.broadcast_state_dict(sd, 0)
        ok_bcast = all(torch.equal(got[k], ref[k]) for k in ref)
        # checkpoint load under torch.distributed: only rank 0 touches the file system (Loader.get_*_weights)
        def read():
            assert rank == 0, "only the source rank may read the checkpoint"
            return {"model": {"hidden_dim": 512}}, ref
        meta, got2 = _shard.broadcast_checkpoint(read, 0)
        ok_bcast = ok_bcast and meta == {"model": {"hidden_dim": 512}} and all(torch.equal(got2[k], ref[k]) for k in ref)
        # sharded "inference": every request's result is a function of the request only
        lengths = [50 + 7 * (i % 11) for i in range(23)]
        mine = _shard.shard_by_length(lengths, world)[rank]
        local = {i: torch.arange(lengths[i]) * (i + 1) for i in mine}
        allr = _shard.gather_in_order(local, len(lengths))
        ok_gather = all(torch.equal(allr[i], torch.arange(lengths[i]) * (i + 1)) for i in range(len(lengths)))
        q.put((rank, ok_bcast, ok_gather, len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_broadcast_shard_gather_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok_b and ok_g for _, ok_b, ok_g, _ in res)
    assert sorted(n for *_, n in res) == [11, 12]
