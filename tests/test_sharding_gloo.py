"""N>1 host logic on CPU: two gloo ranks shard a recording by frame range, each produces its shard's stream
(with the oracle standing in for the GPU encoder -- this test is about the sharding and the size exchange),
sizes are all-gathered, and the concatenation at the exchanged offsets equals the single-process stream."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import x3_oracle as oracle
    sharding = importlib.import_module("x3-rust_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm = oracle.synth(2, 0x58330002, 384000, 0, n)
    s0, s1 = sharding.shard_frames(n, 10000, rank, world)
    stream, _ = oracle.encode(pcm[s0:s1])
    sizes, base = sharding.exchange_sizes(stream.size, dist)
    pending = sharding.exchange_sizes_begin(stream.size, dist)      # the overlapped form bench.py uses
    sizes2, base2 = sharding.exchange_sizes_end(pending)
    assert sizes2 == sizes and base2 == base
    np.save(os.path.join(tmpdir, "shard%d.npy" % rank), stream)
    np.save(os.path.join(tmpdir, "meta%d.npy" % rank), np.array([s0, s1, base] + sizes, dtype=np.int64))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [123456, 10000, 35001])
def test_two_rank_frame_sharding(tmp_path, oracle, n):
    world, port = 2, 29500 + (os.getpid() + n) % 2000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    pcm = oracle.synth(2, 0x58330002, 384000, 0, n)
    ref, _ = oracle.encode(pcm)
    out = np.zeros(ref.size, dtype=np.uint8)
    covered = 0
    for r in range(world):
        meta = np.load(tmp_path / ("meta%d.npy" % r))
        shard = np.load(tmp_path / ("shard%d.npy" % r))
        s0, s1, base = int(meta[0]), int(meta[1]), int(meta[2])
        assert s0 % 10000 == 0 and list(meta[3:]) == list(np.load(tmp_path / "meta0.npy")[3:])
        out[base:base + shard.size] = shard
        covered += s1 - s0
    assert covered == n and np.array_equal(out, ref)


def test_deal_files():
    sharding = importlib.import_module("x3-rust_b200.sharding")
    deal = sharding.deal_files([5760] * 1024, 8)
    assert [len(d) for d in deal] == [128] * 8 and deal[3][0] == 384
    deal = sharding.deal_files([10, 1, 1, 1, 10, 3], 2)
    assert sorted(sum(deal, [])) == list(range(6)) and all(deal)
    assert sharding.shard_frames(1382400000, 10000, 7, 8) == (1209600000, 1382400000)
    assert sharding.shard_frames(25000, 10000, 0, 2) == (0, 10000) and sharding.shard_frames(25000, 10000, 1, 2) == (10000, 25000)
