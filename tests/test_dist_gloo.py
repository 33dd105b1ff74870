"""world_size-2 test of the multi-GPU host logic on CPU (gloo): shard by string,
per-rank match bits, ONE all-reduce of the zero-initialised bitmap == bitwise OR.
The per-rank scan itself is stood in for by the oracle (test infrastructure);
the sharding, bitmap packing and the collective are the product code under test."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    import torch
    import torch.distributed as dist
    from conftest import GOLDEN
    from pire_b200 import workloads as W
    from pire_b200.dist import merge_match_bits, popcount_bits, shard_bounds
    from refpire import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = next(c for c in GOLDEN if c.name == "AppendixA")
    orc = Oracle(case.image)
    spec = W.SynthSpec(n, 256, plants=W.HEADLINE_PLANTS)
    shard, lo = spec.shard(rank, world)
    assert (lo, lo + shard.n_strings) == shard_bounds(n, rank, world)
    final, _, _ = orc.run(shard.host_sample(0, shard.n_strings), fixed_len=256, n=shard.n_strings)
    pad = (-len(final)) % 32
    words = np.packbits(np.concatenate([final, np.zeros(pad, np.uint8)]), bitorder="little").view(np.int32)
    full = merge_match_bits(torch.from_numpy(words.copy()), lo, n)
    np.save(os.path.join(out_dir, "bits%d.npy" % rank), full.numpy())
    assert popcount_bits(full) == (n + 7) // 8
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096, 1000, 24])      # 24 < 32 * world: rank 1's shard is empty but still joins the collective
def test_sharded_bitmap_allreduce(tmp_path, n):
    import torch.multiprocessing as mp
    from conftest import GOLDEN
    from pire_b200 import workloads as W
    from refpire import Oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "bits0.npy")
    b = np.load(tmp_path / "bits1.npy")
    assert (a == b).all()
    case = next(c for c in GOLDEN if c.name == "AppendixA")
    spec = W.SynthSpec(n, 256, plants=W.HEADLINE_PLANTS)
    final, _, _ = Oracle(case.image).run(spec.host_sample(0, n), fixed_len=256, n=n)
    got = np.unpackbits(a.view(np.uint8), bitorder="little")[:n]
    assert (got == final).all()
