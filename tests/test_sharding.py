"""Host-side logic of the N > 1 path on CPU: world-size-2 gloo run (weight broadcast from rank 0,
read sharding, max-over-ranks timing reduction)."""
import hashlib
import os
import socket

import numpy as np
import pytest

from scrappie_b200 import WEIGHTS_DIR
from scrappie_b200.sharding import shard_reads


def test_shards_partition_and_balance():
    rng = np.random.default_rng(3)
    lens = np.clip(rng.lognormal(np.log(8000), 1.0, size=500), 1000, 200000).astype(int)
    for world in (1, 2, 4, 8):
        shards = [shard_reads(lens, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(len(lens)))                # every read exactly once
        loads = np.array([lens[s].sum() for s in shards])
        assert loads.max() - loads.min() <= lens.max()                     # LPT bound
    eq = [shard_reads([4000] * 1024, r, 8) for r in range(8)]
    assert all(len(s) == 128 for s in eq)


def test_plan_batches_covers_and_bounds():
    from scrappie_b200.sharding import lognormal_lengths, plan_batches
    lens = lognormal_lengths(700)
    assert lens.min() >= 1000 and lens.max() <= 200000 and (lens % 5 != 0).any()
    batches = plan_batches(lens, max_reads=256, max_samples=1 << 20)
    assert np.array_equal(np.sort(np.concatenate(batches)), np.arange(700))
    for b in batches:
        assert len(b) <= 256
        assert lens[b].sum() <= (1 << 20) or len(b) == 1
        assert (np.diff(lens[b]) <= 0).all()                 # sorted: CTA groups are homogeneous
    assert len(plan_batches([300000, 10, 10], max_samples=1000)) == 2


def _worker(rank, world, port, blob_path, out_dir):
    import torch.distributed as dist
    from scrappie_b200.sharding import broadcast_blob, max_over_ranks, shard_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # only rank 0 is given a readable path: the others must get the bytes over the wire
    blob = broadcast_blob(blob_path if rank == 0 else "/nonexistent/weights.bin", rank, dist)
    mine = shard_reads([4000 + (i % 7) * 500 for i in range(101)], rank, world)
    t = max_over_ranks([1.0 + rank, 5.0 - rank], dist)
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write("%s %s %r\n" % (hashlib.md5(blob.tobytes()).hexdigest(), ",".join(map(str, mine)), t))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_shards_gloo_world2(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    blob_path = os.path.join(WEIGHTS_DIR, "rgrgr_r94.bin")
    mp.spawn(_worker, args=(2, port, blob_path, str(tmp_path)), nprocs=2, join=True)
    want = hashlib.md5(open(blob_path, "rb").read()).hexdigest()
    seen = []
    for rank in range(2):
        md5, idx, t = open(tmp_path / ("rank%d.txt" % rank)).read().split(" ", 2)
        assert md5 == want
        seen += [int(x) for x in idx.split(",")]
        assert eval(t) == [2.0, 5.0]
    assert sorted(seen) == list(range(101))
