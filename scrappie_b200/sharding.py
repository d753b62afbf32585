"""Multi-GPU plumbing of the raw basecalling path: reads shard across ranks, weights are broadcast once.

The reference parallelises over reads with OpenMP (src/scrappie_raw.c:355-387); here each rank (one process per
GPU) owns a shard of the reads and there is no collective on the per-read path.  The only exchange is the model
weight blob at start-up: rank 0 reads it, every other rank receives it (`torch.distributed` broadcast: NCCL between
GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_reads(lengths, rank, world):
    """Indices of the reads rank `rank` owns.  Reads are dealt longest-first to the currently least loaded rank
    (greedy LPT), so the shards carry nearly equal numbers of samples even for log-normal read lengths; for
    equal-length reads this degenerates to round-robin.  Deterministic, identical on every rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))        # ties -> lowest rank
        owner[i] = r
        load[r] += lengths[i]
    return np.flatnonzero(owner == rank)


def plan_batches(lengths, max_reads=256, max_samples=1 << 20):
    """Dynamic batching for mixed-length reads: returns a list of index arrays, one per batch.

    Reads are sorted by length (longest first) and cut into consecutive runs of at most `max_reads` reads and
    `max_samples` samples, so that (i) the reads that step together inside one scan CTA (groups of 4 consecutive
    reads of a batch) have nearly the same number of blocks -- a GRU layer costs max(T) steps per group -- and
    (ii) a batch's device workspace (about 1.5 KB per sample for rgrgr_r94) stays bounded.  A read longer than
    `max_samples` gets a batch of its own: a read cannot be split, its recurrence is serial
    (src/layers.c:373-470)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    batches, cur, cur_samples = [], [], 0
    for i in order:
        n = int(lengths[i])
        if cur and (len(cur) >= max_reads or cur_samples + n > max_samples):
            batches.append(np.array(cur, dtype=np.int64))
            cur, cur_samples = [], 0
        cur.append(int(i))
        cur_samples += n
    if cur:
        batches.append(np.array(cur, dtype=np.int64))
    return batches


def lognormal_lengths(nreads, seed=4, median=8000.0, sigma=1.0, lo=1000, hi=200000):
    """Read lengths of BASELINE config 4: log-normal, clipped to [1k, 200k] samples, seeded."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.lognormal(np.log(median), sigma, size=nreads), lo, hi).astype(np.int64)


def broadcast_blob(path, rank, dist=None, device="cpu"):
    """The weight blob as a uint8 numpy array on every rank; only rank 0 touches the file system."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.fromfile(path, dtype=np.uint8)
    if rank == 0:
        blob = torch.from_numpy(np.fromfile(path, dtype=np.uint8)).to(device)
        size = torch.tensor([blob.numel()], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, 0)
    if rank != 0:
        blob = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, 0)
    return blob.cpu().numpy()


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (timing reduction)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
