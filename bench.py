#!/usr/bin/env python
"""Benchmark of the raw basecalling hot path (BASELINE.json: raw samples/s, rgrgr_r94).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

One "step" = one pass of the hot path (network forward + Viterbi decode) over the whole workload on each GPU:
`--reads` synthetic reads of `--samples` samples, processed in batches of `--batch` reads that run concurrently on
their own CUDA streams (BASELINE config 2: rgrgr_r94, 1024 x 4000-sample reads, batch 256).  With N > 1 (torchrun,
one rank per GPU) every rank processes its own `--reads` reads (weak scaling); rank 0 loads the weight blob and
broadcasts it over NCCL at start-up; there is no collective on the per-read path.

  value   samples/s with the signals already resident in HBM; device time from CUDA events on the launching stream,
          K steps streamed back to back over `--sets` buffer sets, max over ranks.
  e2e     samples/s through the documented drop-in call, sb2_basecall_batch, from ORDINARY
          (pageable) host arrays: staging into pinned memory, H2D, forward, decode, homopolymer fix-up + overlapper on
          the device, D2H of the base strings, all inside the timed region (wall clock, max over ranks).  Workspaces
          come from the engine's pool, so the call allocates nothing in steady state.  `e2e.persistent` is the same
          work through caller-owned sb2_batch objects with pre-filled pinned buffers (round 1's figure).
  parity  base strings of the timed e2e run compared with the reference's own CPU implementation (oracle/_ref) on the
          same reads; a mismatch makes the run exit non-zero.
  other_configs   BASELINE configs 3 (rnnrf_r94) and 4 (mixed lengths), and with N > 1 config 5 (100 000 reads dealt
          out by shard_reads), measured after the headline with the same method.
"""
import argparse
import json
import os
import queue
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# hardware work queues of the CUDA context (default 8): every batch in flight has its own stream and wants its own queue
# (scrappie_b200 sets the same default; here because torch creates the context first)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

FLOP_PER_READ_STEP_RECURRENT = {"rgrgr_r94": 55296, "rgrgr_r941": 55296, "rgrgr_r10": 55296, "rnnrf_r94": 75264}
# algorithmic HBM bytes per block when every intermediate is materialised once:
# conv out H*4 W; per layer X H*4 R + Xin 3H*4 W, then Xin 3H*4 R + X H*4 W; posterior stride*4 W + R; traceback W
BYTES_PER_BLOCK = {"rgrgr_r94": 96 * 4 + 5 * (2 * 96 * 4 + 2 * 288 * 4) + 2 * 1028 * 4 + 1028,
                   "rgrgr_r941": 96 * 4 + 5 * (2 * 96 * 4 + 2 * 288 * 4) + 2 * 1028 * 4 + 1028,
                   "rnnrf_r94": 112 * 4 + 5 * (3 * 112 * 4 + 2 * 336 * 4) + 2 * 28 * 4 + 8}
FLOP_PER_BLOCK_TOTAL = {"rgrgr_r94": 753408, "rnnrf_r94": 760704}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="rgrgr_r94")
    ap.add_argument("--reads", type=int, default=1024)
    ap.add_argument("--samples", type=int, default=4000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cpu-sample-reads", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="only the headline workload (default: configs 3 and 4, and config 5 when N > 1, are measured too)")
    ap.add_argument("--sustained-seconds", type=float, default=5.0,
                    help="length of the additional long run that shows the figure under sustained clocks (0 = skip)")
    ap.add_argument("--sets", type=int, default=4,
                    help="buffer sets: consecutive steps alternate between the sets and are not synchronised with each "
                         "other, so step n + 1 overlaps the tail of step n (1 = one set, L2 flushed between steps)")
    ap.add_argument("--workload", default="fixed", choices=["fixed", "mixed", "sharded"],
                    help="fixed: --reads reads of --samples samples (BASELINE config 2); mixed: --reads reads with "
                         "log-normal lengths in [1k, 200k] samples, length-bucketed dynamic batching (config 4); sharded: "
                         "--total-reads reads of --samples samples dealt out to the ranks by shard_reads (config 5)")
    ap.add_argument("--total-reads", type=int, default=100000)
    ap.add_argument("--shard-as", default=None, metavar="R/W",
                    help="sharded workload only: take the shard rank R of W ranks would own, whatever this job's size is "
                         "(reproduces one rank of an N-GPU run on one GPU)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from the warm-up to the end of the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_median": float(np.median(pw)) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(nreads, nsamples, seed0):
    from scrappie_b200.synthetic import synthetic_read
    return [synthetic_read(seed0 + i, nsamples) for i in range(nreads)]


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline / parity check (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------

def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def reference_library():
    """(path, -march label) of the compiled reference for THIS host: the x86-64-v4 (AVX-512) build when the CPU has
    it -- the closest shippable stand-in for the reference's -march=native (CMakeLists.txt:97) -- else x86-64-v3."""
    base = os.path.join(ROOT, "oracle", "_ref")
    v4 = os.path.join(base, "v4", "libref_bench.so")
    if os.path.exists(v4) and {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= _cpu_flags():
        return v4, "x86-64-v4"
    v3 = os.path.join(base, "libref_bench.so")
    return (v3, "x86-64-v3") if os.path.exists(v3) else (None, None)


_REF = {}


def _ref_lib():
    import ctypes as C
    if "lib" not in _REF:
        so, march = reference_library()
        L = None
        if so is not None:
            L = C.CDLL(so)
            L.ref_bench_run.restype = C.c_double
            L.ref_bench_run.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p]
            L.ref_bench_free.argtypes = [C.c_void_p]
            blas = None
            try:
                L.ref_bench_blas_config.restype = C.c_char_p
                blas = L.ref_bench_blas_config().decode().strip()
            except AttributeError:
                pass
            _REF["build"] = {"march": march, "blas": blas, "flags": "-O3 -march=%s -fopenmp -DUSE_SSE2 -DNDEBUG "
                             "(reference: -O3 -march=native, CMakeLists.txt:97)" % march}
        _REF["lib"] = L
    return _REF["lib"]


def run_reference(model, sigs, nthreads=0, want_bases=False):
    """Times oracle/_ref (the reference's own C sources + OpenBLAS, one read per OpenMP thread).
    Returns (seconds, nbases, nblocks, threads[, bases])."""
    import ctypes as C
    L = _ref_lib()
    if L is None:
        return None
    concat = np.concatenate(sigs).astype(np.float32)
    lens = np.array([len(s) for s in sigs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    nb, nk = C.c_size_t(0), C.c_size_t(0)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the
    # reference's own recommendation -- one read per core -- means; README.md:66-71)
    threads = nthreads or len(os.sched_getaffinity(0))
    out = (C.c_void_p * len(sigs))() if want_bases else None
    secs = L.ref_bench_run(model.encode(), concat.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(sigs),
                           threads, C.byref(nb), C.byref(nk), out)
    if not want_bases:
        return secs, nb.value, nk.value, threads
    bases = []
    for p in out:
        bases.append(C.string_at(p).decode() if p else None)
        if p:
            L.ref_bench_free(p)
    return secs, nb.value, nk.value, threads, bases


def run_oracle_port(model, sigs):
    """Fallback when oracle/_ref is absent: the scalar C restatement, 1 thread."""
    from oracle.oracle import Oracle
    o = Oracle()
    t0 = time.time()
    nb, bases = 0, []
    for s in sigs:
        _, _, b, _ = o.basecall_raw(model, s)
        bases.append(b)
        nb += len(b or "")
    return time.time() - t0, nb, 0, 1, bases


# posterior tolerance of the parity tests (tests/test_gpu_parity.py): log space; rnnrf's CRF scores on |x| <= 14
POSTERIOR_TOL = {"rnnrf_r94": 2.5e-4}


def explain_mismatch(eng, model, sig, got):
    """A read whose GPU base string differs from the reference's.  north_star's parity statement has two parts: the
    Viterbi decode is bit-exact GIVEN the posterior, and the posterior is within a stated fp32 tolerance.  Posteriors
    ~2e-5 apart (the distance the reference's own OpenBLAS build has from a plain-C fp32 evaluation) can resolve a
    near-tie of the Viterbi recursion differently, so a differing base string is a defect only if one of the two parts
    fails.  Checked here for this read: (1) the GPU posterior against the reference's, (2) the REFERENCE's decoder,
    homopolymer fix-up and overlapper applied to the GPU posterior against the GPU's base string."""
    from oracle.oracle import Reference
    ref = Reference()
    b = eng.batch(model, [len(sig)])
    b.upload([sig])
    b.forward()
    b.decode()
    gpost = b.posterior(0)
    gscore = float(b.paths()[1][0])
    b.close()
    rpost = ref.posterior(model, sig)
    ns = {"rnnrf_r94": 25, "rgrgr_r10": 4097}.get(model, 1025)
    err = float(np.abs(gpost[:, :ns] - rpost[:, :ns]).max())
    if model == "rnnrf_r94":
        rscore, _ = ref.decode_crf(rpost)
        _, path = ref.decode_crf(gpost)
        bases = ref.crfpath_to_basecall(path, gpost.shape[0])
    else:
        rscore, _ = ref.decode_transducer(rpost, ns)
        _, path = ref.decode_transducer(gpost, ns)
        path = ref.homopolymer_path(gpost, ns, path)
        bases, _ = ref.overlapper(path, ns - 1)
    tol = POSTERIOR_TOL.get(model, 1e-4)
    return {"posterior_max_abs_err": err, "posterior_tolerance": tol, "viterbi_score_gpu": gscore, "viterbi_score_reference": rscore,
            "reference_decoder_on_gpu_posterior_gives_gpu_bases": bases == got,
            "explained": bool(err <= tol and bases == got)}


def check_parity(model, sigs, got_bases, eng=None):
    """Base strings of the GPU run vs the CPU reference on the same reads (python/test/test_scrappy.py:72-75 makes
    the same comparison for the reference's own Python binding).  `bases_identical` is the plain comparison; a read
    that differs is examined by explain_mismatch, and `ok` is false -- the run exits non-zero -- unless every such
    read is a Viterbi near-tie between posteriors that agree within the tolerance."""
    res = run_reference(model, sigs, want_bases=True)
    against = "reference (oracle/_ref)"
    if res is None:
        res = run_oracle_port(model, sigs)
        against = "oracle port"
    want = res[4]
    bad = [i for i, (g, w) in enumerate(zip(got_bases, want)) if g != w]
    out = {"reads_checked": len(sigs), "bases_identical": not bad, "mismatching_reads": bad[:8], "against": against,
           "bases_checked": int(sum(len(w or "") for w in want)), "ok": not bad}
    if bad and eng is not None and against.startswith("reference"):
        out["near_ties"] = [dict(read=i, **explain_mismatch(eng, model, sigs[i], got_bases[i])) for i in bad[:8]]
        out["ok"] = len(bad) <= 8 and all(t["explained"] for t in out["near_ties"])
    return out


def cpu_baseline(model, sigs, nreads_sample):
    sample = (sigs * ((nreads_sample + len(sigs) - 1) // len(sigs)))[:nreads_sample]   # the workload's reads, repeated
    res = run_reference(model, sample)
    kind = "reference"
    if res is None:
        sample = sigs[:max(8, nreads_sample // 16)]
        res = run_oracle_port(model, sample)
        kind = "port"
    secs, nbases, _, threads = res[:4]
    nsamp = sum(len(s) for s in sample)
    out = {"value": nsamp / secs, "unit": "samples/s", "cores": threads, "kind": kind,
           "kbases_per_s": nbases / secs / 1e3,
           "sample": "%d of the workload's reads (%d samples), one read per OpenMP thread, 1 BLAS thread, %.2f s wall"
                     % (len(sample), nsamp, secs)}
    out.update(_REF.get("build", {}))
    return out


def main_reference(args, rank, world):
    if rank != 0:
        return
    # one step = one pass over the same workload as the GPU arm (args.reads reads), capped so that K steps stay bounded
    nstep_reads = min(args.reads, args.cpu_sample_reads)
    sigs = make_workload(nstep_reads, args.samples, 1000)
    have_ref = _ref_lib() is not None
    for _ in range(max(1, min(args.warmup, 1))):
        run_reference(args.model, sigs[:32]) if have_ref else None
    times, kind, threads, nbases = [], "reference", 1, 0
    for _ in range(args.steps):
        res = run_reference(args.model, sigs)
        if res is None:
            kind = "port"
            res = run_oracle_port(args.model, sigs[:32])
            nsamp = sum(len(s) for s in sigs[:32])
        else:
            nsamp = sum(len(s) for s in sigs)
        times.append(res[0]); nbases = res[1]; threads = res[3]
    ms = 1e3 * float(np.mean(times))
    value = nsamp / (ms / 1e3)
    cpu = {"value": value, "unit": "samples/s", "cores": threads, "kind": kind,
           "sample": "%d reads x %d samples per step" % (len(sigs), args.samples)}
    cpu.update(_REF.get("build", {}))
    line = {"impl": "reference", "metric": "raw samples/sec (%s)" % args.model, "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload label as the GPU arm's line (build_groups); what is specific to this arm stands beside it
            "config": {"workload": "%s raw, %d synthetic %d-sample reads per GPU, batch=%d (%d concurrent batches), forward + Viterbi decode"
                                   % (args.model, args.reads, args.samples, args.batch, (args.reads + args.batch - 1) // args.batch),
                       "cpu_arm": "%s: bounded sample of %d reads per step, one read per OpenMP thread, 1 BLAS thread" % (kind, len(sigs))},
            "kbases_per_s": nbases / (ms / 1e3) / 1e3,
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------

class Ranks(object):
    """torch.distributed plumbing of the bench: barrier, max / gather of timings."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def gather(self, value):
        """`value` of every rank, as a list (every rank gets it)."""
        if self.dist is None:
            return [float(value)]
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device="cuda")
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def spread(values):
    v = sorted(values)
    return {"min": v[0], "median": float(np.median(v)), "max": v[-1]}


def d2h_bytes(model, nread, max_nblock, max_nbase):
    # finish_on_device (csrc/engine.cu): base count + score per read, then either the whole base-string area
    # (batches whose area is <= 2 MB) or a 2-D copy as wide as the longest call
    klen = 1 if model == "rnnrf_r94" else (6 if model == "rgrgr_r10" else 5)
    stride = (klen * (max_nblock + 1) + 1 + 15) // 16 * 16
    area = nread * stride
    return nread * 8 + (area if area <= (2 << 20) else nread * ((max_nbase + 1 + 15) // 16 * 16))


def build_groups(args, model, workload, rank, world):
    """The rank's reads and their division into batches.  Returns (sigs, groups, description)."""
    import scrappie_b200 as sb  # noqa: F401
    from scrappie_b200.sharding import lognormal_lengths, plan_batches, shard_reads
    from scrappie_b200.synthetic import synthetic_read
    if workload == "mixed":
        lens = lognormal_lengths(args.reads, seed=4 + rank)
        sigs = [synthetic_read(1000 + rank * args.reads + i, int(n)) for i, n in enumerate(lens)]
        plan = plan_batches(lens, max_reads=args.batch, max_samples=args.batch * 4096)
        groups = [[sigs[i] for i in idx] for idx in plan]
        desc = ("%s raw, %d synthetic reads per GPU with log-normal lengths (median 8000, sigma 1.0, clipped to [1k, 200k] "
                "samples; %d samples in total, longest %d), length-bucketed dynamic batching into %d batches of <= %d reads"
                % (model, args.reads, sum(len(s) for s in sigs), max(len(x) for x in sigs), len(groups), args.batch))
    elif workload == "sharded":
        # config 5 literally: read i of the job has seed 1000 + i; the job's reads are dealt out by shard_reads and each
        # rank batches what it owns.  Only 2048 distinct signals are synthesised (read i uses signal i % 2048): the
        # content of a read does not change what the path costs, generating 100 000 of them in Python would.
        if args.shard_as:
            rank, world = (int(x) for x in args.shard_as.split("/"))
        owned = shard_reads([args.samples] * args.total_reads, rank, world)
        distinct = make_workload(min(2048, args.total_reads), args.samples, 1000)
        sigs = [distinct[int(i) % len(distinct)] for i in owned]
        groups = [sigs[i:i + args.batch] for i in range(0, len(sigs), args.batch)]
        desc = ("%s raw, %d synthetic %d-sample reads sharded over %d GPU(s) by shard_reads (%d on this rank, batches of %d)"
                % (model, args.total_reads, args.samples, world, len(sigs), args.batch))
    else:
        sigs = make_workload(args.reads, args.samples, 1000 + rank * args.reads)
        groups = [sigs[i:i + args.batch] for i in range(0, len(sigs), args.batch)]
        desc = ("%s raw, %d synthetic %d-sample reads per GPU, batch=%d (%d concurrent batches), forward + Viterbi decode"
                % (model, args.reads, args.samples, args.batch, len(groups)))
    return sigs, groups, desc


def plan_resident(group_lens, workload, steps, warmup, nsets, batch):
    """Which batch objects are held in HBM for the device-resident measurement and how often each of them runs.
    `group_lens[k]` = the read lengths of the step's batch k.  Returns (obj_group, reps, mult, nsets, steps, warmup):
    object i is a workspace of batch obj_group[i] that runs reps[i] times in the timed region; the first len(mult)
    objects are one copy of every resident batch, and copy j of them stands for mult[j] batches of the step.  Over the
    timed region every batch of the step is computed exactly `steps` times.

    fixed    `nsets` copies ("buffer sets") of every batch, the steps alternating between them.
    sharded  (config 5) the shard's hundreds of batches stream through at most 16 resident full-size batches plus the
             shard's last partial one: every batch is computed in every step, but only 17 workspaces are held.
    mixed    (config 4) a batch of 130 000-sample reads takes 16 x as long as a batch of 4 000-sample reads with as many
             samples (the scan is serial in time): the nsets * nbatch objects are dealt out in proportion to the batches'
             estimated duration, so that the long batches are in flight as often as a continuously fed basecaller
             would have them."""
    nbatch = len(group_lens)
    res_idx, mult = list(range(nbatch)), [1] * nbatch
    if workload == "sharded":
        nsets, steps, warmup = 1, max(1, min(steps, 2)), 1          # one pass = the whole shard, already many batches
        full = [k for k, g in enumerate(group_lens) if len(g) == batch]
        tail = [k for k, g in enumerate(group_lens) if len(g) != batch]
        keep = full[:16]
        res_idx = keep + tail
        mult = [len(full) // len(keep) + (1 if j < len(full) % len(keep) else 0) for j in range(len(keep))] + [1] * len(tail)
    nres = len(res_idx)
    nsets = max(1, min(nsets, steps))
    obj_group = res_idx * nsets
    set_reps = [steps // nsets + (1 if k < steps % nsets else 0) for k in range(nsets)]
    reps = [set_reps[i // nres] * mult[i % nres] for i in range(len(obj_group))]
    if workload == "mixed":
        est = [2.0 + 0.006 * max(group_lens[k]) / 5.0 for k in res_idx]                      # ms, from the longest read
        copies = [int(min(steps, max(1, round(nsets * nres * e / sum(est))))) for e in est]
        obj_group = list(res_idx) + [k for k, c in zip(res_idx, copies) for _ in range(c - 1)]
        seen, reps = {}, []
        for k in obj_group:
            j = seen.get(k, 0)
            seen[k] = j + 1
            c = copies[res_idx.index(k)]
            reps.append(steps // c + (1 if j < steps % c else 0))
    return obj_group, reps, mult, nsets, steps, warmup


def measure(eng, ranks, args, model, workload, steps, warmup, nsets, detail):
    """One workload through all three measurements: device-resident streaming (`value`), the documented call from
    pageable host arrays (`e2e`), the persistent-batch path (`e2e.persistent`), plus the parity check of the e2e run's
    base strings.  Returns a dict; `detail` adds the per-kernel figures of the headline."""
    import scrappie_b200 as sb
    torch = ranks.torch
    rank = ranks.rank
    sigs, groups, desc = build_groups(args, model, workload, rank, ranks.world)
    nbatch = len(groups)
    obj_group, reps, mult, nsets, steps, warmup = plan_resident([[len(x) for x in g] for g in groups], workload, steps,
                                                               warmup, nsets, args.batch)
    nres = len(mult)                                    # the first nres objects are one copy of every resident batch
    params = sb.default_params()
    total_samples = sum(len(s) for s in sigs)

    # ---- device-resident throughput -------------------------------------------------------------------------------
    # The resident batches run back to back on their own streams with no synchronisation in between (a continuously
    # fed basecaller): step n + 1 starts while step n is still decoding.  A step streams GBs of activations, far more
    # than L2 holds.
    batches, pinned = [], []
    for g in [groups[k] for k in obj_group]:
        b = eng.batch(model, [len(s) for s in g])
        pb = sb.PinnedBuffer(b.total_samples_padded)
        pb.array[:] = 0
        for r, s in enumerate(g):
            pb.array[b.sample_offset[r]:b.sample_offset[r] + len(s)] = s
        b.upload_concat(pb.ptr, pinned_async=False)
        batches.append(b)
        pinned.append(pb)
    total_blocks = sum(b.total_blocks * m for b, m in zip(batches[:nres], mult))
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if nsets > 1 or workload == "sharded":
        sb.multi_stream_time(batches, params, nrep=max(1 if workload == "sharded" else 3, (warmup + nsets - 1) // nsets))   # per object
        ranks.barrier()
        launches0 = eng.launches
        step_ms = sb.multi_stream_time(batches, params, nrep=reps) / steps
        launches = eng.launches - launches0
        ranks.barrier()
        l2_note = ("not flushed: %d steps back to back, each streaming its own %.1f GB of activations through HBM "
                   "(L2 is 126 MB); %d buffer sets alternate" % (steps, total_blocks * 26e3 / 1e9, nsets))
    else:
        sb.multi_time(batches, params, nrep=max(3, warmup), flush_l2=True)
        ranks.barrier()
        launches0 = eng.launches
        ms = sb.multi_time(batches, params, nrep=steps, flush_l2=True)
        launches = eng.launches - launches0
        ranks.barrier()
        step_ms = float(np.mean(ms))
        l2_note = "flushed between timed steps (384 MB overwrite, outside the timed region)"
    clocks = sampler.stop() if sampler else None
    out = {"workload": desc, "l2": l2_note, "buffer_sets": nsets, "steps": steps, "nbatch": nbatch,
           "total_samples": total_samples, "total_blocks": total_blocks, "launches": int(launches), "clocks": clocks,
           "resident_batches": "%d workspaces for the step's %d batches" % (len(obj_group), nbatch)}

    # stage intervals of one synchronised step of the first set (diagnostics, outside the timed region)
    if detail:
        sb.multi_time(batches[:nres], params, nrep=1, flush_l2=True)
        out["stage_concurrent"] = [b.stage_ms() for b in batches[:nres]]

    # ---- end to end, persistent batch objects + pre-filled pinned buffers (round 1's e2e) ---------------------------
    results = [None] * len(batches)

    def work_persistent(i, nrep):
        for _ in range(nrep):
            results[i] = batches[i].basecall(pinned[i].ptr, True, params, lazy=True)

    def run_threads(fn, reps_):
        th = [threading.Thread(target=fn, args=(i, reps_[i])) for i in range(len(reps_))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    run_threads(work_persistent, [1] * len(batches))
    ranks.barrier()
    t0 = time.perf_counter()
    run_threads(work_persistent, reps)
    torch.cuda.synchronize()
    e2e_persistent_s = (time.perf_counter() - t0) / steps
    h2d = sum(b.total_samples_padded * 4 * m for b, m in zip(batches[:nres], mult))
    d2h = sum(d2h_bytes(model, b.nread, max(b.nblock), int(res.nbase.max())) * m for b, res, m in zip(batches[:nres], results[:nres], mult))
    persistent_bases = results[0].bases(0)

    def sustained_and_solo():
        # Per-kernel figures from ONE batch timed alone (CUDA events on its stream, L2 flushed), after the timed regions
        # and BEFORE the sustained run (a kernel timed alone is compared with the burst peaks; after five seconds under
        # the power cap the clocks are 20 % lower).  In the concurrent step the stage intervals of different batches
        # overlap, so they are not launch durations.
        batches[0].time(params, nrep=3, flush_l2=True)
        out["stage_solo"] = batches[0].stage_ms()
        out["batch0"] = {"nread": batches[0].nread, "cols": batches[0].total_blocks, "ostride": batches[0].ostride,
                         "nsamp": batches[0].total_samples_padded}
        # ---- sustained: the streaming loop for >= N seconds (power-capped clocks instead of the burst's).  Runs AFTER
        # the short timed regions, so that those see the clocks a fresh job sees.
        if args.sustained_seconds > 0:
            k_long = int(np.ceil(args.sustained_seconds * 1e3 / (step_ms * steps)))
            n_long = steps * k_long
            smp = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
            ranks.barrier()
            long_ms = sb.multi_stream_time(batches, params, nrep=[r * k_long for r in reps])
            ranks.barrier()
            lc = smp.stop() if smp else None
            long_all = ranks.gather(long_ms)
            out["sustained"] = {"value": ranks.world * total_samples * n_long / (max(long_all) * 1e-3), "unit": "samples/s",
                                "steps": n_long, "seconds": max(long_all) * 1e-3, "ms_per_step": max(long_all) / n_long,
                                "clocks": lc}

    def close_batches():
        for b in batches:
            b.close()
        for pb in pinned:
            pb.close()
        del batches[:], pinned[:]

    results = None
    if not detail:
        close_batches()                                 # the larger workloads need the memory for the pooled workspaces

    # ---- end to end through the documented call: sb2_basecall_batch from pageable host arrays ----------------------
    # A team of C host threads (examples/batch_caller.c -> scrappie_b200/libsb2_caller.so), each taking the next batch
    # -- work items (step, batch), a step's batches longest first -- and making ONE sb2_basecall_batch call on it, exactly
    # what INTEGRATION.md section 2.1 tells a maintainer of `scrappie raw` to do.  No Python inside the timed region.
    order = sorted(range(nbatch), key=lambda k: -sum(len(s) for s in groups[k]))
    job = sb.CallerJob(eng, model, groups, order)
    # calls in flight (each caller sleeps on its batch's completion event): 16 when every call takes the same ~12 ms
    # (12 callers measure 5 % less, 20 the same); the mixed workload's calls take 3 .. 250 ms and want 48 of them
    nworker = min(nbatch * max(nsets, 8), int(os.environ.get("BENCH_MIXED_WORKERS", "48") if workload == "mixed" else os.environ.get("BENCH_E2E_WORKERS", "16")))
    # pool warm-up: every caller must have had a workspace made for it (device buffers, pinned staging, graphs) and the
    # workspaces must have seen the largest batch.  A workspace made before the pool's high-water marks were final is
    # let go when it is handed back, so a caller reaches its steady state with its third call: enough passes for that
    # -- in two runs: when the first one returns every workspace has been handed back, the high-water marks are final
    # and the workspaces made before they were are gone; the second one makes the missing ones at their final size
    job.run(2, nworker, params)
    job.run(max(2, (warmup + 1) // 2, (2 * nworker + nbatch - 1) // nbatch), nworker, params)
    ranks.barrier()
    reallocs0 = eng.reallocs
    # mixed lengths: one call on a batch of 130 000-sample reads takes ~0.2 s, as long as a dozen steps -- three times
    # the steps, so that filling and draining the callers' pipeline is not most of what is timed
    e2e_steps = 3 * steps if workload == "mixed" else steps
    e2e_total_s, nbases_doc, doc_bases, _ = job.run(e2e_steps, nworker, params, want_bases=(rank == 0))
    torch.cuda.synchronize()
    e2e_s = e2e_total_s / e2e_steps
    reallocs_timed = eng.reallocs - reallocs0           # 0: the documented call allocates nothing in steady state
    ranks.barrier()

    # ---- parity gate on the base strings the timed run produced -----------------------------------------------------
    starts = np.concatenate([[0], np.cumsum([len(g) for g in groups])])
    if workload == "fixed":
        pick = [(0, r) for r in range(len(groups[0]))] + [(k, r) for k in range(1, nbatch) for r in range(0, len(groups[k]), 8)]
    else:
        rng = np.random.default_rng(17)
        flat = [(k, r) for k in range(nbatch) for r in range(len(groups[k]))]
        pick = [flat[i] for i in sorted(rng.choice(len(flat), size=min(48, len(flat)), replace=False))]
    if rank == 0:
        got = [doc_bases[int(starts[k]) + r] for k, r in pick]
        parity = check_parity(model, [groups[k][r] for k, r in pick], got, eng)
        parity["persistent_path_agrees"] = (persistent_bases == doc_bases[0])
        out["parity"] = parity
    doc_bases = None
    job = None
    if detail:
        sustained_and_solo()
        close_batches()
    eng.trim_pool()                                     # the next workload sizes its own workspaces
    step_all = ranks.gather(step_ms)
    e2e_all = ranks.gather(e2e_s)
    e2ep_all = ranks.gather(e2e_persistent_s)
    W = ranks.world
    out.update({
        "ms_per_step": max(step_all), "value": W * total_samples / (max(step_all) * 1e-3),
        "ms_per_step_ranks": spread(step_all),
        "blocks_per_s": W * total_blocks / (max(step_all) * 1e-3),
        "kbases_per_s": W * nbases_doc / max(e2e_all) / 1e3,
        "e2e": {"value": W * total_samples / max(e2e_all), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": max(e2e_all) * 1e3, "ms_per_step_ranks": spread([x * 1e3 for x in e2e_all]),
                "api": "sb2_basecall_batch (pooled workspaces) from pageable host arrays, %d C host threads (examples/batch_caller.c)" % nworker,
                "workspace_allocations_in_timed_region": reallocs_timed, "steps": e2e_steps,
                "persistent": {"value": W * total_samples / max(e2ep_all), "ms_per_step": max(e2ep_all) * 1e3,
                               "api": "sb2_batch_basecall on caller-owned batches, pre-filled pinned buffers"}},
    })
    return out


def write_peak_gbs(torch):
    """Write-only HBM bandwidth measured now (fill of a 4 GiB buffer, best of 5): kernels that mostly write are bounded
    by it rather than by the half-read half-write copy figure of MEASURED_PEAKS.json."""
    x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x.fill_(1.0)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del x
    torch.cuda.empty_cache()
    return best


def main_b200(args, rank, world, local_rank):
    import scrappie_b200 as sb
    from scrappie_b200.sharding import broadcast_blob

    ranks = Ranks(rank, world, local_rank)
    torch = ranks.torch
    torch.cuda.set_device(local_rank)
    eng = sb.Engine(local_rank)
    default_run = (args.workload == "fixed" and args.model == "rgrgr_r94" and not args.no_other_configs)
    models = [args.model] + (["rnnrf_r94"] if default_run else [])
    # weights: rank 0 reads the blobs, every other rank receives them over NCCL (init only)
    for m in models:
        eng.load_blob(m, broadcast_blob(os.path.join(sb.WEIGHTS_DIR, m + ".bin"), rank, ranks.dist,
                                        device="cuda" if world > 1 else "cpu"))

    head = measure(eng, ranks, args, args.model, args.workload, args.steps, args.warmup, args.sets, detail=True)

    others = {}
    if default_run:
        def brief(m):
            keys = ("workload", "value", "ms_per_step", "ms_per_step_ranks", "kbases_per_s", "steps", "buffer_sets", "parity", "launches", "resident_batches")
            d = {k: m[k] for k in keys if k in m}
            d["e2e"] = {k: m["e2e"][k] for k in ("value", "ms_per_step", "steps", "api", "workspace_allocations_in_timed_region",
                                                 "h2d_bytes_per_step", "d2h_bytes_per_step")}
            d["e2e"]["persistent"] = m["e2e"]["persistent"]["value"]
            d["unit"] = "samples/s"
            return d
        others["rnnrf_r94 (config 3)"] = brief(measure(eng, ranks, args, "rnnrf_r94", "fixed", 8, 4, args.sets, detail=False))
        others["mixed lengths (config 4)"] = brief(measure(eng, ranks, args, "rgrgr_r94", "mixed", 12, 6, 6, detail=False))
        if world > 1:
            others["sharded 100k (config 5)"] = brief(measure(eng, ranks, args, "rgrgr_r94", "sharded", 2, 1, 1, detail=False))

    if rank != 0:
        ranks.close()
        return

    pk, pk_src = peaks()
    hbm = pk["hbm_gbs"]
    solo, b0 = head["stage_solo"], head["batch0"]
    model = args.model
    H = 112 if model == "rnnrf_r94" else 96
    cols, nstate_stride, nsamp0 = b0["cols"], b0["ostride"], b0["nsamp"]
    scan_avg_ms = float(np.mean([solo["scan%d" % l] for l in range(1, 6)]))
    flop_per_launch = FLOP_PER_READ_STEP_RECURRENT[model] * cols
    achieved = flop_per_launch / (scan_avg_ms * 1e-3) / 1e12
    # reads per scan CTA: groups of 8 reads (4 groups at H = 96, 3 at H = 112) for batches of >= 48 reads, else 2 x 4
    big = b0["nread"] >= 48
    reads_per_cta = (24 if H == 112 else 32) if big else 8
    scan_ctas = (b0["nread"] + reads_per_cta - 1) // reads_per_cta
    peak = pk["bf16_tflops"]
    traffic = None
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json")) \
        if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    measured = json.load(open(os.path.join(ROOT, "profiles", tfiles[-1]))) if tfiles else {}
    for k, v in measured.items():
        if k.startswith("gru_scan") and model != "rnnrf_r94":
            traffic = v["dram_read_bytes"] + v["dram_write_bytes"]
    wpeak = write_peak_gbs(torch)

    def hbm_kernel(name, ms, nbytes, wbytes=None):
        k = {"kernel": name, "bound": "hbm", "avg_launch_ms": ms, "bytes_per_launch": nbytes,
             "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm}
        if wbytes is not None:
            k["write_bytes_per_launch"] = wbytes
            k["write_achieved"] = wbytes / (ms * 1e-3) / 1e9
            k["write_peak"] = wpeak
            k["write_peak_source"] = "fill of a 4 GiB buffer measured in this run (not a driver-written peak)"
            k["write_frac"] = k["write_achieved"] / wpeak
        return k
    aff_ms = float(np.mean([solo["affine%d" % l] for l in range(1, 6)]))
    tb_bytes = (nstate_stride - 4 + 4) if model != "rnnrf_r94" else 8
    kernels = [
        hbm_kernel("conv_act", solo["conv"], nsamp0 * 4 + cols * H * 4),
        hbm_kernel("affine_tc (GRU input transform)", aff_ms, cols * (H + 3 * H) * 4, cols * 3 * H * 4),
        hbm_kernel("head (FF + softmax + robust log)", solo["head_gemm"] + solo["head_finish"], cols * (H + nstate_stride) * 4,
                   cols * nstate_stride * 4),
        hbm_kernel("decode (Viterbi + traceback)", solo["decode"], cols * (nstate_stride * 4 + tb_bytes + 4)),
    ]
    stage = head["stage_concurrent"]
    stage_sum = {k: float(np.mean([st[k] for st in stage])) for k in stage[0]}
    step_ms = head["ms_per_step"]
    bpb = BYTES_PER_BLOCK.get(model, 0)
    line = {
        "metric": "raw samples/sec (%s)" % model, "value": head["value"],
        "unit": "samples/s", "n_gpus": world, "steps": head["steps"], "warmup": args.warmup, "ms_per_step": step_ms,
        "ms_per_step_ranks": head["ms_per_step_ranks"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": head["workload"], "l2": head["l2"], "buffer_sets": head["buffer_sets"],
                   "scan_impl": os.environ.get("SCRAPPIE_B200_SCAN", "default"),
                   "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                   "scan_groups": "%d reads per CTA, %d per group" % (reads_per_cta, 8 if big else 4),
                   "parallelism": "reads sharded, %d rank(s), NCCL weight broadcast at init only" % world},
        "kbases_per_s": head["kbases_per_s"],
        "blocks_per_s": head["blocks_per_s"],
        "network_tflops": FLOP_PER_BLOCK_TOTAL.get(model, 0) * head["blocks_per_s"] / 1e12,
        "e2e": head["e2e"],
        "parity": head.get("parity"),
        "sustained": head.get("sustained"),
        # the whole step against the HBM roofline: algorithmic bytes if every intermediate is written and read once
        # (conv out, X / Xin per layer, posterior, traceback) over the measured step time
        "step_hbm": {"bytes_per_block": bpb, "achieved": bpb * head["blocks_per_s"] / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": bpb * head["blocks_per_s"] / 1e9 / hbm / world},
        "gpu_launches": head["launches"],
        "clocks": head["clocks"],
        "roofline": {"kernel": "gru_scan (recurrent sW/sW2 products + gates): 5 of the 13 launches per batch, "
                               "the largest share of the step's SM time (profiles/*_summary.md)",
                     "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "%s bf16 dense burst (MEASURED_PEAKS.json), kernel timed alone; the scan is a chain of "
                                    "dependent 12-instruction UMMA groups, latency- not throughput-bound (DESIGN.md section 4)" % pk_src,
                     "flop_per_launch": flop_per_launch, "avg_launch_ms": scan_avg_ms,
                     # a scan launch of one batch occupies one SM per `reads_per_cta` reads, not the GPU: the
                     # other SMs run the other batches' kernels at the same time
                     "reads_per_cta": reads_per_cta,
                     "sms_used_per_launch": scan_ctas, "frac_of_sm_share": achieved / (peak * min(scan_ctas, 148) / 148.0),
                     "stage_ms_solo_batch": solo, "stage_ms_per_batch_concurrent": stage_sum,
                     "other_kernels": kernels},
    }
    if others:
        line["other_configs"] = others
    if world == 1 and not args.no_cpu_baseline:
        sigs = make_workload(min(args.reads, 1024), args.samples, 1000)
        line["cpu_baseline"] = cpu_baseline(model, sigs, args.cpu_sample_reads)
    print(json.dumps(line))
    ranks.close()
    bad = [name for name, p in [("headline", line.get("parity"))] + [(k, v.get("parity")) for k, v in others.items()]
           if p is not None and not p["ok"]]
    if bad:
        sys.stderr.write("bench.py: base sequences differ from the reference (not a near-tie within the posterior tolerance) in: %s\n" % ", ".join(bad))
        sys.exit(3)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
