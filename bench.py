#!/usr/bin/env python
"""Benchmark of the raw basecalling hot path (BASELINE.json: raw samples/s, rgrgr_r94).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

One "step" = one pass of the hot path (network forward + Viterbi decode) over the whole
workload on each GPU: `--reads` synthetic reads of `--samples` samples, processed in batches
of `--batch` reads that run concurrently on their own CUDA streams (BASELINE config 2:
rgrgr_r94, 1024 x 4000-sample reads, batch 256).  With N > 1 (torchrun, one rank per GPU) every
rank processes its own `--reads` reads (weak scaling); rank 0 loads the weight blob and
broadcasts it over NCCL at start-up; there is no collective on the per-read path.

  value  samples/s with the signals already resident in HBM; device time from CUDA events on
         the launching stream (L2 flushed between steps, outside the timed region), max over ranks.
  e2e    samples/s through the C-ABI batch basecall with HOST buffers: pinned H2D of the
         signals, forward, decode, D2H of paths/scores, homopolymer fix-up and overlapper on
         the host, all inside the timed region (wall clock, max over ranks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

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
    ap.add_argument("--sets", type=int, default=4,
                    help="buffer sets: consecutive steps alternate between the sets and are not synchronised with each "
                         "other, so step n + 1 overlaps the tail of step n (1 = one set, L2 flushed between steps)")
    ap.add_argument("--workload", default="fixed", choices=["fixed", "mixed"],
                    help="fixed: --reads reads of --samples samples (BASELINE config 2); mixed: --reads reads with "
                         "log-normal lengths in [1k, 200k] samples, length-bucketed dynamic batching (config 4)")
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(nreads, nsamples, seed0):
    from scrappie_b200.synthetic import synthetic_read
    return [synthetic_read(seed0 + i, nsamples) for i in range(nreads)]


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------

def run_reference(model, sigs, nthreads=0):
    """Times oracle/_ref (the reference's own C sources + OpenBLAS, one read per OpenMP thread).
    Returns (seconds, nbases, nblocks, threads)."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libref_bench.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    L.ref_bench_run.restype = C.c_double
    L.ref_bench_run.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p]
    concat = np.concatenate(sigs).astype(np.float32)
    lens = np.array([len(s) for s in sigs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    nb, nk = C.c_size_t(0), C.c_size_t(0)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the
    # reference's own recommendation -- one read per core -- means; README.md:66-71)
    threads = nthreads or len(os.sched_getaffinity(0))
    secs = L.ref_bench_run(model.encode(), concat.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(sigs),
                           threads, C.byref(nb), C.byref(nk), None)
    return secs, nb.value, nk.value, threads


def run_oracle_port(model, sigs):
    """Fallback when oracle/_ref is absent: the scalar C restatement, 1 thread."""
    from oracle.oracle import Oracle
    o = Oracle()
    t0 = time.time()
    nb = 0
    for s in sigs:
        _, _, bases, _ = o.basecall_raw(model, s)
        nb += len(bases or "")
    return time.time() - t0, nb, 0, 1


def cpu_baseline(model, sigs, nreads_sample):
    sample = (sigs * ((nreads_sample + len(sigs) - 1) // len(sigs)))[:nreads_sample]   # the workload's reads, repeated
    res = run_reference(model, sample)
    kind = "reference"
    if res is None:
        sample = sigs[:max(8, nreads_sample // 16)]
        res = run_oracle_port(model, sample)
        kind = "port"
    secs, nbases, _, threads = res
    nsamp = sum(len(s) for s in sample)
    return {"value": nsamp / secs, "unit": "samples/s", "cores": threads, "kind": kind,
            "kbases_per_s": nbases / secs / 1e3,
            "sample": "%d of the workload's reads (%d samples), one read per OpenMP thread, 1 BLAS thread, %.2f s wall"
                      % (len(sample), nsamp, secs)}


def main_reference(args, rank, world):
    if rank != 0:
        return
    # one step = one pass over the same workload as the GPU arm (args.reads reads), capped so that K steps stay bounded
    nstep_reads = min(args.reads, args.cpu_sample_reads)
    sigs = make_workload(nstep_reads, args.samples, 1000)
    for _ in range(max(1, min(args.warmup, 1))):
        run_reference(args.model, sigs[:32]) if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_bench.so")) else None
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
    line = {"impl": "reference", "metric": "raw samples/sec (%s)" % args.model, "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s raw, synthetic %d-sample reads (bounded sample: %d reads per step), CPU %s"
                                   % (args.model, args.samples, len(sigs), kind)},
            "kbases_per_s": nbases / (ms / 1e3) / 1e3,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": kind,
                             "sample": "%d reads x %d samples per step" % (len(sigs), args.samples)},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------

def main_b200(args, rank, world, local_rank):
    import torch
    import scrappie_b200 as sb

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = sb.Engine(local_rank)
    # weights: rank 0 reads the blob, every other rank receives it over NCCL (init only)
    from scrappie_b200.sharding import broadcast_blob, max_over_ranks
    blob_path = os.path.join(sb.WEIGHTS_DIR, args.model + ".bin")
    eng.load_blob(args.model, broadcast_blob(blob_path, rank, dist, device="cuda" if world > 1 else "cpu"))

    if args.workload == "mixed":
        from scrappie_b200.sharding import lognormal_lengths, plan_batches
        from scrappie_b200.synthetic import synthetic_read
        lens = lognormal_lengths(args.reads, seed=4 + rank)
        sigs = [synthetic_read(1000 + rank * args.reads + i, int(n)) for i, n in enumerate(lens)]
        plan = plan_batches(lens, max_reads=args.batch, max_samples=args.batch * 4096)
        groups = [[sigs[i] for i in idx] for idx in plan]
        nbatch = len(groups)
    else:
        sigs = make_workload(args.reads, args.samples, 1000 + rank * args.reads)
        nbatch = (args.reads + args.batch - 1) // args.batch
        groups = [sigs[i * args.batch:(i + 1) * args.batch] for i in range(nbatch)]
    nsets = max(1, min(args.sets, args.steps))
    groups = groups * nsets                              # set k = batches[k * nbatch : (k + 1) * nbatch], same reads
    batches = [eng.batch(args.model, [len(s) for s in g]) for g in groups]
    pinned = []
    for b, g in zip(batches, groups):
        pb = sb.PinnedBuffer(b.total_samples_padded)
        pb.array[:] = 0
        for r, s in enumerate(g):
            pb.array[b.sample_offset[r]:b.sample_offset[r] + len(s)] = s
        pinned.append(pb)
        b.upload_concat(pb.ptr, pinned_async=False)
    params = sb.default_params()
    total_samples = sum(len(s) for s in sigs)
    total_blocks = sum(b.total_blocks for b in batches[:nbatch])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident throughput ------------------------------------------------
    # One step = one pass over the rank's `--reads` reads (nbatch concurrent batches).  With two buffer sets the
    # steps alternate between the sets and run back to back on the batches' own streams with no synchronisation
    # in between (a continuously fed basecaller): step n + 1 starts while step n is still decoding.  The working
    # set of a step (GBs of activations) is far larger than L2, so nothing is served from cache across steps.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if nsets > 1:
        # set k runs steps k, k + nsets, ...: exactly args.steps steps in all
        set_reps = [args.steps // nsets + (1 if k < args.steps % nsets else 0) for k in range(nsets)]
        reps = [set_reps[i // nbatch] for i in range(len(batches))]
        sb.multi_stream_time(batches, params, nrep=max(3, (args.warmup + nsets - 1) // nsets))
        barrier()
        launches0 = eng.launches
        step_ms = sb.multi_stream_time(batches, params, nrep=reps) / args.steps
        launches = eng.launches - launches0
        barrier()
        l2_note = ("not flushed: %d steps back to back, each streaming its own %.1f GB of activations through HBM "
                   "(L2 is 126 MB); %d buffer sets alternate" %
                   (args.steps, sum(b.total_blocks for b in batches[:nbatch]) * 26e3 / 1e9, nsets))
    else:
        sb.multi_time(batches, params, nrep=max(3, args.warmup), flush_l2=True)
        barrier()
        launches0 = eng.launches
        ms = sb.multi_time(batches, params, nrep=args.steps, flush_l2=True)
        launches = eng.launches - launches0
        barrier()
        step_ms = float(np.mean(ms))
        l2_note = "flushed between timed steps (384 MB overwrite, outside the timed region)"
    clocks = sampler.stop() if sampler else None
    # stage intervals of one synchronised step of the first set (diagnostics, outside the timed region)
    sb.multi_time(batches[:nbatch], params, nrep=1, flush_l2=True)
    stage = [b.stage_ms() for b in batches[:nbatch]]

    # ---- end to end through the C-ABI batch basecall, host buffers ------------------------
    # One host thread per batch object; each basecalls its batch steps / nsets times (pinned host signal in, base
    # strings out, every time), the threads free-running like the streams above.
    results = [None] * len(batches)

    def work(i, nrep):
        for _ in range(nrep):
            results[i] = batches[i].basecall(pinned[i].ptr, True, params, lazy=True)

    def e2e_run(reps_):
        th = [threading.Thread(target=work, args=(i, reps_[i])) for i in range(len(batches))]
        for t in th:
            t.start()
        for t in th:
            t.join()

    e2e_reps = reps if nsets > 1 else [args.steps] * len(batches)
    e2e_run([1] * len(batches))
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_reps)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    nbases = int(sum(int(res.nbase.sum()) for res in results[:nbatch]))
    h2d = sum(b.total_samples_padded * 4 for b in batches[:nbatch])
    def d2h_bytes(b, res):
        # finish_on_device (csrc/engine.cu): base count + score per read, then either the whole base-string area
        # (batches whose area is <= 2 MB) or a 2-D copy as wide as the longest call
        klen = 1 if args.model == "rnnrf_r94" else (6 if args.model == "rgrgr_r10" else 5)
        stride = (klen * (max(b.nblock) + 1) + 1 + 15) // 16 * 16
        area = b.nread * stride
        return b.nread * 8 + (area if area <= (2 << 20) else b.nread * ((int(res.nbase.max()) + 1 + 15) // 16 * 16))
    d2h = sum(d2h_bytes(b, res) for b, res in zip(batches[:nbatch], results[:nbatch]))

    # ---- reduce over ranks (max time) ------------------------------------------------------
    step_ms, e2e_s = max_over_ranks([step_ms, e2e_s], dist, device="cuda" if world > 1 else "cpu")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    # Per-kernel figures from ONE batch timed alone (CUDA events on its stream, L2 flushed), after the timed region:
    # in the concurrent step the stage intervals of different batches overlap, so they are not launch durations.
    _, _, _ = batches[0].time(params, nrep=3, flush_l2=True)
    solo = batches[0].stage_ms()
    H = 112 if args.model == "rnnrf_r94" else 96
    nstate_stride = batches[0].ostride
    cols = batches[0].total_blocks
    nsamp0 = batches[0].total_samples_padded
    scan_avg_ms = float(np.mean([solo["scan%d" % l] for l in range(1, 6)]))
    flop_per_launch = FLOP_PER_READ_STEP_RECURRENT[args.model] * cols
    achieved = flop_per_launch / (scan_avg_ms * 1e-3) / 1e12
    grp_env = os.environ.get("SCRAPPIE_B200_SCAN_GROUPS", "0")
    if args.model == "rnnrf_r94":
        reads_per_cta = 12 if (batches[0].nread >= 96 and grp_env in ("0", "3")) else 8
    else:
        reads_per_cta = 16 if (batches[0].nread >= 128 and grp_env in ("0", "4")) else 8
    scan_ctas = (batches[0].nread + reads_per_cta - 1) // reads_per_cta
    peak = pk["bf16_tflops"]
    hbm = pk["hbm_gbs"]
    traffic = None
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json")) \
        if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    measured = json.load(open(os.path.join(ROOT, "profiles", tfiles[-1]))) if tfiles else {}
    for k, v in measured.items():
        if k.startswith("gru_scan") and args.model != "rnnrf_r94":
            traffic = v["dram_read_bytes"] + v["dram_write_bytes"]

    # write-only HBM bandwidth of this pool's B200 (tools/hbm_write_probe.py, r29): kernels that mostly write are
    # bounded by it rather than by the half-read half-write copy figure of MEASURED_PEAKS.json
    HBM_WRITE_GBS = 3904.0

    def hbm_kernel(name, ms, nbytes, wbytes=None):
        k = {"kernel": name, "bound": "hbm", "avg_launch_ms": ms, "bytes_per_launch": nbytes,
             "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / hbm}
        if wbytes is not None:
            k["write_bytes_per_launch"] = wbytes
            k["write_achieved"] = wbytes / (ms * 1e-3) / 1e9
            k["write_peak"] = HBM_WRITE_GBS
            k["write_frac"] = k["write_achieved"] / HBM_WRITE_GBS
        return k
    aff_ms = float(np.mean([solo["affine%d" % l] for l in range(1, 6)]))
    tb_bytes = (nstate_stride - 4 + 4) if args.model != "rnnrf_r94" else 8
    kernels = [
        hbm_kernel("conv_act", solo["conv"], nsamp0 * 4 + cols * H * 4),
        hbm_kernel("affine_tc (GRU input transform)", aff_ms, cols * (H + 3 * H) * 4, cols * 3 * H * 4),
        hbm_kernel("head (FF + softmax + robust log)", solo["head_gemm"] + solo["head_finish"], cols * (H + nstate_stride) * 4,
                   cols * nstate_stride * 4),
        hbm_kernel("decode (Viterbi + traceback)", solo["decode"], cols * (nstate_stride * 4 + tb_bytes + 4)),
    ]
    stage_sum = {k: float(np.mean([st[k] for st in stage])) for k in stage[0]}
    line = {
        "metric": "raw samples/sec (%s)" % args.model, "value": world * total_samples / (step_ms * 1e-3),
        "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("%s raw, %d synthetic %d-sample reads per GPU, batch=%d (%d concurrent batches), "
                                "forward + Viterbi decode" % (args.model, args.reads, args.samples, args.batch, nbatch))
                               if args.workload == "fixed" else
                               ("%s raw, %d synthetic reads per GPU with log-normal lengths (median 8000, sigma 1.0, clipped "
                                "to [1k, 200k] samples; %d samples in total, longest %d), length-bucketed dynamic batching "
                                "into %d concurrent batches of <= %d reads" % (args.model, args.reads, total_samples,
                                                                              max(len(x) for x in sigs), nbatch, args.batch)),
                   "l2": l2_note, "buffer_sets": nsets,
                   "scan_impl": os.environ.get("SCRAPPIE_B200_SCAN", "default"),
                   "parallelism": "reads sharded, %d rank(s), NCCL weight broadcast at init only" % world},
        "kbases_per_s": world * nbases / e2e_s / 1e3,
        "blocks_per_s": world * total_blocks / (step_ms * 1e-3),
        "network_tflops": world * FLOP_PER_BLOCK_TOTAL.get(args.model, 0) * total_blocks / (step_ms * 1e-3) / 1e12,
        "e2e": {"value": world * total_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
        # the whole step against the HBM roofline: algorithmic bytes if every intermediate is written and read once
        # (conv out, X / Xin per layer, posterior, traceback) over the measured step time
        "step_hbm": {"bytes_per_block": BYTES_PER_BLOCK.get(args.model), "achieved": (BYTES_PER_BLOCK.get(args.model, 0) * total_blocks
                     / (step_ms * 1e-3) / 1e9), "peak": hbm, "unit": "GB/s",
                     "frac": BYTES_PER_BLOCK.get(args.model, 0) * total_blocks / (step_ms * 1e-3) / 1e9 / hbm},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gru_scan (recurrent sW/sW2 products + gates): 5 of the 13 launches per batch, "
                               "the largest share of the step (profiles/*_summary.md)",
                     "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "%s bf16 dense burst (MEASURED_PEAKS.json), kernel timed alone; the scan is a chain of "
                                    "dependent 12-instruction UMMA groups, latency- not throughput-bound (DESIGN.md section 4)" % pk_src,
                     "flop_per_launch": flop_per_launch, "avg_launch_ms": scan_avg_ms,
                     # a scan launch of one batch occupies one SM per 16 (H = 96) or 8 reads, not the GPU: the
                     # other SMs run the other batches' kernels at the same time
                     "sms_used_per_launch": scan_ctas, "frac_of_sm_share": achieved / (peak * min(scan_ctas, 148) / 148.0),
                     "stage_ms_solo_batch": solo, "stage_ms_per_batch_concurrent": stage_sum,
                     "other_kernels": kernels},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.model, sigs, args.cpu_sample_reads)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
