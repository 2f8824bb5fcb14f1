#!/usr/bin/env python
"""bench.py — KITTI-shaped clouds/sec through subsample + radius search + KPConv KFE encoder (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (B200, libaprb200.so)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path on this box's host cores

A step = one synthetic KITTI-shaped PAIR (2 clouds, first_subsampling_dl = 0.3; BASELINE.json configs[1]) through
the whole hot path: 3 grid subsamplings + 10 radius searches (the pyramid of datasets/dataloader.py:93-176) and the
11 KFE encoder blocks (models/architectures.py:149-153), random-init weights, features = ones.
`value`  : device-resident (level-0 points already in HBM), CUDA-event time per step, L2 flushed between steps.
`e2e`    : same metric from HOST buffers: pinned H2D of the points, the pipeline, D2H of the encoder output.
N > 1    : one process per GPU (torchrun), independent pairs per rank, no collective on the data path (weak scaling).
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

import torch  # noqa: E402

from apr_b200 import synth  # noqa: E402
from apr_b200.config import kitti_config, nuscenes_config  # noqa: E402

METRIC = "KITTI-shaped clouds/sec (subsample+radius+KPConv KFE)"
WORKLOADS = {   # --workload: (synthetic generator kind, distant pairs, config.workload name)      BASELINE.json configs
    "kitti": ("kitti", False, "kitti_pair_kfe_encoder"),          # configs[1] (the metric's configuration)
    "lokitti": ("kitti", True, "lokitti_distant_pair_kfe_encoder"),   # configs[2] pair geometry (pose distance 5-50 m)
    "nuscenes": ("nusc", False, "nuscenes_pair_kfe_encoder"),     # configs[3]
}
UNIT = "clouds/s"
LIMITS_FALLBACK = [57, 56, 57, 55]       # SURVEY.md §6 (80th percentile), used only if calibration is skipped


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops_sustained"], src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback (B200_PROFILING.md)")


def raw_pairs(n_pairs, seed0):
    return [synth.pair_raw(seed0 + i, "kitti") for i in range(n_pairs)]


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Rows sampled before this call (warm-up) are ignored."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.rows = self.rows[self.first:] if len(self.rows) > self.first else self.rows
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_pyramid(ref, cfg, limits, p0, l0):
    """collate_fn_descriptor for one pair through the reference's cpp_wrappers object code (oracle/_ref)."""
    from oracle.ref import collate_ref
    pyr = collate_ref(p0, l0, cfg, limits, ref.subsample_batch, ref.batch_query)
    return dict(points=[torch.from_numpy(p) for p in pyr["points"]],
                neighbors=[torch.from_numpy(n).long() for n in pyr["neighbors"]],   # dataloader.py:164-166
                pools=[torch.from_numpy(n).long() for n in pyr["pools"]],
                features=torch.ones(len(p0), 1))


def reference_encoder(batch, cfg, sd):
    """fp32 torch-CPU restatement of the KFE encoder blocks (the reference's models/ cannot travel to this box)."""
    from oracle import blocks_ref
    with torch.no_grad():
        return blocks_ref.encoder_ref(batch, sd, cfg)


def reference_step(ref, oracle_mod, cfg, sd, limits, p0, l0):
    return reference_encoder(reference_pyramid(ref, cfg, limits, p0, l0), cfg, sd)


def reference_run(ref, cfg, sd, limits, pairs, n_steps, budget_s):
    """The reference's CPU path over n_steps pairs with every host core: the pyramids are built pair-parallel (one
    single-threaded cpp_wrappers call chain per pair, as DataLoader(num_workers) runs collate_fn_descriptor:
    datasets/dataloader.py:252-260; the ctypes calls release the GIL), then the encoder runs pair by pair on all
    threads. Returns (pairs done, seconds). Stops early once budget_s is exceeded (at least one pair)."""
    from concurrent.futures import ThreadPoolExecutor
    cores = torch.get_num_threads()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, min(cores, n_steps))) as ex:
        batches = list(ex.map(lambda i: reference_pyramid(ref, cfg, limits, *pairs[i % len(pairs)]), range(n_steps)))
    t_pyr = time.perf_counter() - t0
    done = 0
    for b in batches:
        reference_encoder(b, cfg, sd)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if done < n_steps:                                        # only `done` pyramids were needed: charge their share
        dt -= t_pyr * (1.0 - done / n_steps)
    return done, dt


CALIB_PAIRS = 10     # both arms calibrate the neighbourhood limits on the same first pairs of the synthetic set


def reference_limits(ref, cfg, pairs):
    """calibrate_neighbors (datasets/dataloader.py:200-232) restated on the reference's own search: the limits the
    reference would use on this data — the same pairs our arm calibrates on, so both arms run the same widths."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.ref import calibrate_ref
    return [int(v) for v in calibrate_ref(pairs, cfg, ref.subsample_batch, ref.batch_query, samples_threshold=10 ** 9)]


def reference_gpu_eager_run(ref, cfg, sd, limits, pairs, dev, budget_s):
    """The reference's actual deployment shape (lib/tester.py:49-57, datasets/dataloader.py:252-260): pyramids on the HOST by
    the cpp_wrappers object code (pair-parallel, like DataLoader workers), then the KFE encoder as eager PyTorch ops on the
    GPU (oracle/blocks_ref.py, the fp32 restatement of models/blocks.py, fed CUDA tensors: int64 indices shipped H2D like
    lib/trainer.py:299-305 does). Returns (pairs, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import blocks_ref
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    def to_dev(b):
        return {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in b.items()}
    with torch.no_grad():
        blocks_ref.encoder_ref(to_dev(reference_pyramid(ref, cfg, limits, *pairs[0])), sd_dev, cfg)   # warm-up (cuBLAS, allocator)
        torch.cuda.synchronize(dev)
        cores = torch.get_num_threads()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=max(1, min(cores, len(pairs)))) as ex:
            batches = list(ex.map(lambda pr: reference_pyramid(ref, cfg, limits, *pr), pairs))
        t_pyr = time.perf_counter() - t0
        done = 0
        for b in batches:
            y = blocks_ref.encoder_ref(to_dev(b), sd_dev, cfg).cpu()
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    if done < len(pairs):
        dt -= t_pyr * (1.0 - done / len(pairs))
    return done, dt, t_pyr * done / len(pairs)


def reference_setup(cfg, n_pairs, seed0, kind_name="kitti", distant=False):
    from oracle.ref import Oracle, RefL1
    ref = RefL1() if RefL1.available() else Oracle()
    # the pyramid half is the reference's own object code when oracle/_ref is built; the encoder half is ALWAYS the
    # builder's fp32 torch restatement (oracle/blocks_ref.py): models/blocks.py cannot travel to the GPU box
    kind = "reference+port" if RefL1.available() else "port"
    pairs = []
    for i in range(n_pairs):
        a, b = synth.pair_raw(seed0 + i, kind_name, distant)
        raw = np.concatenate([a, b]); lens = np.array([len(a), len(b)], np.int32)
        pairs.append(ref.subsample_batch(raw, lens, sampleDl=cfg.first_subsampling_dl))   # first-level voxelisation
    from apr_b200.architectures import KPFCNNEncoder
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg)
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    return ref, kind, pairs, sd


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0                                            # rank 0 alone runs the CPU arm
    cfg = kitti_config()
    try:                                                    # torchrun exports OMP_NUM_THREADS=1: use every host core
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    ref, kind, pairs, sd = reference_setup(cfg, CALIB_PAIRS, 0)
    cores = torch.get_num_threads()
    limits = reference_limits(ref, cfg, pairs) if not args.no_calibrate else LIMITS_FALLBACK   # same pairs as our arm
    warm = min(args.warmup, 1)
    for i in range(warm):
        reference_step(ref, None, cfg, sd, limits, *pairs[i % len(pairs)])
    done, dt = reference_run(ref, cfg, sd, limits, pairs, args.steps, 420.0)   # keep the whole run within a few minutes
    value = 2.0 * done / dt
    sample = (f"{done} step(s) x 1 KITTI-shaped pair; pyramids = "
              f"{'reference object code oracle/_ref (nanoflann)' if kind != 'port' else 'oracle C port'}, built "
              f"pair-parallel on up to {cores} threads; encoder = fp32 torch-CPU restatement oracle/blocks_ref.py on {cores} threads; "
              f"limits calibrated (calibrate_neighbors on the reference search) over the same {CALIB_PAIRS} pairs as the GPU arm")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": warm, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "kitti_pair_kfe_encoder", "pairs_per_step": 1, "points_stacked": int(len(pairs[0][0])),
                       "limits": limits},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ our arm
def algorithmic_work(trace):
    """Per-kernel algorithmic bytes / flops for ONE step from the op trace."""
    w = {"kpconv_flops": 0.0, "kp_weighted_bytes": 0.0, "nb_query_bytes": 0.0, "subsample_bytes": 0.0}
    n = {"kpconv": 0, "nb": 0}
    for rec in trace:
        if rec[0] == "kpconv":
            _, nq, ns, h, k, cin, cout = rec
            w["kpconv_flops"] += 2.0 * nq * k * cin * cout
            w["kp_weighted_bytes"] += 4.0 * (ns * cin + nq * k * cin) + 4.0 * nq * h + 12.0 * (nq + ns)
            n["kpconv"] += 1
        elif rec[0] == "nb":
            _, nq, ns, width = rec
            w["nb_query_bytes"] += 12.0 * nq + 16.0 * ns + 4.0 * nq * width
            n["nb"] += 1
        elif rec[0] == "sub":
            _, nin, m = rec
            w["subsample_bytes"] += 12.0 * nin + 12.0 * m
    return w, n


def encoder_shapes(enc, n_levels):
    """Op trace [(kind, ...)] of one pass derived from the block list and the per-level point counts."""
    from apr_b200 import blocks
    tr = []
    for m in enc.encoder_blocks:
        l = m.layer_ind
        strided = 'strided' in m.block_name
        nq, ns = (n_levels[l + 1] if strided else n_levels[l]), n_levels[l]
        c = m.KPConv
        if isinstance(m, blocks.ResnetBottleneckBlock):
            # 5th field: which kernel runs it on the native path — "gemm" = persistent tcgen05 GEMM (product stored),
            # "nrm" = gemm_nrm_f16_kernel (statistics pass + recompute; pipeline.cu: K_main + K_shortcut <= 1024)
            has_sc = isinstance(m.unary_shortcut, blocks.UnaryBlock)
            k_total = m.out_dim // 4 + (m.in_dim if has_sc else 0)
            closing = "nrm" if k_total <= 1024 else "gemm"
            if isinstance(m.unary1, blocks.UnaryBlock):
                tr.append(("linear", ns, m.in_dim, m.out_dim // 4, "gemm"))
            tr.append(("linear", nq, m.out_dim // 4, m.out_dim, closing))
            if has_sc:
                tr.append(("linear", nq, m.in_dim, m.out_dim, closing))
        tr.append(("kpconv", nq, ns, None, c.K, c.in_channels, c.out_channels))
    return tr


def run_ours(args):
    from concurrent.futures import ThreadPoolExecutor
    from apr_b200 import _native, blocks, dataloader, ops
    from apr_b200.architectures import KPFCNNEncoder
    from apr_b200.pipeline import KFEPipeline
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _native.require_cuda()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = kitti_config() if args.workload != "nuscenes" else nuscenes_config()
    blocks.LINEAR_MODE = "tf32"
    for kv in args.opt:                                       # A/B switches of the native library (aprb_set_option)
        k, v = kv.split("=")
        _native.check(_native.lib().aprb_set_option(k.encode(), int(v)), "aprb_set_option")
    S = max(1, args.streams)
    P = max(1, args.batch)                # pairs stacked per aprb_kfe_forward call (super-batch, per-pair InstanceNorm)

    # ---- inputs: distinct pairs per rank, first-level 0.3 m voxelisation done up front (not part of the path)
    single_dev, pairs_dev, pairs_host = [], [], []
    from apr_b200.shard import shard_indices
    n_distinct = max(args.pairs, S, P + 2 if P > 1 else 0)
    seeds = shard_indices(n_distinct * world, rank, world)              # round-robin over the global pair list
    kind, distant, wl_name = WORKLOADS[args.workload]
    by_seed = {}

    def voxelised(sd):
        if sd not in by_seed:
            a, b = synth.pair_raw(sd, kind, distant)
            raw = torch.from_numpy(np.concatenate([a, b])).to(dev)
            lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
            p0, l0 = ops.grid_subsample(raw, lens, cfg.first_subsampling_dl)
            by_seed[sd] = (p0.contiguous().clone(), l0.clone(), (a, b))
        return by_seed[sd]

    for sd in seeds:
        single_dev.append(voxelised(sd)[:2])
    for j in range(max(args.pairs, S)):                                 # call j carries P distinct pairs, stacked
        sel = [single_dev[(j + t) % n_distinct] for t in range(P)]
        p0 = torch.cat([x[0] for x in sel]).contiguous(); l0 = torch.cat([x[1] for x in sel]).contiguous()
        pairs_dev.append((p0, l0))
        pairs_host.append((p0.cpu().pin_memory(), l0.cpu().pin_memory()))
    torch.manual_seed(0); np.random.seed(0)
    enc = KPFCNNEncoder(cfg).to(dev).eval()
    # neighbourhood limits: calibrate_neighbors (dataloader.py:200-232) over the first CALIB_PAIRS pairs of the synthetic
    # set — the same pairs, on every rank and in the reference arm, so that all arms run the same widths
    calib = [voxelised(sd)[:2] for sd in range(CALIB_PAIRS)] if kind == "kitti" and not distant else single_dev
    limits = dataloader.calibrate_neighbors_device(calib, cfg, samples_threshold=10 ** 9) if not args.no_calibrate else LIMITS_FALLBACK
    limits = [int(x) for x in limits]
    torch.cuda.synchronize(dev)

    # ---- S native pipelines, one CUDA stream and one host thread each (aprb_kfe_forward releases the GIL)
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    ups = "nearest" if args.upsamples == "nearest" else True
    pipes = [KFEPipeline(enc, cfg, limits, build_upsamples=ups, stream=streams[k], clouds_per_segment=2 if P > 1 else 0)
             for k in range(S)]
    pool = ThreadPoolExecutor(max_workers=S) if S > 1 else None
    main_stream = torch.cuda.current_stream(dev)
    done_ev = [torch.cuda.Event() for _ in range(S)]

    def fan_out(fn, i):
        """Run fn(k, pair_index) for the S pipelines concurrently; device-side ordering via events."""
        start = torch.cuda.Event()
        start.record(main_stream)
        def work(k):
            torch.cuda.set_device(dev)                       # worker threads start on device 0
            streams[k].wait_event(start)
            r = fn(k, (i * S + k))
            done_ev[k].record(streams[k])
            return r
        res = list(pool.map(work, range(S))) if pool else [work(0)]
        for k in range(S):
            main_stream.wait_event(done_ev[k])
        return res

    def step_dev(i):
        return fan_out(lambda k, j: pipes[k].forward(*pairs_dev[j % len(pairs_dev)]), i)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sampler=None):
        if sampler is not None:
            sampler.start()          # forks nvidia-smi: do it BEFORE the warm-up so the fork's page-table churn is absorbed there
        for i in range(warmup):
            flush.zero_()
            fn(i)
        if sampler is not None:
            sampler.mark()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        l0 = _native.launch_count()
        t0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()                                    # L2 flush (256 MiB), outside the timed event pair
            ev[i][0].record(main_stream)
            r = fn(warmup + i)
            ev[i][1].record(main_stream)
        barrier()
        wall = time.perf_counter() - t0
        per = [a.elapsed_time(b) for a, b in ev]
        timed.last_steps = [round(t, 3) for t in per]
        return sum(per), wall, _native.launch_count() - l0, r

    clk = ClockSampler(local)
    ms_dev, wall_dev, launches, _ = timed(step_dev, args.steps, args.warmup, sampler=clk)
    steps_dev = timed.last_steps
    clocks = clk.stop()

    # the same region with the FULL [N_l, limit] upsample matrices of the reference's collate (the default builds only their
    # column 0): recorded beside the headline so that the cost of the columns nothing reads is in the line, not hidden
    full_ups = None
    if args.upsamples == "nearest":
        pipes_full = [KFEPipeline(enc, cfg, limits, build_upsamples=True, stream=streams[k], clouds_per_segment=2 if P > 1 else 0)
                      for k in range(S)]
        n_full = max(4, args.steps // 2)
        ms_f, _, _, _ = timed(lambda i: fan_out(lambda k, j: pipes_full[k].forward(*pairs_dev[j % len(pairs_dev)]), i), n_full, 4)
        tf_ = torch.tensor([ms_f], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
        full_ups = {"value": 2.0 * S * P * n_full * world / (float(tf_.item()) * 1e-3), "unit": UNIT, "steps": n_full,
                    "how": "same timed region with build_upsamples = full [N_l, limit] matrices (--upsamples full makes it the headline)"}
        del pipes_full
        torch.cuda.empty_cache()

    out_rows_cap = max(4096, int(pairs_host[0][0].shape[0]) // 4)     # the last level keeps ~6 % of the level-0 rows
    out_cols = pipes[0]._out_cols
    host_out_by_dt = {dt: [[torch.empty((out_rows_cap, out_cols), dtype=dt).pin_memory() for _ in range(2)] for _ in range(S)]
                      for dt in (torch.float16, torch.float32)}
    host_out = host_out_by_dt[torch.float16 if args.e2e_out == "f16" else torch.float32]
    checks = [0.0] * S

    def timed_e2e(steps, warmup):
        """End to end through the host-buffer entry point, free-running: every stream/host thread pushes its own `steps`
        calls back to back (H2D -> path -> D2H -> stream sync per call), so one call's copies overlap the other
        streams' kernels instead of all calls of a step finishing, and copying out, at the same moment. Timed with one
        CUDA event pair on the main stream around the whole region (all streams fork from / join into it)."""
        def run(n_calls, first):
            start = torch.cuda.Event(); start.record(main_stream)
            def work(k):
                # two calls in flight per stream: call c is queued (H2D -> path -> D2H on the pipeline's copy stream into
                # one of two pinned buffers), then the result of call c-1 is awaited and read
                torch.cuda.set_device(dev)
                streams[k].wait_event(start)
                hb = db = 0
                pending = None
                for c in range(n_calls):
                    hp, hl = pairs_host[((first + c) * S + k) % len(pairs_host)]
                    y, ticket = pipes[k].forward_host_async(hp, hl, host_out[k][c & 1])
                    hb += hp.numel() * 4 + hl.numel() * 4; db += y.numel() * y.element_size()
                    if pending is not None:
                        pipes[k].wait_host(pending[1]); checks[k] = float(pending[0][0, 0])
                    pending = (y, ticket)
                if pending is not None:
                    pipes[k].wait_host(pending[1]); checks[k] = float(pending[0][0, 0])
                done_ev[k].record(streams[k])
                return hb, db
            res = list(pool.map(work, range(S))) if pool else [work(0)]
            for k in range(S):
                main_stream.wait_event(done_ev[k])
            return sum(r[0] for r in res), sum(r[1] for r in res)
        run(warmup, 0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(main_stream)
        hb, db = run(steps, warmup)
        e1.record(main_stream)
        barrier()
        return e0.elapsed_time(e1), time.perf_counter() - t0, (hb // steps, db // steps)

    ms_e2e, wall_e2e, io = timed_e2e(args.steps, max(args.warmup, 3))
    # the other output dtype beside it (the reference hands back fp32; fp16 is the lossless-above-2^-14 compact form)
    other_dt = torch.float32 if args.e2e_out == "f16" else torch.float16
    host_out = host_out_by_dt[other_dt]
    ms_e2e_other, _, io_other = timed_e2e(max(3, args.steps // 2), 3)
    steps_other = max(3, args.steps // 2)

    # ---- strict drop-in leg: ONE pair per step exactly as the reference's collate + model run it —
    # numpy in / numpy out through the cpp_wrappers-compatible modules (every subsample_batch / batch_query call pays its
    # own H2D + D2H), the collated batch moved to the device, the module-path encoder, the output copied back.
    def dropin_leg(n_steps):
        from apr_b200 import dataloader as dl
        items = [voxelised(sd) for sd in seeds[:min(len(seeds), 3)]]
        np_pairs = []
        for p0, l0, _ in items:
            h = p0.cpu().numpy(); n0 = int(l0[0])
            np_pairs.append((h[:n0].copy(), h[n0:].copy()))
        def one(i):
            src, tgt = np_pairs[i % len(np_pairs)]
            batch = dl.collate_fn_descriptor(dl.make_list_data(src, tgt), cfg, limits)
            gpu = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v)
                   for k, v in batch.items() if k in ("points", "neighbors", "pools", "upsamples", "stack_lengths")}
            gpu["features"] = batch["features"].to(dev)
            with torch.no_grad():
                return enc(gpu).cpu()
        for i in range(2 * len(np_pairs)):                  # warm-up over every distinct pair: workspaces and the page-locked
            one(i)                                           # staging blocks of torch's caching host allocator (size classes)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(n_steps):
            y = one(i)
        torch.cuda.synchronize(dev)
        return 2.0 * n_steps / (time.perf_counter() - t0), int(y.shape[0])

    dropin_value, _ = dropin_leg(6) if rank == 0 else (None, 0)

    # ---- from RAW scans: pinned host [n, 4] float32 xyzr rows (datasets/kitti.py:191-194) -> H2D ->
    # first-level voxelisation at 0.3 m with open3d's voxel_down_sample semantics (kitti.py:588-589) -> the path -> D2H (fp32)
    def raw_leg(steps, warmup):
        raw_host = []
        for j in range(S):                                              # one stacked call per stream, P pairs each
            scans = []
            for t in range(P):
                a, b = voxelised(seeds[(j + t) % n_distinct])[2]
                scans += [a, b]
            xyz = np.concatenate(scans)
            raw4 = np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1)
            raw_host.append((torch.from_numpy(raw4).pin_memory(),
                             torch.tensor([len(c) for c in scans], dtype=torch.int32).pin_memory()))
        outs = host_out_by_dt[torch.float32]
        def run(n_calls):
            start = torch.cuda.Event(); start.record(main_stream)
            def work(k):
                torch.cuda.set_device(dev)
                streams[k].wait_event(start)
                for c in range(n_calls):
                    pipes[k].forward_from_raw(raw_host[k][0], raw_host[k][1], outs[k][c & 1])
                done_ev[k].record(streams[k])
            list(pool.map(work, range(S))) if pool else work(0)
            for k in range(S):
                main_stream.wait_event(done_ev[k])
        run(warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        run(steps)
        e1.record(main_stream)
        barrier()
        return e0.elapsed_time(e1), int(sum(r[0].numel() * 4 + r[1].numel() * 4 for r in raw_host)), int(sum(len(r[0]) for r in raw_host))

    raw_steps = max(3, args.steps // 2)
    ms_raw, raw_h2d, raw_points = raw_leg(raw_steps, 2)

    # ---- per-kernel pass (one stream, CUDA events around every launch on the launching stream) for the roofline
    _native.prof_enable(True); _native.prof_report()
    for i in range(args.steps):
        flush.zero_()
        pipes[0].forward(*pairs_dev[i % len(pairs_dev)])
    prof = _native.prof_report()
    _native.prof_enable(False)
    # the KPConv contractions are timed under their own label; the tensor-path figures below read the persistent GEMM as one
    # kernel (both labels), the `kernels` table and the whole-operator KPConv figure keep them apart
    prof_split = dict(prof)
    if "kpconv_gemm_kernel" in prof:
        c_k, m_k = prof.pop("kpconv_gemm_kernel")
        c_g, m_g = prof.get("gemm_tf32_kernel", (0, 0.0))
        prof["gemm_tf32_kernel"] = (c_g + c_k, m_g + m_k)
    pyr = pipes[0].pyramid()
    n_levels = [int(p.shape[0]) for p in pyr["points"]]
    trace = encoder_shapes(enc, n_levels)
    kp_flops = sum(2.0 * r[1] * r[4] * r[5] * r[6] for r in trace if r[0] == "kpconv")
    lin_flops = sum(2.0 * r[1] * r[2] * r[3] for r in trace if r[0] == "linear")
    # tensor-path flops: KPConv contractions that run on tcgen05 (K*Cin % 32 == 0) + the unary Linear layers
    recompute = int(dict(kv.split("=") for kv in args.opt).get("gemm_apply", 1)) != 0
    nrm_flops = sum(2.0 * r[1] * r[2] * r[3] for r in trace if r[0] == "linear" and r[4] == "nrm") if recompute else 0.0
    # flops of the launches timed under "gemm_tf32_kernel": KPConv contractions on tcgen05, unary1, stored-path closing Linears
    tc_flops = sum(2.0 * r[1] * r[4] * r[5] * r[6] for r in trace if r[0] == "kpconv" and (r[4] * r[5]) % 32 == 0) \
        + lin_flops - nrm_flops
    # recompute kernels: operands twice (both passes) + fp16 output + fp16 residual rows where the shortcut is not a product
    nrm_blocks = {}
    for r in trace:
        if r[0] == "linear" and r[4] == "nrm":
            nrm_blocks.setdefault((r[1], r[3]), []).append(r[2])
    nrm_bytes = sum(2 * 2.0 * n * sum(ks) + 2.0 * n * c + (2.0 * n * c if len(ks) == 1 else 0.0) for (n, c), ks in nrm_blocks.items())
    opts = dict(kv.split("=") for kv in args.opt)
    f16_mode = int(opts.get("act_f16", 1)) != 0 and int(opts.get("kpconv_f16", 1)) != 0
    # kp_weighted4 (every KPConv but the first, whose Cin = 1 runs kp_weighted_c1): gathers x [Ns, Cin] and writes the
    # weighted tile [Nq, K*Cin], both fp16 in the default mode (fp32 otherwise), + int32 indices + fp32 points
    eb = 2.0 if f16_mode else 4.0
    kpw_bytes = sum(eb * (r[2] * r[5] + r[1] * r[4] * r[5]) + 4.0 * r[1] * limits[0] + 12.0 * (r[1] + r[2])
                    for r in trace if r[0] == "kpconv" and r[5] > 1)

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_dev, ms_e2e, ms_e2e_other, ms_raw], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_e2e_other, ms_raw = t.tolist()
    clouds = 2.0 * S * P * args.steps * world
    value = clouds / (ms_dev * 1e-3)
    e2e = clouds / (ms_e2e * 1e-3)

    pk = peaks()

    def traffic_of(kernel):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this
        workload (profiles/r02_traffic.json), or None when no capture covers the kernel / the options differ."""
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")) as f:
                t = json.load(f)
            ok = t.get("workload") == wl_name and t.get("pairs_per_call") == P and not args.opt
            return float(t["kernels"][kernel]["dram_bytes_per_launch"]) if ok and kernel in t.get("kernels", {}) else None
        except Exception:
            return None

    total_kernel_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1]) if prof else (None, (0, 0.0))
    roof = None
    if top[0] is not None:
        name, (cnt, tot_ms) = top
        per_step_s = tot_ms / args.steps * 1e-3
        share = tot_ms / max(total_kernel_ms, 1e-9)
        if name in ("gemm_tf32_kernel", "sgemm_rowscale_kernel"):
            # fp16 operands (default: act_f16 = kpconv_f16 = 1) run at the bf16/fp16 dense rate, TF32 at half of it
            peak = pk["bf16"] if f16_mode else pk["bf16"] / 2.0
            ach = tc_flops / per_step_s / 1e12 if per_step_s > 0 else 0.0
            roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic_of(name), "launches_per_step": cnt / args.steps,
                    "note": f"algorithmic flops = 2*Nq*K*Cin*Cout over the KPConv contractions + 2*N*Cin*Cout over the unary "
                            f"Linears this kernel runs (unary1 and the stored-path blocks; the recomputed closing Linears, "
                            f"{nrm_flops / 1e9:.1f} GFLOP/call, are gemm_nrm launches) = {tc_flops / 1e9:.1f} GFLOP/call ({P} pair(s)); peak = {pk['src']} bf16 sustained"
                            f"{'' if f16_mode else ' / 2 (TF32)'}; the HBM-bound launches of this kernel (levels 0-1) are read in "
                            f"profiles/; share of kernel time {share:.2f}"}
        else:
            by = kpw_bytes if name == "kp_weighted_kernel" else (nrm_bytes if name.startswith("gemm_nrm") else 0.0)
            ach = by / per_step_s / 1e9 if per_step_s > 0 else 0.0
            roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                    "traffic": traffic_of(name), "launches_per_step": cnt / args.steps,
                    "note": (f"algorithmic bytes of the recompute kernels (fp16 operands of both passes + fp16 output + fp16 residual "
                             f"rows) = {by / 1e6:.0f} MB/call over the stats + apply launches; peak = {pk['src']} HBM copy bandwidth; "
                             f"share of kernel time {share:.2f}") if name.startswith("gemm_nrm") else
                            f"algorithmic bytes = {int(eb)}*(Ns*Cin + Nq*K*Cin) + 4*Nq*H + 12*(Nq+Ns) = {by / 1e6:.0f} MB/call "
                            f"({P} pair(s)); peak = {pk['src']} HBM copy bandwidth; the kernel is instruction-issue bound "
                            f"(profiles/), so this fraction is its distance from the HBM floor; share of kernel time {share:.2f}"}
    # the tensor path beside the dominant kernel (north_star: tensor-pipe evidence for the KPConv contraction)
    roof_tensor = None
    if "gemm_tf32_kernel" in prof and prof["gemm_tf32_kernel"][1] > 0:
        cnt_g, ms_g = prof["gemm_tf32_kernel"]
        peak = pk["bf16"] if f16_mode else pk["bf16"] / 2.0
        ach = tc_flops / (ms_g / args.steps * 1e-3) / 1e12
        roof_tensor = {"kernel": "gemm_tf32_kernel", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                       "frac": ach / peak, "traffic": traffic_of("gemm_tf32_kernel"), "launches_per_step": cnt_g / args.steps,
                       "note": f"persistent tcgen05 GEMM: KPConv contractions + unary1 (+ stored-path closing Linears) = "
                               f"{tc_flops / 1e9:.1f} GFLOP/call; blended over its HBM-bound level-0/1 and tensor-bound level-2/3 "
                               f"launches (per launch: profiles/r02_ncu_traffic_summary.txt)"}
    kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in
               sorted(prof_split.items(), key=lambda kv: -kv[1][1])[:15]}
    # KPConv as ONE operator (north_star subsystem 3): row sums + weighting stage + contraction, all 11 KPConvs of the encoder
    op_names = [k for k in ("rowsum_pos_kernel", "kp_weighted_kernel", "kp_weighted_c1_kernel", "kpconv_gemm_kernel",
                            "sgemm_rowscale_kernel") if k in prof_split]
    op_ms = sum(prof_split[k][1] for k in op_names) / args.steps
    roof_kpconv = None
    if op_ms > 0 and "kpconv_gemm_kernel" in prof_split:
        peak = pk["bf16"] if f16_mode else pk["bf16"] / 2.0
        ach = kp_flops / (op_ms * 1e-3) / 1e12
        wf_bytes = sum((2.0 if f16_mode else 4.0) * r[1] * r[4] * r[5] for r in trace if r[0] == "kpconv" and r[5] > 1)
        roof_kpconv = {"operator": "kpconv (row sums + weighting + contraction)", "kernels": op_names, "bound": "tensor",
                       "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "ms_per_call": op_ms,
                       "weighted_tile_bytes_per_call": wf_bytes,
                       "note": f"2*Nq*K*Cin*Cout over the encoder's KPConvs = {kp_flops / 1e9:.1f} GFLOP/call ({P} pair(s)) over the summed "
                               f"kernel time of the operator's stages; the weighted tile [Nq, K*Cin] ({wf_bytes / 1e9:.2f} GB/call) is "
                               f"written by the weighting kernel and read back by the contraction: the stage split that bounds this "
                               f"figure (DESIGN.md 4.2, 4.4b: both one-kernel forms were built and are slower)"}

    # ---- the memory-bound stages north_star names, each against the HBM peak with SURVEY.md 8d's algorithmic bytes
    Bc = int(pairs_dev[0][1].shape[0])
    L = len(n_levels)
    nb_bytes = 0.0
    for l in range(L):
        W = limits[l]
        nb_bytes += 24.0 * n_levels[l] + 8.0 * Bc + 4.0 * n_levels[l] * W                                    # conv(l)
        if l + 1 < L:
            nb_bytes += 12.0 * (n_levels[l + 1] + n_levels[l]) + 8.0 * Bc + 4.0 * n_levels[l + 1] * W          # pool(l)
            nb_bytes += 12.0 * (n_levels[l] + n_levels[l + 1]) + 8.0 * Bc + 4.0 * n_levels[l] * (1 if args.upsamples == "nearest" else W)   # upsample(l)
    sub_bytes = sum(12.0 * (n_levels[l] + n_levels[l + 1]) + 8.0 * Bc for l in range(L - 1))
    def stage(names, nbytes, what):
        ms = sum(prof[n][1] for n in names if n in prof) / args.steps
        cnt = sum(prof[n][0] for n in names if n in prof) / args.steps
        ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernels": [n for n in names if n in prof], "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": None, "ms_per_call": ms, "launches_per_call": cnt,
                "algorithmic_bytes_per_call": nbytes, "note": what}
    roof_stages = {
        "radius_search": stage(["nb_query_kernel", "nb_nearest_kernel", "nb_count_kernel", "nb_scatter_kernel", "nb_grid_params_kernel"], nb_bytes,
                               "10 searches on 4 cell lists; bytes = 12*Nq + 12*Ns + 8*B + 4*Nq*W per search (SURVEY 8d; W = 1 for the "
                               "upsample searches when only their nearest column is built). The query "
                               "kernel is instruction-issue bound (candidate scan + (d2, index) sort), not HBM bound: this fraction is "
                               "its distance from the HBM floor"),
        "grid_subsample": stage(["sub_params_kernel", "sub_keys_kernel", "cub_radix_sort_pairs32", "cub_radix_sort_pairs64",
                                 "sub_flags_kernel", "sub_lens_kernel", "sub_emit_kernel"], sub_bytes,
                                "3 subsamplings; bytes = 12*N_in + 12*M_out + 8*B each (SURVEY 8d); 6.7 MB per call spread over "
                                "~20 dependent launches: launch-latency bound at this size"),
    }

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if f16_mode else "tf32", "data": "synthetic",
            "config": {"workload": wl_name, "pairs_per_step": S * P, "pairs_per_call": P,
                       "concurrent_streams": S, "points_stacked": int(pairs_dev[0][0].shape[0]), "level_points": n_levels,
                       "limits": limits,
                       "upsample_matrices": ("column 0 only ([N_l, 1]: the nearest support, all the reference reads of an upsample "
                                             "matrix — closest_pool inds[:, 0], blocks.py:71-83; SURVEY 8f-3; bit-identical to "
                                             "column 0 of the full search); --upsamples full builds [N_l, limit]"
                                             if args.upsamples == "nearest" else "full [N_l, limit] matrices like the reference's collate"),
                       "parallelism": f"pairs x{world}", "path": "native (aprb_kfe_forward)",
                       "l2": "flushed between steps (256 MiB memset outside the event pair)",
                       "precision": ("fp16 operands (10-bit mantissa, as TF32) with fp32 accumulation in TMEM for every contraction; "
                                     "normalised activations stored in fp16 (exact: they are TF32-rounded); neighbour search, "
                                     "subsampling, influence weights, statistics in fp32/integer") if f16_mode else
                                    "TF32 operands, fp32 accumulation; activations fp32",
                       "kpconv_gflop_per_pair": kp_flops / 1e9 / P, "linear_gflop_per_pair": lin_flops / 1e9 / P},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                    "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps,
                    "out_dtype": args.e2e_out,
                    "how": "aprb_kfe_forward_host_async from pinned host buffers (H2D, path, D2H of the encoder output through the "
                           "pipeline's copy stream; each result awaited and read one call later), streams free-running over "
                           "the K steps (working set per call >> L2: no flush needed), one CUDA event pair around the region; "
                           + ("host output in fp16: the final activation is rounded to a 10-bit mantissa, so it converts back to "
                              "the same fp32 values (|v| >= 2^-14), half the PCIe bytes; --e2e-out f32 copies fp32"
                              if args.e2e_out == "f16" else "host output in fp32")},
            "e2e_other_dtype": {"value": 2.0 * S * P * steps_other * world / (ms_e2e_other * 1e-3), "unit": UNIT,
                                "out_dtype": "f32" if args.e2e_out == "f16" else "f16", "steps": steps_other,
                                "d2h_bytes_per_step": int(io_other[1])},
            "e2e_from_raw": {"value": 2.0 * S * P * raw_steps * world / (ms_raw * 1e-3), "unit": UNIT, "steps": raw_steps,
                             "h2d_bytes_per_step": raw_h2d, "raw_points_per_step": raw_points,
                             "how": "KFEPipeline.forward_from_raw: pinned raw [n, 4] float32 scans (x, y, z, reflectance) -> H2D -> "
                                    "aprb_voxel_downsample_raw at first_subsampling_dl (open3d 0.10 voxel_down_sample semantics: "
                                    "double arithmetic, origin = min - voxel/2; parity unpinned, open3d absent) -> pyramid + encoder "
                                    "-> D2H of the fp32 output; one stream sync per call, S streams free-running"},
            "dropin_e2e": ({"value": dropin_value, "unit": UNIT, "sample": "6 steps x 1 pair on rank 0, after a warm-up over every distinct pair",
                            "how": "strict drop-in, one pair per step like the reference: numpy in / numpy out through "
                                   "apr_b200.dataloader.collate_fn_descriptor (13 cpp_wrappers-compatible calls, each with its own "
                                   "H2D + D2H, int64 indices), H2D of the collated batch, the module-path encoder "
                                   "(apr_b200.blocks), D2H of the fp32 output; wall clock"} if dropin_value else None),
            "value_full_upsample_matrices": full_ups,
            "roofline": roof, "roofline_tensor": roof_tensor, "roofline_kpconv_operator": roof_kpconv,
            "roofline_stages": roof_stages, "kernels": kernels,
            "wall_ms_per_step": 1e3 * wall_dev / args.steps, "ms_steps": steps_dev}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cfg_r = kitti_config()
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        ref, kind, pairs, sd = reference_setup(cfg_r, 2, 0)
        reference_step(ref, None, cfg_r, sd, limits, *pairs[1])          # untimed warm-up (thread pools, allocator)
        done, dt = reference_run(ref, cfg_r, sd, limits, pairs, 8, 15.0)   # bounded sample: ~10-20 s of CPU work
        line["cpu_baseline"] = {"value": 2.0 * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"{done} KITTI-shaped pairs in {dt:.1f} s, same limits as the GPU arm: pyramids by "
                                          "oracle/_ref (reference object code, pair-parallel over the host threads as DataLoader "
                                          "workers would) + KFE encoder by oracle/blocks_ref.py (the builder's fp32 torch-CPU "
                                          "restatement of models/blocks.py, all threads)"}
        # the reference's deployment shape beside it: host pyramids (reference object code) + eager PyTorch blocks on THIS GPU
        pairs8 = [pairs[i % len(pairs)] for i in range(8)]
        gdone, gdt, gpyr = reference_gpu_eager_run(ref, cfg_r, sd, limits, pairs8, dev, 20.0)
        line["ref_gpu_eager"] = {"value": 2.0 * gdone / gdt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                 "host_pyramid_share": gpyr / gdt,
                                 "sample": f"{gdone} KITTI-shaped pairs in {gdt:.1f} s: pyramids on the host by oracle/_ref "
                                           "(pair-parallel), int64 batch shipped to the device, encoder = oracle/blocks_ref.py run "
                                           "as eager PyTorch CUDA ops (fp32, cuBLAS): the reference's own GPU path "
                                           "(lib/tester.py:49-57), restated"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if pool:
        pool.shutdown()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ config 3: full KPFCNN
def run_kpfcnn(args):
    """BASELINE configs[2]: full KPFCNN encoder-decoder forward (models/architectures.py:137-212) on LoKITTI-shaped distant
    pairs, sharded by pair over the ranks (no collective). Per call: P stacked pairs through KPFCNNPipeline (native KFE
    encoder, then bottleneck / GCN / cross-saliency / decoder stream-ordered in the same call). `value` device-resident,
    `e2e` from pinned host points to pinned host (feats_f, scores) with one stream sync per call."""
    from concurrent.futures import ThreadPoolExecutor
    from apr_b200 import _native, blocks, dataloader, ops
    from apr_b200.architectures import KPFCNN
    from apr_b200.pipeline import KPFCNNPipeline
    from apr_b200.shard import shard_indices
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _native.require_cuda()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = kitti_config() if args.workload != "nuscenes" else nuscenes_config()
    blocks.LINEAR_MODE = "tf32"
    S, P = max(1, args.streams), max(1, args.batch)
    kind, distant, wl_name = WORKLOADS[args.workload]
    n_distinct = max(args.pairs, S, P + 2)
    single = []
    for sd in shard_indices(n_distinct * world, rank, world):
        a, b = synth.pair_raw(sd, kind, distant)
        raw = torch.from_numpy(np.concatenate([a, b])).to(dev)
        lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
        p0, l0 = ops.grid_subsample(raw, lens, cfg.first_subsampling_dl)
        single.append((p0.contiguous().clone(), l0.clone()))
    calls = []
    for j in range(max(args.pairs, S)):
        sel = [single[(j + t) % n_distinct] for t in range(P)]
        p0 = torch.cat([x[0] for x in sel]).contiguous(); l0 = torch.cat([x[1] for x in sel]).contiguous()
        calls.append((p0, l0, p0.cpu().pin_memory(), l0.cpu().pin_memory()))
    limits = [int(x) for x in dataloader.calibrate_neighbors_device(single, cfg, samples_threshold=10 ** 9)]
    torch.manual_seed(0); np.random.seed(0)
    net = KPFCNN(cfg).to(dev).eval()
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    pipes = [KPFCNNPipeline(net, cfg, limits, stream=streams[k], clouds_per_segment=2) for k in range(S)]
    pool = ThreadPoolExecutor(max_workers=S) if S > 1 else None
    main_stream = torch.cuda.current_stream(dev)
    done_ev = [torch.cuda.Event() for _ in range(S)]
    n0max = max(c[0].shape[0] for c in calls)
    host_out = [torch.empty((n0max, cfg.final_feats_dim + 2), dtype=torch.float32).pin_memory() for _ in range(S)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run(n_calls, first, from_host):
        start = torch.cuda.Event(); start.record(main_stream)
        def work(k):
            torch.cuda.set_device(dev)
            streams[k].wait_event(start)
            hb = db = 0
            for c in range(n_calls):
                p0, l0, hp, hl = calls[((first + c) * S + k) % len(calls)]
                if from_host:
                    with torch.cuda.stream(streams[k]):
                        p0 = hp.to(dev, non_blocking=True); l0 = hl.to(dev, non_blocking=True)
                ff, so, ss = pipes[k].forward(p0, l0)
                if from_host:
                    with torch.cuda.stream(streams[k]):
                        out = host_out[k][:ff.shape[0]]
                        out.copy_(torch.cat([ff, so[:, None], ss[:, None]], dim=1), non_blocking=True)
                    streams[k].synchronize()
                    hb += hp.numel() * 4 + hl.numel() * 4; db += out.numel() * 4
            done_ev[k].record(streams[k])
            return hb, db
        res = list(pool.map(work, range(S))) if pool else [work(0)]
        for k in range(S):
            main_stream.wait_event(done_ev[k])
        return sum(r[0] for r in res), sum(r[1] for r in res)

    def timed(steps, warmup, from_host):
        run(warmup, 0, from_host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0c = _native.launch_count()
        e0.record(main_stream)
        io = run(steps, warmup, from_host)
        e1.record(main_stream)
        barrier()
        return e0.elapsed_time(e1), io, _native.launch_count() - l0c

    clk = ClockSampler(local); clk.start()
    ms_dev, _, launches = timed(args.steps, args.warmup, False)
    clocks = clk.stop()
    ms_e2e, io, _ = timed(args.steps, 3, True)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = t.tolist()
    clouds = 2.0 * S * P * args.steps * world
    if rank == 0:
        print(json.dumps({"metric": METRIC.replace("KFE)", "full KPFCNN encoder-decoder)"), "value": clouds / (ms_dev * 1e-3), "unit": UNIT,
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                          "config": {"workload": wl_name.replace("kfe_encoder", "kpfcnn_forward"), "pairs_per_step": S * P,
                                     "pairs_per_call": P, "concurrent_streams": S, "points_stacked": int(calls[0][0].shape[0]),
                                     "limits": limits, "parallelism": f"pairs x{world}",
                                     "path": "KPFCNNPipeline: native aprb_kfe_forward + bottleneck / GCN / decoder stream-ordered",
                                     "l2": "working set per call >> L2, streams free-running"},
                          "clocks": clocks, "gpu_launches": int(launches),
                          "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(io[0] // args.steps),
                                  "d2h_bytes_per_step": int(io[1] // args.steps),
                                  "how": "pinned host points -> H2D -> KPFCNNPipeline.forward -> D2H of [N0, final_feats_dim + 2] fp32 "
                                         "(feats_f, scores_overlap, scores_saliency), one stream sync per call"}}))
    if pool:
        pool.shutdown()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------ training step
def run_train(args):
    """BASELINE configs[4]: one APR training step per rank and step — differentiable KPFCNN forward (our KPConv kernels),
    NPR generative loss + descriptor/score surrogate, backward (scatter-add data gradients), bucketed NCCL all-reduce
    overlapped with backward, SGD (lr 0.01, momentum 0.98, wd 1e-6: main.py:66-75). Not the headline metric."""
    from apr_b200 import _native, dataloader, ops, train
    from apr_b200.architectures import KPFCNN
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _native.require_cuda()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = kitti_config()
    kind, distant, wl_name = WORKLOADS[args.workload]
    from apr_b200.shard import shard_indices
    seeds = shard_indices(max(args.pairs, 2) * world, rank, world)
    items = []
    for sd in seeds:
        a, b = synth.pair_raw(sd, kind, distant)
        d = synth.pair_pose(sd, distant)
        raw = torch.from_numpy(np.concatenate([a, b])).to(dev)
        lens = torch.tensor([len(a), len(b)], dtype=torch.int32, device=dev)
        p0, l0 = ops.grid_subsample(raw, lens, cfg.first_subsampling_dl)
        apc, apc_l = ops.grid_subsample(raw, lens, 0.2)             # denser 'aggregated' cloud standing in for src/tgt_nghb
        n_src = int(l0[0])
        shift = torch.tensor([d, 0.0, 0.0], device=dev)             # tgt sensor frame -> src sensor frame
        nn_idx = ops.radius_neighbors(p0[:n_src], p0[n_src:] + shift, l0[:1], l0[1:], 0.45, 1)
        has = nn_idx[:, 0] < (p0.shape[0] - n_src)
        corr = torch.stack([torch.nonzero(has).flatten(), nn_idx[has, 0].long()], 1)
        items.append((p0.contiguous(), l0, n_src, corr, apc[:int(apc_l[0])].contiguous(), apc[int(apc_l[0]):].contiguous()))
    limits = [int(x) for x in dataloader.calibrate_neighbors_device([(it[0], it[1]) for it in items], cfg)]
    torch.manual_seed(0); np.random.seed(0)
    net = KPFCNN(cfg).to(dev)
    head = train.NPRHead(cfg.final_feats_dim, cfg.point_generation_ratio).to(dev)
    params = [p for p in list(net.parameters()) + list(head.parameters()) if p.requires_grad]
    red = train.GradBucketReducer(params, bucket_mb=25.0)
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.98, weight_decay=1e-6)
    torch.backends.cuda.matmul.allow_tf32 = True                    # the dense contractions run as TF32 library GEMMs
    losses = []

    def step(i):
        p0, l0, n_src, corr, apc_s, apc_t = items[i % len(items)]
        batch = dataloader.build_pyramid_device(p0, l0, cfg, limits)
        opt.zero_grad(set_to_none=True)
        ff, so, ss = train.kpfcnn_forward_train(net, batch)
        loss = train.surrogate_desc_loss(ff[:n_src], ff[n_src:], corr, so, ss, n_src) \
            + train.npr_loss(head, ff[:n_src], p0[:n_src], apc_s) + train.npr_loss(head, ff[n_src:], p0[n_src:], apc_t)
        loss.backward()
        red.finish()
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0c = _native.launch_count()
    e0.record()
    for i in range(args.steps):
        losses.append(step(args.warmup + i))
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    nparam = sum(p.numel() for p in params)
    if rank == 0:
        print(json.dumps({"metric": "APR training steps: pairs/sec (KFE + NPR head, fwd+bwd+allreduce+SGD)", "mode": "train",
                          "value": world * args.steps / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
                          "config": {"workload": wl_name + "_train", "pairs_per_rank_per_step": 1, "params": nparam,
                                     "allreduce_bytes_per_step": 4 * nparam if world > 1 else 0, "buckets": len(red.buckets),
                                     "limits": limits, "points_stacked": int(items[0][0].shape[0])},
                          "gpu_launches": int(_native.launch_count() - l0c),
                          "loss_first_last": [float(sum(losses[:3]) / 3), float(sum(losses[-3:]) / 3)],
                          "loss_note": "mean of the first 3 / last 3 timed steps (the warm-up steps already trained)",
                          "contractions": "tcgen05 TF32 GEMM (aprb_linear_tf32) for out = wf W, dwf = g W^T, dW = wf^T g where the "
                                          "shape allows; InstanceNorm+LeakyReLU forward/backward native; allreduce = bucketed NCCL, "
                                          "launched from gradient hooks during backward"}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=3, help="distinct synthetic pairs cycled through (per rank)")
    ap.add_argument("--no-calibrate", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=5, help="calls in flight per GPU (one CUDA stream + host thread each)")
    ap.add_argument("--batch", type=int, default=8, help="pairs stacked per call (super-batch; per-pair InstanceNorm segments)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = the headline metric (default); train = BASELINE configs[4] training step")
    ap.add_argument("--e2e-out", default="f16", choices=["f16", "f32"],
                    help="dtype of the encoder output copied to the host in the e2e leg (f16 is lossless for |v| >= 2^-14)")
    ap.add_argument("--opt", action="append", default=[], help="native tuning switch name=value (aprb_set_option)")
    ap.add_argument("--upsamples", default="nearest", choices=["nearest", "full"],
                    help="upsample searches of the device-resident pyramid: 'nearest' = only column 0 of each matrix (all the "
                         "reference reads of them: closest_pool, blocks.py:71-83; SURVEY 8f-3), 'full' = [N_l, limit] matrices")
    ap.add_argument("--net", default="kfe", choices=["kfe", "kpfcnn"],
                    help="kfe = the KFE encoder (the headline metric); kpfcnn = full encoder-decoder forward (BASELINE configs[2])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.mode == "train":
        return run_train(args)
    if args.net == "kpfcnn":
        return run_kpfcnn(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
