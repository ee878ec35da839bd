#!/usr/bin/env python
"""bench.py -- WSPR windows/s through coarse search + fine sync + soft-symbol demodulation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--windows M]

Workload (BASELINE.json configs[2], the largest single-GPU configuration): M = 10 000
independent synthetic WSPR windows per GPU (45 000 complex64 samples at 375 sps each, 3.6 GB,
far larger than L2), SNR ~ U(-30, 0) dB in 2500 Hz, carrier offset U(-6, 6) Hz, linear drift
U(-3, 3) Hz, random start; FDR(hbw=10, maxdrift=4, maxfreqs=200, thr=10).  One step = one pass
of the whole hot path over the batch: spectrogram + normalizer + peak pick, coarse search,
the refinement chain and the soft symbols of all 17 jiggled shifts of every gated candidate
(the decoder-independent upper bound of what the reference evaluates; Fano is excluded from
the metric on both sides).

  value  device-resident input, results left on the device (kernels only, CUDA events)
  e2e    the same step through the C ABI with pinned HOST buffers: host->device copy of the
         samples and device->host copy of every result inside the timed region

With N > 1 (torchrun) each rank runs the same per-GPU workload on its own GPU (weak scaling,
no data-path collective); time is the max over ranks, value the sum of windows / that time.
The host-fed arm (e2e) at N > 1 processes the same N x M windows per step but deals them to
the ranks in proportion to each rank's measured host-fed rate, because the host links of one
box are not equally fast when every GPU copies at once (--no-balance: equal split).
The synthetic inputs are made with the library's own encoder (uwspr_b200.channel_symbols);
oracle/ is used by the reference arm and the cpu_baseline leg only.
--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified
sources; the C restatement if it was not built) on all host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gr-uwspr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FL = 45000
try:
    _ALL_CPUS = set(os.sched_getaffinity(0))
except Exception:
    _ALL_CPUS = set(range(os.cpu_count() or 1))
PARAMS = dict(fs=375, fl=FL, spb=256, maxdrift=4, maxfreqs=200, halfbandwidth=10, cf=1500, threshold=10)
METRIC = "WSPR windows/sec (coarse+fine sync+demod)"
FLOP_SPEC, FLOP_COARSE, FLOP_POINT = 9.09e6, 5 * 26 * (2 * PARAMS["maxdrift"] + 126) * 162 * 9.0, 1.327e6


# ------------------------------------------------------------------ synthetic windows
SEED_BASE = 20190222


def message_bytes(rng):
    """50 random payload bits packed MSB-first into 7 bytes (last 6 bits zero)"""
    b = np.zeros(56, np.uint8)
    b[:50] = rng.integers(0, 2, 50, dtype=np.uint8)
    return np.packbits(b)


def gen_windows_torch(nwin, seed, device, maxdrift=3.0, batch=500):
    """synthetic windows of SURVEY 8(d) generated on the GPU (torch is plumbing here: the data
    generator is not part of the measured path).  Returns a (nwin, FL) complex64 CUDA tensor
    and the per-window truth."""
    import torch
    import uwspr_b200 as ub   # the library's own encoder: nothing under oracle/ feeds the measured arms
    rng = np.random.default_rng([SEED_BASE, seed])
    out = torch.empty((nwin, FL), dtype=torch.complex64, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_BASE + 7919 * seed)
    truth = []
    df = 375.0 / 256.0
    k = torch.arange(162 * 256, device=device)
    sym_idx = (k // 256)
    for b0 in range(0, nwin, batch):
        nb = min(batch, nwin - b0)
        msgs = [message_bytes(rng) for _ in range(nb)]
        syms = np.stack([ub.channel_symbols(m) for m in msgs]).astype(np.float64)
        f0 = rng.uniform(-6, 6, nb)
        drift = rng.uniform(-maxdrift, maxdrift, nb)
        start = 375 + rng.integers(0, 2561, nb)
        snr = rng.uniform(-30, 0, nb)
        st = torch.from_numpy(syms).to(device)
        f = (torch.from_numpy(f0).to(device)[:, None] + (st[:, sym_idx] - 1.5) * df
             + (torch.from_numpy(drift).to(device)[:, None] / 2.0) * ((sym_idx[None, :].double() - 81.0) / 81.0))
        phase = 2 * np.pi * torch.cumsum(f, dim=1) / 375.0
        sig = torch.polar(torch.ones_like(phase), phase).to(torch.complex64)
        sigma = np.sqrt((375.0 / 2500.0) / 10 ** (snr / 10.0) / 2.0)
        noise = torch.randn((nb, FL, 2), generator=gen, device=device, dtype=torch.float32)
        x = torch.view_as_complex(noise) * torch.from_numpy(sigma.astype(np.float32)).to(device)[:, None]
        # add frame i at sample start[i] (one vectorised scatter-add for the batch)
        pos = torch.from_numpy(start).to(device)[:, None] + k[None, :]
        ok = pos < FL
        flat = (torch.arange(nb, device=device)[:, None] * FL + pos)[ok]
        torch.view_as_real(x).view(-1, 2).index_add_(0, flat, torch.view_as_real(sig)[ok])
        out[b0:b0 + nb] = x
        truth += [dict(msg=m, f0=a, drift=d, start=int(s), snr=q) for m, a, d, s, q in zip(msgs, f0, drift, start, snr)]
    return out, truth


# ------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    path, lo, hi, use_ref, params = args
    sys.path.insert(0, ROOT)
    xs = np.load(path, mmap_mode="r")
    t_fano = 0.0
    ncand = 0
    if use_ref:
        from oracle import ref_binding as rb
        fdr = rb.RefFDR(**params)
        sd = rb.RefSD(params["fs"], params["fl"], params["spb"], params["maxdrift"], params["maxfreqs"], params["cf"], logdir="/tmp")
        rb.fano_stats(reset=True)
        t0 = time.perf_counter()
        for w in range(lo, hi):
            c, blobs, calls, fanos = rb.pipeline(fdr, sd, np.ascontiguousarray(xs[w]))
            ncand += len(c)
        dt = time.perf_counter() - t0
        t_fano = rb.fano_stats()[0]
    else:
        from oracle import port_binding as ob
        f = ob.OracleFDR(**params)
        t0 = time.perf_counter()
        for w in range(lo, hi):
            x = np.ascontiguousarray(xs[w])
            c = f.transform(x)
            t1 = time.perf_counter()
            ob.demodulate(x, c, cf=params["cf"], run_fano=True)
            ncand += len(c)
            _ = t1
        dt = time.perf_counter() - t0
    return hi - lo, dt, t_fano, ncand


def cpu_reference_run(xs_host, n_sample, cores):
    """windows/s of the reference CPU path on `cores` processes over the first n_sample windows.
    Decoder (Fano) time is measured by interposition and subtracted (it is outside the metric)."""
    import multiprocessing as mp
    from oracle import ref_binding as rb
    use_ref = rb.available()
    n_sample = min(n_sample, len(xs_host))
    path = "/dev/shm/uwspr_bench_%d.npy" % os.getpid()
    np.save(path, xs_host[:n_sample])
    try:
        bounds = [(n_sample * i) // cores for i in range(cores + 1)]
        jobs = [(path, bounds[i], bounds[i + 1], use_ref, PARAMS) for i in range(cores) if bounds[i + 1] > bounds[i]]
        ctx = mp.get_context("spawn")
        t0 = time.perf_counter()
        with ctx.Pool(len(jobs)) as pool:
            res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    finally:
        os.unlink(path)
    busy = max(r[1] - r[2] for r in res)          # slowest worker, decoder time removed
    n = sum(r[0] for r in res)
    fano = sum(r[2] for r in res)
    return dict(value=n / busy, unit="windows/s", cores=len(jobs), kind="reference" if use_ref else "port",
                sample="%d of the workload's windows, %d per core; decoder (Fano) time %.2f s of %.2f core-s subtracted; "
                       "stub FFT/PMT (see oracle/stubs); wall %.1f s incl. process start"
                       % (n, n // len(jobs), fano, sum(r[1] for r in res), wall),
                per_core=n / sum(r[1] - r[2] for r in res), candidates=sum(r[3] for r in res))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """median SM clock and the throttle reasons seen between wall-clock times t0 and t1
        (the timed region); samples are stamped when read from nvidia-smi's 50 ms loop"""
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 8 and (t0 is None or t0 - 0.05 <= t <= t1 + 0.05)]
        if not rows:
            rows = [r for t, r in self.rows if len(r) >= 8]
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None  # noqa: E731
        sm = sorted(v for v in (num(r[1]) for r in rows) if v is not None)
        mx = [v for v in (num(r[2]) for r in rows) if v is not None]
        pw = [v for v in (num(r[3]) for r in rows) if v is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, reasons=sorted(reasons), samples=len(sm))


def recorded_traffic(nwin):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json, written from `ncu --set full` by tools/ncu_summary.py --traffic);
    only returned when the capture was taken at this workload size"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if int(t.get("windows", -1)) == int(nwin):
            return float(t["k_fine_dram_bytes"])
    except Exception:
        pass
    return None


def bind_to_gpu_numa_node(local):
    """pins this rank to the CPUs of the NUMA node its GPU hangs off, so that the pinned host
    buffers it allocates next are local to the GPU's PCIe root (first touch).  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------ main arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import testdata as td
    cores = len(_ALL_CPUS) or os.cpu_count() or 1
    n = min(args.windows, max(cores, args.ref_windows_per_core * cores))
    xs = np.stack([td.synth_window(1000, w, maxdrift=3.0)[0] for w in range(min(n, 4 * cores))])
    # the sample is tiled from a few hundred distinct windows: CPU time per window is data independent
    # to first order (candidate count and gate outcomes vary), so distinct windows are kept to >= 4 per core
    reps = (n + len(xs) - 1) // len(xs)
    xs = np.concatenate([xs] * reps)[:n]
    vals = []
    for step in range(args.warmup + args.steps):
        r = cpu_reference_run(xs, n, cores)
        if step >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    line = dict(metric=METRIC, value=v, unit="windows/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * n / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="%d synthetic WSPR windows (SNR U(-30,0) dB, drift U(-3,3) Hz), FDR hbw=10 maxdrift=4 thr=10; bounded sample of the 10k-window workload" % n,
                            windows=n, **{k: PARAMS[k] for k in ("maxdrift", "halfbandwidth", "threshold", "maxfreqs")}),
                cpu_baseline=dict(vals[-1], value=v),
                e2e=dict(value=v, unit="windows/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import uwspr_b200 as ub
    from uwspr_b200.sharding import balanced_counts, gather_counts, gather_floats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    # --overlap: the generated frames, back to back, are one stream with a frame every 45 000 samples;
    # windows are taken every 22 500 samples (BASELINE.json configs[3]: 50 % sliding-window overlap),
    # so 2n-1 windows share the bytes of n and every sample crosses PCIe once
    nfr = args.windows
    stride = FL // 2 if args.overlap else FL
    nwin = 2 * nfr - 1 if args.overlap else nfr
    # Host-fed arm at N > 1: the box's host links are not equally fast when every GPU copies at once (pairs of
    # GPUs share a PCIe uplink, half of them sit behind a socket hop), so the N x nwin windows of a step are
    # split over the ranks in proportion to each rank's measured host-fed rate instead of equally; a rank
    # therefore keeps up to 1.5 x nwin windows on the host.  The device-resident arm stays at nwin per GPU.
    balance = world > 1 and not args.overlap and not args.no_balance
    nmax = int(1.5 * nwin) if balance else nwin
    ctx = ub.Context(device=local, max_windows=nmax, max_candidates=max(4 * nmax, 1024), **PARAMS)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    xs_dev, truth = gen_windows_torch(nfr, seed=rank, device=dev)
    dptr = (xs_dev.data_ptr(), nfr * FL)
    # pinned host copy for the end-to-end arm
    xs_host_t = torch.empty((nmax if balance else nfr, FL), dtype=torch.complex64, pin_memory=True)
    xs_host_t[:nfr].copy_(xs_dev)
    if balance and nmax > nfr:
        extra, truth_extra = gen_windows_torch(nmax - nfr, seed=1000 + rank, device=dev)
        xs_host_t[nfr:].copy_(extra)
        truth = list(truth) + list(truth_extra)
        del extra
    torch.cuda.synchronize(dev)
    xs_host = xs_host_t.numpy()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    stage_ms = np.zeros(4)
    totals = []

    def step_dev():
        totals.append(ctx.coarse_fine(dptr, nwin=nwin, stride=stride, fetch=False))
        stage_ms[:] += ctx.last_timing()

    e2e_out = {}

    out_bufs = ctx.result_buffers(nmax)   # pinned, allocated once (as a streaming caller would)
    n_e2e = [nwin]                        # this rank's windows per host-fed step

    def step_e2e():
        e2e_out["r"] = ctx.coarse_fine(xs_host.reshape(-1), nwin=n_e2e[0], stride=stride, out=out_bufs)

    for _ in range(args.warmup):
        step_dev()
    stage_ms[:] = 0
    l0 = ctx.launch_count()
    t_w0 = time.time()
    ms = timed(step_dev, args.steps)
    t_w1 = time.time()
    launches = ctx.launch_count() - l0
    st = stage_ms / args.steps
    ncand = totals[-1]
    for _ in range(min(args.warmup, 1) or 1):
        step_e2e()
    e2e_counts = [nwin] * world
    if balance:
        # three rounds: the rates seen under an equal split are already those of full contention for the
        # fast links; the later rounds correct the slow ones (they sped up once the fast ranks had finished)
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.synchronize(dev)
            rates = gather_floats(n_e2e[0] / (time.perf_counter() - t0), dist)
            e2e_counts = balanced_counts(world * nwin, rates, lo=max(1, nwin // 4), hi=nmax)
            n_e2e[0] = e2e_counts[rank]
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop(t_w0, t_w1) if rank == 0 else None
    npk, cands, refined, jig, soft = e2e_out["r"]
    h2d = nfr * FL * 8
    # context for the end-to-end number: the plain pinned-host -> device copy rate of this box
    scratch = torch.empty_like(xs_dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch.copy_(xs_host_t[:nfr], non_blocking=True)
    barrier()   # every rank copies at the same time, as in the host-fed steps
    ev0.record(stream)
    scratch.copy_(xs_host_t[:nfr], non_blocking=True)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    pcie_gbs = h2d / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    pcie_all = gather_floats(pcie_gbs, dist if world > 1 else None)
    del scratch
    d2h = npk.nbytes + cands.nbytes + refined.nbytes + jig.nbytes + soft.nbytes

    counts = gather_counts(nwin, dist if world > 1 else None)
    total_windows = sum(counts)
    value = total_windows * args.steps / (ms / 1e3)
    e2e_value = total_windows * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # correctness of the timed workload: decoded messages against the generator's truth
    dec = ub.decode_candidates(refined, jig, soft)
    base = np.concatenate([[0], np.cumsum(npk)])
    win_of = np.searchsorted(base, [g for g, _, _ in dec], side="right") - 1
    frame_of = (lambda w: w // 2 if w % 2 == 0 else None) if args.overlap else (lambda w: w)
    good = sum(frame_of(w) is not None and bytes(m) == bytes(truth[frame_of(w)]["msg"]) for (g, m, _), w in zip(dec, win_of))
    gated = int(refined["worth_a_try"].sum())
    evals = 12 * ncand - 2 * int((cands["m_type"] == 1).sum()) + (10 + 17) * gated  # sync_and_demodulate calls-points

    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    alg_bytes = 360000.0 * nwin + 2818.0 * ncand
    fine_ms = float(st[2])
    alg_flops = nwin * FLOP_SPEC + ncand * FLOP_COARSE + evals * FLOP_POINT
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported baseline: rank 0 at N = 1 only
        try:
            os.sched_setaffinity(0, _ALL_CPUS)  # undo the NUMA pinning: the baseline uses every core
        except Exception:
            pass
        cores = len(_ALL_CPUS) or os.cpu_count() or 1
        cpu = cpu_reference_run(xs_host, min(nwin, max(cores, 24 * cores)), cores)
    line = dict(
        metric=METRIC, value=value, unit="windows/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=("one synthetic stream per GPU, a frame every 45 000 samples, windows every 22 500 (50 % overlap, BASELINE.json configs[3] geometry)"
                              if args.overlap else
                              "10k synthetic WSPR windows swept over SNR -30..0 dB with random drift, 1 B200 (BASELINE.json configs[2]); per GPU"),
                    windows_per_gpu=nwin, input_bytes_per_gpu=h2d, l2="inputs larger than L2 (3.6 GB vs 126 MB)",
                    jiggles="all 17 per gated candidate", host_numa_node=numa, candidates=ncand, gated=gated, sync_evaluations=evals,
                    **{k: PARAMS[k] for k in ("maxdrift", "halfbandwidth", "threshold", "maxfreqs")}),
        e2e=dict(value=e2e_value, unit="windows/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=int(d2h), ms_per_step=ms_e2e / args.steps,
                 pcie_h2d_gbs=pcie_gbs, pcie_h2d_gbs_per_rank=[round(v, 2) for v in pcie_all],
                 pcie_bound_windows_per_s=sum(pcie_all) * 1e9 / (h2d / nwin),
                 windows_per_rank=e2e_counts,
                 split=("windows split over the ranks in proportion to each rank's measured host-fed rate "
                        "(uwspr_b200.sharding.balanced_counts); same total as the equal split" if balance else "equal")),
        gpu_launches=int(launches),
        stage_ms=dict(spectrogram_normalizer=float(st[0]), coarse_search=float(st[1]), fine_sync_demod=fine_ms, call=float(st[3])),
        roofline=dict(bound="hbm", kernel="k_fine (fine sync + soft symbols)", achieved=alg_bytes / (fine_ms * 1e-3) / 1e9, peak=hbm_peak,
                      unit="GB/s", frac=alg_bytes / (fine_ms * 1e-3) / 1e9 / hbm_peak, traffic=recorded_traffic(nwin),
                      peak_source="MEASURED_PEAKS.json (measured copy)" if peaks else "fallback 6650 GB/s",
                      note="the path is FP32-pipe bound, not HBM bound: see roofline_fp32"),
        roofline_fp32=dict(bound="fp32 (non-fused)", achieved=alg_flops / (ms / args.steps * 1e-3) / 1e12, peak=37.2, unit="TFLOP/s",
                           frac=alg_flops / (ms / args.steps * 1e-3) / 1e12 / 37.2, peak_fma=74.4,
                           note="whole step: algorithmic flops of the reference operation count (SURVEY 8(d)) / step time. "
                                "The reference's sums are decided bit for bit by separate fp32 mul and add roundings, so the "
                                "ceiling is one mul OR add per lane per clock: 148 SM x 128 lanes x 1.965 GHz = 37.2 T/s "
                                "(tools/fp32_pipes.cu measures 34.7-35.8); the FMA peak (74.4) is not reachable without "
                                "changing results"),
        decoded=dict(messages=len(dec), correct=int(good), windows=int(len(npk)), frames=nfr),
        clocks=clocks, cpu_baseline=cpu,
    )
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line.  Libraries (NCCL prints its version with printf) write to
    # file descriptor 1, so fd 1 is pointed at stderr and Python's sys.stdout keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=10000, help="windows per GPU")
    ap.add_argument("--ref-windows-per-core", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="host-fed arm at N > 1: equal split instead of link-rate balanced")
    ap.add_argument("--overlap", action="store_true", help="windows every 22 500 samples of one stream (not the default workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
