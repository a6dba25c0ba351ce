#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: audio-seconds / second for
log-mel + Conv1D subsampling (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path (peak -> log-mel -> 3x separable conv -> lengths -> mask) over
one batch of BASELINE.json configs[2]: 256 synthetic 16 kHz utterances of 1..15 s, zero padded to
15 s, with lengths.  Every rank processes its own batch (weak scaling, no data-path collective);
`value` is total real (un-padded) audio seconds over all ranks / max-over-ranks device time.

`value`    inputs resident in HBM before the timed region (CUDA events, max over ranks);
`e2e`      the same metric through the public API from pinned HOST buffers: H2D of the batch and
           D2H of the encoder input, mask and lengths inside the timed region, every step;
`roofline` the dominant kernel (logmel_kernel) timed live with CUDA events on its stream;
`cpu_baseline` / `--impl reference`: the reference's CPU path restated with torch CPU ops
           (oracle/torch_port.py — TensorFlow is not installable here), on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
WORKLOAD = "configs[2]: log-mel + 3x sepconv1d subsampling to d=192, batch 256 x 1..15 s padded to 15 s, lengths+masks"
BATCH = 256
N_LO, N_HI = 16000, 240000
FALLBACK_HBM_GBS = 6650.0
DISTRIBUTION = "AR(1) rho=0.97 'tilt', peak 0.5, k/32768, seeds 2+rank"


def common_config(batch: int) -> dict:
    """The `config` object both arms print (identical by construction: the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "distribution": DISTRIBUTION,
            "l2": "inputs larger than L2: 246 MB padded waveforms (128 MB of samples read) per step and two steps in "
                  "flight vs 126 MB L2; no flush needed"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=os.environ.get("TASR_BENCH_MATH", "auto"), choices=["auto", "fp32", "tf32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("TASR_CPU_SAMPLE", "256")),
                    help="utterances of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE.json configs[1]/[3]/[4] measurements")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--full-intermediates", action="store_true",
                    help="materialise the log-mel tensor and every layer's activations over the whole padded batch (A/B of the lean default)")
    ap.add_argument("--two-pass", action="store_true",
                    help="separate tasr_absmax_f32 pass before the log-mel kernel (A/B of the single-pass default)")
    ap.add_argument("--streams", type=int, default=2, help="batches in flight (CUDA-graph replays on this many streams)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_batch(rank: int, batch: int):
    from telugu_asr_b200.synth import draw_lengths, make_waveforms
    lens = draw_lengths(batch, N_LO, N_HI, seed=2 + rank)
    wav, lens = make_waveforms(lens, seed=2 + rank, dist="tilt", n_max=N_HI)
    return wav, lens


def make_weights():
    """glorot-uniform weights of the reference's shapes, bias U(-0.1,0.1), seed 7."""
    import math
    rng = np.random.default_rng(7)
    out, cin = [], 80
    for cout in (192, 384, 192):
        ldw, lpw = math.sqrt(6.0 / (9 * cin + 9)), math.sqrt(6.0 / (cin + cout))
        out.append((rng.uniform(-ldw, ldw, (9, cin)).astype(np.float32),
                    rng.uniform(-lpw, lpw, (cin, cout)).astype(np.float32),
                    rng.uniform(-0.1, 0.1, (cout,)).astype(np.float32)))
        cin = cout
    return out


# ------------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period: float = 0.004):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            self._h = h or pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU baseline (the reference's CPU path, restated) — rank 0 only
# ------------------------------------------------------------------------------------------
def _cpu_modes(cores: int):
    """(name, frontend_torch kwargs, torch intra-op threads): how the host cores can be used.  `intra_op` = one
    featurizer call at a time with multi-threaded torch kernels; `utterance_parallel` = `cores` single-threaded
    featurizer calls at a time, the reference's own parallelism (src/dataset.py:227 num_parallel_calls=AUTOTUNE),
    convs multi-threaded; `one_thread` = everything on one core."""
    return [("intra_op", dict(workers=1), cores), ("utterance_parallel", dict(workers=cores, conv_threads=cores), cores),
            ("one_thread", dict(workers=1), 1)]


def _time_cpu(fn, budget_s: float, min_passes: int = 2, max_passes: int = 200):
    times = []
    t_begin = time.perf_counter()
    while len(times) < min_passes or (time.perf_counter() - t_begin < budget_s and len(times) < max_passes):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(wav, lens, weights, n_sample: int, budget_s: float = 8.0):
    """The CPU path on a bounded sample: the first n_sample utterances of the workload in each way the host cores can be
    used (about `budget_s` seconds for the all-core modes, a quarter of the sample for the one-thread mode), plus
    BASELINE.json configs[0] (log-mel of ONE 10 s utterance).  `value` = the best all-core mode."""
    import torch
    from oracle import torch_port
    n_sample = max(1, min(n_sample, wav.shape[0]))
    cores = os.cpu_count() or 1
    modes = {}
    for name, kw, thr in _cpu_modes(cores):
        n = n_sample if thr > 1 else max(1, n_sample // 4)
        w, l = wav[:n], lens[:n]
        audio_s = float(l.sum()) / SAMPLE_RATE
        torch.set_num_threads(thr)
        torch_port.frontend_torch(w[:4], l[:4], weights, **kw)          # warm-up (thread pools, FFT plans)
        times = _time_cpu(lambda: torch_port.frontend_torch(w, l, weights, **kw), budget_s if thr > 1 else budget_s / 2)
        modes[name] = {"value": audio_s / (sum(times) / len(times)), "best_pass": audio_s / min(times), "threads": thr,
                       "utterances": int(n), "passes": len(times), "cpu_seconds": round(sum(times), 2)}
    # configs[0]: the reference's own CPU-runnable case, log-mel of one synthetic 10 s utterance
    from telugu_asr_b200.synth import make_waveforms
    w10, _ = make_waveforms([160000], seed=0, dist="tilt")
    x10 = torch.from_numpy(w10[0])
    cfg0 = {}
    for thr in (1, cores):
        torch.set_num_threads(thr)
        torch_port.logmel_torch(x10)
        times = _time_cpu(lambda: torch_port.logmel_torch(x10), 1.0, min_passes=5, max_passes=400)
        cfg0[f"threads_{thr}"] = {"ms": 1e3 * statistics.median(times), "audio_s_per_s": 10.0 / statistics.median(times)}
    torch.set_num_threads(cores)
    best = max(("intra_op", "utterance_parallel"), key=lambda k: modes[k]["value"])
    return {"value": modes[best]["value"], "unit": "audio-seconds/s", "cores": cores, "kind": "port", "mode": best,
            "modes": modes, "config0_logmel_1x10s": cfg0,
            "sample": f"first {n_sample} utterances of the workload, repeated for ~{budget_s:.0f} s per all-core mode "
                      f"({n_sample // 4 or 1} utterances for the one-thread mode): per-utterance featurizer calls + zero-pad "
                      "collate + 3 separable convs, torch CPU ops (oracle/torch_port.py); TensorFlow reference not "
                      "installable offline"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (restated; see module doc), with all the host
    threads it can use - the faster of the two all-core modes of _cpu_modes, chosen in the warm-up."""
    if rank != 0:
        return
    wav, lens = make_batch(0, args.batch)
    weights = make_weights()
    import torch
    from oracle import torch_port
    cores = os.cpu_count() or 1
    n_sample = max(1, min(args.cpu_sample, args.batch))
    warm = max(1, args.warmup)
    first, best = None, None
    for name, kw, thr in _cpu_modes(cores)[:2]:
        torch.set_num_threads(thr)
        torch_port.frontend_torch(wav[:4], lens[:4], weights, **kw)                 # thread pools, FFT plans
        t0 = time.perf_counter()
        torch_port.frontend_torch(wav[:n_sample], lens[:n_sample], weights, **kw)
        dt = time.perf_counter() - t0
        if first is None or dt < first:
            first, best = dt, (name, kw, thr)
    name, kw, thr = best
    torch.set_num_threads(thr)
    # EXACTLY K timed steps (and W warm-ups); each step is a bounded sample of the workload, sized so that the whole
    # run stays within about two minutes of CPU time: the first n utterances of the batch, all host threads.
    budget = 120.0
    if first * (args.steps + warm) > budget:
        n_sample = max(4, int(n_sample * budget / (first * (args.steps + warm))))
    w, l = wav[:n_sample], lens[:n_sample]
    audio_s = float(l.sum()) / SAMPLE_RATE
    for _ in range(warm):
        torch_port.frontend_torch(w, l, weights, **kw)
    steps = args.steps
    t0 = time.perf_counter()
    for _ in range(steps):
        torch_port.frontend_torch(w, l, weights, **kw)
    dt = (time.perf_counter() - t0) / steps
    val = audio_s / dt
    sample = (f"each step = first {n_sample} of {args.batch} utterances ({audio_s:.0f} audio-s), sized so that K steps + W "
              f"warm-ups take about two minutes of CPU time at most; mode {name} ({cores} host threads)")
    line = {
        "impl": "reference", "metric": "audio-seconds/s, log-mel + conv1d subsampling", "value": val,
        "unit": "audio-seconds/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args.batch),
        "sample": sample,
        "cpu_baseline": {"value": val, "unit": "audio-seconds/s", "cores": cores, "kind": "port", "mode": name,
                         "sample": sample + "; torch CPU restatement of the TensorFlow path (oracle/torch_port.py)"},
        "e2e": {"value": val, "unit": "audio-seconds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# BASELINE.json configs[1], [3], [4]: CUDA-event numbers beside the headline (configs[2])
# ------------------------------------------------------------------------------------------
def measure_configs(fe, dev, rank, world, dist, weights):
    """configs[1]: log-mel only, 64 x 10 s, on every GPU (per-GPU figure from the slowest rank).
    configs[3]: 1024 x 30 s log-mel + subsampling, STRONG split 1024/N per rank.
    configs[4]: corner and centre points of the length x batch sweep, strong split of the batch over the ranks.
    Waveforms are made on the device (eight seeded 'tilt' utterances tiled with per-utterance scale factors: making
    4096 x 30 s on the host would take minutes); every point runs as one CUDA-graph replay per step (the launch mode of
    the headline), CUDA events around >= 20 ms of back-to-back steps after two warm-ups, max over ranks."""
    import torch
    import telugu_asr_b200 as tasr
    from telugu_asr_b200.synth import make_waveforms
    peak, _ = measured_peak()
    base_np, _ = make_waveforms([480000] * 8, seed=3, dist="tilt")
    base = torch.from_numpy(base_np).to(dev)

    def batch_of(n_utt, n_samples, first):
        idx = (torch.arange(n_utt, device=dev) + first)
        w = base[idx % 8, :n_samples] * (1.0 - 0.002 * (idx % 17).to(torch.float32))[:, None]
        return w.contiguous(), torch.full((n_utt,), n_samples, dtype=torch.int32, device=dev)

    def reduce_max(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def time_steps(fn, sync):
        for _ in range(2):
            fn()
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); sync()
        n = int(min(200, max(3, 20.0 / max(a.elapsed_time(b), 1e-3))))
        if dist is not None:
            tn = torch.tensor([n], dtype=torch.int64, device=dev)
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
            n = int(tn[0])
            dist.barrier()
        sync()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        sync()
        return a.elapsed_time(b) / n, n

    sync = torch.cuda.synchronize
    out = {}
    # configs[1] -- the reference-signature featurizer (two passes: peak, then log-mel in the reference's op order)
    feat = fe.featurizer
    w, l = batch_of(64, 160000, 0)
    fbuf = torch.empty((64, 998, 80, 1), dtype=torch.float32, device=dev)
    ms, n = time_steps(lambda: feat.featurize_batch(w, l, out=fbuf, t_max=998), sync)
    ms = reduce_max(ms)
    alg = 64 * (4 * 160000 + 320 * 998)
    out["configs[1]"] = {"workload": "log-mel only, 64 x 10 s per GPU, SpeechFeaturizer two-pass path (absmax + log-mel kernel)",
                         "ms_per_step": ms, "steps": n, "audio_s_per_s_per_gpu": 640.0 / (ms * 1e-3),
                         "algorithmic_bytes": alg, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "note": "61 MB per step: the batch stays L2-resident between steps (126 MB L2), stated not flushed"}
    del fbuf

    def frontend_point(total_utt, seconds):
        """total_utt utterances of `seconds` s split over the ranks (rank r takes a contiguous block)."""
        per = total_utt // world + (1 if rank < total_utt % world else 0)
        first = rank * (total_utt // world) + min(rank, total_utt % world)
        n_s = 16000 * seconds
        ms_local = 0.0
        nsteps = 0
        if per > 0:
            w_, l_ = batch_of(per, n_s, first)
            cap_ = tasr.CapturedFrontEnd(fe, per, n_s, dev)
            cap_.load(w_, l_)
            del w_
            sync()
        if dist is None:
            ms_local, nsteps = time_steps(cap_.replay, sync)
        else:
            # every rank must take part in the collectives inside time_steps, with or without work
            ms_local, nsteps = time_steps(cap_.replay if per > 0 else (lambda: None), sync)
            if per == 0:
                ms_local = 0.0
        ms_ = reduce_max(ms_local)
        if per > 0:
            del cap_
        torch.cuda.empty_cache()
        T = 1 + (n_s - 400) // 160
        t3 = T
        for _ in range(3):
            t3 = (t3 - 9) // 2 + 1
        alg_ = total_utt * (4 * n_s + 4 * 192 * t3 + 4 * t3)
        return {"utterances": total_utt, "seconds_each": seconds, "ms_per_step": ms_, "steps": nsteps,
                "audio_s_per_s": total_utt * seconds / (ms_ * 1e-3), "audio_s_per_s_per_gpu": total_utt * seconds / (ms_ * 1e-3) / world,
                "pipeline_hbm_frac": alg_ / world / (ms_ * 1e-3) / 1e9 / peak}

    out["configs[3]"] = dict(frontend_point(1024, 30), workload=f"log-mel + subsampling, 1024 x 30 s, strong split {1024 // world} per GPU over {world} GPU(s)",
                             scaling="strong")
    pts = []
    for seconds, total in ((1, 1), (1, 4096), (30, 1), (30, 4096), (10, 256)):
        pts.append(frontend_point(total, seconds))
    out["configs[4]"] = {"workload": f"length x batch sweep corners + centre, batch split over {world} GPU(s) (strong), vs cpu_baseline",
                         "points": pts}
    return out


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # plain `python bench.py --gpus N`: re-launch one rank per GPU
        port = 29500 + (os.getpid() % 2000)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__), *sys.argv[1:]])
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import telugu_asr_b200 as tasr
    from telugu_asr_b200 import _native

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    lib = _native.lib()
    wav_np, lens_np = make_batch(rank, args.batch)
    weights = make_weights()
    audio_s = float(lens_np.sum()) / SAMPLE_RATE
    max_len = int(lens_np.max())

    math_mode = args.math
    fe = None
    if math_mode in ("auto", "tf32"):
        try:
            fe = tasr.FrontEnd(math="tf32", lean_intermediates=not args.full_intermediates, single_pass=not args.two_pass)
            fe.set_weights(weights, dev)
            fe.subsampling._ensure_plans()
            math_mode = "tf32"
        except NotImplementedError:
            if math_mode == "tf32":
                raise
            fe = None
    if fe is None:
        math_mode = "fp32"
        fe = tasr.FrontEnd(math="fp32")
        fe.set_weights(weights, dev)

    wav = torch.from_numpy(wav_np).to(dev)
    lens = torch.from_numpy(lens_np).to(dev)
    stream = torch.cuda.current_stream()

    def eager_step():
        return fe(wav, lens, max_length=max_len)

    cap, graph_note = None, None
    if not args.no_graph:
        # static shape [B, N_max]; lengths stay on the device; consecutive steps alternate between the slots/streams
        try:
            cap = tasr.InterleavedFrontEnd(fe, args.batch, wav.shape[1], dev, n_streams=max(1, args.streams))
            for i in range(len(cap.slots)):
                cap.load(i, wav, lens)
            cap.join()
            torch.cuda.synchronize()
        except Exception as e:   # the same kernels, launched one by one: a slower but valid number rather than none
            cap, graph_note = None, f"CUDA-graph capture failed ({e!r}); kernel-by-kernel launches"
            torch.cuda.synchronize()
    step_no = [0]

    def step():
        if cap is None:
            return eager_step()
        i = step_no[0]
        step_no[0] += 1
        return cap.replay(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    # The clock sampler runs over a pre-roll, the timed region and a post-roll of the SAME steps (>= 30 ms of load on
    # either side): K steps alone can be a few milliseconds, less than one NVML sample.
    for _ in range(args.warmup):
        out = step()
    barrier()
    t_probe0 = time.perf_counter()
    for _ in range(8):
        out = step()
    if cap is not None:
        cap.join()
    torch.cuda.synchronize()
    est_ms = max(1e-3, (time.perf_counter() - t_probe0) * 1e3 / 8)
    roll = int(min(400, max(4, 30.0 / est_ms)))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    for _ in range(roll):
        out = step()
    if cap is not None:
        cap.join()
    l0 = lib.tasr_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if cap is not None:
        cap.fork()                      # the slot streams start after e0 ...
    for _ in range(args.steps):
        out = step()
    if cap is not None:
        cap.join()                      # ... and e1 is recorded after all of them
    e1.record()
    launches = int(lib.tasr_launch_count() - l0)
    for _ in range(roll):
        step()
    if cap is not None:
        cap.join()
    barrier()
    clocks = sampler.stop()
    clocks["window"] = f"{roll} untimed steps + the {args.steps} timed steps + {roll} untimed steps, sampled every 4 ms"
    if cap is not None:
        launches = cap.kernels_per_replay * args.steps      # replayed kernels do not pass through the C ABI counter
    ms_total = e0.elapsed_time(e1)

    # ---- per-stage device times (outside the timed region): CUDA events between the launches ----
    _native.stage_marks = []
    n_stage_steps = 20
    for _ in range(n_stage_steps):
        eager_step()
    torch.cuda.synchronize()
    marks, _native.stage_marks = _native.stage_marks, None
    stage_us = {}
    for (n0_, e0_), (n1_, e1_) in zip(marks[:-1], marks[1:]):
        if n1_ != "begin":
            stage_us.setdefault(n1_, []).append(e0_.elapsed_time(e1_) * 1e3)
    stage_us = {k: statistics.median(v) for k, v in stage_us.items()}

    # ---- end to end: pinned host -> H2D -> path -> D2H, every step -------------------------
    # (a) the public host-to-host API (telugu_asr_b200.FrontEndPipeline): the batch leaves pinned host
    #     memory as the ragged int16 PCM the reference's loader decodes (data_util.py:31), is unpacked,
    #     featurised and subsampled on the device and comes back as [B,T3,d] + mask + lengths;
    #     copy-in / compute / copy-out of consecutive steps overlap on three streams.
    from telugu_asr_b200.synth import to_pcm16
    utts = [to_pcm16(wav_np[b, : lens_np[b]]) for b in range(args.batch)]
    e2e_steps = max(3, min(args.steps, 100))

    def run_pipe(packed_output):
        pipe_ = tasr.FrontEndPipeline(fe, args.batch, wav_np.shape[1], dev, pcm16=True, slots=2, packed_output=packed_output)
        for s_ in range(2):
            pipe_.stage(s_, utts)           # host packing is the loader's job: outside the timed region
        for i_ in range(8):
            tk_ = pipe_.submit(i_ % 2)
        tk_.wait()
        barrier()
        return pipe_

    # the padded return ([B, T3, 192] + mask, the reference's tensor shapes) first, as the secondary figure ...
    pipe = run_pipe(False)
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for i in range(e2e_steps):
        tk = pipe.submit(i % 2)
    pipe.drain()
    q1.record()
    barrier()
    padded_return = {"ms_per_step": q0.elapsed_time(q1) / e2e_steps, "d2h_bytes_per_step": int(pipe.d2h_bytes)}
    h_o, h_m, h_l = tk.wait()
    enc, mask, len3 = out
    padded_ok = bool(torch.equal(h_o, enc.cpu()) and torch.equal(h_m, mask.cpu()) and torch.equal(h_l, len3.cpu()))
    del pipe
    # ... then the headline: only the valid rows come back ([sum(len3), 192] + offsets + len3)
    pipe = run_pipe(True)
    e2e_windows = []
    for w_ in range(3 if os.environ.get("TASR_E2E_WINDOWS") else 1):
        n_alloc0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        l0 = lib.tasr_launch_count()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        g0.record()
        for i in range(e2e_steps):
            tk = pipe.submit(i % 2)
        t_host1 = time.perf_counter()
        pipe.drain()                            # last D2H has landed
        g1.record()
        barrier()
        e2e_ms = g0.elapsed_time(g1)
        e2e_launches = int(lib.tasr_launch_count() - l0)
        if pipe.graph:                          # replayed kernels do not pass through the C ABI counter
            e2e_launches = int(pipe.kernels_per_submit) * e2e_steps
        e2e_windows.append({"ms_per_step": e2e_ms / e2e_steps, "host_submit_ms_per_step": (t_host1 - t_host0) * 1e3 / e2e_steps,
                            "cudaMallocs": int(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - n_alloc0)})
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    h_rows, h_offs, h_l = tk.wait()
    enc_c, l3_c = enc.cpu(), len3.cpu()
    e2e_ok = bool(padded_ok and torch.equal(h_l, l3_c) and int(h_offs[-1]) == int(l3_c.clamp(min=0).sum()) and
                  all(torch.equal(h_rows[int(h_offs[b_]): int(h_offs[b_ + 1])], enc_c[b_, : max(int(l3_c[b_]), 0)]) for b_ in range(args.batch)))
    # what the host link gives this rank while every rank uses it: pinned H2D of one step's bytes, timed alone
    probe = torch.empty((int(h2d),), dtype=torch.uint8).pin_memory()
    probe_d = torch.empty((int(h2d),), dtype=torch.uint8, device=dev)
    for _ in range(2):
        probe_d.copy_(probe, non_blocking=True)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(10):
        probe_d.copy_(probe, non_blocking=True)
    r1.record()
    barrier()
    link_gbs = h2d * 10 / (r0.elapsed_time(r1) * 1e-3) / 1e9

    # the first encoder block on what the path delivers: a `stages` entry, outside the headline
    enc_block_us = None
    try:
        blk = tasr.EncoderBlock(input_dim=192, num_heads=6, head_dim=32, fc_factor=1)
        blk.build(dev, seed=3)
        blk.prepare(enc.shape[1])
        for _ in range(3):
            blk(enc, lengths=len3)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(20):
            blk(enc, lengths=len3)
        b1.record()
        torch.cuda.synchronize()
        enc_block_us = b0.elapsed_time(b1) / 20 * 1e3
    except Exception as e:      # never a reason to lose the headline
        enc_block_us = f"failed: {e!r}"

    # (b) the same without the ingest work: padded float32 batch, one stream, no overlap
    pb = tasr.PinnedBatch(args.batch, wav_np.shape[1], dev)
    pb.host_wav.copy_(torch.from_numpy(wav_np))
    pb.host_len.copy_(torch.from_numpy(lens_np))
    h_out = torch.empty(enc.shape, dtype=enc.dtype).pin_memory()
    h_mask = torch.empty(mask.shape, dtype=mask.dtype).pin_memory()
    h_len = torch.empty(len3.shape, dtype=len3.dtype).pin_memory()

    def e2e_step_padded():
        w, l = pb.to_device(non_blocking=True)
        o, m, l3 = fe(w, l, max_length=max_len)
        h_out.copy_(o, non_blocking=True)
        h_mask.copy_(m, non_blocking=True)
        h_len.copy_(l3, non_blocking=True)

    pad_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step_padded()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(pad_steps):
        e2e_step_padded()
    p1.record()
    barrier()
    pad_ms = p0.elapsed_time(p1) / pad_steps

    # ---- BASELINE.json configs[1], [3], [4] (outside the headline timed region, like `stages`) -----------------
    configs = None
    if not args.no_configs:
        configs = measure_configs(fe, dev, rank, world, dist if world > 1 else None, weights)

    # ---- validation (outside every timed region): NCCL gathers results, nothing on the data path -----
    # Every rank runs a small COMMON batch (same seed) and its own shard of a split batch; the per-utterance
    # checksums (integer sums of the float bit patterns) and lengths are all_gathered to every rank, and rank 0
    # checks that all GPUs produced identical bits and that the union of the shards equals the unsharded run.
    validation = None
    if world > 1:
        from telugu_asr_b200.synth import draw_lengths as _dl, make_waveforms as _mw
        vl = _dl(16 * world, 8000, 64000, seed=99)
        vw, vl = _mw(vl, seed=99, dist="tilt")

        def _run(idx):
            nm = -(-int(vl[idx].max()) // 4) * 4
            o, m, l3 = fe(torch.from_numpy(np.ascontiguousarray(vw[idx][:, :nm])).to(dev), torch.from_numpy(vl[idx]).to(dev),
                          max_length=int(vl[idx].max()))
            cs = torch.stack([o[j, : int(l3[j])].contiguous().view(torch.int32).to(torch.int64).sum() for j in range(len(idx))])
            return cs, l3.to(torch.int64)

        all_idx = np.arange(len(vl))
        cs_all, l3_all = _run(all_idx)                                   # the whole batch on this GPU
        mine = np.asarray(tasr.shard_by_length(vl, world)[rank])
        cs_mine, l3_mine = _run(mine)                                    # this rank's shard
        full = torch.zeros((len(vl), 2), dtype=torch.int64, device=dev)
        full[torch.from_numpy(mine).to(dev), 0] = cs_mine
        full[torch.from_numpy(mine).to(dev), 1] = l3_mine
        dist.all_reduce(full, op=dist.ReduceOp.SUM)                      # union of the shards (disjoint rows)
        gathered = [torch.zeros_like(cs_all) for _ in range(world)]
        dist.all_gather(gathered, cs_all)                                # every GPU's result for the common batch
        same_bits = all(bool(torch.equal(g_, cs_all)) for g_ in gathered)
        union_ok = bool(torch.equal(full[:, 0], cs_all) and torch.equal(full[:, 1], l3_all))
        validation = {"collective": "nccl all_gather / all_reduce of per-utterance checksums and lengths (validation only)",
                      "utterances": int(len(vl)), "all_gpus_bit_identical": same_bits, "shard_union_equals_unsharded": union_ok}

    # ---- reduce over ranks: max time, sum of audio ------------------------------------------
    t = torch.tensor([ms_total, e2e_ms, audio_s, float(h2d), float(d2h), float(launches), -float(link_gbs)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax, tsum = t, t
    ms_total_max, e2e_ms_max = float(tmax[0]), float(tmax[1])
    audio_total = float(tsum[2])

    if rank == 0:
        value = audio_total * args.steps / (ms_total_max * 1e-3)
        e2e_val = audio_total * e2e_steps / (e2e_ms_max * 1e-3)
        # roofline of the dominant kernel: algorithmic bytes = read each valid sample once + write
        # each valid log-mel row once (4*N + 4*80*T per utterance)
        T = np.maximum(0, 1 + (lens_np.astype(np.int64) - 400) // 160)
        alg_bytes = float((4 * lens_np.astype(np.int64) + 320 * T).sum())
        peak, peak_src = measured_peak()
        # per-stage algorithmic bytes: valid samples / valid rows only for the ragged stages
        nvalid = [T]
        for _ in range(3):
            nvalid.append(np.maximum(0, (nvalid[-1] - 9) // 2 + 1))
        t_pad = [1498, 745, 369, 181]
        ch = [80, 192, 384, 192]
        stage_bytes = {"absmax_kernel": float(4 * lens_np.astype(np.int64).sum()), "logmel_kernel": alg_bytes}
        lean = bool(getattr(fe, "lean_intermediates", False)) and math_mode == "tf32"
        for i in range(3):   # x read once where valid + y written once (lean intermediates: where valid; else the whole padded tensor)
            rows_out = int(nvalid[i + 1].sum()) if (lean and i < 2) else args.batch * t_pad[i + 1]
            stage_bytes[f"sepconv_layer{i + 1}"] = float(4 * (int(nvalid[i].sum()) * ch[i] + rows_out * ch[i + 1]))
        for k in list(stage_us):
            if k.startswith("logmel") and k not in stage_bytes:
                stage_bytes[k] = alg_bytes
        stages = {k: {"us": round(v, 1), "gbs": (round(stage_bytes[k] / (v * 1e-6) / 1e9, 1) if k in stage_bytes else None),
                      "hbm_frac": (round(stage_bytes[k] / (v * 1e-6) / 1e9 / peak, 3) if k in stage_bytes else None)}
                  for k, v in stage_us.items()}
        # the dominant kernel = the stage with the largest CUDA-event time
        dom = max((k for k in stage_us if k in stage_bytes), key=lambda k: stage_us[k])
        k_ms = stage_us[dom] * 1e-3
        dom_bytes = stage_bytes[dom]
        achieved = dom_bytes / (k_ms * 1e-3) / 1e9
        traffic, ncu_lm, tj = None, None, {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tj = json.load(fh)
            ncu_lm = next((v for k, v in tj.items() if k.startswith(dom) and isinstance(v, dict)), None)
            traffic = ncu_lm.get("dram_bytes_per_launch") if ncu_lm else None
        except Exception:
            pass
        # the whole step against the contract of SURVEY.md 8(d): the waveform read once, [B,T3,192] + mask written once
        t3 = nvalid[3]
        pipe_bytes = float((4 * lens_np.astype(np.int64) + 4 * 192 * t3 + 4 * t3).sum())
        ms_step = ms_total_max / args.steps
        line = {
            "metric": "audio-seconds/s, log-mel + conv1d subsampling", "value": value, "unit": "audio-seconds/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(args.batch),
            "implementation": {
                       "audio_seconds_per_step_per_gpu": audio_s,
                       "pointwise_math": math_mode + (" (tcgen05 kind::tf32, fp32 accumulate)" if math_mode == "tf32" else " (CUDA-core FMA)"),
                       "logmel_math": getattr(fe.featurizer, "logmel_math_used", "fp32 CUDA cores"),
                       "parallelism": f"dp{world} by utterance, no data-path collective",
                       "featurizer": ("single pass: the log-mel kernel finds max|x| while it stages the samples and the first separable conv "
                                      "applies 2 log(gain) and the floor (float32-rounding differences only, tests/test_gpu_parity.py); "
                                      "no separate peak pass") if getattr(fe, "single_pass", False) and math_mode == "tf32"
                                     else "two passes: tasr_absmax_f32, then the log-mel kernel",
                       "intermediates": ("lean: the log-mel tensor and the activations of layers 1-2 are not written far inside the "
                                         "collate padding (no kernel reads them there); encoder input, mask and lengths are bit-identical "
                                         "to the fully materialised run (tests/test_gpu_parity.py)") if lean else "fully materialised",
                       "launch": (f"one CUDA-graph replay per step, {len(cap.slots)} steps in flight on {len(cap.slots)} streams "
                                  "(telugu_asr_b200.InterleavedFrontEnd); roofline.kernel_ms and `stages` are CUDA-event times "
                                  "of the same kernels launched one by one on one stream right after the timed region")
                                 if cap is not None else (graph_note or "kernel-by-kernel launches")},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": k_ms,
                         # the whole step against SURVEY.md 8(d)'s pipeline contract (waveform in, encoder input + mask out)
                         "pipeline_algorithmic_bytes_per_step": pipe_bytes,
                         "pipeline_achieved": pipe_bytes / (ms_step * 1e-3) / 1e9,
                         "pipeline_frac": pipe_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         # share of the step's kernel time (sum of the single-stream stage times: comparable with the ncu
                         # launch list, which serialises the kernels); the timed step overlaps two batches, so kernel_ms
                         # divided by ms_per_step would overstate it
                         "kernel_share_of_step": (k_ms * 1e3) / max(sum(stage_us.values()), 1e-9),
                         "kernel_ms_over_ms_per_step": k_ms / (ms_total / args.steps),
                         "ncu": ({"source": tj.get("source"), **{kk: ncu_lm.get(kk) for kk in ("dram_pct", "issue_active_pct", "fma_pipe_pct", "tensor_pipe_pct", "warp_instructions", "registers", "warps_active_pct")}}
                                 if ncu_lm else None),
                         "note": "frac = the dominant kernel's algorithmic bytes (SURVEY.md 8d: what it must read and write once) / its "
                                 "CUDA-event time / the measured copy bandwidth; pipeline_frac = the same for the whole step against the "
                                 "fused contract (waveform read once, encoder input + mask written once: 73.6 kB per audio-second); "
                                 "kernel_ms is the kernel's single-stream time, the timed step overlaps two batches; see DESIGN.md 4"},
            "e2e": {"value": e2e_val, "unit": "audio-seconds/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_ms_max / e2e_steps, "gpu_launches": e2e_launches,
                    "matches_device_resident_result": e2e_ok, "windows": e2e_windows,
                    "pcie_h2d_gbs": h2d / (e2e_ms_max / e2e_steps * 1e-3) / 1e9,
                    # pinned-memory H2D of one step's bytes, every rank at once and nothing else running: the slowest rank's rate
                    "host_link_ceiling_gbs": -float(tmax[6]),
                    "host_link_frac": (h2d / (e2e_ms_max / e2e_steps * 1e-3) / 1e9) / max(-float(tmax[6]), 1e-9),
                    "padded_return": {**padded_return, "value": audio_s / (padded_return["ms_per_step"] * 1e-3),
                                      "note": "rank 0; same pipeline returning the zero-padded [B,T3,192] + mask + len3"},
                    "api": "telugu_asr_b200.FrontEndPipeline(packed_output=True).submit: ragged int16 PCM in pinned host memory (valid "
                           "samples only) -> H2D -> unpack -> log-mel (single pass) -> 3x sepconv -> lengths/mask -> valid rows packed "
                           "-> D2H of [sum(len3),192] f32 + len3 (offsets are host arithmetic on the staged lengths); "
                           "double-buffered slots; the device side of a submit is one CUDA-graph launch per slot",
                    "f32_padded_single_stream": {"value": audio_s / (pad_ms * 1e-3), "ms_per_step": pad_ms,
                                                 "h2d_bytes_per_step": int(pb.h2d_bytes), "note": "rank 0; padded float32 [B,N_max] H2D, no overlap"}},
            "stages": stages,
            "next_stage": {"encoder_block": {"us": enc_block_us, "what": "EncoderBlock (MHSA 6x32 with RoPE + padding mask, FFN, two LayerNorms; "
                                             "tcgen05 TF32 dense layers, FP32 attention) on this batch's [B,T3,192] + len3, eager launches, rank 0"}},
            "configs": configs,
            "validation": validation,
            "gpu_launches": int(float(tsum[5])) if world > 1 else launches,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(wav_np, lens_np, weights, args.cpu_sample)
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "audio-seconds/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
