#!/usr/bin/env python
"""bench.py -- TDoA hypercubes scored/sec (SRP-PHAT + shift-stack), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--sub-batches S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The headline workload is config C2 of SURVEY.md section 8d (7 mics, 5 speakers, 3 s @ 48 kHz, one desk geometry,
G ~ 2e4 hypercubes).  One STEP = S sub-batches of B synthetic mixtures per GPU (default 12 x 64 = 768 mixtures, so
the timed region lasts ~0.5 s); every sub-batch runs the complete device path
    asw_srp_score (STFT+PHAT+cross-spectra -> GCC lag tables -> SRP gather, max over windows)
    asw_map_topk  (MAX_POWER and the K best hypercubes per mixture)
    asw_peaks_find (fill_powermap + find_valid_peak_new: thresholded 3-D local maxima -> peak hypercubes)
    asw_select_patches + asw_build_shift_table (local_source_adaptive -> dense per-sub-batch patch table)
    asw_shift_stack of every coarse hypercube patch of every mixture into a ring of (128, M, T) network-input
                  buffers, 9 of them (1152 patches) per launch
with scoring / pruning / stacking of consecutive sub-batches software-pipelined on three streams.  Every sub-batch
selects its own patches on the device.  `value` starts with inputs resident in HBM; `e2e` starts from pinned host
buffers (16-bit PCM, what the data is) and ends with the maps / top-K / patch lists back on the host.

The same JSON line carries the other BASELINE configs as sub-objects, each with its CPU (oracle port) time:
    c1  one mixture through the drop-in Mic_Array.Apply_SRP_PHAT (host tensor in, Patch list out), 44.1 kHz
    c3  fine stage: 64 mixtures -> device pruning -> asw_subdivide -> fused shift-stack + normalize_input
    c5  16 mics, 10 s clips, dense grid (G ~ 1e5): scoring stages
    hypercube_sharded (N > 1)  C5 geometry, hypercubes sharded over the ranks, GCC-table all-gather + top-K merge
Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tdoa_hypercubes_scored_per_sec"
UNIT = "hypercubes/s"
WORKLOAD = "C2: coarse width-8 (half-width 4) hypercube SRP-PHAT + Spotform_Big_Patch shift-stack, 7 mics, 5 speakers, 3 s @ 48 kHz"
N_MICS, N_SPK, T_SAMPLES, FS = 7, 5, 144000, 48000
GEOM_SEED = 1
C1_FS, C1_T, C1_SPK = 44100, 132300, 3            # BASELINE configs[0]
C5_MICS, C5_T, C5_SPK = 16, 480000, 4             # BASELINE configs[4]
MAX_LAG = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="mixtures per GPU per sub-batch")
    ap.add_argument("--sub-batches", type=int, default=12, help="sub-batches per step (a step = sub-batches x batch mixtures)")
    ap.add_argument("--unique-batches", type=int, default=4, help="distinct resident sub-batches the step cycles through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip", default="", help="comma list of sub-benchmarks to skip: c1,c3,c5,sharded,e2e,variants")
    ap.add_argument("--launch-batches", type=int, default=9,
                    help="128-patch network batches stacked by one shift-stack launch (ring of that many buffers)")
    ap.add_argument("--c3-group", type=int, default=16, help="mixtures per group of the c3 fine-stage pipeline")
    ap.add_argument("--count-sync", type=int, default=1,
                    help="1: the host reads each sub-batch's patch count (4 bytes, behind the pruning of the NEXT "
                         "sub-batch's scoring) and launches exactly the patches selected; 0: capacity-sized counted "
                         "launches, rows beyond the device count exit on the device")
    ap.add_argument("--stft-path", default="auto", choices=["auto", "split", "generic"],
                    help="kernels of the STFT + PHAT + cross-spectra stage (auto = fused register kernel for <= 8 mics)")
    ap.add_argument("--streams", type=int, default=3, choices=[1, 3],
                    help="3: scoring (SM/shared-memory bound), pruning (latency bound, B CTAs) and shift-stack (HBM "
                         "bound) of consecutive sub-batches run on their own streams and overlap; 1: fully serial")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def source_hash(*names):
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "acousticswarms_speech_b200", "csrc", n), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every few ms over the timed regions (NVML in a thread)."""

    def __init__(self, gpu_index=0, period_s=0.004):
        import threading
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._ok = True
        except Exception:
            return
        self._period = period_s
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self._period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if not self._ok:
            return out
        self._stop.set()
        self._t.join(timeout=2)
        if self.samples:
            out["sm_mhz"] = statistics.median(self.samples)
            out["sm_mhz_min"] = min(self.samples)
            out["samples"] = len(self.samples)
        out["reasons"] = sorted(self.reasons)
        return out


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's CPU path; the only place bench.py touches oracle/)
def pcm_content(x):
    """16-bit PCM content (what the reference's PCM_16 wav datasets hold) carried as float32 = pcm / 32768."""
    return (np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0).astype(np.float32)


def cpu_reference_setup(n_mics=N_MICS, fs=FS, seed=GEOM_SEED):
    from acousticswarms_speech_b200 import synth
    from acousticswarms_speech_b200.constants import freq_bins, n_fft
    from oracle import cpu_reference, geometry_oracle
    scene = synth.desk_array(n_mics, np.random.default_rng(seed), fs)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, FS=fs)
    ref = cpu_reference.CpuReferencePath(geo, freq_bins, fs, n_fft)
    return scene, geo, ref


def cpu_reference_step(ref, mix):
    """One mixture through the reference's CPU path: Apply_SRP_PHAT + the coarse shift loop."""
    patches, m = ref.apply_srp_phat(mix)
    ref.shift_stack(mix, patches)
    return patches


def run_reference(args, rank):
    if rank != 0:
        return
    from acousticswarms_speech_b200 import synth
    scene, geo, ref = cpu_reference_setup()
    G = geo.grids.shape[0]
    steps = max(1, args.steps)
    mixes = [pcm_content(synth.mixture(scene, N_SPK, T_SAMPLES, seed=100 + i)) for i in range(min(steps, 8))]
    for i in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(ref, mixes[0])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_reference_step(ref, mixes[i % len(mixes)])
    dt = time.perf_counter() - t0
    v = G * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mixtures_per_step": 1, "hypercubes": G,
                       "note": "reference CPU path (oracle port of SRP_Map_WINDOW_torch + pruning + shift loop), "
                               "bounded sample: one mixture of the workload per step (its API has no batch axis), "
                               "steering table precomputed (setup excluded, README.md:144)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.threads, "kind": "port",
                             "sample": f"{steps} mixtures, 1 per step, Apply_SRP_PHAT + coarse shift loop"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class C2Pipeline:
    """The C2 device path for one sub-batch, software-pipelined over three streams (see module docstring)."""

    def __init__(self, args, node, fe, dev, B, M, T, G, cap, world):
        import torch
        self.torch = torch
        self.args, self.node, self.fe, self.dev = args, node, fe, dev
        self.B, self.M, self.T, self.G, self.cap, self.world = B, M, T, G, cap, world
        self.K = fe.topk
        self.MAXPATCH = node.native_select.max_patches
        self.MAXP = node.native_peaks.max_peaks
        self.tables = [(torch.zeros((cap, M), device=dev, dtype=torch.int32),
                        torch.zeros((cap,), device=dev, dtype=torch.int32),
                        torch.zeros((1,), device=dev, dtype=torch.int32)) for _ in range(2)]
        self.table_free = [None, None]
        self.maps = [torch.empty((B, G), device=dev) for _ in range(2)]
        self.map_free = [None, None]
        self.step_no = 0
        self.sel_pin = torch.empty((B, self.MAXPATCH * (M + 1) + 1), dtype=torch.int32).pin_memory()
        self.map_pin = torch.empty((B, G), dtype=torch.float32).pin_memory()
        self.val_pin = torch.empty((B, self.K), dtype=torch.float32).pin_memory()
        self.idx_pin = torch.empty((B, self.K), dtype=torch.int32).pin_memory()
        self.peaks_pin = torch.empty((B, self.MAXP), dtype=torch.int32).pin_memory()
        self.count_pin = torch.empty((B,), dtype=torch.int32).pin_memory()
        self.ntot_pin = [torch.zeros((1,), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.d2h_bytes = int(B * G * 4 + B * self.K * 8 + B * self.MAXP * 4 + B * 4 + self.sel_pin.numel() * 4)
        # the one collective of the mixture-sharded path: all ranks' (value, index) top-K lists, packed into a single
        # all-gather per sub-batch and issued on a side stream so it overlaps the shift-stack
        self.collective = world > 1
        self.NPACK = 4
        self.packs = [torch.empty((B, 2 * self.K), device=dev) for _ in range(self.NPACK)] if world > 1 else None
        self.pack_free = [None] * self.NPACK
        self.gathered = torch.empty((world * B, 2 * self.K), device=dev) if world > 1 else None
        self.comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
        # priorities: the SM/latency-bound stages get SMs first, the HBM-bound shift-stack fills what is left
        three = args.streams == 3
        stack_pri = int(os.environ.get("ASW_BENCH_STACK_PRIORITY", "0"))
        score_pri = int(os.environ.get("ASW_BENCH_SCORE_PRIORITY", "-1"))
        self.stack_stream = torch.cuda.Stream(device=dev, priority=stack_pri) if three else None
        self.prune_stream = torch.cuda.Stream(device=dev, priority=-1) if three else None
        self.score_stream = torch.cuda.Stream(device=dev, priority=score_pri) if three else None
        self.serial = not three
        self.count_sync = bool(args.count_sync)
        # "plain" | "norm" (fused normalize_input from correlation tables) | "norm_exact" (per-patch pass) |
        # "norm_grouped" (one tiled pass per mixture serving all of its patches)
        self.mode = "plain"
        self.corr = None
        self.corr_tabs = None
        self.max_lag = MAX_LAG
        self.stage_events = []
        self.launched_patches = []   # patches covered by every timed shift-stack launch (count-sync mode)
        self.trace = None            # list of per-sub-batch event dicts while the pipelined schedule is being traced

    def enable_norm(self):
        from acousticswarms_speech_b200 import native
        if self.corr is None:
            self.max_lag = native.CorrTables.lag_for_geometry(self.node.mic_pos, self.node.FS)
            self.corr = native.CorrTables(self.M, self.dev, max_lag=self.max_lag)
            self.corr_tabs = [self.torch.empty((self.B, self.corr.table_len), device=self.dev, dtype=self.torch.float64)
                              for _ in range(2)]

    def compute(self, src, events=None, to_host=False):
        torch = self.torch
        caller = torch.cuda.current_stream(self.dev)
        main = caller if self.serial else self.score_stream
        if main is not caller:
            main.wait_stream(caller)                     # inputs produced on the caller's stream (e.g. uploads)
        with torch.cuda.stream(main):
            self._compute(src, events, to_host, main)

    def _compute(self, src, events, to_host, main):
        import torch.distributed as dist
        from acousticswarms_speech_b200 import native
        torch, fe, node = self.torch, self.fe, self.node
        B, M, K, MAXPATCH = self.B, self.M, self.K, self.MAXPATCH
        slot = self.step_no & 1
        self.step_no += 1
        if self.map_free[slot] is not None:
            main.wait_event(self.map_free[slot])         # prune of sub-batch i-2 is done with this map buffer
        timing_stages = events is not None and self.serial
        if timing_stages:                                # single-stream pass: stage boundaries for the `stages` object
            st_ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            st_ev[0].record(main)
            self.stage_events.append(st_ev)
        tr = None
        if self.trace is not None and not self.serial:
            tr = {k: torch.cuda.Event(enable_timing=True) for k in ("score0", "score1", "prune0", "prune1", "stack0", "stack1")}
            self.trace.append(tr)
            tr["score0"].record(main)
        m, val, idx = fe.score(src, out=self.maps[slot])
        if tr:
            tr["score1"].record(main)
        if self.collective:    # every rank learns every mixture's top-K
            ps = (self.step_no - 1) % self.NPACK
            pack = self.packs[ps]
            if self.pack_free[ps] is not None:
                main.wait_event(self.pack_free[ps])      # the gather that read this buffer NPACK sub-batches ago
            pack[:, :K] = val
            pack[:, K:] = idx.view(torch.float32)
            packed = torch.cuda.Event()
            packed.record()
            self.comm_stream.wait_event(packed)
            with torch.cuda.stream(self.comm_stream):
                dist.all_gather_into_tensor(self.gathered, pack)
                self.pack_free[ps] = torch.cuda.Event()
                self.pack_free[ps].record(self.comm_stream)
        pstream = main if self.serial else self.prune_stream
        sstream = main if self.serial else self.stack_stream
        if not self.serial:                              # val / idx are read on the prune stream (D2H below)
            val.record_stream(pstream)
            idx.record_stream(pstream)
        scored = torch.cuda.Event()
        scored.record(main)
        if timing_stages:
            self.stage_events[-1][1].record(main)
        pstream.wait_event(scored)
        with torch.cuda.stream(pstream):
            shifts_dev, mi_dev, ntot_dev = self.tables[slot]
            if self.table_free[slot] is not None:
                pstream.wait_event(self.table_free[slot])     # stack of sub-batch i-2 is done with this table
            if tr:
                tr["prune0"].record(pstream)
            peaks, count, _ = node.native_peaks.find(m)                            # fill_powermap + find_valid_peak_new
            n_p, off_p, wid_p, pk_p = node.native_select.select(m, peaks, count)   # local_source_adaptive
            native.build_shift_table(n_p, off_p, self.cap, shifts_dev, mi_dev, ntot_dev)
            if self.count_sync:
                self.ntot_pin[slot].copy_(ntot_dev, non_blocking=True)
            if self.mode == "norm":
                self.corr.compute(src, out=self.corr_tabs[slot])
            # the shift-stack (and the host's count read) wait for the table only, not for the result copies below
            pruned = torch.cuda.Event()
            pruned.record(pstream)
            if tr:
                tr["prune1"].record(pstream)
            if timing_stages:
                self.stage_events[-1][2].record(pstream)
            self.map_free[slot] = pruned
            if to_host:
                sp = self.sel_pin
                sp[:, 0].copy_(n_p, non_blocking=True)
                sp[:, 1:1 + MAXPATCH * (M - 1)].copy_(off_p.view(B, -1), non_blocking=True)
                sp[:, 1 + MAXPATCH * (M - 1):1 + MAXPATCH * M].copy_(wid_p, non_blocking=True)
                sp[:, 1 + MAXPATCH * M:].copy_(pk_p, non_blocking=True)
                self.peaks_pin.copy_(peaks, non_blocking=True)
                self.count_pin.copy_(count, non_blocking=True)
                self.map_pin.copy_(m, non_blocking=True)
                self.val_pin.copy_(val, non_blocking=True)
                self.idx_pin.copy_(idx, non_blocking=True)
                self.map_free[slot] = torch.cuda.Event()      # the map buffer is free once it has been copied out
                self.map_free[slot].record(pstream)
            for t in (peaks, count, n_p, off_p, wid_p, pk_p):
                t.record_stream(pstream)
        n_rows = self.cap
        if self.count_sync:
            # 4 bytes: how many patches this sub-batch selected.  The wait overlaps the shift-stack of the PREVIOUS
            # sub-batch, which is still running on its stream; no launch is issued for rows nobody selected.
            pruned.synchronize()
            n_rows = min(int(self.ntot_pin[slot][0]), self.cap)
        sstream.wait_event(pruned)
        with torch.cuda.stream(sstream):
            if tr:
                tr["stack0"].record(sstream)
            if self.mode == "plain":
                fe.stack_counted(src, shifts_dev, mi_dev, ntot_dev, n_rows, events=events)
            else:
                fe.stack_norm_counted(src, shifts_dev, mi_dev, ntot_dev, n_rows,
                                      tables=self.corr_tabs[slot] if self.mode == "norm" else None,
                                      max_lag=self.max_lag, events=events, grouped=self.mode == "norm_grouped")
            self.table_free[slot] = torch.cuda.Event()
            self.table_free[slot].record(sstream)
            if tr:
                tr["stack1"].record(sstream)
            if timing_stages:
                self.stage_events[-1][3].record(sstream)
        if events is not None:
            self.launched_patches.append(n_rows)

    def join(self):
        cur = self.torch.cuda.current_stream(self.dev)
        if not self.serial:
            cur.wait_stream(self.score_stream)
            cur.wait_stream(self.prune_stream)
            cur.wait_stream(self.stack_stream)
        if self.comm_stream is not None:
            cur.wait_stream(self.comm_stream)


def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    from acousticswarms_speech_b200 import _lib, native, synth
    from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
    from acousticswarms_speech_b200.pipeline import FrontEnd
    from acousticswarms_speech_b200.srp_phat import SRP_PHAT

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    skip = set(x for x in args.skip.split(",") if x)
    pk, pk_kind = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            x = float(t.item())
        return x

    B, NSUB, NU = args.batch, args.sub_batches, max(1, args.unique_batches)
    scene = synth.desk_array(N_MICS, np.random.default_rng(GEOM_SEED), FS)
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=FS, n_fft=n_fft, grid_size=0.05,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
    node.native.set_stft_path(args.stft_path)
    fe = FrontEnd(node, dev, launch_batches=args.launch_batches, ring=2 * args.launch_batches)
    G = node.grids.shape[0]
    M, T = N_MICS, T_SAMPLES

    # synthetic mixtures: every rank gets its own NU x B mixtures (weak scaling over mixtures), 16-bit PCM content
    pcm_pins, mix_pins, mix_devs = [], [], []
    for u in range(NU):
        raw = synth.mixtures(scene, N_SPK, T, seeds=[100_000 * rank + 1000 * u + 100 + b for b in range(B)])
        pcm = torch.from_numpy(np.clip(np.rint(raw * 32768.0), -32768, 32767).astype(np.int16))
        f32 = pcm.to(torch.float32) / 32768.0
        pcm_pins.append(pcm.pin_memory())
        mix_pins.append(f32.pin_memory())
        mix_devs.append(mix_pins[-1].to(dev))
    mix_host0 = mix_pins[0].numpy()

    # warm-up pass: how many coarse patches each distinct sub-batch selects (sizes the shift-table capacity; the timed
    # sub-batches select their own patches on the device and never read these lists)
    n_per = []
    for x in mix_devs:
        n_sel, _, _, _ = fe.select(fe.score(x)[0])
        n_per.append(int(n_sel.clamp(max=node.native_select.max_patches).sum()))
    torch.cuda.synchronize()
    MAXPATCH = node.native_select.max_patches
    cap = min(B * MAXPATCH, ((int(max(n_per) * 1.25) + fe.net_batch - 1) // fe.net_batch) * fe.net_batch)
    pipe = C2Pipeline(args, node, fe, dev, B, M, T, G, cap, world)
    patches_per_step = sum(n_per[j % NU] for j in range(NSUB))

    host_issue_ms = [0.0]

    def run_step(events=None):
        for j in range(NSUB):
            pipe.compute(mix_devs[j % NU], events=events)

    def timed(n_steps, events=None):
        """value: inputs already resident in HBM."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(n_steps):
            run_step(events)
        host_issue_ms[0] = 1e3 * (time.perf_counter() - t_host) / n_steps
        pipe.join()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    # e2e: every sub-batch copies its B mixtures from pinned host memory (ring of NBUF device buffers on a copy stream,
    # so the PCIe transfer of sub-batch i+1 overlaps the kernels of sub-batch i) and returns maps, top-K, peaks and
    # patch lists to the host.
    copy_stream = torch.cuda.Stream(device=dev, priority=-1)     # the PCM expansion kernel must not queue behind the stack
    NBUF = 3
    in_bufs = [torch.empty_like(mix_devs[0]) for _ in range(NBUF)]
    pcm_bufs = [torch.empty((B, M, T), device=dev, dtype=torch.int16) for _ in range(NBUF)]

    def timed_e2e(n_steps, pcm):
        """pcm=True: the host ships int16 PCM (what the data is) and asw_pcm16_to_f32 expands it on the device."""
        barrier()
        main = torch.cuda.current_stream(dev)
        copied = [torch.cuda.Event() for _ in range(NBUF)]
        consumed = [torch.cuda.Event() for _ in range(NBUF)]
        total = n_steps * NSUB
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_stream.wait_event(e0)

        def upload(i):
            s = i % NBUF
            if pcm:
                pcm_bufs[s].copy_(pcm_pins[i % NU], non_blocking=True)
                native.pcm16_to_f32(pcm_bufs[s], in_bufs[s])
            else:
                in_bufs[s].copy_(mix_pins[i % NU], non_blocking=True)

        with torch.cuda.stream(copy_stream):
            upload(0)
            copied[0].record()
        for i in range(total):
            cur, nxt = i % NBUF, (i + 1) % NBUF
            if i + 1 < total:
                with torch.cuda.stream(copy_stream):
                    if i + 1 >= NBUF:
                        copy_stream.wait_event(consumed[nxt])   # sub-batch i+1-NBUF was the last user of this buffer
                    upload(i + 1)
                    copied[nxt].record()
            main.wait_event(copied[cur])
            pipe.compute(in_bufs[cur], to_host=True)
            if pipe.stack_stream is not None:
                consumed[cur].record(pipe.stack_stream)     # the shift-stack is the last reader of the input buffer
            else:
                consumed[cur].record()
        pipe.join()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    def timed_h2d(n_steps, pcm):
        """The box's host -> device ceiling for the same bytes: plain pinned copies, nothing else running."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps * NSUB):
            if pcm:
                pcm_bufs[i % NBUF].copy_(pcm_pins[i % NU], non_blocking=True)
            else:
                in_bufs[i % NBUF].copy_(mix_pins[i % NU], non_blocking=True)
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    # ---- headline: value --------------------------------------------------------------------------
    W = max(args.warmup, 3)
    for _ in range(W):
        run_step()
    pipe.join()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = _lib.launch_count()
    ms = timed(args.steps)
    launches = _lib.launch_count() - l0
    host_ms = host_issue_ms[0]
    clocks = sampler.stop() if sampler else None

    ablation = None
    if world > 1:      # what the per-sub-batch top-K all-gather costs: same run without it
        pipe.collective = False
        timed(1)
        ms_nc = timed(args.steps)
        pipe.collective = True
        ablation = {"no_collective_ms_per_step": ms_nc / args.steps,
                    "note": "same timed loop with the per-sub-batch top-K all-gather (and its two pack kernels) switched off"}

    # ---- kernel-level timing of the dominant kernel with events on the launching stream; single-stream pass so the
    # kernel is timed alone, not while sharing SMs with the scoring kernels
    def kernel_pass(mode, n_steps=1):
        pipe.mode = mode
        if mode != "plain":
            pipe.enable_norm()
        saved = (pipe.serial, pipe.stack_stream)
        pipe.serial, pipe.stage_events, pipe.launched_patches = True, [], []
        events = []
        timed(1)
        timed(n_steps, events=events)
        torch.cuda.synchronize()
        pipe.serial = saved[0]
        k = [(a.elapsed_time(b), n) for a, b, n in events]
        rows_full = fe.net_batch * fe.launch_batches
        if pipe.count_sync:     # every launch covers exactly the rows it was given: use those of at least one network batch
            sel = [(t, n) for t, n in k if n >= fe.net_batch]
        else:       # capacity-sized launches: only those whose rows all lie below the sub-batch's patch count
            per = (cap + rows_full - 1) // rows_full
            sel = [(t, n) for i, (t, n) in enumerate(k) if ((i % per) + 1) * rows_full <= n_per[((i // per) % NSUB) % NU]]
        st = pipe.stage_events[1:] or pipe.stage_events
        rows = sum(n for _, n in sel)
        out = {"avg_launch_ms": sum(t for t, _ in sel) / max(1, len(sel)), "full_launches_timed": len(sel),
               "rows_per_launch": rows / max(1, len(sel)), "ms_per_128_rows": sum(t for t, _ in sel) / max(1, rows) * 128,
               "score_ms": sum(e[0].elapsed_time(e[1]) for e in st) / len(st),
               "prune_ms": sum(e[1].elapsed_time(e[2]) for e in st) / len(st),
               "stack_ms": sum(e[2].elapsed_time(e[3]) for e in st) / len(st)}
        pipe.mode = "plain"
        return out

    def variant(kp, kernel):
        byts = 4.0 * kp["rows_per_launch"] * M * T
        a = byts / (kp["avg_launch_ms"] / 1e3) / 1e9
        return {"kernel": kernel, "achieved": a, "frac": a / pk["hbm_gbs"], "avg_launch_ms": kp["avg_launch_ms"],
                "algorithmic_bytes_per_launch": byts, "patches_per_launch": kp["rows_per_launch"],
                "ms_per_128_patches": kp["ms_per_128_rows"], "launches_timed": kp["full_launches_timed"]}

    kp_plain = kernel_pass("plain")
    variants = {"plain": variant(kp_plain, "shift_stack_vec_kernel<false>")}
    fused_step = None
    if "variants" not in skip:
        kp_norm = kernel_pass("norm")
        kp_exact = kernel_pass("norm_exact")
        variants["fused_norm_tables"] = variant(kp_norm, "shift_ref_stats_kernel (table look-ups) + shift_stack_vec_kernel<true>")
        variants["fused_norm_exact"] = variant(kp_exact, "shift_ref_stats_kernel (exact pass) + shift_stack_vec_kernel<true>")
        kp_grp = kernel_pass("norm_grouped")
        variants["fused_norm_grouped"] = variant(kp_grp, "shift_stats_grouped_kernel (one tiled pass per mixture) + "
                                                         "shift_stack_vec_kernel<true>")
        # the whole pipelined step with the fused normalize_input: statistics from the grouped pass (a few dozen coarse
        # patches per mixture), and for comparison from correlation tables rebuilt per sub-batch on the pruning stream
        pipe.mode = "norm_grouped"
        timed(1)
        ms_grp = timed(args.steps)
        pipe.mode = "norm"
        timed(1)
        ms_norm = timed(args.steps)
        pipe.mode = "plain"
        fused_step = {"ms_per_step": ms_grp / args.steps,
                      "value": world * NSUB * B * G / (ms_grp / args.steps / 1e3), "unit": UNIT,
                      "statistics": "asw_shift_stack_norm_grouped",
                      "stack_stage_ms_per_sub_batch": kp_grp["stack_ms"], "prune_stage_ms_per_sub_batch": kp_grp["prune_ms"],
                      "with_correlation_tables": {"ms_per_step": ms_norm / args.steps,
                                                  "stack_stage_ms_per_sub_batch": kp_norm["stack_ms"],
                                                  "prune_stage_ms_per_sub_batch": kp_norm["prune_ms"]},
                      "note": "same step, shift-stack fused with normalize_input (what shift_and_sep feeds the network).  A "
                              "mixture's ~35 coarse patches take their statistics from one tiled pass over the mixture "
                              "(exact integer sums); per-mixture correlation tables (~15 us per mixture, "
                              "with_correlation_tables) pay off in the fine stage where they serve ~800 patches, see c3"}

    # ---- where the pipelined step's time goes: one traced step (events around every stage on its own stream), then
    # interval arithmetic on the host.  A stage's interval starts when its stream reaches it, so waits for SMs held by
    # another stage's kernels show up as stretched intervals, not as gaps.
    overlap = None
    if args.streams == 3:
        timed(1)
        pipe.trace = []
        base = torch.cuda.Event(enable_timing=True)
        barrier()
        base.record()
        run_step()
        pipe.join()
        torch.cuda.synchronize()
        tr, pipe.trace = pipe.trace, None

        def iv(a, b):
            return [(base.elapsed_time(t[a]), base.elapsed_time(t[b])) for t in tr]

        def union(ivs):
            tot, end = 0.0, -1.0
            for lo, hi in sorted(ivs):
                if hi > end:
                    tot += hi - max(lo, end)
                    end = hi
            return tot

        def inter(xs, ys):
            return union(xs) + union(ys) - union(xs + ys)
        sc, prn, stk = iv("score0", "score1"), iv("prune0", "prune1"), iv("stack0", "stack1")
        wall = max(hi for _, hi in stk) - min(lo for lo, _ in sc)
        n = len(tr)
        overlap = {"sub_batches_traced": n, "wall_ms_per_sub_batch": wall / n,
                   "score_busy_ms_per_sub_batch": union(sc) / n, "prune_busy_ms_per_sub_batch": union(prn) / n,
                   "stack_busy_ms_per_sub_batch": union(stk) / n,
                   "score_and_stack_concurrent_ms_per_sub_batch": inter(sc, stk) / n,
                   "prune_and_stack_concurrent_ms_per_sub_batch": inter(prn, stk) / n,
                   "nothing_running_ms_per_sub_batch": (wall - union(sc + prn + stk)) / n,
                   "serial_sum_ms_per_sub_batch": kp_plain["score_ms"] + kp_plain["prune_ms"] + kp_plain["stack_ms"],
                   "note": "intervals measured with CUDA events on the three streams during one pipelined step; a stage's "
                           "interval includes the time its kernels wait for SMs (the scoring kernels and the shift-stack do "
                           "not co-reside: each owns the register file while it runs), so `concurrent` is time the two "
                           "stages were both in flight, and wall - serial_sum is what the pipelining really saves"}

    # ---- self-check (untimed): the 3-stream pipeline must produce exactly what the serial schedule produces -- maps,
    # top-K, peak counts, patch counts, the device-built shift tables and the stacked ring buffers
    def snapshot(serial):
        saved = pipe.serial
        pipe.serial = serial or saved
        out = []
        for j in range(2):                                            # both buffer slots
            slot = pipe.step_no & 1
            for bf in fe._bufs:                                       # rows no launch of this sub-batch writes: zeros
                bf.zero_()
            pipe.compute(mix_devs[j % NU], to_host=True)
            pipe.join()
            torch.cuda.synchronize()
            sh, mi_t, nt = pipe.tables[slot]
            n = int(nt.item())
            out.append([pipe.map_pin.clone(), pipe.val_pin.clone(), pipe.idx_pin.clone(), pipe.count_pin.clone(),
                        pipe.sel_pin[:, 0].clone(), sh[:n].cpu(), mi_t[:n].cpu(), torch.tensor([n]),
                        torch.stack([bf.view(torch.int32).sum() for bf in fe._bufs]).cpu()])
        pipe.serial = saved
        return out
    ref_out, pipe_out = snapshot(True), snapshot(False)
    if not all(torch.equal(x, y) for a, b2 in zip(ref_out, pipe_out) for x, y in zip(a, b2)):
        raise RuntimeError("bench self-check failed: the pipelined schedule and the serial schedule disagree")

    # ---- end to end from pinned host memory -------------------------------------------------------
    e2e = e2e_f32 = None
    if "e2e" not in skip:
        es = max(2, min(args.steps, 10))
        timed_e2e(1, True)
        ms_e2e_pcm = timed_e2e(es, True) / es
        timed_e2e(1, False)
        ms_e2e_f32 = timed_e2e(es, False) / es
        timed_h2d(1, True)
        ms_h2d_pcm = timed_h2d(es, True) / es
        ms_h2d_f32 = timed_h2d(es, False) / es
        per_step = world * NSUB * B * G

        def e2e_obj(ms_step, ms_h2d, bytes_sub, what):
            gbs = NSUB * bytes_sub / (ms_step / 1e3) / 1e9
            ceil = NSUB * bytes_sub / (ms_h2d / 1e3) / 1e9
            return {"value": per_step / (ms_step / 1e3), "unit": UNIT, "ms_per_step": ms_step,
                    "h2d_bytes_per_step": int(NSUB * bytes_sub), "d2h_bytes_per_step": int(NSUB * pipe.d2h_bytes),
                    "ingest": what, "h2d_gbs_per_gpu": gbs, "h2d_ceiling_gbs_per_gpu": ceil,
                    "frac_of_h2d_ceiling": gbs / ceil,
                    "h2d_ceiling_note": f"plain pinned cudaMemcpyAsync of the same bytes on all {world} rank(s) at once, "
                                        "nothing else running, max over ranks",
                    "pipeline": "pinned host -> device copy of sub-batch i+1 overlaps the kernels of sub-batch i (3 input buffers)"}
        e2e = e2e_obj(ms_e2e_pcm, ms_h2d_pcm, B * M * T * 2,
                      "16-bit PCM shipped as int16, expanded on the device by asw_pcm16_to_f32 (exact)")
        e2e_f32 = e2e_obj(ms_e2e_f32, ms_h2d_f32, B * M * T * 4, "float32 samples")

    # ---- the other BASELINE configs ------------------------------------------------------------------
    extra = {}
    cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if "c3" not in skip:
        extra["c3"] = bench_c3(args, fe, node, mix_devs, dev, pk, reduce_max, barrier, world)
    if "c1" not in skip and rank == 0:
        extra["c1"] = bench_c1(dev, cpu)
    if world > 1:
        dist.barrier()
    if "c5" not in skip or ("sharded" not in skip and world > 1):
        c5, sharded = bench_c5(args, dev, rank, world, cpu, "c5" not in skip, "sharded" not in skip and world > 1,
                               reduce_max, barrier)
        if c5 is not None:
            extra["c5"] = c5
        if sharded is not None:
            extra["hypercube_sharded"] = sharded

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    per_step = world * NSUB * B * G
    ms_step = ms / args.steps
    value = per_step / (ms_step / 1e3)
    stack_bytes = 4.0 * patches_per_step * M * T
    score_bytes = NSUB * (4.0 * B * M * T + 4.0 * B * G) + 4.0 * G * node.native.P
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (16-bit PCM content carried as float32)",
        "config": {"workload": WORKLOAD, "mixtures_per_gpu_per_step": NSUB * B, "sub_batches_per_step": NSUB,
                   "mixtures_per_sub_batch": B, "distinct_resident_sub_batches": NU, "hypercubes": G, "mics": M,
                   "speakers": N_SPK, "samples": T, "fs": FS, "coarse_patches_per_step_per_gpu": patches_per_step,
                   "net_batch": fe.net_batch, "network_batches_per_stack_launch": fe.launch_batches, "streams": args.streams, "count_sync": int(pipe.count_sync),
                   "parallelism": f"mixtures sharded over {world} GPU(s)",
                   "l2": f"every sub-batch reads {B * M * T * 4 / 1e6:.0f} MB of inputs it last touched {NU} sub-batches "
                         f"({NU * B * M * T * 4 / 1e6:.0f} MB) ago and writes {n_per[0] * M * T * 4 / 1e9:.2f} GB of stacked "
                         "output: far beyond the 126 MB L2 (no explicit flush)",
                   "shift_table_capacity": cap,
                   "prune": "peak picking (fill_powermap + find_valid_peak_new) and greedy hypercube selection "
                            "(local_source_adaptive) run on the device inside every sub-batch; the shift-stack reads the "
                            "device-built patch table of its own sub-batch; with count_sync the host reads the 4-byte "
                            "patch count (behind the previous sub-batch's shift-stack) and launches exactly that many "
                            "patches, otherwise capacity-sized launches skip the rows beyond the device count"},
        "gpu_launches": int(launches), "host_issue_ms_per_step": host_ms,
        "roofline": {"kernel": "shift_stack_vec_kernel", "bound": "hbm", "achieved": variants["plain"]["achieved"],
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": variants["plain"]["frac"],
                     "peak_kind": pk_kind + " (copy, burst)",
                     "algorithmic_bytes_per_launch": variants["plain"]["algorithmic_bytes_per_launch"],
                     "avg_launch_ms": variants["plain"]["avg_launch_ms"], "traffic": None,
                     "variants": variants,
                     "step": {"algorithmic_bytes_per_step": stack_bytes + score_bytes,
                              "achieved": (stack_bytes + score_bytes) / (ms_step / 1e3) / 1e9,
                              "frac": (stack_bytes + score_bytes) / (ms_step / 1e3) / 1e9 / pk["hbm_gbs"],
                              "note": "whole pipelined step: 4 M T bytes per stacked patch + scoring's compulsory traffic "
                                      "(audio in, map out, lag table once) over the step time"}},
        "stages": {"note": "per GPU and per sub-batch, from a single-stream pass (stages not overlapped); SURVEY 8(d) stage metrics",
                   "srp_score_ms": kp_plain["score_ms"], "prune_ms": kp_plain["prune_ms"], "shift_stack_ms": kp_plain["stack_ms"],
                   "srp_hypercubes_per_s": B * G / (kp_plain["score_ms"] / 1e3),
                   "patches_stacked_per_s": (patches_per_step / NSUB) / (kp_plain["stack_ms"] / 1e3),
                   "shift_stack_frac_of_nominal_8tbs": variants["plain"]["achieved"] / 8000.0},
        "selfcheck": "pipelined sub-batch == serial sub-batch (maps, top-K, peak / patch counts, shift tables, stacked ring checksums)",
        "clocks": clocks,
    }
    if overlap:
        line["stages"]["pipelined"] = overlap
    if fused_step:
        line["fused_norm_step"] = fused_step
    if ablation:
        line["ablation"] = ablation
    if e2e:
        line["e2e"] = e2e
        line["e2e_f32"] = e2e_f32
    line.update(extra)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):      # ncu --set full capture, valid only for the kernel source it was taken from
        try:
            with open(tp) as fh:
                tr = json.load(fh)
            if tr.get("shift_stack_cu_sha16") == source_hash("shift_stack.cu"):
                line["roofline"]["traffic"] = tr.get("shift_stack_vec_kernel_bytes_per_launch")
                line["roofline"]["traffic_source"] = tr.get("source")
            else:
                line["roofline"]["traffic_source"] = "profiles/traffic.json is from another version of shift_stack.cu: not quoted"
        except Exception:
            pass
    if cpu:
        try:
            t0 = time.perf_counter()
            _, geo, ref = cpu_reference_setup()
            cpu_reference_step(ref, mix_host0[0])                    # warm-up
            best = None
            for i in range(3):
                t1 = time.perf_counter()
                cpu_reference_step(ref, mix_host0[(i + 1) % B])
                dt = time.perf_counter() - t1
                best = dt if best is None else min(best, dt)
            line["cpu_baseline"] = {"value": G / best, "unit": UNIT, "cores": ref.threads, "kind": "port",
                                    "sample": "1 mixture of the workload per run (Apply_SRP_PHAT + coarse shift loop), 1 warm-up + best of 3",
                                    "seconds_per_mixture": best, "setup_seconds_excluded": time.perf_counter() - t0 - 4 * best}
            if "c3" in line and line["c3"] is not None:
                line["c3"]["cpu_baseline"] = cpu_c3(ref, geo, mix_host0[1])
        except Exception as e:  # the CPU leg must never lose the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
def bench_c3(args, fe, node, mix_devs, dev, pk, reduce_max, barrier, world):
    """BASELINE configs[2]: fine refinement over the surviving hypercubes of 64 mixtures per GPU.  Per group of 16
    mixtures: score -> peaks -> greedy selection (every selected coarse patch is a candidate: no separator weights
    exist to thin them, sep/helpers/local_utils_3d.py:339-388) -> ONE asw_subdivide launch (search_area) -> fine shift
    table on the device -> per-mixture correlation tables -> fused shift-stack + normalize_input of every fine and
    centre patch (what shift_and_sep(Strict=1) feeds the network, sep/Mic_Array.py:263).  The host reads one 4-byte
    patch count per group, behind the previous group's stacking."""
    import torch
    from acousticswarms_speech_b200 import native
    GB, NMIX = args.c3_group, 64
    B, M, T = mix_devs[0].shape
    groups = []
    for x in mix_devs:
        for g0 in range(0, B, GB):
            groups.append(x[g0:g0 + GB])
    while len(groups) * GB < NMIX:
        groups = groups + groups
    groups = groups[:NMIX // GB]
    P = node.native_select.max_patches
    cap = GB * 1280
    max_lag = native.CorrTables.lag_for_geometry(node.mic_pos, node.FS)
    corr = native.CorrTables(M, dev, max_lag=max_lag)
    front_stream, stack_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    fe_fine = type(fe)(node, dev, net_batch=fe.net_batch, launch_batches=8, ring=16)    # its own ring and map buffers

    def front(x):
        """Everything before the stack, on the front stream; returns the device tables + an event."""
        with torch.cuda.stream(front_stream):
            smap, _, _ = fe_fine.score(x, out=torch.empty((x.shape[0], node.grids.shape[0]), device=dev))
            n_sel, off, wid, pk_ = fe_fine.select(smap)
            shifts, mi, ci, cstart, ntot, status, cnt = fe_fine.fine_table(n_sel, off, wid, cap)
            tabs = corr.compute(x)
            pin = torch.zeros((3,), dtype=torch.int32).pin_memory()
            pin[0:1].copy_(ntot, non_blocking=True)
            pin[1:2].copy_(status.max().reshape(1).to(torch.int32), non_blocking=True)
            pin[2:3].copy_(cnt.max().reshape(1).to(torch.int32), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(front_stream)
        return {"x": x, "shifts": shifts, "mi": mi, "ntot": ntot, "tabs": tabs, "pin": pin, "ev": ev,
                "keep": (smap, n_sel, off, wid, pk_, ci, cstart, status, cnt)}

    def run(events=None, serial=False):
        front_stream.wait_stream(torch.cuda.current_stream(dev))
        stack_stream.wait_stream(torch.cuda.current_stream(dev))
        total, held = 0, []
        nxt = front(groups[0])
        for g in range(len(groups)):
            cur = nxt
            held.append(cur)
            if g + 1 < len(groups) and not serial:
                nxt = front(groups[g + 1])              # runs while group g is being stacked
            cur["ev"].synchronize()                      # 4-byte count (+ capacity flags)
            n, st, mx = (int(v) for v in cur["pin"])
            if st != 0 or mx > 128 or n >= cap:
                raise RuntimeError(f"c3: fine-table capacity exceeded (status {st}, leaves {mx}, rows {n} of {cap})")
            total += n
            stack_stream.wait_event(cur["ev"])
            with torch.cuda.stream(stack_stream):
                fe_fine.stack_norm_counted(cur["x"], cur["shifts"], cur["mi"], cur["ntot"], n, tables=cur["tabs"],
                                           max_lag=max_lag, events=events)
                done = torch.cuda.Event()
                done.record(stack_stream)
            if serial:
                done.synchronize()
                if g + 1 < len(groups):
                    nxt = front(groups[g + 1])
        torch.cuda.current_stream(dev).wait_stream(stack_stream)
        torch.cuda.current_stream(dev).wait_stream(front_stream)
        return total, held

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    reps = 3
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        total, held = run()
        del held
    e1.record()
    barrier()
    ms = reduce_max(e0.elapsed_time(e1)) / reps
    # kernel-level: fused launches alone (serial pass), and the front stage alone
    events = []
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    total, held = run(events=events, serial=True)
    t1.record()
    torch.cuda.synchronize()
    ms_serial = t0.elapsed_time(t1)
    k = [(a.elapsed_time(b), n) for a, b, n in events if n >= fe.net_batch]
    k_rows = sum(n for _, n in k)
    k_ms = sum(t for t, _ in k) / max(1, k_rows) * fe.net_batch          # per 128 patches
    stack_ms = sum(a.elapsed_time(b) for a, b, n in events)
    byts = 4.0 * total * M * T
    # spot check of the device-built fine table against the reference's host loop on one candidate-rich mixture is
    # done by tests/test_gpu_integration.py::test_c3_shape_fine_refinement_over_a_batch; here: table invariants
    h = held[0]
    n0 = int(h["pin"][0])
    mi0 = h["mi"][:n0].cpu().numpy()
    sh0 = h["shifts"][:n0].cpu().numpy()
    assert (sh0[:, 0] == 0).all() and mi0.min() >= 0 and mi0.max() < GB and (np.diff(mi0) >= 0).all()
    return {"workload": "C3: fine width-2 Spotform_Small_Patch_Parallel refinement over surviving hypercubes, 64 mixtures per GPU, "
                        "7 mics, 3 s @ 48 kHz", "mixtures_per_gpu": NMIX, "fine_patches_per_gpu": total,
            "fine_patches_per_mixture": total / NMIX, "ms_per_pass": ms, "corr_table_max_lag": max_lag,
            "mixtures_per_group": GB, "patches_per_s": world * total / (ms / 1e3),
            "hbm": {"algorithmic_bytes": byts, "achieved_gbs": byts / (ms / 1e3) / 1e9,
                    "frac": byts / (ms / 1e3) / 1e9 / pk["hbm_gbs"],
                    "kernel_ms_per_128_patches": k_ms, "patches_per_launch": fe_fine.net_batch * fe_fine.launch_batches,
                    "kernel_achieved_gbs": 4.0 * fe.net_batch * M * T / (k_ms / 1e3) / 1e9,
                    "kernel_frac": 4.0 * fe.net_batch * M * T / (k_ms / 1e3) / 1e9 / pk["hbm_gbs"],
                    "kernel": "shift_ref_stats_kernel (table look-ups) + shift_stack_vec_kernel<true>"},
            "serial_pass_ms": ms_serial, "serial_stack_ms": stack_ms, "serial_front_ms": ms_serial - stack_ms,
            "note": "timed pass = score + peaks + selection + asw_subdivide + fine table + correlation tables + fused "
                    "shift-stack/normalize of all fine patches, two groups in flight (front of group g+1 overlaps the "
                    "stacking of group g); serial_* = the same with the stages back to back"}


def cpu_c3(ref, geo, mix):
    """CPU leg of c3 on a bounded sample: the reference's fine stage for ONE mixture -- search_area for every coarse
    patch (sep/helpers/local_utils_3d.py:212-335) timed in full, the shift loop + normalize_input timed on the first
    128 fine patches and scaled to the mixture's fine-patch count."""
    import torch
    from oracle import shift_oracle, subdivide_oracle
    t0 = time.perf_counter()
    patches = cpu_reference_step(ref, mix)
    t1 = time.perf_counter()
    fine, _ = subdivide_oracle.small_patch_list(patches, geo.mic_pos)
    t2 = time.perf_counter()
    sample = fine[:128]
    data = ref.shift_stack(mix, sample)
    x = data[:len(sample)]
    x = torch.round(x * 2 ** 15) / 2 ** 15                     # normalize_input (SpeakerLocalization/network.py:28-40)
    r = x.mean(1, keepdim=True)
    x = (x - r.mean(2, keepdim=True)) / r.std(2, keepdim=True)
    t3 = time.perf_counter()
    per_patch = (t3 - t2) / len(sample)
    total = (t2 - t1) + per_patch * len(fine)
    return {"value": len(fine) / total, "unit": "fine patches/s", "cores": ref.threads, "kind": "port",
            "sample": f"one mixture: search_area of its {len(patches)} coarse patches in full ({t2 - t1:.1f} s, {len(fine)} fine "
                      f"patches), shift loop + normalize_input on {len(sample)} of them ({per_patch * 1e3:.1f} ms each) scaled",
            "seconds_per_mixture": total}


def bench_c1(dev, cpu):
    """BASELINE configs[0]: one 7-mic, 3 s, 44.1 kHz mixture with 3 speakers through the drop-in
    Mic_Array.Apply_SRP_PHAT (sep/Mic_Array.py:152-194): host tensor in, Patch list out, one call per mixture."""
    import torch
    from acousticswarms_speech_b200 import synth
    from acousticswarms_speech_b200.mic_array import Mic_Array
    scene = synth.desk_array(N_MICS, np.random.default_rng(0), C1_FS)
    t0 = time.perf_counter()
    ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi, fs=C1_FS, device=dev)
    t_setup = time.perf_counter() - t0
    mixes = [torch.from_numpy(pcm_content(synth.mixture(scene, C1_SPK, C1_T, seed=s))) for s in range(4)]
    for m in mixes:
        ma.Apply_SRP_PHAT(m)
    torch.cuda.synchronize()
    ts, npatch = [], 0
    for i in range(40):
        t = time.perf_counter()
        patches, _ = ma.Apply_SRP_PHAT(mixes[i % 4])
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
        npatch = len(patches)
    G = ma.SRP_node.grids.shape[0]
    med = float(np.median(ts))
    out = {"workload": "C1: Mic_Array.Apply_SRP_PHAT on one synthetic 7-mic, 3 s, 44.1 kHz mixture with 3 speakers",
           "hypercubes": G, "latency_us_median": med * 1e6, "latency_us_min": min(ts) * 1e6, "calls": len(ts),
           "hypercubes_per_s": G / med, "patches_returned": npatch, "geometry_setup_s": t_setup,
           "h2d_bytes_per_call": int(N_MICS * C1_T * 4),
           "note": "wall clock around the reference-facing call, host float32 tensor in, list[Patch] out, device synchronised"}
    if cpu:
        try:
            from acousticswarms_speech_b200.constants import freq_bins, n_fft
            from oracle import cpu_reference, geometry_oracle
            t0 = time.perf_counter()
            geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, FS=C1_FS)
            ref = cpu_reference.CpuReferencePath(geo, freq_bins, C1_FS, n_fft)
            t_set = time.perf_counter() - t0
            ref.apply_srp_phat(mixes[0].numpy())
            t1 = time.perf_counter()
            pr, _ = ref.apply_srp_phat(mixes[1].numpy())
            dt = time.perf_counter() - t1
            out["cpu_baseline"] = {"value": G / dt, "unit": UNIT, "cores": ref.threads, "kind": "port",
                                   "sample": "one C1 mixture, Apply_SRP_PHAT only, 1 warm-up + 1 run",
                                   "seconds_per_call": dt, "setup_seconds_excluded": t_set, "patches": len(pr)}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": f"failed: {e!r}"}
    return out


def sharded_entry(node, mic_positions, mix, K, full_map, fval, fidx, dev, rank, world, time_it, label):
    """One hypercube-sharded pass (TableExchangeSRP: transform stage sharded over mixtures, NCCL all-gather of the GCC
    lag tables, per-rank gather + top-K, NCCL all-gather + merge) timed against the unsharded pass on one GPU, and
    compared with it bit for bit (merged top-K on every rank, every rank's map slice)."""
    import torch
    import torch.distributed as dist
    from acousticswarms_speech_b200 import constants, dist as adist, native
    M, T = mix.shape[1], mix.shape[2]
    win = constants.window_length(T)
    lag = native.pair_lags(node.grids, mic_positions, 48000, 343.0)
    sh, handle = adist.native_table_exchange_srp(lag, M, dev)
    ms_sh = time_it(lambda: sh.topk(mix, K), 5)
    ms_single = time_it(lambda: native.map_topk(node.native.score(mix, win), K), 5)
    val, idx = sh.topk(mix, K)
    slice_map = sh.score_slice(mix)
    torch.cuda.synchronize()
    ok_topk = bool(torch.equal(val, fval) and torch.equal(idx, fidx))
    ok_map = bool(torch.equal(slice_map, full_map[:, sh.g0:sh.g1]))
    flags = torch.tensor([int(ok_topk), int(ok_map)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if not bool(flags.min().item()):
        raise RuntimeError(f"hypercube-sharded pass differs from the unsharded one ({label}; rank {rank}: top-K {ok_topk}, map {ok_map})")
    G, B = full_map.shape[1], mix.shape[0]
    out = {"workload": label + ": hypercubes sharded over the ranks, transform stage sharded over mixtures, NCCL all-gather of "
                       "the GCC lag tables, per-rank gather + top-K, NCCL all-gather + merge of the top-K lists (all inside "
                       "the timed region)",
           "ranks": world, "hypercubes": G, "mixtures": B, "ms_sharded": ms_sh, "ms_one_gpu_unsharded": ms_single,
           "speedup_vs_one_gpu": ms_single / ms_sh, "hypercubes_per_s": B * G / (ms_sh / 1e3),
           "selfcheck": "merged top-K (values and indices) and every rank's map slice are bit-identical to the unsharded "
                        "result on the same rank (all ranks agreed)"}
    del sh, handle
    return out


def bench_c5(args, dev, rank, world, cpu, do_c5, do_sharded, reduce_max, barrier):
    """BASELINE configs[4]: 16-mic distributed array, 10 s clips, 2.5 cm grid (G ~ 1e5, P = 120): scoring stages.
    With N > 1 the same geometry is used for the hypercube-sharded pass of configs[3]'s exchange scheme: the transform
    stage sharded over mixtures, an NCCL all-gather of the GCC lag tables, every rank gathers ITS hypercubes, then an
    all-gather + merge of the per-rank top-K -- all inside the timed region -- and, untimed, the merged top-K and the
    map slices are compared bit for bit with the unsharded result computed on the same rank."""
    import torch
    import torch.distributed as dist
    from acousticswarms_speech_b200 import constants, dist as adist, native, synth
    from acousticswarms_speech_b200.srp_phat import SRP_PHAT
    scene = synth.table_array(C5_MICS, np.random.default_rng(16))
    t0 = time.perf_counter()
    node = SRP_PHAT(scene.mic_positions, constants.freq_bins, scene.roi, FS=48000, n_fft=constants.n_fft,
                    grid_size=0.025, grid_size_z=0.05, threshold=list(constants.SRP_THRESHOLDS), WIDTH=8, device=dev)
    t_setup = time.perf_counter() - t0
    G, P = node.grids.shape[0], node.native.P
    Bc = 4
    mix_host = np.stack([pcm_content(synth.mixture(scene, C5_SPK, C5_T, seed=8 + b)) for b in range(Bc)])
    mix = torch.from_numpy(mix_host).to(dev)                    # the same mixtures on every rank
    win = constants.window_length(C5_T)
    K = 128
    c5 = sharded = None

    def time_it(fn, reps):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1)) / reps

    full_map = node.native.score(mix, win)
    fval, fidx = native.map_topk(full_map, K)
    if do_c5:
        Nw = node.native.num_windows(C5_T, win)
        gcc = node.native.gcc(mix, win)
        ms_score = time_it(lambda: node.native.score(mix, win), 5)
        ms_gcc = time_it(lambda: node.native.gcc(mix, win, out=gcc), 5)
        ms_gather = time_it(lambda: node.native.gather(gcc, Nw), 5)
        ms_topk = time_it(lambda: native.map_topk(full_map, K), 5)
        peaks_, count_, _ = node.native_peaks.find(full_map)
        ms_peaks = time_it(lambda: node.native_peaks.find(full_map), 5)
        c5 = {"workload": "C5: 16-mic distributed array, 10 s clips, 2.5 cm x 5 cm grid (stress pair count and hypercube volume)",
              "mics": C5_MICS, "pairs": P, "hypercubes": G, "samples": C5_T, "windows": Nw, "mixtures_per_call": Bc,
              "score_ms": ms_score, "hypercubes_per_s": Bc * G / (ms_score / 1e3),
              "stages_ms": {"stft_phat_cc_gcc": ms_gcc, "srp_gather": ms_gather, "topk": ms_topk, "peaks": ms_peaks},
              "geometry_setup_s": t_setup, "peak_clusters_first_mixture": int(count_[0]),
              "note": "every rank times the same 4 mixtures on the full grid (replicas); max over ranks"}
        if cpu:
            try:
                from oracle import srp_oracle
                sub = np.linspace(0, G - 1, 4096).astype(np.int64)
                t1 = time.perf_counter()
                srp_oracle.score(mix_host[0][:, :2 * win], node.grids[sub], scene.mic_positions, constants.freq_bins,
                                 48000, constants.n_fft)
                dt = time.perf_counter() - t1
                nwin_s = 3                                        # windows in 2 * win samples
                est = dt * (G / len(sub)) * (Nw / nwin_s)
                c5["cpu_baseline"] = {"value": G / est, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                      "sample": f"numpy restatement (the reference's own table would be {16 * G * 198 * P / 1e9:.0f} GB: "
                                                f"it cannot run this config), {len(sub)} of {G} hypercubes x {nwin_s} of {Nw} windows of one "
                                                f"mixture in {dt:.1f} s, scaled linearly", "seconds_per_mixture_estimated": est}
            except Exception as e:
                c5["cpu_baseline"] = {"value": None, "sample": f"failed: {e!r}"}
    if do_sharded:
        sharded = sharded_entry(node, scene.mic_positions, mix, K, full_map, fval, fidx, dev, rank, world, time_it,
                                "C4 exchange scheme on the C5 geometry (16 mics, 10 s, 4 mixtures)")
        # BASELINE configs[3] as named: 256 mixtures of the 7-mic C2 geometry, hypercubes sharded over the ranks
        del full_map, mix
        scene7 = synth.desk_array(N_MICS, np.random.default_rng(GEOM_SEED), FS)
        node7 = SRP_PHAT(scene7.mic_positions, constants.freq_bins, scene7.roi, FS=FS, n_fft=constants.n_fft, grid_size=0.05,
                         threshold=list(constants.SRP_THRESHOLDS), WIDTH=8, device=dev)
        base = torch.from_numpy(np.stack([pcm_content(synth.mixture(scene7, N_SPK, T_SAMPLES, seed=40 + b))
                                          for b in range(8)])).to(dev)
        mix7 = torch.cat([torch.roll(base, shifts=17 * i, dims=2) for i in range(32)], 0).contiguous()   # 256 mixtures
        map7 = node7.native.score(mix7, constants.window_length(T_SAMPLES))
        v7, i7 = native.map_topk(map7, K)
        sharded_c4 = sharded_entry(node7, scene7.mic_positions, mix7, K, map7, v7, i7, dev, rank, world, time_it,
                                   "C4: 256 mixtures of the 7-mic C2 geometry (3 s @ 48 kHz)")
        sharded_c4["note"] = ("for this many mixtures of a small grid, sharding the MIXTURES (the bench's headline value: no "
                              "table exchange, one top-K all-gather) is the faster split; the hypercube split pays the "
                              "all-gather of 176 MB of lag tables and wins when the grid is large and the batch small "
                              "(hypercube_sharded, C5 geometry)")
        sharded["c4_256_mixtures"] = sharded_c4
    return c5, sharded


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world)


if __name__ == "__main__":
    main()
