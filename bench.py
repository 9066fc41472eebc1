#!/usr/bin/env python
"""bench.py — beamformed audio-seconds per second on B200 (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

A "step" is `passes_per_step` passes of the fused hot path over one batch of synthetic multichannel audio
(`--workload`): n_streams independent streams x hops_per_pass hops, resident in HBM, every pass continuing the
streams' state (the batch is larger than L2; passes are repeated so that the timed region exceeds one second).
  value        whole-job audio-seconds / second over all ranks (device-resident inputs, CUDA events, max over ranks)
  e2e          same metric through the host-buffer C-ABI call (pinned host -> H2D -> kernels -> D2H), beside the
               measured pinned-copy ceiling of the box
  roofline     algorithmic bytes ((M+1)*4 B per sample) / measured kernel time, vs MEASURED_PEAKS.json
  cpu_baseline the reference's CPU path (oracle/_ref: the reference's unmodified node sources; FFTW/Eigen are plain
               stand-ins) on the host cores, bounded sample
  workloads    the other BASELINE.json configs (C1, C3 lcmv/gss, C4, C5), each timed briefly the same way
With --impl reference only the CPU path runs (no product library is imported or loaded).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

SR = 48000
CPU_NOTE = ("reference node sources compiled unmodified against oracle/shim; FFTW and Eigen are plain stand-ins there (radix-2 Stockham FFT, "
            "eager heap-allocating matrices, -O2): real FFTW/Eigen would be faster, so GPU/CPU ratios are upper bounds")

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on.  1184 = 8 x 148 streams ("batched 1k streams").
    "c2": dict(name="C2: MVDR 8-mic 1024-pt, energy-thresholded bins, batched synthetic streams", algo="mvdr", mics="circ8",
               n_streams=1184, hops=188, passes=16, kernel="sel_stream_kernel<mvdr>"),
    # same workload with the gate opened 5x (freq_mag_threshold 0.0002): the selection density SURVEY.md section 8d expected
    "c2hi": dict(name="C2 variant: MVDR 8-mic 1024-pt, freq_mag_threshold 0.0002 (dense selection)", algo="mvdr", mics="circ8", n_streams=1184,
                 hops=188, passes=8, kernel="sel_stream_kernel<mvdr>", params=dict(freq_mag_threshold=0.0002)),
    "c1": dict(name="C1: DAS 3-mic (aira3) 1024-pt, batched synthetic streams", algo="das", mics="aira3", n_streams=2048,
               hops=188, passes=48, kernel="das_pairs_kernel"),
    "c3l": dict(name="C3: LCMV 8-mic, 3 interferers", algo="lcmv", mics="circ8", n_streams=1184, hops=188, passes=8,
                interferers=(80.0, -60.0, 150.0), kernel="sel_pairs_kernel<lcmv>"),
    "c3g": dict(name="C3: GSS 8-mic, 3 interferers", algo="gss", mics="circ8", n_streams=1184, hops=188, passes=12,
                interferers=(80.0, -60.0, 150.0), kernel="sel_pairs_kernel<gss>"),
    "c4": dict(name="C4: PhaseMPF 2-mic (binaural) 4096-pt, phase mask + MCRA bi-channel post-filter", algo="phasempf", mics="binaural",
               n_streams=1184, hops=47, hop=2048, passes=12, kernel="frames_kernel_n<phasempf,4096>"),
    # BASELINE.json configs[4]: steered-response sweep; streams are sharded across ranks and the maps are gathered (NCCL)
    "c5": dict(name="C5: 64-mic (8x8 grid, 4 cm) steered-response DAS sweep over 360 directions, 1024-pt", algo="das", mics="grid64",
               n_streams=100, hops=188, passes=1, kernel="srp_power_tc_kernel", srp_dirs=360),
    # SURVEY.md section 8f rank 2 (single-channel nodes: 2 x 4 algorithmic bytes per sample)
    "mcra": dict(name="MCRA node (mcra.launch), first microphone of aira3, 1024-pt", algo="mcra", mics="aira3", n_streams=2368, hops=188, passes=8,
                 kernel="mcra_pairs_kernel", alg_channels=1),
    "ref": dict(name="rosjack_ref passthrough (window^2 overlap-add), first microphone of aira3", algo="ref", mics="aira3", n_streams=2368,
                hops=188, passes=64, kernel="ref_kernel", alg_channels=1),
    # SURVEY.md section 8f rank 1
    "gsc": dict(name="GSC 3-mic (aira3) 1024-pt: per-microphone alignment + 128-tap NLMS (gsc.launch)", algo="gsc", mics="aira3", n_streams=4736,
                hops=94, passes=1, kernel="gsc_nlms_kernel (+ gsc_align_kernel<1024>)"),
    "ph": dict(name="Phase 3-mic (aira3) 1024-pt phase mask", algo="phase", mics="aira3", n_streams=1184, hops=188, passes=8,
               kernel="phase_n_kernel<phase,1024,3>"),
}
EXTRA = ("c1", "c3l", "c3g", "c4", "c5")   # the other BASELINE.json configs, reported under "workloads" by the default run


def load_by_path(name, rel):
    """Import a pure-Python file of the package WITHOUT importing the package (the reference arm must not load the product)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_config(wl, tables):
    """The `config` object of the JSON line: a pure function of the workload table, identical in both arms."""
    M = len(tables.GEOMETRIES[wl["mics"]])
    Hh = wl.get("hop", 512)
    B, T, P = wl["n_streams"], wl["hops"], wl.get("passes", 1)
    cfg = {"workload": wl["name"], "algo": "srp" if "srp_dirs" in wl else wl["algo"], "n_mics": M, "fft_win": 2 * Hh, "hop": Hh, "sample_rate": SR,
           "streams_per_gpu": B, "hops_per_pass": T, "passes_per_step": P, "audio_s_per_step_per_gpu": B * T * Hh * P / SR,
           "input_bytes_per_pass_per_gpu": B * M * T * Hh * 4, "l2_policy": "inputs of a pass larger than L2 (126 MB), no flush"}
    if "srp_dirs" in wl:
        cfg["directions"] = wl["srp_dirs"]
    if wl.get("params"):
        cfg["params"] = dict(wl["params"])
    return cfg


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU path (reference arm and cpu_baseline): no product code
# ----------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU path on a bounded sample of a workload: `streams` synthetic streams of `hops` hops, generated
    ONCE, then processed by oracle/_ref/<node>_ref (the reference's own node sources) on all host threads, one stream per
    thread at a time.  Falls back to the oracle port when the binaries did not travel with the tree."""

    def __init__(self, wl, tables, synth, cores):
        import oracle_lib
        import ref_lib
        self.wl, self.cores = wl, cores
        self.algo = wl["algo"]
        self.H = wl.get("hop", 512)
        self.srp = "srp_dirs" in wl
        fields = tables.plain_config_fields(wl["algo"], mics=wl["mics"], hop=self.H, interferers=wl.get("interferers", ()), **wl.get("params", {}))
        self.cfg = oracle_lib.config_from_fields(fields)
        self.kind = "reference" if (ref_lib.available(self.algo) and not self.srp) else "port"
        self.ref_lib, self.oracle_lib = ref_lib, oracle_lib
        xy = tables.GEOMETRIES[wl["mics"]]
        if self.srp:   # the sweep has no reference node: the oracle evaluates the reference DAS formula per direction (very slow)
            self.n, self.hops = cores, 2
            self.thetas = (-180.0 + 360.0 * np.arange(wl["srp_dirs"]) / wl["srp_dirs"]).astype(np.float32).astype(np.float64)
        else:
            self.n = 2 * cores
            self.hops = min(wl["hops"], 188) if self.H <= 512 else wl["hops"]
            if self.algo in ("das", "ref", "mcra", "phase"):
                self.hops *= 4    # cheap nodes: longer streams so that process start-up and file I/O stay negligible
        L = self.hops * self.H
        if self.kind == "port":
            oracle_lib.lib()
        with ThreadPoolExecutor(max_workers=cores) as ex:   # numpy's sin releases the GIL: the sample is generated in parallel, once
            parts = list(ex.map(lambda b: synth.synth_batch(xy, 1, L, seed=0xC0FFEE + 31 * b)[0], range(self.n)))
        self.x = np.stack(parts)
        self.audio_s = self.n * L / SR
        self.sample = "%d streams x %d hops (%.1f s of audio) of the same workload, generated once, %d threads" % (self.n, self.hops, self.audio_s, cores)

    def _one(self, b):
        if self.srp:
            self.oracle_lib.Oracle(self.cfg).srp(self.x[b], self.thetas)
        elif self.kind == "reference":
            self.ref_lib.run_ref(self.algo, self.cfg, self.x[b])
        else:
            self.oracle_lib.Oracle(self.cfg).process(self.x[b])
        return 0

    def step(self):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.cores) as ex:
            list(ex.map(self._one, range(self.n)))
        return time.perf_counter() - t0

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        dts = [self.step() for _ in range(steps)]
        tot = sum(dts)
        return self.audio_s * steps / tot, tot


def reference_arm(args, wl, tables, synth, cores):
    config = workload_config(wl, tables)
    ref = CpuReference(wl, tables, synth, cores)
    value, tot = ref.run(args.steps, min(args.warmup, 2))
    print(json.dumps({
        "impl": "reference", "metric": "beamformed audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": ref.kind, "sample": ref.sample + "; each step processes the sample once"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": CPU_NOTE if ref.kind == "reference" else "CPU restatement of the reference (oracle port); oracle/_ref was not shipped"}))
    return 0


def cpu_baseline(wl, tables, synth, cores, steps=3):
    ref = CpuReference(wl, tables, synth, cores)
    v, tot = ref.run(steps, 1 if not ref.srp else 0)
    return {"value": v, "unit": "audio-s/s", "cores": cores, "kind": ref.kind, "sample": ref.sample + ", %d passes over it, %.1f s wall" % (steps, tot),
            "note": CPU_NOTE if ref.kind == "reference" else "oracle port (the sweep is an extension: the reference DAS formula evaluated per direction)"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def device_synth(torch, mic_xy, n_streams, n_samples, seed, device):
    """Synthetic plane-wave batch generated on the device (same signal model as beamform_b200.synth)."""
    from beamform_b200.synth import mic_delays
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.Generator(np.random.PCG64(seed))
    M = len(mic_xy)
    x = torch.empty((n_streams, M, n_samples), dtype=torch.float32, device=device)
    n = torch.arange(n_samples, dtype=torch.float64, device=device) / SR
    env = torch.ones(n_samples, dtype=torch.float64, device=device)
    env[:2048] = 0.0
    chunk = 32
    for b0 in range(0, n_streams, chunk):
        nb = min(chunk, n_streams - b0)
        acc = 1e-3 * torch.randn((nb, M, n_samples), dtype=torch.float32, device=device, generator=g)
        for (lo, hi, amp) in ((-40.0, 40.0, 0.1), (60.0, 300.0, 0.05)):
            theta = rng.uniform(lo, hi, size=nb)
            f0 = rng.uniform(120.0, 260.0, size=nb)
            tau = np.stack([mic_delays(mic_xy, th) for th in theta])          # [nb][M]
            tau_t = torch.tensor(tau, dtype=torch.float64, device=device)[:, :, None]
            f0_t = torch.tensor(f0, dtype=torch.float64, device=device)[:, None, None]
            t = n[None, None, :] - tau_t
            s = torch.zeros((nb, M, n_samples), dtype=torch.float64, device=device)
            for h in range(1, 13):
                s += (1.0 / h) * torch.sin(2 * np.pi * h * f0_t * t + float(rng.uniform(0, 2 * np.pi)))
            acc += (amp * env[None, None, :] * s).to(torch.float32)
        x[b0:b0 + nb] = acc
    return x


class Ctx:
    pass


def copy_ceiling(torch, xh, xd, yh, yd, reps=2):
    """Pinned-memory copy ceiling of this rank: the step's H2D and D2H transfers alone, concurrently on two streams (what the
    host-buffer entry overlaps them with), no kernels.  Returns (seconds per step, h2d GB/s, d2h GB/s)."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            e0.record()
            xd.copy_(xh, non_blocking=True)
            e1.record()
        with torch.cuda.stream(s2):
            e2.record()
            yh.copy_(yd, non_blocking=True)
            e3.record()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r = (dt, xh.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9, yh.numel() * 4 / (e2.elapsed_time(e3) * 1e-3) / 1e9)
        best = r if best is None or r[0] < best[0] else best
    return best


def run_workload(name, wl, steps, warmup, ctx, want_cpu, want_e2e, main):
    """One workload on this rank's GPU (all ranks call this in lock step).  Returns the result dict on rank 0, else None."""
    torch, dist, bf = ctx.torch, ctx.dist, ctx.bf
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    tables = bf.tables
    mic_xy = bf.GEOMETRIES[wl["mics"]]
    M, B, T, P = len(mic_xy), wl["n_streams"], wl["hops"], wl.get("passes", 1)
    Hh = wl.get("hop", 512)
    L = T * Hh
    srp = "srp_dirs" in wl
    D = wl.get("srp_dirs", 0)
    config = workload_config(wl, tables)
    if srp:
        config["collective"] = "all_gather of maps (NCCL)" if world > 1 else "none"
        from beamform_b200.shard import gather_maps
        thetas = (-180.0 + 360.0 * np.arange(D) / D).astype(np.float32)
    cfg = bf.make_config(wl["algo"], mics=wl["mics"], hop=Hh, interferers=wl.get("interferers", ()), device=ctx.local_rank, **wl.get("params", {}))
    beam = bf.Beamformer(cfg, n_streams=B)
    x = device_synth(torch, mic_xy, B, L, seed=0xBEA4F0 + 1000 * rank, device=dev)
    y = torch.empty((B, T, D) if srp else (B, L), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()

    def one_pass():
        if srp:
            beam.srp_device(x.data_ptr(), thetas, y.data_ptr(), T, stream_ptr=stream.cuda_stream)
            return gather_maps(y, world * B) if world > 1 else y
        beam.process_device(x.data_ptr(), y.data_ptr(), T, stream_ptr=stream.cuda_stream)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup * P):
        one_pass()
    fence()
    sampler = ClockSampler(ctx.local_rank) if main else None
    if sampler:
        sampler.start()
    beam.set_profiling(True)
    l0 = beam.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps * P):
        one_pass()
    e1.record(stream)
    fence()
    clocks = sampler.stop() if sampler else None
    launches = beam.kernel_launches - l0
    kern_ms, kern_n = beam.get_profile()
    beam.set_profiling(False)
    t_max = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_max.item())
    audio_s_pass = world * B * L / SR
    value = audio_s_pass * P * steps / (elapsed_ms * 1e-3)
    ms_per_pass = elapsed_ms / (steps * P)

    # ---- e2e: host buffers through the C ABI, H2D + kernels + D2H inside the timed region ----
    e2e, e2e_err = None, None
    if want_e2e:
        xh = yh = None
        try:   # pinned host buffers: with 8 ranks on one box this is several GB per rank; a failure must not lose the run
            xh = torch.empty((B, M, L), dtype=torch.float32, pin_memory=True)
            yh = torch.empty(tuple(y.shape), dtype=torch.float32, pin_memory=True)
        except Exception as ex:
            e2e_err = repr(ex)[:200]
        ok = torch.tensor([0 if e2e_err else 1], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # every rank takes the same branch
        if int(ok.item()) == 1:
            xh.copy_(x)
            ev, nev = bf.make_events([])
            lib = bf.lib()
            xd = torch.empty_like(x) if srp else None

            def e2e_step():
                if srp:
                    xd.copy_(xh, non_blocking=True)
                    beam.srp_device(xd.data_ptr(), thetas, y.data_ptr(), T, stream_ptr=stream.cuda_stream)
                    yh.copy_(y, non_blocking=True)
                    torch.cuda.synchronize()
                else:
                    rc = lib.bf_process_batch(beam._h, xh.data_ptr(), M * L, L, yh.data_ptr(), L, T, ev, nev)
                    assert rc == 0, lib.bf_last_error()

            e2e_step()
            fence()
            n_e2e = 3 if main else 2
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            e2e = {"value": audio_s_pass * n_e2e / float(t_e.item()), "unit": "audio-s/s", "h2d_bytes_per_step": B * M * L * 4,
                   "d2h_bytes_per_step": int(yh.numel()) * 4, "steps": n_e2e,
                   "note": "one e2e step = one pass (one host batch) through bf_process_batch" if not srp else "one e2e step = H2D + sweep + D2H of the maps",
                   "checksum": float(yh.double().sum()) if srp else float(yh[:, -Hh:].double().abs().sum())}
            if main:   # what the box's pinned copies alone allow, all ranks copying at once
                fence()
                dtc, h2d, d2h = copy_ceiling(torch, xh, xd if srp else x, yh, y)
                t_c = torch.tensor([dtc], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
                ceil = audio_s_pass / float(t_c.item())
                e2e["copy_ceiling"] = {"value": ceil, "unit": "audio-s/s", "h2d_gbs_rank0": h2d, "d2h_gbs_rank0": d2h, "frac": e2e["value"] / ceil,
                                       "how": "the step's pinned H2D and D2H copies alone, concurrently on two streams, every rank at once, max over ranks"}
        elif e2e_err is None:
            e2e_err = "pinned host allocation failed on another rank"
        del xh, yh

    stats = None
    if rank == 0 and main and wl["algo"] in ("mvdr", "lcmv", "gss"):
        # measured selection density of the workload (outside the timed region): the per-bin solves only run for the
        # (bin, frame) items that pass the magnitude gate, so the cost of these nodes scales with it
        try:
            nb = min(B, 16)
            flags = torch.zeros((nb, T, 2 * Hh), dtype=torch.uint8, device=dev)
            probe = bf.Beamformer(cfg, n_streams=nb)
            probe.set_capture(flags.data_ptr())
            probe.process_device(x.data_ptr(), y.data_ptr(), T, stream_ptr=stream.cuda_stream, in_stream_stride=M * L, in_mic_stride=L, out_stream_stride=L)
            torch.cuda.synchronize()
            per_frame = float((flags & 1)[:, :, :Hh + 1].float().sum().item() / (nb * T))
            stats = {"selected_items_per_frame": per_frame, "selected_fraction_of_half_spectrum": per_frame / (Hh + 1), "probe_streams": nb}
            del probe, flags
        except Exception as ex:
            stats = {"error": repr(ex)[:120]}

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # hbm-bound workloads: the dominant kernel's launch; the sweep: both kernels of a step (spectra images + contraction)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_%s.json" % name)))
            traffic = tj.get("dram_bytes_all_kernels_per_step" if srp else "dram_bytes_per_launch")
        except Exception:
            pass
        if srp:
            peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
            alg = 8.0 * D * M * 513 * B * T          # 8*D*M*(N/2+1) FLOP per frame
            achieved = alg / (ms_per_pass * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_ratio": (traffic / (B * M * L * 4.0 + B * T * D * 4.0)) if traffic else None,
                    "traffic_note": "DRAM bytes of both kernels of a step (ncu): the spectra leave srp_spectra_kernel as bf16 hi/lo operand images and are read back once",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400",
                    "kernel": "srp_power_tc_kernel (tcgen05 BF16x3) + srp_spectra_kernel", "kernel_ms_per_launch": ms_per_pass,
                    "algorithmic_flops_per_launch": alg,
                    "note": "algorithmic FLOPs (8*D*M*513 per frame); the BF16x3 split issues 3x as many tensor FLOPs (K padded 2*M -> 128)"}
        else:
            peak = float(peaks.get("hbm_gbs", 6650.0))
            alg = (wl.get("alg_channels", M) + 1) * 4.0 * B * L
            k_ms = kern_ms / max(1, kern_n)
            achieved = alg / (k_ms * 1e-3) / 1e9 if kern_n else None
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": traffic, "traffic_ratio": (traffic / alg) if traffic else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                    "frac_of_nominal_8000_gbs": (achieved / 8000.0) if achieved else None,   # the ~8 TB/s the north_star quotes
                    "kernel": wl.get("kernel"), "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg}
        cpu = None
        if want_cpu and world == 1:
            cpu = cpu_baseline(wl, tables, ctx.synth, ctx.cores, steps=3 if main else 2)
        out = {"value": value, "unit": "audio-s/s", "ms_per_step": elapsed_ms / steps, "steps": steps, "warmup": warmup,
               "dtype": "f32 (lcmv solves f64)" if wl["algo"] == "lcmv" else ("bf16x3 split of f32 (tensor cores), f32 accumulate" if srp else "f32"),
               "config": config, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches}
        if clocks:
            out["clocks"] = clocks
        if stats:
            out["workload_stats"] = stats
        if e2e_err:
            out["e2e_error"] = e2e_err
    beam.close()
    del beam, x, y
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--hops", type=int, default=0, help="override hops per pass")
    ap.add_argument("--passes", type=int, default=0, help="override passes per step")
    ap.add_argument("--param", action="append", default=[], metavar="KEY=VALUE", help="override a node parameter of the workload (launch-file key)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the brief runs of the other BASELINE configs (key `workloads`)")
    args = ap.parse_args()

    wl = dict(WORKLOADS[args.workload])
    if args.streams:
        wl["n_streams"] = args.streams
    if args.hops:
        wl["hops"] = args.hops
    if args.passes:
        wl["passes"] = args.passes
    if args.param:
        wl["params"] = dict(wl.get("params", {}), **{k: float(v) for k, v in (kv.split("=", 1) for kv in args.param)})
        wl["name"] += " [" + ", ".join(args.param) + "]"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        tables = load_by_path("bf_tables", "beamform_b200/tables.py")
        synth = load_by_path("bf_synth", "beamform_b200/synth.py")
        return reference_arm(args, wl, tables, synth, cores)

    import torch
    import torch.distributed as dist
    import beamform_b200 as bf
    from beamform_b200 import synth, tables
    bf.tables = tables
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.bf, ctx.synth = torch, dist, bf, synth
    ctx.dev = torch.device("cuda", local_rank)
    ctx.world, ctx.rank, ctx.local_rank, ctx.cores = world, rank, local_rank, cores
    if world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)

    res = run_workload(args.workload, wl, args.steps, args.warmup, ctx, want_cpu=not args.no_cpu, want_e2e=not args.no_e2e, main=True)
    extra = {}
    if args.workload == "c2" and not args.no_extra and not (args.streams or args.hops or args.param):
        k2, w2 = max(2, args.steps // 5), 3
        for nm in EXTRA:
            w = dict(WORKLOADS[nm])
            w["passes"] = max(1, w["passes"] // 4)   # brief: a quarter of the passes per step
            try:
                r = run_workload(nm, w, k2, w2, ctx, want_cpu=not args.no_cpu, want_e2e=not args.no_e2e, main=False)
            except Exception as ex:   # a failing side workload must not lose the headline line (every rank fails alike)
                r = {"error": repr(ex)[:300]}
            if rank == 0 and r is not None:
                roof = r.get("roofline") or {}
                r["frac"] = roof.get("frac")
                r["ms"] = roof.get("kernel_ms_per_launch")
                r["traffic_ratio"] = roof.get("traffic_ratio")
                extra[nm] = r
    if rank == 0:
        line = {"metric": "beamformed audio-sec/sec", "value": res["value"], "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": res["dtype"], "data": "synthetic"}
        for k in ("config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "workload_stats", "e2e_error"):
            if k in res:
                line[k] = res[k]
        if extra:
            line["workloads"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
