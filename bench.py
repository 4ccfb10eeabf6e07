#!/usr/bin/env python
"""bench.py -- lossless ALPC encode throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU baseline: the oracle port on host cores
    python bench.py --config {2,3,4,5}                       # which BASELINE config is the workload

A step = one pass of the whole encode path over one batch of synthetic PCM.  Workloads (BASELINE.json configs):
  2 (default at N = 1): one 1-hour 44.1 kHz 16-bit stereo multitone+noise stream (3600 frames), level 5.
  3 (default at N > 1): batch corpus of 3-minute 44.1 kHz stereo tracks sharded over ranks, weak scaling:
      every rank encodes --tracks (default 20) x 180 s per step; no collective on the data path, one
      all_gather of per-rank byte lengths per step for the final concatenation.
  4: hi-res 96 kHz stereo sweep+noise, bit_depth 24 in the header, level 9 (maximum LPC order), --seconds long.
  5: 8 kHz mono speech-like signal, 4096 tracks of 8 s (small frames), level 5.
Every workload enters through Encoder::encode's documented input (interleaved f32).
`value` = PCM GB/s (2 bytes x interleaved samples / time; 3 bytes for the nominal 24-bit config 4) over all
ranks with inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI call (H2D of the f32
samples and D2H of the .flo bytes inside the timed region).  `parity` compares the timed output, frame by
frame, with the CPU oracle on the sample the CPU baseline encodes anyway; a mismatch fails the run.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

SEED = 0xF10 + 2

# BASELINE.json configs: (sample_rate, channels, header bit depth, level, signal, noise lsb, nominal PCM bytes / sample)
CONFIGS = {
    2: dict(sr=44100, ch=2, bits=16, level=5, kind="multitone", noise=64, pcm_bytes=2,
            name="1-hour 44.1 kHz 16-bit stereo synthetic multitone+noise, single stream (BASELINE config 2)"),
    3: dict(sr=44100, ch=2, bits=16, level=5, kind="multitone", noise=64, pcm_bytes=2,
            name="batch corpus shard of 3-minute 44.1 kHz 16-bit stereo tracks (BASELINE config 3)"),
    4: dict(sr=96000, ch=2, bits=24, level=9, kind="sweep", noise=32, pcm_bytes=3,
            name="hi-res 96 kHz 24-bit stereo synthetic sweep+noise at maximum LPC order (BASELINE config 4)"),
    5: dict(sr=8000, ch=1, bits=16, level=5, kind="speech", noise=16, pcm_bytes=2,
            name="8 kHz mono telephone-band speech-like signal, many short tracks (BASELINE config 5)"),
}


def track_lengths(cfg_id: int, args, world: int) -> list[int]:
    """Sample frames per track of the WHOLE job (all ranks)."""
    c = CONFIGS[cfg_id]
    if cfg_id == 2:
        return [args.seconds * c["sr"]] * world
    if cfg_id == 3:
        return [180 * c["sr"]] * (args.tracks * world)
    if cfg_id == 4:
        return [args.seconds * c["sr"]] * world
    return [8 * c["sr"]] * (4096 * world)


def peaks() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        if os.environ.get("FLO_BENCH_NO_CLOCKS"):      # diagnosis only: a line without clocks is not a valid bench line
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        """Summary of the samples taken inside [t0, t1] (the timed region); if the region was shorter than
        the sampling period, of the samples since `t0 - 1 s` (warm-up steps of the same workload)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for t, r in self.rows if t0 <= t <= t1]
        window = "timed region"
        if len(inside) < 3:
            inside = [r for t, r in self.rows if t0 - 1.0 <= t <= t1 + 0.05]
            window = "timed region + preceding warm-up (region shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import flo_oracle as oracle                      # checker / CPU baseline only
    oracle.lib()
    return oracle


def cpu_encode(jobs, sr: int, ch: int, bits: int, level: int, threads: int):
    """Times the oracle (CPU port of the reference algorithm) on `jobs` (interleaved f32 arrays made of whole
    frames; frames are independent in the reference: encoder.rs:53-61), `threads` jobs at a time.
    Returns (seconds of wall clock, the .flo images)."""
    from concurrent.futures import ThreadPoolExecutor
    oracle = oracle_mod()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda s: oracle.encode(s, sr, ch, bits, level, b""), jobs))
    return time.perf_counter() - t0, outs


def flo_frames(img) -> list[tuple[int, int]]:
    """(position, size) of every frame of a .flo image, from its TOC (writer.rs:39-100, 193-224)."""
    mv = memoryview(img)
    assert bytes(mv[:4]) == b"FLO!", "not a flo image"
    toc_size = struct.unpack_from("<Q", mv, 38)[0]
    n = struct.unpack_from("<I", mv, 70)[0]
    data0 = 70 + toc_size
    out = []
    for i in range(n):
        _idx, off, size, _ts = struct.unpack_from("<IQII", mv, 74 + 20 * i)
        out.append((data0 + off, size))
    return out


def cpu_jobs_for(cfg_id: int, c: dict, pcm_tracks, threads: int, sample_s: int):
    """The bounded CPU sample of the workload: a list of (track index, first frame, f32 samples of whole frames)."""
    import numpy as np
    sr, ch = c["sr"], c["ch"]
    per = sr * ch
    jobs = []
    if cfg_id in (2, 4):
        # slices of the first `sample_s` seconds of the stream, one per thread, each a run of whole frames
        total_s = pcm_tracks[0].numel() // per
        sample_s = min(sample_s, total_s)
        run = max(1, sample_s // threads)
        for i in range(min(threads, max(1, sample_s // run))):
            seg = pcm_tracks[0][i * run * per:(i + 1) * run * per].cpu().numpy()
            jobs.append((0, i * run, seg.astype(np.float32) * np.float32(1 / 32768)))
    else:
        # whole tracks
        per_track_s = pcm_tracks[0].numel() // per
        ntr = max(threads, min(len(pcm_tracks), max(1, sample_s // max(1, per_track_s))))
        ntr = min(ntr, len(pcm_tracks))
        for t in range(ntr):
            jobs.append((t, 0, pcm_tracks[t].cpu().numpy().astype(np.float32) * np.float32(1 / 32768)))
    return jobs


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5], help="BASELINE config (0 = 2 at N=1, 3 at N>1)")
    ap.add_argument("--seconds", type=int, default=3600, help="audio seconds per GPU per step (configs 2 and 4)")
    ap.add_argument("--tracks", type=int, default=20, help="config 3: tracks of 180 s per GPU per step")
    ap.add_argument("--level", type=int, default=-1, help="override the config's compression level")
    ap.add_argument("--cpu-sample-seconds", type=int, default=0, help="0 = auto (about 15 s of CPU work)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-level / other-config side measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    cfg_id = args.config or (2 if world == 1 else 3)
    c = CONFIGS[cfg_id]
    SR, CH, BITS = c["sr"], c["ch"], c["bits"]
    level = c["level"] if args.level < 0 else args.level
    peak, peak_src = peaks()
    corpus_n = track_lengths(cfg_id, args, world)
    per_gpu_tracks = len(corpus_n) // world
    per_gpu_seconds = sum(corpus_n[:per_gpu_tracks]) / SR
    workload = c["name"] + (f": {per_gpu_tracks} x {corpus_n[0] // SR} s per GPU per step" if cfg_id in (3, 5) else "")
    config = {"workload": workload, "baseline_config": cfg_id, "audio_seconds_per_gpu_per_step": per_gpu_seconds,
              "tracks_per_gpu_per_step": per_gpu_tracks, "sample_rate": SR, "channels": CH, "bit_depth": BITS,
              "level": level, "entry": "Encoder::encode (interleaved f32)",
              "frames_per_gpu_per_step": sum(-(-n // SR) for n in corpus_n[:per_gpu_tracks]),
              "l2": "inputs (%.2f GB f32 per GPU per step) exceed the 126 MB L2; no flush needed" % (4e-9 * sum(corpus_n[:per_gpu_tracks]) * CH),
              "parallelism": f"{n_gpus} x independent shards, no data-path collective"}

    import numpy as np

    # ---------------------------------------------------------------- reference arm (CPU) ----------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import synth_pcm16 as np_synth
        threads = host_threads()
        sample_s = args.cpu_sample_seconds or (60 * threads if cfg_id != 4 else 12 * threads)
        per = SR * CH
        if cfg_id in (2, 4):
            run = max(1, sample_s // threads)
            pcm = np_synth(run * threads * SR, CH, SR, seed=SEED, kind=c["kind"], noise_lsb=c["noise"])
            jobs = [pcm[i * run * per:(i + 1) * run * per].astype(np.float32) * np.float32(1 / 32768) for i in range(threads)]
        else:
            tlen = corpus_n[0]
            ntr = max(threads, sample_s // max(1, tlen // SR))
            jobs = [np_synth(tlen, CH, SR, seed=SEED + 131 * t, kind=c["kind"], noise_lsb=c["noise"]).astype(np.float32) * np.float32(1 / 32768)
                    for t in range(ntr)]
        secs = sum(j.size for j in jobs) / per
        for _ in range(min(args.warmup, 1)):
            cpu_encode(jobs[:threads], SR, CH, BITS, level, threads)
        dts = []
        for _ in range(args.steps):
            dt, _outs = cpu_encode(jobs, SR, CH, BITS, level, threads)
            dts.append(dt)
        dt = sum(dts) / len(dts)
        samples = secs * per
        val = c["pcm_bytes"] * samples / dt / 1e9
        line = {"impl": "reference", "metric": "lossless encode PCM throughput", "value": val, "unit": "GB/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                "audio_seconds_per_step": secs,
                "ms_per_step_scaled_to_workload": dt * 1e3 * per_gpu_seconds / secs,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 Levinson",
                "data": "synthetic", "config": config, "x_realtime": secs / dt,
                "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                                 "sample": f"{secs:.0f} s of the workload as {len(jobs)} independent jobs of whole frames, "
                                           f"{threads} at a time, one thread each (ms_per_step is the time of this sample); "
                                           "C oracle (literal restatement of the reference encoder, gcc -O3), "
                                           "the Rust reference cannot be built here (no cargo)"},
                "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- CUDA arm -----------------------
    import torch
    import torch.distributed as dist
    import flo_b200
    import synth_torch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one block of host cores per rank: the ranks' copy threads and pinned buffers do not share cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            k = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * k:(local_rank + 1) * k]) or set(cores))
        except Exception:
            pass
    ctx = flo_b200.Context(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # 20 ms period; the summary only uses samples inside the timed region
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # synthetic PCM directly in HBM (integer-only generator, identical to tests/helpers.synth_pcm16)
    from flo_b200 import shard
    ranges = shard.partition_tracks([shard.frames_of_track(n * CH, SR, CH) for n in corpus_n], world)
    t_lo, t_hi = ranges[rank]                                                       # this rank's contiguous shard
    tracks_n = corpus_n[t_lo:t_hi]
    pcm_tracks = [synth_torch.synth_pcm16_long(n, CH, SR, SEED + 131 * (t_lo + i), c["kind"], c["noise"], dev)
                  for i, n in enumerate(tracks_n)]
    f32_tracks = [p.to(torch.float32) * (1.0 / 32768.0) for p in pcm_tracks]    # reflo/src/audio.rs:247-254 (exact)
    n_list = [int(t.numel()) for t in f32_tracks]
    nt = len(n_list)
    total_inter = sum(n_list)
    bound = ctx.output_bound(n_list, [SR] * nt, [CH] * nt)
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    ptrs = [t.data_ptr() for t in f32_tracks]

    # The exchange of a step is issued by a helper thread while the main thread is already inside the next encode call
    # (ctypes releases the GIL for the call's 3 ms): its ~0.2 ms of tensor set-up and NCCL launch no longer leave the GPU
    # idle between steps.  One helper, so every rank issues its all_gathers in step order.
    xpool = None
    if world > 1:
        from concurrent.futures import ThreadPoolExecutor
        xpool = ThreadPoolExecutor(max_workers=1, initializer=lambda: torch.cuda.set_device(dev))

    def step():
        off, ln = ctx.encode_batch_device(ptrs, n_list, [SR] * nt, [CH] * nt, [BITS] * nt, d_out.data_ptr(), bound, level=level)
        if world > 1:       # per-track byte lengths for the final concatenation (the only exchange), non-blocking
            pending.append(xpool.submit(shard.exchange_lengths_async, [int(v) for v in ln], ranges))
        return off, ln

    pending = []

    def sync_all():
        while pending:      # the concatenation offsets of every step are complete before the clock stops
            pending.pop().result().result()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        off, ln = step()
    out_bytes = int(ln.sum())
    sync_all()
    t_region0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_ms, dev_ms, launches = [], [], 0
    ev0.record()
    for _ in range(args.steps):
        off, ln = step()
        t = ctx.last_timing()
        enc_ms.append(t["encode_ms"]); dev_ms.append(t["device_ms"]); launches += t["launches"]
    ev1.record()
    sync_all()
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None
    counters = ctx.last_counters()
    ms_total = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = c["pcm_bytes"] * total_inter * world / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (k_encode_frames): algorithmic bytes = f32 in + .flo out
    k_ms = sum(enc_ms) / len(enc_ms)
    alg_bytes = 4.0 * total_inter + out_bytes
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None                      # DRAM bytes per launch from the committed ncu capture of this exact workload
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and cfg_id == 2 and tj["frames"] == args.seconds and tj["level"] == level:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_encode_frames", "kernel_ms": k_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "latency / issue bound, not bandwidth bound (exhaustive candidate search of the reference: "
                        "~190 thread-instructions per channel-sample at level 5, FP64-pipe FIR); DRAM traffic stays "
                        "near the algorithmic bytes (input read once, 16-bit planes live in L2); see DESIGN.md section 6"}

    # ---- byte parity of the timed output against the CPU oracle + CPU baseline on the same sample ----
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        threads = host_threads()
        sample_s = args.cpu_sample_seconds or (60 * threads if cfg_id != 4 else 12 * threads)
        jobs = cpu_jobs_for(cfg_id, c, pcm_tracks, threads, sample_s)
        dt, outs = cpu_encode([j[2] for j in jobs], SR, CH, BITS, level, threads)
        secs = sum(j[2].size for j in jobs) / (SR * CH)
        host_img = d_out.cpu().numpy()
        gpu_frames = {}
        checked, same = 0, True
        first_bad = None
        for (t, f0, _x), ref in zip(jobs, outs):
            if t not in gpu_frames:
                gpu_frames[t] = (int(off[t]), flo_frames(host_img[int(off[t]):int(off[t]) + int(ln[t])]))
            base, gfr = gpu_frames[t]
            for i, (rp, rs) in enumerate(flo_frames(ref)):
                gp, gs = gfr[f0 + i]
                ok = gs == rs and bytes(host_img[base + gp:base + gp + gs]) == ref[rp:rp + rs]
                checked += 1
                if not ok and same:
                    same, first_bad = False, {"track": t, "frame": f0 + i}
        parity = {"checker": "CPU oracle (C restatement of the reference encoder)", "frames_checked": checked,
                  "audio_seconds_checked": secs, "identical": same, "first_mismatch": first_bad}
        # single-thread figure: what one Encoder::encode call is in the reference
        x1 = jobs[0][2][:min(jobs[0][2].size, 20 * SR * CH)]
        dt1, _ = cpu_encode([x1], SR, CH, BITS, level, 1)
        cpu = {"value": c["pcm_bytes"] * secs * SR * CH / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"{secs:.0f} s of the same input as {len(jobs)} jobs of whole frames, {threads} at a time, one thread each "
                         "(C oracle, gcc -O3)",
               "x_realtime": secs / dt,
               "single_thread": {"value": c["pcm_bytes"] * x1.size / dt1 / 1e9, "unit": "GB/s", "x_realtime": x1.size / (SR * CH) / dt1,
                                 "sample": f"{x1.size // (SR * CH)} s, one thread (one Encoder::encode call of the reference is single-threaded)"}}
        del host_img

    # side measurements (not the headline): other compression levels, the PCM16 entry, the other BASELINE configs
    extras = None
    if world == 1 and cfg_id == 2 and not args.no_extras:
        extras = {"note": "encode-kernel ms per step and % of measured HBM peak on the same 1-hour stream; "
                          "levels 0-3 try fixed predictors only (encoder.rs:204)", "levels": {}}
        for lv in (0, 2, 4, 5, 7, 9):
            ts = []
            for _ in range(3):
                o2, l2 = ctx.encode_batch_device(ptrs, n_list, [SR], [CH], [16], d_out.data_ptr(), bound, level=lv)
                ts.append(ctx.last_timing()["encode_ms"])
            ab = 4.0 * total_inter + float(l2.sum())
            extras["levels"][str(lv)] = {"kernel_ms": min(ts), "flo_bytes": int(l2.sum()),
                                         "pct_of_hbm_peak": 100.0 * ab / (min(ts) * 1e-3) / 1e9 / peak}
        ts = []
        for _ in range(3):
            o2, l2 = ctx.encode_batch_device([pcm_tracks[0].data_ptr()], n_list, [SR], [CH], [16], d_out.data_ptr(), bound,
                                             level=level, fmt=flo_b200.FMT_PCM16)
            ts.append(ctx.last_timing()["encode_ms"])
        extras["pcm16_entry"] = {"kernel_ms": min(ts), "same_bytes_as_f32_entry": int(l2.sum()) == out_bytes,
                                 "pct_of_hbm_peak": 100.0 * (2.0 * total_inter + float(l2.sum())) / (min(ts) * 1e-3) / 1e9 / peak}

        # companion row N2: decode the file just written, device-resident, and check the round trip
        d_dec = torch.empty(total_inter, dtype=torch.float32, device=dev)
        ts = []
        for _ in range(3):
            n_dec, _info = ctx.decode_device(d_out.data_ptr() + int(o2[0]), int(l2[0]), d_dec.data_ptr(), d_dec.numel())
            ts.append(ctx.last_timing())
        q = torch.trunc(torch.clamp(f32_tracks[0] * 32767.0, -32767.0, 32767.0)) * torch.tensor(1.0 / 32767.0, dtype=torch.float32, device=dev)
        k_dec = min(t["encode_ms"] for t in ts)
        extras["decode"] = {"kernel": "k_dec_units", "kernel_ms": k_dec, "device_pass_ms": min(t["device_ms"] for t in ts),
                            "launches": ts[-1]["launches"], "round_trip_exact": bool(n_dec == total_inter and torch.equal(d_dec, q)),
                            "pct_of_hbm_peak": 100.0 * (4.0 * total_inter + float(l2[0])) / (k_dec * 1e-3) / 1e9 / peak}
        del d_dec, q

        # companion row N4: the analysis metadata libflo::encode() attaches, on the same resident f32 stream (wall time of
        # the C-ABI call incl. its synchronisation; parity with the oracle: tests/test_gpu_analysis.py)
        from flo_b200 import analysis as fa
        cap = fa.peaks_count(total_inter, SR, CH, 50)
        d_peaks = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        x0 = f32_tracks[0]

        def best_ms(f, reps=3):
            f(); torch.cuda.synchronize()
            best = None
            for _ in range(reps):
                t0 = time.perf_counter(); f(); torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0) if best is not None else time.perf_counter() - t0
            return best * 1e3
        lufs_box = [None]
        pk_ms = best_ms(lambda: fa.extract_waveform_peaks_device(x0.data_ptr(), x0.numel(), CH, SR, 50, d_peaks.data_ptr(), cap, ctx=ctx))
        lu_ms = best_ms(lambda: lufs_box.__setitem__(0, fa.integrated_loudness_device(x0.data_ptr(), x0.numel(), CH, SR, ctx=ctx)))
        extras["analysis"] = {"waveform_peaks": {"peaks": cap, "call_ms": pk_ms, "pct_of_hbm_peak": 100.0 * 4.0 * x0.numel() / (pk_ms * 1e-3) / 1e9 / peak},
                              "r128_integrated_loudness": {"lufs": lufs_box[0], "call_ms": lu_ms, "passes_over_the_input": 2}}
        del d_peaks

        # the other BASELINE configs, device-resident (their own bench lines: bench.py --config 3|4|5)
        extras["configs"] = {}
        for cid, kw in ((3, dict(tracks=20)), (4, dict(seconds=1184)), (5, dict())):
            cc = CONFIGS[cid]

            lens = track_lengths(cid, argparse.Namespace(seconds=kw.get("seconds", 3600), tracks=kw.get("tracks", 20)), 1)
            pcs = [synth_torch.synth_pcm16_long(n, cc["ch"], cc["sr"], SEED + 131 * i, cc["kind"], cc["noise"], dev) for i, n in enumerate(lens)]
            xs = [p.to(torch.float32) * (1.0 / 32768.0) for p in pcs]
            nn = [int(t.numel()) for t in xs]
            bb = ctx.output_bound(nn, [cc["sr"]] * len(nn), [cc["ch"]] * len(nn))
            oo = torch.empty(bb, dtype=torch.uint8, device=dev)
            best = None
            for _ in range(4):
                _o, l3 = ctx.encode_batch_device([t.data_ptr() for t in xs], nn, [cc["sr"]] * len(nn), [cc["ch"]] * len(nn),
                                                 [cc["bits"]] * len(nn), oo.data_ptr(), bb, level=cc["level"])
                tt = ctx.last_timing()
                best = tt if best is None or tt["device_ms"] < best["device_ms"] else best
            tot = sum(nn)
            extras["configs"][str(cid)] = {
                "workload": cc["name"], "tracks": len(nn), "audio_seconds": sum(lens) / cc["sr"], "level": cc["level"],
                "device_ms": best["device_ms"], "encode_kernel_ms": best["encode_ms"],
                "value_GBps": cc["pcm_bytes"] * tot / best["device_ms"] / 1e6, "x_realtime": sum(lens) / cc["sr"] / best["device_ms"] * 1e3,
                "pct_of_hbm_peak_encode_kernel": 100.0 * (4.0 * tot + float(l3.sum())) / best["encode_ms"] / 1e6 / peak,
                "compression_ratio": 2.0 * tot / float(l3.sum())}
            del pcs, xs, oo

    # e2e: host buffers through the reference-facing C-ABI call (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        host_in = [torch.empty(n, dtype=torch.float32, pin_memory=True) for n in n_list]
        for h, d in zip(host_in, f32_tracks):
            h.copy_(d)
        torch.cuda.synchronize()
        specs = [flo_b200.TrackSpec(h.numpy(), SR, CH, BITS, b"") for h in host_in]
        for _ in range(2):                                         # warm-up (arena growth, pinned output pool)
            with ctx.encode_batch(specs, level, views=True) as res:
                e2e_out = res.total_bytes()
        sync_all()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 5))
        for _ in range(reps):
            # the call returns when the .flo images are in host memory; they are read (first/last byte) and released
            with ctx.encode_batch(specs, level, views=True) as res:
                assert all(a[0] == 0x46 and a[-1] is not None for a in res.arrays)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        tm = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())

        # what the copies alone cost on this box with all ranks copying at once: the same bytes host -> device and
        # device -> host as plain cudaMemcpyAsync on two streams (no kernel) -- the ceiling e2e can reach
        h_out = torch.empty(e2e_out, dtype=torch.uint8, pin_memory=True)
        d_in = torch.empty(total_inter, dtype=torch.float32, device=dev)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        big = torch.cat([h.view(-1) for h in host_in]) if len(host_in) > 1 else host_in[0]
        big = big.pin_memory() if not big.is_pinned() else big

        def copies():
            with torch.cuda.stream(s_in):
                d_in.copy_(big, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out[:e2e_out], non_blocking=True)
            s_in.synchronize(); s_out.synchronize()
        copies()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            copies()
        dtc = (time.perf_counter() - t0) / reps
        tc = torch.tensor([dtc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceil_s = float(tc.item())
        del h_out, d_in, big

        e2e = {"value": c["pcm_bytes"] * total_inter * world / e2e_s / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 4 * total_inter, "d2h_bytes_per_step": e2e_out,
               "ms_per_step": e2e_s * 1e3, "x_realtime": per_gpu_seconds * world / e2e_s,
               "pcie_ceiling_gbs": c["pcm_bytes"] * total_inter * world / ceil_s / 1e9,
               "pcie_ceiling_ms_per_step": ceil_s * 1e3, "frac_of_pcie_ceiling": ceil_s / e2e_s,
               "pcie_ceiling_note": "the same H2D and D2H bytes as plain cudaMemcpyAsync from / to pinned memory on two "
                                    "streams, all ranks at once, no kernel"}

        # the same call from pageable memory (what a Rust &[f32] is): staged through the library's pinned ring
        if world == 1:
            pag = [np.array(h.numpy(), copy=True) for h in host_in]
            specs_p = [flo_b200.TrackSpec(a, SR, CH, BITS, b"") for a in pag]
            for _ in range(2):
                with ctx.encode_batch(specs_p, level, views=True) as res:
                    same_p = res.total_bytes() == e2e_out
            t0 = time.perf_counter()
            for _ in range(reps):
                with ctx.encode_batch(specs_p, level, views=True) as res:
                    pass
            dtp = (time.perf_counter() - t0) / reps
            e2e["pageable_input"] = {"value": c["pcm_bytes"] * total_inter / dtp / 1e9, "unit": "GB/s", "ms_per_step": dtp * 1e3,
                                     "same_bytes": bool(same_p), "slowdown_vs_pinned": dtp / e2e_s}
            del pag, specs_p

        if extras is not None and world == 1:
            # same call with reflo's 16-bit PCM as the host buffer (flo_encode_pcm16): half the H2D bytes
            host_pcm = [torch.empty(n, dtype=torch.int16, pin_memory=True) for n in n_list]
            for h, d in zip(host_pcm, pcm_tracks):
                h.copy_(d)
            torch.cuda.synchronize()
            specs16 = [flo_b200.TrackSpec(h.numpy(), SR, CH, 16, b"") for h in host_pcm]
            for _ in range(2):
                with ctx.encode_batch(specs16, level, flo_b200.FMT_PCM16, views=True) as res:
                    same = res.total_bytes() == e2e_out
            t0 = time.perf_counter()
            for _ in range(reps):
                with ctx.encode_batch(specs16, level, flo_b200.FMT_PCM16, views=True) as res:
                    pass
            dt16 = (time.perf_counter() - t0) / reps
            extras["e2e_pcm16_entry"] = {"value": 2.0 * total_inter / dt16 / 1e9, "unit": "GB/s", "ms_per_step": dt16 * 1e3,
                                         "h2d_bytes_per_step": 2 * total_inter, "same_bytes_as_f32_entry": bool(same)}
            del host_pcm, specs16
        del host_in, specs

    rc = 0
    if rank == 0:
        line = {"metric": "lossless encode PCM throughput", "value": value, "unit": "GB/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 Levinson", "data": "synthetic",
                "config": config, "x_realtime": per_gpu_seconds * world / (ms_step * 1e-3),
                "pct_of_hbm_peak": 100.0 * achieved / peak, "flo_bytes_per_step_per_gpu": out_bytes,
                "compression_ratio": 2.0 * total_inter / out_bytes, "roofline": roofline, "cpu_baseline": cpu,
                "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "device_ms_per_step": sum(dev_ms) / len(dev_ms), "analysis_counters_last_step": counters,
                "extras": extras}
        print(json.dumps(line))
        if parity is not None and not parity["identical"]:
            print("bench.py: the timed output differs from the CPU oracle: " + json.dumps(parity), file=sys.stderr)
            rc = 3
    if world > 1:
        if xpool is not None:
            xpool.shutdown()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
