#!/usr/bin/env python
"""bench.py -- lossless ALPC encode throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU baseline: the oracle port on host cores

A step = one pass of the whole encode path over one batch of synthetic PCM:
  N = 1 : BASELINE config 2 -- one 1-hour 44.1 kHz 16-bit stereo multitone+noise stream (3600 frames),
          given to Encoder::encode as interleaved f32 (the documented entry), level 5.
  N > 1 : BASELINE config 3 (batch corpus of 3-min tracks) sharded over ranks, weak scaling: every rank
          encodes 20 x 180 s tracks per step (the same 3600 frames / GPU as N = 1); no collective on the
          data path, one all_gather of per-rank byte lengths per step for the final concatenation.
`value` = PCM GB/s (2 bytes x interleaved samples / time) over all ranks with inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (H2D of the f32 samples and D2H of the .flo bytes
inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

SR, CH, LEVEL = 44100, 2, 5
TRACK_SECONDS = 180
SEED = 0xF10 + 2


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


def cpu_encode_seconds(pcm_np, seconds: int, threads: int, level: int = LEVEL):
    """Times the oracle (CPU port of the reference algorithm) on `seconds` one-second slices of the
    stream, `threads` slices at a time (frames are independent in the reference: encoder.rs:53-61)."""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import flo_oracle as oracle                      # checker / CPU baseline only
    oracle.lib()
    per = SR * CH
    # one slice per thread, each a contiguous run of whole frames
    seconds = max(threads, seconds // threads * threads)
    run = seconds // threads
    slices = [np.ascontiguousarray(pcm_np[i * run * per:(i + 1) * run * per]).astype(np.float32) * np.float32(1 / 32768)
              for i in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda s: len(oracle.encode(s, SR, CH, 16, level, b"")), slices))
    dt = time.perf_counter() - t0
    return dt, seconds, sum(outs)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=int, default=3600, help="audio seconds per GPU per step")
    ap.add_argument("--level", type=int, default=LEVEL)
    ap.add_argument("--cpu-sample-seconds", type=int, default=0, help="0 = auto (about 15 s of CPU work)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-level / PCM16-entry side measurements")
    args = ap.parse_args()
    level = args.level

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    peak, peak_src = peaks()
    seconds = args.seconds
    n_inter = seconds * SR * CH
    workload = (f"1-hour 44.1 kHz 16-bit stereo synthetic multitone+noise, single stream (BASELINE config 2)"
                if world == 1 else
                f"batch corpus shard: {seconds // TRACK_SECONDS} x {TRACK_SECONDS} s 44.1 kHz 16-bit stereo tracks per GPU per step (BASELINE config 3)")
    config = {"workload": workload, "audio_seconds_per_gpu_per_step": seconds, "sample_rate": SR, "channels": CH,
              "level": level, "entry": "Encoder::encode (interleaved f32)", "frames_per_gpu_per_step": seconds,
              "l2": "inputs (1.27 GB f32 per step) are 10x the 126 MB L2; no flush needed",
              "parallelism": f"{n_gpus} x independent shards, no data-path collective"}

    import numpy as np

    # ---------------------------------------------------------------- reference arm (CPU) ----------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import synth_pcm16 as np_synth
        threads = host_threads()
        sample_s = args.cpu_sample_seconds or 60 * threads
        pcm = np_synth(sample_s * SR, CH, SR, seed=SEED, kind="multitone", noise_lsb=64)
        for _ in range(min(args.warmup, 1)):
            cpu_encode_seconds(pcm, threads, threads, level)
        dts = []
        for _ in range(args.steps):
            dt, secs, nbytes = cpu_encode_seconds(pcm, sample_s, threads, level)
            dts.append(dt)
        dt = sum(dts) / len(dts)
        samples = secs * SR * CH
        val = 2.0 * samples / dt / 1e9
        line = {"impl": "reference", "metric": "lossless encode PCM throughput", "value": val, "unit": "GB/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 Levinson",
                "data": "synthetic", "config": config, "x_realtime": secs / dt,
                "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                                 "sample": f"{secs} s of the stream as {threads} independent slices, one thread each; "
                                           "C oracle (literal restatement of the reference encoder, gcc -O2), "
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
    ctx = flo_b200.Context(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # 20 ms period; the summary only uses samples inside the timed region
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # synthetic PCM directly in HBM (integer-only generator, identical to tests/helpers.synth_pcm16)
    from flo_b200 import shard
    if world == 1:
        corpus_n = [seconds * SR]
    else:
        corpus_n = [TRACK_SECONDS * SR] * (world * (seconds // TRACK_SECONDS))      # the whole job's tracks
    ranges = shard.partition_tracks([shard.frames_of_track(n * CH, SR, CH) for n in corpus_n], world)
    t_lo, t_hi = ranges[rank]                                                       # this rank's contiguous shard
    tracks_n = corpus_n[t_lo:t_hi]
    pcm_tracks = [synth_torch.synth_pcm16_long(n, CH, SR, SEED + 131 * (t_lo + i), "multitone", 64, dev)
                  for i, n in enumerate(tracks_n)]
    f32_tracks = [p.to(torch.float32) * (1.0 / 32768.0) for p in pcm_tracks]    # reflo/src/audio.rs:247-254 (exact)
    n_list = [int(t.numel()) for t in f32_tracks]
    total_inter = sum(n_list)
    bound = ctx.output_bound(n_list, [SR] * len(n_list), [CH] * len(n_list))
    d_out = torch.empty(bound, dtype=torch.uint8, device=dev)
    ptrs = [t.data_ptr() for t in f32_tracks]

    def step():
        off, ln = ctx.encode_batch_device(ptrs, n_list, [SR] * len(ptrs), [CH] * len(ptrs), [16] * len(ptrs),
                                          d_out.data_ptr(), bound, level=level)
        if world > 1:       # per-track byte lengths for the final concatenation (the only exchange), non-blocking
            pending.append(shard.exchange_lengths_async([int(v) for v in ln], ranges))
        return off, ln

    pending = []

    def sync_all():
        while pending:      # the concatenation offsets of every step are complete before the clock stops
            pending.pop().result()
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
        step()
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
    value = 2.0 * total_inter * world / (ms_step * 1e-3) / 1e9

    # roofline of the dominant kernel (k_encode_frames): algorithmic bytes = f32 in + .flo out
    k_ms = sum(enc_ms) / len(enc_ms)
    alg_bytes = 4.0 * total_inter + out_bytes
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None                      # DRAM bytes per launch from the committed ncu capture of this exact workload
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and tj["frames"] == seconds and tj["level"] == level:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_encode_frames", "kernel_ms": k_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "issue/ALU-pipe bound at level 5 (10 exhaustive candidates per channel: ~200 thread-instructions per "
                        "channel-sample); DRAM traffic is 1.07 x the algorithmic bytes (some write-back of the L2-resident "
                        "16-bit planes, no re-reads of the input); see DESIGN.md section 6"}

    # side measurements (not the headline): other compression levels and the PCM16 entry, same stream
    extras = None
    if world == 1 and not args.no_extras:
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
                            "pct_of_hbm_peak": 100.0 * (4.0 * total_inter + float(l2[0])) / (k_dec * 1e-3) / 1e9 / peak,
                            "note": "flo_decode_device on the level-5 file of the same stream; latency-bound (one lane per "
                                    "channel of a frame, reader warp + predictor warp), see DESIGN.md section 9.2"}
        del d_dec, q

    # e2e: host buffers through the reference-facing C-ABI call (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        host_in = [torch.empty(n, dtype=torch.float32, pin_memory=True) for n in n_list]
        for h, d in zip(host_in, f32_tracks):
            h.copy_(d)
        torch.cuda.synchronize()
        specs = [flo_b200.TrackSpec(h.numpy(), SR, CH, 16, b"") for h in host_in]
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
        e2e = {"value": 2.0 * total_inter * world / float(tm.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 4 * total_inter, "d2h_bytes_per_step": e2e_out,
               "ms_per_step": float(tm.item()) * 1e3, "x_realtime": seconds * world / float(tm.item())}
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        sample_s = args.cpu_sample_seconds or 60 * threads
        pcm_np = pcm_tracks[0][:sample_s * SR * CH].cpu().numpy()
        dt, secs, _ = cpu_encode_seconds(pcm_np, sample_s, threads, level)
        cpu = {"value": 2.0 * secs * SR * CH / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"first {secs} s of the same stream as {threads} slices, one thread each (C oracle, gcc -O2)",
               "x_realtime": secs / dt}

    if rank == 0:
        line = {"metric": "lossless encode PCM throughput", "value": value, "unit": "GB/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 Levinson", "data": "synthetic",
                "config": config, "x_realtime": seconds * world / (ms_step * 1e-3),
                "pct_of_hbm_peak": 100.0 * achieved / peak, "flo_bytes_per_step_per_gpu": out_bytes,
                "compression_ratio": 2.0 * total_inter / out_bytes, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "device_ms_per_step": sum(dev_ms) / len(dev_ms), "analysis_counters_last_step": counters,
                "extras": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
