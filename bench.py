#!/usr/bin/env python
"""bench.py - headline benchmark: Msamples/s of the DTMF Goertzel bank (channels x 8 kHz).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # spandsp's own CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic input: every channel receives
T = 79 968 samples (784 DTMF blocks, ~10 s of audio) of synthetic dtmf_tx-like digits in AWGN.
N = 1 runs BASELINE.json configs[1] (65 536 channels on one B200); N > 1 runs configs[4]'s shape
(131 072 channels per GPU, 1 048 576 at N = 8), channels sharded over ranks with no data-path
collective; the per-step NCCL traffic is the gather of the detected-digit event records to rank 0.

JSON line (rank 0): see the task contract.  `value` = samples/s with the input resident in HBM;
`e2e` = the same through span_b200_bank_rx_host() with pinned HOST input (H2D inside the timed
region) and the event records read back (D2H).  `roofline` refers to the filter-bank kernel alone.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

T_SAMPLES = 79968            # 784 blocks of 102 samples
DIGIT_SAMPLES = 840          # 50 ms on + 55 ms off (src/dtmf.c:68-69)
ON_SAMPLES = 400
ROW = [697.0, 770.0, 852.0, 941.0]
COL = [1209.0, 1336.0, 1477.0, 1633.0]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# synthetic input

def synth_dtmf_torch(torch, channels, T, seed, device, out=None, chunk=2048):
    """[channels, T] int16 on `device`: random DTMF digits (-10 dBm0 per tone, 50/55 ms cadence)
    plus -30 dBm0 white Gaussian noise.  Same shape as dtmf_tx + awgn."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if out is None:
        out = torch.empty((channels, T), dtype=torch.int16, device=device)
    amp = 32768.0 * 10.0 ** ((-10.0 - 3.14) / 20.0)
    noise_rms = 32768.0 * 10.0 ** ((-30.0 - 3.14 - 3.02) / 20.0)
    t = torch.arange(T, device=device, dtype=torch.float32)
    k = torch.arange(T, device=device) // DIGIT_SAMPLES
    on = ((torch.arange(T, device=device) % DIGIT_SAMPLES) < ON_SAMPLES).to(torch.float32)
    ndig = int(k.max().item()) + 1
    rowf = torch.tensor(ROW, device=device)
    colf = torch.tensor(COL, device=device)
    w = 2.0 * np.pi / 8000.0
    for c0 in range(0, channels, chunk):
        c1 = min(channels, c0 + chunk)
        d = torch.randint(0, 16, (c1 - c0, ndig), generator=g, device=device)
        fr = rowf[d >> 2][:, k]
        fc = colf[d & 3][:, k]
        x = amp * (torch.sin(w * fr * t) + torch.sin(w * fc * t)) * on
        x += noise_rms * torch.randn((c1 - c0, T), generator=g, device=device)
        out[c0:c1] = x.round_().clamp_(-32768, 32767).to(torch.int16)
        del d, fr, fc, x
    return out


def synth_dtmf_numpy(channels, T, seed):
    rng = np.random.default_rng(seed)
    amp = 32768.0 * 10.0 ** ((-10.0 - 3.14) / 20.0)
    noise_rms = 32768.0 * 10.0 ** ((-30.0 - 3.14 - 3.02) / 20.0)
    t = np.arange(T, dtype=np.float64)
    k = np.arange(T) // DIGIT_SAMPLES
    on = ((np.arange(T) % DIGIT_SAMPLES) < ON_SAMPLES).astype(np.float64)
    ndig = int(k.max()) + 1
    out = np.empty((channels, T), dtype=np.int16)
    w = 2.0 * np.pi / 8000.0
    for c in range(channels):
        d = rng.integers(0, 16, ndig)
        fr = np.asarray(ROW)[d >> 2][k]
        fc = np.asarray(COL)[d & 3][k]
        x = amp * (np.sin(w * fr * t) + np.sin(w * fc * t)) * on + rng.normal(0.0, noise_rms, T)
        out[c] = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    return out


# ---------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region (NVML, 2 ms
    period in a thread; the timed region is only tens of milliseconds long)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.err = None
        self.t0 = None
        self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except Exception:
                return self.idx
        return self.idx

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, mx, pw, rs, time.perf_counter()))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=5)
        allsamples = self.samples
        if self.t0 is not None and self.t1 is not None:
            inside = [x for x in allsamples if self.t0 <= x[4] <= self.t1]
            if inside:
                self.samples = inside
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: %s" % (self.err or "nvml")]}
        sm = [x[0] for x in self.samples]
        reasons = set()
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        for x in self.samples:
            for bit, name in bits.items():
                if x[3] & bit:
                    reasons.add(name)
        pw = [x[2] for x in self.samples if x[2] is not None]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(pw)) if pw else None}


# ---------------------------------------------------------------------------------------------
# CPU reference / baseline

def cpu_reference(channels, T, passes, threads, seed=4242):
    """Time the reference's own dtmf_rx() (oracle/_ref fast build = spandsp's default flags) - or the
    plain-C port where the reference could not be compiled - on `threads` host threads.
    Whole-buffer calls (the reference's best case).  Returns (Msamples/s, kind, seconds)."""
    from oracle import pyoracle as po
    if po.available("fast"):
        o, kind = po.load("fast"), "reference"
    else:
        po.build(ref=False, port=True)
        o, kind = po.load("port"), "port"
    amp = synth_dtmf_numpy(channels, T, seed)
    p = po.make_params(po.DET_DTMF, po.MODE_DIGITS_CB, T)
    o.run(p, amp[: max(1, threads)], nthreads=threads, want_events=False)       # warm
    secs = 0.0
    for _ in range(passes):
        _, _, s = o.run(p, amp, nthreads=threads, want_events=False)
        secs += s
    return channels * T * passes / secs / 1e6, kind, secs


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------

def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    channels = max(threads * 16, 256)
    channels_total = 65536 if args.gpus == 1 else 131072 * args.gpus
    T = T_SAMPLES
    passes = 1
    # size one step to ~2 s
    v0, kind, s0 = cpu_reference(channels, T, 1, threads)
    passes = max(1, int(1.0 / max(s0, 1e-3)))
    for _ in range(args.warmup):
        cpu_reference(channels, T, 1, threads)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        v, kind, s = cpu_reference(channels, T, passes, threads)
        secs += s
    wall = time.perf_counter() - t0
    value = channels * T * passes * args.steps / secs / 1e6
    line = {
        "impl": "reference",
        "metric": "Msamples/s DTMF Goertzel (chans x 8kHz)", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d-channel DTMF Goertzel (8 bins, 102-sample blocks), T=%d" % (channels_total, T),
                   "sample": "%d channels x %d samples x %d passes per step, whole-buffer dtmf_rx() calls" % (channels, T, passes)},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind,
                         "sample": "%d ch x %d samples x %d passes x %d steps" % (channels, T, passes, args.steps)},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels-per-gpu", type=int, default=0)
    ap.add_argument("--samples", type=int, default=T_SAMPLES)
    ap.add_argument("--variant", type=int, default=-1, help="staging variant knob (tuning)")
    ap.add_argument("--slice", type=int, default=-1, help="blocks per time slice (tuning)")
    ap.add_argument("--packed", type=int, default=-1, help="f32x2 adds on/off (tuning)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-g711", action="store_true")
    ap.add_argument("--realtime", type=int, default=1, help="1: realtime (on/off + level) events, 0: digit events")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from spandsp_b200 import build as sb_build
    sb_build.build()
    from spandsp_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    C = args.channels_per_gpu or (65536 if args.gpus == 1 else 131072)
    T = args.samples
    ctx = engine.Context(local)
    bank = engine.Bank.dtmf(ctx, C)
    if args.realtime:
        bank.dtmf_realtime(True)
    if args.variant >= 0:
        bank.tune(1, args.variant)
    if args.slice >= 0:
        bank.tune(0, args.slice)
    if args.packed >= 0:
        bank.tune(3, args.packed)
    bank.tune(4, 1)

    d_amp = synth_dtmf_torch(torch, C, T, 1234567 + rank, dev)
    torch.cuda.synchronize()
    # One explicit stream for the library's launches AND the timing events (a NULL stream would make the
    # library use its own context stream, which torch's events do not see).
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream

    ev_cap = C * (T // 102 // 2 + 1)
    d_events = None
    comm_stream = None
    if world > 1:
        # Two record buffers and a side stream for NCCL: the gather of step k-1 travels over NVLink while the
        # kernels of step k execute.  ready[i]: buffer i holds a step's records (recorded on the work stream
        # BEFORE the next step is launched, so the gather does not wait for that next step); done[i]: the
        # gather that read buffer i has finished (the work stream waits for it before overwriting the buffer).
        d_events = [torch.empty((ev_cap, 6), dtype=torch.int32, device=dev) for _ in range(2)]
        comm_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [None, None]
    pending = {"n": None, "buf": None, "slot": 0, "total": 0}

    def gather_events(buf, n_local, slot):
        """NCCL: event counts all-gathered, records gathered to rank 0 (padded to the max count)."""
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready[slot])
            cnt = torch.tensor([n_local], dtype=torch.int64, device=dev)
            allc = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(allc, cnt)
            counts = [int(c.item()) for c in allc]
            mx = max(max(counts), 1)
            send = buf[:mx]
            if rank == 0:
                bufs = [torch.empty_like(send) for _ in range(world)]
                dist.gather(send, bufs, dst=0)
            else:
                dist.gather(send, None, dst=0)
            ev = torch.cuda.Event()
            ev.record(comm_stream)
            done[slot] = ev
        return sum(counts)

    def flush_gather():
        if pending["n"] is not None:
            pending["total"] = gather_events(pending["buf"], pending["n"], pending["slot"])
            pending["n"] = None

    step_no = [0]

    def step_device():
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        if world > 1:
            flush_gather()                                  # step k-1's records travel while step k computes
            slot = step_no[0] & 1
            buf = d_events[slot]
            step_no[0] += 1
            if done[slot] is not None:
                work_stream.wait_event(done[slot])          # the gather of step k-2 has released this buffer
            pending["n"] = bank.events_to_device(buf.data_ptr(), ev_cap, stream)
            ready[slot].record(work_stream)
            pending["buf"] = buf
            pending["slot"] = slot
            return pending["total"]
        n, ov = bank.event_count()
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # NVML start-up takes longer than the timed region; start early
    for _ in range(args.warmup):
        step_device()
    if world > 1:
        flush_gather()
    bank.kernel_ms()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    total_events = 0
    for _ in range(args.steps):
        total_events = step_device()
    if world > 1:
        flush_gather()
        total_events = pending["total"]
        work_stream.wait_stream(comm_stream)                # the last gather is inside the timed region
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = bank.kernel_ms()
    launches = bank.last_launches * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = C * world * T * args.steps / (ms / 1e3) / 1e6

    # ---- end-to-end arm: pinned host input through span_b200_bank_rx_host ----------------------
    e2e = None
    if not args.no_e2e:
        h_amp = torch.empty((C, T), dtype=torch.int16, pin_memory=True)
        h_amp.copy_(d_amp)
        torch.cuda.synchronize()
        h_events_t = torch.empty((ev_cap, 6), dtype=torch.int32, pin_memory=True)      # pinned: D2H at link speed
        h_events = h_events_t.numpy().view(engine.EVENT_DTYPE).reshape(-1)
        bank.reset()
        if args.realtime:
            bank.dtmf_realtime(True)
        e2e_steps = max(2, min(args.steps, 5))

        def step_host():
            bank.rx_host((h_amp.data_ptr(), T), stream, samples=T)
            ev = bank.events(out=h_events)
            return len(ev)

        for _ in range(2):
            nev = step_host()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(e2e_steps):
            nev = step_host()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max(e0.elapsed_time(e1), wall * 1e3)
        t = torch.tensor([ems], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        e2e = {"value": C * world * T * e2e_steps / (ems / 1e3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(C * T * 2), "d2h_bytes_per_step": int(nev * 24 + 8),
               "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
               "path": "span_b200_bank_rx_host (pinned int16 [channel][sample]) + span_b200_bank_events"}
        del h_amp
        # ---- same, with 8-bit u-law input (G.711 expand fused into the kernel load).
        # Reported beside e2e, not instead of it: the reference API takes int16.
        gpath = os.path.join(ROOT, "tests", "golden", "g711_golden.npz")
        if os.path.exists(gpath) and not args.no_g711:
            enc = torch.from_numpy(np.load(gpath)["encode_ulaw"]).to(dev)
            h_u8 = torch.empty((C, T), dtype=torch.uint8, pin_memory=True)
            for c0 in range(0, C, 2048):
                c1 = min(C, c0 + 2048)
                h_u8[c0:c1].copy_(enc[(d_amp[c0:c1].to(torch.int32) + 32768).long()])
            torch.cuda.synchronize()
            del enc
            bank.reset()
            if args.realtime:
                bank.dtmf_realtime(True)

            def step_host_g711():
                bank.rx_host_g711((h_u8.data_ptr(), T), False, stream, samples=T)
                return len(bank.events(out=h_events))

            for _ in range(2):
                nev8 = step_host_g711()
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(e2e_steps):
                nev8 = step_host_g711()
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            gms = max(e0.elapsed_time(e1), wall * 1e3)
            t = torch.tensor([gms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gms = float(t.item())
            e2e["g711_ulaw"] = {"value": C * world * T * e2e_steps / (gms / 1e3) / 1e6, "unit": "Msamples/s",
                                "h2d_bytes_per_step": int(C * T), "d2h_bytes_per_step": int(nev8 * 24 + 8),
                                "ms_per_step": gms / e2e_steps, "path": "span_b200_bank_rx_host_g711 (pinned u-law bytes)"}
            del h_u8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = measured_peaks()
    bytes_per_launch = 2.0 * C * T                       # 2 B per input sample, nothing else
    kern_s = (kern_ms / max(kern_n, 1)) / 1e3
    achieved = bytes_per_launch / kern_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                "kernel": "bank_kernel_staged<DtmfDet>", "kernel_ms": kern_ms / max(kern_n, 1), "launches_timed": kern_n,
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src + " (of measured)"}
    prof = os.path.join(ROOT, "profiles", "dtmf_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu:
        threads = host_threads()
        ch = max(threads * 16, 256)
        v0, kind, s0 = cpu_reference(ch, T, 1, threads)
        passes = max(1, int(20.0 / max(s0 * threads, 1e-3)))         # ~20 core-seconds of CPU work
        v, kind, s = cpu_reference(ch, T, passes, threads)
        cpu = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": kind,
               "sample": "%d channels x %d samples x %d passes, whole-buffer dtmf_rx() calls, %.1f core-seconds"
                         % (ch, T, passes, s * threads)}

    line = {
        "metric": "Msamples/s DTMF Goertzel (chans x 8kHz)", "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d-channel DTMF Goertzel (8 bins, 102-sample blocks), T=%d samples/channel/step"
                               % (C * world, T),
                   "channels_per_gpu": C, "events": "realtime (on/off, level, duration)" if args.realtime else "digits",
                   "events_per_step": int(total_events), "l2": "input %.1f GB per GPU per step, larger than L2" % (C * T * 2 / 1e9),
                   "kernel_path": bank.last_path, "parallelism": "channels sharded x%d, NCCL event gather" % world},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
