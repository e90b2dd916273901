#!/usr/bin/env python
"""bench.py - headline benchmark: Msamples/s of the DTMF Goertzel bank (channels x 8 kHz).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # spandsp's own CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic input.  The input is what SURVEY.md 8(d) specifies,
made on the device by the library's own bit-exact signal sources: channel c (global number) sends 95 digits drawn
from "123A456B789C*0#D" by a linear congruential sequence seeded with c through dtmf_tx() (default level and timing),
plus awgn() at -30 dBm0 seeded 1234567 + c; T = 79 968 samples (784 DTMF blocks, ~10 s of audio).
N = 1 runs BASELINE.json configs[1] (65 536 channels on one B200); N > 1 runs configs[4]'s shape (131 072 channels per
GPU, 1 048 576 at N = 8), channels sharded over ranks with no data-path collective; the per-step NCCL traffic is the
gather of the detected-digit records to rank 0, done by the library (span_b200_bank_gather_*: exact-count
ncclSend/ncclRecv of 12-byte records, overlapped with the next step's kernels).

JSON line (rank 0): see the task contract.  `value` = samples/s with the input resident in HBM; `e2e` = the same
through span_b200_bank_rx_host() with pinned HOST input (H2D inside the timed region) and the records read back
(D2H); `roofline` refers to the filter-bank kernel alone; `parity_check` = 256 random channels of the very buffer
that was timed, run through the reference (oracle/_ref strict build; the plain-C restatement where that is absent)
and compared record for record with what the GPU reported (at N > 1: with what rank 0 gathered).
At N = 1 the line also carries "configs": {"cfg3": ..., "cfg4": ...} - BASELINE.json configs[2] (32 768-channel
super_tone_rx with the Hong Kong descriptor of global-tones.xml) and configs[3] (8 192-channel V.29 receive) - each
with its own value / roofline / cpu_baseline / e2e / parity_check.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_SAMPLES = 79968            # 784 blocks of 102 samples
DIGITS = "123A456B789C*0#D"
PARITY_CHANNELS = 256


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

def digit_strings(chan0, channels, ndigits=95):
    """[channels, ndigits] uint8: channel c's digits from x <- 1103515245*x + 12345 mod 2^31 seeded with c, digit = bits 16..19."""
    x = np.arange(chan0, chan0 + channels, dtype=np.uint64)
    out = np.empty((channels, ndigits), dtype=np.uint8)
    table = np.frombuffer(DIGITS.encode(), dtype=np.uint8)
    for k in range(ndigits):
        x = (x * np.uint64(1103515245) + np.uint64(12345)) & np.uint64(0x7FFFFFFF)
        out[:, k] = table[((x >> np.uint64(16)) & np.uint64(15)).astype(np.int64)]
    return out


def make_dtmf_input(torch, engine, ctx, channels, T, chan0, dev, stream):
    """cfg2 / cfg5 input on the device: dtmf_tx() + awgn() per channel (span_b200_gen.h), int16 [channels][T]."""
    d = torch.empty((channels, T), dtype=torch.int16, device=dev)
    tx = engine.DtmfTxBank(ctx, channels)
    digits = digit_strings(chan0, channels)
    lens = np.full(channels, digits.shape[1], dtype=np.int32)
    rc = engine.lib().span_b200_dtmf_tx_bank_put_each(tx.h, 0, channels, digits.ctypes.data, digits.shape[1], lens.ctypes.data)
    assert rc == 0, rc
    torch.cuda.synchronize()
    tx.tx_device(d.data_ptr(), T, T, True, stream)
    noise = engine.AwgnBank(ctx, channels, -30.0, seed0=1234567 + chan0)
    noise.add_device(d.data_ptr(), T, T, stream)
    noise.sync()
    tx.close()
    noise.close()
    return d


CFG3_CADENCES = [        # nominal cadences of six tones of the Hong Kong set: (f1, l1, f2, l2, d1, d2, d3, d4, repeat), tone id
    ((350, -13, 440, -13, 3000, 0, 0, 0, 1), 0),         # dial tone, continuous
    ((480, -13, 620, -13, 500, 500, 0, 0, 1), 2),        # busy
    ((480, -13, 620, -13, 250, 250, 0, 0, 1), 3),        # congestion
    ((480, -13, 620, -13, 3000, 0, 0, 0, 1), 4),         # number unobtainable, continuous
    ((400, -13, 0, 0, 3000, 0, 0, 0, 1), 6),             # "XXX" of the reference test's fill_descriptor
    ((1100, -13, 0, 0, 500, 3000, 0, 0, 1), 7),          # FAX calling tone
]


def make_tone_input(torch, engine, ctx, channels, T, dev, stream):
    """cfg3 input: channel c plays cadence c mod 6 at -13 dBm0 (tone_gen) + awgn -40 dBm0 seeded 7654321 + c."""
    d = torch.empty((channels, T), dtype=torch.int16, device=dev)
    gen = engine.ToneGenBank(ctx, channels)
    descs = np.asarray([CFG3_CADENCES[c % len(CFG3_CADENCES)][0] for c in range(channels)], dtype=np.int32)
    gen.init_each(descs)
    torch.cuda.synchronize()
    gen.tx_device(d.data_ptr(), T, T, True, stream)
    noise = engine.AwgnBank(ctx, channels, -40.0, seed0=7654321)
    noise.add_device(d.data_ptr(), T, T, stream)
    noise.sync()
    gen.close()
    noise.close()
    return d


def make_v29_input(torch, engine, ctx, channels, T, dev, stream):
    """cfg4 input: v29_tx(9600, no TEP) at -13 dBm0 of the sequence seeded c + 1, + awgn -43 dBm0 seeded 1234567 + c."""
    d = torch.empty((channels, T), dtype=torch.int16, device=dev)
    tx = engine.V29TxBank(ctx, channels, 9600, False)
    tx.power(-13.0)
    tx.set_prbs(seed0=1)
    torch.cuda.synchronize()
    tx.tx_device(d.data_ptr(), T, T, True, stream)
    noise = engine.AwgnBank(ctx, channels, -43.0, seed0=1234567)
    noise.add_device(d.data_ptr(), T, T, stream)
    noise.sync()
    tx.close()
    noise.close()
    return d


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
# CPU reference / baseline (the one place besides the parity check where oracle/ is executed)

def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_oracle(prefer):
    """(oracle, kind): the compiled reference (`prefer` = 'fast' for timing, 'strict' for parity), else the restatement."""
    from oracle import pyoracle as po
    if po.available(prefer):
        return po.load(prefer), "reference"
    po.build(ref=False, port=True)
    return po.load("port"), "port"


def cpu_sample_dtmf(channels, T):
    """A sample of the cfg2 workload made on the host by the reference's own generators (same recipe as the device input)."""
    from oracle import pyoracle as po
    o, _ = load_oracle("fast")
    amp = np.zeros((channels, T), dtype=np.int16)
    digits = digit_strings(0, channels)
    if hasattr(o.lib, "ref_dtmf_tx_calls"):
        for c in range(channels):
            row, _, _ = po.dtmf_tx_calls(o, [T], digits[c].tobytes().decode(), fill=0)
            amp[c] = po.awgn_run(o, T, 1234567 + c, -30.0, into=row.copy())
    else:
        import hashlib  # noqa: F401  (no generator in the restatement: a tone pair per digit is enough to time the detector)
        rng = np.random.default_rng(4242)
        t = np.arange(T)
        for c in range(channels):
            amp[c] = (8000*np.sin(2*np.pi*697*t/8000) + 8000*np.sin(2*np.pi*1209*t/8000) + rng.normal(0, 700, T)).astype(np.int16)
    return amp


def cpu_reference_dtmf(amp, passes, threads):
    """Time the reference's own dtmf_rx() on `threads` host threads, whole-buffer calls (its best case)."""
    from oracle import pyoracle as po
    o, kind = load_oracle("fast")
    p = po.make_params(po.DET_DTMF, po.MODE_DIGITS_CB, amp.shape[1])
    o.run(p, amp[: max(1, threads)], nthreads=threads, want_events=False)       # warm
    secs = 0.0
    for _ in range(passes):
        _, _, s = o.run(p, amp, nthreads=threads, want_events=False)
        secs += s
    return amp.shape[0] * amp.shape[1] * passes / secs / 1e6, kind, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    channels = max(threads * 16, 256)
    channels_total = 65536 if args.gpus == 1 else 131072 * args.gpus
    T = T_SAMPLES
    amp = cpu_sample_dtmf(channels, T)
    v0, kind, s0 = cpu_reference_dtmf(amp, 1, threads)
    passes = max(1, int(1.0 / max(s0, 1e-3)))            # size one step to ~1 s
    for _ in range(args.warmup):
        cpu_reference_dtmf(amp, 1, threads)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        v, kind, s = cpu_reference_dtmf(amp, passes, threads)
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
                   "sample": "%d channels x %d samples x %d passes per step, whole-buffer dtmf_rx() calls" % (channels, T, passes),
                   "input": "dtmf_tx (95 LCG digits per channel) + awgn -30 dBm0, made by the reference's own generators"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind,
                         "sample": "%d ch x %d samples x %d passes x %d steps" % (channels, T, passes, args.steps)},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# parity checks: the GPU's records against the reference on channels sampled from the timed buffer

def rows_by_channel(cols, wanted):
    """cols = (channel, block, kind, a, b, c) arrays -> {channel: [(kind, a, b, c), ...]} for the wanted channels, in order."""
    ch, _, kind, a, b, c = cols
    sel = np.isin(ch, wanted)
    out = {int(w): [] for w in wanted}
    for i in np.nonzero(sel)[0]:
        out[int(ch[i])].append((int(kind[i]), int(a[i]), int(b[i]), int(c[i])))
    return out


def oracle_rows(events, digit_only_a=True):
    out = []
    for e in events:
        kind = int(e["kind"])
        if kind == 1 and digit_only_a:
            out.append((kind, int(e["a"]), 0, 0))
        else:
            out.append((kind, int(e["a"]), int(e["b"]), int(e["c"])))
    return out


def parity_tonebank(det, mode, rows_amp, chan_ids, gpu_cols, tones=None):
    """Run the sampled rows through the reference and compare per channel.  Returns the parity_check object."""
    from oracle import pyoracle as po
    o, kind = load_oracle("strict")
    T = rows_amp.shape[1]
    p = po.make_params(det, mode, T, tones=tones) if tones is not None else po.make_params(det, mode, T)
    ev, _, _ = o.run(p, np.ascontiguousarray(rows_amp), nthreads=host_threads())
    got = rows_by_channel(gpu_cols, chan_ids)
    mism = 0
    nev = 0
    first = None
    for i, c in enumerate(chan_ids):
        want = oracle_rows(ev[i])
        have = got[int(c)]
        nev += len(want)
        if want != have:
            mism += 1
            if first is None:
                k = next((j for j in range(min(len(want), len(have))) if want[j] != have[j]), min(len(want), len(have)))
                first = {"channel": int(c), "index": k, "want": want[k] if k < len(want) else None, "have": have[k] if k < len(have) else None,
                         "n_want": len(want), "n_have": len(have)}
    out = {"channels": int(len(chan_ids)), "events": int(nev), "mismatches": int(mism),
           "oracle": "oracle/_ref strict (the reference's sources, -fno-fast-math)" if kind == "reference" else "oracle/tonebank_oracle.c (restatement)"}
    if first is not None:
        out["first_mismatch"] = first
    return out


# ---------------------------------------------------------------------------------------------

def time_steps(torch, step, steps, barrier):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    r = None
    for _ in range(steps):
        r = step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1), r


def time_steps_host(torch, step, steps, barrier):
    """Steps that involve the host (H2D, D2H, host waits): CUDA events and the wall clock, whichever is longer."""
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    r = None
    for _ in range(steps):
        r = step()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    return max(e0.elapsed_time(e1), wall * 1e3), r


def pinned_like(torch, ctx, d, numa):
    """Pinned host copy of a device tensor; on the GPU's NUMA node when numa (span_b200_host_alloc)."""
    if numa:
        try:
            arr = ctx.host_alloc(tuple(d.shape), np.dtype(str(d.dtype).replace("torch.", "")))
            t = torch.from_numpy(arr)
            t.copy_(d)
            torch.cuda.synchronize()
            return t, arr
        except Exception:  # noqa: BLE001
            pass
    t = torch.empty(tuple(d.shape), dtype=d.dtype, pin_memory=True)
    t.copy_(d)
    torch.cuda.synchronize()
    return t, None


def run_cfg3(torch, engine, ctx, dev, work_stream, args, peaks, peak_src):
    """BASELINE configs[2]: 32 768-channel super_tone_rx with the Hong Kong descriptor (tests/golden/global_tones_hk.json =
    the descriptor tests/super_tone_rx_tests.c builds from spandsp/global-tones.xml), tone + segment reports."""
    from oracle import pyoracle as po
    stream = work_stream.cuda_stream
    C_, T = 32768, 80000
    tones = json.load(open(os.path.join(ROOT, "tests", "golden", "global_tones_hk.json")))
    d_amp = make_tone_input(torch, engine, ctx, C_, T, dev, stream)
    bank = engine.Bank.super_tone(ctx, C_, tones, want_segments=True)
    bank.set_wire(True, 0)
    bank.tune(4, 1)

    def step():
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        return bank.event_count()[0]

    def barrier():
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    bank.kernel_ms()
    ms, nev = time_steps(torch, step, args.steps, barrier)
    kern_ms, kern_n = bank.kernel_ms()
    launches = bank.last_launches * args.steps
    bytes_per_launch = 2.0 * C_ * T
    kms = kern_ms / max(kern_n, 1)
    achieved = bytes_per_launch / (kms / 1e3) / 1e9
    out = {"workload": "32768-channel super_tone_rx, Hong Kong set of global-tones.xml (8 tones, %d monitored frequencies), T=%d" % (bank.bins, T),
           "input": "tone_gen cadences of six of the set's tones at -13 dBm0 (channel c plays c mod 6) + awgn -40 dBm0 seed 7654321+c, made on the device",
           "value": C_ * T * args.steps / (ms / 1e3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms / args.steps, "steps": args.steps,
           "events_per_step": int(nev), "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "traffic": None, "kernel": "bank_kernel_staged<SuperToneDet<%d>>" % ((bank.bins + 1) // 2), "kernel_ms": kms,
                        "launches_timed": kern_n, "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src}}
    # parity: fresh detectors over the timed buffer, 256 random channels against the reference
    bank.reset()
    bank.rx_device(d_amp.data_ptr(), T, T, stream)
    cols = engine.wire_unpack(bank.events_wire())
    rng = np.random.default_rng(3)
    ids = np.sort(rng.choice(C_, PARITY_CHANNELS, replace=False))
    rows = d_amp[torch.from_numpy(ids).to(dev)].cpu().numpy()
    out["parity_check"] = parity_tonebank(po.DET_SUPER_TONE, po.MODE_SEGMENTS, rows, ids, cols, tones=tones)
    # e2e
    if not args.no_e2e:
        h_amp, h_arr = pinned_like(torch, ctx, d_amp, args.numa)
        h_ev = torch.empty((int(nev) * 2 + 1024, 3), dtype=torch.int32, pin_memory=True)
        h_ev_np = h_ev.numpy().view(engine.WIRE_DTYPE).reshape(-1)
        bank.reset()

        def step_host():
            bank.rx_host((h_amp.data_ptr(), T), stream, samples=T)
            return len(bank.events_wire(out=h_ev_np))

        for _ in range(2):
            n2 = step_host()
        e2e_steps = max(2, min(args.steps, 5))
        ems, n2 = time_steps_host(torch, step_host, e2e_steps, barrier)
        out["e2e"] = {"value": C_ * T * e2e_steps / (ems / 1e3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(C_ * T * 2),
                      "d2h_bytes_per_step": int(n2 * 12 + 8), "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
                      "path": "span_b200_bank_rx_host (pinned int16 [channel][sample]) + span_b200_bank_events_wire"}
        del h_amp
        if h_arr is not None:
            ctx.host_free(h_arr)
    # CPU baseline: the reference's super_tone_rx on the host threads, a sample of the same buffer
    if not args.no_cpu:
        threads = host_threads()
        ch = max(threads * 8, 64)
        sample = d_amp[:ch].cpu().numpy()
        o, kind = load_oracle("fast")
        p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, T, tones=tones)
        o.run(p, sample[:threads], nthreads=threads, want_events=False)
        _, _, s0 = o.run(p, sample, nthreads=threads, want_events=False)
        passes = max(1, int(8.0 / max(s0 * threads, 1e-3)))
        secs = 0.0
        for _ in range(passes):
            _, _, s = o.run(p, sample, nthreads=threads, want_events=False)
            secs += s
        out["cpu_baseline"] = {"value": ch * T * passes / secs / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
                               "sample": "%d channels x %d samples x %d passes, whole-buffer super_tone_rx() calls" % (ch, T, passes)}
    bank.close()
    del d_amp
    torch.cuda.empty_cache()
    return out


def run_cfg4(torch, engine, ctx, dev, work_stream, args, peaks, peak_src):
    """BASELINE configs[3]: 8 192-channel V.29 9600 bit/s receive, every channel its own transmitter and noise."""
    from oracle import pyoracle as po
    stream = work_stream.cuda_stream
    C_, T = 8192, 80000
    d_amp = make_v29_input(torch, engine, ctx, C_, T, dev, stream)
    bank = engine.V29Bank(ctx, C_, 9600)
    bank.set_signal_cutoff(-45.5)

    def barrier():
        torch.cuda.synchronize()

    def step():
        # every step is the whole workload from freshly restarted receivers (v29_rx_restart is part of the API)
        bank.restart(9600)
        bank.rx_device(d_amp.data_ptr(), T, T, stream)

    for _ in range(args.warmup):
        step()
    barrier()
    # the kernel alone: CUDA events on the launching stream around the rx call (restart is a separate, earlier launch)
    ks = []
    for _ in range(args.steps):
        bank.restart(9600)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        e1.record()
        torch.cuda.synchronize()
        ks.append(e0.elapsed_time(e1))
    kms = float(np.mean(ks))
    ms, _ = time_steps(torch, step, args.steps, barrier)
    nbits, _ = bank.counts()
    bytes_per_launch = 2.0 * C_ * T
    achieved = bytes_per_launch / (kms / 1e3) / 1e9
    out = {"workload": "8192-channel V.29 9600 bit/s receive (RRC FIR pair + T/2 adaptive equalizer), T=%d" % T,
           "input": "v29_tx of a 23-bit sequence seeded c+1 at -13 dBm0 + awgn -43 dBm0 seed 1234567+c, cutoff -45.5 dBm0, made on the device",
           "value": C_ * T * args.steps / (ms / 1e3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms / args.steps, "steps": args.steps,
           "put_bit_calls_per_step": int(nbits.sum()), "gpu_launches": int(2 * args.steps),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "traffic": None, "kernel": "modem_rx_kernel<RxV29>", "kernel_ms": kms, "launches_timed": len(ks),
                        "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                        "note": "latency- and issue-bound serial receiver, far from the HBM roofline by construction (DESIGN.md 7)"}}
    # parity: 256 random channels of the timed buffer through the reference's v29_rx; put_bit streams must be identical
    rng = np.random.default_rng(4)
    ids = np.sort(rng.choice(C_, PARITY_CHANNELS, replace=False))
    rows = np.ascontiguousarray(d_amp[torch.from_numpy(ids).to(dev)].cpu().numpy())
    o, okind = load_oracle("strict")
    pc = {"channels": int(len(ids)), "oracle": "oracle/_ref strict (the reference's v29rx.c)" if okind == "reference" else "unavailable"}
    if okind == "reference":
        cap = 2 * T + 64
        rbits = np.zeros((len(ids), cap), dtype=np.int8)
        rn = np.zeros(len(ids), dtype=np.int32)
        o.lib.ref_v29_run_batch.restype = C.c_double
        o.lib.ref_v29_run_batch(C.c_void_p(rows.ctypes.data), C.c_int64(T), C.c_int(len(ids)), C.c_int(T), C.c_int(T), C.c_int(9600),
                                C.c_float(-45.5), C.c_int(host_threads()), C.c_void_p(rbits.ctypes.data), C.c_int64(cap), C.c_void_p(rn.ctypes.data))
        mism = 0
        trained = 0
        total = 0
        for i, c in enumerate(ids):
            have = bank.bits(int(c))
            want = rbits[i, : rn[i]]
            total += int(rn[i])
            trained += int((want == -4).any())
            if len(have) != len(want) or not (have == want).all():
                mism += 1
        pc.update({"put_bit_calls": total, "mismatches": mism, "channels_trained": trained,
                   "note": "at -43 dBm0 noise with the -45.5 dBm0 cutoff the reference's carrier detector fires on the noise alone in part of "
                           "the channels, which then report TRAINING_FAILED and park; the GPU reproduces each of them"})
    out["parity_check"] = pc
    if not args.no_e2e:
        h_amp, h_arr = pinned_like(torch, ctx, d_amp, args.numa)
        wcap = (int(nbits.max()) + 31) // 32 + 2
        scap = 64
        h_words = torch.empty((C_, wcap), dtype=torch.int32, pin_memory=True)       # the packed put_bit streams of every channel
        h_status = torch.empty((C_, scap, 2), dtype=torch.int32, pin_memory=True)
        h_n = np.zeros(C_, dtype=np.int32)
        h_ns = np.zeros(C_, dtype=np.int32)
        fn = engine.lib().span_b200_v29_bank_output_packed

        def step_host():
            bank.restart(9600)
            bank.rx_host((h_amp.data_ptr(), T), stream, samples=T)
            return fn(bank.h, h_words.data_ptr(), wcap, h_n.ctypes.data, h_status.data_ptr(), scap, h_ns.ctypes.data)

        for _ in range(2):
            step_host()
        e2e_steps = max(2, min(args.steps, 5))
        ems, mx = time_steps_host(torch, step_host, e2e_steps, barrier)
        out["e2e"] = {"value": C_ * T * e2e_steps / (ems / 1e3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(C_ * T * 2),
                      "d2h_bytes_per_step": int(C_ * (4 * ((int(mx) + 31) // 32) + 8 * int(h_ns.max()) + 8)), "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
                      "path": "span_b200_v29_bank_rx_host (pinned int16, fed in four pieces of time: the copy of a piece overlaps the kernel of the one before) + span_b200_v29_bank_output_packed (every channel's put_bit stream: "
                              "data bits 32 to a word + status reports)"}
        del h_amp
        if h_arr is not None:
            ctx.host_free(h_arr)
    if not args.no_cpu:
        threads = host_threads()
        ch = max(threads * 2, 16)
        sample = np.ascontiguousarray(d_amp[:ch].cpu().numpy())
        o, kind = load_oracle("fast")
        if kind == "reference":
            o.lib.ref_v29_run_batch.restype = C.c_double

            def run_once():
                return o.lib.ref_v29_run_batch(C.c_void_p(sample.ctypes.data), C.c_int64(T), C.c_int(ch), C.c_int(T), C.c_int(T), C.c_int(9600),
                                               C.c_float(-45.5), C.c_int(threads), None, C.c_int64(0), None)
            s0 = run_once()
            passes = max(1, int(8.0 / max(s0 * threads, 1e-3)))
            secs = sum(run_once() for _ in range(passes))
            out["cpu_baseline"] = {"value": ch * T * passes / secs / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
                                   "sample": "%d channels x %d samples x %d passes, whole-buffer v29_rx() calls" % (ch, T, passes)}
        else:
            out["cpu_baseline"] = None
    bank.close()
    del d_amp
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 sub-benchmarks (N = 1 only)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--numa", type=int, default=1, help="1: staging memory of the e2e arm on the GPU's NUMA node (span_b200_host_alloc)")
    ap.add_argument("--nccl-ctas", type=int, default=4, help="cap on the thread blocks NCCL may use for the record gather")
    ap.add_argument("--gather", default="peer_copy", choices=["peer_copy", "nccl"],
                    help="how the records travel to rank 0: the root's copy engines over NVLink (CUDA IPC), or ncclSend/ncclRecv")
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

    C_ = args.channels_per_gpu or (65536 if args.gpus == 1 else 131072)
    T = args.samples
    chan0 = rank * C_
    ctx = engine.Context(local)
    bank = engine.Bank.dtmf(ctx, C_)
    if args.realtime:
        bank.dtmf_realtime(True)
    if args.variant >= 0:
        bank.tune(1, args.variant)
    if args.slice >= 0:
        bank.tune(0, args.slice)
    if args.packed >= 0:
        bank.tune(3, args.packed)
    bank.tune(4, 1)
    bank.set_wire(True, chan0)

    # One explicit stream for the library's launches AND the timing events (a NULL stream would make the
    # library use its own context stream, which torch's events do not see).
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    d_amp = make_dtmf_input(torch, engine, ctx, C_, T, chan0, dev, stream)
    torch.cuda.synchronize()

    comm = None
    if world > 1:
        # the library's own communicator: the id travels over torch.distributed, the records over the library's NCCL calls
        uid = torch.from_numpy(engine.Comm.unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        comm = engine.Comm(ctx, uid.cpu().numpy(), world, rank, max_ctas=args.nccl_ctas)
        comm.set_transport(args.gather)
        bank.attach_comm(comm, 0)

    state = {"k": 0, "total": 0}

    def step_device():
        """rx(k); the records of call k-1 travel (exact-count send/recv) while the kernels of call k run."""
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        if world > 1:
            if state["k"] > 0:
                state["total"], _ = bank.gather_end()
            bank.gather_begin()
            state["k"] += 1
            return state["total"]
        return bank.event_count()[0]

    def drain():
        if world > 1 and state["k"] > 0:
            state["total"], _ = bank.gather_end()
            state["k"] = 0
            bank.gathered()                                 # waits for the transfer

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
    drain()
    bank.kernel_ms()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    t0 = time.perf_counter()
    total_events = 0
    for _ in range(args.steps):
        total_events = step_device()
    if world > 1:
        drain()                                             # the last gather is inside the timed region
        total_events = state["total"]
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        ms = max(ms, wall_ms)                               # the gather's completion is waited for on the host
    kern_ms, kern_n = bank.kernel_ms()
    launches = bank.last_launches * args.steps
    kernel_path = bank.last_path
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = C_ * world * T * args.steps / (ms / 1e3) / 1e6

    # ---- parity check: fresh detectors over the timed buffer; 256 random channels against the reference -------------
    parity = None
    if not args.no_parity:
        from oracle import pyoracle as po
        bank.reset()
        if args.realtime:
            bank.dtmf_realtime(True)
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        rng = np.random.default_rng(2)
        ids = np.sort(rng.choice(C_ * world, PARITY_CHANNELS, replace=False))          # global channel numbers, same on every rank
        mine = ids[(ids >= chan0) & (ids < chan0 + C_)]
        rows_local = d_amp[torch.from_numpy(mine - chan0).to(dev)] if len(mine) else torch.empty((0, T), dtype=torch.int16, device=dev)
        if world > 1:
            bank.gather_begin()
            bank.gather_end()
            sizes = [int(((ids >= r * C_) & (ids < (r + 1) * C_)).sum()) for r in range(world)]
            pad = torch.zeros((max(sizes), T), dtype=torch.int16, device=dev)
            pad[: rows_local.shape[0]] = rows_local
            pad8 = pad.view(torch.uint8)                    # (torch's NCCL binding has no int16)
            bufs = [torch.empty_like(pad8) for _ in range(world)] if rank == 0 else None
            dist.gather(pad8, bufs, dst=0)
            if rank == 0:
                rows = torch.cat([b.view(torch.int16)[:n] for b, n in zip(bufs, sizes)]).cpu().numpy()
                cols = engine.wire_unpack(bank.gathered_host())
        else:
            rows = rows_local.cpu().numpy()
            cols = engine.wire_unpack(bank.events_wire())
        if rank == 0:
            parity = parity_tonebank(po.DET_DTMF, po.MODE_REALTIME if args.realtime else po.MODE_DIGITS_CB, rows, ids, cols)
            parity["records_checked_from"] = "rank 0's gathered buffer (all ranks' records)" if world > 1 else "span_b200_bank_events_wire"
        barrier()

    # ---- end-to-end arm: pinned host input through span_b200_bank_rx_host ----------------------
    def all_ok(ok):
        """Every rank managed (a failed allocation on one rank must not leave the others waiting in a collective)."""
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    e2e = None
    if not args.no_e2e:
        if comm is not None:
            comm.sync()
        nev_cap = C_ * (T // 102 // 2 + 1)
        h_amp = h_arr = h_events = None
        err = None
        try:
            h_amp, h_arr = pinned_like(torch, ctx, d_amp, args.numa)
            h_events_t = torch.empty((nev_cap, 3), dtype=torch.int32, pin_memory=True)      # pinned: D2H at link speed
            h_events = h_events_t.numpy().view(engine.WIRE_DTYPE).reshape(-1)
        except Exception as e:  # noqa: BLE001
            err = "%s: %s" % (type(e).__name__, str(e).splitlines()[0])
        if not all_ok(err is None):
            e2e = {"error": "pinned host staging could not be allocated on every rank (%s)" % (err or "another rank")}
        else:
            bank.reset()
            if args.realtime:
                bank.dtmf_realtime(True)
            e2e_steps = max(2, min(args.steps, 5))

            def step_host():
                bank.rx_host((h_amp.data_ptr(), T), stream, samples=T)
                return len(bank.events_wire(out=h_events))

            for _ in range(2):
                nev = step_host()
            ems, nev = time_steps_host(torch, step_host, e2e_steps, barrier)
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
            e2e = {"value": C_ * world * T * e2e_steps / (ems / 1e3) / 1e6, "unit": "Msamples/s",
                   "h2d_bytes_per_step": int(C_ * T * 2), "d2h_bytes_per_step": int(nev * 12 + 8),
                   "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
                   "path": "span_b200_bank_rx_host (pinned int16 [channel][sample], copy pipelined with the kernels over channel ranges) "
                           "+ span_b200_bank_events_wire (12-byte records)",
                   "numa": {"gpu_node": ctx.numa_node, "staging": "span_b200_host_alloc (GPU's node)" if h_arr is not None else "cudaHostAlloc (default policy)"}}
            # ---- same, with 8-bit u-law input (G.711 expand fused into the kernel load).
            # Reported beside e2e, not instead of it: the reference API takes int16.  The u-law bytes go into the first half
            # of the staging buffer the int16 arm used (no second pinned allocation).
            gpath = os.path.join(ROOT, "tests", "golden", "g711_golden.npz")
            if os.path.exists(gpath) and not args.no_g711:
                enc = torch.from_numpy(np.load(gpath)["encode_ulaw"]).to(dev)
                h_u8 = h_amp.view(torch.uint8).reshape(-1)[: C_ * T].view(C_, T)
                for c0 in range(0, C_, 2048):
                    c1 = min(C_, c0 + 2048)
                    h_u8[c0:c1].copy_(enc[(d_amp[c0:c1].to(torch.int32) + 32768).long()])
                torch.cuda.synchronize()
                del enc
                bank.reset()
                if args.realtime:
                    bank.dtmf_realtime(True)

                def step_host_g711():
                    bank.rx_host_g711((h_u8.data_ptr(), T), False, stream, samples=T)
                    return len(bank.events_wire(out=h_events))

                for _ in range(2):
                    nev8 = step_host_g711()
                gms, nev8 = time_steps_host(torch, step_host_g711, e2e_steps, barrier)
                t = torch.tensor([gms], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                gms = float(t.item())
                e2e["g711_ulaw"] = {"value": C_ * world * T * e2e_steps / (gms / 1e3) / 1e6, "unit": "Msamples/s",
                                    "h2d_bytes_per_step": int(C_ * T), "d2h_bytes_per_step": int(nev8 * 12 + 8),
                                    "ms_per_step": gms / e2e_steps, "path": "span_b200_bank_rx_host_g711 (pinned u-law bytes)"}
                del h_u8
        del h_amp
        if h_arr is not None:
            ctx.host_free(h_arr)

    if rank != 0:
        if comm is not None:
            comm.sync()
        barrier()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_src = measured_peaks()
    bytes_per_launch = 2.0 * C_ * T                      # 2 B per input sample, nothing else
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
        sample = d_amp[:ch].cpu().numpy()                # the CPU baseline runs on rows of the very buffer the GPU was timed on
        v0, kind, s0 = cpu_reference_dtmf(sample, 1, threads)
        passes = max(1, int(20.0 / max(s0 * threads, 1e-3)))         # ~20 core-seconds of CPU work
        v, kind, s = cpu_reference_dtmf(sample, passes, threads)
        cpu = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": kind,
               "sample": "%d channels x %d samples x %d passes, whole-buffer dtmf_rx() calls, %.1f core-seconds"
                         % (ch, T, passes, s * threads)}

    configs = None
    if world == 1 and not args.no_configs:
        bank.close()
        del d_amp
        torch.cuda.empty_cache()
        configs = {}
        for name, fn in (("cfg3", run_cfg3), ("cfg4", run_cfg4)):
            try:
                configs[name] = fn(torch, engine, ctx, dev, work_stream, args, peaks, peak_src)
            except Exception as e:  # noqa: BLE001
                configs[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {
        "metric": "Msamples/s DTMF Goertzel (chans x 8kHz)", "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d-channel DTMF Goertzel (8 bins, 102-sample blocks), T=%d samples/channel/step"
                               % (C_ * world, T),
                   "input": "dtmf_tx (95 LCG digits per channel, seed = channel) + awgn -30 dBm0 seed 1234567+channel, made on the device "
                            "by the library's bit-exact dtmf_tx / awgn banks (SURVEY 8d cfg2)",
                   "channels_per_gpu": C_, "events": "realtime (on/off, level, duration)" if args.realtime else "digits",
                   "events_per_step": int(total_events), "l2": "input %.1f GB per GPU per step, larger than L2" % (C_ * T * 2 / 1e9),
                   "kernel_path": kernel_path,
                   "parallelism": "channels sharded x%d; records gathered to rank 0 by span_b200_bank_gather_* (12-byte records, "
                                  "counts by ncclAllGather, records by %s, NCCL capped at %d CTAs)"
                                  % (world, "the root's copy engines over NVLink (exact counts, CUDA IPC)" if args.gather == "peer_copy"
                                     else "exact-count ncclSend/ncclRecv", args.nccl_ctas)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity_check": parity,
    }
    if configs is not None:
        line["configs"] = configs
    print(json.dumps(line))
    if comm is not None:
        comm.sync()
    barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
