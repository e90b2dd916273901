"""ctypes access to the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Two interchangeable back ends expose the same flat C interface (oracle/ref_harness.h):

* ``ref``  - oracle/_ref/libspandsp_ref_{strict,fast}.so : the reference's OWN sources
             (compiled in place from /root/reference/src by oracle/Makefile) behind
             oracle/ref_harness.c.  ``strict`` is the pinned oracle, ``fast`` is the
             reference's default flag set and serves as the CPU baseline.
* ``port`` - oracle/libtonebank_oracle.so : our plain-C restatement of the algorithm
             (oracle/tonebank_oracle.c), buildable anywhere.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing in spandsp_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

DET_DTMF, DET_BELL_MF, DET_R2_MF, DET_SUPER_TONE = 0, 1, 2, 3
MODE_DIGITS_CB, MODE_REALTIME, MODE_POLL, MODE_SEGMENTS = 0, 1, 2, 3
EV_DIGIT, EV_TONE, EV_SEGMENT = 1, 2, 5

MAX_ST_TONES = 32
MAX_ST_ELEMENTS = 128


class RefEvent(C.Structure):
    _fields_ = [("chunk", C.c_int32), ("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32)]


EVENT_DTYPE = np.dtype([("chunk", "<i4"), ("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")])


class RefParams(C.Structure):
    _fields_ = [
        ("detector", C.c_int32),
        ("mode", C.c_int32),
        ("chunk", C.c_int32),
        ("fillin_every", C.c_int32),
        ("dtmf_set_parms", C.c_int32),
        ("dtmf_filter_dialtone", C.c_int32),
        ("dtmf_twist", C.c_float),
        ("dtmf_reverse_twist", C.c_float),
        ("dtmf_threshold", C.c_float),
        ("r2_fwd", C.c_int32),
        ("st_tones", C.c_int32),
        ("st_tone_segs", C.c_int32 * MAX_ST_TONES),
        ("st_elements", C.c_int32 * (4 * MAX_ST_ELEMENTS)),
    ]


class RefFinal(C.Structure):
    _fields_ = [("status", C.c_int32), ("ndigits", C.c_int32), ("digits", C.c_char * 256)]


FINAL_DTYPE = np.dtype([("status", "<i4"), ("ndigits", "<i4"), ("digits", "S256")])


def make_params(detector, mode=MODE_DIGITS_CB, chunk=160, fillin_every=0, dtmf_parms=None, r2_fwd=1, tones=None):
    """tones: list of tones, each a list of (f1, f2, min_ms, max_ms) elements (super-tone)."""
    p = RefParams()
    p.detector = detector
    p.mode = mode
    p.chunk = chunk
    p.fillin_every = fillin_every
    if dtmf_parms is not None:
        p.dtmf_set_parms = 1
        p.dtmf_filter_dialtone = int(dtmf_parms.get("filter_dialtone", -1))
        p.dtmf_twist = float(dtmf_parms.get("twist", -1.0))
        p.dtmf_reverse_twist = float(dtmf_parms.get("reverse_twist", -1.0))
        p.dtmf_threshold = float(dtmf_parms.get("threshold", -99.0))
    p.r2_fwd = int(r2_fwd)
    if tones:
        assert len(tones) <= MAX_ST_TONES
        k = 0
        p.st_tones = len(tones)
        for t, elements in enumerate(tones):
            p.st_tone_segs[t] = len(elements)
            for e in elements:
                assert k < MAX_ST_ELEMENTS
                for i in range(4):
                    p.st_elements[4 * k + i] = int(e[i])
                k += 1
    return p


def build(ref=True, port=True, quiet=True):
    """Compile the oracles (make -C oracle).  The _ref part is skipped where /root/reference is absent."""
    targets = []
    if port:
        targets.append("port")
    if ref and os.path.exists("/root/reference/src/dtmf.c"):
        targets.append("ref")
    if targets:
        subprocess.run(["make", "-C", HERE] + targets, check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


class Oracle:
    """One loaded oracle library."""

    def __init__(self, path, name):
        self.name = name
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        L = self.lib
        L.ref_run.restype = C.c_double
        L.ref_run.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ref_abi_version.restype = C.c_int
        for fn in ("ref_dtmf_generate", "ref_bell_mf_generate", "ref_r2_mf_generate", "ref_cadence_generate"):
            if hasattr(L, fn):
                getattr(L, fn).restype = C.c_int
        if hasattr(L, "ref_goertzel_fac"):
            L.ref_goertzel_fac.restype = C.c_float
            L.ref_goertzel_fac.argtypes = [C.c_float, C.c_int]
            L.ref_goertzel_blocks.restype = C.c_int
            L.ref_goertzel_blocks.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    # ---- detectors -------------------------------------------------------------------
    def run(self, params, amp, nthreads=1, ev_cap=None, want_events=True):
        """amp: int16 [channels, n] (C-contiguous rows, row stride may exceed n via slicing of a 2-D array).

        Returns (events list per channel as structured arrays, final structured array, seconds)."""
        amp = np.asarray(amp)
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
        channels, n = amp.shape
        stride = amp.strides[0] // 2
        if ev_cap is None:
            ev_cap = max(16, n // 100 + 16)
        fin = np.zeros(channels, dtype=FINAL_DTYPE)
        if want_events:
            ev = np.zeros((channels, ev_cap), dtype=EVENT_DTYPE)
            cnt = np.zeros(channels, dtype=np.int64)
            secs = self.lib.ref_run(C.byref(params), amp.ctypes.data, stride, channels, n, nthreads,
                                    ev.ctypes.data, ev_cap, cnt.ctypes.data, fin.ctypes.data)
            if (cnt > ev_cap).any():
                raise RuntimeError("oracle event capacity exceeded: %d > %d" % (cnt.max(), ev_cap))
            events = [ev[c, :cnt[c]] for c in range(channels)]
        else:
            secs = self.lib.ref_run(C.byref(params), amp.ctypes.data, stride, channels, n, nthreads,
                                    None, 0, None, fin.ctypes.data)
            events = None
        return events, fin, secs

    # ---- generators (reference back end only) ----------------------------------------
    def dtmf_generate(self, digits, n, level=-10, twist=0, on_ms=-1, off_ms=-1, noise_seed=0, noise_dbm0=-100.0):
        out = np.zeros(n, dtype=np.int16)
        self.lib.ref_dtmf_generate(C.c_void_p(out.ctypes.data), C.c_int(n), digits.encode(), C.c_int(level), C.c_int(twist),
                                   C.c_int(on_ms), C.c_int(off_ms), C.c_int(noise_seed), C.c_float(noise_dbm0))
        return out

    def bell_mf_generate(self, digits, n, noise_seed=0, noise_dbm0=-100.0):
        out = np.zeros(n, dtype=np.int16)
        self.lib.ref_bell_mf_generate(C.c_void_p(out.ctypes.data), C.c_int(n), digits.encode(),
                                      C.c_int(noise_seed), C.c_float(noise_dbm0))
        return out

    def r2_mf_generate(self, digits, n, fwd=True, on_samples=800, off_samples=800, noise_seed=0, noise_dbm0=-100.0):
        out = np.zeros(n, dtype=np.int16)
        self.lib.ref_r2_mf_generate(C.c_void_p(out.ctypes.data), C.c_int(n), digits.encode(), C.c_int(int(fwd)),
                                    C.c_int(on_samples), C.c_int(off_samples), C.c_int(noise_seed), C.c_float(noise_dbm0))
        return out

    def cadence_generate(self, steps, n, noise_seed=0, noise_dbm0=-100.0):
        """steps: list of (f1_hz, f2_hz, level_dbm0, length_ms)."""
        flat = np.asarray(steps, dtype=np.int32).reshape(-1)
        out = np.zeros(n, dtype=np.int16)
        self.lib.ref_cadence_generate(C.c_void_p(out.ctypes.data), C.c_int(n), C.c_void_p(flat.ctypes.data),
                                      C.c_int(len(steps)), C.c_int(noise_seed), C.c_float(noise_dbm0))
        return out

    def awgn_add(self, amp, seed, level_dbm0):
        assert amp.dtype == np.int16 and amp.flags.c_contiguous
        self.lib.ref_awgn_add(C.c_void_p(amp.ctypes.data), C.c_int(amp.size), C.c_int(seed), C.c_float(level_dbm0))
        return amp

    def goertzel_fac(self, freq, samples):
        return float(self.lib.ref_goertzel_fac(freq, samples))

    def goertzel_blocks(self, freq, samples, amp):
        amp = np.ascontiguousarray(amp, dtype=np.int16)
        out = np.zeros(len(amp) // samples + 1, dtype=np.float32)
        nb = self.lib.ref_goertzel_blocks(freq, samples, amp.ctypes.data, len(amp), out.ctypes.data)
        return out[:nb]

    def super_tone_bins(self, params):
        fac = np.zeros(64, dtype=np.float32)
        self.lib.ref_super_tone_bins.restype = C.c_int
        n = self.lib.ref_super_tone_bins(C.byref(params), C.c_void_p(fac.ctypes.data), C.c_int(64))
        return fac[:n]


V29_SYM_DTYPE = np.dtype([("re", "<f4"), ("im", "<f4"), ("tre", "<f4"), ("tim", "<f4"), ("state", "<i4")])


def v29_generate(o, n, bit_rate=9600, tep=False, power_dbm0=-13.0, lfsr_seed=1, lead=0, noise_seed=1234567, noise_dbm0=-50.0):
    """v29_tx of PRBS data (+ awgn), `lead` samples of silence first.  Reference back ends only."""
    amp = np.zeros(n, dtype=np.int16)
    o.lib.ref_v29_generate.restype = C.c_int
    o.lib.ref_v29_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(bit_rate), C.c_int(int(tep)), C.c_float(power_dbm0),
                           C.c_uint32(lfsr_seed), C.c_int(lead), C.c_int(noise_seed), C.c_float(noise_dbm0))
    return amp


def v29_run(o, amp, bit_rate=9600, chunk=160, cutoff=-100.0, want_qam=True, restart_at=-1, restart_old_train=0):
    """One channel through the reference's v29_rx.  Returns dict(bits int8[], syms, eq_coeff[66], final[8])."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    bits = np.zeros(n * 3 // 2 + 64, dtype=np.int8)
    syms = np.zeros(n * 2 // 5 + 16, dtype=V29_SYM_DTYPE)
    nb = C.c_int32(0)
    ns = C.c_int32(0)
    eq = np.zeros(66, dtype=np.float32)
    fin = np.zeros(8, dtype=np.int32)
    o.lib.ref_v29_run_ex.restype = C.c_int
    rc = o.lib.ref_v29_run_ex(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(int(want_qam)),
                           C.c_int(restart_at), C.c_int(restart_old_train),
                           C.c_void_p(bits.ctypes.data), C.c_int(bits.size), C.byref(nb), C.c_void_p(syms.ctypes.data), C.c_int(syms.size),
                           C.byref(ns), C.c_void_p(eq.ctypes.data), C.c_void_p(fin.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_v29_run failed")
    return {"bits": bits[:nb.value].copy(), "syms": syms[:ns.value].copy(), "eq_coeff": eq, "final": fin}


def v29_run_batch(o, amp, bit_rate=9600, chunk=160, cutoff=-100.0, nthreads=1):
    """Many channels, timing only (CPU baseline).  Returns seconds."""
    amp = np.asarray(amp)
    assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
    o.lib.ref_v29_run_batch.restype = C.c_double
    return o.lib.ref_v29_run_batch(C.c_void_p(amp.ctypes.data), C.c_int64(amp.strides[0] // 2), C.c_int(amp.shape[0]), C.c_int(amp.shape[1]),
                                   C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(nthreads), None, C.c_int64(0), None)


def v29_tables(lib, fn_name):
    re = np.zeros(48 * 27, np.float32)
    im = np.zeros(48 * 27, np.float32)
    si = np.zeros(2048, np.float32)
    sq = np.zeros(193, np.uint16)
    g = np.zeros(9, np.float32)
    it = np.zeros(9, np.int32)
    getattr(lib, fn_name)(C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data), C.c_void_p(si.ctypes.data), C.c_void_p(sq.ctypes.data),
                          C.c_void_p(g.ctypes.data), C.c_void_p(it.ctypes.data))
    return {"rrc_re": re, "rrc_im": im, "sine": si, "sqrt": sq, "godard": g, "ints": it}


def v17_generate(o, n, bit_rate=14400, tep=False, power_dbm0=-13.0, lfsr_seed=1, lead=0, burst1=-1, gap=0, burst2=0,
                 noise_seed=1234567, noise_dbm0=-50.0):
    """v17_tx of PRBS data (+ awgn): silence, a long-trained burst, optionally a gap and a short-trained burst."""
    amp = np.zeros(n, dtype=np.int16)
    o.lib.ref_v17_generate.restype = C.c_int
    rc = o.lib.ref_v17_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(bit_rate), C.c_int(int(tep)), C.c_float(power_dbm0),
                                C.c_uint32(lfsr_seed), C.c_int(lead), C.c_int(burst1), C.c_int(gap), C.c_int(burst2),
                                C.c_int(noise_seed), C.c_float(noise_dbm0))
    if rc < 0:
        raise RuntimeError("ref_v17_generate failed")
    return amp


def v17_run(o, amp, bit_rate=14400, chunk=160, cutoff=-100.0, want_qam=True, restart_at=-1, restart_short=1):
    """One channel through the reference's v17_rx.  Returns dict(bits int8[], syms, eq_coeff[66], final[10])."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    bits = np.zeros(n * 2 + 64, dtype=np.int8)
    syms = np.zeros(n * 2 // 5 + 16, dtype=V29_SYM_DTYPE)
    nb = C.c_int32(0)
    ns = C.c_int32(0)
    eq = np.zeros(66, dtype=np.float32)
    fin = np.zeros(10, dtype=np.int32)
    o.lib.ref_v17_run.restype = C.c_int
    rc = o.lib.ref_v17_run(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(int(want_qam)),
                           C.c_int(restart_at), C.c_int(restart_short),
                           C.c_void_p(bits.ctypes.data), C.c_int(bits.size), C.byref(nb), C.c_void_p(syms.ctypes.data), C.c_int(syms.size),
                           C.byref(ns), C.c_void_p(eq.ctypes.data), C.c_void_p(fin.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_v17_run failed")
    return {"bits": bits[:nb.value].copy(), "syms": syms[:ns.value].copy(), "eq_coeff": eq, "final": fin}


def v17_run_batch(o, amp, bit_rate=14400, chunk=160, cutoff=-100.0, nthreads=1):
    """Many channels, timing only (CPU baseline).  Returns seconds."""
    amp = np.asarray(amp)
    assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
    o.lib.ref_v17_run_batch.restype = C.c_double
    return o.lib.ref_v17_run_batch(C.c_void_p(amp.ctypes.data), C.c_int64(amp.strides[0] // 2), C.c_int(amp.shape[0]), C.c_int(amp.shape[1]),
                                   C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(nthreads))


def v17_tables(lib, fn_name):
    re = np.zeros(192 * 27, np.float32)
    im = np.zeros(192 * 27, np.float32)
    g = np.zeros(9, np.float32)
    it = np.zeros(12, np.int32)
    con = np.zeros(244 * 2, np.float32)
    maps = np.zeros(4 * 36 * 36 * 8, np.uint8)
    m48 = np.zeros(36 * 36, np.uint8)
    getattr(lib, fn_name)(C.c_void_p(re.ctypes.data), C.c_void_p(im.ctypes.data), C.c_void_p(g.ctypes.data), C.c_void_p(it.ctypes.data),
                          C.c_void_p(con.ctypes.data), C.c_void_p(maps.ctypes.data), C.c_void_p(m48.ctypes.data))
    return {"rrc_re": re, "rrc_im": im, "godard": g, "ints": it, "constellations": con, "maps": maps, "map4800": m48}


def v27ter_generate(o, n, bit_rate=4800, tep=False, power_dbm0=-13.0, lfsr_seed=1, lead=0, burst1=-1, gap=0, burst2=0,
                    noise_seed=1234567, noise_dbm0=-50.0):
    """v27ter_tx of PRBS data (+ awgn): silence, a burst, optionally a gap and a second burst."""
    amp = np.zeros(n, dtype=np.int16)
    o.lib.ref_v27ter_generate.restype = C.c_int
    rc = o.lib.ref_v27ter_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(bit_rate), C.c_int(int(tep)), C.c_float(power_dbm0),
                                   C.c_uint32(lfsr_seed), C.c_int(lead), C.c_int(burst1), C.c_int(gap), C.c_int(burst2),
                                   C.c_int(noise_seed), C.c_float(noise_dbm0))
    if rc < 0:
        raise RuntimeError("ref_v27ter_generate failed")
    return amp


def v27ter_run(o, amp, bit_rate=4800, chunk=160, cutoff=-100.0, want_qam=True, restart_at=-1, restart_old_train=0):
    """One channel through the reference's v27ter_rx.  Gardner hops (qam_report with NULL pointers) appear as
    symbols with NaN coordinates and state = the integrator value.  Returns dict(bits, syms, eq_coeff[64], final[10])."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    bits = np.zeros(n * 2 + 64, dtype=np.int8)
    syms = np.zeros(n * 2 // 5 + 16, dtype=V29_SYM_DTYPE)
    nb = C.c_int32(0)
    ns = C.c_int32(0)
    eq = np.zeros(64, dtype=np.float32)
    fin = np.zeros(10, dtype=np.int32)
    o.lib.ref_v27ter_run.restype = C.c_int
    rc = o.lib.ref_v27ter_run(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(int(want_qam)),
                              C.c_int(restart_at), C.c_int(restart_old_train),
                              C.c_void_p(bits.ctypes.data), C.c_int(bits.size), C.byref(nb), C.c_void_p(syms.ctypes.data), C.c_int(syms.size),
                              C.byref(ns), C.c_void_p(eq.ctypes.data), C.c_void_p(fin.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_v27ter_run failed")
    return {"bits": bits[:nb.value].copy(), "syms": syms[:ns.value].copy(), "eq_coeff": eq, "final": fin}


def v27ter_run_batch(o, amp, bit_rate=4800, chunk=160, cutoff=-100.0, nthreads=1):
    """Many channels, timing only (CPU baseline).  Returns seconds."""
    amp = np.asarray(amp)
    assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
    o.lib.ref_v27ter_run_batch.restype = C.c_double
    return o.lib.ref_v27ter_run_batch(C.c_void_p(amp.ctypes.data), C.c_int64(amp.strides[0] // 2), C.c_int(amp.shape[0]), C.c_int(amp.shape[1]),
                                      C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff), C.c_int(nthreads))


def v27ter_tables(lib, fn_name):
    r48 = np.zeros(8 * 27, np.float32)
    i48 = np.zeros(8 * 27, np.float32)
    r24 = np.zeros(12 * 27, np.float32)
    i24 = np.zeros(12 * 27, np.float32)
    it = np.zeros(8, np.int32)
    getattr(lib, fn_name)(C.c_void_p(r48.ctypes.data), C.c_void_p(i48.ctypes.data), C.c_void_p(r24.ctypes.data), C.c_void_p(i24.ctypes.data),
                          C.c_void_p(it.ctypes.data))
    return {"rrc4800_re": r48, "rrc4800_im": i48, "rrc2400_re": r24, "rrc2400_im": i24, "ints": it}


FSK_FINAL = 28


def fsk_generate(o, n, spec=1, level_dbm0=1.0, lfsr_seed=1, char_bits=0, parity=0, idle=2, lead=0, burst=-1, noise_seed=1234567, noise_dbm0=-100.0):
    """fsk_tx of PRBS data (raw bits, or start-stop characters when char_bits > 0) + awgn.  level_dbm0 > 0: the spec's tx level."""
    amp = np.zeros(n, dtype=np.int16)
    o.lib.ref_fsk_generate.restype = C.c_int
    rc = o.lib.ref_fsk_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(spec), C.c_float(level_dbm0), C.c_uint32(lfsr_seed),
                                C.c_int(char_bits), C.c_int(parity), C.c_int(idle), C.c_int(lead), C.c_int(burst),
                                C.c_int(noise_seed), C.c_float(noise_dbm0))
    if rc < 0:
        raise RuntimeError("ref_fsk_generate failed")
    return amp


def fsk_run(o, amp, spec=1, framing_mode=1, chunk=160, cutoff=-100.0, frame=(0, 0, 0), restart=(-1, 0, 0), fillin=(-1, 0)):
    """One channel through the reference's fsk_rx.  Returns dict(out int16[], final int32[28], window int32[2,128,2])."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    out = np.zeros(n * 2 + 64, dtype=np.int16)
    nout = C.c_int32(0)
    fin = np.zeros(FSK_FINAL, dtype=np.int32)
    win = np.zeros((2, 128, 2), dtype=np.int32)
    o.lib.ref_fsk_run.restype = C.c_int
    rc = o.lib.ref_fsk_run(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(spec), C.c_int(framing_mode), C.c_float(cutoff),
                           C.c_int(frame[0]), C.c_int(frame[1]), C.c_int(frame[2]),
                           C.c_int(restart[0]), C.c_int(restart[1]), C.c_int(restart[2]), C.c_int(fillin[0]), C.c_int(fillin[1]),
                           C.c_void_p(out.ctypes.data), C.c_int(out.size), C.byref(nout), C.c_void_p(fin.ctypes.data), C.c_void_p(win.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_fsk_run failed")
    return {"out": out[:nout.value].copy(), "final": fin, "window": win}


def fsk_run_batch(o, amp, spec=1, framing_mode=1, chunk=160, nthreads=1):
    """Many channels, timing only (CPU baseline).  Returns seconds."""
    amp = np.asarray(amp)
    assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
    o.lib.ref_fsk_run_batch.restype = C.c_double
    return o.lib.ref_fsk_run_batch(C.c_void_p(amp.ctypes.data), C.c_int64(amp.strides[0] // 2), C.c_int(amp.shape[0]), C.c_int(amp.shape[1]),
                                   C.c_int(chunk), C.c_int(spec), C.c_int(framing_mode), C.c_int(nthreads))


def fsk_tables(o):
    sine = np.zeros(257, np.int16)
    derived = np.zeros((11, 4), np.int32)
    presets = np.zeros((11, 5), np.int32)
    o.lib.ref_fsk_tables(C.c_void_p(sine.ctypes.data), C.c_void_p(derived.ctypes.data), C.c_void_p(presets.ctypes.data))
    return {"sine": sine, "derived": derived, "presets": presets}


MCT_FINAL = 16


def mct_generate(o, n, tone_type, freq=0.0, level_dbm0=1.0, mod_freq=0.0, lead=0, burst=-1, flags=40, lfsr_seed=1,
                 noise_seed=1234567, noise_dbm0=-100.0, into=None):
    """modem_connect_tones_tx (types 1..5, 8, 9) or V.21 ch 2 flags + data (type 6) ADDED into a buffer, then awgn.
    level_dbm0 > 0 / freq <= 0 / mod_freq <= 0: the generator's defaults."""
    amp = np.zeros(n, dtype=np.int16) if into is None else into
    assert amp.dtype == np.int16 and len(amp) == n and amp.flags["C_CONTIGUOUS"]
    o.lib.ref_mct_generate.restype = C.c_int
    rc = o.lib.ref_mct_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(tone_type), C.c_float(freq), C.c_float(level_dbm0),
                                C.c_float(mod_freq), C.c_int(lead), C.c_int(burst), C.c_int(flags), C.c_uint32(lfsr_seed),
                                C.c_int(noise_seed), C.c_float(noise_dbm0))
    if rc < 0:
        raise RuntimeError("ref_mct_generate failed")
    return amp


def mct_run(o, amp, tone_type, chunk=160, use_callback=True):
    """One channel through the reference's modem_connect_tones_rx.  Returns dict(ev int32[n,3] = (call, tone, level),
    final int32[16], fsk_final int32[28])."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    cap = 4096
    ev = np.zeros((cap, 3), dtype=np.int32)
    nev = C.c_int32(0)
    fin = np.zeros(MCT_FINAL, dtype=np.int32)
    ffin = np.zeros(FSK_FINAL, dtype=np.int32)
    o.lib.ref_mct_run.restype = C.c_int
    rc = o.lib.ref_mct_run(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(tone_type), C.c_int(1 if use_callback else 0),
                           C.c_void_p(ev.ctypes.data), C.c_int(cap), C.byref(nev), C.c_void_p(fin.ctypes.data), C.c_void_p(ffin.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_mct_run failed")
    if nev.value > cap:
        raise RuntimeError("mct event buffer overflow")
    return {"ev": ev[:nev.value].copy(), "final": fin, "fsk_final": ffin}


def mct_run_batch(o, amp, tone_type, chunk=160, nthreads=1):
    """Many channels, timing only (CPU baseline).  Returns seconds."""
    amp = np.asarray(amp)
    assert amp.dtype == np.int16 and amp.ndim == 2 and amp.strides[1] == 2
    o.lib.ref_mct_run_batch.restype = C.c_double
    return o.lib.ref_mct_run_batch(C.c_void_p(amp.ctypes.data), C.c_int64(amp.strides[0] // 2), C.c_int(amp.shape[0]), C.c_int(amp.shape[1]),
                                   C.c_int(chunk), C.c_int(tone_type), C.c_int(nthreads))


def dtmf_tx_calls(o, max_lens, digits, digits2=None, put2_before_call=0, level=None, timing=None, fill=0x5555):
    """The reference's dtmf_tx(): one put, then len(max_lens) calls.  Returns (amp with `fill` where nothing was
    written, lens returned, put results)."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    puts = np.zeros(2, dtype=np.int32)
    o.lib.ref_dtmf_tx_calls.restype = C.c_int
    rc = o.lib.ref_dtmf_tx_calls(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)),
                                 C.c_char_p(digits.encode()), C.c_char_p(digits2.encode()) if digits2 is not None else None,
                                 C.c_int(put2_before_call),
                                 C.c_int(0 if level is None else 1), C.c_int(0 if level is None else level[0]), C.c_int(0 if level is None else level[1]),
                                 C.c_int(0 if timing is None else 1), C.c_int(0 if timing is None else timing[0]), C.c_int(0 if timing is None else timing[1]),
                                 C.c_void_p(out_lens.ctypes.data), C.c_void_p(puts.ctypes.data))
    if rc != 0:
        raise RuntimeError("ref_dtmf_tx_calls failed")
    return amp, out_lens, puts


def awgn_run(o, n, seed, level, dbov=False, into=None):
    """n x the reference's awgn(); added (saturating) to `into` if given."""
    amp = np.zeros(n, dtype=np.int16) if into is None else into
    assert amp.dtype == np.int16 and len(amp) == n and amp.flags["C_CONTIGUOUS"]
    o.lib.ref_awgn_run.restype = C.c_int
    rc = o.lib.ref_awgn_run(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(seed), C.c_float(level), C.c_int(1 if dbov else 0),
                            C.c_int(0 if into is None else 1))
    if rc != 0:
        raise RuntimeError("ref_awgn_run failed")
    return amp


def gen_tables(o):
    sine = np.zeros(2048, np.float32)
    rates = np.zeros(8, np.int32)
    gains = np.zeros(3, np.float32)
    o.lib.ref_gen_tables(C.c_void_p(sine.ctypes.data), C.c_void_p(rates.ctypes.data), C.c_void_p(gains.ctypes.data))
    return {"sine": sine, "rates": rates, "gains": gains}


SIG_FINAL = 31


def sig_generate(o, n, tone_type, steps, tone_db=-100.0, freq_offset=0.0, noise_seed=1234567, noise_dbm0=-100.0, into=None):
    """sig_tone_tx() script ADDED into a buffer: steps = [(mode bits, samples), ...]; then awgn."""
    amp = np.zeros(n, dtype=np.int16) if into is None else into
    assert amp.dtype == np.int16 and len(amp) == n and amp.flags["C_CONTIGUOUS"]
    st = np.asarray(steps, dtype=np.int32).reshape(-1, 2)
    o.lib.ref_sig_generate.restype = C.c_int
    rc = o.lib.ref_sig_generate(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(tone_type), C.c_void_p(st.ctypes.data), C.c_int(len(st)),
                                C.c_float(tone_db), C.c_float(freq_offset), C.c_int(noise_seed), C.c_float(noise_dbm0))
    if rc < 0:
        raise RuntimeError("ref_sig_generate failed")
    return amp


def sig_run(o, amp, tone_type, chunk=160, lens=None, modes=((0, 0x40),)):
    """One channel through the reference's sig_tone_rx (a copy of amp is processed in place).  modes = ((call, mode), ...).
    Returns dict(out int16[], ev int32[n,3] = (call, signalling_state, duration), final int32[31])."""
    out = np.array(amp, dtype=np.int16, copy=True)
    n = len(out)
    cap = 8192
    ev = np.zeros((cap, 3), dtype=np.int32)
    nev = C.c_int32(0)
    fin = np.zeros(SIG_FINAL, dtype=np.int32)
    md = np.asarray(modes, dtype=np.int32).reshape(-1, 2)
    ln = None if lens is None else np.asarray(lens, dtype=np.int32)
    o.lib.ref_sig_run.restype = C.c_int
    rc = o.lib.ref_sig_run(C.c_void_p(out.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_void_p(ln.ctypes.data) if ln is not None else None,
                           C.c_int(0 if ln is None else len(ln)), C.c_int(tone_type), C.c_void_p(md.ctypes.data), C.c_int(len(md)),
                           C.c_void_p(ev.ctypes.data), C.c_int(cap), C.byref(nev), C.c_void_p(fin.ctypes.data))
    if rc != 0 or nev.value > cap:
        raise RuntimeError("ref_sig_run failed")
    return {"out": out, "ev": ev[:nev.value].copy(), "final": fin}


def tone_burst(o, f1, l1, f2, l2, on_ms, off_ms, max_samples=1000):
    """One test burst of tests/dtmf_rx_tests.c / bell_mf_rx_tests.c (my_*_gen_init + my_*_generate), reference tone_gen()."""
    amp = np.zeros(max_samples, dtype=np.int16)
    o.lib.ref_tone_burst.restype = C.c_int
    n = o.lib.ref_tone_burst(C.c_void_p(amp.ctypes.data), C.c_int(max_samples), C.c_int(int(f1)), C.c_int(int(l1)), C.c_int(int(f2)), C.c_int(int(l2)),
                             C.c_int(int(on_ms)), C.c_int(int(off_ms)))
    return amp[:n]


def super_tone_range_stimulus(o, first_level=-80, last_level=-1, chunks=100):
    """The dial tone sweep of tests/super_tone_rx_tests.c:detection_range_tests()."""
    n = (last_level - first_level + 1) * chunks * 160
    amp = np.zeros(n, dtype=np.int16)
    o.lib.ref_super_tone_range_stimulus.restype = C.c_int
    got = o.lib.ref_super_tone_range_stimulus(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(first_level), C.c_int(last_level), C.c_int(chunks))
    assert got == n
    return amp


_cache = {}


def _path(kind):
    if kind == "port":
        return os.path.join(HERE, "libtonebank_oracle.so")
    return os.path.join(HERE, "_ref", "libspandsp_ref_%s.so" % kind)


def available(kind):
    return os.path.exists(_path(kind))


def load(kind="strict"):
    """kind: 'strict' (pinned reference build), 'fast' (reference default flags) or 'port'."""
    if kind not in _cache:
        p = _path(kind)
        if not os.path.exists(p):
            raise FileNotFoundError("oracle library %s is not built (run `make -C oracle`)" % p)
        _cache[kind] = Oracle(p, kind)
    return _cache[kind]


def events_tuple(ev, with_chunk=False):
    """Structured event array -> list of plain tuples for easy comparison."""
    if with_chunk:
        return [(int(e["chunk"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev]
    return [(int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev]

def tone_gen_calls(o, max_lens, desc, fill=0x5555):
    """tone_gen() with descriptor desc = (f1, l1, f2, l2, d1, d2, d3, d4, repeat): len(max_lens) calls.
    Returns (amp with `fill` where nothing was written, lens returned)."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    d = np.asarray(desc, dtype=np.int32)
    assert d.shape == (9,)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    fn = o.lib.ref_tone_gen_calls
    fn.restype = C.c_int
    if fn(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)), C.c_void_p(d.ctypes.data),
          C.c_void_p(out_lens.ctypes.data)) != 0:
        raise RuntimeError("tone_gen_calls failed")
    return amp, out_lens


def v29_tx_calls(o, max_lens, bit_rate=9600, tep=False, power_dbm0=-14.0, lfsr_seed=None, bits=None, nbits=0, restart=(-1, 9600, False),
                 fill=0x5555):
    """v29_tx(): len(max_lens) calls.  Data: the 23-bit sequence seeded lfsr_seed, or `bits` (uint8, LSB first, nbits of
    them, then end of data).  restart = (before call, bit rate, tep).  Returns (amp, lens returned, status bits)."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    status = C.c_int32(0)
    b = None if bits is None else np.ascontiguousarray(bits, dtype=np.uint8)
    fn = o.lib.ref_v29_tx_calls
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)), C.c_int(bit_rate), C.c_int(int(tep)),
            C.c_float(power_dbm0), C.c_int(0 if bits is None else 1), C.c_uint32(1 if lfsr_seed is None else lfsr_seed),
            C.c_void_p(None if b is None else b.ctypes.data), C.c_int(nbits),
            C.c_int(restart[0]), C.c_int(restart[1]), C.c_int(int(restart[2])),
            C.c_void_p(out_lens.ctypes.data), C.byref(status))
    if rc != 0:
        raise RuntimeError("v29_tx_calls failed")
    return amp, out_lens, status.value


def v29_tx_tables(o):
    t = np.zeros((10, 9), dtype=np.float32)
    o.lib.ref_v29_tx_tables(C.c_void_p(t.ctypes.data))
    return t
