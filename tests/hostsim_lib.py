"""ctypes loader for tests/hostsim (the modem receivers compiled for the host).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostsim", "modem_hostsim.cu")
LIB = os.path.join(HERE, "hostsim", "libmodem_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "spandsp_b200", "csrc")

SYM_DTYPE = np.dtype([("re", "<f4"), ("im", "<f4"), ("tre", "<f4"), ("tim", "<f4"), ("state", "<i4"), ("bit_pos", "<i4")])


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("sb_modem.cuh", "sb_v29_rx.cuh", "sb_v29_rx_body.cuh", "sb_v17_rx.cuh", "sb_v27ter_rx.cuh", "sb_fsk_rx.cuh", "sb_mct_rx.cuh", "sb_gen.cuh", "sb_sig_rx.cuh")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    subprocess.run([os.environ.get("NVCC", "nvcc"), "-O2", "-std=c++17", "-x", "cu", "-Wno-deprecated-gpu-targets",
                    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-o", LIB, SRC], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def run(modem, amp, bit_rate, chunk=160, cutoff=-100.0, restart_at=-1, restart_mode=1):
    """modem: 'v17', 'v29' or 'v27ter'.  Same result layout as pyoracle.v17_run / v29_run."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    bits = np.zeros(n * 2 + 64, dtype=np.int8)
    syms = np.zeros(n * 2 // 5 + 16, dtype=SYM_DTYPE)
    nb = C.c_int32(0)
    ns = C.c_int32(0)
    eq = np.zeros(66, dtype=np.float32)
    fin = np.zeros(10, dtype=np.int32)
    fn = getattr(lib(), "hostsim_%s_run" % modem)
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(bit_rate), C.c_float(cutoff),
            C.c_int(restart_at), C.c_int(restart_mode),
            C.c_void_p(bits.ctypes.data), C.c_int(bits.size), C.byref(nb), C.c_void_p(syms.ctypes.data), C.c_int(syms.size),
            C.byref(ns), C.c_void_p(eq.ctypes.data), C.c_void_p(fin.ctypes.data))
    if rc != 0:
        raise RuntimeError("hostsim run failed")
    return {"bits": bits[:nb.value].copy(), "syms": syms[:ns.value].copy(), "eq_coeff": eq, "final": fin}


def fsk_run(amp, spec5, framing_mode=1, chunk=160, cutoff=-100.0, frame=(0, 0, 0), restart=(-1, None, 0), fillin=(-1, 0)):
    """The FSK receiver of sb_fsk_rx.cuh on the host; same result layout as pyoracle.fsk_run.
    spec5 = (freq_zero, freq_one, tx_level, min_level, baud_rate x 100)."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    out = np.zeros(n * 2 + 64, dtype=np.int16)
    nout = C.c_int32(0)
    fin = np.zeros(28, dtype=np.int32)
    win = np.zeros((2, 128, 2), dtype=np.int32)
    s5 = np.asarray(spec5, dtype=np.int32)
    r5 = np.asarray(restart[1] if restart[1] is not None else spec5, dtype=np.int32)
    fn = lib().hostsim_fsk_run
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_void_p(s5.ctypes.data), C.c_int(framing_mode), C.c_float(cutoff),
            C.c_int(frame[0]), C.c_int(frame[1]), C.c_int(frame[2]),
            C.c_int(restart[0]), C.c_void_p(r5.ctypes.data), C.c_int(restart[2]), C.c_int(fillin[0]), C.c_int(fillin[1]),
            C.c_void_p(out.ctypes.data), C.c_int(out.size), C.byref(nout), C.c_void_p(fin.ctypes.data), C.c_void_p(win.ctypes.data))
    if rc != 0:
        raise RuntimeError("hostsim fsk run failed")
    return {"out": out[:nout.value].copy(), "final": fin, "window": win}


def mct_run(amp, tone_type, chunk=160):
    """The modem connect tone detector of sb_mct_rx.cuh on the host; same result layout as pyoracle.mct_run
    (final has a 17th entry, the accumulated hit)."""
    amp = np.ascontiguousarray(amp, dtype=np.int16)
    n = len(amp)
    cap = 4096
    ev = np.zeros((cap, 3), dtype=np.int32)
    nev = C.c_int32(0)
    fin = np.zeros(17, dtype=np.int32)
    ffin = np.zeros(28, dtype=np.int32)
    fn = lib().hostsim_mct_run
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_int(tone_type), C.c_void_p(ev.ctypes.data), C.c_int(cap),
            C.byref(nev), C.c_void_p(fin.ctypes.data), C.c_void_p(ffin.ctypes.data))
    if rc != 0 or nev.value > cap:
        raise RuntimeError("hostsim mct run failed")
    return {"ev": ev[:nev.value].copy(), "final": fin, "fsk_final": ffin}


def dtmf_tx_calls(max_lens, digits, digits2=None, put2_before_call=0, level=None, timing=None, fill=0x5555):
    """The DTMF transmitter of sb_gen.cuh on the host; same arguments and results as pyoracle.dtmf_tx_calls."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    puts = np.zeros(2, dtype=np.int32)
    fn = lib().hostsim_dtmf_tx_calls
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)),
            C.c_char_p(digits.encode()), C.c_char_p(digits2.encode()) if digits2 is not None else None, C.c_int(put2_before_call),
            C.c_int(0 if level is None else 1), C.c_int(0 if level is None else level[0]), C.c_int(0 if level is None else level[1]),
            C.c_int(0 if timing is None else 1), C.c_int(0 if timing is None else timing[0]), C.c_int(0 if timing is None else timing[1]),
            C.c_void_p(out_lens.ctypes.data), C.c_void_p(puts.ctypes.data))
    if rc != 0:
        raise RuntimeError("hostsim dtmf_tx failed")
    return amp, out_lens, puts


def awgn_run(n, seed, level, dbov=False, into=None):
    """The noise source of sb_gen.cuh on the host; same arguments and results as pyoracle.awgn_run."""
    amp = np.zeros(n, dtype=np.int16) if into is None else into
    fn = lib().hostsim_awgn_run
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_int(n), C.c_int(seed), C.c_float(level), C.c_int(1 if dbov else 0), C.c_int(0 if into is None else 1))
    if rc != 0:
        raise RuntimeError("hostsim awgn failed")
    return amp


def sig_run(amp, tone_type, chunk=160, lens=None, modes=((0, 0x40),)):
    """The signalling tone receiver of sb_sig_rx.cuh on the host; same arguments and results as pyoracle.sig_run."""
    out = np.array(amp, dtype=np.int16, copy=True)
    n = len(out)
    cap = 8192
    ev = np.zeros((cap, 3), dtype=np.int32)
    nev = C.c_int32(0)
    fin = np.zeros(31, dtype=np.int32)
    md = np.asarray(modes, dtype=np.int32).reshape(-1, 2)
    ln = None if lens is None else np.asarray(lens, dtype=np.int32)
    fn = lib().hostsim_sig_run
    fn.restype = C.c_int
    rc = fn(C.c_void_p(out.ctypes.data), C.c_int(n), C.c_int(chunk), C.c_void_p(ln.ctypes.data) if ln is not None else None,
            C.c_int(0 if ln is None else len(ln)), C.c_int(tone_type), C.c_void_p(md.ctypes.data), C.c_int(len(md)),
            C.c_void_p(ev.ctypes.data), C.c_int(cap), C.byref(nev), C.c_void_p(fin.ctypes.data))
    if rc != 0 or nev.value > cap:
        raise RuntimeError("hostsim sig run failed")
    return {"out": out, "ev": ev[:nev.value].copy(), "final": fin}

def tone_gen_calls(max_lens, desc, fill=0x5555):
    """tone_gen() with descriptor desc = (f1, l1, f2, l2, d1, d2, d3, d4, repeat): len(max_lens) calls.
    Returns (amp with `fill` where nothing was written, lens returned)."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    d = np.asarray(desc, dtype=np.int32)
    assert d.shape == (9,)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    fn = lib().hostsim_tone_gen_calls
    fn.restype = C.c_int
    if fn(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)), C.c_void_p(d.ctypes.data),
          C.c_void_p(out_lens.ctypes.data)) != 0:
        raise RuntimeError("tone_gen_calls failed")
    return amp, out_lens


def v29_tx_calls(max_lens, bit_rate=9600, tep=False, power_dbm0=-14.0, lfsr_seed=None, bits=None, nbits=0, restart=(-1, 9600, False),
                 fill=0x5555):
    """v29_tx(): len(max_lens) calls.  Data: the 23-bit sequence seeded lfsr_seed, or `bits` (uint8, LSB first, nbits of
    them, then end of data).  restart = (before call, bit rate, tep).  Returns (amp, lens returned, status bits)."""
    max_lens = np.asarray(max_lens, dtype=np.int32)
    amp = np.full(int(max_lens.sum()), fill, dtype=np.int16)
    out_lens = np.zeros(len(max_lens), dtype=np.int32)
    status = C.c_int32(0)
    b = None if bits is None else np.ascontiguousarray(bits, dtype=np.uint8)
    fn = lib().hostsim_v29_tx_calls
    fn.restype = C.c_int
    rc = fn(C.c_void_p(amp.ctypes.data), C.c_void_p(max_lens.ctypes.data), C.c_int(len(max_lens)), C.c_int(bit_rate), C.c_int(int(tep)),
            C.c_float(power_dbm0), C.c_int(0 if bits is None else 1), C.c_uint32(1 if lfsr_seed is None else lfsr_seed),
            C.c_void_p(None if b is None else b.ctypes.data), C.c_int(nbits),
            C.c_int(restart[0]), C.c_int(restart[1]), C.c_int(int(restart[2])),
            C.c_void_p(out_lens.ctypes.data), C.byref(status))
    if rc != 0:
        raise RuntimeError("v29_tx_calls failed")
    return amp, out_lens, status.value


def v29_tx_tables():
    t = np.zeros((10, 9), dtype=np.float32)
    lib().hostsim_v29_tx_tables(C.c_void_p(t.ctypes.data))
    return t
