"""Synthetic telephony signals for tests and the benchmark (numpy / torch, our own code).

These are NOT the reference's generators: they only have to produce plausible int16 input
(dual tones with the nominal DTMF/MF/R2 frequencies, cadenced supervisory tones, Gaussian
noise).  What the detectors make of them is decided by the oracle, sample for sample.
"""
import numpy as np

DTMF_DIGITS = "123A456B789C*0#D"
DTMF_ROW = [697.0, 770.0, 852.0, 941.0]
DTMF_COL = [1209.0, 1336.0, 1477.0, 1633.0]
BELL_MF_FREQS = [700.0, 900.0, 1100.0, 1300.0, 1500.0, 1700.0]
R2_FWD_FREQS = [1380.0, 1500.0, 1620.0, 1740.0, 1860.0, 1980.0]
R2_BACK_FREQS = [1140.0, 1020.0, 900.0, 780.0, 660.0, 540.0]


def dbm0_to_amp(level):
    """Peak amplitude of a sine of `level` dBm0 (0 dBm0 sine peak = 32768*10^(-3.14/20))."""
    return 32768.0 * 10.0 ** ((level - 3.14) / 20.0)


def dual_tone(n, f1, f2, a1, a2, phase=0.0):
    t = np.arange(n, dtype=np.float64)
    x = a1 * np.sin(2 * np.pi * f1 * t / 8000.0 + phase)
    if f2:
        x = x + a2 * np.sin(2 * np.pi * f2 * t / 8000.0 + 1.3 * phase)
    return x


def finish(x, rng, noise_dbm0):
    if noise_dbm0 is not None and noise_dbm0 > -99:
        rms = 32768.0 * 10.0 ** ((noise_dbm0 - 3.14 - 3.02) / 20.0)
        x = x + rng.normal(0.0, rms, size=x.shape)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def dtmf_channels(channels, n, seed=0, on=400, off=440, level=(-25, -5), twist=(-5, 5), noise=(-50, -25),
                  jitter=True):
    """[channels, n] int16 of random DTMF digit trains."""
    rng = np.random.default_rng(seed)
    out = np.zeros((channels, n), dtype=np.int16)
    digits = []
    for c in range(channels):
        x = np.zeros(n, dtype=np.float64)
        pos = int(rng.integers(0, 200)) if jitter else 0
        dl = []
        while pos + on < n:
            d = int(rng.integers(0, 16))
            lv = float(rng.uniform(*level))
            tw = float(rng.uniform(*twist))
            o = on + (int(rng.integers(-80, 240)) if jitter else 0)
            g = off + (int(rng.integers(-100, 300)) if jitter else 0)
            o = min(o, n - pos)
            x[pos:pos + o] = dual_tone(o, DTMF_ROW[d >> 2], DTMF_COL[d & 3], dbm0_to_amp(lv), dbm0_to_amp(lv + tw),
                                       float(rng.uniform(0, 6.28)))
            dl.append(DTMF_DIGITS[d])
            pos += o + g
        out[c] = finish(x, rng, float(rng.uniform(*noise)))
        digits.append("".join(dl))
    return out, digits


def mf_channels(channels, n, freqs, seed=0, on=560, off=560, level=(-20, -5), noise=(-55, -35)):
    rng = np.random.default_rng(seed)
    out = np.zeros((channels, n), dtype=np.int16)
    for c in range(channels):
        x = np.zeros(n, dtype=np.float64)
        pos = int(rng.integers(0, 300))
        while pos + on < n:
            i, j = rng.choice(6, size=2, replace=False)
            lv = float(rng.uniform(*level))
            o = min(on + int(rng.integers(-60, 400)), n - pos)
            x[pos:pos + o] = dual_tone(o, freqs[i], freqs[j], dbm0_to_amp(lv), dbm0_to_amp(lv + rng.uniform(-4, 4)),
                                       float(rng.uniform(0, 6.28)))
            pos += o + off + int(rng.integers(-100, 500))
        out[c] = finish(x, rng, float(rng.uniform(*noise)))
    return out


def cadence_channels(channels, n, cadences, seed=0, noise=(-60, -45)):
    """cadences: list of cadences; each a list of (f1, f2, level_dbm0, length_ms); f1 == 0 -> silence."""
    rng = np.random.default_rng(seed)
    out = np.zeros((channels, n), dtype=np.int16)
    for c in range(channels):
        cad = cadences[c % len(cadences)]
        x = np.zeros(n, dtype=np.float64)
        pos = int(rng.integers(0, 400))
        k = 0
        while pos < n:
            f1, f2, lv, ms = cad[k % len(cad)]
            ln = min(ms * 8, n - pos)
            if f1:
                x[pos:pos + ln] = dual_tone(ln, f1, f2, dbm0_to_amp(lv), dbm0_to_amp(lv), float(rng.uniform(0, 6.28)))
            pos += ln
            k += 1
        out[c] = finish(x, rng, float(rng.uniform(*noise)))
    return out


def random_tones(rng, nfreqs=8, ntones=6):
    """A random super-tone descriptor: list of tones, each a list of (f1, f2, min_ms, max_ms)."""
    freqs = sorted(int(f) for f in rng.choice(np.arange(300, 2000, 40), size=nfreqs, replace=False))
    tones = []
    for _ in range(ntones):
        steps = int(rng.integers(1, 5))
        t = []
        for s in range(steps):
            if s % 2 == 0:
                f1 = int(rng.choice(freqs))
                f2 = int(rng.choice(freqs)) if rng.random() < 0.5 else 0
                if f2 == f1:
                    f2 = 0
                ms = int(rng.integers(100, 1000))
                t.append((f1, f2, int(ms * 0.8), int(ms * 1.2)))
            else:
                ms = int(rng.integers(100, 1000))
                t.append((0, 0, int(ms * 0.8), int(ms * 1.2)))
        tones.append(t)
    return tones
