"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tonebank_golden.npz")


def golden():
    return np.load(GOLDEN)


def oracle_rows(events_per_channel, with_chunk=True):
    """Oracle events -> sorted list of (channel, [chunk,] kind, a, b, c) rows."""
    rows = []
    for c, ev in enumerate(events_per_channel):
        for e in ev:
            if with_chunk:
                rows.append((c, int(e["chunk"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
            else:
                rows.append((c, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
    return rows


def golden_rows(arr, with_chunk=True):
    if with_chunk:
        return [tuple(int(x) for x in r) for r in arr]
    return [(int(r[0]), int(r[2]), int(r[3]), int(r[4]), int(r[5])) for r in arr]


def normalise(rows):
    """Make oracle rows and engine rows comparable: digit events carry only the digit."""
    out = []
    for r in rows:
        c, kind, a, b, cc = r[0], r[-4], r[-3], r[-2], r[-1]
        if kind == po.EV_DIGIT:
            b, cc = 0, 0
        out.append((c, kind, a, b, cc))
    return out


def engine_rows(ev):
    return [(int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev]


def run_engine_chunked(bank, amp, chunk, torch, via_host=False):
    """Feed amp [channels, n] to a bank in `chunk`-sample calls; return all events as rows
    (channel, kind, a, b, c) in per-channel time order."""
    channels, n = amp.shape
    per_chan = [[] for _ in range(channels)]
    if not via_host:
        d = torch.from_numpy(np.ascontiguousarray(amp)).cuda()
    pos = 0
    while pos < n:
        ln = min(chunk, n - pos)
        if via_host:
            bank.rx_host(np.ascontiguousarray(amp[:, pos:pos + ln]))
        else:
            bank.rx_device(d.data_ptr() + 2 * pos, amp.shape[1], ln)
        for e in bank.events():
            per_chan[int(e["channel"])].append((int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
        pos += ln
    return [r for ch in per_chan for r in ch]
