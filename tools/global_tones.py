"""Read a tone set out of the reference's spandsp/global-tones.xml the way tests/super_tone_rx_tests.c does
(parse_tone / parse_tone_set, :97-274) and print the super_tone_rx descriptor it leads to, followed by the two
tones of super_tone_rx_fill_descriptor() (:361-373), as JSON: [[[f1, f2, min_ms, max_ms], ...], ...].

    python tools/global_tones.py /root/reference/spandsp/global-tones.xml hk > tests/golden/global_tones_hk.json

The quirks of the test's parser are kept: it looks for "ringback-tone" while the file says "ringing-tone" (ringing
tones are skipped), container steps contribute only their children, frequencies are rounded by + 0.5 and truncation,
and the duration window is (length +- tolerance) -+ 30 ms unless a recognition length is given."""
import json
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

NAMES = ("dial-tone", "ringback-tone", "busy-tone", "number-unobtainable-tone", "congestion-tone", "waiting-tone")
F = np.float32


def scan_freq(x):
    """sscanf(x, "%f [%f%%]", &f1, &f_tol) then sscanf(x, "%f+%f [%f%%]", &f1, &f2, &f_tol)"""
    f1, f2 = 0.0, 0.0
    m = re.match(r"\s*([-+]?[0-9.]+)", x)
    if m:
        f1 = float(F(m.group(1)))
        m2 = re.match(r"\s*[-+]?[0-9.]+\+([0-9.]+)", x)
        if m2:
            f2 = float(F(m2.group(1)))
    return f1, f2


def scan_len(x, default_tol):
    """sscanf(x, "%f [%f%%]", &length, &tol)"""
    m = re.match(r"\s*([-+]?[0-9.]+)(?:\s*\[\s*([-+]?[0-9.]+)%\])?", x)
    length = float(F(m.group(1))) if m else 0.0
    tol = float(F(m.group(2))) if m and m.group(2) else default_tol
    return length, tol


def parse_tone(node, elements):
    for step in node:
        if step.tag != "step":
            continue
        f1, f2 = scan_freq(step.get("freq")) if step.get("freq") else (0.0, 0.0)
        length, length_tol = scan_len(step.get("length"), 10.0) if step.get("length") else (0.0, 10.0)
        rl, _ = scan_len(step.get("recognition-length"), 10.0) if step.get("recognition-length") else (0.0, 10.0)
        if f1 or f2 or length:
            if length == 0.0:
                lo = int(rl * 1000.0 + 0.5) if rl else 700
                hi = 0
            else:
                lo = int(rl * 1000.0 + 0.5) if rl else int((length * 1000.0 + 0.5) * (1.0 - length_tol / 100.0) - 30)
                hi = int((length * 1000.0 + 0.5) * (1.0 + length_tol / 100.0) + 30)
            elements.append([int(f1 + 0.5), int(f2 + 0.5), lo, hi])
        parse_tone(step, elements)


def tone_set(path, uncode):
    root = ET.parse(path).getroot()
    tones = []
    for ts in root:
        if ts.tag == "tone-set" and ts.get("uncode") == uncode:
            for tone in ts:
                if tone.tag in NAMES:
                    el = []
                    parse_tone(tone, el)
                    tones.append(el)
    tones.append([[400, 0, 700, 0]])                                # super_tone_rx_fill_descriptor(): "XXX"
    tones.append([[1100, 0, 400, 600], [0, 0, 2800, 3200]])         # ... and the FAX tone
    return tones


if __name__ == "__main__":
    print(json.dumps(tone_set(sys.argv[1], sys.argv[2])))
