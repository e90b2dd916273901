"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/*.h declares.  No compute calls here."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*;", txt):
        n = m.group(1)
        if n in ("defined", "sizeof", "__attribute__"):
            continue
        names.add(n)
    return names


def test_exports(engine_lib):
    import ctypes
    L = ctypes.CDLL(engine_lib.LIB_PATH)
    inc = os.path.join(ROOT, "include")
    missing = []
    total = 0
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        for fn in sorted(declared_functions(os.path.join(inc, h))):
            if fn.endswith("_t"):
                continue        # callback typedefs
            total += 1
            if not hasattr(L, fn):
                missing.append("%s:%s" % (h, fn))
    assert total > 20
    assert not missing, "declared but not exported: %s" % missing


def test_no_gpu_fails_loudly(engine_lib):
    """Without a CUDA device context creation must fail with an explanation, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        engine_lib.Context(0)
    except engine_lib.EngineError as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("Context() succeeded without a GPU")


def test_product_does_not_touch_oracle():
    """Nothing under spandsp_b200/ may import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "spandsp_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("strict CPU oracle", "").lower() \
                    or f == "sb_common.cuh" or "oracle/" not in txt, "%s mentions oracle/" % f
                assert "tonebank_oracle" not in txt and "pyoracle" not in txt and "libspandsp_ref" not in txt, f


def declared_prototypes(path):
    """name -> number of parameters, for every function prototype in a header."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    out = {}
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(([^;{}()]*(?:\([^()]*\)[^;{}()]*)*)\)\s*;", txt):
        name, args = m.group(1), m.group(2).strip()
        if name in ("defined", "sizeof", "__attribute__") or name.endswith("_t"):
            continue
        depth = 0
        n = 1
        for ch in args:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 0:
                n += 1
        out[name] = 0 if args in ("", "void") else n
    return out


def test_python_binding_matches_headers(engine_lib):
    """Every entry of the ctypes signature table in spandsp_b200/engine.py has as many arguments as the prototype in
    include/*.h declares (a drifted binding would corrupt the call without any error)."""
    L = engine_lib.lib()
    inc = os.path.join(ROOT, "include")
    protos = {}
    for h in sorted(os.listdir(inc)):
        if h.endswith(".h") and h != "spandsp_b200_dropin.h":
            protos.update(declared_prototypes(os.path.join(inc, h)))
    checked = 0
    for name, nargs in protos.items():
        fn = getattr(L, name, None)
        if fn is None or fn.argtypes is None:
            continue
        assert len(fn.argtypes) == nargs, "%s: header declares %d parameters, binding passes %d" % (name, nargs, len(fn.argtypes))
        checked += 1
    assert checked > 100


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI for C callers: every header under include/ must compile as C99 on its own."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        src = tmp_path / "use.c"
        src.write_text('#include "%s"\nint main(void) { return 0; }\n' % h)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, "%s: %s" % (h, r.stderr[:500])


def test_c_program_links_against_dropin(tmp_path, engine_lib):
    """A plain C caller written against the reference's names links against the library; without a GPU the
    constructors return NULL and the library says why (no fallback)."""
    import subprocess
    import torch
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
/* the reference's own header names: a caller written against spandsp compiles with no source edit */
#include <spandsp/telephony.h>
#include <spandsp/async.h>
#include <spandsp/dtmf.h>
#include <spandsp/bell_r2_mf.h>
#include <spandsp/super_tone_rx.h>
#include <spandsp/tone_detect.h>
#include <spandsp/v29rx.h>
#include <spandsp/v17rx.h>
#include <spandsp/v27ter_rx.h>
#include <spandsp/fsk.h>
#include <spandsp/modem_connect_tones.h>
#include <spandsp/sig_tone.h>

static void digits(void *user, const char *d, int len) { (void) user; (void) d; (void) len; }
static void tone(void *user, int code, int level, int delay) { (void) user; (void) code; (void) level; (void) delay; }

int main(void)
{
    int16_t amp[160];
    char buf[8];
    dtmf_rx_state_t *s;
    modem_connect_tones_rx_state_t *m;
    sig_tone_rx_state_t *g;

    memset(amp, 0, sizeof(amp));
    if (SIG_STATUS_TRAINING_SUCCEEDED != -4  ||  SAMPLE_RATE != 8000)
        return 5;
    s = dtmf_rx_init(NULL, digits, NULL);
    m = modem_connect_tones_rx_init(NULL, MODEM_CONNECT_TONES_FAX_CNG, tone, NULL);
    g = sig_tone_rx_init(NULL, SIG_TONE_2280HZ, tone, NULL);
    if (s == NULL  ||  m == NULL  ||  g == NULL)
    {
        printf("no-gpu: %s\n", span_b200_last_error());
        return (s == NULL  &&  m == NULL  &&  g == NULL)  ?  3  :  4;
    }
    dtmf_rx(s, amp, 160);
    modem_connect_tones_rx(m, amp, 160);
    sig_tone_rx(g, amp, 160);
    printf("digits=%d hit=%d\n", (int) dtmf_rx_get(s, buf, 7), modem_connect_tones_rx_get(m));
    dtmf_rx_free(s);
    modem_connect_tones_rx_free(m);
    sig_tone_rx_free(g);
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(engine_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                        "-L", libdir, "-lspandsp_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:800]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "digits=0 hit=0" in r.stdout, (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 3 and "no-gpu:" in r.stdout and len(r.stdout.strip()) > len("no-gpu:"), (r.returncode, r.stdout)
