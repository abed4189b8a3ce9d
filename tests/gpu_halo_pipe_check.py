"""tc_halo_pipe_kernel vs the SIMT conv (genie_debug_tc_selftest: two ragged segments, residual, pre-activation) at
the generator's C = 128 shapes, with timings.  GENIE_HALO_PIPE=0 runs the one-tile kernel for comparison."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genie-tts_b200")]
os.environ["GENIE_SELFTEST_TIME"] = "1"
from genie_tts import _native as N  # noqa: E402

L = N.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1440000
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cout = int(sys.argv[3]) if len(sys.argv) > 3 else cin
bad = 0
for (taps, dil) in ((3, 1), (3, 5), (7, 1), (7, 3), (11, 1), (11, 5)):
    e, r = C.c_float(0), C.c_float(0)
    print(f"M={M} C={cin}->{cout} taps={taps} dil={dil}", flush=True)
    rc = L.genie_debug_tc_selftest(M, cin, cout, taps, dil, 1, 0, C.byref(e), C.byref(r))
    rel = e.value / max(r.value, 1e-9)
    print(f"   rc={rc} max err {e.value:.3e} (ref max {r.value:.3e}, rel {rel:.2e})", flush=True)
    if rc != 0 or not (rel < 5e-3):
        bad += 1
        print("   FAILED:", N.last_error() if hasattr(N, "last_error") else "", flush=True)
# a small ragged case: tiles that end inside a segment, fewer tiles than CTAs
for (m, taps, dil) in ((1000, 7, 3), (129, 3, 1), (40000, 11, 5)):
    e, r = C.c_float(0), C.c_float(0)
    rc = L.genie_debug_tc_selftest(m, cin, cout, taps, dil, 1, 0, C.byref(e), C.byref(r))
    rel = e.value / max(r.value, 1e-9)
    print(f"M={m} taps={taps} dil={dil}: rc={rc} rel {rel:.2e}", flush=True)
    if rc != 0 or not (rel < 5e-3):
        bad += 1
print("BAD" if bad else "OK", bad)
sys.exit(1 if bad else 0)
