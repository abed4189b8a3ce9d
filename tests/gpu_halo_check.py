"""Debug: tcgen05 halo-conv selftest.  argv: case indices (default all).  GENIE_TC_HALO = debug flags."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genie-tts_b200")]
if not os.environ.get("NO_TIME"):
    os.environ["GENIE_SELFTEST_TIME"] = "1"
from genie_tts import _native as N
L = N.lib()
cases = [(5000, 64, 64, 3, 1), (5000, 64, 64, 3, 8), (5000, 64, 64, 7, 3), (180000, 256, 256, 3, 1), (180000, 256, 256, 11, 5),
         (1440000, 128, 128, 7, 3), (2880000, 64, 64, 7, 1), (5760000, 32, 32, 11, 1), (5760000, 32, 32, 3, 5),
         (11520000, 16, 16, 7, 1), (11520000, 16, 16, 11, 5), (3000, 192, 512, 7, 1)]
sel = [int(a) for a in sys.argv[1:]] or range(len(cases))
for i in sel:
    (M, Cin, Cout, nt, dil) = cases[i]
    e, r = C.c_float(0), C.c_float(0)
    print(f"M={M} Cin={Cin} Cout={Cout} taps={nt} dil={dil}", flush=True)
    rc = L.genie_debug_tc_selftest(M, Cin, Cout, nt, dil, 1, 0, C.byref(e), C.byref(r))
    print(f"   err {e.value:.3e} ref_max {r.value:.3e} rc={rc}", flush=True)
