import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genie-tts_b200")]
os.environ["GENIE_SELFTEST_TIME"] = "1"
from genie_tts import _native as N
L = N.lib()
cases = [(24200, 512, 1536, 1, 1), (24200, 2048, 512, 1, 1), (100, 512, 1536, 1, 1),
         (180000, 256, 256, 3, 1), (180000, 256, 256, 11, 5), (1440000, 128, 128, 7, 3), (2880000, 64, 64, 7, 1),
         (5760000, 32, 32, 11, 1), (11520000, 16, 16, 7, 1)]
for (M, Cin, Cout, nt, dil) in cases:
    for mode, exact in ((1, 0), (2, 1)):
        e, r = C.c_float(0), C.c_float(0)
        print(f"M={M} Cin={Cin} Cout={Cout} taps={nt} dil={dil} mode={mode}", flush=True)
        rc = L.genie_debug_tc_selftest(M, Cin, Cout, nt, dil, mode, exact, C.byref(e), C.byref(r))
        print(f"   err {e.value:.3e} rc={rc}", flush=True)
