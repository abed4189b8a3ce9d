import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genie-tts_b200")]
from genie_tts import _native as N
L = N.lib()
cases = [  # M, Cin, Cout, ntaps, dil
    (300, 64, 64, 1, 1), (300, 64, 32, 1, 1), (300, 128, 128, 3, 1), (1000, 256, 256, 3, 3), (777, 192, 384, 5, 1),
    (500, 512, 1536, 1, 1), (2000, 32, 32, 11, 5), (2000, 16, 16, 7, 3), (260, 2048, 512, 1, 1), (100, 512, 2048, 1, 1),
    (640, 96, 192, 1, 1), (640, 48, 48, 3, 1), (640, 24, 24, 11, 1), (333, 192, 768, 3, 1), (64, 768, 192, 3, 1),
]
bad = 0
for (M, Cin, Cout, nt, dil) in cases:
    for mode, exact in ((1, 0), (2, 1), (3, 0)):
        e, r = C.c_float(0), C.c_float(0)
        rc = L.genie_debug_tc_selftest(M, Cin, Cout, nt, dil, mode, exact, C.byref(e), C.byref(r))
        tol = {1: 2e-2, 2: 6e-5, 3: 6e-5}[mode]
        flag = "" if (rc == 0 and e.value < tol) else "  <-- FAIL " + (L.genie_last_error() or b"").decode()
        bad += bool(flag)
        print(f"M={M} Cin={Cin} Cout={Cout} taps={nt} dil={dil} mode={mode} exact_w={exact}: err {e.value:.3e} (ref max {r.value:.2f}){flag}")
sys.exit(1 if bad else 0)
