"""One selftest case (for ncu): python tests/gpu_halo_one.py M C taps dil"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genie-tts_b200")]
from genie_tts import _native as N
L = N.lib()
M, c, taps, dil = (int(x) for x in sys.argv[1:5])
e, r = C.c_float(0), C.c_float(0)
rc = L.genie_debug_tc_selftest(M, c, c, taps, dil, 1, 0, C.byref(e), C.byref(r))
print(rc, e.value, r.value)
