import os, sys, ctypes as C
os.environ["GENIE_SELFTEST_TIME"] = "1"
lib = C.CDLL(sys.argv[1])
lib.genie_last_error.restype = C.c_char_p
cases = [(24200, 512, 1536, 1, 1, 2, 1), (24200, 2048, 512, 1, 1, 2, 1), (180000, 256, 256, 3, 1, 1, 0),
         (1440000, 128, 128, 7, 3, 1, 0), (2880000, 64, 64, 7, 1, 1, 0), (5760000, 32, 32, 11, 1, 1, 0),
         (11520000, 16, 16, 7, 1, 1, 0)]
for (M, Cin, Cout, nt, dil, mode, exact) in cases:
    e, r = C.c_float(0), C.c_float(0)
    print(f"M={M} Cin={Cin} Cout={Cout} taps={nt} mode={mode}", flush=True)
    lib.genie_debug_tc_selftest(M, Cin, Cout, nt, dil, mode, exact, C.byref(e), C.byref(r))
