"""Model load time (is a device-layout cache next to the model dir worth it?): wall time of
B200Model(model_dir) = parse 4 .onnx tables + read .bin files + H2D + weight-norm fold / repack / tcgen05 packing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model
from genie_tts.weights import read_model_dir
for ver, seed in (("v2", 0), ("v2ProPlus", 1)):
    d = fixture_dir(ver, seed)
    B200Model(d).close()                       # CUDA context, page cache
    t0 = time.perf_counter(); read_model_dir(d); t1 = time.perf_counter()
    ts = []
    for _ in range(3):
        t = time.perf_counter(); m = B200Model(d); ts.append(time.perf_counter() - t); info = m.info(); m.close()
    print(f"{ver}: host parse {1e3 * (t1 - t0):.0f} ms; B200Model() {1e3 * min(ts):.0f} ms "
          f"(weights {info['weight_bytes'] / 1e6:.0f} MB on device)")
