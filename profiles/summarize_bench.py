"""One line per bench JSON file: value / e2e / stage times / dominant-kernel roofline.  usage: summarize_bench.py f.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:                                   # noqa: BLE001
        print(f"{f}: no JSON line ({e})")
        continue
    ro = r.get("roofline", {})
    print(f"{f}: value {r['value']:.1f} e2e {r['e2e']['value']:.1f} {r['unit']} | ms/step {r['ms_per_step']:.1f} "
          f"(isolated {r.get('isolated_step_ms', 0):.1f}) | stages "
          + " ".join(f"{k}={v:.1f}" for k, v in r.get("stage_ms", {}).items())
          + f" | attention {ro.get('avg_launch_us', 0):.2f} us = {ro.get('achieved', 0):.0f} GB/s = {ro.get('frac', 0):.3f} of HBM"
          + f" | first audio {r.get('first_audio_ms_p50_batch1', 0):.1f} ms | clocks {r.get('clocks', {}).get('reasons')}")
