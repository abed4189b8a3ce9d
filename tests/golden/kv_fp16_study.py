"""fp16-KV what-if on the CPU port: exact-match % of greedy tokens vs the fp32-cache port, on ja100 / sharp100."""
import sys, os, json
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/genie-tts_b200', '/root/repo/tests/golden']
import numpy as np, torch
from concurrent.futures import ProcessPoolExecutor
from make_acceptance_golden import case_inputs

def one(a):
    case, i, mode = a
    torch.set_num_threads(1)
    from conftest import fixture_dir
    from oracle import gsv_port as P
    ver, fs, items, steps = case_inputs(case)
    pm = P.PortModel(fixture_dir(ver, fs))
    pr, tx = items[i]
    r = P.t2s_generate(pm, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"], max_steps=steps, kv_fp16=mode)
    return i, r.y_full[0], r.idx

if __name__ == "__main__":
    out = {}
    for case in ("ja100", "sharp100"):
        g = np.load(f"/root/repo/tests/golden/acceptance_{case}.npz")
        for mode in (None, "v", "kv"):
            with ProcessPoolExecutor(8) as ex:
                res = sorted(ex.map(one, [(case, i, mode) for i in range(100)]))
            exact, div = 0, []
            for i, y, idx in res:
                ref = g["y_full"][i, :g["y_len"][i]]
                if len(ref) == len(y) and np.array_equal(ref, y) and idx == g["idx"][i]:
                    exact += 1
                else:
                    m = min(len(ref), len(y)); neq = np.nonzero(ref[:m] != y[:m])[0]
                    Ly = int(g["y_len"][i] - g["idx"][i] - 2)
                    t = (int(neq[0]) if len(neq) else m - 1) - Ly
                    div.append((i, t, float(g["gap"][i, t]) if len(neq) else float(g["stop_margin"][i, t])))
            out[f"{case}/{mode}"] = {"exact": exact, "divergences": div}
            print(case, mode, exact, div, flush=True)
    json.dump(out, open("/tmp/kv16_study.json", "w"), indent=1)
