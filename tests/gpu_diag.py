"""Stage-by-stage diagnostic of the CUDA path against the CPU port (GPU box).
Prints max-abs errors; exits non-zero when something is badly off."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir  # noqa: E402
from genie_tts.engine import B200Model, SamplingParams  # noqa: E402
from oracle import gsv_port as P  # noqa: E402
from synth import make_prompt_inputs, make_text_inputs, make_zp_noise  # noqa: E402


def err(name, a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    if a.shape != b.shape:
        print(f"  {name}: SHAPE {a.shape} vs {b.shape}")
        return 1e9
    e = np.abs(a - b).max() if a.size else 0.0
    print(f"  {name}: max-abs err {e:.3e} (ref max {np.abs(b).max():.3e})")
    return e


def run(version, fseed, pkw, tkw, steps):
    print(f"=== {version} fixture seed {fseed}")
    d = fixture_dir(version, fseed)
    t = time.time()
    m = B200Model(d)
    print("  load %.2fs" % (time.time() - t), m.info())
    pm = P.PortModel(d)
    pr = make_prompt_inputs(**pkw)
    tx = make_text_inputs(**tkw)
    prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"], pr.get("sv_emb"))
    codes, ge, gea = prompt.read()
    ref_codes = P.vq_prompts(pm, torch.as_tensor(pr["ssl_content"])).numpy()
    print("  prompt codes equal:", np.array_equal(codes, ref_codes), (codes != ref_codes).sum(), "/", len(codes))
    if pm.is_v2pp:
        rge, rgea = P.prompt_encoder_v2pp(pm, pr["ref_audio"], pr["sv_emb"])
        err("ge", ge, rge.numpy())
        err("ge_advanced", gea, rgea.numpy())
    else:
        rge = P.ref_enc_v2(pm, pr["ref_audio"])
        rgea = None
        err("ge", ge, rge.numpy())
    # T2S
    m.keep(True)
    m.record_logits(True)
    sp = SamplingParams(greedy=True, max_steps=steps)
    ys, idxs = m.t2s_generate([prompt], [tx["text_seq"]], [tx["text_bert"]], sp)
    lg = m.read_logits().reshape(-1, 1025)
    r = P.t2s_generate(pm, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                       max_steps=steps, keep_logits=True)
    x_ref = P.t2s_encode_text(pm, torch.as_tensor(pr["ref_seq"]).reshape(-1), torch.as_tensor(tx["text_seq"]).reshape(-1),
                              torch.as_tensor(pr["ref_bert"]), torch.as_tensor(tx["text_bert"])).numpy()
    err("x", m.read_kept("x"), x_ref)
    n = min(len(lg), len(r.logits))
    for i in sorted(set([0, 1, n - 1])):
        err(f"logits[{i}]", lg[i], r.logits[i])
    print("  y_full equal:", np.array_equal(ys[0], r.y_full[0]), "idx", idxs[0], r.idx)
    if not np.array_equal(ys[0], r.y_full[0]):
        print("   gpu :", ys[0][-steps - 2:].tolist())
        print("   port:", r.y_full[0][-steps - 2:].tolist())
    m.record_logits(False)
    ys2, _ = m.t2s_generate([prompt], [tx["text_seq"]], [tx["text_bert"]], sp)
    print("  graph-replay path equals eager path:", np.array_equal(ys2[0], ys[0]), m.last_timing())
    # VITS
    sem = r.y_full[0][-(r.idx + 1):-1] if r.idx > 0 else r.y_full[0][:-1]
    sem = sem[sem < 1024][:steps]
    zp = make_zp_noise(7, len(sem))
    audio = m.vits_decode([prompt], [tx["text_seq"]], [sem], [zp])[0]
    col = {}
    ref_audio = P.vits_decode(pm, tx["text_seq"], sem, rge, rgea, zp_noise=torch.as_tensor(zp), collect=col)
    for k in ("z", "g_pre", "g_s0", "g_s1", "g_s2", "g_s3", "g_s4"):
        err(k, m.read_kept(k), col[k].numpy())
    # intermediate
    mp, logs = P.enc_p(pm, torch.as_tensor(sem), torch.as_tensor(tx["text_seq"]).reshape(-1),
                       (rgea if rgea is not None else rge).reshape(-1, 1))
    st = m.read_kept("stats").reshape(-1, 384)
    err("m_p", st[:, :192], mp.t().numpy())
    err("logs_p", st[:, 192:], logs.t().numpy())
    e = err("audio", audio, ref_audio)
    snr = 10 * np.log10((ref_audio.astype(np.float64) ** 2).sum() / max(((audio - ref_audio).astype(np.float64) ** 2).sum(), 1e-30))
    print(f"  audio SNR {snr:.1f} dB, amp {np.abs(ref_audio).max():.3f}")
    m.keep(False)
    prompt.close()
    m.close()
    return e


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() // 2))
    bad = 0
    bad += run("v2", 0, dict(seed=11, Lr=20, Ts=60, n_audio=64000, bert=True), dict(seed=12, Lt=15, bert=True), 12) > 2e-3
    if "--all" in sys.argv:
        bad += run("v2ProPlus", 1, dict(seed=31, Lr=20, Ts=60, n_audio=64000, v2pp=True), dict(seed=32, Lt=15), 12) > 2e-3
    sys.exit(1 if bad else 0)
