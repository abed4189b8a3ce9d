"""CPU-only suite: oracle vs golden vectors, model-dir format, host logic, C-ABI exports."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from synth import make_prompt_inputs, make_text_inputs, make_zp_noise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "genie-tts_b200")


# ---------------------------------------------------------------- model-dir format
def test_fixture_layout_matches_converter_sizes(v2_dir, v2pp_dir):
    sz = lambda d, f: os.path.getsize(os.path.join(d, f))  # noqa: E731
    assert sz(v2_dir, "t2s_shared_fp16.bin") == 153413634
    assert sz(v2_dir, "vits_fp16.bin") == 80843520
    assert sz(v2_dir, "t2s_encoder_fp32.bin") == 11465732
    assert sz(v2pp_dir, "vits_fp16.bin") == 124345856
    assert sz(v2pp_dir, "prompt_encoder_fp16.bin") == 44262912


def test_weight_tables(v2_dir, v2pp_dir):
    from genie_tts.weights import read_model_dir
    t = read_model_dir(v2_dir)
    assert not t.is_v2pp and len(t.t2s.tensors) == 291 and len(t.vits.tensors) == 668 and len(t.encoder.tensors) == 7
    w = t.t2s["transformer_encoder.layers.3.self_attn.in_proj_weight"]
    assert w.dtype == np.float16 and w.shape == (1536, 512)
    assert t.encoder["encoder.bert_proj.weight"].dtype == np.float32
    t2 = read_model_dir(v2pp_dir)
    assert t2.is_v2pp and len(t2.vits.tensors) == 650 and len(t2.prompt_encoder.tensors) == 23
    with pytest.raises(FileNotFoundError):
        read_model_dir(os.path.join(v2_dir, "missing"))


def test_weights_only_onnx_roundtrip(tmp_path):
    from genie_tts.onnx_reader import load_model, write_weights_only_model
    p = str(tmp_path / "w.onnx")
    rows = [("a.weight", [3, 4], 1, "x.bin", 0, 48), ("b.bias", [5], 1, "x.bin", 48, 20)]
    write_weights_only_model(p, rows)
    m = load_model(p)
    got = [(t.name, list(t.dims), t.data_type, t.external["location"], int(t.external["offset"]),
            int(t.external["length"])) for t in m.graph.initializers]
    assert got == [tuple(r) for r in rows] and all(t.is_external for t in m.graph.initializers)


def test_reader_on_reference_graph_templates():
    from fixture_models import have_templates, template_path
    if not have_templates("v2"):
        pytest.skip("reference graph templates not staged (oracle/_ref/graphs)")
    from genie_tts.onnx_reader import load_model
    m = load_model(template_path("v2", "t2s_stage_decoder_fp32"))
    assert len(m.graph.nodes) == 1822 and len(m.graph.initializers) == 291 and m.opset == 20
    assert [i.name for i in m.graph.inputs][:3] == ["iy", "iy_emb", "past_k_layer_0"]


# ---------------------------------------------------------------- oracle pinned on golden vectors
@pytest.mark.parametrize("case,ver", [("v2_small", "v2"), ("v2pp_small", "v2pp")])
def test_port_matches_golden(case, ver, v2_dir, v2pp_dir):
    """oracle/gsv_port.py (the travelling oracle) vs vectors produced by the reference's graph files."""
    from oracle import gsv_port as P
    from test_gpu_parity import CASES
    d = v2pp_dir if ver == "v2pp" else v2_dir
    g = np.load(os.path.join(GOLD, case + ".npz"))
    pkw, tkw, steps = CASES[case]
    pr, tx = make_prompt_inputs(**pkw), make_text_inputs(**tkw)
    pm = P.PortModel(d)
    r = P.t2s_generate(pm, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                       max_steps=steps, keep_logits=True)
    assert np.array_equal(r.y_full[0], g["y_full"]) and r.idx == int(g["idx"])
    assert np.array_equal(r.tokens, g["tokens"])
    assert np.abs(r.logits[0] - g["logits_first"]).max() < 1e-4
    assert np.abs(r.logits[-1] - g["logits_last"]).max() < 1e-4
    zp = make_zp_noise(pkw["seed"] + 100, steps + 2)
    if pm.is_v2pp:
        ge, gea = P.prompt_encoder_v2pp(pm, pr["ref_audio"], pr["sv_emb"])
        assert np.abs(ge.numpy().reshape(-1) - g["ge"]).max() < 1e-4
    else:
        ge, gea = P.ref_enc_v2(pm, pr["ref_audio"]), None
    audio = P.vits_decode(pm, tx["text_seq"], g["semantic"], ge, gea, zp_noise=torch.as_tensor(zp))
    assert np.abs(audio - g["audio"]).max() < 1e-4


def test_interpreter_reproduces_golden(v2_dir):
    """The graph interpreter on the reference's own graph files regenerates the committed vectors."""
    from fixture_models import have_templates
    if not have_templates("v2"):
        pytest.skip("reference graph templates not staged (oracle/_ref/graphs)")
    from oracle import ref_pipeline as R
    from test_gpu_parity import CASES
    pkw, tkw, steps = CASES["v2_small"]
    g = np.load(os.path.join(GOLD, "v2_small.npz"))
    pr, tx = make_prompt_inputs(**pkw), make_text_inputs(**tkw)
    s = R.load_sessions(v2_dir)
    R.set_sampler_mode(s, greedy=True, zp_noise=make_zp_noise(pkw["seed"] + 100, steps + 2))
    toks = R.t2s_cpu(s, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                     max_steps=steps)
    assert np.array_equal(toks, g["tokens"])
    audio = R.vocode(s, tx["text_seq"], R.strip_eos(toks), ref_audio_32k=pr["ref_audio"])
    assert np.abs(audio - g["audio"]).max() < 1e-5


def test_sampler_semantics_of_port():
    """stage#[1775-1821]: penalty once per distinct token from the raw logit; ties kept by top-k."""
    from oracle.gsv_port import sample_token
    lg = torch.zeros(1025)
    lg[5], lg[7], lg[9] = 2.0, 2.0 * 1.35, -1.0
    tok, stop = sample_token(lg.clone(), torch.tensor([7, 7, 7]))        # 7 penalised once -> 2.0, tie with 5
    assert tok == 5 and not stop
    lg2 = torch.zeros(1025); lg2[1024] = 9.0
    assert sample_token(lg2, torch.tensor([0]))[1] is True


# ---------------------------------------------------------------- host logic
def test_loop_quirks_and_eos_strip():
    from genie_tts.Core.Inference import finish_t2s, strip_eos
    y = np.arange(100, 120)
    assert finish_t2s(y, 0).shape == (1, 1, 20) and finish_t2s(y, 0)[0, 0, -1] == 0       # idx 0 -> whole y
    assert finish_t2s(y, 5).tolist() == [[[115, 116, 117, 118, 0]]]
    t = np.array([[[3, 4, 1024, 5, 1024]]])
    assert strip_eos(t).tolist() == [[[3, 4]]]
    assert strip_eos(np.array([[[3, 4]]])).tolist() == [[[3, 4]]]


def test_utils_and_language():
    from genie_tts.Utils.Language import normalize_language
    from genie_tts.Utils.Utils import LRUCacheDict
    assert normalize_language("JA") == "Japanese" and normalize_language("zh-CN") == "Chinese"
    assert normalize_language("klingon") == "klingon"
    c = LRUCacheDict(2)
    c["a"], c["b"] = 1, 2
    _ = c["a"]
    c["c"] = 3
    assert list(c.keys()) == ["a", "c"]


def test_text_splitter():
    from genie_tts.Utils.TextSplitter import TextSplitter
    s = TextSplitter()
    assert s.split("") == []
    assert s.split("こんにちは。今日はいい天気ですね、散歩に行きましょう！はい") == \
        ["こんにちは。", "今日はいい天気ですね、散歩に行きましょう！", "はい"]
    assert s.split("a.b.c") == ["a.b.c"]                    # below min_len: merged


def test_public_api_surface_and_errors(tmp_path):
    import genie_tts as genie
    for name in ["load_character", "unload_character", "set_reference_audio", "tts_async", "tts", "stop",
                 "convert_to_onnx", "clear_reference_audio_cache", "start_server", "wait_for_playback_done",
                 "load_predefined_character", "download_genie_data"]:           # reference __init__.py:16-29
        assert callable(getattr(genie, name))
    assert callable(genie.start_server_per_gpu)                                  # extension: one server process per GPU
    from genie_tts import _native as N
    if N.lib().genie_device_count() == 0:
        with pytest.raises(Exception):
            genie.start_server_per_gpu(block=False)                              # fails loudly without a GPU
    with pytest.raises(FileNotFoundError):
        genie.load_character("x", str(tmp_path / "nope"), "ja")
    (tmp_path / "empty").mkdir()
    with pytest.raises(FileNotFoundError):
        genie.load_character("x", str(tmp_path / "empty"), "ja")
    with pytest.raises(ValueError):
        genie.set_reference_audio("nobody", "a.wav", "text")      # no language, unknown character


def test_product_fails_loudly_without_gpu(v2_dir):
    from genie_tts import _native as N
    if N.lib().genie_device_count() > 0:
        pytest.skip("GPU present")
    from genie_tts.engine import B200Model
    with pytest.raises(N.GenieNativeError):
        B200Model(v2_dir)
    import genie_tts as genie
    with pytest.raises(ValueError):
        genie.load_character("c", v2_dir, "xx")
    from genie_tts.ModelManager import model_manager
    assert model_manager.load_character("c", v2_dir, "Japanese") is False      # logged, not raised (reference :304-309)


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "gsv_port" in src or "onnx_interp" in src:
                    bad.append(f)
    assert not bad, bad


# ---------------------------------------------------------------- C-ABI
def test_cabi_exports_every_declared_symbol():
    from genie_tts import _native as N
    hdr = open(os.path.join(ROOT, "include", "genie_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(genie_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert lib.genie_version() >= 100


# ---------------------------------------------------------------- multi-process (gloo, world_size 2)
def test_gloo_world_size_2_bench_aggregation(tmp_path):
    """The N > 1 leg of bench.py on CPU: every rank draws the SAME workload (yield-independent scaling), times its
    own replica, and rank 0 reports max-over-ranks time and summed units (bench.aggregate_over_ranks)."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path[:0] = [{ROOT!r}, {PKG!r}, {os.path.join(ROOT, 'tests')!r}]\n"
        "import numpy as np\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "cfg = bench.CONFIGS[2]\n"
        "pr, texts, berts = bench.make_workload(cfg, 6, rank=r)\n"
        "sig = torch.tensor([float(sum(int(t['text_seq'].sum()) for t in texts))])\n"
        "lo, hi = sig.clone(), sig.clone()\n"
        "dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)\n"
        "assert lo.item() == hi.item(), 'every rank must synthesise the same sentences'\n"
        "dt, units = bench.aggregate_over_ranks([0.5 * (r + 1), 0.25], [10.0, 3.0 + r], w, torch.device('cpu'))\n"
        "assert dt == [0.5 * w, 0.25] and units == [10.0 * w, sum(3.0 + k for k in range(w))], (dt, units)\n"
        "dist.barrier(); print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sk:                       # a free port: a fixed one can still be in TIME_WAIT
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ---------------------------------------------------------------- dynamic batching front
class _FakeSynth:
    """Stands in for GENIE.tts_batch: tags each waveform with its phoneme count, records batch sizes."""

    def __init__(self, delay=0.02, fail_on=None):
        self.batches, self.delay, self.fail_on = [], delay, fail_on

    def tts_batch(self, model, prompts, text_seqs, text_berts=None, sampling=None, zp_noise=None):
        import time
        time.sleep(self.delay)
        self.batches.append((len(prompts), None if text_berts is None else [b.shape for b in text_berts]))
        if self.fail_on is not None and any(len(s) == self.fail_on for s in text_seqs):
            raise ValueError("bad request in batch")
        return [np.full(3, float(len(s)) + 1000.0 * p, np.float32) for p, s in zip(prompts, text_seqs)]


def test_batch_scheduler_batches_orders_and_survives_errors():
    from concurrent.futures import wait
    from genie_tts.Scheduler import BatchScheduler, ReplicaPool
    synth = _FakeSynth()
    sch = BatchScheduler(model=None, synthesizer=synth, max_batch=16, max_wait_ms=30.0)
    futs = [sch.submit(i % 3, np.arange(5 + i), np.ones((5 + i, 1024), np.float32) if i == 7 else None) for i in range(40)]
    wait(futs, timeout=30)
    for i, f in enumerate(futs):                               # every request gets ITS result, whatever the batching
        assert f.result()[0] == 5 + i + 1000.0 * (i % 3)
    st = sch.stats.summary()
    assert st["requests"] == 40 and st["batches"] < 40 and st["max_batch"] <= 16 and st["latency_ms_p50"] > 0
    assert sum(n for n, _ in synth.batches) == 40
    # a batch containing BERT rows passes zero rows for the others (reference: zeros == no features)
    shapes = [b for _, b in synth.batches if b is not None]
    assert shapes and all(s[1] == 1024 for batch in shapes for s in batch)
    # an exception fails only that batch; the worker keeps serving
    synth.fail_on = 9
    bad = sch.submit(0, np.arange(9))
    with pytest.raises(ValueError):
        bad.result(timeout=10)
    synth.fail_on = None
    assert sch.submit(0, np.arange(4)).result(timeout=10)[0] == 4.0
    with pytest.raises(ValueError):
        sch.submit(0, np.zeros(0))
    assert sch.load == 0
    sch.close()
    with pytest.raises(RuntimeError):
        sch.submit(0, np.arange(3))
    # replica pool: least-loaded dispatch, one prompt handle per replica
    s0, s1 = (BatchScheduler(None, _FakeSynth(delay=0.05), max_batch=4, max_wait_ms=1.0, name=f"r{i}") for i in range(2))
    pool = ReplicaPool([s0, s1])
    fs = [pool.submit([0, 1], np.arange(10)) for _ in range(16)]
    wait(fs, timeout=30)
    used = {int(f.result()[0]) // 1000 for f in fs}
    assert used == {0, 1}
    with pytest.raises(ValueError):
        pool.submit([0], np.arange(3))
    pool.close()


# ---------------------------------------------------------------- continuous batching (slot pool) host logic
class _FakePoolModel:
    """Host-side stand-in for the C-ABI slot pool (genie_t2s_pool_*): a request decodes for len(text) % 7 + 3 steps;
    records how many slots were busy at each admission, so the test can see requests JOIN a pool that is decoding."""

    def __init__(self, fail_len=None):
        self.state, self.left, self.meta = [], [], []
        self.admit_busy, self.vits_batches, self.fail_len = [], [], fail_len
        self.lock = None

    def pool_create(self, n_slots, kv_capacity, max_prompt_tokens, max_steps=500):
        self.state, self.left, self.meta = [0] * n_slots, [0] * n_slots, [None] * n_slots

    def pool_admit(self, slots, prompts, text_seqs, text_berts=None, samplings=None):
        if self.fail_len is not None and any(len(s) == self.fail_len for s in text_seqs):
            raise ValueError("bad request in admission")
        self.admit_busy.append(sum(1 for s in self.state if s == 1))
        for sl, p, seq in zip(slots, prompts, text_seqs):
            assert self.state[sl] == 0
            self.state[sl], self.left[sl], self.meta[sl] = 1, len(seq) % 7 + 3, (p, len(seq))

    def pool_step(self, n):
        import time
        time.sleep(0.002)
        for i, s in enumerate(self.state):
            if s == 1:
                self.left[i] -= n
                if self.left[i] <= 0:
                    self.state[i] = 2
        return sum(1 for s in self.state if s == 1)

    def pool_poll(self):
        return np.asarray(self.state), np.zeros(len(self.state), np.int32)

    def pool_read(self, slot):
        p, n = self.meta[slot]
        return np.asarray([7, 8] + [n % 1000] * 6, np.int64), 3       # 2 prompt tokens + 6 generated, idx 3

    def pool_release(self, slot):
        self.state[slot] = 0

    def vits_decode(self, prompts, text_seqs, sems, *a, **k):
        self.vits_batches.append(len(prompts))
        return [np.full(2, float(len(s)) + 1000.0 * p.tag, np.float32) for p, s in zip(prompts, text_seqs)]


def test_continuous_batcher_admits_into_running_pool_and_isolates_requests():
    import time
    from concurrent.futures import wait
    from types import SimpleNamespace
    from genie_tts.Scheduler import ContinuousBatcher, ReplicaPool
    m = _FakePoolModel()
    prompts = [SimpleNamespace(tag=t, ref_len=10, n_prompt_tokens=20) for t in range(3)]
    cb = ContinuousBatcher(m, n_slots=8, kv_capacity=600, max_prompt_tokens=64, max_steps=500, steps_per_tick=1,
                           max_admit=4, vits_max_batch=4, vits_window_ms=1.0)
    futs = []
    for i in range(40):                                         # trickle in: later requests arrive mid-decode
        futs.append(cb.submit(prompts[i % 3], np.arange(5 + i) % 700))
        if i % 5 == 4:
            time.sleep(0.004)
    wait(futs, timeout=30)
    for i, f in enumerate(futs):                                # every request gets ITS result, whatever shares the pool
        assert f.result()[0] == 5 + i + 1000.0 * (i % 3)
    assert any(b > 0 for b in m.admit_busy), "no admission happened while other slots were decoding"
    assert max(m.vits_batches) <= 4 and sum(m.vits_batches) == 40
    assert cb.load == 0 and all(s == 0 for s in m.state)
    # oversize / invalid requests fail alone
    with pytest.raises(ValueError):
        cb.submit(prompts[0], np.arange(700) % 700).result(timeout=10)         # needs more KV rows than a slot holds
    with pytest.raises(ValueError):
        cb.submit(prompts[0], np.asarray([3, 900])).result(timeout=10)         # phoneme id out of range
    # a failing admission fails only its own requests; the pool keeps serving
    m.fail_len = 9
    bad = cb.submit(prompts[0], np.arange(9))
    with pytest.raises(ValueError):
        bad.result(timeout=10)
    m.fail_len = None
    assert cb.submit(prompts[1], np.arange(4)).result(timeout=10)[0] == 1004.0
    # cancel_all drops what is queued / decoding (reference /stop)
    slow = [cb.submit(prompts[0], np.arange(6)) for _ in range(6)]
    cb.cancel_all()
    for f in slow:
        with pytest.raises(RuntimeError):
            f.result(timeout=10)
    assert cb.submit(prompts[2], np.arange(3)).result(timeout=10)[0] == 2003.0
    assert cb.load == 0
    # replica pool over two continuous batchers
    cb2 = ContinuousBatcher(_FakePoolModel(), n_slots=4, kv_capacity=600, max_prompt_tokens=64, steps_per_tick=1,
                            vits_window_ms=1.0, name="r1")
    pool = ReplicaPool([cb, cb2])
    fs = [pool.submit([prompts[0], prompts[1]], np.arange(10)) for _ in range(24)]
    wait(fs, timeout=30)
    assert {int(f.result()[0]) // 1000 for f in fs} == {0, 1}
    pool.close()
    with pytest.raises(RuntimeError):
        cb.submit(prompts[0], np.arange(3))


# ---------------------------------------------------------------- bench.py contract (reference arm runs on CPU)
def test_bench_reference_arm_json_contract():
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio-sec/sec" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    assert "workload" in line["config"] and line["config"]["bench_config"] == 2      # same object as the b200 arm's
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
