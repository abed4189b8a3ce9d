#!/usr/bin/env python
"""Benchmark of the GPT-SoVITS synthesis hot path (BASELINE.json metric: audio-sec/sec).

Workload (BASELINE.json configs[1], README latency-set shape): GPT-SoVITS V2,
random-init weights in the converter's on-disk layout, 100 Japanese-shaped
sentences per GPU (ref 60 phones, target 40-60 phones, 264 HuBERT frames -> 132
prompt tokens, fixed budget of 90 semantic tokens = 3.6 s audio each), synthetic
inputs at the G2P/HuBERT boundary.  One step = one pass of the
whole path (T2S encode + prefill + 90 decode steps with Philox sampling + SoVITS
decode) over the batch.

  value : whole-job audio-sec/sec with inputs resident in HBM (C-ABI io_on_device=1)
  e2e   : same metric through the reference-facing call GENIE.tts_batch with HOST
          numpy buffers (H2D of phoneme ids, D2H of tokens and float32 audio inside
          the timed region)
  --config {2,3,4} : BASELINE.json configs[1] (default, the headline), configs[2] (V2ProPlus English paragraph,
          batch 64, KV up to ~1440) and configs[3] (V2 with 1024-d BERT rows, batch 256 per GPU)
  --impl reference : the reference's own CPU path (its ONNX graph files executed
          by oracle/onnx_interp.py over the restated host loop; falls back to the
          torch port if the graphs were not staged) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "genie-tts_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "audio-sec/sec"
# BASELINE.json configs[1..3] (0-based: configs[0] is the reference's own batch-1 CPU case -> --impl reference).
# Synthetic inputs at the G2P / HuBERT / SV / RoBERTa output boundary, SURVEY.md §8d.
CONFIGS = {
    2: dict(version="v2", fixture_seed=0, sentences=100, tokens=90, Lr=60, Ts=264, n_audio=169600, Lt=(40, 60),
            bert=False,
            workload="configs[1]: GPT-SoVITS V2, 100 sentences x ~20 chars (Lr=60, Lt~U{40..60}, 132 prompt tokens, "
                     "90-token budget, KV 242->332), 1 B200"),
    3: dict(version="v2ProPlus", fixture_seed=1, sentences=64, tokens=500, Lr=300, Ts=1000, n_audio=640000,
            Lt=(100, 140), bert=False,
            workload="configs[2]: GPT-SoVITS V2ProPlus English paragraph, sentence-split: batch 64 x (Lt~U{100..140}, "
                     "20 s reference = 300 ref phones + 500 prompt tokens, 500-step budget = the reference's loop "
                     "bound, KV 900->1441 per utterance)"),
    4: dict(version="v2", fixture_seed=0, sentences=256, tokens=90, Lr=60, Ts=264, n_audio=169600, Lt=(40, 60),
            bert=True,
            workload="configs[3]: GPT-SoVITS V2 Chinese shape: 256 sentences per GPU with N(0,1) 1024-d BERT rows for "
                     "reference and target text (bert_proj GEMM in the encoder), 90-token budget"),
}
DTYPE = ("f32 accumulate/activations; T2S GEMMs fp16 hi+lo split operands (tcgen05), KV cache rows fp16 (q/scores/"
         "accumulators f32); SoVITS convs single-pass fp16 operands (tcgen05)")
# algorithmic work (measured on the reference graphs)
GEN_GFLOP_PER_AUDIO_S = {"v2": 8.26 + 16.52 + 8.26 + 4.13 + 2.06 + 1.36,   # HiFi-GAN stages 0-4 + ups
                         "v2ProPlus": 89.0}
# dram__bytes_read.sum + dram__bytes_write.sum of one decode_attention launch from the committed `ncu --set full`
# capture of the CURRENT kernel (profiles/), with the algorithmic bytes of that same launch
ATT_NCU = {"file": "profiles/r02/ncu_full_decode_attention16_bulk.txt (one branch launch of the two-branch step: 50 "
                   "utterances, kv_len ~245, fp16 rows)", "dram_bytes": 25341184, "algorithmic_bytes": 25.0e6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def make_workload(cfg, n_sent=None, rank=0, seed0=1234):
    """(prompt inputs, [text inputs]) of a config.  Every rank draws the SAME sentences (``rank`` is deliberately
    unused): produced audio per step is then identical on every GPU and the 1/2/4/8 scaling curve measures the
    machine, not the per-rank EOS-strip yield."""
    from synth import make_prompt_inputs, make_text_inputs
    n_sent = n_sent or cfg["sentences"]
    v2pp = cfg["version"] == "v2ProPlus"
    pr = make_prompt_inputs(seed=seed0, Lr=cfg["Lr"], Ts=cfg["Ts"], n_audio=cfg["n_audio"], bert=cfg["bert"], v2pp=v2pp)
    rng = np.random.default_rng(seed0 + 1)
    lo, hi = cfg["Lt"]
    texts = [make_text_inputs(seed=seed0 + 10 + i, Lt=int(rng.integers(lo, hi + 1)), bert=cfg["bert"])
             for i in range(n_sent)]
    berts = [t["text_bert"] for t in texts] if cfg["bert"] else None
    return pr, texts, berts


def aggregate_over_ranks(times, units, world, device):
    """Whole-job figures of an N-rank run: time = MAX over ranks (device-timed spans), units = SUM over ranks."""
    import torch
    import torch.distributed as dist
    tt = torch.tensor(list(times), dtype=torch.float64, device=device)
    uu = torch.tensor(list(units), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    return tt.tolist(), uu.tolist()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_step(cfg, pr, tx, sessions_or_port, kind, tokens):
    """One sentence through the reference CPU path; returns audio seconds produced."""
    import torch
    v2pp = cfg["version"] == "v2ProPlus"
    if kind == "reference":
        from oracle import ref_pipeline as R
        s = sessions_or_port
        toks = R.t2s_cpu(s, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                         max_steps=tokens)
        sem = R.strip_eos(toks)
        sem = np.where(sem >= 1024, 0, sem)
        if v2pp:
            if "ge" not in pr:
                pr["ge"], pr["ge_advanced"] = R.prompt_global_emb(s, pr["ref_audio"], pr["sv_emb"])
            audio = R.vocode(s, tx["text_seq"], sem, ge=pr["ge"], ge_advanced=pr["ge_advanced"])
        else:
            audio = R.vocode(s, tx["text_seq"], sem, ref_audio_32k=pr["ref_audio"])
    else:
        from oracle import gsv_port as P
        pm = sessions_or_port
        r = P.t2s_generate(pm, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                           max_steps=tokens, force_tokens=tokens,
                           noise_fn=lambda i: torch.randn(1025))
        sem = r.tokens.reshape(-1)
        sem = sem[sem < 1024]
        if v2pp:
            ge, gea = P.prompt_encoder_v2pp(pm, pr["ref_audio"], pr["sv_emb"])
        else:
            ge, gea = P.ref_enc_v2(pm, pr["ref_audio"]), None
        audio = P.vits_decode(pm, tx["text_seq"], sem, ge, gea, zp_noise=torch.randn(1, 192, 2 * len(sem)))
    return len(audio) / 32000.0


def load_cpu_reference(cfg, model_dir):
    import torch
    from fixture_models import have_templates
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if have_templates(cfg["version"]) and os.path.getsize(os.path.join(model_dir, "vits_fp32.onnx")) > 100000:
        from oracle import ref_pipeline as R
        s = R.load_sessions(model_dir)
        R.set_sampler_mode(s, greedy=False)
        return s, "reference"
    from oracle import gsv_port as P
    return P.PortModel(model_dir), "port"


def line_config(cfg, args, world):
    """The `config` object of the JSON line — identical for both arms (the driver compares them)."""
    depth = max(1, args.pipeline)
    return {"workload": cfg["workload"], "bench_config": args.config,
            "sentences_per_gpu": args.sentences or cfg["sentences"], "tokens_per_sentence": cfg["tokens"],
            "sampling": "top_k=15 T=1.0 rep=1.35 Philox",
            "l2": "working set (KV cache, vocoder activations) >> 126 MB L2, no explicit flush",
            "parallelism": f"dp{world} by utterance, replicas only; every rank synthesises the same sentences "
                           "(yield-independent scaling)",
            "pipeline": f"{depth} steps in flight per GPU on {depth} execution contexts, stage-aligned: prefills one "
                        "after the other, the decodes of all steps in flight at the same time, then the vocoder passes; "
                        "stage_ms / rooflines / share_of_step are measured on isolated steps"}


def cpu_sample_tokens(cfg):
    """Bounded CPU sample: the config's own token budget, capped so one sentence stays within ~20 s of CPU work."""
    return min(cfg["tokens"], 90)


def run_reference_arm(args, cfg, model_dir, rank):
    import torch
    if rank != 0:
        return
    pr, texts, _ = make_workload(cfg, 4)
    obj, kind = load_cpu_reference(cfg, model_dir)
    cores = torch.get_num_threads()
    tokens = cpu_sample_tokens(cfg)
    sample = (f"1 sentence of the workload per step ({tokens}-token budget, batch 1: the reference path is "
              f"single-stream by construction), {'reference ONNX graphs on the torch-CPU interpreter' if kind == 'reference' else 'torch port'}"
              " (onnxruntime 1.22.1 unavailable offline)")
    for i in range(args.warmup):
        cpu_reference_step(cfg, pr, texts[i % len(texts)], obj, kind, tokens)
    t0 = time.perf_counter()
    audio_s = 0.0
    for i in range(args.steps):
        audio_s += cpu_reference_step(cfg, pr, texts[i % len(texts)], obj, kind, tokens)
    dt = time.perf_counter() - t0
    v = audio_s / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": line_config(cfg, args, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)        # a multiple of the default pipeline depth (whole waves)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configs index + 1: 2 = headline (V2 JA 100 sentences), 3 = V2ProPlus EN "
                         "paragraph batch 64 long KV, 4 = V2 ZH BERT batch 256 per GPU")
    ap.add_argument("--sentences", type=int, default=0, help="override the config's batch size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kv-fp32", action="store_true", help="keep the KV cache rows in fp32 (default: fp16 rows)")
    ap.add_argument("--pipeline", type=int, default=3,
                    help="batches in flight per GPU (execution contexts on the same weights): 1 = one step after "
                         "the other; 3 (default): three steps in flight, their decode stages run at the same time")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = CONFIGS[args.config]
    TOKENS = cfg["tokens"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from conftest import fixture_dir
    # rank 0 writes the fixture once; other ranks wait on the marker
    if local_rank == 0:
        model_dir = fixture_dir(cfg["version"], cfg["fixture_seed"])
    else:
        from conftest import FIXTURE_ROOT
        model_dir = os.path.join(FIXTURE_ROOT, f"{cfg['version']}_seed{cfg['fixture_seed']}")
        while not os.path.exists(os.path.join(model_dir, ".complete")):
            time.sleep(0.5)

    if args.impl == "reference":
        run_reference_arm(args, cfg, model_dir, rank)
        return

    import torch
    import torch.distributed as dist
    from genie_tts import _native as N
    from genie_tts.Core.Inference import GENIE, finish_t2s, strip_eos
    from genie_tts.engine import B200Model, SamplingParams
    N.require_gpu()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = B200Model(model_dir, device=local_rank)
    if args.kv_fp32:
        model.set_option("kv_fp16", 0)
    for opt in ("decode_branches", "decode_split_min", "prefill_single"):        # experiments: GENIE_OPT_decode_branches=3 ...
        if os.environ.get("GENIE_OPT_" + opt):
            model.set_option(opt, int(os.environ["GENIE_OPT_" + opt]))
    pr, texts, berts = make_workload(cfg, args.sentences or cfg["sentences"], rank=rank)
    prompt = model.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"], pr.get("sv_emb"))  # untimed
    B = len(texts)
    prompts = [prompt] * B
    seqs = [t["text_seq"].reshape(-1) for t in texts]
    lens = np.asarray([len(s) for s in seqs], dtype=np.int32)
    sp = SamplingParams(greedy=False, seed=2026, max_steps=TOKENS, fixed_steps=TOKENS)
    genie = GENIE()

    # ---- device-resident leg: `pipeline` execution contexts on the same weights, one set of I/O buffers each
    dev = torch.device("cuda", local_rank)
    depth = max(1, args.pipeline)
    ctxs = model.pipeline_contexts(depth)
    seq_dev = torch.from_numpy(np.concatenate(seqs)).to(dev)
    bert_dev = torch.from_numpy(np.concatenate(berts, axis=0)).to(dev) if berts is not None else None
    y_ld = prompt.n_prompt_tokens + TOKENS + 2
    io = [(torch.zeros((B, y_ld), dtype=torch.int64, device=dev),
           torch.zeros(B * TOKENS * 1280, dtype=torch.float32, device=dev)) for _ in ctxs]
    stage_ms = {"prefill": [], "decode": [], "vits": [], "generator": []}

    dbg = os.environ.get("BENCH_DEBUG") == "1"

    def step_device(k=0, record=True):
        ctx = ctxs[k]
        y_dev, audio_dev = io[k]
        w0 = time.perf_counter()
        y_len, idx = ctx.t2s_generate_device(prompts, seq_dev, lens, sp, y_dev, text_bert_cat=bert_dev)
        w1 = time.perf_counter()
        t = ctx.last_timing()
        # host glue of the reference (Inference.py:41-44,108-109) on the small token matrix
        y = y_dev.cpu().numpy()
        sems = [strip_eos(finish_t2s(y[b, :y_len[b]], int(idx[b]))).reshape(-1) for b in range(B)]
        sems = [s if len(s) else np.zeros(1, np.int64) for s in sems]
        sl = np.asarray([len(s) for s in sems], dtype=np.int32)
        sem_dev = torch.from_numpy(np.concatenate(sems)).to(dev)
        w2 = time.perf_counter()
        alen = ctx.vits_decode_device(prompts, seq_dev, lens, sem_dev, sl, audio_dev, seed=sp.seed)
        w3 = time.perf_counter()
        t2 = ctx.last_timing()
        if dbg and rank == 0:
            print(f"[dbg] ctx{k} t2s call {1e3 * (w1 - w0):.2f} ms (stages {t['t2s_ms']:.2f}), glue {1e3 * (w2 - w1):.2f} ms, "
                  f"vits call {1e3 * (w3 - w2):.2f} ms (stage {t2['vits_ms']:.2f})", file=sys.stderr)
        if record:
            stage_ms["prefill"].append(t["prefill_ms"]); stage_ms["decode"].append(t["decode_ms"])
            stage_ms["vits"].append(t2["vits_ms"]); stage_ms["generator"].append(t2["generator_ms"])
        return float(alen.sum()) / 32000.0, t2

    from genie_tts.Core.Inference import MERGE_WAVE_VOCODER as merge_vocoder

    def vocode_device(k, y_len, idx):
        y_dev, audio_dev = io[k]
        y = y_dev.cpu().numpy()
        sems = [strip_eos(finish_t2s(y[b, :y_len[b]], int(idx[b]))).reshape(-1) for b in range(B)]
        sems = [x if len(x) else np.zeros(1, np.int64) for x in sems]
        sl = np.asarray([len(x) for x in sems], dtype=np.int32)
        sem_dev = torch.from_numpy(np.concatenate(sems)).to(dev)
        alen = ctxs[k].vits_decode_device(prompts, seq_dev, lens, sem_dev, sl, audio_dev, seed=sp.seed)
        return float(alen.sum()) / 32000.0

    seq_dev_w = torch.cat([seq_dev] * depth) if depth > 1 else seq_dev
    audio_dev_w = (torch.zeros(depth * B * TOKENS * 1280, dtype=torch.float32, device=dev)
                   if depth > 1 and merge_vocoder else None)

    def vocode_device_wave(toks):
        """ONE vocoder pass for the whole wave (as GENIE.tts_batch_stream does): utterances of all steps in flight,
        each on the Philox stream of its position in its own batch."""
        w = len(toks)
        sems, ids = [], []
        for k, (y_len, idx) in enumerate(toks):
            y = io[k][0].cpu().numpy()
            ss = [strip_eos(finish_t2s(y[b, :y_len[b]], int(idx[b]))).reshape(-1) for b in range(B)]
            sems += [x if len(x) else np.zeros(1, np.int64) for x in ss]
            ids += list(range(B))
        sl = np.asarray([len(x) for x in sems], dtype=np.int32)
        sem_dev = torch.from_numpy(np.concatenate(sems)).to(dev)
        alen = ctxs[0].vits_decode_device(prompts * w, seq_dev_w[:len(seq_dev) * w], np.tile(lens, w), sem_dev, sl,
                                          audio_dev_w, seed=sp.seed, noise_ids=ids)
        return float(alen.sum()) / 32000.0

    def run_steps_device(n):
        """n steps, `depth` in flight, stage-aligned like GENIE.tts_batch_stream: per wave the prefills one after the
        other, ALL decodes at the same time (one host thread per context), then one vocoder pass for the wave."""
        if depth == 1:
            return sum(step_device(0, record=False)[0] for _ in range(n))
        from concurrent.futures import ThreadPoolExecutor
        total, done = 0.0, 0

        def dec(k):
            ctxs[k].t2s_decode_steps(TOKENS)
            return ctxs[k].t2s_read_device(io[k][0])
        with ThreadPoolExecutor(max_workers=depth) as ex:
            while done < n:
                w = min(depth, n - done)
                for k in range(w):
                    ctxs[k].t2s_prefill_device(prompts, seq_dev, lens, sp, text_bert_cat=bert_dev)
                toks = list(ex.map(dec, range(w)))
                if merge_vocoder:
                    total += vocode_device_wave(toks)
                else:
                    for k in range(w):
                        total += vocode_device(k, *toks[k])
                done += w
        return total

    def step_host(k=0):
        auds = genie.tts_batch(ctxs[k], prompts, seqs, berts, sampling=sp)
        return sum(len(a) for a in auds) / 32000.0, sum(a.nbytes for a in auds)

    # warm-up: every context once (graphs, workspaces), then W isolated steps on context 0 — their stage events are
    # the per-stage figures of the line (under overlap a stage's event span also contains the other batch's kernels)
    for k in range(depth):
        step_device(k, record=False)
    for _ in range(args.warmup):
        step_device(0)
    last_t = None
    t_iso0 = time.perf_counter()
    _, last_t = step_device(0)
    iso_step_ms = 1000 * (time.perf_counter() - t_iso0)
    if depth > 1:
        run_steps_device(depth)         # one untimed wave: the wave-sized vocoder workspace exists before the clock starts
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = N.lib().genie_launch_count()
    N.lib().genie_profiler_range(1)     # no-op unless run under `ncu --profile-from-start off`; outside the timed
    barrier()                           # region: the first cudaProfilerStart of a process costs tens of ms
    t0 = time.perf_counter()
    audio_s = run_steps_device(args.steps)
    barrier()
    dt = time.perf_counter() - t0
    N.lib().genie_profiler_range(0)
    launches = N.lib().genie_launch_count() - launches0
    clk = clocks.stop()

    # ---- e2e leg (host buffers through the reference-facing call, same pipelining)
    for k in range(depth):
        step_host(k)
    for _ in genie.tts_batch_stream(model, ((prompts, seqs, berts) for _ in range(depth)), sampling=sp, depth=depth):
        pass                            # one untimed wave
    barrier()
    t1 = time.perf_counter()
    e_audio, d2h = 0.0, 0
    # the public throughput call: a stream of batches, `depth` in flight (GENIE.tts_batch_stream)
    for auds in genie.tts_batch_stream(model, ((prompts, seqs, berts) for _ in range(args.steps)), sampling=sp, depth=depth):
        e_audio += sum(len(a) for a in auds) / 32000.0
        d2h = sum(a.nbytes for a in auds) + B * y_ld * 8
    barrier()
    dt_e = time.perf_counter() - t1

    # ---- batch-1 first-audio latency (BASELINE.json metric part 3): one sentence of the config, host buffers,
    # text front end excluded (untimed in the reference comparison too), the config's token budget
    lat = []
    for i in range(12):
        torch.cuda.synchronize()
        tl0 = time.perf_counter()
        genie.tts_batch(model, [prompt], [seqs[i % B]], [berts[i % B]] if berts is not None else None, sampling=sp)
        lat.append(1000 * (time.perf_counter() - tl0))
    first_audio_ms = float(np.median(lat[2:]))
    first_audio_p99 = float(np.max(lat[2:]))
    t_b1 = model.last_timing()          # stage events of the last batch-1 call

    (dt, dt_e), (audio_s, e_audio) = aggregate_over_ranks([dt, dt_e], [audio_s, e_audio], world, dev)

    # ---- dominant kernel, timed live: decode_attention sits inside the step's CUDA graph, so the library
    # replays it on the final KV cache (24 layers back to back, GBs >> L2) between CUDA events on its stream
    model.set_option("time_attention", 20)
    model.t2s_generate_device(prompts, seq_dev, lens, sp, io[0][0], text_bert_cat=bert_dev)
    t_att = model.last_timing()
    model.set_option("time_attention", 0)

    if rank == 0:
        hbm_peak, tf_peak, which = peaks()
        sm = {k: float(np.mean(v)) for k, v in stage_ms.items()}
        audio_per_step = audio_s / args.steps / world
        step_ms = 1000 * dt / args.steps
        solo_ms = sm["prefill"] + sm["decode"] + sm["vits"]     # one step alone: what the shares below refer to
        S = np.asarray([cfg["Lr"] + len(q) + prompt.n_prompt_tokens for q in seqs], dtype=np.float64)   # prefill rows
        # (1) decode attention (HBM): algorithmic bytes = K and V rows of every cached token of one layer, fp32.
        # The replay runs on the FINAL cache (S + TOKENS rows per utterance): its bytes and its time belong together.
        att_us, att_mb = t_att["decode_attention_us"], t_att["decode_attention_kv_mb"]
        att_gbs = att_mb * 1e6 / (att_us * 1e-6) / 1e9 if att_us > 0 else 0.0
        # inside the step the cache grows linearly from S to S + TOKENS: the AVERAGE launch streams these bytes
        kvb = model.kv_bytes_per_element
        kv_mb_avg = float((S + TOKENS / 2.0).sum()) * 2 * 512 * kvb / 1e6
        att_us_avg = att_us * kv_mb_avg / att_mb if att_mb > 0 else 0.0     # time is linear in bytes
        att_share = att_us_avg * 1e-3 * 24 * TOKENS / solo_ms
        # (2) generator convs (tensor): GFLOP per audio-second of the graph over the generator stage
        n_gen = max(1, last_t["generator_launches"])
        gen_tf = GEN_GFLOP_PER_AUDIO_S[cfg["version"]] * 1e9 * audio_per_step / (sm["generator"] * 1e-3) / 1e12
        # (3) whole decode step (HBM): fp16 weights + fp32 KV read of the AVERAGE step
        kv_mb_step = kv_mb_avg * 24
        dec_gbs = (152.364 + kv_mb_step) * 1e6 / (sm["decode"] / TOKENS * 1e-3) / 1e9
        # (4) prefill (tensor): 150.99 MFLOP per position + 24*4*S^2*512 attention
        pre_tf = float((S * 150.99e6 + 24 * 4 * S * S * 512).sum()) / (sm["prefill"] * 1e-3) / 1e12
        narrow_gbs = (last_t.get("narrow_conv_mb", 0.0) * 1e6 / (last_t["narrow_conv_ms"] * 1e-3) / 1e9
                      if last_t.get("narrow_conv_ms") else 0.0)
        b1_ms_tok = t_b1["decode_ms"] / max(1, t_b1["steps"])
        b1_T = cfg["Lr"] + len(seqs[11 % B]) + prompt.n_prompt_tokens + TOKENS / 2
        b1_gbs = (152.364e6 + 24 * 2 * 512 * 4 * b1_T) / (b1_ms_tok * 1e-3) / 1e9     # batch <= 4 keeps fp32 rows
        h2d = int(sum(s.nbytes for s in seqs) + lens.nbytes + (sum(b.nbytes for b in berts) if berts is not None else 0))
        line = {
            "metric": METRIC, "value": audio_s / dt, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": line_config(cfg, args, world),
            "isolated_step_ms": iso_step_ms,
            # yield-independent companion of `value`: decode tokens x 40 ms (value counts the audio actually
            # produced, i.e. after the reference's slicing quirks and EOS strip)
            "token_audio_s_per_s": world * B * TOKENS * 0.04 * args.steps / dt,
            "e2e": {"value": e_audio / dt_e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "t2s_tokens_per_s": world * B * TOKENS / (sm["decode"] * 1e-3),        # decode stage of an isolated step
            "stage_ms": sm,
            "first_audio_ms_p50_batch1": first_audio_ms, "first_audio_ms_max_batch1": first_audio_p99,
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "decode_attention_kernel<fused> (one query per utterance x head over "
                         f"the {'fp16' if kvb == 2 else 'fp32'} KV cache; largest single kernel of the step)",
                         "achieved": att_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": att_gbs / hbm_peak,
                         "peak_source": which, "bytes_per_launch": att_mb * 1e6, "avg_launch_us": att_us,
                         "launches_per_step": 24 * TOKENS, "share_of_step": att_share,
                         "bytes_per_launch_step_average": kv_mb_avg * 1e6,
                         "timing": "replayed alone after the timed steps on the FINAL KV cache as ONE launch over all "
                                   "utterances, 20 x 24 layers back to back, CUDA events on the launching stream (inside "
                                   "the step it is a CUDA-graph node, issued as two half-batch launches on the two branch "
                                   "streams); share_of_step scales that time to the step-average KV length",
                         "traffic": ATT_NCU["dram_bytes"],
                         "traffic_note": f"{ATT_NCU['file']}: dram read+write of one launch whose algorithmic bytes "
                                         f"are {ATT_NCU['algorithmic_bytes'] / 1e6:.1f} MB"},
            "rooflines": [
                {"stage": "sovits generator convs (tc_conv_gemm / tc_halo_conv, tcgen05)", "bound": "tensor",
                 "achieved": gen_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": gen_tf / tf_peak,
                 "launches_per_step": n_gen, "avg_launch_ms": sm["generator"] / n_gen,
                 "share_of_step": sm["generator"] / solo_ms},
                # narrow generator stages (<= 32 channels: tc_halo_conv, fused transposed convs, conv_post): every
                # tensor of every conv counted once per read / write, over the stage events of the last timed step
                {"stage": "sovits generator narrow stages (C <= 32)", "bound": "hbm", "achieved": narrow_gbs,
                 "peak": hbm_peak, "unit": "GB/s", "frac": narrow_gbs / hbm_peak,
                 "ms": last_t.get("narrow_conv_ms"), "share_of_step": (last_t.get("narrow_conv_ms") or 0.0) / solo_ms},
                {"stage": "t2s decode step (all kernels, CUDA graph), bytes of the step-average KV length",
                 "bound": "hbm", "achieved": dec_gbs,
                 "peak": hbm_peak, "unit": "GB/s", "frac": dec_gbs / hbm_peak, "share_of_step": sm["decode"] / solo_ms},
                {"stage": "t2s prefill (tc_conv_gemm split-fp16 + attention)", "bound": "tensor", "achieved": pre_tf,
                 "peak": tf_peak, "unit": "TFLOP/s", "frac": pre_tf / tf_peak, "share_of_step": sm["prefill"] / solo_ms},
                # batch 1 (first-audio path): one persistent kernel per token; bytes = fp16 weights + fp32 KV read
                {"stage": "t2s decode batch 1 (t2s_step_persistent_kernel)", "bound": "hbm",
                 "achieved": b1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": b1_gbs / hbm_peak,
                 "ms_per_token": b1_ms_tok, "stage_ms": {"prefill": t_b1["prefill_ms"], "decode": t_b1["decode_ms"],
                                                         "vits": t_b1["vits_ms"]}},
            ],
        }
        if not args.no_cpu_baseline:
            obj, kind = load_cpu_reference(cfg, model_dir)
            tokens = cpu_sample_tokens(cfg)
            tcpu = time.perf_counter()
            a = cpu_reference_step(cfg, pr, texts[0], obj, kind, tokens)
            dcpu = time.perf_counter() - tcpu
            line["cpu_baseline"] = {"value": a / dcpu, "unit": "audio-s/s", "cores": torch.get_num_threads(),
                                    "kind": kind, "sample": f"1 sentence ({tokens} tokens, batch 1), {dcpu:.1f} s of CPU work"}
        print(json.dumps(line))
    prompt.close()
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
