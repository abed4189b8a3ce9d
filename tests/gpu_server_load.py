"""BASELINE config 5: FastAPI server, continuous batching, N concurrent synthetic requests over every GPU of the box.

  python tests/gpu_server_load.py [--clients 512] [--rounds 3] [--gpus 8] [--procs 8] [--out gpurun_out/config5_load.json]

The server (genie_tts.Server: the reference's REST surface on SynthesisService, one ContinuousBatcher per GPU) runs in
its OWN process, started by this script (`--procs P`: P server processes on P ports, each owning gpus / P GPUs, with
one client process each - one Python process serves about 1500 requests/s before its interpreter lock saturates);
the clients are `--clients` concurrent aiohttp coroutines, each sending `--rounds` /tts requests back to back (closed loop).  A request is one ~20-character sentence
(JA20 shape: 40-60 phonemes from a synthetic front end, 90-token budget, Philox sampling with a fresh seed), the
response is the sentence's raw s16 PCM stream.  Reported: first-audio latency (request sent -> first PCM byte; the
reference streams one chunk per sentence, so this is the sentence's whole synthesis) p50 / p99, request latency,
sustained audio-seconds per second, per-replica scheduler counters."""
import argparse
import asyncio
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
PORT = 18777


def serve(n_gpus: int, first_device: int = 0, port: int = PORT) -> None:
    import uvicorn
    from conftest import fixture_dir
    from genie_tts import Internal, Server
    from genie_tts.Audio.ReferenceAudio import ReferenceAudio
    from genie_tts.GetPhonesAndBert import set_text_frontend
    from genie_tts.Service import SynthesisService
    from genie_tts.engine import SamplingParams
    from synth import make_prompt_inputs

    def frontend(text, language):                    # synthetic G2P at the boundary: 40-60 phonemes per sentence
        h = sum(map(ord, text))
        rng = np.random.default_rng(h)
        seq = rng.integers(0, 732, (1, 40 + h % 21)).astype(np.int64)
        seq[0, 0] = 3
        return seq, None

    set_text_frontend(frontend)
    svc = SynthesisService(devices=list(range(first_device, first_device + n_gpus)), n_slots=int(os.getenv("LOAD_SLOTS", "256")), kv_capacity=448, max_prompt_tokens=160,
                           max_steps=90, sampling=SamplingParams(max_steps=90, fixed_steps=90))
    Server.set_service(svc)
    svc.load_character("Mika", fixture_dir("v2", 0), "Japanese")
    pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
    ref = ReferenceAudio.from_features("synthetic-ref", pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
    Internal.set_reference_features("Mika", ref)
    svc.set_reference("Mika", ref)
    st = svc.submit("Mika", "warm up", False)        # builds pools, prompts and graphs on every replica
    list(st.chunks(timeout=300))
    uvicorn.run(Server.app, host="127.0.0.1", port=port, log_level="warning")


async def run_clients(n_clients: int, rounds: int, port: int = PORT):
    import aiohttp
    url = f"http://127.0.0.1:{port}/tts"
    first, total, nbytes = [], [], []
    conn = aiohttp.TCPConnector(limit=0)
    async with aiohttp.ClientSession(connector=conn, timeout=aiohttp.ClientTimeout(total=600)) as sess:
        async def client(k):
            for r in range(rounds):
                t0 = time.perf_counter()
                async with sess.post(url, json={"character_name": "Mika", "text": f"要求{k}の{r}番目の文です。"}) as resp:
                    assert resp.status == 200, resp.status
                    got, t_first = 0, None
                    async for chunk in resp.content.iter_any():
                        if t_first is None and chunk:
                            t_first = time.perf_counter()
                        got += len(chunk)
                t1 = time.perf_counter()
                first.append(1000 * ((t_first or t1) - t0))
                total.append(1000 * (t1 - t0))
                nbytes.append(got)
        t0 = time.perf_counter()
        await asyncio.gather(*[client(k) for k in range(n_clients)])
        wall = time.perf_counter() - t0
        async with sess.get(f"http://127.0.0.1:{port}/stats") as resp:
            stats = await resp.json()
    return first, total, nbytes, wall, stats


def wait_up(srv, port):
    import urllib.request
    for _ in range(600):                              # model load x n_gpus + warm-up
        try:
            urllib.request.urlopen(f"http://127.0.0.1:{port}/stats", timeout=2).read()
            return
        except Exception:
            if srv.poll() is not None:
                raise RuntimeError("server process died, see gpurun_out/config5_server*.err")
            time.sleep(0.5)
    raise RuntimeError("server did not come up")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--serve", type=int, default=0)
    ap.add_argument("--first-device", type=int, default=0)
    ap.add_argument("--port", type=int, default=PORT)
    ap.add_argument("--client-of", type=int, default=-1, help="internal: run the clients of server process i")
    ap.add_argument("--clients", type=int, default=512)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--procs", type=int, default=1,
                    help="server processes (each owns gpus / procs GPUs and its own port; clients are spread evenly, "
                         "one client process per server process)")
    ap.add_argument("--contexts", type=int, default=0, help="schedulers per GPU (0: the service's default)")
    ap.add_argument("--slots", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5_load.json"))
    a = ap.parse_args()
    if a.serve:
        serve(a.serve, a.first_device, a.port)
        return
    if a.client_of >= 0:                              # one client process: warm, wait for the go file, run, dump raw numbers
        asyncio.run(run_clients(a.clients, 2, a.port))
        open(a.out + ".ready", "w").close()
        while not os.path.exists(a.out + ".go"):
            time.sleep(0.005)
        t_start = time.time()
        first, total, nbytes, wall, stats = asyncio.run(run_clients(a.clients, a.rounds, a.port))
        with open(a.out, "w") as f:
            json.dump({"first": first, "total": total, "nbytes": nbytes, "t_start": t_start, "t_end": time.time(),
                       "stats": stats}, f)
        return
    from genie_tts import _native as N
    n_gpus = a.gpus or N.lib().genie_device_count()
    if a.contexts:
        os.environ["GENIE_CONTEXTS_PER_GPU"] = str(a.contexts)
    os.environ["LOAD_SLOTS"] = str(a.slots)
    procs = max(1, min(a.procs, n_gpus))
    per = n_gpus // procs
    out_dir = os.path.join(ROOT, "gpurun_out")
    me = os.path.abspath(__file__)
    servers, clients = [], []
    from conftest import fixture_dir
    fixture_dir("v2", 0)                              # written once here, not by P server processes at the same time
    try:
        for i in range(procs):
            servers.append(subprocess.Popen(
                [sys.executable, me, "--serve", str(per), "--first-device", str(i * per), "--port", str(PORT + i)],
                stdout=subprocess.DEVNULL, stderr=open(os.path.join(out_dir, f"config5_server{i}.err"), "w")))
        for i, srv in enumerate(servers):
            wait_up(srv, PORT + i)
        if procs == 1:
            asyncio.run(run_clients(a.clients, 2))                            # warm: every graph bucket, every replica
            first, total, nbytes, wall, stats = asyncio.run(run_clients(a.clients, a.rounds))
        else:
            tmp = [os.path.join(out_dir, f"config5_client{i}.json") for i in range(procs)]
            for t in tmp:
                for suffix in ("", ".ready", ".go"):
                    if os.path.exists(t + suffix):
                        os.remove(t + suffix)
            for i in range(procs):
                clients.append(subprocess.Popen(
                    [sys.executable, me, "--client-of", str(i), "--port", str(PORT + i), "--clients",
                     str(a.clients // procs), "--rounds", str(a.rounds), "--out", tmp[i]]))
            while not all(os.path.exists(t + ".ready") for t in tmp):         # every client process has warmed its server
                if any(c.poll() not in (None, 0) for c in clients):
                    raise RuntimeError("a client process failed")
                time.sleep(0.05)
            for t in tmp:
                open(t + ".go", "w").close()
            for c in clients:
                if c.wait(timeout=900) != 0:
                    raise RuntimeError("a client process failed")
            parts = [json.load(open(t)) for t in tmp]
            first = sum((q["first"] for q in parts), [])
            total = sum((q["total"] for q in parts), [])
            nbytes = sum((q["nbytes"] for q in parts), [])
            wall = max(q["t_end"] for q in parts) - min(q["t_start"] for q in parts)
            stats = {"mika": sum((q["stats"]["mika"] for q in parts), [])}
            for t in tmp:
                for suffix in ("", ".ready", ".go"):
                    if os.path.exists(t + suffix):
                        os.remove(t + suffix)
    finally:
        for srv in servers:
            srv.terminate()
        for srv in servers:
            try:
                srv.wait(timeout=20)
            except Exception:
                srv.kill()
        for c in clients:
            if c.poll() is None:
                c.kill()
    audio_s = sum(nbytes) / 2 / 32000.0
    res = {"config": "BASELINE configs[4]: FastAPI server, continuous batching, closed-loop synthetic clients",
           "n_gpus": n_gpus, "server_processes": procs, "contexts_per_gpu": len(stats["mika"]) // n_gpus,
           "slots": a.slots, "clients": a.clients, "requests": len(first), "rounds_per_client": a.rounds,
           "sentence": "JA20 shape: 40-60 phonemes, 132 prompt tokens, 90-token budget, Philox sampling",
           "first_audio_ms_p50": float(np.percentile(first, 50)), "first_audio_ms_p99": float(np.percentile(first, 99)),
           "request_ms_p50": float(np.percentile(total, 50)), "request_ms_p99": float(np.percentile(total, 99)),
           "audio_s": audio_s, "wall_s": wall, "audio_s_per_s": audio_s / wall, "replicas": stats}
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "replicas"}))


if __name__ == "__main__":
    main()
