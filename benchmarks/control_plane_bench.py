"""Micro-benchmarks of the host control plane (no GPU needed): RPC latency / throughput, bulk chunk fetch, codecs, torrent
broadcast of parameter-sized payloads, and the per-job overhead of the three kinds of master.

    python benchmarks/control_plane_bench.py [--out profiles/control_plane.json]

Why these numbers: in the reference EVERY round pays this plane — a broadcast of ``w`` (6.3 KB mnist8m, 16 KB epsilon, 378 KB rcv1,
SURVEY §2.6) plus a task-binary broadcast, a ``LaunchTask`` RPC per partition and a ``StatusUpdate`` carrying the gradient back,
all through a single-threaded driver event loop.  The fused backend moves that traffic into the kernel (NVLink); what remains on
this plane is the closure API, whose cost per job is what is measured below.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200 import AsyncSession  # noqa: E402
from asyncframework_b200.core.serializer import CompressionCodec  # noqa: E402
from asyncframework_b200.deploy import LocalStandaloneCluster  # noqa: E402
from asyncframework_b200.rpc import RpcEndpoint, RpcEnv  # noqa: E402


class Echo(RpcEndpoint):
    def __init__(self):
        self.n = 0

    def receive(self, msg, sender):
        self.n += 1

    def receive_and_reply(self, msg, ctx):
        ctx.reply(msg)


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def bench_rpc(encrypt: bool) -> dict:
    kw = dict(secret="k", encrypt=True) if encrypt else {}
    with RpcEnv("s", **kw) as server, RpcEnv("c", server=False, **kw) as client:
        ep = Echo()
        server.setup_endpoint("echo", ep)
        ref = client.endpoint_ref(server.address, "echo")
        for _ in range(200):
            ref.ask_sync(1)
        lat = []
        for _ in range(3000):
            t0 = time.perf_counter()
            ref.ask_sync(1)
            lat.append((time.perf_counter() - t0) * 1e6)
        n = 50_000
        t0 = time.perf_counter()
        for i in range(n):
            ref.send(i)
        while ep.n < n:
            time.sleep(0.001)
        send_rate = n / (time.perf_counter() - t0)
        out = {"ask_p50_us": round(pct(lat, 0.5), 1), "ask_p99_us": round(pct(lat, 0.99), 1), "oneway_msgs_per_s": round(send_rate)}
        blocks = {"4MiB": os.urandom(4 << 20), "64MiB": os.urandom(64 << 20)}
        server.set_stream_manager(blocks.get)
        for name, blob in blocks.items():
            client.fetch_chunk(server.address, name)
            reps = 8 if name == "4MiB" else 3
            t0 = time.perf_counter()
            for _ in range(reps):
                client.fetch_chunk(server.address, name)
            out[f"chunk_fetch_{name}_MBps"] = round(reps * len(blob) / (time.perf_counter() - t0) / 1e6)
        return out


def bench_codecs() -> dict:
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(4_000_000) * 0.01).astype(np.float32)                    # a parameter-like payload: 16 MB
    pixels = (rng.integers(0, 256, 16_000_000) * (rng.random(16_000_000) < 0.19)).astype(np.uint8)   # mnist-like sparsity: 16 MB
    out = {}
    for name in ("lz4", "zstd", "zlib"):
        codec = CompressionCodec.create(name)
        for label, arr in (("float_params", w), ("sparse_pixels", pixels)):
            raw = arr.tobytes()
            t0 = time.perf_counter()
            blob = codec.compress(raw)
            tc = time.perf_counter() - t0
            t0 = time.perf_counter()
            CompressionCodec.decompress(blob)
            td = time.perf_counter() - t0
            out[f"{name}_{label}"] = {"ratio": round(len(raw) / len(blob), 2), "compress_MBps": round(len(raw) / tc / 1e6),
                                       "decompress_MBps": round(len(raw) / td / 1e6)}
    return out


def bench_jobs_and_broadcast(master: str, label: str) -> dict:
    out = {}
    with AsyncSession(master, app_name=f"bench-{label}") as sc:
        ds = sc.parallelize(range(4000), 4).cache()
        for _ in range(5):
            ds.count()
        n = 60
        t0 = time.perf_counter()
        for _ in range(n):
            ds.count()
        out["blocking_job_ms"] = round((time.perf_counter() - t0) / n * 1e3, 3)
        # the ASYNC round shape: broadcast w, submit a non-blocking reduce over all workers, drain the results
        ctx = sc.async_context(4)
        for d, name in ((784, "w_mnist8m_6KB"), (47_236, "w_rcv1_378KB"), (4_000_000, "w_32MB")):
            w = torch.zeros(d, dtype=torch.float64)
            rounds = 30 if d < 1_000_000 else 4
            t_b = t_r = 0.0
            for r in range(rounds + 2):
                t0 = time.perf_counter()
                bc = sc.broadcast(w)
                t1 = time.perf_counter()
                ds.map_partitions(lambda it, bc=bc: [float(bc.value[0]) + sum(1 for _ in it)]).async_reduce(lambda a, b: a + b, ctx)
                for _ in range(4):
                    ctx.collect_all(timeout=30)
                t2 = time.perf_counter()
                if r >= 2:
                    t_b += t1 - t0
                    t_r += t2 - t1
            out[name] = {"broadcast_ms": round(t_b / rounds * 1e3, 3), "async_round_ms": round(t_r / rounds * 1e3, 3)}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "control_plane.json"))
    a = ap.parse_args()
    res = {"host": f"{os.cpu_count()} vCPUs (shared build container)", "rpc_plain": bench_rpc(False), "rpc_aes_gcm": bench_rpc(True),
           "codecs": bench_codecs(), "masters": {}}
    res["masters"]["local[4]"] = bench_jobs_and_broadcast("local[4]", "local")
    res["masters"]["local-cluster[2]"] = bench_jobs_and_broadcast("local-cluster[2]", "lc")
    with LocalStandaloneCluster(num_workers=2, cores_per_worker=1) as c:
        res["masters"]["spark:// (2 workers)"] = bench_jobs_and_broadcast(c.master_url, "standalone")
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
