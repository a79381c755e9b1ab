#!/usr/bin/env python
"""Latency-aware roofline of the fused round at N GPUs: a small fluid (bandwidth-sharing) timeline model.

The HBM roofline alone (bytes / measured copy bandwidth) explains N = 1 (96.7 %) but not N = 8 (62 %): at 8 GPUs a round
streams only ~50 us of data and ~29 us of per-round serial work (inter-kernel gap, NVLink pull, fill/drain of the compute
phase, NVLink tail) is no longer negligible.  This script turns the measured phase timestamps (profiles/README.md, run11)
into a timeline of W co-resident workers per GPU that share HBM bandwidth while computing, reproduces the measured
iteration rates, and predicts what each scheduling change would buy — the plan for the next round in numbers.

    python benchmarks/round_timeline_model.py [--gpus 8] [--json profiles/round_timeline_model.json]

Model (per GPU; GPUs are symmetric, contention at the server is not modelled):
  worker cycle = serial phase S (gap + pull + fill/drain + tail; uses no HBM bandwidth)  ->  compute phase (D bytes)
  while n workers compute, each streams at min(cap, BW / n); cap = alpha * BW is what ONE worker's share of the SMs can
  draw on its own (1 CTA/SM when two workers co-reside: alpha < 1; 2 CTAs/SM when alone: alpha = measured 0.967).
Variants:
  lockstep      W workers start together every round (what run11 shows: the two kernels launch back to back)
  antiphase     workers offset by half a cycle (device-side handshake instead of the one-time stagger sleep)
  persist       lockstep, inter-kernel gap replaced by the in-kernel round barrier (the serial persistent rounds of csrc/kernels/dense_stream.cu)
  persist+anti  both
  pipelined     ONE worker per GPU owning all SMs (2 CTAs/SM), round r+1's pull/compute overlapping round r's tail
                (double-buffered g_local / counters): serial work hides behind streaming entirely
"""
from __future__ import annotations

import argparse
import json

ROWS_PER_ITER = 810_000            # 8.1 M rows x sampling rate 0.1
ROW_BYTES = 784 * 4
BW = 6569.6e9                      # measured copy bandwidth (MEASURED_PEAKS.json)
ALPHA_FULL = 0.967                 # one worker with 2 CTAs/SM, measured at N = 1
# measured per-round serial terms at N = 8 (device timestamps, run11), microseconds
GAP_US, PULL_US, FILL_DRAIN_US, TAIL_US = 11.5, 3.7, 10.0, 4.0
PERSIST_GAP_US = 2.0               # round barrier spin + prologue instead of a kernel launch (estimate)


def simulate(workers: int, bytes_per_round: float, serial_us: float, alpha: float, offsets_us, rounds: int = 400) -> float:
    """Fluid simulation; returns the steady-state period (us) of one worker round."""
    cap = alpha * BW
    t = 0.0
    # state: (phase, remaining) phase 0 = serial (remaining us), 1 = compute (remaining bytes)
    st = [[0, serial_us + off] for off in offsets_us]
    done = [0] * workers
    marks = {}
    while min(done) < rounds:
        n = sum(1 for p, _ in st if p == 1)
        rate = min(cap, BW * ALPHA_FULL / n) if n else 0.0                      # bytes/s per computing worker
        dt = min((r if p == 0 else r / rate * 1e6) for p, r in st)
        t += dt
        for i, (p, r) in enumerate(st):
            if p == 0:
                r -= dt
                if r <= 1e-9:
                    st[i] = [1, bytes_per_round]
                else:
                    st[i][1] = r
            else:
                r -= rate * dt * 1e-6
                if r <= 1e-3:
                    done[i] += 1
                    st[i] = [0, serial_us]
                    if i == 0 and done[0] in (rounds // 2, rounds):
                        marks[done[0]] = t
                else:
                    st[i][1] = r
    return (marks[rounds] - marks[rounds // 2]) / (rounds - rounds // 2)


def predict(gpus: int, alpha: float):
    out = {}
    D_gpu = ROWS_PER_ITER * ROW_BYTES / gpus
    serial = GAP_US + PULL_US + FILL_DRAIN_US + TAIL_US
    serial_p = PERSIST_GAP_US + PULL_US + FILL_DRAIN_US + TAIL_US
    ideal_us = D_gpu / BW * 1e6
    out["hbm_roofline_it_s"] = 1e6 / ideal_us
    # one worker per GPU, one launch per round (measured: w1)
    out["w1_launch"] = 1e6 / simulate(1, D_gpu, serial, ALPHA_FULL, [0.0])
    out["w1_persist"] = 1e6 / simulate(1, D_gpu, serial_p, ALPHA_FULL, [0.0])
    # two workers per GPU, half the rows each
    half = D_gpu / 2
    out["w2_lockstep"] = 1e6 / simulate(2, half, serial, alpha, [0.0, 0.0])
    cyc = serial + half / (alpha * BW) * 1e6
    out["w2_antiphase"] = 1e6 / simulate(2, half, serial, alpha, [0.0, cyc / 2])
    out["w2_persist_lockstep"] = 1e6 / simulate(2, half, serial_p, alpha, [0.0, 0.0])
    cyc_p = serial_p + half / (alpha * BW) * 1e6
    out["w2_persist_antiphase"] = 1e6 / simulate(2, half, serial_p, alpha, [0.0, cyc_p / 2])
    # pipelined single worker: serial work of round r overlaps the streaming of round r+1; only the part of the fill/drain
    # that cannot overlap (the first/last chunk of each CTA, ~half of it) stays exposed
    exposed = FILL_DRAIN_US / 2
    out["w1_pipelined"] = 1e6 / max(D_gpu / (ALPHA_FULL * BW) * 1e6 + exposed, serial_p)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = {"gpus": a.gpus, "measured_run11_it_s": {"w1": 12749, "w2": 12823} if a.gpus == 8 else None, "by_alpha": {}}
    keys = ["hbm_roofline_it_s", "w1_launch", "w1_persist", "w2_lockstep", "w2_antiphase", "w2_persist_lockstep", "w2_persist_antiphase",
            "w1_pipelined"]
    print(f"N = {a.gpus} GPUs, predicted iterations/s (measured run11 at N = 8: w1 12 749, w2 12 823)")
    print("alpha = fraction of the HBM bandwidth ONE of two co-resident workers (1 CTA/SM) can draw on its own")
    print(f"{'alpha':>6} " + " ".join(f"{k:>21}" for k in keys))
    for alpha in (0.5, 0.6, 0.7, 0.8, 0.9):
        p = predict(a.gpus, alpha)
        res["by_alpha"][alpha] = {k: round(v) for k, v in p.items()}
        print(f"{alpha:>6} " + " ".join(f"{p[k]:>21.0f}" for k in keys))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
