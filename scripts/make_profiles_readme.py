#!/usr/bin/env python
"""Regenerate profiles/README_round1.md (the round-1 tables) from the JSON results copied under profiles/bench/.  The round-2 summary, profiles/README.md, is written by hand from profiles/bench/r2/."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
hbm = peaks["hbm_gbs"]


def load(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def table(title, files):
    out = [f"### {title}", "",
           "| file | impl | algo | N | workers/GPU | iterations/s | ms/iteration | GB/s per GPU | of measured HBM | e2e iterations/s |",
           "|---|---|---|---|---|---|---|---|---|---|"]
    for f in files:
        d = load(f)
        if not d or "value" not in d:
            continue
        e2e = (d.get("e2e") or {}).get("value")
        out.append(f"| {os.path.basename(f)} | {d['impl']} | {d['config']['model'].split()[0]} | {d['n_gpus']} | "
                   f"{d['config'].get('workers_per_gpu')} | {d['value']:.1f} | {d['ms_per_step']:.4f} | {d['hbm_gbps_per_gpu']:.0f} | "
                   f"{d['hbm_gbps_per_gpu'] / hbm * 100:.1f}% | {'' if e2e is None else f'{e2e:.1f}'} |")
    return out + [""]


out = ["# Measured results (B200, this repo)", "",
       "All numbers are device-timed with CUDA events on the launching stream (worker streams fork/join on those events), max over",
       "ranks, after warm-up; every round streams a fresh Bernoulli sample (inputs >> the 126 MB L2) out of HBM-resident shards;",
       "SM clocks / throttle reasons are sampled through NVML during the timed region (none seen: 1965 MHz, `reasons: []`).",
       f"Roofline denominator: measured copy bandwidth `hbm_gbs` = {hbm} GB/s from `MEASURED_PEAKS.json` (\"of measured\").",
       "Headline workload (BASELINE.json config 2): ASGD logistic regression, mnist8m shape (8.1 M x 784 fp32), sampling rate 0.1;",
       "one iteration = every worker completes one round = 810 000 rows sampled, gradient computed, pushed and applied.", ""]
for run, title in (("run11", "Final build (prologue overlap, early clock ticket, guided scheduling) — 8-GPU box.  PDL was still the default in this session: every row except `_nopdl` ran with programmatic dependent launch; the `_w1` vs `_w1_nopdl` pair (-3 %) is why the shipped default is PDL off"),
                   ("run9", "Previous build (guided scheduling + early prefetch + PDL) — 8-GPU box"),
                   ("", "First scaling pass (register double-buffered kernel, 2 workers/GPU) — 8-GPU box")):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench", run, "*.json"))) if run else \
        sorted(glob.glob(os.path.join(ROOT, "profiles", "bench", "sc_*.json")))
    if files:
        out += table(title, files)
        ours = {}
        nccl = {}
        for f in files:
            d = load(f)
            if not d or "value" not in d:
                continue
            base = os.path.basename(f)
            if d["impl"] == "ours" and d["config"]["model"].startswith("ASGD") and "_w1" not in base and "graph" not in base \
                    and "nopdl" not in base and "seedk" not in base:
                ours[d["n_gpus"]] = max(ours.get(d["n_gpus"], 0), d["value"])
            if d["impl"] == "nccl":
                nccl[d["n_gpus"]] = d["value"]
        if 1 in ours:
            out.append("Strong scaling (fixed 8.1 M-row dataset): " + ", ".join(
                f"N={n}: {v:.0f} it/s ({v / ours[1] / n * 100:.0f} %)" for n, v in sorted(ours.items())))
        if nccl:
            out.append("")
            out.append("vs the PyTorch + NCCL + cuBLAS baseline (`baseline/nccl_sync.py`, same box): " + ", ".join(
                f"N={n}: {ours[n] / v:.1f}x" for n, v in sorted(nccl.items()) if n in ours))
        out.append("")
cfgs = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench", "*", "configs*.jsonl")))
if cfgs:
    out += ["## Other BASELINE configs (8 GPUs, `benchmarks/run_config.py`)", "",
            "| config | algo | delay intensity | workers | iterations/s | ms/iteration | device ms for the run | objective start -> end | mode |", "|---|---|---|---|---|---|---|---|---|"]
    eq = {}
    for c in cfgs:
        for line in open(c):
            r = json.loads(line)
            total_ms = r['ms_per_iteration'] * r['iterations']
            out.append(f"| {r['config']} | {r['algo']} | {r['delay_intensity']} | {r['workers']} | {r['iterations_per_s']:.0f} | "
                       f"{r['ms_per_iteration']:.4f} | {total_ms:.1f} | {r['objective_start']:.4f} -> {r['objective_end']:.4f} | {r.get('mode', 'fixed rounds per worker')} ({os.path.basename(os.path.dirname(c))}) |")
            if r['config'] == 'straggler' and 'run11' in c:
                eq.setdefault(r['delay_intensity'], {})[r['algo']] = (total_ms, r['objective_end'])
    out.append("")
    if eq:
        out += ["Wall-clock (device time) to the same training loss, equal work (300 x 8 sampled batches; run11): " + "; ".join(
            f"delay {d}: ASAGA {v['asaga'][0]:.1f} ms (loss {v['asaga'][1]:.4f}) vs synchronous SAGA {v['saga_sync'][0]:.1f} ms "
            f"(loss {v['saga_sync'][1]:.4f}) = {v['saga_sync'][0] / v['asaga'][0]:.2f}x" for d, v in sorted(eq.items()) if 'asaga' in v and 'saga_sync' in v) + ".", ""]
tm = os.path.join(ROOT, "profiles", "round_timeline_model.json")
if os.path.exists(tm):
    t = json.load(open(tm))
    out += ["## Latency-aware roofline at 8 GPUs (MODEL, not a measurement: `benchmarks/round_timeline_model.py`)", "",
            "The HBM roofline (bytes / measured copy bandwidth = 20.7 k it/s at N=8) ignores the ~29 us of serial work per round (inter-kernel gap 11.5,",
            "pull 3.7, fill/drain 10, tail 4 us; device timestamps above).  A fluid timeline of the co-resident workers with those measured terms",
            "reproduces the measured rates (`w1_launch` / `w2_lockstep` 12.6 k vs measured 12.7-12.8 k it/s: the two workers of a GPU run in lock-step, so",
            "their serial phases do not hide behind each other) and predicted (in round 1) what persistent / pipelined round kernels could buy",
            "(alpha = share of the HBM bandwidth one of two co-resident workers can draw alone, unknown until measured):", "",
            "| alpha | " + " | ".join(k for k in next(iter(t["by_alpha"].values()))) + " |",
            "|---|" + "---|" * len(next(iter(t["by_alpha"].values())))]
    for a, row in t["by_alpha"].items():
        out.append(f"| {a} | " + " | ".join(str(v) for v in row.values()) + " |")
    out.append("")
rp = os.path.join(ROOT, "profiles", "roofline_per_path.md")      # hand-written analysis of the committed timestamps: included verbatim
if os.path.exists(rp):
    out += open(rp).read().rstrip("\n").split("\n") + [""]
hr = os.path.join(ROOT, "profiles", "host_runtime.json")
if os.path.exists(hr):
    h = json.load(open(hr))
    lp, rq, wt = h["libsvm_parser"], h["result_queue"], h["worker_table"]
    out += ["## Host runtime (C++, measured on the build container's 8 shared vCPUs, so +-30 % run to run: `benchmarks/host_runtime_bench.py`)", "",
            f"LIBSVM ingest (K14; {lp['file_mb']} MB file, {lp['rows']} rows, {lp['nnz']} non-zeros, file -> CSR NumPy arrays, median of 7):", "",
            "| parser | MB/s |", "|---|---|"]
    for k in sorted((k for k in lp if k.startswith("native_")), key=lambda k: int(k.split("_")[1][:-1])):
        out.append(f"| C++ mmap parser, {k.split('_')[1][:-1]} thread(s) | {lp[k]} |")
    out += [f"| pure-Python parser of the same grammar (`parse_libsvm_record`) | {lp['python_mb_s']} |", "",
            f"Result queue ({rq['producers']} producer threads, {rq['records']} records): {rq['single_take_mrec_s']} M records/s with single takes, "
            f"{rq['batch64_take_mrec_s']} M records/s with `take_batch(64)` (`AsyncContext.drain`).  Worker table (P = {wt['P']}, through ctypes): "
            f"{wt['task_done_mops_s']} M `task_done`/s, {wt['available_workers_scan_mops_s']} M `available_workers` scans/s.", ""]
cp = os.path.join(ROOT, "profiles", "control_plane.json")
if os.path.exists(cp):
    c = json.load(open(cp))
    out += ["## Host control plane (closure API on `local[N]` / `local-cluster[n]` / `spark://`; measured on the build container: "
            "`benchmarks/control_plane_bench.py`)", "",
            "What the reference pays EVERY round on its JVM/Netty plane (broadcast of `w`, task-binary broadcast, LaunchTask + StatusUpdate RPCs, "
            "single-threaded driver loop) is what the fused backend moved into the kernel; this is the cost of the same plane here, for the closure API.", "",
            "| RPC (loopback TCP, two envs in one process) | ask p50 | ask p99 | one-way msgs/s | chunk fetch 4 MiB | chunk fetch 64 MiB |", "|---|---|---|---|---|---|"]
    for k, label in (("rpc_plain", "plain"), ("rpc_aes_gcm", "HMAC auth + AES-256-GCM")):
        r = c[k]
        out.append(f"| {label} | {r['ask_p50_us']} us | {r['ask_p99_us']} us | {r['oneway_msgs_per_s']} | {r['chunk_fetch_4MiB_MBps']} MB/s | {r['chunk_fetch_64MiB_MBps']} MB/s |")
    out += ["", "| master | blocking job (4 tasks) | round with w = 6 KB (mnist8m): broadcast / async reduce | w = 378 KB (rcv1) | w = 32 MB |", "|---|---|---|---|---|"]
    for m, v in c["masters"].items():
        f = lambda x: f"{x['broadcast_ms']} / {x['async_round_ms']} ms"  # noqa: E731
        out.append(f"| `{m}` | {v['blocking_job_ms']} ms | {f(v['w_mnist8m_6KB'])} | {f(v['w_rcv1_378KB'])} | {f(v['w_32MB'])} |")
    out += ["", "| codec (16 MB payloads) | float parameters: ratio, compress / decompress MB/s | 19 %-dense pixel bytes: ratio, compress / decompress MB/s |", "|---|---|---|"]
    for name in ("lz4", "zstd", "zlib"):
        a, b = c["codecs"][f"{name}_float_params"], c["codecs"][f"{name}_sparse_pixels"]
        out.append(f"| {name} | {a['ratio']}x, {a['compress_MBps']} / {a['decompress_MBps']} | {b['ratio']}x, {b['compress_MBps']} / {b['decompress_MBps']} |")
    out += ["", "(an asynchronous round = `sc.broadcast(w)` + non-blocking `async_reduce` over 4 partitions + draining the 4 results; on cluster "
            "masters the broadcast is a torrent broadcast — serialise, compress, 4 MiB pieces with Adler32, peer-to-peer fetch by 2 executor processes.)", ""]
sc_ = os.path.join(ROOT, "profiles", "straggler_cpu_cluster.json")
if os.path.exists(sc_):
    d = json.load(open(sc_))
    out += [f"### Straggler sweep through the closure API on executor processes (`{d['master']}`, CPU, {d['rows']} x {d['features']}, "
            f"{d['workers']} workers; `examples/straggler_sweep.py --master ...`)", "",
            "Equal work per row pair (240 worker-updates): ASAGA applies every worker's result as it arrives, synchronous SAGA waits for all workers "
            "each iteration.  Delay = multiple of the mean task time injected into worker 0 (-1 = production-cluster model, 25 % stragglers).", "",
            "| delay | ASAGA elapsed ms | ASAGA objective | sync SAGA elapsed ms | sync SAGA objective |", "|---|---|---|---|---|"]
    runs = {(r["algorithm"], r["delay"]): r for r in d["runs"]}
    for c in sorted({r["delay"] for r in d["runs"]}, key=lambda x: (x < 0, x)):
        a, s = runs[("asaga", c)], runs[("saga_sync", c)]
        out.append(f"| {c} | {a['elapsed_ms']:.0f} | {a['objective']:.4f} | {s['elapsed_ms']:.0f} | {s['objective']:.4f} |")
    out += ["", "The synchronous run slows down with the straggler (it waits for it every iteration); the asynchronous one is nearly flat and, for the "
            "same amount of gradient work and the same step size, ends ~20x closer to the optimum (240 applied updates vs 60) — the reference's claim, reproduced on this framework's CPU closure path "
            "(per-update overhead of the process transport is ~4 ms here, so absolute times are dominated by it).", ""]
ev = os.path.join(ROOT, "profiles", "bench", "eval_bench_r1.log")
if os.path.exists(ev):
    out += ["## Objective evaluation (K12): tcgen05/TMEM GEMM vs exact CUDA-core kernel (1 GPU, 2 M x 784)", "",
            "| T snapshots | tcgen05 ms | HBM GB/s | useful TFLOP/s | CUDA-core ms | speed-up | max rel. diff |", "|---|---|---|---|---|---|---|"]
    for line in open(ev):
        if line.startswith("EVAL_RESULT "):
            r = json.loads(line[len("EVAL_RESULT "):])
            out.append(f"| {r['T']} | {r['tc_ms']:.3f} | {r['tc_hbm_gbps']:.0f} | {r['tc_useful_tflops']:.1f} | {r['cc_ms']:.2f} | "
                       f"{r['speedup']:.1f}x | {r['max_rel_diff']:.1e} |")
    out += ["", "ncu (`eval_tc_r1_ncu_raw.csv`): 492 us for 3.14 GB read = 6.38 TB/s DRAM (78 % of ncu's nominal peak), tensor pipe active 16 %,",
            "116 registers/thread, 2 TMEM accumulator buffers x 64 columns.", ""]
out += ["## Kernel evidence", "",
        "* `dense_round_r1_ncu_raw.csv`, `dense_round_r2_ncu_raw.csv` — `ncu --set full` raw metrics of the fused dense round kernel: 413-419 us",
        "  under ncu, 2.62 GB DRAM read for 2.54 GB algorithmic bytes (3 % extra: labels, ids), DRAM throughput 76.5-77.7 % of ncu's nominal",
        "  peak (6.25-6.35 TB/s), dominant stall `long_scoreboard` (HBM), tensor pipe 0 % (CUDA-core GEMV by design, DESIGN.md §2.1).",
        "* `launches_r2.csv` — launch list of the timed region: only `dense_round_kernel<7,256>`, one launch per worker round.",
        "* `sass/INDEX.md` + `sass/*.txt` — per-kernel SASS opcode histograms: `UTCHMMA`/`LDTM`/`UTMALDG` in `eval_tc_kernel`, `UBLKCP`/`SYNCS` in the",
        "  TMA round variants, `REDG.E.ADD.F32x4...STRONG.SYS` + `LDG.E.128.STRONG.SYS` (NVLink peer reductions/loads) and",
        "  `ATOMG.E.ADD.64.STRONG.SYS` (logical clock) in every round kernel, `ACQBULK`/`PREEXIT` (programmatic dependent launch).",
        "* `sanitizer/` — compute-sanitizer memcheck / racecheck / synccheck summaries.", ""]
open(os.path.join(ROOT, "profiles", "README_round1.md"), "w").write("\n".join(out) + "\n")
print("profiles/README_round1.md written,", len(out), "lines")
