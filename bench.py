#!/usr/bin/env python
"""bench.py - the headline measurement: ResNet-50 v1.5 bf16 batch-256 images/s on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repository's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference CPU plugin (oracle/_ref)

One "step" = one inference request of the workload batch (BASELINE.json configs[2]; SURVEY.md 8d
config 3).  N > 1 is launched by torch.distributed.run, one rank per GPU; every rank owns a full
replica of the compiled network and its own batch (the reference's MULTI scaling model,
multi_device_exec_network.cpp:49-136; batch rows are independent so there is NO data-path
collective) -> "scaling": "weak".  torch.distributed is used only for the barrier and the
max-over-ranks reduction of the timings.

Numbers on the JSON line:
  value      images/s with the U8 input already resident in HBM: K CUDA-graph replays (ingest ->
             49 tcgen05 conv kernels (max pool in the stem's epilogue, shortcuts as extra K) + avg pool +
             FC -> egress) timed with CUDA events on the request's stream.
  e2e        images/s through the DROP-IN boundary: the UNMODIFIED reference benchmark_app (oracle/_ref/bin/benchmark_app,
             samples/benchmark_app/main.cpp:457-629) loads plugin/lib/libb200Plugin.so through InferenceEngine::Core and runs
             `-d B200 -b 256 -api async`: every request copies the U8 NCHW batch from pinned host memory to the device and
             reads the fp32 [B,1000] scores back.  N > 1: rank 0 runs ONE benchmark_app process over `-d MULTI:B200.0,..,B200.N-1`
             (the reference's own scheduler: whole requests to idle devices) while the other ranks idle.
             `e2e_python_mirror` keeps the same measurement through the Python ctypes mirror (diagnostic).
  strong     the batch-sharding scheduler of the plugin: ONE process, `-d B200.ALL`-style device set of N GPUs, global batch 256
             split into N contiguous slices per request ("scaling": "strong"), measured by the same benchmark_app.
  roofline   the tcgen05 implicit-GEMM kernel (all Convolution/FullyConnected launches of one step): algorithmic FLOPs /
             (their share of the step x ms_per_step of the timed CUDA-graph replay), against the measured bf16 peak.
  cpu_baseline  the reference CPU plugin on this box's host cores through the same benchmark_app, bounded sample (rank 0, N = 1).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ResNet-50 img/s bf16 b256"
UNIT = "img/s"
GFLOP_PER_IMG = 8.178368512  # conv + FC, 2*MAC, layer by layer (recomputed by the compiler below)


# ------------------------------------------------------------------------------------------------
def peaks():
    """Roofline denominators: driver-measured file, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            flat = {}

            def walk(o):
                if isinstance(o, dict):
                    for k, v in o.items():
                        if isinstance(v, (int, float)):
                            flat.setdefault(k, float(v))
                        else:
                            walk(v)
            walk(d)
            out = {"source": "measured"}
            out["hbm_gbs"] = flat.get("hbm_gbs", 6650.0)
            out["bf16_tflops"] = flat.get("bf16_tflops", 1590.0)
            out["bf16_tflops_sustained"] = flat.get("bf16_tflops_sustained", flat.get("bf16_tflops", 1400.0))
            return out
        except Exception:
            pass
    return {"source": "fallback", "hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        """Samples whose arrival time falls inside [t0, t1] (the timed region); the sampler itself is started before the warm-up so that
        nvidia-smi's start-up latency does not eat the window.  Falls back to the samples nearest to the window."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = list(self.lines)
        if t0 is not None and t1 is not None:
            inside = [ln for (ts, ln) in lines if t0 - 0.01 <= ts <= t1 + 0.03]
            if not inside and lines:     # window shorter than the sampling period: the nearest two samples
                mid = 0.5 * (t0 + t1)
                inside = [ln for (ts, ln) in sorted(lines, key=lambda p: abs(p[0] - mid))[:2]]
            sel = inside
        else:
            sel = [ln for (_, ln) in lines]
        for ln in sel:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": round(statistics.median(pw), 1),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BENCHMARK_APP = os.path.join(REF_DIR, "bin", "benchmark_app")
PLUGIN_SO = os.path.join(ROOT, "plugin", "lib", "libb200Plugin.so")


def write_ir(batch, tmp):
    """ResNet-50 IR v10 at `batch`, written by a HELPER process: the process that times the reference arm never imports dldt_b200
    (and so never maps libb200kernels.so)."""
    xml = os.path.join(tmp, "resnet50_b%d.xml" % batch)
    code = "import sys; sys.path.insert(0, %r); from dldt_b200 import ir, zoo; ir.serialize(zoo.resnet50(batch=%d), %r)" % (ROOT, batch, xml)
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    return xml


def benchmark_app(xml, device, seconds, extra=(), api="async", threads=None, visible=None):
    """Runs the reference's own benchmark_app and parses its report.  Returns dict(fps, latency_ms, nireq, out)."""
    import re
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(REF_DIR, "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    if visible is not None:
        env["CUDA_VISIBLE_DEVICES"] = visible
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)      # torchrun exports OMP_NUM_THREADS=1 to its children
    else:
        env.pop("OMP_NUM_THREADS", None)
    cmd = [BENCHMARK_APP, "-m", xml, "-d", device, "-api", api, "-t", str(seconds)] + list(extra)
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=os.path.join(REF_DIR, "lib"), timeout=900)
    if p.returncode != 0:
        raise RuntimeError("benchmark_app failed (%d): %s" % (p.returncode, (p.stdout[-800:] + p.stderr[-800:]).strip()))
    fps = re.search(r"Throughput:\s+([0-9.]+)\s+FPS", p.stdout)
    lat = re.search(r"Latency:\s+([0-9.]+)\s+ms", p.stdout)
    nireq = re.search(r"(\d+) inference requests", p.stdout)
    cnt = re.search(r"Count:\s+(\d+) iterations", p.stdout)
    return {"fps": float(fps.group(1)) if fps else None, "latency_ms": float(lat.group(1)) if lat else None,
            "nireq": int(nireq.group(1)) if nireq else None, "iterations": int(cnt.group(1)) if cnt else None, "out": p.stdout}


def reference_cpu_run(batch, seconds, tmp=None):
    """Times the UNMODIFIED reference CPU plugin (oracle/_ref: libinference_engine + libMKLDNNPlugin built by oracle/build_ref.sh) with the
    reference's own benchmark_app: `-d CPU -api async` (automatic streams and request count, all host cores, bf16 where the host has
    avx512_bf16 - the CPU plugin's default), U8 image input, on ResNet-50 at `batch`; plus BASELINE.json config 1: batch 1 fp32 sync
    latency.  Returns dict(value img/s, ms_per_step, cores, kind, sample, latency_b1_fp32_ms)."""
    cores = os.cpu_count() or 1
    if os.path.exists(BENCHMARK_APP) and os.path.exists(os.path.join(REF_DIR, "lib", "libMKLDNNPlugin.so")):
        tmp = tmp or tempfile.mkdtemp(prefix="b200_bench_ref_")
        r = benchmark_app(write_ir(batch, tmp), "CPU", seconds)
        bf16 = "avx512_bf16" in open("/proc/cpuinfo").read()
        lat1 = None
        try:
            l1 = benchmark_app(write_ir(1, tmp), "CPU", 4, extra=["-enforcebf16=false"], api="sync")
            lat1 = l1["latency_ms"]
        except Exception:
            pass
        per = batch * 1e3 / r["fps"]
        return {"value": r["fps"], "unit": UNIT, "cores": cores, "kind": "reference", "ms_per_step": per, "latency_b1_fp32_ms": lat1,
                "sample": "reference benchmark_app -d CPU -api async -b %d -t %d (MKLDNNPlugin from oracle/_ref, ResNet-50 v1.5 %s, U8 input, %s requests, "
                          "automatic streams, %d host threads, %s iterations)" % (batch, seconds, "bf16 (plugin default on avx512_bf16)" if bf16 else "fp32",
                                                                                 r["nireq"], cores, r["iterations"])}
    # the reference build did not travel: time the numpy restatement instead (single process, numpy's BLAS threads)
    import numpy as np
    from dldt_b200 import zoo
    from oracle import interp
    b = min(batch, 4)
    f = zoo.resnet50(batch=b)
    x = np.random.default_rng(0).uniform(0, 255, (b, 3, 224, 224)).astype(np.float32)
    interp.run(f, {"data": x})
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        interp.run(f, {"data": x})
    per = (time.perf_counter() - t0) / n * 1e3
    return {"value": b * 1e3 / per, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": per, "latency_b1_fp32_ms": None,
            "sample": "numpy oracle port (oracle/interp.py), ResNet-50 v1.5 fp32, %d passes of batch %d" % (n, b)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    # steps/warmup scale the bounded sample: about 1 s of CPU time per requested step, clamped to [10, 60] s
    seconds = int(min(60, max(10, args.steps)))
    cb = reference_cpu_run(args.ref_batch, seconds)
    line = {"impl": "reference", "metric": METRIC, "value": round(cb["value"], 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(cb["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if "bf16" in cb["sample"] else "f32", "data": "synthetic",
            "config": {"workload": "ResNet-50 v1.5 224x224 random-init, U8 NCHW input, reference CPU plugin through benchmark_app -api async, batch %d per "
                                   "request (bounded sample: %d s)" % (args.ref_batch, seconds)},
            "cpu_baseline": {"value": round(cb["value"], 2), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]},
            "latency_b1_fp32_ms": cb["latency_b1_fp32_ms"],
            "e2e": {"value": round(cb["value"], 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import numpy as np
    # Whatever libraries print on stdout (NCCL's version banner under torchrun) must not reach the driver: fd 1 is pointed at stderr for
    # the whole run and the one JSON line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from dldt_b200 import Core, zoo

    dist = None
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION (the image's default): rank 0's stdout must be the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a second, CPU-only group for the long waits: an NCCL barrier parks a spinning kernel on every waiting rank's GPU, which would
        # steal SMs from the one process that drives all N GPUs in the benchmark_app legs below
        cpu_group = dist.new_group(backend="gloo")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    def host_barrier():
        torch.cuda.synchronize(local_rank)
        if dist is not None:
            dist.barrier(group=cpu_group)

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    f = zoo.resnet50(batch=B)
    core = Core()
    net = core.load_network(f, "B200.%d" % local_rank, {"B200_INFERENCE_PRECISION": "BF16"})
    nreq = max(2, args.inflight)
    reqs = [net.create_infer_request(input_precisions={"data": "U8"}) for _ in range(nreq)]
    rng = np.random.default_rng(1234 + rank)
    for r in reqs:
        r.input_blobs["data"][...] = rng.integers(0, 256, r.input_blobs["data"].shape, dtype=np.uint8)
        r.infer()   # first run: allocates staging, captures the CUDA graph, leaves the input resident in HBM
    req = reqs[0]
    launches_per_step = req.num_launches()
    flops_per_step = net.stat("flops")
    pk = peaks()

    # ---- value: device-resident, CUDA events on the request's stream ----
    sampler = ClockSampler(local_rank)
    if not args.no_clock_sampler:
        sampler.start()                                # before the warm-up: nvidia-smi needs ~100 ms to deliver its first line
    for _ in range(max(args.warmup, 3)):
        req.launch_device()
    req.wait()
    barrier()
    t_begin = time.time()
    ms_total = req.time_launches(args.steps)          # b200_request_time: event, K graph launches, event, sync
    t_end = time.time()
    barrier()
    clocks = sampler.stop(t_begin, t_end)
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    # diagnostics (outside the timed region): the same launch timed in short bursts, to expose clock drift / interference
    bursts = [round(req.time_launches(10) / 10, 4) for _ in range(4)]
    value = world * B * 1e3 / ms_step

    # ---- e2e: public API, pinned host buffers, `nreq` requests in flight ----
    for k in range(max(args.warmup, 3)):
        r = reqs[k % nreq]
        r.wait()
        r.start_async()
    for r in reqs:
        r.wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        r = reqs[k % nreq]
        r.wait()
        r.start_async()
    for r in reqs:
        r.wait()
    torch.cuda.synchronize(local_rank)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_mirror = world * B * args.steps / e2e_s
    h2d = int(sum(b.nbytes for b in req.input_blobs.values()))
    d2h = int(sum(b.nbytes for b in req.output_blobs.values()))

    # ---- e2e through the drop-in boundary: the reference's own benchmark_app over libb200Plugin.so (rank 0; the other ranks idle) ----
    e2e_app, strong = None, None
    have_app = os.path.exists(BENCHMARK_APP) and os.path.exists(PLUGIN_SO)
    host_barrier()
    if rank == 0 and have_app and not args.no_benchmark_app:
        tmp = tempfile.mkdtemp(prefix="b200_bench_")
        xml = write_ir(B, tmp)
        try:
            dev = "B200.%d" % local_rank if world == 1 else "MULTI:" + ",".join("B200.%d" % k for k in range(world))
            benchmark_app(xml, dev, 1)                       # untimed warm-up run of the same command (file cache, plugin load, clocks)
            r = benchmark_app(xml, dev, args.app_seconds)
            e2e_app = {"value": r["fps"], "latency_ms": r["latency_ms"], "nireq": r["nireq"], "iterations": r["iterations"],
                       "how": "oracle/_ref/bin/benchmark_app -m resnet50.xml -d %s -b %d -api async -t %d (unmodified reference app + Core + "
                              "plugin/lib/libb200Plugin.so)" % (dev, B, args.app_seconds)}
        except Exception as e:
            e2e_app = {"value": None, "error": str(e)[:300]}
        try:
            # batch sharding inside the plugin: one process, global batch B split over the N GPUs of this run
            # (-d B200.ALL = DEVICE_ID=ALL = every visible GPU; the set is narrowed to this run's N GPUs through CUDA_VISIBLE_DEVICES)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            ids = [v for v in vis.split(",") if v][:world] if vis else [str(k) for k in range(world)]
            r = benchmark_app(xml, "B200.ALL", args.app_seconds, visible=",".join(ids))
            strong = {"value": r["fps"], "unit": UNIT, "scaling": "strong", "global_batch": B, "n_gpus": world, "latency_ms": r["latency_ms"],
                      "nireq": r["nireq"], "iterations": r["iterations"],
                      "how": "benchmark_app -d B200.ALL (%d visible GPUs) -b %d -api async -t %d: every request of batch %d is scattered as %d contiguous "
                             "slices of %d images, one CUDA graph per GPU, gathered on completion" % (world, B, args.app_seconds, B, world, B // world)}
        except Exception as e:
            strong = {"value": None, "scaling": "strong", "error": str(e)[:300]}
    host_barrier()
    e2e_value = e2e_app["value"] if (e2e_app and e2e_app.get("value")) else e2e_mirror

    # ---- roofline of the dominant kernel class (tcgen05 implicit GEMM), CUDA events per launch ----
    eg = net.get_exec_graph_info()
    tc_us, tc_flops, all_us = 0.0, 0.0, 0.0
    passes = 3
    for _ in range(passes):
        pc = req.get_perf_counts()
        for e in eg:
            us = pc[e["name"]]["real_time"]
            all_us += us
            if e["impl"].startswith("tcgen05"):
                tc_us += us
                tc_flops += e["flops"]
    n_tc = sum(1 for e in eg if e["impl"].startswith("tcgen05"))
    # per-kernel times come from an eager, event-bracketed pass (no overlap between kernels); the timed region replays the captured
    # CUDA graph (programmatic dependent launch overlaps prologues).  The tcgen05 launches' duration inside the timed step is therefore
    # taken as their SHARE of the eager pass x ms_per_step, so that the figure follows from the line's own ms_per_step.
    tc_share = tc_us / all_us if all_us else 0.0
    tc_ms_in_step = tc_share * ms_step
    achieved = (tc_flops / passes) / (tc_ms_in_step * 1e-3) / 1e12 if tc_ms_in_step > 0 else 0.0
    achieved_eager = tc_flops / (tc_us * 1e-6) / 1e12 if tc_us > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    # traffic: DRAM read+write bytes per tcgen05 launch from the committed ncu pass over one step (profiles/r1_traffic.json, written
    # from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` by tools/summarise_ncu.py; never measured under this run)
    traffic = None
    try:
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        tj = json.load(open(tpath if os.path.exists(tpath) else os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = {"dram_bytes_per_launch": tj["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"),
                   "source": tj.get("source")}
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "conv_tc3_kernel (CTA-pair) + conv_tc2_kernel, tcgen05 implicit GEMM; all %d Convolution/FullyConnected "
                                               "launches of a step" % n_tc,
                "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "peak_source": "%s bf16 sustained" % pk["source"], "traffic": traffic,
                "share_of_step": round(tc_share, 4), "achieved_eager_event_sum": round(achieved_eager, 2),
                "how": "sum of algorithmic FLOPs of the tcgen05 launches / (their share of an eager event-bracketed pass x ms_per_step of the timed graph replay)",
                "step_achieved": round(value / world * GFLOP_PER_IMG / 1e3, 2), "step_frac": round(value / world * GFLOP_PER_IMG / 1e3 / peak, 4)}

    # ---- batch-1 p50 latency (second half of BASELINE.json's metric): sync requests through the public API ----
    lat = None
    if not args.no_latency:
        f1 = zoo.resnet50(batch=1)
        net1 = core.load_network(f1, "B200.%d" % local_rank, {"B200_INFERENCE_PRECISION": "BF16"})
        r1 = net1.create_infer_request(input_precisions={"data": "U8"})
        r1.input_blobs["data"][...] = rng.integers(0, 256, r1.input_blobs["data"].shape, dtype=np.uint8)
        for _ in range(20):
            r1.infer()
        ts = []
        for _ in range(200):
            t1 = time.perf_counter()
            r1.start_async()
            r1.wait()
            ts.append((time.perf_counter() - t1) * 1e3)
        dev_ms = r1.time_launches(50) / 50
        lat = {"p50_ms": round(statistics.median(ts), 4), "p99_ms": round(sorted(ts)[int(len(ts) * 0.99) - 1], 4), "device_only_ms": round(dev_ms, 4),
               "how": "200 synchronous StartAsync+Wait of batch 1 (U8 input from pinned host memory, fp32 scores back), median as benchmark_app -api sync"}
        del r1, net1

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "ResNet-50 v1.5 224x224 random-init (nGraph-style graph, BN folded), U8 NCHW input, batch %d per GPU" % B,
                       "global_batch": B * world, "parallelism": "replica per GPU, batch-sharded requests, no collective",
                       "l2": "no explicit flush: one step streams a %.0f MB activation arena (> 126 MB L2)" % (net.stat("arena_bytes") / 1e6),
                       "gflop_per_img": round(flops_per_step / B / 1e9, 4)},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "through": "reference benchmark_app + InferenceEngine::Core + libb200Plugin.so" if (e2e_app and e2e_app.get("value")) else
                               "python ctypes mirror (benchmark_app or the plugin did not travel)",
                    "benchmark_app": e2e_app},
            "e2e_python_mirror": {"value": round(e2e_mirror, 1), "unit": UNIT, "requests_in_flight": nreq,
                                  "how": "dldt_b200.InferRequest.start_async/wait, one process per GPU, same C ABI underneath"},
            "strong": strong,
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "clocks": clocks, "burst_ms_per_step": bursts, "latency_b1": lat}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = reference_cpu_run(args.ref_batch, args.ref_seconds)
            line["cpu_baseline"] = {"value": round(cb["value"], 2), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                                    "latency_b1_fp32_ms": cb["latency_b1_fp32_ms"]}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
    if dist is not None:
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (the metric is quoted on 256)")
    ap.add_argument("--ref-batch", type=int, default=256, help="batch of the CPU reference requests (the workload's own)")
    ap.add_argument("--ref-seconds", type=int, default=12, help="duration of the cpu_baseline leg (benchmark_app -t)")
    ap.add_argument("--app-seconds", type=int, default=5, help="duration of each benchmark_app leg on the B200 plugin")
    ap.add_argument("--no-benchmark-app", action="store_true", help="skip the benchmark_app legs (e2e falls back to the Python mirror)")
    ap.add_argument("--inflight", type=int, default=3, help="infer requests in flight in the e2e leg (measured on B200: 2 -> 74.4k, 3 -> 76.6k, 4 -> 76.6k img/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 latency leg")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
