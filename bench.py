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
             53 conv + pool + FC kernels -> egress) timed with CUDA events on the request's stream.
  e2e        images/s through the public API (InferRequest.start_async / wait) with three requests in
             flight: every step copies the U8 NCHW batch from pinned host memory to the device and
             reads the fp32 [B,1000] scores back (wall clock between synchronisations).
  roofline   the tcgen05 implicit-GEMM kernel (all Convolution/FullyConnected launches of one step):
             algorithmic FLOPs / sum of their CUDA-event durations, against the measured bf16 peak.
  cpu_baseline  the reference CPU plugin on this box's host cores, bounded sample (rank 0, N = 1).
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
IE_RUN = os.path.join(REF_DIR, "bin", "ie_run")


def reference_cpu_run(batch, warmup, steps, bf16=True):
    """Times the UNMODIFIED reference CPU plugin (oracle/_ref: libinference_engine + libMKLDNNPlugin built
    by oracle/build_ref.sh) on ResNet-50 through Core::LoadNetwork(net,"CPU") + StartAsync/Wait
    (tools/ie_run.cpp).  Returns dict(value img/s, ms_per_step, cores, kind, sample)."""
    import numpy as np
    cores = os.cpu_count() or 1
    if os.path.exists(IE_RUN) and os.path.exists(os.path.join(REF_DIR, "lib", "libMKLDNNPlugin.so")):
        from dldt_b200 import ir, zoo
        f = zoo.resnet50(batch=batch)
        tmp = tempfile.mkdtemp(prefix="b200_bench_ref_")
        xml = os.path.join(tmp, "resnet50.xml")
        ir.serialize(f, xml)
        x = np.random.default_rng(0).integers(0, 256, (batch, 3, 224, 224), dtype=np.uint8)
        xin = os.path.join(tmp, "data.bin")
        x.tofile(xin)
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = os.path.join(REF_DIR, "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
        # all host threads (torchrun exports OMP_NUM_THREADS=1 to its children; only rank 0 runs this leg)
        env["OMP_NUM_THREADS"] = str(cores)
        cmd = [IE_RUN, "--model", xml, "--device", "CPU", "--output-dir", tmp, "--niter", str(warmup + steps),
               "--plugins", os.path.join(REF_DIR, "lib", "plugins.xml"), "--input", "data=" + xin, "--input-precision", "data=U8",
               "--config", "ENFORCE_BF16=%s" % ("YES" if bf16 else "NO")]
        p = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if p.returncode != 0 and bf16:   # host without avx512_bf16: the CPU plugin rejects ENFORCE_BF16=YES
            cmd[-1] = "ENFORCE_BF16=NO"
            bf16 = False
            p = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("reference run failed: " + p.stderr[-1500:])
        man = json.load(open(os.path.join(tmp, "manifest.json")))
        ms = man["infer_ms"][warmup:]
        per = sum(ms) / len(ms)
        return {"value": batch * 1e3 / per, "unit": UNIT, "cores": cores, "kind": "reference", "ms_per_step": per,
                "sample": "reference MKLDNNPlugin (oracle/_ref), ResNet-50 v1.5 %s, U8 input, %d sync requests of batch %d after %d warm-up, "
                          "1 stream x %d OMP threads" % ("bf16 (ENFORCE_BF16=YES)" if bf16 else "fp32", len(ms), batch, warmup, cores)}
    # the reference build did not travel: time the numpy restatement instead (single process, numpy's BLAS threads)
    from dldt_b200 import zoo
    from oracle import interp
    b = min(batch, 4)
    f = zoo.resnet50(batch=b)
    x = np.random.default_rng(0).uniform(0, 255, (b, 3, 224, 224)).astype(np.float32)
    interp.run(f, {"data": x})
    t0 = time.perf_counter()
    n = max(1, min(steps, 3))
    for _ in range(n):
        interp.run(f, {"data": x})
    per = (time.perf_counter() - t0) / n * 1e3
    return {"value": b * 1e3 / per, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": per,
            "sample": "numpy oracle port (oracle/interp.py), ResNet-50 v1.5 fp32, %d passes of batch %d" % (n, b)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = reference_cpu_run(args.ref_batch, args.warmup, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": round(cb["value"], 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(cb["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if "bf16" in cb["sample"] else "f32", "data": "synthetic",
            "config": {"workload": "ResNet-50 v1.5 224x224 random-init, U8 NCHW input, CPU plugin, batch %d per step (bounded sample of the b256 workload; "
                                   "throughput is per image)" % args.ref_batch},
            "cpu_baseline": {"value": round(cb["value"], 2), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]},
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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

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
    e2e_value = world * B * args.steps / e2e_s
    h2d = int(sum(b.nbytes for b in req.input_blobs.values()))
    d2h = int(sum(b.nbytes for b in req.output_blobs.values()))

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
    achieved = tc_flops / (tc_us * 1e-6) / 1e12 if tc_us > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    # traffic: DRAM read+write bytes per tcgen05 launch from the committed ncu pass over one step (profiles/r1_traffic.json, written
    # from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` by tools/summarise_ncu.py; never measured under this run)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = {"dram_bytes_per_launch": tj["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"),
                   "source": tj.get("source")}
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "conv_tc3_kernel (CTA-pair) + conv_tc2_kernel, tcgen05 implicit GEMM; all %d Convolution/FullyConnected "
                                               "launches of a step" % n_tc,
                "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "peak_source": "%s bf16 sustained" % pk["source"], "traffic": traffic,
                "share_of_step": round(tc_us / all_us, 4) if all_us else None,
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
                    "requests_in_flight": nreq},
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "clocks": clocks, "burst_ms_per_step": bursts, "latency_b1": lat}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = reference_cpu_run(args.ref_batch, 1, args.ref_steps)
            line["cpu_baseline"] = {"value": round(cb["value"], 2), "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]}
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
    ap.add_argument("--ref-batch", type=int, default=32, help="batch of the CPU reference sample")
    ap.add_argument("--ref-steps", type=int, default=8, help="timed CPU requests of the cpu_baseline leg")
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
