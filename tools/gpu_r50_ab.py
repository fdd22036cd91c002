"""A/B of kernel-selection knobs on ResNet-50 b256 bf16 (not a bench value): for every configuration, device-resident ms/step,
the per-kernel table, and the output's rel-L2 / top-1 agreement against the first configuration (same weights, same input).
usage: python tools/gpu_r50_ab.py "B200_TC3=0" "" "B200_TC3_HALO=1" ...   (each argument = space-separated env assignments)"""
import os, sys, time, subprocess, json
import numpy as np
sys.path.insert(0, ".")

BATCH = int(os.environ.get("AB_BATCH", "256"))


def run_one(tag):
    from dldt_b200 import Core, zoo
    f = zoo.resnet50(batch=BATCH)
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    req = net.create_infer_request(input_precisions={"data": "U8"})
    req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
    out = list(req.infer().values())[0]
    np.save("gpurun_out/ab_%s.npy" % tag, out)
    best = 1e9
    for it in range(3):
        best = min(best, req.time_launches(10) / 10)
    print("[%s] R50 b%d: %.3f ms/step -> %.0f img/s (device resident)" % (tag, BATCH, best, BATCH / best * 1e3), flush=True)
    pc = req.get_perf_counts()
    tot = 0.0
    for e in net.get_exec_graph_info():
        us = pc[e["name"]]["real_time"]
        tot += us
        print("  %8.1f us %-14s %-44s %7.1f TFLOP/s %7.1f GB/s  %s" % (us, e["type"], e["impl"], e["flops"] / us / 1e6, e["bytes"] / us / 1e3, e["name"]))
    print("[%s] per-kernel total %.3f ms" % (tag, tot / 1e3), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    ref = None
    for k, cfg in enumerate(sys.argv[1:]):
        env = dict(os.environ)
        for kv in cfg.split():
            a, b = kv.split("=")
            env[a] = b
        tag = "cfg%d" % k
        print("===== %s: %s" % (tag, cfg or "(default)"), flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--one", tag], env=env, capture_output=True, text=True, timeout=240)
            print(r.stdout, flush=True)
            if r.returncode != 0:
                print("  [exit %d] %s" % (r.returncode, r.stderr[-600:]), flush=True)
                continue
        except subprocess.TimeoutExpired:
            print("  TIMEOUT", flush=True)
            continue
        out = np.load("gpurun_out/ab_%s.npy" % tag)
        os.remove("gpurun_out/ab_%s.npy" % tag)
        if ref is None:
            ref = out
        else:
            err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
            top1 = float((out.argmax(1) == ref.argmax(1)).mean())
            print("[%s] vs cfg0: rel-L2 %.3e  top-1 agreement %.4f  finite %s" % (tag, err, top1, bool(np.isfinite(out).all())), flush=True)
