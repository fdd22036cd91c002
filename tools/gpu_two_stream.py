"""Experiment (not a bench value): ResNet-50 bf16 global batch 256 as S independent sub-batches of 256/S on S streams of ONE GPU, each
its own CUDA graph, replayed concurrently from S host threads - do the ramp / tail of one stream's kernels overlap the other's work?
usage: python tools/gpu_two_stream.py [S ...]      env: as tools/gpu_r50_ab.py"""
import sys, threading, time
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo

TOTAL = 256
ITERS = 40


def run(S):
    b = TOTAL // S
    net = Core().load_network(zoo.resnet50(batch=b), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    reqs = [net.create_infer_request(input_precisions={"data": "U8"}) for _ in range(S)]
    for k, r in enumerate(reqs):
        r.input_blobs["data"][...] = np.random.default_rng(k).integers(0, 256, r.input_blobs["data"].shape, dtype=np.uint8)
        r.infer()
    single = min(reqs[0].time_launches(ITERS) / ITERS for _ in range(3))
    best = 1e9
    for rep in range(4):
        ms = [0.0] * S
        bar = threading.Barrier(S + 1)

        def work(k):
            bar.wait()
            ms[k] = reqs[k].time_launches(ITERS)
        th = [threading.Thread(target=work, args=(k,)) for k in range(S)]
        for t in th: t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th: t.join()
        wall = (time.perf_counter() - t0) * 1e3
        best = min(best, max(max(ms), 0.0))
        print("  S=%d rep %d: per-stream device ms %s, wall %.1f ms" % (S, rep, ["%.1f" % m for m in ms], wall), flush=True)
    print("[S=%d] batch %d per stream: alone %.3f ms/step (%.0f img/s); %d streams concurrently: %.3f ms per %d images -> %.0f img/s"
          % (S, b, single, b / single * 1e3, S, best / ITERS, TOTAL, TOTAL * ITERS / best * 1e3), flush=True)


if __name__ == "__main__":
    for s in [int(a) for a in sys.argv[1:]] or [1, 2]:
        run(s)
