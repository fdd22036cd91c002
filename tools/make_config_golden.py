"""Generate tests/golden/config_cases.npz: the full-width BASELINE.json configurations (tests/config_cases.py) on the UNMODIFIED
reference CPU plugin (oracle/_ref MKLDNNPlugin, built from /root/reference by oracle/build_ref.sh).  Build container only:
    python tools/make_config_golden.py [resnet50 mobilenet_v2 bert ssd]

Per configuration the file holds the fp32 reference outputs (ENFORCE_BF16=NO; the oracle of record) and the CALIBRATION of the
north-star gate: the same graph and inputs through the reference's own bf16 mode (ENFORCE_BF16=YES on this avx512_bf16 host),
i.e. what MKLDNNPlugin itself loses at 16 bits (rel-L2 and top-1 agreement against its fp32 mode).  Test infrastructure only.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import config_cases as cc  # noqa: E402
from tools import refrun              # noqa: E402

PATH = os.path.join(ROOT, "tests", "golden", "config_cases.npz")


def run_classifier(kind, out, meta):
    x = cc.classifier_inputs(kind)
    chunk = cc.CHUNK[kind]
    f = cc.classifier(kind, chunk)
    l32, l16 = [], []
    t0 = time.time()
    for i in range(0, len(x), chunk):
        o32, _ = refrun.run(f, {"data": x[i:i + chunk]}, "CPU", {"ENFORCE_BF16": "NO"})
        o16, _ = refrun.run(f, {"data": x[i:i + chunk]}, "CPU", {"ENFORCE_BF16": "YES"})
        l32.append(o32["fc_bias"])
        l16.append(o16["fc_bias"])
        if (i // chunk) % 8 == 0:
            print("  %s %d/%d (%.0f s)" % (kind, i, len(x), time.time() - t0), flush=True)
    l32, l16 = np.concatenate(l32), np.concatenate(l16)
    out[kind + "/logits_fp32"] = l32[:cc.N_STORED].astype(np.float32)
    out[kind + "/top1_fp32"] = l32.argmax(1).astype(np.int16)
    out[kind + "/top1_ref_bf16"] = l16.argmax(1).astype(np.int16)
    # margin between the two largest fp32 logits: lets the test tell a real flip from a near-tie
    srt = np.sort(l32, 1)
    out[kind + "/margin_fp32"] = (srt[:, -1] - srt[:, -2]).astype(np.float32)
    meta[kind] = {"images": int(len(x)), "ref_bf16_rel_l2": cc.rel_l2(l16, l32), "ref_bf16_rel_l2_stored": cc.rel_l2(l16[:cc.N_STORED], l32[:cc.N_STORED]),
                  "ref_bf16_top1_agreement": float((l16.argmax(1) == l32.argmax(1)).mean()), "distinct_top1_classes": int(len(set(l32.argmax(1).tolist()))),
                  "logit_std": float(l32.std())}
    print(kind, meta[kind], flush=True)


def run_bert(out, meta):
    f, inputs = cc.bert()
    o32, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "NO"})
    o16, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "YES"})
    k = next(iter(o32))
    out["bert/" + k] = o32[k].astype(np.float32)
    meta["bert"] = {"output": k, "ref_bf16_rel_l2": cc.rel_l2(o16[k], o32[k])}
    print("bert", meta["bert"], flush=True)


def run_ssd(out, meta):
    f, inputs = cc.ssd()
    o32, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "NO"})
    o16, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "YES"})
    k = next(iter(o32))
    out["ssd/" + k] = o32[k].astype(np.float32)
    r32, r16 = cc.detection_rows(o32[k]), cc.detection_rows(o16[k])
    mg, mr, miss = cc.match_detections(r16, r32)
    n = min(len(r32), len(r16))
    meta["ssd"] = {"output": k, "rows": int(len(r32)), "ref_bf16_rows": int(len(r16)),
                   "ref_bf16_positional_label_agreement": float((r32[:n, :2] == r16[:n, :2]).all(1).mean()),
                   "ref_bf16_matched": int(len(mr)), "ref_bf16_unmatched": int(miss),
                   "ref_bf16_box_rel_l2": cc.rel_l2(mg[:, 3:], mr[:, 3:]) if len(mr) else None,
                   "score_range": [float(r32[:, 2].min()), float(r32[:, 2].max())], "labels": int(len(set(r32[:, 1].tolist())))}
    print("ssd", meta["ssd"], flush=True)


def main():
    assert refrun.available(), "build the reference first: bash oracle/build_ref.sh"
    which = sys.argv[1:] or ["resnet50", "mobilenet_v2", "bert", "ssd"]
    out, meta = {}, {}
    if os.path.exists(PATH):   # regenerate selectively
        old = np.load(PATH)
        for k in old.files:
            if k != "__meta__":
                out[k] = old[k]
        meta = json.loads(bytes(old["__meta__"]).decode())
    for w in which:
        for k in [k for k in out if k.startswith(w + "/")]:
            del out[k]
        if w in ("resnet50", "mobilenet_v2"):
            run_classifier(w, out, meta)
        elif w == "bert":
            run_bert(out, meta)
        elif w == "ssd":
            run_ssd(out, meta)
    meta["reference"] = "opencv/dldt @ /root/reference, MKLDNNPlugin built by oracle/build_ref.sh; fp32 = ENFORCE_BF16=NO, calibration = ENFORCE_BF16=YES"
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(PATH, **out)
    print("wrote %s (%.1f KB)" % (PATH, os.path.getsize(PATH) / 1e3))


if __name__ == "__main__":
    main()
