"""The BASELINE.json configurations at FULL WIDTH (the real topologies, random-init), shared by
  tools/make_config_golden.py   runs them on the UNMODIFIED reference CPU plugin (oracle/_ref, MKLDNNPlugin) in fp32
                                (ENFORCE_BF16=NO, the oracle of record) AND in its own bf16 mode (ENFORCE_BF16=YES, the
                                calibration: what the reference itself loses at 16 bits on the same graph)
                                -> tests/golden/config_cases.npz
  tests/test_gpu_configs.py     the CUDA path (-m gpu) against those outputs at the north-star tolerances.
Only the batch is reduced (the CPU side has to finish in minutes); inputs are regenerated from the seed on the GPU box.
Reference harness this mirrors: inference-engine/tests/functional/shared_test_classes/include/shared_test_classes/base/
layer_test_utils.hpp:86-108 (threshold compare of plugin output against a reference implementation on generated inputs).
"""
import numpy as np

from dldt_b200 import zoo

CHUNK = {"resnet50": 16, "mobilenet_v2": 16}       # images per compiled network (both sides)
N_IMAGES = {"resnet50": 1024, "mobilenet_v2": 1024}  # >= 1000 images for the top-1 statistic
N_STORED = 64                                        # images whose full fp32 logits are stored (rel-L2 is checked on these)
BERT_BATCH, BERT_SEQ = 2, 384
SSD_BATCH = 2


def images(n, seed, res=224):
    """Synthetic 'image-like' inputs in [0,1]: a random low-resolution colour grid (2..8 cells) upsampled to res x res with a
    per-image contrast / brightness and a little pixel noise, so that different images drive the network differently
    (white noise averages out in the global pooling of a random-init network)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 3, res, res), np.float32)
    for i in range(n):
        g = int(rng.integers(2, 9))
        low = rng.uniform(0, 1, (3, g, g)).astype(np.float32)
        rep = -(-res // g)
        img = np.repeat(np.repeat(low, rep, 1), rep, 2)[:, :res, :res]
        img = img * rng.uniform(0.3, 1.0) + rng.uniform(0, 0.3)
        out[i] = np.clip(img + rng.normal(0, 0.02, img.shape), 0, 1)
    return out


def classifier(kind, batch):
    return zoo.resnet50(batch=batch) if kind == "resnet50" else zoo.mobilenet_v2(batch=batch)


def classifier_inputs(kind):
    return images(N_IMAGES[kind], {"resnet50": 50, "mobilenet_v2": 52}[kind])


def bert():
    f = zoo.bert_base(batch=BERT_BATCH, seq=BERT_SEQ)
    rng = np.random.default_rng(384)
    mask = np.ones((BERT_BATCH, BERT_SEQ), np.int32)
    mask[1, 300:] = 0                                      # one padded sequence: the -10000 mask path is exercised
    types = np.zeros((BERT_BATCH, BERT_SEQ), np.int32)
    types[:, 200:] = 1
    return f, {"input_ids": rng.integers(0, 30522, (BERT_BATCH, BERT_SEQ), dtype=np.int32), "attention_mask": mask, "token_type_ids": types}


# SSD300-ResNet34: the location heads are scaled so that the box decoder's exp() stays finite and the confidence heads so that the
# softmax is neither saturated (every score 1.0 -> order decided by ties) nor uniform (every score 1/81): scores 0.5..0.98 over ~20 labels
SSD_HEAD_SCALE, SSD_CONF_SCALE = 0.002, 0.002


def ssd():
    f = zoo.ssd300_resnet34(batch=SSD_BATCH, head_scale=SSD_HEAD_SCALE, conf_scale=SSD_CONF_SCALE)
    return f, {"data": images(SSD_BATCH, 300, res=300)}


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def detection_rows(det):
    d = np.asarray(det).reshape(-1, 7)
    n = int((d[:, 0] >= 0).sum())
    return d[:n]


def match_detections(got, ref):
    """Rows of `got` matched to rows of `ref` by (image, label) and, inside such a group, by nearest box centre (the prior id is
    not part of the output).  Returns (matched_got, matched_ref, n_unmatched_ref)."""
    mg, mr, miss = [], [], 0
    keys = {(int(r[0]), int(r[1])) for r in ref}
    for key in sorted(keys):
        rg = got[(got[:, 0] == key[0]) & (got[:, 1] == key[1])]
        rr = ref[(ref[:, 0] == key[0]) & (ref[:, 1] == key[1])]
        used = np.zeros(len(rg), bool)
        for r in rr:
            if not len(rg) or used.all():
                miss += 1
                continue
            cg = np.stack([(rg[:, 3] + rg[:, 5]) / 2, (rg[:, 4] + rg[:, 6]) / 2], 1)
            cr = np.array([(r[3] + r[5]) / 2, (r[4] + r[6]) / 2])
            d = np.where(used, np.inf, ((cg - cr) ** 2).sum(1))
            j = int(d.argmin())
            scale = max(abs(r[5] - r[3]), abs(r[6] - r[4]), 1e-3)
            if np.sqrt(d[j]) > 0.25 * scale:
                miss += 1
                continue
            used[j] = True
            mg.append(rg[j])
            mr.append(r)
    return np.asarray(mg, np.float32).reshape(-1, 7), np.asarray(mr, np.float32).reshape(-1, 7), miss
