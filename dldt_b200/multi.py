"""Batch sharding across the GPUs of one box (mode B): every device holds a replica of the compiled
network for batch N/G, one user request of batch N is scattered as contiguous dim-0 slices and gathered back.  No data-path
collective exists on this path: rows of the batch are independent for every operation (BN is folded, softmax/LayerNorm are
per-sample, DetectionOutput loops per image - mkldnn_detection_output_node.cpp:152,218), so the only cross-shard step is the
re-packing of DetectionOutput's single [1,1,N*K,7] tensor (image ids renumbered, -1 terminator moved to the end).

The reference scales the other way (MULTI: whole requests to whichever device is idle, multi_device_exec_network.cpp:110-136);
bench.py reports that mode.  Both use the helpers below, which are plain host code and are covered by world-size-2 gloo tests.
"""
import numpy as np

from .ie import Core


def shard_range(global_batch, world, rank):
    """Contiguous [begin, end) slice of the batch owned by `rank`; the first (global_batch % world) ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(global_batch, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def merge_detection_outputs(parts, offsets):
    """parts[k]: DetectionOutput result [1,1,R_k,7] of shard k (rows [image_id, label, conf, xmin, ymin, xmax, ymax], first unused
    row has image_id -1, the rest are zero: mkldnn_detection_output_node.cpp:505-536); offsets[k]: index of the shard's first
    image in the global batch.  Returns the tensor the unsharded network would have produced: [1,1,sum(R_k),7]."""
    rows = []
    total = 0
    for part, off in zip(parts, offsets):
        p = np.asarray(part, np.float32).reshape(-1, 7)
        total += p.shape[0]
        end = np.nonzero(p[:, 0] < 0)[0]
        valid = p[: int(end[0])] if end.size else p
        valid = valid.copy()
        valid[:, 0] += off
        rows.append(valid)
    out = np.zeros((total, 7), np.float32)
    cat = np.concatenate(rows, axis=0) if rows else np.zeros((0, 7), np.float32)
    out[: cat.shape[0]] = cat
    if cat.shape[0] < total:
        out[cat.shape[0], 0] = -1
    return out.reshape(1, 1, total, 7)


def merge_outputs(name, parts, offsets, batch_axis_outputs=True):
    """Gather step for one output: DetectionOutput-shaped tensors are re-packed, everything else is concatenated on dim 0."""
    first = np.asarray(parts[0])
    if first.ndim == 4 and first.shape[0] == 1 and first.shape[1] == 1 and first.shape[3] == 7:
        return merge_detection_outputs(parts, offsets)
    return np.concatenate([np.asarray(p) for p in parts], axis=0)


class BatchShardedNetwork:
    """One replica per device, built for the shard's batch by `builder(batch) -> Function`."""

    def __init__(self, builder, global_batch, devices, config=None, input_precisions=None):
        self.global_batch = global_batch
        self.devices = list(devices)
        core = Core()
        self.ranges = [shard_range(global_batch, len(self.devices), r) for r in range(len(self.devices))]
        self.nets, self.requests = [], []
        for dev, (b, e) in zip(self.devices, self.ranges):
            if e == b:
                self.nets.append(None)
                self.requests.append(None)
                continue
            net = core.load_network(builder(e - b), dev, config or {})
            self.nets.append(net)
            self.requests.append(net.create_infer_request(input_precisions=input_precisions))

    def infer(self, inputs):
        live = [(r, rng) for r, rng in zip(self.requests, self.ranges) if r is not None]
        for req, (b, e) in live:           # scatter + launch everything before the first wait: the devices run concurrently
            for name, arr in inputs.items():
                req.set_blob(name, np.asarray(arr)[b:e])
            req.start_async()
        for req, _ in live:
            req.wait()
        names = list(live[0][0].output_blobs)
        offsets = [b for _, (b, _) in live]
        return {n: merge_outputs(n, [req.output_blobs[n].copy() for req, _ in live], offsets) for n in names}
