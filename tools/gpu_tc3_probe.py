"""GPU probe for the CTA-pair kernel (conv_tc3.cu): every case runs in its own process (a trap or a hang in one case cannot take
the others down) with B200_TC3=1 so that small shapes take the pair kernel too.  Prints rel-L2 against the numpy oracle."""
import os, subprocess, sys
import numpy as np

CASES = {
    # name: (N, C, H, W, O, k, stride, pad, residual, act, precision, extra env)
    "mat_ws_bn256_tail":   (4, 64, 28, 28, 256, 1, 1, 0, False, "relu", "BF16", {}),
    "mat_ws_bn64":         (4, 256, 28, 28, 64, 1, 1, 0, False, "relu", "BF16", {}),
    "mat_res_ws_bn256":    (3, 256, 20, 20, 256, 1, 1, 0, True, "relu", "BF16", {}),
    "mat_nows_bn256_x2":   (2, 512, 14, 14, 512, 1, 1, 0, False, "relu", "BF16", {}),
    "mat_bn128_nows":      (2, 128, 14, 14, 128, 1, 1, 0, False, "relu", "BF16", {"B200_NO_WS": "1"}),
    "mat_res_nows":        (2, 512, 14, 14, 512, 1, 1, 0, True, "relu", "BF16", {}),
    "mat_swish_fp16":      (2, 64, 16, 16, 192, 1, 1, 0, False, "swish", "FP16", {}),
    "im2col_s2":           (2, 128, 30, 30, 256, 3, 2, 1, False, "relu", "BF16", {}),
    "im2col_3x3_7x7":      (5, 512, 7, 7, 512, 3, 1, 1, False, "relu", "BF16", {}),
    "im2col_norows_56":    (2, 64, 56, 56, 64, 3, 1, 1, False, "relu", "BF16", {"B200_TC3_HALO": "0"}),
    "halo_56_c128_o256":   (1, 128, 56, 56, 256, 3, 1, 1, False, "relu", "BF16", {}),
    "halo_28_res_nows":    (3, 128, 28, 28, 128, 3, 1, 1, True, "relu", "BF16", {"B200_NO_WS": "1"}),
    "halo_14_c256_b16":    (16, 256, 14, 14, 256, 3, 1, 1, False, "relu", "BF16", {}),
    "halo_14_partial_res": (3, 128, 14, 14, 128, 3, 1, 1, True, "relu", "BF16", {"B200_NO_WS": "1", "B200_TC3_HALO": "2"}),
    "s2d_stem_64":         (2, 3, 64, 64, 64, 7, 2, 3, False, "relu", "BF16", {}),
    "s2d_stem_w124":       (1, 3, 40, 248, 64, 7, 2, 3, False, "relu", "BF16", {}),
    "s2d_stem_3seg_fallbk": (1, 3, 40, 520, 64, 7, 2, 3, False, "relu", "BF16", {}),
    "s2d_stem_300_fallbk": (1, 3, 300, 300, 64, 7, 2, 3, False, "relu", "BF16", {}),
    "s2d_stem_fp16_rows":  (3, 3, 50, 40, 64, 7, 2, 3, False, "relu", "FP16", {}),
    "halo_56_ws":          (2, 64, 56, 56, 64, 3, 1, 1, False, "relu", "BF16", {}),
    "halo_28":             (3, 128, 28, 28, 128, 3, 1, 1, False, "relu", "BF16", {}),
    "halo_14_c256":        (5, 256, 14, 14, 256, 3, 1, 1, False, "relu", "BF16", {}),
    "halo_28_res_fp16":    (3, 128, 28, 28, 128, 3, 1, 1, True, "relu", "FP16", {}),
}


def run_case(name):
    sys.path.insert(0, ".")
    from dldt_b200 import Core, Function
    from oracle import interp
    from tests.util import rel_l2
    n, c, h, w, o, k, s, p, res, act, prec, _ = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % 2**31)
    f = Function("conv")
    x = f.parameter((n, c, h, w))
    y = f.convolution(x, (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32), (s, s), (p, p), (p, p), (1, 1))
    y = f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32))
    if res:
        y = f.relu(f.add(y, x))
    elif act == "relu":
        y = f.relu(y)
    else:
        y = f.swish(y)
    f.result(y)
    xin = rng.standard_normal((n, c, h, w)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    impl = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    err = rel_l2(got, ref)
    # where do the errors sit (rows mode debugging): per output column w and per channel block
    d = np.abs(got - ref)
    bad_w = np.nonzero(d.max(axis=(0, 1, 2)) > 0.1)[0]
    bad_h = np.nonzero(d.max(axis=(0, 1, 3)) > 0.1)[0]
    bad_n = np.nonzero(d.max(axis=(1, 2, 3)) > 0.1)[0]
    print("%-20s %s rel_l2 %.3e  impl %s  bad n=%s h=%s w=%s" % (name, "OK  " if err < 1e-2 else "BAD ", err, impl, bad_n[:8], bad_h[:8], bad_w[:8]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        sys.exit(0)
    for name, case in CASES.items():
        env = dict(os.environ, B200_TC3="1")
        env.update(case[-1])
        try:
            r = subprocess.run([sys.executable, __file__, name], env=env, capture_output=True, text=True, timeout=90)
            out = (r.stdout.strip() or "") + ("" if r.returncode == 0 else "  [exit %d] %s" % (r.returncode, r.stderr.strip()[-400:]))
        except subprocess.TimeoutExpired:
            out = "%-20s TIMEOUT (hang)" % name
        print(out, flush=True)
