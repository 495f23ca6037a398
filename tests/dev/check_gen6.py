"""gen_kernel_v6 (tensor-core many-stream generator): greedy sequences vs the fp64 ring oracle and vs gen_kernel_v3,
continuation across calls, and the time per step at 256 streams.  python tests/dev/check_gen6.py [n] [steps] [time_streams]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net
from wavenet_b200 import _lib
from wavenet_b200._lib import check
from wavenet_b200.wavenet import _ptr, _stream

n = int(sys.argv[1]) if len(sys.argv) > 1 else 130
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tstreams = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cfg = make_cfg("C")
w = O.init_weights(cfg, np.random.default_rng(7), np.float64)
window = np.random.default_rng(3).integers(0, 256, (n, O.input_width(cfg))).astype(np.int32)


def run(win, parts, v6, mode=_lib.WN_GEN_GREEDY):
    os.environ["WN_GEN_V6"] = "1" if v6 else "0"
    net = make_net(cfg, w, faster=True, head_act="reference")
    net.prime(win)
    outs = []
    for st in parts:
        out = torch.empty((win.shape[0], st), dtype=torch.int32, device="cuda")
        check(net._libh.wn_gen_run(net._gen, _ptr(net._params), st, mode, 5, _ptr(out), _stream()))
        outs.append(out.cpu().numpy())
    torch.cuda.synchronize()
    return np.concatenate(outs, axis=1)


if n > 0:
    a = run(window, [steps], True)
    print("v6 ran", a.shape, a[0, :12])
    c = run(window, [steps], False)
    dv = (a != c).any(axis=1)
    print("streams differing from v3:", int(dv.sum()), "of", n, "first mismatch steps:",
          sorted(set(int(np.argmax(a[i] != c[i])) for i in np.nonzero(dv)[0]))[:10])
    t0 = time.time()
    want = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64).generate_greedy(window, steps)
    print("oracle %.1f s" % (time.time() - t0))
    print("v6 vs oracle: streams differing", int((a != want).any(axis=1).sum()), "| v3 vs oracle:", int((c != want).any(axis=1).sum()))
    b = run(window, [steps // 3, steps // 2, steps - steps // 3 - steps // 2], True)
    print("continuation across calls identical:", bool(np.array_equal(a, b)))
    a2 = run(window, [steps], True)
    print("run-to-run identical:", bool(np.array_equal(a, a2)))
    s6 = run(window, [steps], True, _lib.WN_GEN_SAMPLE)
    s3 = run(window, [steps], False, _lib.WN_GEN_SAMPLE)
    print("sample mode: streams differing from v3:", int((s6 != s3).any(axis=1).sum()))

if tstreams > 0:
    win = np.random.default_rng(0).integers(0, 256, (tstreams, O.input_width(cfg))).astype(np.int32)
    for v6 in (True, False):
        os.environ["WN_GEN_V6"] = "1" if v6 else "0"
        net = make_net(cfg, w, faster=True, head_act="reference")
        T = 2000
        out = torch.empty((tstreams, T), dtype=torch.int32, device="cuda")
        for rep in range(2):
            net.prime(win)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(net._libh.wn_gen_run(net._gen, _ptr(net._params), T, _lib.WN_GEN_SAMPLE, 0, _ptr(out), _stream()))
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s: %d streams x %d steps: %.2f us/step, %.2f M samples/s" % ("v6" if v6 else "v3", tstreams, T, 1e3 * ms / T,
                                                                             tstreams * T / ms / 1e3))
        del net
