"""Manual check (not a test): the cluster generator continues its state across calls and agrees with the single-CTA kernel."""
import os, sys, subprocess
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import wavenet_oracle as O
from bench import config_c
from wavenet_b200 import _lib
from wavenet_b200._lib import check
from wavenet_b200.wavenet import _ptr, _stream
from wavenet_b200.faster_wavenet import FasterWaveNet


def run(n, parts):
    net = FasterWaveNet(config_c(), seed=0)
    net.set_weights(O.init_weights(O.config_C(), np.random.default_rng(7), np.float32))
    net.to_gpu(0)
    window = np.random.default_rng(3).integers(0, 256, (n, net.input_width)).astype(np.int32)
    net.prime(window)
    outs = []
    for steps in parts:          # consecutive wn_gen_run calls continue from the generator state
        out = torch.empty((n, steps), dtype=torch.int32, device="cuda")
        check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_GREEDY, 0, _ptr(out), _stream()))
        outs.append(out.cpu().numpy())
    return np.concatenate(outs, axis=1)


if __name__ == "__main__":
    if len(sys.argv) > 1:        # child: save the greedy sequence
        n, parts = int(sys.argv[1]), [int(v) for v in sys.argv[3].split(",")]
        np.save(sys.argv[2], run(n, parts))
        sys.exit(0)
    for n in (1, 5, 13):
        a, b, c = "/tmp/gen_a_%d.npy" % n, "/tmp/gen_b_%d.npy" % n, "/tmp/gen_c_%d.npy" % n
        subprocess.check_call([sys.executable, __file__, str(n), a, "300"])
        subprocess.check_call([sys.executable, __file__, str(n), b, "77,123,100"])
        subprocess.check_call([sys.executable, __file__, str(n), c, "300"], env=dict(os.environ, WN_GEN_V4="0"))
        x, y, z = np.load(a), np.load(b), np.load(c)
        print("n=%d: one call vs three calls identical: %s; cluster vs single-CTA kernel identical: %s (%d of %d)" %
              (n, np.array_equal(x, y), np.array_equal(x, z), (x == z).sum(), x.size))
