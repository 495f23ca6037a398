# -*- coding: utf-8 -*-
"""Import-time model factory of the reference's train_audio/model.py:1-59."""
import json
import os
import sys

sys.path.append(os.path.split(os.path.dirname(os.path.abspath(__file__)))[0])
from args import args  # noqa: E402
from wavenet_b200.wavenet import WaveNet, Params  # noqa: E402
from wavenet_b200.faster_wavenet import FasterWaveNet  # noqa: E402

try:
    os.mkdir(args.model_dir)
except Exception:
    pass

# live file name is wavenet.json (model.py:13); README/tests say params.json (README.md:34): accept both
filename = args.model_dir + "/wavenet.json"
if not os.path.isfile(filename) and os.path.isfile(args.model_dir + "/params.json"):
    filename = args.model_dir + "/params.json"
if os.path.isfile(filename):
    print("loading", filename)
    with open(filename) as f:
        try:
            params = Params(json.load(f))
        except Exception:
            raise Exception("could not load {}".format(filename))
else:
    params = Params()
    params.quantization_steps = 256
    params.sampling_rate = 8000

    params.causal_conv_no_bias = True
    params.causal_conv_filter_width = 2
    params.causal_conv_channels = [256]

    params.residual_conv_dilation_no_bias = True
    params.residual_conv_projection_no_bias = True
    params.residual_conv_filter_width = 2
    params.residual_conv_channels = [128, 128, 128, 128, 128, 128, 128, 128]
    params.residual_num_blocks = 1

    params.softmax_conv_no_bias = False
    params.softmax_conv_channels = [256, 256]

    params.optimizer = "adam"
    params.momentum = 0.9
    params.weight_decay = 0
    params.gradient_clipping = 1.0

    with open(filename, "w") as f:
        json.dump(params.to_dict(), f, indent=4)

if args.fast:
    wavenet = FasterWaveNet(params)
else:
    wavenet = WaveNet(params)

params.dump()
wavenet.load(args.model_dir)

if args.gpu_device == -1:
    raise Exception("the B200 backend has no CPU mode (-g -1): the reference's CPU arithmetic lives in oracle/ for tests only")
wavenet.to_gpu(args.gpu_device)
wavenet.set_precision(args.precision)
if getattr(args, "deterministic", False):
    wavenet.set_deterministic(True)
