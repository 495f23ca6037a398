# -*- coding: utf-8 -*-
"""Command line of the reference's train_audio/args.py:5-19 (module-level singleton `args`)."""
import argparse

parser = argparse.ArgumentParser()
parser.add_argument("-g", "--gpu_device", type=int, default=0)
parser.add_argument("-w", "--wav-dir", type=str, default="wav")
parser.add_argument("-m", "--model-dir", type=str, default="model")

# generation
parser.add_argument("-o", "--output_dir", type=str, default="generated_audio")
parser.add_argument("-s", "--seconds", type=float, default=1.0)
parser.add_argument("--lr", type=float, default=0.001, help="learning_rate")
# live flag of the reference is --fast (args.py:14); its README says --use_faster_wavenet (README.md:44)
parser.add_argument("--fast", "--use_faster_wavenet", dest="fast", action="store_true", default=False)

# seed
parser.add_argument("--seed", type=int, default=None)

# B200 backend extras (not in the reference)
parser.add_argument("--precision", type=str, default="fp16x2", choices=["fp16x2", "tf32", "fp32"],
                    help="fp16x2: tcgen05 on split fp16 operands, fp32-grade (default, meets the reference's fp32 arithmetic to "
                         "1e-4 logits / 1e-3 gradients); tf32: single-pass tcgen05 (faster, 1e-2 logits); fp32: exact SIMT FFMA")
parser.add_argument("--greedy", action="store_true", default=False, help="argmax decoding instead of sampling")
parser.add_argument("--deterministic", action="store_true", default=False,
                    help="bit-reproducible training: fixed-order gradient reductions instead of atomics (fused fp16x2 shape, +5 %% step time)")

args = parser.parse_args()
