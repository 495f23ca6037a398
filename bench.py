#!/usr/bin/env python
"""Benchmark of the musyoku/wavenet hot paths on B200 (see BASELINE.json / SURVEY.md 8d).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code path

A "step" is one train.py step of the 30-layer config-C network on a synthetic mu-law batch of 32 x 16000 samples per
GPU: forward + cross-entropy + backward + (NCCL all-reduce when N > 1) + gradient clipping + Adam.  `value` is whole-job
audio samples/s with the batch resident in HBM; `e2e` is the same step driven through the public Python API from pinned
host buffers (H2D of samples/targets, D2H of the loss inside the timed region).

The headline runs in the fp16x2 precision mode (tcgen05 on split fp16 operands, fp32-grade: the mode that meets the 1e-4
logit / 1e-3 gradient gates); `precision_modes` reports the single-pass tf32 mode and the exact-fp32 SIMT mode next to it.

The reference arm executes the REFERENCE'S OWN SOURCE (oracle/_ref: mechanical py2->py3 transform of
/root/reference/*.py, NumPy stand-in for the Chainer primitives -- im2col + tensordot like Chainer's CPU path) on the
host cores; when oracle/_ref is absent it falls back to the NumPy oracle port.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FWD_FLOP_PER_POS = 2473984
LAYER_FLOP_PER_POS = 2 * (128 * 128 + 64 * 64 + 256 * 64)   # one residual layer incl. skip conv
B_PER_GPU, WIDTH = 32, 16000
METRIC = "train audio samples/s (config C 30-layer, fwd+loss+bwd+clip+Adam)"
UNIT = "samples/s"
DTYPE_NAMES = {"fp16x2": "f16x2 (split fp16 operands hi+lo on tcgen05, fp32 accumulate; fp32-grade)",
               "tf32": "tf32", "fp32": "f32"}


def config_c():
    from wavenet_b200.wavenet import Params
    p = Params()
    p.causal_conv_channels = [64]
    p.residual_conv_channels = [64] * 10
    p.residual_num_blocks = 3
    p.softmax_conv_channels = [256, 256, 256]
    return p


def config_b():
    from wavenet_b200.wavenet import Params
    p = Params()                       # reference default network, train_audio/model.py:24-43
    p.causal_conv_channels = [256]
    p.residual_conv_channels = [128] * 8
    p.residual_num_blocks = 1
    p.softmax_conv_channels = [256, 256]
    return p


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_batch(rank, B, W):
    """default_rng(rank).integers(0,256,(B,W)); targets = next sample (train.py:21)."""
    rng = np.random.default_rng(rank)
    x = rng.integers(0, 256, (B, W + 1)).astype(np.int32)
    return np.ascontiguousarray(x[:, :W]), np.ascontiguousarray(x[:, 1:])


# ---- CPU legs: the reference's own code path on the host cores ------------------------------------------
_CFG_KEYS = ["quantization_steps", "causal_conv_no_bias", "causal_conv_filter_width", "causal_conv_channels",
             "residual_conv_dilation_no_bias", "residual_conv_projection_no_bias", "residual_conv_filter_width",
             "residual_conv_channels", "residual_num_blocks", "softmax_conv_no_bias", "softmax_conv_channels",
             "weight_decay", "momentum", "gradient_clipping"]


def load_reference():
    """(R, RF, RD) = the executed reference modules, or None when oracle/_ref is unavailable."""
    try:
        from oracle.ref_build import import_reference
        return import_reference()
    except Exception:
        return None


def _ref_params(R, cfg):
    p = R.Params()
    for k in _CFG_KEYS:
        setattr(p, k, getattr(cfg, k))
    return p


class CpuTrainer(object):
    """One train.py:58-80 step (one-hot -> forward blocks -> slice -> cross_entropy -> backprop) on the host CPU:
    through the executed reference source when available (kind "reference"), else the NumPy oracle port."""

    def __init__(self, which, B, W, train_width):
        from oracle import wavenet_oracle as O
        self.O, self.B, self.W, self.tw = O, B, W, train_width
        self.cfg = O.config_C() if which == "C" else O.config_B()
        self.x, self.tgt = synth_batch(0, B, W)
        self.tgt = np.ascontiguousarray(self.tgt[:, W - train_width:])
        self.ref = load_reference()
        if self.ref is not None:
            R, RF, RD = self.ref
            np.random.seed(1234)
            self.net = R.WaveNet(_ref_params(R, self.cfg))       # Chainer default init (LeCunNormal), wavenet.py:379-455
            self.net.update_laerning_rate(1e-3)
            self.kind = "reference"
        else:
            self.w = O.init_weights(self.cfg, np.random.default_rng(1234), np.float32)
            self.st = O.new_adam_state(self.w)
            self.kind = "port"

    def step(self):
        if self.ref is not None:
            R, RF, RD = self.ref
            net, W, tw = self.net, self.W, self.tw
            onehot = RD.onehot_pixel_image(self.x, quantization_steps=256)
            out = net.forward_causal_block(onehot)
            out, skip = net.forward_residual_block(out)
            if W - tw >= 1:
                skip = net.slice_1d(skip, W - tw)
            y = net.forward_softmax_block(skip, apply_softmax=False)
            loss = net.cross_entropy(y, self.tgt)
            net.backprop(loss)
            return float(loss.data)
        O = self.O
        fw = O.forward_loss(self.cfg, self.w, self.x, self.tgt, train_width=self.tw, dtype=np.float32)
        g = O.backward(self.cfg, fw)
        O.clip_and_adam(self.cfg, self.w, g, self.st, lr=1e-3)
        return float(fw["loss"])

    def describe(self, what):
        impl = ("the reference's own wavenet.py executed on the host (oracle/_ref: py2->py3 transform, NumPy stand-in for the "
                "Chainer primitives: im2col + tensordot convs, define-by-run backward, clip + Adam)" if self.kind == "reference"
                else "NumPy oracle port (oracle/_ref unavailable)")
        return "%s; %s; BLAS threads = all host cores" % (what, impl)


def cpu_generation_sample(n_steps=60):
    """faster_wavenet.py incremental steps at config C on the host (rolled windows, ELU head over the whole window, H2D-free):
    time per step after the priming call."""
    from oracle import wavenet_oracle as O
    cfg = O.config_C()
    Win = O.input_width(cfg)
    ref = load_reference()
    audio = np.full((Win,), 127, dtype=np.int32)
    if ref is not None:
        R, RF, RD = ref
        np.random.seed(1234)
        net = RF.FasterWaveNet(_ref_params(R, cfg))
        fn = lambda a: net._forward_one_step(RD.onehot_pixel_image(a[-Win:].reshape(1, -1), 256), apply_softmax=True,
                                             as_numpy=True)[0, :, 0, -1]
        kind = "reference"
    else:
        lit = O.LiteralFastGenerator(cfg, O.init_weights(cfg, np.random.default_rng(1234), np.float32), np.float32)
        fn = lambda a: lit._forward_one_step(O.onehot_pixel_image(a[-Win:].reshape(1, -1), 256), apply_softmax=True)[0, :, 0, -1]
        kind = "port"
    t0 = time.perf_counter()
    p = fn(audio)
    audio = np.append(audio, [int(np.argmax(p))])
    prime_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(n_steps):
        p = fn(audio)
        audio = np.append(audio, [int(np.argmax(p))])
    per = (time.perf_counter() - t0) / n_steps
    return {"value": 1.0 / per, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": kind,
            "sample": "%d incremental _forward_one_step calls at config C after the priming call (%.2f s), batch 1; "
                      "16000 samples would take %.0f s at this rate (extrapolated)" % (n_steps, prime_s, 16000 * per),
            "s_per_sample": per}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B, W = 1, 8192
    tr = CpuTrainer("C", B, W, W)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tr.step()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = B * W / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic mu-law indices, random-init weights",
        "config": {"workload": "config C (30 layers d=1..512 x3, 64 residual / 256 skip, head 256-256-256) train step on a bounded "
                               "CPU sample of %d x %d samples per step (the GPU arm steps 32 x 16000 per GPU; samples/s normalises)"
                               % (B, W)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": tr.kind,
                         "sample": tr.describe("%d x %d samples per step" % (B, W))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16x2", choices=["fp16x2", "tf32", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="sequences per GPU")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--no-gen", action="store_true", help="skip the generation side metrics")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-modes", action="store_true", help="skip the other precision modes")
    ap.add_argument("--gen-steps", type=int, default=16000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from wavenet_b200 import _lib                      # the product arm never touches oracle/ (only the CPU legs do)
    from wavenet_b200.wavenet import WaveNet, _ptr, _stream
    from wavenet_b200.faster_wavenet import FasterWaveNet

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()

    B, W = args.batch, args.width
    params = config_c()
    net = FasterWaveNet(params, seed=1234)             # the facade's own initialiser (LeCunNormal, bias 0, wavenet.py:379-455)
    net.to_gpu(local_rank)
    net.set_precision(args.precision)
    if world > 1:
        from wavenet_b200.dist import init_comm
        init_comm(net)                                 # NCCL communicator inside libwavenet_b200.so (wn_comm_init)
    net.update_laerning_rate(1e-3)
    x_h, t_h = synth_batch(rank, B, W)
    x_d = torch.from_numpy(x_h).cuda()
    t_d = torch.from_numpy(t_h).cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    step = lambda: net.train_step(x_d, t_d)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.wn_launch_count(1)
    ms = timed(step, args.steps)
    launches = int(lib.wn_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    value = B * W * world / (ms / 1e3)
    tc = bool(lib.wn_tc_active(net._h))
    eff_prec = args.precision if tc else "fp32"

    # ---- e2e: host buffers through the public API -------------------------------------
    xp, tp = torch.from_numpy(x_h).pin_memory(), torch.from_numpy(t_h).pin_memory()
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    xs, ts = torch.empty_like(x_d), torch.empty_like(t_d)

    def e2e_step():
        xs.copy_(xp, non_blocking=True)
        ts.copy_(tp, non_blocking=True)
        loss = net.train_step(xs, ts)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss every step (train.py:83)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = B * W * world / (e2e_ms / 1e3)

    # ---- forward+loss only (BASELINE config 2 wording) and the dominant kernel ------------
    def fwd_only():
        _lib.check(lib.wn_forward_loss(net._h, _ptr(net._params), _ptr(x_d), _ptr(t_d), W, _ptr(net._loss), None,
                                       _stream()))
    fwd_only()
    fwd_ms = timed(fwd_only, args.steps)

    def residual_only():
        _lib.check(lib.wn_forward_residual_block(net._h, _ptr(net._params), None, None, None, _stream()))
    residual_only()
    res_ms = timed(residual_only, args.steps)
    peaks = measured_peaks()
    n_layers = 30

    def stored_traffic(name):
        """DRAM bytes per launch from a committed `ncu --set full` capture (profiles/): a STORED figure, not measured in this
        run -- the run itself measures `achieved` with CUDA events."""
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(tpath):
            with open(tpath) as f:
                return json.load(f).get("dram_bytes_per_launch"), "stored ncu capture profiles/" + name
        return None, None

    roofline_layer = roofline_tensor = None
    if eff_prec == "fp16x2":
        # dominant kernel family of the step: the fused residual-layer kernels; the forward one is timed alone here
        def layers_only():
            for l in range(n_layers):
                _lib.check(lib.wn_tcs_layer_forward(net._h, l, _stream()))
        layers_only()
        lay_ms = timed(layers_only, args.steps) / n_layers
        # split rows are 4 B per channel: read x(t) once (x(t-d) re-read hits L2), write x_out and z (split) and the 16-bit
        # fixed-point sigmoid tape
        bytes_per_pos = 3 * 64 * 4 + 64 * 2
        alg_bytes = bytes_per_pos * B * W
        achieved = alg_bytes / (lay_ms / 1e3) / 1e9
        traffic, tsrc = stored_traffic("r02_ncu_tcs_layer_kernel_final.json")
        layer_flops = 2 * (128 * 128 + 64 * 64) * B * W
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
                    "kernel": "tcs_layer_kernel (fused residual layer forward, fp16x2), %.1f us per launch, 30 per step"
                              % (1e3 * lay_ms),
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": "%s HBM copy bandwidth" % peaks["source"],
                    "tensor": {"achieved_tflops_algorithmic": layer_flops / (lay_ms / 1e3) / 1e12,
                               "executed_mma_tflops": 3 * layer_flops / (lay_ms / 1e3) / 1e12,
                               "peak_tflops": peaks["bf16_sustained"],
                               "note": "three kind::f16 MMAs per product (hi.hi + hi.lo + lo.hi); per-layer GEMMs (K=128/64) are "
                                       "HBM-bound, see DESIGN.md section 5"},
                    "residual_stack_forward_ms": res_ms}

        def skip_only():
            _lib.check(lib.wn_tcs_skip_gemm(net._h, _stream()))
        skip_only()
        skip_ms = timed(skip_only, args.steps)
        skip_flops = 2 * 64 * n_layers * 256 * B * W
        skip_tf = skip_flops / (skip_ms / 1e3) / 1e12
        roofline_tensor = {"bound": "tensor", "achieved": 3 * skip_tf, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                           "frac": 3 * skip_tf / peaks["bf16_burst"], "algorithmic_tflops": skip_tf,
                           "kernel": "tcs_gemm_kernel<256,4> skip sum (K=1920, N=256, K blocks flushed to registers), %.1f us per launch"
                                     % (1e3 * skip_ms),
                           "peak_source": "%s bf16 burst figure (cuBLAS); achieved counts the three executed fp16 MMAs per product"
                                          % peaks["source"]}
    elif eff_prec == "tf32":
        scratch = torch.zeros_like(net._grads)

        def gates_only():
            for l in range(n_layers - 1):
                _lib.check(lib.wn_tc_gate_backward_layer(net._h, l, _ptr(scratch), _stream()))
        gates_only()
        gate_ms = timed(gates_only, args.steps) / (n_layers - 1)
        gate_bytes = (3 * 64 * 4 + 64 * 2 + 128 * 4) * B * W    # read dout, dzs, z (fp32) + sigmoid (fp16); write dafg
        gate_achieved = gate_bytes / (gate_ms / 1e3) / 1e9

        def layers_only():
            for l in range(n_layers):
                _lib.check(lib.wn_tc_layer_forward(net._h, l, _stream()))
        layers_only()
        lay_ms = timed(layers_only, args.steps) / n_layers
        bytes_per_pos = 3 * 64 * 4 + 64 * 2
        alg_bytes = bytes_per_pos * B * W
        achieved = alg_bytes / (lay_ms / 1e3) / 1e9
        traffic, tsrc = stored_traffic("r01_ncu_tc_gate_bwd_kernel.json")
        gate_flops = 2 * (64 * 64 + 64 * 64) * B * W
        roofline = {"bound": "hbm", "achieved": gate_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gate_achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": tsrc,
                    "kernel": "tc_gate_bwd_kernel (dz GEMM + gate derivative + dWp, one residual layer), %.1f us per launch"
                              % (1e3 * gate_ms),
                    "algorithmic_bytes_per_launch": gate_bytes,
                    "peak_source": "%s HBM copy bandwidth" % peaks["source"],
                    "tensor": {"achieved_tflops": gate_flops / (gate_ms / 1e3) / 1e12, "peak_tflops": peaks["bf16_sustained"] / 2.0}}
        ltraffic, ltsrc = stored_traffic("r01_ncu_tc_layer_kernel.json")
        roofline_layer = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": achieved / peaks["hbm_gbs"], "traffic": ltraffic, "traffic_source": ltsrc,
                          "kernel": "tc_layer_kernel (fused residual layer forward), %.1f us per launch" % (1e3 * lay_ms),
                          "algorithmic_bytes_per_launch": alg_bytes, "residual_stack_forward_ms": res_ms}

        def skip_only():
            _lib.check(lib.wn_tc_skip_gemm(net._h, _stream()))
        skip_only()
        skip_ms = timed(skip_only, args.steps)
        skip_tf = 2 * 64 * n_layers * 256 * B * W / (skip_ms / 1e3) / 1e12
        roofline_tensor = {"bound": "tensor", "achieved": skip_tf, "peak": peaks["bf16_burst"] / 2.0, "unit": "TFLOP/s",
                           "frac": skip_tf / (peaks["bf16_burst"] / 2.0),
                           "kernel": "tc_gemm_kernel<256> skip sum (K=1920, N=256), %.1f us per launch" % (1e3 * skip_ms),
                           "peak_source": "kind::tf32 taken as half of the %s bf16 burst figure (cuBLAS)" % peaks["source"]}
    else:
        achieved = LAYER_FLOP_PER_POS * n_layers * B * W / (res_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": 72.0, "unit": "TFLOP/s", "frac": achieved / 72.0,
                    "traffic": None, "kernel": "residual stack forward, fp32 SIMT (%.3f ms per pass)" % res_ms,
                    "peak_source": "fp32 SIMT nominal 72 TF/s (no tensor pipe in use)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAMES[eff_prec],
        "data": "synthetic mu-law indices default_rng(rank), random-init LeCunNormal weights (facade initialiser, seed 1234)",
        "config": {"workload": "config C (30 layers d=1..512 x3, 64 residual / 256 skip, head 256-256-256), "
                               "%d x %d samples per GPU, full-width teacher-forced train step" % (B, W),
                   "global_batch": B * world, "width": W, "parallelism": "dp%d" % world, "precision_mode": eff_prec,
                   "allreduce": (None if world == 1 else ("fused one-shot peer-memory all-reduce + norm (wn_allreduce_clip_adam_step)"
                                                          if lib.wn_comm_peer_enabled(net._h) else "ncclAllReduce behind the C ABI")),
                   "l2": "working set (activation tape ~20 GB) is far larger than the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(x_h.nbytes + t_h.nbytes),
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "roofline_layer": roofline_layer, "roofline_tensor": roofline_tensor,
        "train_tflops": 3 * FWD_FLOP_PER_POS * B * W * world / (ms / 1e3) / 1e12,
        "fwd_loss": {"ms": fwd_ms, "samples_per_s": B * W * world / (fwd_ms / 1e3),
                     "tflops": FWD_FLOP_PER_POS * B * W * world / (fwd_ms / 1e3) / 1e12},
    }

    # ---- strong scaling (BASELINE config 5): GLOBAL batch 256 x 16000 at every N -- 256 / (32 N) micro-batches of 32 per rank
    # with gradient accumulation, then ONE all-reduce + clip + Adam; the all-reduce is amortised over 8 micro-batches at N = 1
    # and over one at N = 8 ---------------------------------------------------------------------------------------------
    if B * world <= 256 and 256 % (B * world) == 0 and W == WIDTH:
        k = 256 // (B * world)
        mbs = [(x_d, t_d)] * k
        net.train_step_accumulated(mbs)
        sms = timed(lambda: net.train_step_accumulated(mbs), 2)
        line["strong_scaling"] = {"global_batch": 256, "width": W, "micro_batches_per_rank": k, "ms_per_step": sms,
                                  "samples_per_s": 256 * W / (sms / 1e3), "scaling": "strong",
                                  "note": "one optimiser step = k x (forward + loss + backward) + gradient accumulation + "
                                          "one all-reduce + clip + Adam"}

    # ---- the other precision modes on the same step (N == 1): parity class next to speed -------------------
    if world == 1 and not args.no_modes:
        modes = {eff_prec: {"ms_per_step": ms, "samples_per_s": value}}
        for other, nsteps in (("fp16x2", args.steps), ("tf32", args.steps), ("fp32", 2)):
            if other in modes:
                continue
            net.set_precision(other)
            for _ in range(2):
                step()
            oms = timed(step, nsteps)
            modes[other] = {"ms_per_step": oms, "samples_per_s": B * W / (oms / 1e3)}
        net.set_precision(args.precision)
        modes["fp16x2"]["parity"] = "logits <= 1e-4, gradients <= 1e-3 vs the reference (tests/test_gpu_fp16x2.py, test_gpu_reference.py)"
        modes["tf32"]["parity"] = "logits <= 1e-2 with argmax agreement; gradients ~4e-2 (single-pass tensor cores)"
        modes["fp32"]["parity"] = "exact fp32 SIMT FFMA: logits <= 1e-4, gradients <= 1e-3"
        line["precision_modes"] = modes

    # ---- generation side metrics (BASELINE configs 3 and 4), N streams sharded, no collective ---
    if not args.no_gen:
        gen = {}
        Win = net.input_width
        # batch_256 shards 256 streams over the ranks (BASELINE config 4, strong scaling: 256 / N streams per GPU);
        # batch_256_per_gpu (N > 1 only) keeps 256 streams on EVERY GPU (weak scaling, the units-per-rank-fixed reading of 5)
        cases = [("batch_1", 1, 1), ("batch_256", 256, max(1, 256 // world))]
        if world > 1:
            cases.append(("batch_256_per_gpu", 256 * world, 256))
        # the serving-scale cases on the tensor-core generator (gen_kernel_v6), weak-scaled: as many streams per GPU as stay
        # resident with 8-CTA clusters (15 x 128 = 1920 on a B200) and with 4-CTA clusters (33 x 128 = 4224)
        cap8, cap4 = int(lib.wn_gen_mma_capacity(8)), int(lib.wn_gen_mma_capacity(4))
        if cap8 > 296:
            cases.append(("many_streams_8cta_per_gpu", cap8 * world, cap8))
        if cap4 > cap8:
            cases.append(("many_streams_4cta_per_gpu", cap4 * world, cap4))
        for key, n_total, n in cases:
            if n_total == 1 and rank != 0:
                continue
            if n_total == 1:
                window = np.full((1, Win), 127, dtype=np.int32)                       # generate.py:21
            else:
                window = np.random.default_rng(0).integers(0, 256, (n_total, Win)).astype(np.int32)[rank * n:(rank + 1) * n]
            steps = args.gen_steps
            net.prime(window)
            out = torch.empty((n, steps), dtype=torch.int32, device="cuda")

            def run():
                _lib.check(lib.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_SAMPLE, 0, _ptr(out), _stream()))
            net.prime(window)
            run()
            torch.cuda.synchronize()
            net.prime(window)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            gms = e0.elapsed_time(e1)
            if world > 1 and n_total > 1:
                t = torch.tensor([gms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                gms = float(t[0])
            # end to end through the public call: host window -> device, priming pass, the sampling loop, samples -> host
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            host_samples = net.generate(window, steps, mode="sample", seed=0).cpu()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - w0
            if world > 1 and n_total > 1:
                t = torch.tensor([e2e_s], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t[0])
            del host_samples
            total_streams = n * world if n_total > 1 else 1
            us = 1e3 * gms / steps
            sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
            clustered = n <= 15          # gen_kernel_v4: one 8-CTA cluster per stream while all clusters are co-resident
            many = False                 # gen_kernel_v5 (16 streams per 8-CTA cluster) is opt-in (WN_GEN_V5=1): slower than v3 today
            mma = n > 296 and os.environ.get("WN_GEN_V6", "1") != "0"   # gen_kernel_v6: automatic beyond 2 x sm_count = 296 streams
            mma_cs = 8 if (mma and n <= cap8) else 4                    # 8-CTA clusters while they are all resident, else 4-CTA
            n_ctas = 8 * n if clustered else (mma_cs * -(-n // 128) if mma else (8 * -(-n // 16) if many else -(-n // (1 if n <= 148 else 2))))
            per_cta = (8 // mma_cs) * 32 * 32768 if mma else (33 * 32768 if (clustered or many) else 1270272 * 4)   # packed weight bytes per CTA per step
            gen[key] = {
                "samples_per_s": total_streams * steps / (gms / 1e3), "streams": total_streams, "streams_per_gpu": n,
                "scaling": "weak" if key.endswith("_per_gpu") else ("strong" if n_total > 1 else "single stream"), "steps": steps,
                "us_per_step": us, "cycles_per_sample_per_stream": us * sm_mhz,
                "e2e": {"samples_per_s": total_streams * steps / e2e_s, "seconds": e2e_s,
                        "what": "FasterWaveNet.generate(host window): H2D of the window, priming pass, sampling loop, D2H of the samples",
                        "h2d_bytes": int(window.nbytes), "d2h_bytes": int(n * steps * 4)},
                "kernel": "gen_kernel_v4 (8-CTA cluster per stream, output-split matvecs, st.async exchanges)" if clustered
                          else ("gen_kernel_v6<%d> (tcgen05: streams = MMA M dimension, 128 per %d-CTA cluster, fp16 hi|lo operands, "
                                "weights split by output rows, z all-gathered with bulk DSMEM copies)" % (mma_cs, mma_cs) if mma else ("gen_kernel_v5 (16 streams per 8-CTA cluster: weights in registers swept over the streams, "
                                "transposing-butterfly reductions, st.async exchanges)" if many
                                else "gen_kernel_v3 (one CTA per 1-2 streams)")),
                "weight_stream_gbs_per_sm": per_cta / (us * 1e-6) / 1e9,
                "weight_stream_gbs_all_ctas": n_ctas * world * per_cta / (us * 1e-6) / 1e9,
                "bound": ("dependency-chain latency (30 layers x 1 cluster exchange + head per sample)" if clustered else
                          "per layer: DSMEM all-gather of z (24-28 KB into every CTA at ~17-21 B/clk) + chains of tiny MMAs that run at the MMA "
                          "latency (55-75 cycles each) + TMEM reads of the epilogues (64 B/clk), all on one dependency chain per cluster" if mma else
                          ("FP32 FMA issue + one cluster exchange per layer (per CTA and step: 16 streams x 155 k MACs)" if many
                           else "dependency-chain latency (30 layers x 2 block barriers + head per sample)"))}
        if rank == 0 and world == 1 and not args.no_cpu:
            gen["cpu_baseline"] = cpu_generation_sample()
            gen["batch_1"]["vs_cpu_reference"] = gen["batch_1"]["samples_per_s"] / gen["cpu_baseline"]["value"]
        line["fast_gen"] = gen

    # ---- BASELINE config 1 shape: reference default network (train_audio/model.py:24-43), one train.py step on a 1 s 16 kHz
    # clip, batch 1 -- on the GPU and (CPU leg) through the reference's own code on the host ----
    if rank == 0 and world == 1:
        netb = WaveNet(config_b(), seed=0)
        netb.to_gpu(local_rank)
        netb.set_precision(args.precision)
        netb.update_laerning_rate(1e-3)
        xb_h, tb_h = synth_batch(0, 1, 16000)
        tw = 16000 - 257
        xb = torch.from_numpy(xb_h).cuda()
        tb = torch.from_numpy(np.ascontiguousarray(tb_h[:, 16000 - tw:])).cuda()
        stepb = lambda: netb.train_step(xb, tb, train_width=tw)
        for _ in range(3):
            stepb()
        msb = timed(stepb, 10)
        c1 = {"workload": "R256/G128 x 8 layers, batch 1 x 16000, train_width 15743, fwd+bwd+clip+Adam",
              "ms_per_step": msb, "samples_per_s": 16000 / (msb / 1e3), "tc_active": bool(lib.wn_tc_active(netb._h)),
              "precision_mode": args.precision, "tflops": 3 * 3276800 * 16000 / (msb / 1e3) / 1e12}
        del netb
        if not args.no_cpu:
            trb = CpuTrainer("B", 1, 16000, tw)
            trb.step()
            t0 = time.perf_counter()
            trb.step()
            sb = time.perf_counter() - t0
            c1["cpu_baseline"] = {"value": 16000 / sb, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": trb.kind,
                                  "sample": trb.describe("one full config-1 step (1 x 16000, train_width 15743) after one warm-up, %.1f s" % sb)}
            c1["vs_cpu_reference"] = c1["samples_per_s"] / c1["cpu_baseline"]["value"]
        line["config1_reference_default_net"] = c1

    # ---- CPU baseline of the headline workload (rank 0, N == 1 only) -----------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        tr = CpuTrainer("C", 1, 8192, 8192)
        tr.step()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            tr.step()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        line["cpu_baseline"] = {"value": 8192 / best, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": tr.kind,
                                "sample": tr.describe("best of 2 config-C train steps on 1 x 8192 samples (%.1f s each) after one warm-up" % best)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
