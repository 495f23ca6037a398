"""train_audio/generate.py of the reference (9-66), Python 3, driving the B200 backend.

Without --fast the reference's per-sample loop is kept verbatim (slow path: a full
teacher-forced pass per sample, generate.py:35).  With --fast / --use_faster_wavenet the
whole loop of generate.py:24-43 -- incremental forward, softmax, categorical draw, append --
runs inside one persistent CUDA kernel (wn_gen_run) and only the finished samples come back."""
import os
import sys
import time

import numpy as np

sys.path.append(os.path.split(os.path.dirname(os.path.abspath(__file__)))[0])
from args import args  # noqa: E402
from model import params, wavenet  # noqa: E402
from wavenet_b200 import data  # noqa: E402


def generate_audio(sampling_rate=48000, generate_sec=1, remove_silence_frames=False):
    num_layers = len(params.residual_conv_channels)
    receptive_steps_per_unit = params.residual_conv_filter_width ** num_layers
    receptive_steps = (receptive_steps_per_unit - 1) * params.residual_num_blocks + 1
    input_width = receptive_steps
    input_width += len(params.causal_conv_channels)

    generated_quantized_audio = np.full((input_width,), 127, dtype=np.int32)
    n_steps = int(sampling_rate * generate_sec) - 1

    start_time = time.time()
    if args.fast and not remove_silence_frames:
        window = generated_quantized_audio.reshape((1, -1))
        mode = "greedy" if args.greedy else "sample"
        seed = 0 if args.seed is None else args.seed
        samples = wavenet.generate(window, n_steps, mode=mode, seed=seed).cpu().numpy()[0]
        generated_quantized_audio = np.append(generated_quantized_audio, samples, axis=0)
    else:
        # per-sample loop of generate.py:24-43.  remove_silence_frames drops a generated 127 from the signal, so the next
        # call sees the SAME window (and the fast path feeds its last sample again, generate.py:40-43 + faster_wavenet.py:
        # 65-78): that data-dependent feedback stays on the host loop, one incremental step per call.
        step_fn = wavenet._forward_one_step if args.fast else wavenet.forward_one_step
        for time_step in range(1, n_steps + 1):
            padded_quantized_x_batch = generated_quantized_audio[-input_width:].reshape((1, -1))
            softmax = step_fn(padded_quantized_x_batch, apply_softmax=True, as_numpy=True)
            softmax = softmax[0, :, 0, -1].astype(np.float64)
            softmax /= softmax.sum()
            if args.greedy:
                generated_quantized_signal = int(np.argmax(softmax))
            else:
                generated_quantized_signal = np.random.choice(np.arange(params.quantization_steps), p=softmax)
            if generated_quantized_signal == 127 and remove_silence_frames:
                pass
            else:
                generated_quantized_audio = np.append(generated_quantized_audio, [generated_quantized_signal], axis=0)
            if time_step % 10 == 0:
                sys.stdout.write("\rgenerating {:.2f} msec / {:.2f} msec".format(time_step * 1000.0 / sampling_rate, generate_sec * 1000.0))
                sys.stdout.flush()

    print("\ndone in {:.3f} sec".format(time.time() - start_time))

    generated_quantized_audio = generated_quantized_audio[input_width:]
    try:
        os.mkdir(args.output_dir)
    except Exception:
        pass
    filename = "{}/generated.wav".format(args.output_dir)
    data.save_audio_file(filename, generated_quantized_audio, params.quantization_steps, format="16bit_pcm", sampling_rate=sampling_rate)


def main():
    np.random.seed(args.seed)
    generate_audio(generate_sec=args.seconds, sampling_rate=params.sampling_rate)


if __name__ == "__main__":
    main()
