"""train_audio/train.py of the reference (14-139), Python 3, driving the B200 backend.

The loop, batch construction and logging are the reference's; the one-hot image
(train.py:62) is skipped because the backend consumes the int32 samples directly
(pass the one-hot tensor instead and it is accepted all the same)."""
import os
import sys
import time

import numpy as np

sys.path.append(os.path.split(os.path.dirname(os.path.abspath(__file__)))[0])
from args import args  # noqa: E402
from model import params, wavenet  # noqa: E402
from wavenet_b200 import data  # noqa: E402


class stdout:
    BOLD = "\033[1m"
    END = "\033[0m"
    CLEAR = "\033[2K"


def create_batch(signal, batch_size, input_width, target_width):
    """train.py:14-22."""
    indecis = np.random.randint(0, signal.size - target_width - input_width - 1, size=batch_size)
    input_batch = np.empty((batch_size, input_width + target_width), dtype=np.int32)
    target_batch = np.empty((batch_size, target_width), dtype=np.int32)
    for n in range(batch_size):
        start = indecis[n]
        input_batch[n] = signal[start:start + input_width + target_width]
        target_batch[n] = signal[start + input_width + 1:start + input_width + target_width + 1]
    return input_batch, target_batch


def train_audio(filename, batch_size=16, train_width=16, repeat=1000):
    """train.py:24-91."""
    path_to_file = args.wav_dir + "/" + filename
    signals, sampling_rate = data.load_audio_file(path_to_file, quantization_steps=params.quantization_steps)

    num_layers = len(params.residual_conv_channels)
    receptive_width_per_unit = params.residual_conv_filter_width ** num_layers
    receptive_width = (receptive_width_per_unit - 1) * params.residual_num_blocks + 1
    input_width = receptive_width
    input_width += len(params.causal_conv_channels)

    sum_loss = 0
    signals = np.insert(signals, 0, np.full((input_width,), 127, dtype=np.int32), axis=0)

    import torch
    signals_dev = torch.from_numpy(np.ascontiguousarray(signals, dtype=np.int32)).to(wavenet._device)   # resident for the whole file
    for batch_index in range(0, repeat):
        # same np.random stream as create_batch (train.py:15); the gather itself runs on the device
        indecis = np.random.randint(0, signals.size - train_width - input_width - 1, size=batch_size)
        input_batch, target_batch = wavenet.create_batch(signals_dev, indecis, input_width, train_width)
        output = wavenet.forward_causal_block(input_batch)
        output, sum_skip_connections = wavenet.forward_residual_block(output)
        output = wavenet.slice_1d(output, output.data.shape[3] - train_width)
        sum_skip_connections = wavenet.slice_1d(sum_skip_connections, sum_skip_connections.data.shape[3] - train_width)
        output = wavenet.forward_softmax_block(sum_skip_connections, apply_softmax=False)
        loss = wavenet.cross_entropy(output, target_batch)
        wavenet.backprop(loss)

        sum_loss += float(loss.data)
        if batch_index % 10 == 0:
            sys.stdout.write("\r	{} - {} width; {}/{}".format(stdout.BOLD + filename + stdout.END, signals.size, batch_index, repeat))
            sys.stdout.flush()

    wavenet.save(args.model_dir)
    return sum_loss


def main():
    np.random.seed(args.seed)
    wavenet.update_laerning_rate(args.lr)

    files = []
    for fn in os.listdir(args.wav_dir):
        if fn.endswith(".wav"):
            print("loading", fn)
            files.append(fn)

    num_layers = len(params.residual_conv_channels)
    receptive_width_per_unit = params.residual_conv_filter_width ** num_layers
    receptive_width = (receptive_width_per_unit - 1) * params.residual_num_blocks + 1
    receptive_msec = int(receptive_width * 1000.0 / params.sampling_rate)
    print("receptive field width:", receptive_msec, "[millisecond]")
    print("receptive field width:", receptive_width, "[step]")

    batch_size = 16
    train_width = 500
    max_epoch = 2000
    start_time = time.time()
    print("files: {} batch_size: {} train_width: {}".format(len(files), batch_size, train_width))

    for epoch in range(1, max_epoch):
        average_loss = 0
        for i, filename in enumerate(files):
            loss = train_audio(filename, batch_size=batch_size, train_width=train_width, repeat=500)
            average_loss += loss
        average_loss /= len(files)
        sys.stdout.write(stdout.CLEAR)
        sys.stdout.write("\repoch: {} - {:.4e} loss - {} min\n".format(epoch, average_loss, int((time.time() - start_time) / 60)))
        sys.stdout.flush()
        wavenet.save(args.model_dir)


if __name__ == "__main__":
    main()
