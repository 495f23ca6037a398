"""wavenet_b200 -- B200-native (sm_100a) backend for musyoku/wavenet's training stack
and incremental generator.  The CUDA library is mandatory; nothing here computes on the CPU."""
from . import _lib  # noqa: F401
from .wavenet import WaveNet, Params, Variable  # noqa: F401
from .faster_wavenet import FasterWaveNet  # noqa: F401
