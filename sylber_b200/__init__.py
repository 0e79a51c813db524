"""sylber_b200 - B200-native (sm_100a) implementation of SYLBER's `Segmenter` forward path.

Drop-in for `sylber.Segmenter` (reference: sylber/model/sylber.py:28-138): same constructor, same
`__call__(wav_file=None, wav=None, in_second=True)`, same `{segments, segment_features, hidden_states}`
output contract.  All arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
include/sylber_b200.h (libsylber_b200.so); PyTorch is used for device memory, streams and
torch.distributed only.  There is no CPU fallback.
"""
from .segmenter import Segmenter, SpeechModel, KMeansQuantizer  # noqa: F401
from .thresholder import Thresholder  # noqa: F401
from .batching import plan_length_buckets  # noqa: F401
from .distributed import segment_sharded  # noqa: F401
from ._lib import build_library, load_library, library_path  # noqa: F401

__all__ = ["Segmenter", "SpeechModel", "KMeansQuantizer", "Thresholder", "plan_length_buckets", "segment_sharded", "build_library",
           "load_library", "library_path"]
__version__ = "0.1.0"
