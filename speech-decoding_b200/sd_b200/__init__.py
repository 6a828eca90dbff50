"""sd_b200 -- host runtime of the B200-native BrainEncoder + CLIPLoss hot path.

The public, reference-facing API is the drop-in package `speech_decoding`
(speech_decoding.models / speech_decoding.utils.loss) next to this one; this
package holds the ctypes binding of libsd_b200.so and the executor."""
from .ops import set_precision, get_precision, set_impl  # noqa: F401
