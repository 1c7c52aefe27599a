"""B200-native (sm_100a) Perceiver IO attention stack: a drop-in for the forward path of
`perceiver_io/transformer_primitives.py` and the encoder / decoder of `perceiver_io/perceiver.py`
in JOBR0/PerceiverIO_Pytorch.  See DESIGN.md and INTEGRATION.md."""
from .primitives import Attention, CrossAttention, MLP, SelfAttention, make_cross_attention_mask  # noqa: F401
from .perceiver import PerceiverDecoder, PerceiverEncoder, TrainablePositionEncoding  # noqa: F401
from .engine import set_precision  # noqa: F401
from .inputs import PositionedInput, fourier_position_table, positioned_image_input, perceiver_io_forward  # noqa: F401

__all__ = ["Attention", "MLP", "SelfAttention", "CrossAttention", "make_cross_attention_mask",
           "PerceiverEncoder", "PerceiverDecoder", "TrainablePositionEncoding", "set_precision",
           "PositionedInput", "fourier_position_table", "positioned_image_input", "perceiver_io_forward"]
