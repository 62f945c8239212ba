"""cobevt_b200: B200-native (sm_100a) implementation of CoBEVT's fused-axial-attention hot path.

Public surface = the reference's nn.Module classes (see modules.py) + the C ABI in
include/cobevt_b200.h.  Importing this package does not touch the GPU; the first kernel call
loads cobevt_b200/libcobevt_sm100.so and fails loudly if it is missing.
"""
from .modules import (  # noqa: F401
    ResnetEncoder, BEVEmbedding, Attention, CrossWinAttention, CrossViewSwapAttention, FAXModule,
    SwapFusionEncoder, SwapFusionBlock, SwapFusionBlockMask, NaiveDecoder, BevSegHead, STTF,
    CorpBEVT, FaxFusedTransformer, Bottleneck, NaiveCompressor, post_process, set_compute_dtype,
)

__version__ = "0.1.0"
