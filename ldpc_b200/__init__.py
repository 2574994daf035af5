"""ldpc_b200 -- B200-native batched belief-propagation decoding behind the quantumgizmos/ldpc API.

``BpDecoder`` / ``BpOsdDecoder`` keep the reference package's constructor arguments and ``.decode(syndrome)``
(reference src_python/ldpc/bp_decoder, src_python/ldpc/bposd_decoder) and add ``.decode_batch``.  The work is
done by hand-written sm_100a CUDA kernels in ``ldpc_b200/csrc`` behind the C-ABI of ``include/bp_b200.h``.
"""
from .bp_decoder import BpDecoder, BpDecoderBase, SoftInfoBpDecoder, io_test
from .bposd_decoder import BpOsdDecoder
from . import codes
from .monte_carlo import MonteCarloBscSimulation
from .legacy import bp_decoder, bposd_decoder

__all__ = ["BpDecoder", "BpDecoderBase", "SoftInfoBpDecoder", "BpOsdDecoder", "MonteCarloBscSimulation", "bp_decoder", "bposd_decoder",
           "io_test", "codes"]
__version__ = "0.1.0"
