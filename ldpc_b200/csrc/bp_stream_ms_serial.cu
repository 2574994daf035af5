// Instantiates the ms / serial stream kernels (bp_stream.cuh) for every degree bucket.
#include "bp_stream.cuh"
namespace bpb {
StreamKernel pick_stream_ms_serial(int dc, int dv, bool regular, bool llr) {
    return pick_stream_bucket<kMinimumSum, kSerial>(dc, dv, regular, llr);
}
}  // namespace bpb
