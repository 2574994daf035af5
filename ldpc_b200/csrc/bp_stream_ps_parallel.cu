// Instantiates the ps / parallel stream kernels (bp_stream.cuh) for every degree bucket.
#include "bp_stream.cuh"
namespace bpb {
StreamKernel pick_stream_ps_parallel(int dc, int dv, bool regular, bool llr) {
    return pick_stream_bucket<kProductSum, kParallel>(dc, dv, regular, llr);
}
}  // namespace bpb
