// Instantiates the min-sum paired on-chip kernels (bp_pair.cuh) for every degree bucket.
#include "bp_pair.cuh"
namespace bpb {
PairKernel pick_pair_ms(int dc, int dv, bool regular, bool llr, int cta_threads) {
    return pick_pair_bucket<kMinimumSum>(dc, dv, regular, llr, cta_threads);
}
}  // namespace bpb
