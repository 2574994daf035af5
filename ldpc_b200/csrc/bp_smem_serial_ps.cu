// Instantiates the product-sum on-chip SERIAL-schedule kernels (bp_smem_serial.cuh).
#include "bp_smem_serial.cuh"
namespace bpb {
SmemKernel pick_smem_serial_ps(int dc, int dv, bool regular, bool llr) {
    return pick_smem_serial_bucket<kProductSum>(dc, dv, regular, llr);
}
}  // namespace bpb
