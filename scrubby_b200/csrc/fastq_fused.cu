// fastq_fused.cu -- single-pass fused parse -> probe -> compact kernel (placeholder until built).
#include "fastq_records.cuh"

namespace sgpu {
sgpu_status clean_fused(sgpu_ctx *, const sgpu_idset *, const uint8_t *, size_t, int, uint8_t *, size_t, size_t *,
                        uint8_t *, size_t, size_t *, sgpu_counts *, int *used) {
    *used = 0;
    return SGPU_OK;
}
}  // namespace sgpu
