// Instantiates every kernel of plans.cuh so that nvcc / ptxas see them (sm_100a). Never linked.
#include "plans.cuh"
#include "tile2csr_v2.cuh"
#include "rowplans.cuh"
template __global__ void plans::k_plan_build<false>(int, const int *, const int *, const int *, const int *, const int *, const uint16_t *,
                                                    const uint16_t *, const uint16_t *, const uint16_t *, uint16_t *, uint16_t *, int *, int *,
                                                    const int *, unsigned *, uint8_t *, unsigned *);
template __global__ void plans::k_plan_build<true>(int, const int *, const int *, const int *, const int *, const int *, const uint16_t *,
                                                   const uint16_t *, const uint16_t *, const uint16_t *, uint16_t *, uint16_t *, int *, int *,
                                                   const int *, unsigned *, uint8_t *, unsigned *);
