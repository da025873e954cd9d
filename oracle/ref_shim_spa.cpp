/*
 * oracle/ref_shim_spa.cpp -- extern "C" wrapper around the reference's value-bearing serial SPA
 * (src/external/cusparse/spgemm_serialref_spa.h:33), compiled in place from /root/reference.
 * Separate translation unit because that header includes external/cusparse/common.h, which
 * clashes with src/common.h. TEST INFRASTRUCTURE ONLY (see ref_shim.cpp).
 */
#include <stdbool.h>
#include "common.h"
#include "utils.h"
#include "spgemm_serialref_spa.h"

extern "C" int ref_spgemm_serialref(const int *rpA, const int *ciA, const double *vA, int mA, int nA, int nnzA,
                                    const int *rpB, const int *ciB, const double *vB, int mB, int nB, int nnzB,
                                    int *rpC, int *ciC, double *vC, int mC, int nC, int *nnzC, int get_nnzC_only)
{
    return spgemm_serialref(rpA, ciA, vA, mA, nA, nnzA, rpB, ciB, vB, mB, nB, nnzB,
                            rpC, ciC, vC, mC, nC, nnzC, get_nnzC_only != 0);
}
