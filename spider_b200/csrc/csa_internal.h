// Internal helpers shared by the translation units of libcsa_b200.so (error reporting, driver entry points).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/csa_b200.h"

namespace csa {

int set_error(int code, const char* fmt, ...);

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// device pointer to 4 words of host-mapped memory used by the barrier watchdog (nullptr if it cannot be allocated)
uint32_t* debug_record_devptr();

// number of SMs if `device` is compute capability 10.x, else 0 (cached per device)
int sm_count(int device);
// programmatic dependent launch between the library's kernels (CSA_PDL=0 turns it off: A/B knob)
bool pdl_enabled();

}  // namespace csa
