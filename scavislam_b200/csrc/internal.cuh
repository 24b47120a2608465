// internal.cuh -- cross-module hooks inside libsvsb200.so (not part of the C ABI)
#pragma once
#include "../../include/svs_b200.h"

namespace svs {
// device-resident results of the last svs_match on this handle (n = number of candidate points)
__attribute__((visibility("hidden"))) void matcher_device_results(svs_matcher* m, const svs_match_result** d_res, int* n,
                                                                   int* device);
}  // namespace svs
