// svs_nvtx.hpp -- NVTX ranges on the C-ABI entry points, named after the reference's own timers:
// VisionTools::PerformanceMonitor "dense tracking" | "fast" | "match" | "dense point cloud" (stereo_frontend.cpp:118-303)
// and the StopWatch around optimizer.optimize (slam_graph.cpp:344-352); "copyDataToG2o" / "restoreDataFromG2o" are the
// functions of slam_graph.cpp:907-1058 the problem set-up and the write-back replace.  NVTX 3 is header-only and a
// no-op (one pointer test) unless a tool is attached.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace svs {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace svs
