// b200q — register-tiled kernels, complex64 instantiations (see rtile_launch.cuh).
#include "rtile_launch.cuh"

namespace b200q {

int rtile_run_c64(bool ws, void* v0, void* v1, const RtArgs& a, int64_t batch, const RtOp* od,
                   const double2* md, int nslots, double scale, double* out_dev, double* partials,
                   size_t pcap, cudaStream_t s) {
  return rtile_run<float>(ws, v0, v1, a, batch, od, md, nslots, scale, out_dev, partials, pcap, s);
}

}  // namespace b200q
