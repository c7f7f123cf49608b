// b200q — one explicit instantiation of the register-tiled kernel (see rtile_launch.cuh).
#include "rtile_launch.cuh"

namespace b200q {
template int rtile_launch<double, 3, 2, 256, 2, false>(RT_LAUNCH_ARGS);
}  // namespace b200q
