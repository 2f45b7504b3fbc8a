// curve_k256.cu — instantiates every batch kernel for k256 (one translation unit per curve so the
// curves compile in parallel).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_k256() { return Launch<CurveK256>::table(); }
}  // namespace ecb
