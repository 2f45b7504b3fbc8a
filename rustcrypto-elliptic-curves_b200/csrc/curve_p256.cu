// curve_p256.cu — instantiates every batch kernel for p256 (one translation unit per curve so the
// curves compile in parallel).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_p256() { return Launch<CurveP256>::table(); }
}  // namespace ecb
