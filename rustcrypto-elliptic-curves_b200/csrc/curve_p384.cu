// curve_p384.cu — instantiates every batch kernel for p384 (one translation unit per curve so the
// curves compile in parallel).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_p384() { return Launch<CurveP384>::table(); }
}  // namespace ecb
