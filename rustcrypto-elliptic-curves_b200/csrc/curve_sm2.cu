// curve_sm2.cu — instantiates every batch kernel for sm2 (one translation unit per curve so the
// curves compile in parallel).
#include "kernels_impl.cuh"

namespace ecb {
const CurveLaunch* launch_sm2() { return Launch<CurveSM2>::table(); }
}  // namespace ecb
