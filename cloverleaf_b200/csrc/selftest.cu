// selftest.cu -- device self-test of the branch-free fp64 div / rcp / sqrt (common.cuh, Math<false>)
// against nvcc's own IEEE operators, on pseudo-random and adversarial operands.  Exported for the
// test-suite (tests/test_gpu_kernels.py::test_fast_math_matches_ieee); not on the hydro path.
#include "clover_b200.h"
#include "common.cuh"

namespace clv {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// operand family by `kind`: wide exponent range, hydro-like range, adversarial mantissas, zeros
__device__ __forceinline__ double make_operand(unsigned long long h, int kind) {
  unsigned long long mant = h & 0x000fffffffffffffull;
  const unsigned long long sign = (h >> 63) << 63;
  int expo;
  switch (kind & 3) {
    case 0: expo = (int)((h >> 52) & 0x7ff) % 1200 - 600; break;   // 2^-600 .. 2^600
    case 1: expo = (int)((h >> 52) & 0x3f) - 32; break;            // 2^-32 .. 2^31
    case 2: {                                                       // mantissa corner cases
      const int sel = (int)((h >> 52) & 7);
      const unsigned long long pats[8] = {0ull, 1ull, 0x000fffffffffffffull, 0x000ffffffffffffeull,
                                          0x0008000000000000ull, 0x0007ffffffffffffull,
                                          0x0005555555555555ull, 0x000aaaaaaaaaaaaaull};
      mant = pats[sel];
      expo = (int)((h >> 56) & 0x1f) - 16;
      break;
    }
    default: expo = (int)((h >> 52) & 0xf) - 8; break;
  }
  double v = __longlong_as_double((long long)(sign | ((unsigned long long)(expo + 1023) << 52) | mant));
  if ((kind & 3) == 3 && ((h >> 60) & 7) == 0) v = __longlong_as_double((long long)sign);  // +-0
  return v;
}

__global__ void selftest_math_kernel(unsigned long long n, unsigned long long seed, unsigned long long* out) {
  unsigned long long mism = 0, flagged = 0, checked = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long h1 = mix64(seed + 2 * i), h2 = mix64(seed + 2 * i + 1);
    const int kind = (int)(h1 >> 7);
    const double a = make_operand(h1, kind), b = make_operand(h2, kind >> 2);
    bool bad = false, dummy = false;
    const double qf = Math<false>::div(a, b, bad);
    const double qs = a / b;
    if (bad) ++flagged;
    else { ++checked; if (__double_as_longlong(qf) != __double_as_longlong(qs)) ++mism; }
    bad = false;
    const double rf = Math<false>::rcp(b, bad);
    const double rs = Math<true>::rcp(b, dummy);
    if (bad) ++flagged;
    else { ++checked; if (__double_as_longlong(rf) != __double_as_longlong(rs)) ++mism; }
    bad = false;
    const double pa = fabs(a);
    const double sf = Math<false>::sqrt(pa, bad);
    const double ss = Math<true>::sqrt(pa, dummy);
    if (bad) ++flagged;
    else { ++checked; if (__double_as_longlong(sf) != __double_as_longlong(ss)) ++mism; }
  }
  atomicAdd(&out[0], mism);
  atomicAdd(&out[1], flagged);
  atomicAdd(&out[2], checked);
}

}  // namespace clv

using namespace clv;

extern "C" void clover_b200_selftest_math_(long long* n, long long* seed, long long* mismatches,
                                           long long* flagged, long long* checked) {
  ensure_init();
  unsigned long long* d = nullptr;
  CLV_CUDA(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
  CLV_CUDA(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), stream()));
  selftest_math_kernel<<<148 * 8, 256, 0, stream()>>>((unsigned long long)*n, (unsigned long long)*seed, d);
  unsigned long long h[3];
  CLV_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream()));
  CLV_CUDA(cudaStreamSynchronize(stream()));
  CLV_CUDA(cudaFree(d));
  *mismatches = (long long)h[0];
  *flagged = (long long)h[1];
  *checked = (long long)h[2];
}
