// peak_kernel.cu -- measures the fp64 FMA issue peak of the device the solve kernels run on.
//
// The restart solve (solve_t1_kernel / solve_kernel) is bound by fp64 issue and dependency latency, not by HBM
// (DESIGN.md section 5), and MEASURED_PEAKS.json carries no fp64 figure, so bench.py measures the denominator of
// `roofline_solve` itself: every thread runs 8 independent DFMA chains (enough to cover the pipe's latency at 8 warps
// per scheduler), no memory traffic; achieved = 2 * fma count / time.
#include <cuda_runtime.h>

namespace optik {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true for the arguments used: keeps the chains alive
}
}  // namespace optik

// -> fp64 TFLOP/s sustained over about `seconds` of back-to-back launches (CUDA events on `stream`); < 0 on error
extern "C" double optik_measure_fp64_peak(int device, double seconds) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  double* out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  const double flop_per_launch = 2.0 * 64.0 * iters * 256.0 * blocks;
  optik::fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);  // warm-up
  cudaDeviceSynchronize();
  double best = 0.0, total_ms = 0.0;
  int launches = 0;
  while (total_ms < seconds * 1e3 && launches < 100000) {
    cudaEventRecord(e0);
    for (int k = 0; k < 8; k++) optik::fp64_peak_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    total_ms += ms;
    launches += 8;
    const double tf = 8.0 * flop_per_launch / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}
