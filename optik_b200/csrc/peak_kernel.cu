// peak_kernel.cu -- measurement-only kernels: the fp64 FMA issue peak of the device the solve kernels run on, and the HBM
// rate of a pure streaming kernel with the evaluator's read/write mix.
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
// Streaming kernel with the evaluator's read/write MIX and nothing else: per item `rd` 16-byte units are read and `wr`
// 16-byte units written, fully coalesced (unit u of item i sits at [u][i]), grid-stride.  What HBM delivers to a kernel
// that is 80 % stores, next to the 50/50 copy figure of MEASURED_PEAKS.json.
__global__ void __launch_bounds__(256) hbm_mix_kernel(const double2* __restrict__ in, double2* __restrict__ out, unsigned long long items,
                                                      int rd, int wr) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
    double2 acc = make_double2(0.0, 0.0);
    for (int u = 0; u < rd; u++) {
      const double2 v = in[(unsigned long long)u * items + i];
      acc.x += v.x; acc.y += v.y;
    }
    for (int u = 0; u < wr; u++) out[(unsigned long long)u * items + i] = make_double2(acc.x + u, acc.y);
  }
}
}  // namespace optik

// -> GB/s (read + written bytes / time) of the mix kernel over `items` items of rd_units / wr_units 16-byte units each,
// best of `reps` launches; < 0 on error.  Measurement only (bench.py: the ceiling beside the evaluator's roofline).
extern "C" double optik_measure_hbm_mix(int device, unsigned long long items, int rd_units, int wr_units, int reps) {
  if (cudaSetDevice(device) != cudaSuccess || items == 0 || rd_units < 1 || wr_units < 1) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  double2 *in = nullptr, *out = nullptr;
  if (cudaMalloc(&in, items * 16ull * rd_units) != cudaSuccess) return -1.0;
  if (cudaMalloc(&out, items * 16ull * wr_units) != cudaSuccess) { cudaFree(in); return -1.0; }
  cudaMemset(in, 0, items * 16ull * rd_units);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8;
  double best = -1.0;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
    optik::hbm_mix_kernel<<<blocks, 256>>>(in, out, items, rd_units, wr_units);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gbs = (double)items * 16.0 * (rd_units + wr_units) / (ms * 1e-3) / 1e9;
    if (r > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(in); cudaFree(out);
  return best;
}

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
