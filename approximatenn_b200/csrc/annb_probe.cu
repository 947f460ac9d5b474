// annb_probe.cu — FP32-pipe probe: the denominators of the S3 roofline, measured on the device
// the library runs on ("measure an FFMA microbenchmark for the real figure").
//   mode 0  FFMA chains            (one instruction = 2 flops: the usual "FP32 peak")
//   mode 1  FMUL + FADD chains     (separately rounded, what the reference's arithmetic allows
//                                   without packing: 1 flop per instruction)
//   mode 2  FFMA2 chains           (packed fp32x2: 4 flops per instruction)
//   mode 3  FMUL2 + FADD2 chains   (packed, separately rounded: 2 flops per instruction — the
//                                   form the exact distance tree uses)
// Returns TFLOP/s (flops as counted above) of the best of `reps` timed launches.
#include "annb_common.cuh"

typedef unsigned long long u64p;

__device__ __forceinline__ u64p probe_pack(float lo, float hi) {
  u64p r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b) {
  constexpr int CH = 8;                                        // independent chains per thread
  if (MODE < 2) {
    float x[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = (float)(threadIdx.x + c) * 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (MODE == 0) x[c] = __fmaf_rn(x[c], a, b);
        else x[c] = __fadd_rn(__fmul_rn(x[c], a), b);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) s += x[c];
    if (s == 12345.678f) out[0] = s;                           // keeps the chains alive
  } else {
    u64p x[CH];
    const u64p a2 = probe_pack(a, a), b2 = probe_pack(b, b);
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = probe_pack((float)(threadIdx.x + c) * 1e-3f, (float)c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (MODE == 2) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a2), "l"(b2));
        } else {
          // even chains only multiply, odd chains only add: no product ever feeds a sum, so
          // ptxas cannot contract a pair into one FFMA2 (it does that for mul+add of the same
          // value even under -fmad=false)
          asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[c & ~1]) : "l"(a2));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[c | 1]) : "l"(b2));
        }
      }
    }
    u64p s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s ^= x[c];
    if (s == 0x1234567812345678ull) out[0] = 1.f;
  }
}

extern "C" double annb_probe_fp32(int mode, int reps, annb_stream stream) {
  int dev = 0, sms = 0;
  RT_CHECK(cudaGetDevice(&dev));
  RT_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float *out = nullptr;
  RT_CHECK(cudaMalloc((void **)&out, 256));
  cudaEvent_t e0, e1;
  RT_CHECK(cudaEventCreate(&e0));
  RT_CHECK(cudaEventCreate(&e1));
  const int iters = 8192, grid = sms * 8, block = 256;
  const double per_inst[4] = {2.0, 1.0, 4.0, 2.0};
  if (mode < 0 || mode > 3) mode = 0;
  // instructions per thread: iters * 8 chains * (1 or 2 instructions)
  const double inst = (double)iters * 8 * ((mode == 0 || mode == 2) ? 1 : 2);
  const double flops = inst * per_inst[mode] * (double)grid * block;
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    RT_CHECK(cudaEventRecord(e0, stream));
    switch (mode) {
      case 0: fp32_probe_kernel<0><<<grid, block, 0, stream>>>(out, iters, 1.0000001f, 1e-9f); break;
      case 1: fp32_probe_kernel<1><<<grid, block, 0, stream>>>(out, iters, 1.0000001f, 1e-9f); break;
      case 2: fp32_probe_kernel<2><<<grid, block, 0, stream>>>(out, iters, 1.0000001f, 1e-9f); break;
      default: fp32_probe_kernel<3><<<grid, block, 0, stream>>>(out, iters, 1.0000001f, 1e-9f); break;
    }
    LAUNCH_CHECK("fp32_probe");
    RT_CHECK(cudaEventRecord(e1, stream));
    RT_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    RT_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms > 0) {                                      // first launch is the warm-up
      double tf = flops / (ms * 1e-3) / 1e12;
      if (tf > best) best = tf;
    }
  }
  RT_CHECK(cudaEventDestroy(e0));
  RT_CHECK(cudaEventDestroy(e1));
  RT_CHECK(cudaFree(out));
  return best;
}
