// Issue/pipe throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, with and without LDS.128 mixed in.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH> __global__ void __launch_bounds__(256) probe(float* out, const float* in, int iters)
{
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(in[0], in[1], in[2], in[3]);
  __syncthreads();
  float2 a[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) a[c] = make_float2(in[c] + threadIdx.x, in[c + 1]);
  float2 m = make_float2(in[4], in[5]), k = make_float2(in[6], in[7]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 2) {  // LDS.128 broadcast feeding the multiplier, then packed
        const float4 v = sm[(it + r) & 63];
        m = make_float2(v.x, v.y);
        k = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (MODE == 0) {  // scalar: two FFMA
          a[c].x = fmaf(a[c].x, m.x, k.x);
          a[c].y = fmaf(a[c].y, m.y, k.y);
        } else {  // packed: one FFMA2
          a[c] = __ffma2_rn(a[c], m, k);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += a[c].x + a[c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CH> void run(const char* name, int warpsPerSM)
{
  float *out, *in;
  const int blocks = 148 * (warpsPerSM / 8 > 0 ? warpsPerSM / 8 : 1);
  const int threads = warpsPerSM >= 8 ? 256 : warpsPerSM * 32;
  cudaMalloc(&out, sizeof(float) * blocks * 256);
  cudaMalloc(&in, 64);
  float h[16] = {1.0f, 0.999f, 1.0f, 0.5f, 0.9999f, 0.9998f, 1e-3f, 2e-3f};
  cudaMemcpy(in, h, 64, cudaMemcpyHostToDevice);
  const int iters = 20000;
  probe<MODE, CH><<<blocks, threads>>>(out, in, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE, CH><<<blocks, threads>>>(out, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = 2.0 * CH * 8.0 * iters * blocks * threads;  // scalar FMAs
  printf("%-28s warps/SM %2d chains %d: %.3f ms  %.2f TFLOP/s  (%.1f FMA lanes/clk/SM at 1.9 GHz)\n", name, warpsPerSM, CH, ms,
         2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
  cudaFree(in);
}

int main()
{
  for (int w : {4, 8, 16, 32}) {
    run<0, 1>("FFMA x2 (1 dep chain pair)", w);
    run<1, 1>("FFMA2 (1 dep chain)", w);
    run<0, 4>("FFMA x2", w);
    run<1, 4>("FFMA2", w);
    run<2, 4>("FFMA2 + LDS.128 per 4", w);
    run<0, 8>("FFMA x2", w);
    run<1, 8>("FFMA2", w);
  }
  return 0;
}
