// Microbenchmark (B200): warp-instruction issue rates of the FP32 instruction mix the brute-force detector test uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/microbench_fp32_pipes scripts/microbench_fp32_pipes.cu
// Prints, per variant, lane-operations/s and warp-instructions per cycle per SM sub-partition (SMSP).
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rc);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

constexpr int kChains = 8;

// variant 0: FADD only; 1: FADD2 only; 2: FMNMX only; 3: per-ray test scalar (3 FADD + FMNMX + FADD, max-reduced);
// 4: the same with FADD2 updates of two voxels at a time and FMNMX3 reduction
template <int V>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float d) {
  float acc = -1e30f;
  float x[kChains], y[kChains], w[kChains];
  float2 x2[kChains], y2[kChains], w2[kChains];
  for (int i = 0; i < kChains; ++i) {
    x[i] = threadIdx.x * 0.001f + i, y[i] = x[i] * 0.5f, w[i] = x[i] * 0.25f;
    x2[i] = make_float2(x[i], x[i] + d), y2[i] = make_float2(y[i], y[i] + d), w2[i] = make_float2(w[i], w[i] + d);
  }
  const float2 d2 = make_float2(d, d);
  for (int it = 0; it < iters; ++it) {
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) x[i] += d, y[i] += d, w[i] += d;
    } else if (V == 1) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) x2[i] = add2(x2[i], d2), y2[i] = add2(y2[i], d2), w2[i] = add2(w2[i], d2);
    } else if (V == 2) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) x[i] = fmaxf(x[i], y[i] + 0.f), y[i] = fminf(y[i], w[i]), w[i] = fmaxf(w[i], d);
    } else if (V == 3) {
      float t[kChains];
#pragma unroll
      for (int i = 0; i < kChains; ++i) {
        t[i] = fabsf(w[0]) - fmaxf(fabsf(x[0]), fabsf(y[0]));
        x[0] += d, y[0] += d, w[0] += d;
      }
      float tm = t[0];
#pragma unroll
      for (int i = 1; i < kChains; ++i) tm = fmaxf(tm, t[i]);
      acc = fmaxf(acc, tm);
    } else if (V == 4) {
      float t[kChains];
#pragma unroll
      for (int i = 0; i < kChains; i += 2) {
        t[i] = fabsf(w2[0].x) - fmaxf(fabsf(x2[0].x), fabsf(y2[0].x));
        t[i + 1] = fabsf(w2[0].y) - fmaxf(fabsf(x2[0].y), fabsf(y2[0].y));
        x2[0] = add2(x2[0], d2), y2[0] = add2(y2[0], d2), w2[0] = add2(w2[0], d2);
      }
      float tm = max3(t[0], t[1], t[2]);
      tm = max3(tm, t[3], t[4]);
      tm = max3(tm, t[5], t[6]);
      acc = max3(acc, tm, t[7]);
    }
  }
  float s = acc;
  for (int i = 0; i < kChains; ++i) s += x[i] + y[i] + w[i] + x2[i].x + x2[i].y + y2[i].x + y2[i].y + w2[i].x + w2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, double warp_instr_per_iter, double lane_ops_per_iter) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int blocks = sms * 8, threads = 256, iters = 20000;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  probe<V><<<blocks, threads>>>(out, iters, 1e-3f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    probe<V><<<blocks, threads>>>(out, iters, 1e-3f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  const double warps = (double)blocks * threads / 32, s = best * 1e-3;
  const double winstr = warps * iters * warp_instr_per_iter;
  printf("{\"variant\": \"%s\", \"ms\": %.3f, \"warp_instr_per_cycle_per_smsp\": %.3f, \"lane_ops_per_s\": %.4g}\n", name, best,
         winstr / s / (sms * 4.0) / (khz * 1e3), warps * 32 * iters * lane_ops_per_iter / s);
  cudaFree(out);
}

int main() {
  run<0>("FADD x24 per iteration", 24, 24);
  run<1>("FADD2 x24 per iteration (48 lane adds)", 24, 48);
  run<2>("FMNMX x24 per iteration", 24, 24);
  run<3>("scalar test: 8 voxels = 24 FADD + 8 FMNMX + 8 FADD + 8 FMNMX", 48, 8);
  run<4>("packed test: 8 voxels = 12 FADD2 + 8 FMNMX + 8 FADD + 4 FMNMX3", 32, 8);
  return 0;
}
