// Developer microbenchmark: issue throughput of the instruction kinds the Bellman-sweep kernel is made of,
// on one SM sub-partition's worth of warps (results steer which pipe the per-cell arithmetic goes to).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes tools/ubench/pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

#define BODY(NAME, DECL, STMT)                                                              \
  __global__ void __launch_bounds__(512) NAME(float* out, float seed, long long* cyc) {    \
    float x[ILP], y[ILP];                                                                   \
    for (int i = 0; i < ILP; ++i) { x[i] = seed + threadIdx.x + i; y[i] = seed * 0.5f + i; } \
    DECL;                                                                                   \
    long long t0 = clock64();                                                               \
    for (int it = 0; it < ITERS; ++it) {                                                    \
      _Pragma("unroll") for (int i = 0; i < ILP; ++i) { STMT; }                             \
    }                                                                                       \
    long long t1 = clock64();                                                               \
    float s = 0;                                                                            \
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];                                          \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                         \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                        \
  }

BODY(k_fmul, float c = seed, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c)))
BODY(k_fadd, float c = seed, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c)))
BODY(k_ffma, float c = seed, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c), "f"(y[i])))
BODY(k_ffma_imm, , asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(x[i]) : "f"(y[i])))
BODY(k_fmul2, float c = seed,
     asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}"
                  : "+f"(x[i]), "+f"(y[i]) : "f"(c)))
BODY(k_fadd2, float c = seed,
     asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; add.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}"
                  : "+f"(x[i]), "+f"(y[i]) : "f"(c)))
BODY(k_ffma2, float c = seed,
     asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; fma.rn.f32x2 a, a, b, b; mov.b64 {%0, %1}, a;}"
                  : "+f"(x[i]), "+f"(y[i]) : "f"(c)))
BODY(k_fsel, float c = seed,
     asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; selp.f32 %0, %0, %2, p;}" : "+f"(x[i]) : "f"(c), "f"(y[i])))
BODY(k_selp_only, int pr = seed > 0,
     asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 0; selp.f32 %0, %0, %2, p;}" : "+f"(x[i]) : "r"(pr), "f"(y[i])))
BODY(k_fset, float c = seed, asm volatile("set.eq.f32.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c)))
BODY(k_fmnmx, float c = seed, asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c)))
BODY(k_fmnmx3, float c = seed, asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c), "f"(y[i])))
BODY(k_lop3, , asm volatile("{.reg .b32 a, b; mov.b32 a, %0; mov.b32 b, %1; lop3.b32 a, a, b, 0x18181818, 0x6a; mov.b32 %0, a;}" : "+f"(x[i]) : "f"(y[i])))
BODY(k_prmt, , asm volatile("{.reg .b32 a, b; mov.b32 a, %0; mov.b32 b, %1; prmt.b32 a, a, b, 0x4441; mov.b32 %0, a;}" : "+f"(x[i]) : "f"(y[i])))
BODY(k_frnd, , asm volatile("cvt.rni.f32.f32 %0, %0;" : "+f"(x[i])))
BODY(k_imad, int c = (int)seed, asm volatile("{.reg .b32 a, b; mov.b32 a, %0; mov.b32 b, %2; mad.lo.s32 a, a, %1, b; mov.b32 %0, a;}" : "+f"(x[i]) : "r"(c), "f"(y[i])))
// mixes: one FMA-pipe + one ALU-pipe instruction per slot pair
BODY(k_mix_fmul_fsel, int pr = seed > 0,
     asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; mul.rn.f32 %0, %0, %3; selp.f32 %1, %1, %3, p;}"
                  : "+f"(x[i]), "+f"(y[i]) : "r"(pr), "f"(seed)))
BODY(k_mix_fmul2_fsel, int pr = seed > 0,
     asm volatile("{.reg .pred p; .reg .b64 a, b; setp.ne.s32 p, %2, 0; mov.b64 a, {%0, %0}; mov.b64 b, {%3, %3}; mul.rn.f32x2 a, a, b; "
                  "mov.b64 {%0, _}, a; selp.f32 %1, %1, %3, p;}"
                  : "+f"(x[i]), "+f"(y[i]) : "r"(pr), "f"(seed)))
BODY(k_mix_ffma_fset, float c = seed,
     asm volatile("fma.rn.f32 %0, %0, %2, %1; set.eq.f32.f32 %1, %1, %2;" : "+f"(x[i]), "+f"(y[i]) : "f"(c)))

template <typename K>
static void run(const char* name, K kernel, int per_iter, int warps_per_smsp) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 512 * 4 * 148);
  cudaMalloc(&cyc, 148 * 8);
  const int threads = warps_per_smsp * 4 * 32;
  kernel<<<148, threads>>>(out, 1.0f, cyc);
  kernel<<<148, threads>>>(out, 1.0f, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const double warp_instr_per_smsp = double(ITERS) * ILP * per_iter * warps_per_smsp;
  printf("%-18s warps/SMSP %d: %.3f cycles per warp-instruction per SMSP (%.2f instr/clk)\n", name, warps_per_smsp,
         c / warp_instr_per_smsp, warp_instr_per_smsp / c);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4}) {
    run("FMUL", k_fmul, 1, w);
    run("FADD", k_fadd, 1, w);
    run("FFMA 3-reg", k_ffma, 1, w);
    run("FFMA imm", k_ffma_imm, 1, w);
    run("FMUL2", k_fmul2, 1, w);
    run("FADD2", k_fadd2, 1, w);
    run("FFMA2", k_ffma2, 1, w);
    run("FSETP+SEL", k_fsel, 2, w);
    run("SEL(+isetp hoisted)", k_selp_only, 1, w);
    run("FSET", k_fset, 1, w);
    run("FMNMX", k_fmnmx, 1, w);
    run("FMNMX3", k_fmnmx3, 1, w);
    run("LOP3", k_lop3, 1, w);
    run("PRMT", k_prmt, 1, w);
    run("FRND", k_frnd, 1, w);
    run("IMAD", k_imad, 1, w);
    run("FMUL+FSEL", k_mix_fmul_fsel, 2, w);
    run("FMUL2+FSEL", k_mix_fmul2_fsel, 2, w);
    run("FFMA+FSET", k_mix_ffma_fset, 2, w);
  }
  return 0;
}
