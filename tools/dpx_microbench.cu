// DPX / integer-pipe issue-rate microbenchmark for B200 (sm_100a).
// Measures warp-instructions per clock per SM for the instruction classes the DP kernels
// are built from, alone and in pairs (to find which ones share an execution pipe).
// The result (ops/clk/SM) is the denominator "R" of the GCUPS roofline.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dpx_microbench dpx_microbench.cu
// run  : ./dpx_microbench > gpurun_out/dpx_microbench.json
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

enum Op {
  OP_VIADDMNMX_S16X2 = 0, OP_VIMNMX3_S16X2, OP_VIMNMX_S16X2, OP_VIADD_16X2, OP_VIADDMNMX_S32, OP_VIMNMX3_S32,
  OP_IADD, OP_IMAD, OP_LOP3, OP_HSET2, OP_PRMT, OP_SHF, OP_VIADDMNMX_S16X2_RELU, OP_HADD2, OP_HMNMX2, OP_SELP, OP_NOP_COUNT
};

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b, uint32_t c) {
  if (OP == OP_VIADDMNMX_S16X2) return __viaddmax_s16x2(a, b, c);
  if (OP == OP_VIADDMNMX_S16X2_RELU) return __viaddmax_s16x2_relu(a, b, c);
  if (OP == OP_VIMNMX3_S16X2) return __vimax3_s16x2(a, b, c);
  if (OP == OP_VIMNMX_S16X2) { uint32_t r; asm volatile("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_VIADD_16X2) return __vadd2(a, b);
  if (OP == OP_VIADDMNMX_S32) return (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);
  if (OP == OP_VIMNMX3_S32) return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c);
  if (OP == OP_IADD) { uint32_t r; asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_IMAD) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_LOP3) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_HSET2) { uint32_t r; asm volatile("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_PRMT) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_SHF) { uint32_t r; asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_HADD2) { uint32_t r; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_HMNMX2) { uint32_t r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_SELP) { uint32_t r; asm volatile("{.reg .pred p; setp.ne.u32 p, %3, 0; selp.b32 %0, %1, %2, p;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  return a;
}

// OP2 == OP_NOP_COUNT -> single-op kernel. ILP independent chains per op.
template <int OP1, int OP2, int ILP>
__global__ void __launch_bounds__(256) k_bench(uint32_t* out, const uint32_t* in, int iters, long long* cyc) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[ILP], d[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) { a[j] = in[(tid + j) & 1023]; d[j] = in[(tid + 7 * j + 3) & 1023]; }
  uint32_t b = in[(tid + 11) & 1023], c = in[(tid + 13) & 1023];
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        a[j] = apply<OP1>(a[j], b, c);
        if (OP2 != OP_NOP_COUNT) d[j] = apply<OP2>(d[j], c, b);
      }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) r ^= a[j] ^ d[j];
  out[tid] = r;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}

struct Result { std::string name; double ops_per_clk_sm; double ms; double ghz; double lane_ops_per_s; int nops; };

static const char* opname(int op) {
  switch (op) {
    case OP_VIADDMNMX_S16X2: return "VIADDMNMX.S16x2"; case OP_VIMNMX3_S16X2: return "VIMNMX3.S16x2";
    case OP_VIMNMX_S16X2: return "VIMNMX.S16x2(fused pairwise into VIMNMX3)"; case OP_VIADD_16X2: return "VIADD.16x2";
    case OP_VIADDMNMX_S32: return "VIADDMNMX.S32"; case OP_VIMNMX3_S32: return "VIMNMX3.S32";
    case OP_IADD: return "IADD3"; case OP_IMAD: return "IMAD"; case OP_LOP3: return "LOP3";
    case OP_HSET2: return "HSET2"; case OP_PRMT: return "PRMT"; case OP_SHF: return "SHF";
    case OP_VIADDMNMX_S16X2_RELU: return "VIADDMNMX.S16x2.RELU"; case OP_HADD2: return "HFMA2";
    case OP_HMNMX2: return "HMNMX2"; case OP_SELP: return "SEL";
  }
  return "none";
}

template <int OP1, int OP2, int ILP>
Result run(int n_sm, uint32_t* d_out, uint32_t* d_in, long long* d_cyc, int threads = 256) {
  const int blocks_per_sm = 4;
  const int grid = n_sm * blocks_per_sm;
  const int iters = 131072;   // ~35 ms per kernel: long enough for the SM clock to settle
  k_bench<OP1, OP2, ILP><<<grid, threads>>>(d_out, d_in, 64, d_cyc);  // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_bench<OP1, OP2, ILP><<<grid, threads>>>(d_out, d_in, iters, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(2 * grid);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
  // average per-block duration in SM cycles; all blocks are co-resident (one wave)
  double sum = 0; for (int b = 0; b < grid; b++) sum += (double)(cyc[2 * b + 1] - cyc[2 * b]);
  double avg_cyc = sum / grid;
  const int nops = (OP2 == OP_NOP_COUNT) ? 1 : 2;
  double warp_instr_per_sm = (double)iters * 4 * ILP * nops * (threads / 32) * blocks_per_sm;
  if (OP1 == OP_VIMNMX_S16X2 && OP2 == OP_NOP_COUNT) warp_instr_per_sm *= 0.5;  // ptxas fuses max(max(a,b),b) pairs into one VIMNMX3
  Result r;
  r.name = std::string(opname(OP1)) + (OP2 == OP_NOP_COUNT ? "" : std::string("+") + opname(OP2));
  if (threads != 256) r.name += " @" + std::to_string(threads * blocks_per_sm / 32) + " warps/SM";
  r.ops_per_clk_sm = warp_instr_per_sm * 32.0 / avg_cyc;   // lane-ops per clock per SM
  r.ms = ms;
  r.ghz = avg_cyc / (ms * 1e6);
  r.lane_ops_per_s = warp_instr_per_sm * 32.0 * n_sm / (ms * 1e-3);   // wall-clock (CUDA events): the roofline number
  r.nops = nops;
  return r;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int n_sm = prop.multiProcessorCount;
  uint32_t *d_out, *d_in; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(uint32_t) * n_sm * 4 * 256));
  CK(cudaMalloc(&d_in, sizeof(uint32_t) * 1024));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 2 * n_sm * 4));
  std::vector<uint32_t> h(1024);
  for (int i = 0; i < 1024; i++) h[i] = (uint32_t)(i * 2654435761u) & 0x00ff00ffu;
  CK(cudaMemcpy(d_in, h.data(), sizeof(uint32_t) * 1024, cudaMemcpyHostToDevice));
  std::vector<Result> res;
#define RUN1(op) res.push_back(run<op, OP_NOP_COUNT, 8>(n_sm, d_out, d_in, d_cyc))
#define RUN2(o1, o2) res.push_back(run<o1, o2, 6>(n_sm, d_out, d_in, d_cyc))
  RUN1(OP_VIADDMNMX_S16X2); RUN1(OP_VIADDMNMX_S16X2_RELU); RUN1(OP_VIMNMX3_S16X2); RUN1(OP_VIMNMX_S16X2); RUN1(OP_VIADD_16X2);
  RUN1(OP_VIADDMNMX_S32); RUN1(OP_VIMNMX3_S32); RUN1(OP_IADD); RUN1(OP_IMAD); RUN1(OP_LOP3); RUN1(OP_HSET2);
  RUN1(OP_PRMT); RUN1(OP_SHF); RUN1(OP_HADD2);
  RUN2(OP_VIADDMNMX_S16X2, OP_VIMNMX3_S16X2); RUN2(OP_VIADDMNMX_S16X2, OP_VIADD_16X2); RUN2(OP_VIADDMNMX_S16X2, OP_IADD);
  RUN2(OP_VIADDMNMX_S16X2, OP_IMAD); RUN2(OP_VIADDMNMX_S16X2, OP_LOP3); RUN2(OP_VIADDMNMX_S16X2, OP_HSET2);
  RUN2(OP_VIADDMNMX_S16X2, OP_PRMT); RUN2(OP_VIADDMNMX_S16X2, OP_HADD2); RUN2(OP_VIADD_16X2, OP_IMAD); RUN2(OP_VIADD_16X2, OP_LOP3);
  RUN2(OP_HSET2, OP_IMAD); RUN2(OP_HSET2, OP_LOP3); RUN2(OP_IMAD, OP_LOP3); RUN2(OP_VIMNMX3_S32, OP_IMAD);
  RUN1(OP_HMNMX2); RUN2(OP_VIADDMNMX_S16X2, OP_HMNMX2); RUN2(OP_VIADD_16X2, OP_HMNMX2); RUN2(OP_VIADDMNMX_S16X2, OP_SELP);
  // the same mixes at the occupancy the DP fast path runs at (8 warps per SM = 2 per scheduler)
  res.push_back(run<OP_VIADDMNMX_S16X2, OP_NOP_COUNT, 8>(n_sm, d_out, d_in, d_cyc, 64));
  res.push_back(run<OP_VIADDMNMX_S16X2, OP_VIADD_16X2, 6>(n_sm, d_out, d_in, d_cyc, 64));
  res.push_back(run<OP_VIADDMNMX_S16X2, OP_VIMNMX3_S16X2, 6>(n_sm, d_out, d_in, d_cyc, 64));
  double R_alu = 0, R_fma = 0, R_dual = 0;
  for (auto& r : res) {
    if (r.nops == 1 && (r.name.find("VIADDMNMX.S16x2") == 0 || r.name.find("VIMNMX3.S16x2") == 0)) R_alu = std::max(R_alu, r.lane_ops_per_s);
    if (r.nops == 1 && (r.name == "VIADD.16x2" || r.name == "IMAD")) R_fma = std::max(R_fma, r.lane_ops_per_s);
    if (r.name == "VIADDMNMX.S16x2+VIADD.16x2") R_dual = r.lane_ops_per_s;
  }
  printf("{\"gpu\": \"%s\", \"n_sm\": %d, \"clock_rate_khz\": %d,\n", prop.name, n_sm, prop.clockRate);
  printf(" \"R_alu_lane_ops_per_s\": %.6g, \"R_fma_lane_ops_per_s\": %.6g, \"R_alu_plus_fma_lane_ops_per_s\": %.6g,\n", R_alu, R_fma, R_dual);
  printf(" \"note\": \"lane_ops_per_s = warp instructions x 32 / CUDA-event time over all SMs (1024 threads/SM, ILP 8 or 6+6); R_alu is the DPX/ALU-pipe peak used as the GCUPS roofline denominator; the clk-based column divides by clock64() deltas and is informational only\",\n \"results\": [\n");
  for (size_t i = 0; i < res.size(); i++)
    printf("  {\"op\": \"%s\", \"lane_ops_per_s\": %.5g, \"ms\": %.3f, \"lane_ops_per_clk_per_sm_by_clock64\": %.2f}%s\n", res[i].name.c_str(),
           res[i].lane_ops_per_s, res[i].ms, res[i].ops_per_clk_sm, i + 1 < res.size() ? "," : "");
  printf("]}\n");
  return 0;
}
