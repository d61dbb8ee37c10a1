// Compiles the REFERENCE's Non-CDP/NW/main.cu for the host so that its own nw_cpu / generateQuery can pin oracle/nw_linear.c
// (test infrastructure; the source is read in place from /root/reference, only the binary lands in oracle/_ref/).
// The CUDA runtime calls main() makes are stubbed: main() is never run, only nw_cpu and generateQuery are exported.
#include <stddef.h>
#include <string.h>
typedef int cudaError_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t) { *p = 0; return 0; }
static inline cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind) { return 0; }
static inline cudaError_t cudaMemset(void*, int, size_t) { return 0; }
static inline cudaError_t cudaFree(void*) { return 0; }
#define main nw_reference_main
#include "main.cu"
#undef main
void nw_gpu0(unsigned char*, unsigned char*, int*, unsigned int) {}
void nw_gpu1(unsigned char*, unsigned char*, int*, unsigned int) {}
void nw_gpu2(unsigned char*, unsigned char*, int*, unsigned int) {}
void nw_gpu3(unsigned char*, unsigned char*, int*, unsigned int) {}
extern "C" void nw_ref_cpu(unsigned char* reference, unsigned char* query, int* matrix, unsigned int N) { nw_cpu(reference, query, matrix, N); }
extern "C" void nw_ref_generate_query(unsigned char* reference, unsigned char* query, unsigned int N, unsigned int seed) {
  srand(seed);
  for (unsigned int r = 0; r < N; ++r) reference[r] = rand() % 256;      // main.cu:93-95
  generateQuery(reference, query, N);                                    // main.cu:96
}
