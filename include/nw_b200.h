/* nw_b200.h -- C ABI of the full-matrix linear-gap Needleman-Wunsch kernel.
 *
 * Replaces the suite's NW benchmark kernels: Non-CDP/NW/common.h:9-12 declares
 *     void nw_gpu0..3(unsigned char* reference_d, unsigned char* query_d, int* matrix_d, unsigned int N);
 * each of which fills the N x N score matrix  matrix[q*N + r]  of the recurrence in NW/main.cu:9-27 (nw_cpu):
 *     top     = q == 0 ? (r+1)*DELETION  : matrix[(q-1)*N + r]
 *     left    = r == 0 ? (q+1)*INSERTION : matrix[q*N + r-1]
 *     topleft = q == 0 ? r*DELETION : (r == 0 ? q*INSERTION : matrix[(q-1)*N + r-1])
 *     matrix[q*N + r] = max(topleft + (query[q] == reference[r] ? MATCH : MISMATCH), top + INSERTION, left + DELETION)
 * with MATCH 1, MISMATCH -1, INSERTION -1, DELETION -1 (NW/common.h:5-8).  The reference walks the anti-diagonals of 64 x 64
 * tiles with one kernel launch + cudaDeviceSynchronize per tile diagonal (NW/kernel3.cu:86-103: 2*ceil(N/64) launches);
 * this library computes the whole matrix in ONE persistent launch (csrc/nw_matrix.cu).  Plain pointers and sizes only. */
#ifndef NW_B200_H
#define NW_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t match, mismatch, insertion, deletion; } nwx_scores_t;   /* NW/common.h:5-8, run-time here */

/* Drop-in for nw_gpu0..3 (same argument list, device pointers, default scores, default stream).  Like the reference's
 * functions it returns after the work has been queued AND finished (they end with cudaDeviceSynchronize). */
void nw_gpu_b200(unsigned char* reference_d, unsigned char* query_d, int* matrix_d, unsigned int N);

/* Asynchronous form: device pointers, explicit scores (NULL = defaults) and CUDA stream (NULL = default stream).
 * Returns 0, or -1 on invalid arguments / allocation failure (message on stderr). */
int nwx_matrix_async(const unsigned char* reference_d, const unsigned char* query_d, int* matrix_d, unsigned int N,
                     const nwx_scores_t* scores, void* cuda_stream);

/* Host-buffer form (what NW/main.cu:110-160 does around the kernel): H2D of both sequences, the kernel, and -- if
 * matrix_host != NULL -- the D2H of the N*N*4-byte matrix.  ms[0..2] (optional) = H2D, kernel, D2H in milliseconds from
 * CUDA events; checksum (optional) = nwx_checksum of the device matrix.  The device matrix is kept in a per-device scratch
 * that nwx_release frees. */
int nwx_matrix_host(const unsigned char* reference, const unsigned char* query, int* matrix_host, unsigned int N,
                    const nwx_scores_t* scores, float ms[3], uint64_t* checksum);

/* sum over i of (uint32)matrix[i] * (((uint32)i * 2654435761u) | 1u)  mod 2^64, computed on the device (stream-ordered,
 * then synchronised).  Lets a caller compare a 4 GB matrix with a streaming CPU run without copying it back. */
uint64_t nwx_checksum(const int* matrix_d, uint64_t n_elements, void* cuda_stream);

/* Kernel variant (A/B testing): 0 = warp-specialised (default), 1 = one warp per row-block. */
void nwx_set_variant(int variant);

/* Number of kernels launched by this module since load (for launch accounting), and scratch release. */
uint64_t nwx_kernel_launches(void);
void nwx_release(void);

#ifdef __cplusplus
}
#endif
#endif /* NW_B200_H */
