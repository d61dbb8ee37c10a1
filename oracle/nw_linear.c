/* oracle/nw_linear.c -- CPU restatement of the suite's full-matrix linear-gap Needleman-Wunsch (TEST INFRASTRUCTURE ONLY:
 * tests/, bench.py's cpu_baseline leg and smoke() may call it; the product never does).
 *
 * Follows Non-CDP/NW/main.cu:9-27 (nw_cpu) cell by cell; scores are run-time here (NW/common.h:5-8 fixes them to
 * MATCH 1, MISMATCH -1, INSERTION -1, DELETION -1).  Pinned against the reference's own nw_cpu compiled for the host
 * (oracle/_ref/libnw_ref_host.so, built from the sources in place) in tests/test_nw_oracle.py. */
#include <stdint.h>
#include <stdlib.h>

static inline int max3(int a, int b, int c) { int m = a > b ? a : b; return c > m ? c : m; }

/* matrix[q*N + r], q over the query, r over the reference (main.cu:10-11) */
void nw_linear_matrix(const unsigned char* reference, const unsigned char* query, int* matrix, unsigned N,
                      int match, int mismatch, int insertion, int deletion) {
  for (unsigned q = 0; q < N; ++q) {
    for (unsigned r = 0; r < N; ++r) {
      const int top = (q == 0) ? (int)(r + 1) * deletion : matrix[(size_t)(q - 1) * N + r];                      /* :13 */
      const int left = (r == 0) ? (int)(q + 1) * insertion : matrix[(size_t)q * N + (r - 1)];                   /* :14 */
      const int topleft = (q == 0) ? (int)r * deletion : ((r == 0) ? (int)q * insertion : matrix[(size_t)(q - 1) * N + (r - 1)]);   /* :15 */
      const int ins = top + insertion, del = left + deletion;                                                    /* :17-18 */
      const int m = topleft + ((query[q] == reference[r]) ? match : mismatch);                                   /* :19 */
      matrix[(size_t)q * N + r] = max3(ins, del, m);     /* :21-23: (ins > del ? ins : del), then (m > max ? m : max) == max of the three */
    }
  }
}

/* Same recurrence with two rolling rows; returns sum_i (uint32)matrix[i] * (((uint32)i * 2654435761u) | 1u) mod 2^64, the
 * hash include/nw_b200.h defines for nwx_checksum, so a 4 GB matrix can be checked without materialising it on the host. */
uint64_t nw_linear_checksum(const unsigned char* reference, const unsigned char* query, unsigned N,
                            int match, int mismatch, int insertion, int deletion) {
  if (N == 0) return 0;
  int* prev = (int*)malloc(sizeof(int) * N);
  int* cur = (int*)malloc(sizeof(int) * N);
  uint64_t h = 0;
  for (unsigned q = 0; q < N; ++q) {
    for (unsigned r = 0; r < N; ++r) {
      const int top = (q == 0) ? (int)(r + 1) * deletion : prev[r];
      const int left = (r == 0) ? (int)(q + 1) * insertion : cur[r - 1];
      const int topleft = (q == 0) ? (int)r * deletion : ((r == 0) ? (int)q * insertion : prev[r - 1]);
      const int v = max3(top + insertion, left + deletion, topleft + ((query[q] == reference[r]) ? match : mismatch));
      cur[r] = v;
      const uint64_t i = (uint64_t)q * N + r;
      h += (uint64_t)(uint32_t)v * (uint64_t)((((uint32_t)i) * 2654435761u) | 1u);
    }
    int* t = prev; prev = cur; cur = t;
  }
  free(prev); free(cur);
  return h;
}
