// ref_gpu_driver.cpp -- ORACLE / BASELINE infrastructure (never product code).
// Drives the UNMODIFIED reference library (Non-CDP/GASAL2, linked in by build_ref_gpu.sh) through its own
// public API on a batch file written by tests / bench, the way its example program does
// (test_prog/test_prog.cpp:203-437: NB_STREAMS storages, gasal_host_batch_fill per sequence,
// gasal_aln_async, poll gasal_is_aln_async_done), and writes every result array to a binary file.
//   ref_gpu_driver <in.bin> <out.bin> algo start second head tail k_band revcomp match mismatch gapo gape batch_size [reps]
// in.bin : u32 n, u32 q_bytes, u32 t_bytes, q_bytes[], t_bytes[], q_offs[n], t_offs[n], q_lens[n], t_lens[n], q_op[n] (u8), t_op[n] (u8)
// out.bin: u32 n, then i32[n] x 8 (score, q_end, t_end, q_start, t_start, score2, q_end2, t_end2), u32[n] n_cigar_ops,
//          u32 q_bytes, u8[q_bytes] cigar (laid out at the pair's global query offset)
// stdout : one JSON line with wall-clock and CUDA-event timings.
#include "gasal_header.h"
#include <chrono>
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>

struct Batch { uint32_t n, qb, tb; std::vector<uint8_t> q, t, qop, top; std::vector<uint32_t> qo, to, ql, tl; };

static bool load(const char* path, Batch& b) {
  FILE* f = fopen(path, "rb"); if (!f) return false;
  if (fread(&b.n, 4, 1, f) != 1 || fread(&b.qb, 4, 1, f) != 1 || fread(&b.tb, 4, 1, f) != 1) return false;
  b.q.resize(b.qb); b.t.resize(b.tb); b.qo.resize(b.n); b.to.resize(b.n); b.ql.resize(b.n); b.tl.resize(b.n); b.qop.resize(b.n); b.top.resize(b.n);
  bool ok = fread(b.q.data(), 1, b.qb, f) == b.qb && fread(b.t.data(), 1, b.tb, f) == b.tb && fread(b.qo.data(), 4, b.n, f) == b.n &&
            fread(b.to.data(), 4, b.n, f) == b.n && fread(b.ql.data(), 4, b.n, f) == b.n && fread(b.tl.data(), 4, b.n, f) == b.n &&
            fread(b.qop.data(), 1, b.n, f) == b.n && fread(b.top.data(), 1, b.n, f) == b.n;
  fclose(f); return ok;
}

int main(int argc, char** argv) {
  if (argc < 15) { fprintf(stderr, "usage: see source header\n"); return 2; }
  Batch b; if (!load(argv[1], b)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  char* fake_argv[] = {(char*)"ref_gpu_driver", NULL};
  Parameters* args = new Parameters(1, fake_argv);
  args->algo = (algo_type)atoi(argv[3]); args->start_pos = (comp_start)atoi(argv[4]); args->secondBest = (Bool)atoi(argv[5]);
  args->semiglobal_skipping_head = (data_source)atoi(argv[6]); args->semiglobal_skipping_tail = (data_source)atoi(argv[7]);
  args->k_band = atoi(argv[8]); args->isReverseComplement = atoi(argv[9]) != 0;
  args->sa = atoi(argv[10]); args->sb = atoi(argv[11]); args->gapo = atoi(argv[12]); args->gape = atoi(argv[13]);
  const uint32_t batch_size = (uint32_t)atoi(argv[14]);
  const int reps = argc > 15 ? atoi(argv[15]) : 1;
  gasal_subst_scores sub; sub.match = args->sa; sub.mismatch = args->sb; sub.gap_open = args->gapo; sub.gap_extend = args->gape;
  gasal_copy_subst_scores(&sub);
  uint32_t max_q = 0, max_t = 0;
  for (uint32_t i = 0; i < b.n; i++) { if (b.ql[i] > max_q) max_q = b.ql[i]; if (b.tl[i] > max_t) max_t = b.tl[i]; }
  const int NB = 2;
  gasal_gpu_storage_v v = gasal_init_gpu_storage_v(NB);
  gasal_init_streams(&v, max_q + 7, max_t + 7, batch_size, args);
  std::vector<int32_t> res[8]; for (auto& r : res) r.assign(b.n, 0);
  std::vector<uint32_t> ncig(b.n, 0); std::vector<uint8_t> cigar(b.qb, 0);
  cudaEvent_t e0[NB], e1[NB]; for (int i = 0; i < NB; i++) { cudaEventCreate(&e0[i]); cudaEventCreate(&e1[i]); }
  double ev_ms_total = 0, wall_best = 1e30, ev_best = 1e30;
  for (int rep = 0; rep < reps; rep++) {
    ev_ms_total = 0;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t start[NB] = {0, 0}, cnt[NB] = {0, 0}; bool busy[NB] = {false, false};
    uint32_t next = 0; int done_batches = 0, total_batches = (int)((b.n + batch_size - 1) / batch_size);
    while (done_batches < total_batches) {
      for (int s = 0; s < NB; s++) {
        gasal_gpu_storage_t* st = &v.a[s];
        if (!busy[s] && next < b.n) {
          uint32_t qi = 0, ti = 0, j = 0;
          for (uint32_t i = next; i < b.n && j < batch_size; i++, j++) {
            st->current_n_alns++;
            if (st->current_n_alns > st->host_max_n_alns) gasal_host_alns_resize(st, st->host_max_n_alns * 2, args);
            st->host_query_batch_offsets[j] = qi; st->host_target_batch_offsets[j] = ti;
            qi = gasal_host_batch_fill(st, qi, (const char*)&b.q[b.qo[i]], b.ql[i], QUERY);
            ti = gasal_host_batch_fill(st, ti, (const char*)&b.t[b.to[i]], b.tl[i], TARGET);
            st->host_query_batch_lens[j] = b.ql[i]; st->host_target_batch_lens[j] = b.tl[i];
          }
          gasal_op_fill(st, &b.qop[next], j, QUERY); gasal_op_fill(st, &b.top[next], j, TARGET);
          start[s] = next; cnt[s] = j; next += j;
          cudaEventRecord(e0[s], st->str);
          gasal_aln_async(st, qi, ti, j, args);
          cudaEventRecord(e1[s], st->str);
          busy[s] = true;
        }
        if (busy[s] && gasal_is_aln_async_done(st) == 0) {
          float ms = 0; cudaEventElapsedTime(&ms, e0[s], e1[s]); ev_ms_total += ms;
          gasal_res_t* h = st->host_res;
          for (uint32_t j = 0; j < cnt[s]; j++) {
            const uint32_t i = start[s] + j;
            if (h->aln_score) res[0][i] = h->aln_score[j];
            if (args->algo != GLOBAL) {
              if (h->query_batch_end) res[1][i] = h->query_batch_end[j];
              if (h->target_batch_end) res[2][i] = h->target_batch_end[j];
              if (args->start_pos != WITHOUT_START) { res[3][i] = h->query_batch_start[j]; res[4][i] = h->target_batch_start[j]; }
            }
            if (args->secondBest && st->host_res_second) {
              res[5][i] = st->host_res_second->aln_score[j];
              if (args->algo != GLOBAL) { res[6][i] = st->host_res_second->query_batch_end[j]; res[7][i] = st->host_res_second->target_batch_end[j]; }
            }
            if (args->start_pos == WITH_TB && (args->algo == LOCAL || args->algo == GLOBAL)) {
              ncig[i] = h->n_cigar_ops[j];
              const uint32_t off = st->host_query_batch_offsets[j];
              for (uint32_t u = 0; u < ncig[i] && b.qo[i] + u < b.qb; u++) cigar[b.qo[i] + u] = h->cigar[off + u];
            }
          }
          busy[s] = false; done_batches++;
        }
      }
    }
    double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (wall < wall_best) wall_best = wall;
    if (ev_ms_total < ev_best) ev_best = ev_ms_total;
  }
  FILE* f = fopen(argv[2], "wb");
  fwrite(&b.n, 4, 1, f);
  for (auto& r : res) fwrite(r.data(), 4, b.n, f);
  fwrite(ncig.data(), 4, b.n, f); fwrite(&b.qb, 4, 1, f); fwrite(cigar.data(), 1, b.qb, f);
  fclose(f);
  double cells = 0; for (uint32_t i = 0; i < b.n; i++) cells += (double)b.ql[i] * b.tl[i];
  printf("{\"impl\": \"reference-gpu\", \"max_query_len\": %d, \"n\": %u, \"batch_size\": %u, \"streams\": %d, \"reps\": %d, \"wall_ms\": %.3f, \"event_ms_sum\": %.3f, "
         "\"gcups_wall\": %.2f, \"gcups_events\": %.2f}\n", MAX_QUERY_LEN, b.n, batch_size, NB, reps, wall_best, ev_best, cells / wall_best / 1e6, cells / ev_best / 1e6);
  gasal_destroy_streams(&v, args); gasal_destroy_gpu_storage_v(&v);
  return 0;
}
