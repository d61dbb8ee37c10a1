// batcher.cpp -- host batcher / multi-GPU sharder (north_star subsystem 1).
//
// The reference leaves batching to its caller (test_prog.cpp:203-437 splits the input equally over OpenMP
// threads and fills fixed 5000-pair batches in input order) and has no multi-GPU mode.  This module is the
// library-side replacement: given ALL pairs of a job it
//   1. length-bins them (sort by padded query/target length class) so that the threads of a warp walk
//      matrices of the same shape and whole batches qualify for the s16x2 fast path,
//   2. cuts the sorted order into batches bounded in pairs, bytes and DP cells,
//   3. assigns batches to GPUs by LPT (largest processing time first) on the number of DP cells,
//   4. runs one host worker thread per GPU, each with its own storages/streams, streaming batches with
//      pinned cudaMemcpyAsync through gasal_aln_async / gasal_is_aln_async_done, and
//   5. scatters the results back to the caller's original pair order on the host (no device collective:
//      pairs are independent).
// Steps 1-3 are pure host code (gasalx_plan_*), unit-tested without a GPU.
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/gasal_b200.h"

namespace {

struct Plan {
  std::vector<uint32_t> order;         // pair indices in batch order
  std::vector<uint32_t> batch_start;   // n_batches + 1 offsets into `order`
  std::vector<uint64_t> batch_cells;
  std::vector<int32_t> batch_gpu;
  std::vector<uint32_t> batch_max_q, batch_max_t;
  std::vector<uint64_t> batch_q_bytes, batch_t_bytes;
};

inline uint32_t pad8(uint32_t x) { return (x + 7u) & ~7u; }

// Defaults chosen from measurements (1 M pairs of 150 x 600 on one B200, 16 host cores, tools/bench_batcher_sweep.py): staging a batch
// into the pinned pages (read ASCII, pack, write) is host-memory work of ~10 GB/s per thread, a 131072-pair batch takes ~10 ms
// to stage and 3 ms to align, so throughput comes from MANY small batches in flight on many storages, filled by several threads:
// 4 storages x 4 threads x 131072 pairs: 47.8 ms; 16 storages x 8 threads x 32768 pairs: 30.4 ms (the kernels alone: 19.3 ms).
inline int auto_streams(const gasalx_sched_t* sc) { return sc->streams_per_gpu > 0 ? sc->streams_per_gpu : 16; }
inline int auto_threads(const gasalx_sched_t* sc, int ns) {
  int t = sc->host_threads_per_gpu;
  if (t <= 0) {
    const int hw = (int)std::thread::hardware_concurrency(), g = sc->n_gpus > 0 ? sc->n_gpus : 1;
    t = std::max(2, (hw > 0 ? hw : 8) / (2 * g));
    t = std::min(t, 8);
  }
  return std::max(1, std::min(t, ns));
}
inline uint32_t auto_batch_pairs(const gasalx_sched_t* sc, uint64_t n) {
  if (sc->max_batch_pairs) return sc->max_batch_pairs;
  const uint64_t g = sc->n_gpus > 0 ? sc->n_gpus : 1, per = (n + g * 2 * auto_streams(sc) - 1) / (g * 2 * auto_streams(sc));   // two rounds per storage
  return (uint32_t)std::min<uint64_t>(262144, std::max<uint64_t>(16384, (per + 4095) & ~4095ull));
}

void make_plan(const uint32_t* q_lens, const uint32_t* t_lens, uint64_t n, const gasalx_sched_t* sc, Plan& pl) {
  const uint32_t max_pairs = auto_batch_pairs(sc, n);
  const uint64_t max_bytes = sc->max_batch_bytes ? sc->max_batch_bytes : (1ull << 30);
  const uint64_t max_cells = sc->max_batch_cells ? sc->max_batch_cells : (1ull << 36);
  const int n_gpus = sc->n_gpus > 0 ? sc->n_gpus : 1;
  pl.order.resize(n);
  std::iota(pl.order.begin(), pl.order.end(), 0u);
  // 1. length binning: class = (padded query words, padded target words); stable so equal shapes keep input order
  std::vector<uint64_t> key(n);
  for (uint64_t i = 0; i < n; i++) key[i] = ((uint64_t)(pad8(q_lens[i]) >> 3) << 32) | (uint64_t)(pad8(t_lens[i]) >> 3);
  bool uniform = true;
  for (uint64_t i = 1; i < n && uniform; i++) uniform = (key[i] == key[0]);
  if (!sc->keep_input_order && !uniform)
    std::stable_sort(pl.order.begin(), pl.order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  // 2. cut into batches; never mix pairs on different sides of the fast-path query limit
  const uint32_t fast_q8 = 352;   // == kFastMaxQ8 (dp_s16.cuh)
  pl.batch_start.push_back(0);
  uint64_t cells = 0, qb = 0, tb = 0; uint32_t cnt = 0, mq = 0, mt = 0;
  auto close = [&](uint64_t pos) {
    if (cnt == 0) return;
    pl.batch_start.push_back((uint32_t)pos); pl.batch_cells.push_back(cells); pl.batch_max_q.push_back(mq); pl.batch_max_t.push_back(mt);
    pl.batch_q_bytes.push_back(qb); pl.batch_t_bytes.push_back(tb);
    cells = qb = tb = 0; cnt = 0; mq = mt = 0;
  };
  for (uint64_t k = 0; k < n; k++) {
    const uint32_t i = pl.order[k];
    const uint64_t q8 = pad8(q_lens[i]), t8 = pad8(t_lens[i]);
    const bool crosses = cnt > 0 && ((pad8(mq) <= fast_q8) != (q8 <= fast_q8));
    // long pairs run on the wavefront kernel, whose boundary double buffer takes 32 B per query base of the batch
    const bool wave_full = q8 > fast_q8 && qb + q8 > (64ull << 20);
    if (cnt >= max_pairs || qb + q8 > max_bytes || tb + t8 > max_bytes || qb + q8 > 0xFFFFFFF0ull || tb + t8 > 0xFFFFFFF0ull || (cnt > 0 && wave_full) ||
        (cnt > 0 && cells + (uint64_t)q_lens[i] * t_lens[i] > max_cells) || crosses)
      close(k);
    cells += (uint64_t)q_lens[i] * t_lens[i]; qb += q8; tb += t8; cnt++;
    mq = std::max(mq, q_lens[i]); mt = std::max(mt, t_lens[i]);
  }
  close(n);
  // 3. LPT over GPUs on DP cells
  const size_t nb = pl.batch_cells.size();
  pl.batch_gpu.assign(nb, 0);
  std::vector<uint32_t> by_size(nb);
  std::iota(by_size.begin(), by_size.end(), 0u);
  std::stable_sort(by_size.begin(), by_size.end(), [&](uint32_t a, uint32_t b) { return pl.batch_cells[a] > pl.batch_cells[b]; });
  std::vector<uint64_t> load(n_gpus, 0);
  for (uint32_t b : by_size) {
    int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    pl.batch_gpu[b] = g; load[g] += pl.batch_cells[b];
  }
}

// Storages (pinned host pages, device buffers, streams) are expensive to create -- a GB of pinned memory takes a few hundred
// milliseconds -- so each device keeps the set of its last job and reuses it while the parameters match and the capacity
// suffices.  One job at a time per device (the mutex is held for the duration of a worker).
struct DeviceCache {
  std::mutex mu;
  bool valid = false;
  gasal_gpu_storage_v v;
  gasal_params_t params;
  int max_q = 0, max_t = 0, max_pairs = 0;
};
DeviceCache g_cache[64];

bool same_params(const gasal_params_t& a, const gasal_params_t& b) {
  return a.algo == b.algo && a.start_pos == b.start_pos && a.secondBest == b.secondBest && a.isPacked == b.isPacked &&
         a.isReverseComplement == b.isReverseComplement;
}

struct Job {
  const uint8_t* q; const uint8_t* t; const uint64_t* q_offs; const uint64_t* t_offs; const uint32_t* q_lens; const uint32_t* t_lens;
  const uint8_t* q_op; const uint8_t* t_op; const uint32_t* seeds;
  const gasal_params_t* params; gasalx_results_t* out;
};

void worker(int gpu_slot, int device, const Plan& pl, const Job& job, const gasalx_sched_t* sc, std::atomic<int>* err, double* gpu_ms) {
  if (cudaSetDevice(device) != cudaSuccess) { err->store(1); return; }
  std::vector<uint32_t> mine;
  uint32_t max_pairs = 0, mq = 0, mt = 0;
  for (uint32_t b = 0; b < pl.batch_gpu.size(); b++)
    if (pl.batch_gpu[b] == gpu_slot) {
      mine.push_back(b);
      max_pairs = std::max(max_pairs, pl.batch_start[b + 1] - pl.batch_start[b]);
      mq = std::max(mq, pl.batch_max_q[b]); mt = std::max(mt, pl.batch_max_t[b]);
    }
  if (mine.empty()) return;
  // process big batches first so the tail is short
  std::stable_sort(mine.begin(), mine.end(), [&](uint32_t a, uint32_t b) { return pl.batch_cells[a] > pl.batch_cells[b]; });
  const int ns = std::max(1, std::min<int>(auto_streams(sc), (int)mine.size()));      // never more storages than batches
  // size the pages for the largest batch in BYTES (ragged batches are much smaller than max_pairs x max_len)
  uint64_t maxqb = 0, maxtb = 0;
  for (uint32_t b : mine) { maxqb = std::max(maxqb, pl.batch_q_bytes[b]); maxtb = std::max(maxtb, pl.batch_t_bytes[b]); }
  DeviceCache& dc = g_cache[device & 63];
  std::lock_guard<std::mutex> lock(dc.mu);
  const bool fits = dc.valid && dc.v.n == ns && same_params(dc.params, *job.params) && (int)max_pairs <= dc.max_pairs &&
                    (uint64_t)dc.max_pairs * dc.max_q >= maxqb && (uint64_t)dc.max_pairs * dc.max_t >= maxtb;
  if (!fits) {
    // Capacity is tracked in BYTES per side and in pairs, each merged on its own with what the cache already held: a long-read job
    // (few pairs x 100 kb) followed by a short-read job (262144 pairs x 150 b) must not ask for 262144 x 100 kb.
    uint64_t qbytes_cap = std::max<uint64_t>(maxqb, 8), tbytes_cap = std::max<uint64_t>(maxtb, 8);
    int np = (int)max_pairs;
    if (dc.valid) {
      if (dc.v.n == ns && same_params(dc.params, *job.params)) {
        qbytes_cap = std::max(qbytes_cap, (uint64_t)dc.max_pairs * dc.max_q); tbytes_cap = std::max(tbytes_cap, (uint64_t)dc.max_pairs * dc.max_t);
        np = std::max(np, dc.max_pairs);
      }
      gasal_destroy_streams(&dc.v, &dc.params); gasal_destroy_gpu_storage_v(&dc.v); dc.valid = false;
    }
    const int nq = (int)((qbytes_cap + (uint64_t)np - 1) / (uint64_t)np) + 7, nt = (int)((tbytes_cap + (uint64_t)np - 1) / (uint64_t)np) + 7;
    dc.v = gasal_init_gpu_storage_v(ns);
    if (gasalx_try_init_streams(&dc.v, nq, nt, np, job.params) != GASALX_OK) {
      fprintf(stderr, "[GASAL ERROR:] gasalx_align_all: cannot create the storages of device %d: %s", device, gasalx_last_error());
      gasal_destroy_gpu_storage_v(&dc.v); err->store(1); return;
    }
    dc.params = *job.params; dc.max_q = ((nq + 7) / 8) * 8; dc.max_t = ((nt + 7) / 8) * 8; dc.max_pairs = np; dc.valid = true;
  }
  gasal_gpu_storage_v& v = dc.v;
  std::vector<int> running(ns, -1);
  // Staging (one bulk copy per run of neighbouring pairs into the pinned pages) and the scatter of the results are host
  // memory traffic: one thread per GPU moves ~11 GB/s, less than a B200 consumes on short reads.  The storages of a GPU are
  // therefore split over `nthreads` host threads; each owns its storages (and their streams) and pulls batches from one queue.
  const int nthreads = auto_threads(sc, ns);
  std::atomic<size_t> next(0), done(0);
  std::mutex ms_mu;
  double ms_total = 0;
  const gasal_params_t* P = job.params;
  auto run_storages = [&](int s_lo, int s_hi) {
  auto harvest = [&](int s) {
    gasal_gpu_storage_t* st = &v.a[s];
    const uint32_t b = (uint32_t)running[s];
    const uint32_t lo = pl.batch_start[b], cnt = pl.batch_start[b + 1] - lo;
    float ms[3]; if (gasalx_last_batch_ms(st, ms) == 0) { std::lock_guard<std::mutex> g(ms_mu); ms_total += ms[0]; }
    gasal_res_t* h = st->host_res; gasalx_results_t* o = job.out;
    for (uint32_t j = 0; j < cnt; j++) {
      const uint32_t i = pl.order[lo + j];
      if (o->aln_score && h->aln_score) o->aln_score[i] = h->aln_score[j];
      if (o->query_batch_end && h->query_batch_end) o->query_batch_end[i] = h->query_batch_end[j];
      if (o->target_batch_end && h->target_batch_end) o->target_batch_end[i] = h->target_batch_end[j];
      if (o->query_batch_start && h->query_batch_start) o->query_batch_start[i] = h->query_batch_start[j];
      if (o->target_batch_start && h->target_batch_start) o->target_batch_start[i] = h->target_batch_start[j];
    }
    if (P->secondBest && st->host_res_second) {
      gasal_res_t* h2 = st->host_res_second;
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t i = pl.order[lo + j];
        if (o->aln_score2 && h2->aln_score) o->aln_score2[i] = h2->aln_score[j];
        if (o->query_batch_end2 && h2->query_batch_end) o->query_batch_end2[i] = h2->query_batch_end[j];
        if (o->target_batch_end2 && h2->target_batch_end) o->target_batch_end2[i] = h2->target_batch_end[j];
      }
    }
    if (P->start_pos == WITH_TB && o->cigar && o->n_cigar_ops && h->cigar && h->n_cigar_ops) {
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t i = pl.order[lo + j];
        const uint32_t nops = h->n_cigar_ops[j];
        o->n_cigar_ops[i] = nops;
        // the caller's CIGAR buffer mirrors its query buffer: pair i owns pad8(q_len) bytes at q_offs[i] (reference Q6)
        const uint32_t room = pad8(job.q_lens[i]);
        memcpy(o->cigar + job.q_offs[i], h->cigar + st->host_query_batch_offsets[j], std::min(nops, room));
      }
    }
    running[s] = -1; done.fetch_add(1);
  };
  for (;;) {
    bool progressed = false, busy = false;
    for (int s = s_lo; s < s_hi; s++) {
      gasal_gpu_storage_t* st = &v.a[s];
      if (running[s] >= 0) {
        if (gasal_is_aln_async_done(st) == 0) { harvest(s); progressed = true; } else busy = true;
        continue;
      }
      const size_t take = next.fetch_add(1);
      if (take >= mine.size()) continue;
      const uint32_t b = mine[take];
      const uint32_t lo = pl.batch_start[b], cnt = pl.batch_start[b + 1] - lo;
      if (cnt > st->host_max_n_alns) gasal_host_alns_resize(st, (int)cnt, P);
      uint32_t qi = 0, ti = 0;
      for (uint32_t j = 0; j < cnt;) {
        // the job is already in the padded host layout: pairs that are neighbours in the input AND in the batch are staged
        // with ONE bulk copy per sequence buffer instead of one gasal_host_batch_fill per sequence
        const uint32_t i0 = pl.order[lo + j];
        uint32_t run = 1;
        while (j + run < cnt) {
          const uint32_t ip = pl.order[lo + j + run - 1], in = pl.order[lo + j + run];
          if (in != ip + 1 || job.q_offs[in] != job.q_offs[ip] + pad8(job.q_lens[ip]) || job.t_offs[in] != job.t_offs[ip] + pad8(job.t_lens[ip])) break;
          run++;
        }
        const uint32_t il = i0 + run - 1;
        const uint64_t qbytes = job.q_offs[il] + pad8(job.q_lens[il]) - job.q_offs[i0], tbytes = job.t_offs[il] + pad8(job.t_lens[il]) - job.t_offs[i0];
        for (uint32_t r = 0; r < run; r++) {
          const uint32_t i = i0 + r;
          st->host_query_batch_offsets[j + r] = qi + (uint32_t)(job.q_offs[i] - job.q_offs[i0]);
          st->host_target_batch_offsets[j + r] = ti + (uint32_t)(job.t_offs[i] - job.t_offs[i0]);
          st->host_query_batch_lens[j + r] = job.q_lens[i]; st->host_target_batch_lens[j + r] = job.t_lens[i];
          if (P->isReverseComplement) {          // cached storages keep the op bytes of earlier jobs: a side without ops is written as 0
            st->host_query_op[j + r] = job.q_op ? job.q_op[i] : 0;
            st->host_target_op[j + r] = job.t_op ? job.t_op[i] : 0;
          }
          if (job.seeds && st->host_seed_scores) st->host_seed_scores[j + r] = job.seeds[i];
        }
        if (run > 1) {
          qi = gasal_host_batch_add(st, qi, (const char*)job.q + job.q_offs[i0], (uint32_t)qbytes, QUERY);
          ti = gasal_host_batch_add(st, ti, (const char*)job.t + job.t_offs[i0], (uint32_t)tbytes, TARGET);
        } else {
          qi = gasal_host_batch_fill(st, qi, (const char*)job.q + job.q_offs[i0], job.q_lens[i0], QUERY);
          ti = gasal_host_batch_fill(st, ti, (const char*)job.t + job.t_offs[i0], job.t_lens[i0], TARGET);
        }
        j += run;
      }
      st->current_n_alns = cnt;
      gasal_aln_async(st, qi, ti, cnt, P);
      running[s] = (int)b; progressed = true; busy = true;
    }
    if (!busy && next.load() >= mine.size()) break;          // nothing in flight on my storages and the queue is empty
    if (!progressed) std::this_thread::yield();
  }
  };   // run_storages
  std::vector<std::thread> helpers;
  for (int t = 1; t < nthreads; t++)
    helpers.emplace_back([&, t] { if (cudaSetDevice(device) != cudaSuccess) { err->store(1); return; } run_storages(t * ns / nthreads, (t + 1) * ns / nthreads); });
  run_storages(0, ns / nthreads);
  for (auto& h : helpers) h.join();
  if (gpu_ms) *gpu_ms = ms_total;
}

}  // namespace

extern "C" {

void gasalx_sched_default(gasalx_sched_t* s) {
  memset(s, 0, sizeof(*s));
  s->n_gpus = 1; s->streams_per_gpu = 0; s->max_batch_pairs = 0; s->max_batch_bytes = 1ull << 30; s->max_batch_cells = 1ull << 36;   // 0 = automatic
  s->host_threads_per_gpu = 0;
}

// Pure host planning (no CUDA): returns the number of batches; fills order[n], and, for up to `cap` batches,
// batch_start[nb+1], batch_gpu[nb], batch_cells[nb].  Pass NULL outputs to just count.
int64_t gasalx_plan_batches(const uint32_t* q_lens, const uint32_t* t_lens, uint64_t n, const gasalx_sched_t* sched,
                            uint32_t* order, uint32_t* batch_start, int32_t* batch_gpu, uint64_t* batch_cells, uint64_t cap) {
  Plan pl; make_plan(q_lens, t_lens, n, sched, pl);
  const uint64_t nb = pl.batch_cells.size();
  if (order) memcpy(order, pl.order.data(), n * sizeof(uint32_t));
  const uint64_t m = std::min<uint64_t>(nb, cap);
  if (batch_start) memcpy(batch_start, pl.batch_start.data(), (m + (m == nb ? 1 : 0)) * sizeof(uint32_t));
  if (batch_gpu) memcpy(batch_gpu, pl.batch_gpu.data(), m * sizeof(int32_t));
  if (batch_cells) memcpy(batch_cells, pl.batch_cells.data(), m * sizeof(uint64_t));
  return (int64_t)nb;
}

// Align ALL pairs of a job on sched->n_gpus GPUs of this process; results land in the caller's order.
// Sequences: concatenated, each padded to a multiple of 8 bytes (gasal_host_batch_fill layout) at 64-bit byte offsets.
// Returns 0, or -1 on a CUDA/device error.  gpu_ms (optional, n_gpus entries): device time per GPU (CUDA events).
int gasalx_align_all(const uint8_t* q_bytes, const uint8_t* t_bytes, const uint64_t* q_offs, const uint64_t* t_offs,
                     const uint32_t* q_lens, const uint32_t* t_lens, const uint8_t* q_op, const uint8_t* t_op, const uint32_t* seed_scores,
                     uint64_t n, const gasal_params_t* params, const gasalx_sched_t* sched, gasalx_results_t* out, double* gpu_ms) {
  if (n == 0) return 0;
  if (params->isPacked) { fprintf(stderr, "[GASAL ERROR:] gasalx_align_all takes unpacked (ASCII) sequences; pre-packed input goes through gasal_aln_async\n"); return -2; }
  if (n > 0xFFFFFFF0ull) { fprintf(stderr, "[GASAL ERROR:] gasalx_align_all: more than 2^32 pairs in one job\n"); return -1; }
  Plan pl; make_plan(q_lens, t_lens, n, sched, pl);
  Job job = {q_bytes, t_bytes, q_offs, t_offs, q_lens, t_lens, q_op, t_op, seed_scores, params, out};
  const int n_gpus = sched->n_gpus > 0 ? sched->n_gpus : 1;
  int have = 0; cudaGetDeviceCount(&have);
  if (have <= 0) { fprintf(stderr, "[GASAL CUDA ERROR:] no CUDA device (there is no CPU fallback)\n"); exit(EXIT_FAILURE); }
  std::atomic<int> err(0);
  std::vector<std::thread> th;
  std::vector<double> ms(n_gpus, 0.0);
  for (int g = 0; g < n_gpus; g++) {
    const int dev = sched->gpu_ids ? sched->gpu_ids[g] : (g % have);
    th.emplace_back(worker, g, dev, std::cref(pl), std::cref(job), sched, &err, &ms[g]);
  }
  for (auto& t : th) t.join();
  if (gpu_ms) memcpy(gpu_ms, ms.data(), sizeof(double) * n_gpus);
  return err.load() ? -1 : 0;
}

// Free the per-device storage caches of gasalx_align_all (also safe to never call: process exit releases them).
void gasalx_batcher_release(void) {
  for (int d = 0; d < 64; d++) {
    DeviceCache& dc = g_cache[d];
    std::lock_guard<std::mutex> lock(dc.mu);
    if (!dc.valid) continue;
    if (cudaSetDevice(d) == cudaSuccess) { gasal_destroy_streams(&dc.v, &dc.params); gasal_destroy_gpu_storage_v(&dc.v); }
    dc.valid = false;
  }
}

}  // extern "C"
