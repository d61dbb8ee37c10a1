// nvb_cpp_example.cpp -- the call sequence of NvB/sw-benchmark/sw-benchmark.cu:366-383 (batch_score_profile) written against
// include/nvbio_b200_batched.hpp: read a job file, run Gotoh global / semi-global / local through BatchedAlignmentScore::enact,
// print "type score sink_text sink_pattern" per job.  Built by __graft_entry__.build() into build/nvb_cpp_example; used by
// tests/test_nvbio_gpu.py::test_cpp_front_end (job file: u32 n_patterns, u32 n_texts, offsets[n_patterns+1], offsets[n_texts+1], bytes).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "nvbio_b200_batched.hpp"

using namespace nvbio_b200;

template <typename aligner_type>
static void batch_score_profile(const aligner_type aligner, uint32_t n, const std::vector<uint8_t>& p, const std::vector<uint32_t>& po,
                                const std::vector<uint8_t>& t, const std::vector<uint32_t>& to, int type_id) {
  typedef aln::HostStringStream<aligner_type> stream_type;
  typedef aln::BatchedAlignmentScore<stream_type, aln::DeviceThreadScheduler> batch_type;     // our batch type
  std::vector<int32_t> score(n), si(n), sj(n);
  stream_type stream(aligner, n, p.data(), po.data(), (uint32_t)po.size() - 1, t.data(), to.data(), (uint32_t)to.size() - 1, score.data(), si.data(), sj.data());
  const uint64_t temp_size = batch_type::max_temp_storage(stream.max_pattern_length(), stream.max_text_length(), stream.size());
  batch_type batch;
  batch.enact(stream, temp_size, NULL);
  for (uint32_t i = 0; i < n; ++i) printf("%d %d %d %d\n", type_id, score[i], si[i], sj[i]);
}

// the reference's traceback batch (batched.h:436-451): prints "type score source_text source_pattern sink_text sink_pattern n_ops" and the ops
template <typename aligner_type>
static void batch_traceback(const aligner_type aligner, uint32_t n, const std::vector<uint8_t>& p, const std::vector<uint32_t>& po,
                            const std::vector<uint8_t>& t, const std::vector<uint32_t>& to, int type_id) {
  typedef aln::HostTracebackStream<aligner_type> stream_type;
  typedef aln::BatchedAlignmentTraceback<32u, stream_type> batch_type;
  std::vector<nvbx_alignment_t> alns(n);
  stream_type probe(aligner, n, p.data(), po.data(), (uint32_t)po.size() - 1, t.data(), to.data(), (uint32_t)to.size() - 1, alns.data(), NULL, 0);
  const uint32_t stride = probe.ops_stride();
  std::vector<uint8_t> ops((size_t)n * stride);
  stream_type stream(aligner, n, p.data(), po.data(), (uint32_t)po.size() - 1, t.data(), to.data(), (uint32_t)to.size() - 1, alns.data(), ops.data(), stride);
  batch_type batch;
  batch.enact(stream);
  for (uint32_t i = 0; i < n; ++i) {
    printf("%d %d %d %d %d %d %u ", type_id, alns[i].score, alns[i].source_text, alns[i].source_pattern, alns[i].sink_text, alns[i].sink_pattern, alns[i].n_ops);
    for (uint32_t k = 0; k < alns[i].n_ops; ++k) putchar("MID"[ops[(size_t)i * stride + k]]);
    putchar('\n');
  }
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s jobs.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  uint32_t np = 0, nt = 0;
  if (fread(&np, 4, 1, f) != 1 || fread(&nt, 4, 1, f) != 1) return 2;
  std::vector<uint32_t> po(np + 1), to(nt + 1);
  if (fread(po.data(), 4, np + 1, f) != np + 1 || fread(to.data(), 4, nt + 1, f) != nt + 1) return 2;
  std::vector<uint8_t> p(po[np]), t(to[nt]);
  if (fread(p.data(), 1, p.size(), f) != p.size() || fread(t.data(), 1, t.size(), f) != t.size()) return 2;
  fclose(f);
  try {
    aln::SimpleGotohScheme scoring;                      // sw-benchmark.cu:585-589
    scoring.m_match = 2; scoring.m_mismatch = -1; scoring.m_gap_open = -2; scoring.m_gap_ext = -1;
    batch_score_profile(aln::make_gotoh_aligner<aln::GLOBAL>(scoring), np, p, po, t, to, 0);
    batch_score_profile(aln::make_gotoh_aligner<aln::SEMI_GLOBAL>(scoring), np, p, po, t, to, 2);
    batch_score_profile(aln::make_gotoh_aligner<aln::LOCAL>(scoring), np, p, po, t, to, 1);
    batch_score_profile(aln::make_edit_distance_aligner<aln::SEMI_GLOBAL>(), np, p, po, t, to, 3);
    if (argc > 2) batch_traceback(aln::make_gotoh_aligner<aln::LOCAL>(scoring), np, p, po, t, to, 4);     // any second argument: also trace the local alignments back
  } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
