// fasta_io.cpp -- the caller-side I/O step of the GASAL2 path, native and GPU-free:
//   * gasalx_read_fasta_pairs: two FASTA or FASTQ files (plain or gzip; the format is sniffed per file) -> the engine's host layout in one pass
//     (concatenated sequences padded with N to a multiple of 8, 64-bit offsets, lengths, per-sequence ops), ready for
//     gasalx_align_all or for slicing into gasal_host_batch_add.  The reference's only caller reads the files with a serial
//     getline loop into std::vector<std::string> and copies every sequence again (test_prog.cpp:90-137, 289-313).
//     Header conventions are the reference's (test_prog.cpp:79-89): the FIRST character of a header line selects the
//     operation: '>' forward, '<' reverse, '/' complement, '+' reverse-complement; sequences may span several lines.
//   * gasalx_pack_host: 4-bit packing on the host (same layout as the device kernel, pack_rc_seqs.h:13-53) for callers that
//     use isPacked storages -- halves the PCIe bytes and removes the pack kernel from the pipeline.
//   * gasalx_cigar_string: the (count << 2 | op) bytes of gasal_res_t::cigar (REVERSE alignment order) -> "12M1X3D..." with
//     adjacent equal ops merged, exactly what test_prog.cpp:382-428 prints.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/gasal_b200.h"

namespace {

struct Records {
  std::vector<uint8_t> bytes;          // padded, concatenated
  std::vector<uint64_t> offs;
  std::vector<uint32_t> lens;
  std::vector<uint8_t> ops;
  std::string names;                   // '\n'-separated header texts (without the first character)
};

bool slurp(const char* path, std::vector<char>& buf) {
  gzFile f = gzopen(path, "rb");       // reads plain files transparently
  if (!f) return false;
  gzbuffer(f, 1 << 20);
  size_t used = 0;
  buf.resize(1 << 22);
  for (;;) {
    if (buf.size() - used < (1u << 20)) buf.resize(buf.size() * 2);
    const int got = gzread(f, buf.data() + used, (unsigned)std::min<size_t>(buf.size() - used, 1u << 30));
    if (got < 0) { gzclose(f); return false; }
    if (got == 0) break;
    used += (size_t)got;
  }
  gzclose(f);
  buf.resize(used);
  return true;
}

void close_record(Records& r, uint64_t start, int n_code) {
  const uint64_t len = r.bytes.size() - start;
  r.lens.push_back((uint32_t)len);
  const uint32_t pad = (8u - (uint32_t)(len & 7u)) & 7u;          // gasal_host_batch_fill: pad with N_CODE to x8 (host_batch.cpp:100-102)
  r.bytes.insert(r.bytes.end(), pad, (uint8_t)n_code);
}

bool parse(const std::vector<char>& buf, Records& r, int n_code) {
  static const char starts[] = "></+";
  const char* p = buf.data(); const char* end = p + buf.size();
  bool open = false; uint64_t start = 0;
  r.bytes.reserve(buf.size() + buf.size() / 16);
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char* stop = le;
    if (stop > p && stop[-1] == '\r') stop--;
    if (stop > p) {
      const char* hit = strchr(starts, p[0]);
      if (hit && *hit) {
        if (open) close_record(r, start, n_code);
        r.ops.push_back((uint8_t)(hit - starts));
        r.offs.push_back(r.bytes.size());
        r.names.append(p + 1, (size_t)(stop - p - 1)); r.names.push_back('\n');
        start = r.bytes.size(); open = true;
      } else if (open) {
        r.bytes.insert(r.bytes.end(), (const uint8_t*)p, (const uint8_t*)stop);
      } else return false;                                         // sequence data before any header
    }
    p = nl ? nl + 1 : end;
  }
  if (open) close_record(r, start, n_code);
  return true;
}

// FASTQ: strict 4-line records  @name / sequence / + / qualities  (qualities are skipped: the aligner is quality-free)
bool parse_fastq(const std::vector<char>& buf, Records& r, int n_code) {
  const char* p = buf.data(); const char* end = p + buf.size();
  r.bytes.reserve(buf.size() / 2 + 64);
  int line_in_rec = 0;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const char* stop = le;
    if (stop > p && stop[-1] == '\r') stop--;
    if (line_in_rec == 0) {
      if (stop == p) { p = nl ? nl + 1 : end; continue; }           // blank line between records
      if (p[0] != '@') return false;
      r.ops.push_back(0); r.offs.push_back(r.bytes.size());
      r.names.append(p + 1, (size_t)(stop - p - 1)); r.names.push_back('\n');
    } else if (line_in_rec == 1) {
      const uint64_t start = r.bytes.size();
      r.bytes.insert(r.bytes.end(), (const uint8_t*)p, (const uint8_t*)stop);
      close_record(r, start, n_code);
    } else if (line_in_rec == 2) {
      if (stop == p || p[0] != '+') return false;
    }
    line_in_rec = (line_in_rec + 1) & 3;
    p = nl ? nl + 1 : end;
  }
  return line_in_rec == 0 || line_in_rec == 3 ? (r.lens.size() == r.offs.size()) : false;
}

bool looks_like_fastq(const std::vector<char>& buf) {
  for (char c : buf) { if (c == '\n' || c == '\r' || c == ' ') continue; return c == '@'; }
  return false;
}

template <typename T>
T* dup(const std::vector<T>& v) {
  T* p = (T*)malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
  if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

}  // namespace

extern "C" {

int gasalx_read_fasta_pairs(const char* query_path, const char* target_path, int n_code, gasalx_fasta_pairs_t* out) {
  memset(out, 0, sizeof(*out));
  std::vector<char> qb, tb;
  if (!slurp(query_path, qb)) { fprintf(stderr, "[GASAL ERROR:] cannot read %s\n", query_path); return -1; }
  if (!slurp(target_path, tb)) { fprintf(stderr, "[GASAL ERROR:] cannot read %s\n", target_path); return -1; }
  Records q, t;
  const bool okq = looks_like_fastq(qb) ? parse_fastq(qb, q, n_code) : parse(qb, q, n_code);
  const bool okt = looks_like_fastq(tb) ? parse_fastq(tb, t, n_code) : parse(tb, t, n_code);
  if (!okq || !okt) { fprintf(stderr, "[GASAL ERROR:] malformed FASTA / FASTQ input\n"); return -2; }
  if (q.lens.size() != t.lens.size()) {
    fprintf(stderr, "Batch1 and target_batch files should be fasta having same number of sequences\n");   // test_prog.cpp:131
    return -3;
  }
  out->n = q.lens.size();
  out->q = dup(q.bytes); out->t = dup(t.bytes); out->q_bytes = q.bytes.size(); out->t_bytes = t.bytes.size();
  out->q_offs = dup(q.offs); out->t_offs = dup(t.offs); out->q_lens = dup(q.lens); out->t_lens = dup(t.lens);
  out->q_op = dup(q.ops); out->t_op = dup(t.ops);
  out->q_names = (char*)malloc(q.names.size() + 1); memcpy(out->q_names, q.names.c_str(), q.names.size() + 1);
  out->t_names = (char*)malloc(t.names.size() + 1); memcpy(out->t_names, t.names.c_str(), t.names.size() + 1);
  uint32_t mq = 0, mt = 0; int any_op = 0;
  for (uint64_t i = 0; i < out->n; i++) { mq = std::max(mq, q.lens[i]); mt = std::max(mt, t.lens[i]); any_op |= q.ops[i] | t.ops[i]; }
  out->max_query_len = mq; out->max_target_len = mt; out->any_op = any_op != 0;
  return 0;
}

void gasalx_free_fasta_pairs(gasalx_fasta_pairs_t* p) {
  if (!p) return;
  free(p->q); free(p->t); free(p->q_offs); free(p->t_offs); free(p->q_lens); free(p->t_lens); free(p->q_op); free(p->t_op);
  free(p->q_names); free(p->t_names);
  memset(p, 0, sizeof(*p));
}

// gasalx_pack_host lives in host_pack.cpp (SIMD)

// returns the string length (excluding the terminator), or -(needed capacity) when `cap` is too small
int gasalx_cigar_string(const uint8_t* cigar, uint32_t n_ops, char* out, uint32_t cap) {
  static const char opc[4] = {'M', 'X', 'D', 'I'};
  std::string s;
  if (n_ops > 0) {
    int last_op = cigar[n_ops - 1] & 3;
    uint64_t count = cigar[n_ops - 1] >> 2;
    for (int64_t u = (int64_t)n_ops - 2; u >= 0; u--) {
      const int op = cigar[u] & 3;
      if (op == last_op) count += cigar[u] >> 2;
      else { s += std::to_string(count); s.push_back(opc[last_op]); count = cigar[u] >> 2; }
      last_op = op;
    }
    s += std::to_string(count); s.push_back(opc[last_op]);
  }
  if (s.size() + 1 > cap) return -(int)(s.size() + 1);
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// SAM v1 mandatory fields + AS/XS tags; RNEXT/PNEXT/TLEN are "*", 0, 0 (single-end)
int gasalx_sam_record(char* out, uint32_t cap, const char* qname, int flag, const char* rname, int64_t pos1, int mapq,
                      const char* cigar, const char* seq, uint32_t seq_len, const char* qual, int32_t score, int32_t score2) {
  std::string s;
  s += (qname && *qname) ? qname : "*"; s.push_back('\t');
  s += std::to_string(flag); s.push_back('\t');
  s += (rname && *rname) ? rname : "*"; s.push_back('\t');
  s += std::to_string((long long)pos1); s.push_back('\t');
  s += std::to_string(mapq); s.push_back('\t');
  s += (cigar && *cigar) ? cigar : "*"; s += "\t*\t0\t0\t";
  if (seq && seq_len) s.append(seq, seq_len); else s.push_back('*');
  s.push_back('\t');
  if (qual && seq_len) s.append(qual, seq_len); else s.push_back('*');
  s += "\tAS:i:" + std::to_string(score);
  if (score2 >= 0) s += "\tXS:i:" + std::to_string(score2);
  if (s.size() + 1 > cap) return -(int)(s.size() + 1);
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // extern "C"
