/* args_parser.h -- `class Parameters` with the public surface of the reference's src/args_parser.h:24-68
 * (same field names, defaults, CLI flags and failure messages), so that callers written against GASAL2
 * (e.g. its test_prog/test_prog.cpp) compile unchanged.  The library itself only consumes the POD image
 * returned by pod(); see gasal_b200.h. */
#ifndef ARGS_PARSER_H
#define ARGS_PARSER_H
#include <fstream>
#include <iostream>
#include <string>
#include <string.h>
#include "gasal.h"

enum fail_type { NOT_ENOUGH_ARGS, TOO_MANY_ARGS, WRONG_ARG, WRONG_FILES, WRONG_ALGO };

class Parameters {
 public:
  Parameters(int argc, char** argv);
  ~Parameters();
  void print();
  void failure(fail_type f);
  void help();
  void parse();
  void fileopen();

  /* the run-time options the library reads, as plain data */
  gasal_params_t pod() const {
    gasal_params_t p;
    p.algo = (int32_t)algo; p.start_pos = (int32_t)start_pos; p.secondBest = (int32_t)secondBest;
    p.isPacked = isPacked ? 1 : 0; p.isReverseComplement = isReverseComplement ? 1 : 0;
    p.semiglobal_skipping_head = (int32_t)semiglobal_skipping_head; p.semiglobal_skipping_tail = (int32_t)semiglobal_skipping_tail;
    p.k_band = k_band;
    return p;
  }

  int32_t sa;
  int32_t sb;
  int32_t gapo;
  int32_t gape;
  comp_start start_pos;
  int print_out;
  int n_threads;
  int32_t k_band;

  Bool secondBest;

  bool isPacked;
  bool isReverseComplement;

  data_source semiglobal_skipping_head;
  data_source semiglobal_skipping_tail;

  algo_type algo;

  std::string query_batch_fasta_filename;
  std::string target_batch_fasta_filename;

  std::ifstream query_batch_fasta;
  std::ifstream target_batch_fasta;

 private:
  int argc;
  char** argv;
};
#endif
