// compat_args.cpp -- `class Parameters` (include/args_parser.h): the command-line front end GASAL2 callers use.
// Behaviour follows the reference's src/args_parser.cpp (defaults :12-29, help text :73-90, parse :93-240,
// file opening :242-250): same flags, same defaults, same failure messages and exit codes.
#include "../../include/args_parser.h"
#include <cstdlib>

Parameters::Parameters(int argc_, char** argv_)
    : sa(1), sb(4), gapo(6), gape(1), start_pos(WITHOUT_START), print_out(0), n_threads(1), k_band(0), secondBest(FALSE), isPacked(false),
      isReverseComplement(false), semiglobal_skipping_head(TARGET), semiglobal_skipping_tail(TARGET), algo(UNKNOWN), argc(argc_), argv(argv_) {}

Parameters::~Parameters() { query_batch_fasta.close(); target_batch_fasta.close(); }

void Parameters::print() {
  std::cerr << "sa=" << sa << " , sb=" << sb << " , gapo=" << gapo << " , gape=" << gape << std::endl;
  std::cerr << "start_pos=" << start_pos << " , print_out=" << print_out << " , n_threads=" << n_threads << std::endl;
  std::cerr << "semiglobal_skipping_head=" << semiglobal_skipping_head << " , semiglobal_skipping_tail=" << semiglobal_skipping_tail << " , algo=" << algo << std::endl;
  std::cerr << std::boolalpha << "isPacked = " << isPacked << " , secondBest = " << secondBest << std::endl;
  std::cerr << "query_batch_fasta_filename=" << query_batch_fasta_filename << " , target_batch_fasta_filename=" << target_batch_fasta_filename << std::endl;
}

void Parameters::failure(fail_type f) {
  switch (f) {
    case NOT_ENOUGH_ARGS: std::cerr << "Not enough Parameters. Required: -y AL_TYPE file1.fasta file2.fasta. See help (--help, -h) for usage. " << std::endl; break;
    case WRONG_ARG: std::cerr << "Wrong argument. See help (--help, -h) for usage. " << std::endl; break;
    case WRONG_FILES: std::cerr << "File error: either a file doesn't exist, or cannot be opened." << std::endl; break;
    default: break;
  }
  exit(1);
}

void Parameters::help() {
  std::cerr << "Usage: ./test_prog.out [-a] [-b] [-q] [-r] [-s] [-t] [-p] [-n] [-y] <query_batch.fasta> <target_batch.fasta>\n"
            << "Options: -a INT    match score [" << sa << "]\n"
            << "         -b INT    mismatch penalty [" << sb << "]\n"
            << "         -q INT    gap open penalty [" << gapo << "]\n"
            << "         -r INT    gap extension penalty [" << gape << "]\n"
            << "         -s        find the start position\n"
            << "         -t        compute traceback (the start position is then always computed)\n"
            << "         -p        print the alignment results\n"
            << "         -n INT    Number of threads [" << n_threads << "]\n"
            << "         -y AL_TYPE       Alignment type: \"local\", \"semi_global\", \"global\", \"ksw\", \"banded\"\n"
            << "         -x HEAD TAIL     for semi-global alignment: what may be skipped at the heads and tails (NONE, QUERY, TARGET, BOTH)\n"
            << "         -k INT    Band width for \"banded\"\n"
            << "         --help, -h : displays this message.\n"
            << "         --second-best   also report the second best score (WITHOUT_START only).\n"
            << "Single-pack multi-Parameters (e.g. -sp) is not supported." << std::endl;
}

static bool parse_source(const std::string& s, data_source& out) {
  if (s == "NONE") out = NONE; else if (s == "TARGET") out = TARGET; else if (s == "QUERY") out = QUERY; else if (s == "BOTH") out = BOTH; else return false;
  return true;
}

void Parameters::parse() {
  for (int c = 1; c < argc; c++) {
    const std::string a(argv[c]);
    if (a == "--help" || a == "-h") { help(); exit(0); }
  }
  if (argc < 4) failure(NOT_ENOUGH_ARGS);
  int c;
  for (c = 1; c < argc - 2; c++) {
    const std::string cur(argv[c]);
    if (cur.size() >= 2 && cur[0] == '-' && cur[1] == '-') {
      if (cur == "--second-best") secondBest = TRUE;
      continue;
    }
    if (cur.empty() || cur[0] != '-') failure(WRONG_ARG);
    if (cur.size() != 2) failure(WRONG_ARG);
    auto next = [&]() -> std::string { c++; if (c >= argc) failure(NOT_ENOUGH_ARGS); return std::string(argv[c]); };
    switch (cur[1]) {
      case 'y': {
        const std::string v = next();
        if (v == "local") algo = LOCAL; else if (v == "semi_global") algo = SEMI_GLOBAL; else if (v == "global") algo = GLOBAL;
        else if (v == "ksw") algo = KSW; else if (v == "banded") algo = BANDED;   // "banded" is an addition: the reference CLI cannot reach its banded kernel
        break;
      }
      case 'a': sa = std::stoi(next()); break;
      case 'b': sb = std::stoi(next()); break;
      case 'q': gapo = std::stoi(next()); break;
      case 'r': gape = std::stoi(next()); break;
      case 's': start_pos = WITH_START; break;
      case 't': start_pos = WITH_TB; break;
      case 'p': print_out = 1; break;
      case 'n': n_threads = std::stoi(next()); break;
      case 'k': k_band = std::stoi(next()); break;
      case 'x':
        if (!parse_source(next(), semiglobal_skipping_head)) failure(WRONG_ARG);
        if (!parse_source(next(), semiglobal_skipping_tail)) failure(WRONG_ARG);
        break;
      default: break;
    }
  }
  query_batch_fasta_filename = std::string(argv[c]);
  c++;
  target_batch_fasta_filename = std::string(argv[c]);
  fileopen();
}

void Parameters::fileopen() {
  query_batch_fasta.open(query_batch_fasta_filename, std::ifstream::in);
  if (!query_batch_fasta) failure(WRONG_FILES);
  target_batch_fasta.open(target_batch_fasta_filename);
  if (!target_batch_fasta) failure(WRONG_FILES);
}
