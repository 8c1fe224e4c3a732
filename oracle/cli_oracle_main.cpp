// cli_oracle_main.cpp — TEST-ONLY binary: the repository's host code (parameter parsing, index loading, FASTQ
// chunking, SAM/SJ/Log writers: star_b200/csrc/host/) driven by the CPU oracle instead of the CUDA engine.
// Used by tests/ to (a) pin the oracle + host code against the unmodified reference (oracle/_ref/STAR) and
// (b) produce expected outputs on machines without a GPU.  Never shipped, never linked into libstar_b200.so.
#include "star_oracle.h"
int main(int argc, char** argv) { return star_cli_main_engine(argc, argv, star_oracle_engine()); }
