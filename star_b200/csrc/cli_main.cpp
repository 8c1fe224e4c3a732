// cli_main.cpp — the drop-in `STAR` executable: every alignment is computed by the CUDA engine inside libstar_b200.so.
#include "../../include/star_b200.h"
int main(int argc, char** argv) { return star_cli_main(argc, argv); }
