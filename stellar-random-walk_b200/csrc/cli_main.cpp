// stellar-rw -- native stand-in for `spark-submit --class au.csiro.data61.randomwalk.Main <jar>` with the
// same argv (README.md:30-34, Main.scala:18-27): stellar-rw --cmd randomwalk --input <edges> --output <dir> ...
#include "../../include/srw.h"

int main(int argc, char **argv) { return srw_main(argc - 1, (const char *const *)(argv + 1)); }
