// main.cpp — `approx_counter`, the drop-in binary (see cli.cpp).
#include "host_util.h"

int main(int argc, const char **argv) { return apch::cli_main(argc, argv); }
