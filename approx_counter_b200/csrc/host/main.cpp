// main.cpp — `approx_counter`, the drop-in binary (see cli.cpp).
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "host_util.h"

int main(int argc, const char **argv) {
    apch::one_shot_process = true;
    const int rc = apch::cli_main(argc, argv);
    // every output file is closed by now; skip the CUDA runtime's exit handlers (about
    // 0.7 s of context teardown that the driver does anyway when the process ends)
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    std::quick_exit(rc);
}
