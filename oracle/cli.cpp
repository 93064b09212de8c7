// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product (see rng.hpp header).
// Command-line twin of `java -jar PedigreeSim.jar x.par` (Main.java:47-58) over the restatement:
//   psim_oracle_cli x.par [--philox]
#include <cstdio>
#include <cstring>

#include "popio.hpp"
#include "device_twin.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::printf("PedigreeSim: parameter file PedigreeSim.par not found or not readable.\n"); return 0; }
    int rng_mode = 0;
    for (int i = 2; i < argc; i++) if (!std::strcmp(argv[i], "--philox")) rng_mode = 1;
    std::string log;
    orc::run_parameter_file(argv[1], rng_mode, log);
    std::fputs(log.c_str(), stdout);
    return 0;  // the reference always exits 0 (Main.java:106-109,446-448)
}
