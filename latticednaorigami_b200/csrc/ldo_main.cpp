// ldo_main.cpp — command-line driver with the reference's interface (apps/main.cpp:19-117):
//   latticeDNAOrigami_b200 -i file.inp [--replicas N] [--device D]
// Every replica runs the simulation described by the parameter file on one warp of the GPU.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../../include/ldo_host.h"

int main(int argc, char* argv[]) {
    std::string inp;
    int replicas = 1, device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a {argv[i]};
        if ((a == "-i" || a == "--parameter_filename") && i + 1 < argc) inp = argv[++i];
        else if (a == "--replicas" && i + 1 < argc) replicas = std::atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "-h" || a == "--help") {
            std::cout << "\nAllowed options:\n  -i, --parameter_filename  Input file\n  --replicas N              replicas run concurrently (one warp each)\n  --device D                CUDA device\n\n";
            return 1;
        }
        else if (a == "-v" || a == "--version") {
            std::cout << "\nlatticeDNAOrigami_b200\n\n";
            return 1;
        }
    }
    if (inp.empty()) {
        std::cout << "Input parameter file must be provided" << std::endl << "Run with -h to see all options" << std::endl << std::endl;
        return 1;
    }
    ldo_sim* sim = ldo_sim_create(inp.c_str(), replicas, device, 0, 1);
    int rc = sim ? ldo_sim_run(sim) : -1;
    if (rc != 0) {
        // same failure report as apps/main.cpp:105-115
        std::cout << std::endl << "An exception occurred during the run" << std::endl << std::endl;
        std::cerr << ldo_host_last_error() << std::endl;
        std::cout << std::endl << "Ending run unsuccesfully" << std::endl << std::endl;
        if (sim) ldo_sim_destroy(sim);
        return EXIT_FAILURE;
    }
    ldo_sim_destroy(sim);
    return EXIT_SUCCESS;
}
