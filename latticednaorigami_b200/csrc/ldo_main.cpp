// ldo_main.cpp — command-line driver with the reference's interface (apps/main.cpp:19-117):
//   latticeDNAOrigami_b200 -i file.inp [--replicas N] [--device D] [--gpus G]
// Every replica runs the simulation described by the parameter file on one warp of a GPU. With --gpus G the N
// replicas are spread over devices D .. D+G-1, one host thread per GPU; replica-exchange ladders are sharded over
// the GPUs and exchange through one NCCL communicator (the reference: one MPI rank per replica).
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ldo_host.h"

int main(int argc, char* argv[]) {
    std::string inp;
    int replicas = 1, device = 0, gpus = 1;
    for (int i = 1; i < argc; i++) {
        std::string a {argv[i]};
        if ((a == "-i" || a == "--parameter_filename") && i + 1 < argc) inp = argv[++i];
        else if (a == "--replicas" && i + 1 < argc) replicas = std::atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else if (a == "-h" || a == "--help") {
            std::cout << "\nAllowed options:\n  -i, --parameter_filename  Input file\n  --replicas N              replicas run concurrently (one warp each)\n  --device D                first CUDA device\n  --gpus G                  devices to spread the replicas over\n\n";
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
    if (gpus < 1 || replicas < gpus || replicas % gpus != 0) {
        std::cout << "--replicas must be a positive multiple of --gpus" << std::endl;
        return 1;
    }
    std::vector<std::string> errors(gpus);
    std::vector<int> rcs(gpus, -1);
    unsigned char id[128] = {0};
    if (gpus > 1 && ldo_comm_unique_id(id) != 0) {
        std::cout << std::endl << "An exception occurred during the run" << std::endl << std::endl;
        std::cerr << ldo_host_last_error() << std::endl;
        std::cout << std::endl << "Ending run unsuccesfully" << std::endl << std::endl;
        return EXIT_FAILURE;
    }
    auto run_rank = [&](int rank) {
        ldo_sim* sim = ldo_sim_create(inp.c_str(), replicas / gpus, device + rank, rank, gpus);
        int rc = sim ? 0 : -1;
        if (rc == 0 && gpus > 1) rc = ldo_sim_comm_init(sim, id);
        if (rc == 0) rc = ldo_sim_run(sim);
        if (rc != 0) errors[rank] = ldo_host_last_error();
        if (sim) ldo_sim_destroy(sim);
        rcs[rank] = rc;
    };
    if (gpus == 1) {
        run_rank(0);
    }
    else {
        std::vector<std::thread> threads;
        for (int r = 0; r < gpus; r++) threads.emplace_back(run_rank, r);
        for (auto& t: threads) t.join();
    }
    for (int r = 0; r < gpus; r++) {
        if (rcs[r] != 0) {
            // same failure report as apps/main.cpp:105-115
            std::cout << std::endl << "An exception occurred during the run" << std::endl << std::endl;
            std::cerr << errors[r] << std::endl;
            std::cout << std::endl << "Ending run unsuccesfully" << std::endl << std::endl;
            return EXIT_FAILURE;
        }
    }
    return EXIT_SUCCESS;
}
