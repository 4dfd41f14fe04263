// host/main.cpp — `spinwalk sim -c a.ini [b.ini ...] [-p] [-d N]` on the B200 engine.
//
// Command line of the reference's `sim` subcommand (src/spinwalk.cpp:53-56): -c/--configs (mandatory, one or more existing
// files), -p/--use_cpu, -d/--device.  This build has no CPU path: -p is accepted and refused with a clear message.  The other
// subcommands (phantom, config, dwi) are offline generators outside this engine's scope and are not provided here.
// Extensions: -d takes a comma-separated list (spins sharded over several GPUs), --compat selects the reference-arithmetic
// kernel, --sums adds the ensemble sums to the output file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "../include/spinwalk_engine.h"
#include "sim_driver.h"

static void usage()
{
    fprintf(stderr,
            "spinwalk (B200 engine)\nUsage: spinwalk [-g] sim -c CONFIG [CONFIG...] [-p] [-d DEVICE[,DEVICE...]] [--compat] [--sums] [-q]\n"
            "  -c,--configs   config. files as many as you want. e.g. -c config1.ini config2.ini ... configN.ini\n"
            "  -p,--use_cpu   not available: this engine has no CPU path\n"
            "  -d,--device    select GPU device(s) (if there are multiple GPUs)\n"
            "  -g,--gpu_info  print the number of GPUs\n");
}

int main(int argc, char **argv)
{
    std::vector<std::string> configs;
    swk_host::SimOptions opt;
    bool sim = false, use_cpu = false, gpu_info = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "sim") sim = true;
        else if (a == "phantom" || a == "config" || a == "dwi") {
            fprintf(stderr, "The '%s' subcommand is an offline generator of the reference package; this build provides 'sim' only.\n", a.c_str());
            return 1;
        } else if (a == "-g" || a == "--gpu_info") gpu_info = true;
        else if (a == "-p" || a == "--use_cpu") use_cpu = true;
        else if (a == "--compat") opt.compat = true;
        else if (a == "--sums") opt.write_sums = true;
        else if (a == "-q") opt.quiet = true;
        else if (a == "-l" || a == "--log") { if (i + 1 < argc) i++; } // log file of the reference CLI: messages go to stderr here
        else if ((a == "-d" || a == "--device") && i + 1 < argc) {
            opt.devices.clear();
            std::string v = argv[++i];
            for (size_t p = 0; p <= v.size();) {
                const size_t q = std::min(v.find(',', p), v.size());
                opt.devices.push_back(atoi(v.substr(p, q - p).c_str()));
                p = q + 1;
            }
        } else if (a == "-c" || a == "--configs") {
            while (i + 1 < argc && argv[i + 1][0] != '-') configs.push_back(argv[++i]);
        } else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "unknown argument: %s\n", a.c_str()); usage(); return 1; }
    }
    if (gpu_info) {
        printf("Number of GPU(s): %d\n", swk_device_count());
        if (!sim) return 0;
    }
    if (!sim) { usage(); return argc > 1 ? 1 : 0; }
    if (configs.empty()) { fprintf(stderr, "--configs is required\n"); return 1; }
    for (const auto &c : configs)
        if (!std::filesystem::exists(c)) { fprintf(stderr, "--configs: File does not exist: %s\n", c.c_str()); return 1; }
    if (use_cpu) { fprintf(stderr, "-p/--use_cpu: this engine has no CPU path (by design); run without -p\n"); return 1; }
    std::string err;
    if (!swk_host::run_sim(configs, opt, err)) {
        fprintf(stderr, "Simulation failed. See the log file\n%s\n", err.c_str()); // spinwalk.cpp:131-134
        return 1;
    }
    printf("Simulation completed successfully. See the log file\n");
    return 0;
}
