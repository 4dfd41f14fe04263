// host/sim_driver.h — the host side of `spinwalk sim`: what sim::monte_carlo::run does around the kernel
// (src/sim/monte_carlo.cu:199-355), driving the B200 engine through the C-ABI (include/spinwalk_engine.h) instead of
// launching cu_sim itself.  Reads the config (sim_config.h), the phantom / XYZ0 / M0 HDF5 files (h5lite.h), generates the
// start positions exactly like the reference (std::mt19937 + uniform_real_distribution<float>, monte_carlo.cu:142-151),
// runs every FoV / gradient / phase-cycling scale, and writes the reference's output datasets (monte_carlo.cu:168-197).
#pragma once

#include <string>
#include <vector>

namespace swk_host {

struct SimOptions {
    std::vector<int> devices = {0}; // -d: one id like the reference, or a comma-separated list (spins sharded, phantom replicated)
    bool compat = false;            // --compat: the reference CUDA build's arithmetic (bit-exact walks) instead of the fast path
    bool write_sums = false;        // --sums: add the per-(scale, echo, substrate) ensemble sums as dataset "sums"
    bool sums_only = false;         // --sums-only: no per-spin arrays at all (host or device); the output holds sums / scales / TE
    bool device_positions = false;  // --device-positions: default XYZ0 drawn on the device (swk_set_spins(NULL)) instead of std::mt19937 on the host
    bool quiet = false;
};

// ≙ sim::handler::execute (src/sim/handler.cu:8-17): every config file in turn; false at the first failure.
bool run_sim(const std::vector<std::string> &config_files, const SimOptions &opt, std::string &error);

} // namespace swk_host
