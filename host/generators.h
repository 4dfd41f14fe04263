// host/generators.h — the reference's offline generator subcommands, host side:
//   `spinwalk dwi`     src/dwi/handler.cpp:8-17 + src/dwi/pgse.cpp:67-153      PGSE gradient table written into a config file
//   `spinwalk config`  src/config/handler.cpp:14-39 + config_generator.cpp     GRE / SE / bSSFP config + default_config.ini
//   `spinwalk phantom` src/phantom/handler.cpp:10-35                          cylinders / spheres / two pools -> HDF5 (GPU fill)
// dwi and config are pure text and need no device; phantom drives the GPU generator through include/spinwalk_phantom.h.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace swk_host {

// ≙ dMRI::execute_args (src/dwi/handler.h:9-16); CLI: -b b... -v x y z -d start δ Δ -c config (src/spinwalk.cpp:80-84)
struct DwiArgs {
    uint32_t start_ms = 0, delta_ms = 0, DELTA_ms = 0;
    std::vector<float> dir = {0.f, 0.f, 1.f};
    std::vector<double> b_value = {0.0};
    std::string config;
};
bool generate_dwi(const DwiArgs &args, std::string &error);

// ≙ config::execute_args (src/config/handler.h:9-15); CLI: -s seq -p phantoms... -e TE -t timestep -o output (src/spinwalk.cpp:73-78)
struct ConfigArgs {
    std::string seq_name = "default";
    uint32_t TE_us = 0, timestep_us = 0;
    std::vector<std::string> phantoms;
    std::string output = "config_default.ini";
};
bool generate_config(const ConfigArgs &args, std::string &error);

// ≙ phantom::execute_args (src/phantom/handler.h:10-25) with the CLI defaults of src/spinwalk.cpp:33-37
struct PhantomArgs {
    bool cylinder = false, sphere = false, twopools = false, ply = false;
    float radius = 50.f, orientation = 90.f, volume_fraction = 4.f, fov = 1000.f;
    uint32_t resolution = 500;
    float dchi = 0.11e-6f, oxy_level = 0.75f;
    int32_t seed = -1;
    std::string ply_file, output;
    int device = 0;
    bool quiet = false;
};
// defined in phantom_cli.cpp (links libspinwalk_b200.so); not part of libswkhost.so
bool generate_phantom(const PhantomArgs &args, std::string &error);

} // namespace swk_host
