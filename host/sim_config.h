// host/sim_config.h — the `sim` configuration as the reference's config_reader leaves it in `parameters`,
// `parameters_hvec` and its own members (src/sim/config_reader.cpp:25-324, config_reader.h:20-44,
// simulation_parameters.cuh:39-43,176-201).  Same keys, same inheritance, same conversions, same checks.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace swk_host {

struct SimConfig {
    // ---- struct parameters (defaults: simulation_parameters.cuh:178-199) ----
    float    B0 = 9.4f;
    float    linear_phase_cycling = 0.f, quadratic_phase_cycling = 0.f;
    int32_t  timestep_us = -1, TR_us = -1, n_dummy_scan = 0;
    uint32_t n_spins = 1000, n_substrate = 0, n_scales = 1;
    uint64_t seed = 0, max_iterations = 9999;
    bool     cross_fov = true, record_trajectory = false;
    // ---- struct parameters_hvec; *_us hold MICROSECONDS until timing_scale() turns them into timepoints ----
    std::vector<double>  diffusivity;
    std::vector<float>   RF_FA_deg, RF_PH_deg, dephasing_deg, gradientX_mTm, gradientY_mTm, gradientZ_mTm, pXY, T1_ms, T2_ms;
    std::vector<int32_t> TE_us, RF_us, dephasing_us, gradient_us;
    bool times_in_timepoints = false;
    // ---- config_reader members ----
    std::string seq_name, output_dir, config_filename;
    std::vector<std::string> phantom, xyz0, m0, output_files;
    std::vector<float> scales;
    int scale_type = 0; // WHAT_TO_SCALE: 0 FoV, 1 gradient, 2 phase cycling (uninitialised in the reference when the key is absent)

    std::string error; // message of the last failure (the reference logs it and returns false)

    // ≙ config_reader::prepare (config_reader.cpp:25-37): cleanup, read (with PARENT_CONFIG recursion), check, timing_scale.
    // check_files = false skips the "file exists" test and the creation of OUTPUT_DIR (unit tests on bare configs).
    bool prepare(const std::string &config_file, bool check_files = true);

    std::string to_json() const; // every field above, for tests

private:
    bool read(const std::string &config_file);
    bool check(bool check_files);
    void timing_scale();
};

} // namespace swk_host
