#include "sim_driver.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <random>
#include <thread>

#include "../include/spinwalk_engine.h"
#include "h5lite.h"
#include "sim_config.h"

namespace swk_host {

namespace {

struct Pinned { // big host arrays: page-locked when the driver allows it, plain otherwise
    void *p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    bool alloc(size_t n)
    {
        release();
        bytes = n;
        if (n == 0) return true;
        if (swk_alloc_pinned(&p, n) == SWK_OK && p) { pinned = true; return true; }
        p = malloc(n);
        pinned = false;
        return p != nullptr;
    }
    void release()
    {
        if (p) { if (pinned) swk_free_pinned(p); else free(p); }
        p = nullptr;
        bytes = 0;
    }
    ~Pinned() { release(); }
};

struct Phantom {
    std::vector<uint8_t> mask;
    std::vector<float> fieldmap, fov;
    uint64_t dims[3] = {0, 0, 0};
};

// ≙ monte_carlo::read_phantom (monte_carlo.cu:98-123): fieldmap optional, mask and fov mandatory
bool read_phantom(const std::string &file, Phantom &ph, std::string &err)
{
    h5::Reader r;
    if (!r.open(file)) { err = r.error(); return false; }
    ph.fieldmap.clear();
    if (r.exists("fieldmap") && !r.read("fieldmap", ph.fieldmap)) { err = r.error(); return false; }
    if (!r.read("mask", ph.mask)) { err = r.error(); return false; } // int8 masks (MATLAB / h5py) convert to uint8 like H5Dread
    h5::DatasetInfo di;
    if (!r.info("fov", di)) { err = r.error(); return false; }
    if (di.count() != 3) { err = "dataset \"fov\" has different size " + std::to_string(di.count()) + " vs 3"; return false; }
    ph.fov.resize(3);
    if (!r.read("fov", h5::DType::F32, ph.fov.data(), 3)) { err = r.error(); return false; }
    if (!r.info("mask", di)) { err = r.error(); return false; }
    if (di.dims.size() != 3) { err = "dataset \"mask\" must be 3-dimensional"; return false; }
    for (int i = 0; i < 3; i++) ph.dims[i] = di.dims[i];
    if (!ph.fieldmap.empty() && ph.fieldmap.size() != ph.mask.size()) { err = "fieldmap and mask sizes differ"; return false; }
    return true;
}

// ≙ monte_carlo::initialize_position (monte_carlo.cu:125-153)
bool init_positions(const std::string &file, uint64_t seed, const std::vector<float> &fov, std::vector<float> &xyz0, std::string &err)
{
    if (!file.empty()) {
        h5::Reader r;
        h5::DatasetInfo di;
        if (!r.open(file) || !r.info("XYZ", di)) { err = r.error(); return false; }
        if (di.count() != xyz0.size()) { err = "dataset \"XYZ\" has different size " + std::to_string(di.count()) + " vs " + std::to_string(xyz0.size()); return false; }
        if (!r.read("XYZ", h5::DType::F32, xyz0.data(), xyz0.size())) { err = r.error(); return false; }
        for (size_t i = 0; i < xyz0.size(); i++)
            if (xyz0[i] < 0 || xyz0[i] > fov[i % 3]) { err = "Initial positions are outside the FoV."; return false; }
        return true;
    }
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dx(0.01 * fov[0], 0.99 * fov[0]), dy(0.01 * fov[1], 0.99 * fov[1]), dz(0.01 * fov[2], 0.99 * fov[2]);
    for (size_t i = 0; i < xyz0.size() / 3; i++) {
        xyz0[3 * i + 0] = dx(gen);
        xyz0[3 * i + 1] = dy(gen);
        xyz0[3 * i + 2] = dz(gen);
    }
    return true;
}

struct EngineSet {
    std::vector<swk_engine *> e;
    ~EngineSet() { for (auto *x : e) swk_destroy(x); }
};

bool run_one(const std::string &config_file, const SimOptions &opt, std::string &err)
{
    auto t_run = std::chrono::steady_clock::now();
    SimConfig cfg;
    if (!cfg.prepare(config_file)) { err = cfg.error; return false; }

    // ---- parameters::prepare (simulation_parameters.cuh:227-245) through the engine's own helper ----
    swk_params P{};
    P.B0 = cfg.B0;
    P.linear_phase_cycling = cfg.linear_phase_cycling;
    P.quadratic_phase_cycling = cfg.quadratic_phase_cycling;
    P.timestep_us = cfg.timestep_us;
    P.TR_us = cfg.TR_us;
    P.n_dummy_scan = cfg.n_dummy_scan;
    P.n_spins = cfg.n_spins;
    P.n_substrate = cfg.n_substrate;
    P.seed = cfg.seed ? cfg.seed : std::random_device{}();
    P.max_iterations = cfg.max_iterations;
    P.cross_fov = cfg.cross_fov;
    P.record_trajectory = cfg.record_trajectory;
    std::vector<double> sigma(cfg.diffusivity.size());
    if (cfg.RF_FA_deg.empty() || swk_prepare(&P, cfg.RF_FA_deg[0], cfg.T1_ms[0], cfg.diffusivity.data(), (uint32_t)sigma.size(), sigma.data()) != SWK_OK) {
        err = "TR, TIME_STEP and at least one RF pulse must be set";
        return false;
    }
    swk_tables T{};
    T.step_sigma_m = sigma.data();          T.n_step_sigma = (uint32_t)sigma.size();
    T.T1_ms = cfg.T1_ms.data();             T.n_T1 = (uint32_t)cfg.T1_ms.size();
    T.T2_ms = cfg.T2_ms.data();             T.n_T2 = (uint32_t)cfg.T2_ms.size();
    T.pXY = cfg.pXY.data();                 T.n_pXY = (uint32_t)cfg.pXY.size();
    T.RF_FA_deg = cfg.RF_FA_deg.data();     T.n_RF_FA = (uint32_t)cfg.RF_FA_deg.size();
    T.RF_PH_deg = cfg.RF_PH_deg.data();     T.n_RF_PH = (uint32_t)cfg.RF_PH_deg.size();
    T.RF_tp = cfg.RF_us.data();             T.n_RF = (uint32_t)cfg.RF_us.size();
    T.TE_tp = cfg.TE_us.data();             T.n_TE = (uint32_t)cfg.TE_us.size();
    T.dephasing_deg = cfg.dephasing_deg.data(); T.n_dephasing_deg = (uint32_t)cfg.dephasing_deg.size();
    T.dephasing_tp = cfg.dephasing_us.data();   T.n_dephasing = (uint32_t)cfg.dephasing_us.size();
    T.gradX_mTm = cfg.gradientX_mTm.data(); T.n_gradX = (uint32_t)cfg.gradientX_mTm.size();
    T.gradY_mTm = cfg.gradientY_mTm.data(); T.n_gradY = (uint32_t)cfg.gradientY_mTm.size();
    T.gradZ_mTm = cfg.gradientZ_mTm.data(); T.n_gradZ = (uint32_t)cfg.gradientZ_mTm.size();
    T.gradient_tp = cfg.gradient_us.data(); T.n_gradient = (uint32_t)cfg.gradient_us.size();

    // ---- engines: one per device, spins sharded by contiguous id range (never more engines than spins: no empty shard) ----
    const size_t S = cfg.n_spins, K = cfg.scales.size(), E = cfg.TE_us.size(), ns = cfg.n_substrate;
    const size_t G = std::min<size_t>(opt.devices.size(), S);
    EngineSet es;
    for (size_t g = 0; g < G; g++) {
        swk_engine *e = nullptr;
        if (swk_create(opt.devices[g], &e) != SWK_OK) { err = swk_last_error(nullptr); return false; }
        es.e.push_back(e);
        if (swk_set_sequence(e, &P, &T) != SWK_OK) { err = swk_last_error(e); return false; }
    }

    // ---- host arrays, reference layouts (monte_carlo.cu:61-70); none with --sums-only ----
    const size_t trj = cfg.record_trajectory ? (size_t)P.n_timepoints * (size_t)(P.n_dummy_scan + 1) : 1;
    const bool per_spin = !opt.sums_only;
    Pinned M1, XYZ1, Tt;
    if (per_spin && (!M1.alloc(K * S * E * 3 * sizeof(float)) || !XYZ1.alloc(K * S * trj * 3 * sizeof(float)) || !Tt.alloc(K * S * E))) {
        err = "not enough host memory for the outputs (--sums-only writes the ensemble sums without per-spin arrays)";
        return false;
    }
    std::vector<float> xyz0, m0(0);
    std::vector<double> sums(K * E * ns * 4), part(K * E * ns * 4);

    for (size_t ip = 0; ip < cfg.phantom.size(); ip++) {
        if (!opt.quiet) fprintf(stderr, "Simulating phantom: %s\n", cfg.phantom[ip].c_str());
        Phantom ph;
        if (!read_phantom(cfg.phantom[ip], ph, err)) return false;
        const bool dev_pos = opt.device_positions && cfg.xyz0[ip].empty(); // an XYZ0 file always wins
        if (!dev_pos) {
            xyz0.resize(S * 3);
            if (!init_positions(cfg.xyz0[ip], P.seed, ph.fov, xyz0, err)) return false;
        }
        if (!cfg.m0[ip].empty()) { // the reference reads the file (size check) and then overwrites M0 with (0,0,1) anyway (monte_carlo.cu:155-166)
            h5::Reader r;
            h5::DatasetInfo di;
            if (!r.open(cfg.m0[ip]) || !r.info("M", di)) { err = r.error(); return false; }
            if (di.count() != S * 3) { err = "dataset \"M\" has different size " + std::to_string(di.count()) + " vs " + std::to_string(S * 3); return false; }
        }
        auto t_sim = std::chrono::steady_clock::now();
        std::fill(sums.begin(), sums.end(), 0.0);
        std::vector<std::string> errs(G);
        std::vector<std::vector<double>> parts(G, part);
        auto work = [&](size_t g) {
            swk_engine *e = es.e[g];
            const uint64_t first = S * g / G, last = S * (g + 1) / G;
            const float fov[3] = {ph.fov[0], ph.fov[1], ph.fov[2]};
            if (swk_set_phantom(e, ph.mask.data(), ph.fieldmap.empty() ? nullptr : ph.fieldmap.data(), ph.dims, fov, 0) != SWK_OK ||
                swk_set_host_rows(e, S, first) != SWK_OK ||
                swk_run(e, dev_pos ? nullptr : xyz0.data() + 3 * first, nullptr /* M0 = (0,0,1) */, (uint32_t)first, (uint32_t)(last - first), cfg.scales.data(),
                        (uint32_t)K, cfg.scale_type, opt.compat ? SWK_MODE_COMPAT : SWK_MODE_FAST, per_spin ? static_cast<float *>(M1.p) : nullptr,
                        per_spin ? static_cast<float *>(XYZ1.p) : nullptr, per_spin ? static_cast<uint8_t *>(Tt.p) : nullptr, parts[g].data(), nullptr) != SWK_OK)
                errs[g] = swk_last_error(e);
        };
        if (G == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (size_t g = 0; g < G; g++) th.emplace_back(work, g);
            for (auto &t : th) t.join();
        }
        for (size_t g = 0; g < G; g++) {
            if (!errs[g].empty()) { err = errs[g]; return false; }
            for (size_t i = 0; i < sums.size(); i++) sums[i] += parts[g][i]; // rank order: the same result for every run
        }
        if (!opt.quiet)
            fprintf(stderr, "Simulation took %.3f seconds.\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sim).count());

        // ---- save (monte_carlo.cu:168-197): M, XYZ, T, scales, TE ----
        const std::string &out = cfg.output_files[ip];
        std::error_code ec;
        std::filesystem::remove(out, ec);
        std::filesystem::create_directories(std::filesystem::absolute(out).parent_path(), ec);
        h5::Writer w(out);
        if (per_spin) {
            w.add("M", {K, S, E, 3}, h5::DType::F32, M1.p);
            w.add("XYZ", {K, S, trj, 3}, h5::DType::F32, XYZ1.p);
            w.add("T", {K, S, E, 1}, h5::DType::U8, Tt.p);
        }
        w.add("scales", {K, 1, 1, 1}, cfg.scales);
        std::vector<float> te_s;
        for (int32_t tp : cfg.TE_us) te_s.push_back(tp * cfg.timestep_us * 1e-6); // timepoints back to seconds (monte_carlo.cu:192-193)
        w.add("TE", {E, 1, 1, 1}, te_s);
        if (opt.write_sums) w.add("sums", {K, E, ns, 4}, sums);
        // which arithmetic produced the file (the reference has one; this engine has two): 0 = --compat (the reference's, spin by spin), 1 = fast
        const std::vector<uint8_t> mode_v{(uint8_t)(opt.compat ? SWK_MODE_COMPAT : SWK_MODE_FAST)}; // (the writer keeps pointers until close())
        const std::vector<uint64_t> seed_v{(uint64_t)P.seed};
        w.add("swk_mode", {1}, mode_v);
        w.add("swk_seed", {1}, seed_v);
        if (!w.close()) { err = w.error(); return false; }
        if (!opt.quiet) fprintf(stderr, "Saved %s\n", out.c_str());
    }
    if (!opt.quiet)
        fprintf(stderr, "Entire run took %.3f seconds.\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run).count());
    return true;
}

} // namespace

bool run_sim(const std::vector<std::string> &config_files, const SimOptions &opt, std::string &error)
{
    for (const auto &f : config_files)
        if (!run_one(f, opt, error)) return false;
    return true;
}

} // namespace swk_host
