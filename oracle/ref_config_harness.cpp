// oracle/ref_config_harness.cpp — TEST INFRASTRUCTURE.  Runs the reference's own sim::config_reader
// (/root/reference/src/sim/config_reader.cpp, compiled where it lies by oracle/Makefile) on a config file and prints
// what it leaves in `parameters`, `parameters_hvec` and its getters as one JSON object — the same keys as
// swk_host::SimConfig::to_json() (host/sim_config.cpp), so tests can compare the two field by field.
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <boost/log/trivial.hpp>

#include "sim/config_reader.h"
#include "sim/simulation_parameters.cuh"

template <class T>
static std::string join(const std::vector<T> &v)
{
    std::ostringstream o;
    o.precision(17);
    o << "[";
    for (size_t i = 0; i < v.size(); i++) o << (i ? ", " : "") << v[i];
    o << "]";
    return o.str();
}
static std::string jstr(const std::string &s)
{
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        o += c;
    }
    return o + "\"";
}
static std::string join_s(const std::vector<std::string> &v)
{
    std::string o = "[";
    for (size_t i = 0; i < v.size(); i++) o += (i ? ", " : "") + jstr(v[i]);
    return o + "]";
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: ref_config <config.ini>\n");
        return 2;
    }
    sim::config_reader cr;
    parameters p;
    parameters_hvec h;
    bool ok = false;
    try {
        ok = cr.prepare(argv[1], &p, &h);
    } catch (const std::exception &e) {
        std::cout << "{\"ok\": false, \"exception\": " << jstr(e.what()) << "}" << std::endl;
        return 0;
    }
    if (!ok) {
        std::cout << "{\"ok\": false}" << std::endl;
        return 0;
    }
    std::vector<std::string> outs;
    for (size_t i = 0; i < cr.get_filename("PHANTOM").size(); i++) outs.push_back(cr.get_output_filename(i));
    std::ostringstream o;
    o.precision(9);
    o << "{\"ok\": true, \"B0\": " << p.B0 << ", \"linear_phase_cycling\": " << p.linear_phase_cycling << ", \"quadratic_phase_cycling\": " << p.quadratic_phase_cycling
      << ", \"timestep_us\": " << p.timestep_us << ", \"TR_us\": " << p.TR_us << ", \"n_dummy_scan\": " << p.n_dummy_scan << ", \"n_spins\": " << p.n_spins
      << ", \"n_substrate\": " << p.n_substrate << ", \"n_scales\": " << p.n_scales << ", \"seed\": " << p.seed << ", \"max_iterations\": " << p.max_iterations
      << ", \"cross_fov\": " << (p.enCrossFOV ? 1 : 0) << ", \"record_trajectory\": " << (p.enRecordTrajectory ? 1 : 0) << ", \"scale_type\": " << (int)cr.get_scale_type()
      << ", \"diffusivity\": " << join(h.diffusivity) << ", \"RF_FA_deg\": " << join(h.RF_FA_deg) << ", \"RF_PH_deg\": " << join(h.RF_PH_deg)
      << ", \"dephasing_deg\": " << join(h.dephasing_deg) << ", \"gradientX_mTm\": " << join(h.gradientX_mTm) << ", \"gradientY_mTm\": " << join(h.gradientY_mTm)
      << ", \"gradientZ_mTm\": " << join(h.gradientZ_mTm) << ", \"pXY\": " << join(h.pXY) << ", \"T1_ms\": " << join(h.T1_ms) << ", \"T2_ms\": " << join(h.T2_ms)
      << ", \"TE\": " << join(h.TE_us) << ", \"RF_T\": " << join(h.RF_us) << ", \"dephasing_T\": " << join(h.dephasing_us) << ", \"gradient_T\": " << join(h.gradient_us)
      << ", \"scales\": " << join(cr.get_scales()) << ", \"phantom\": " << join_s(cr.get_filename("PHANTOM")) << ", \"xyz0\": " << join_s(cr.get_filename("XYZ0"))
      << ", \"m0\": " << join_s(cr.get_filename("M0")) << ", \"output_files\": " << join_s(outs) << "}";
    std::cout << o.str() << std::endl;
    return 0;
}
