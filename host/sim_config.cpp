#include "sim_config.h"

#include <algorithm>
#include <cstdio>
#include <filesystem>
#include <iterator>
#include <sstream>

#include "ini_file.h"

namespace fs = std::filesystem;

namespace swk_host {

namespace {

// config_reader.cpp:16-23: every space-separated vector is parsed AS FLOAT and then cast (so "100e3" is a valid time,
// and integers above 2^24 lose their low bits); parsing stops at the first token that is not a number.
template <class T>
std::vector<T> str2vec(const std::string &s)
{
    std::istringstream ss(s);
    std::vector<T> v;
    for (std::istream_iterator<float> it(ss), end; it != end; ++it) v.push_back(static_cast<T>(*it));
    return v;
}

std::string idx(const std::string &name, unsigned i) { return name + "[" + std::to_string(i) + "]"; }

template <class T>
std::string join(const std::vector<T> &v)
{
    std::ostringstream o;
    o.precision(17);
    o << "[";
    for (size_t i = 0; i < v.size(); i++) o << (i ? ", " : "") << v[i];
    o << "]";
    return o.str();
}
std::string jstr(const std::string &s)
{
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        o += c;
    }
    return o + "\"";
}
std::string join_s(const std::vector<std::string> &v)
{
    std::string o = "[";
    for (size_t i = 0; i < v.size(); i++) o += (i ? ", " : "") + jstr(v[i]);
    return o + "]";
}

} // namespace

bool SimConfig::prepare(const std::string &config_file, bool check_files)
{
    *this = SimConfig(); // ≙ cleanup() + fresh parameters
    config_filename = config_file;
    try {
        if (!read(config_file)) return false;
    } catch (const std::exception &ex) { // std::stof / stoi / stod on a malformed value (the reference would terminate)
        error = std::string("malformed value in config: ") + ex.what();
        return false;
    }
    if (!check(check_files)) return false;
    timing_scale();
    return true;
}

void SimConfig::timing_scale()
{ // config_reader.cpp:39-46: microseconds -> timepoints by INTEGER division
    for (auto *v : {&TE_us, &RF_us, &dephasing_us, &gradient_us})
        for (auto &x : *v) x = x / timestep_us;
    times_in_timepoints = true;
}

bool SimConfig::read(const std::string &path)
{
    if (!fs::exists(path)) {
        error = "Config-file does not exist: " + path;
        return false;
    }
    IniFile ini;
    if (!ini.load(path)) {
        error = "Failed to read config file: " + path;
        return false;
    }
    auto val = [&](const char *sec, const std::string &key) { return ini.get(sec, key); };

    // ---- GENERAL: the parent is read FIRST, this file then overrides what it sets (config_reader.cpp:78-87) ----
    if (!val("GENERAL", "PARENT_CONFIG").empty()) {
        fs::path parent(val("GENERAL", "PARENT_CONFIG"));
        if (parent.is_relative()) parent = fs::absolute(path).parent_path() / parent;
        if (!read(parent.string())) return false;
    }
    if (!val("GENERAL", "SEQ_NAME").empty()) seq_name = val("GENERAL", "SEQ_NAME");

    // ---- FILES (config_reader.cpp:89-104) ----
    const fs::path here = fs::absolute(path).parent_path();
    for (auto pr : {std::make_pair("PHANTOM", &phantom), std::make_pair("XYZ0", &xyz0), std::make_pair("M0", &m0)}) {
        const std::string name = pr.first;
        std::vector<std::string> &files = *pr.second;
        if (ini.has("FILES", idx(name, 0))) files.clear();
        for (unsigned i = 0; i < 65536 && !val("FILES", idx(name, i)).empty(); i++) files.push_back(val("FILES", idx(name, i)));
        for (auto &f : files)
            if (fs::path(f).is_relative()) f = fs::weakly_canonical(here / f).string();
    }
    if (!val("FILES", "OUTPUT_DIR").empty()) output_dir = val("FILES", "OUTPUT_DIR");
    if (fs::path(output_dir).is_relative()) // relative to the TOP-LEVEL config file, whichever file names it (config_reader.cpp:102-103)
        output_dir = fs::weakly_canonical(fs::absolute(config_filename).parent_path() / output_dir).string();

    // ---- SCAN_PARAMETERS (config_reader.cpp:106-141) ----
    const char *SP = "SCAN_PARAMETERS";
    if (!val(SP, "TR").empty()) TR_us = (int32_t)std::stof(val(SP, "TR"));
    if (!val(SP, "TIME_STEP").empty()) timestep_us = (int32_t)std::stof(val(SP, "TIME_STEP"));
    if (!val(SP, "TE").empty()) TE_us = str2vec<int32_t>(val(SP, "TE"));
    if (!val(SP, "RF_T").empty()) RF_us = str2vec<int32_t>(val(SP, "RF_T"));
    if (!val(SP, "RF_FA").empty()) RF_FA_deg = str2vec<float>(val(SP, "RF_FA"));
    if (!val(SP, "RF_PH").empty()) RF_PH_deg = str2vec<float>(val(SP, "RF_PH"));
    if (!val(SP, "DEPHASING_T").empty()) dephasing_us = str2vec<int32_t>(val(SP, "DEPHASING_T"));
    if (!val(SP, "DEPHASING").empty()) dephasing_deg = str2vec<float>(val(SP, "DEPHASING"));
    if (!val(SP, "GRADIENT_T").empty()) gradient_us = str2vec<int32_t>(val(SP, "GRADIENT_T"));
    if (!val(SP, "GRADIENT_X").empty()) gradientX_mTm = str2vec<float>(val(SP, "GRADIENT_X"));
    if (!val(SP, "GRADIENT_Y").empty()) gradientY_mTm = str2vec<float>(val(SP, "GRADIENT_Y"));
    if (!val(SP, "GRADIENT_Z").empty()) gradientZ_mTm = str2vec<float>(val(SP, "GRADIENT_Z"));
    if (!val(SP, "DUMMY_SCAN").empty()) n_dummy_scan = std::stoi(val(SP, "DUMMY_SCAN"));
    if (!val(SP, "LINEAR_PHASE_CYCLING").empty()) linear_phase_cycling = std::stof(val(SP, "LINEAR_PHASE_CYCLING"));
    if (!val(SP, "QUADRATIC_PHASE_CYCLING").empty()) quadratic_phase_cycling = std::stof(val(SP, "QUADRATIC_PHASE_CYCLING"));

    // ---- SIMULATION_PARAMETERS (config_reader.cpp:144-159) ----
    const char *SM = "SIMULATION_PARAMETERS";
    if (!val(SM, "B0").empty()) B0 = std::stof(val(SM, "B0"));
    if (!val(SM, "SEED").empty()) seed = (uint64_t)(int64_t)std::stoi(val(SM, "SEED"));
    if (!val(SM, "NUMBER_OF_SPINS").empty()) n_spins = (uint32_t)std::stod(val(SM, "NUMBER_OF_SPINS")); // scientific notation allowed
    if (!val(SM, "CROSS_FOV").empty()) cross_fov = std::stoi(val(SM, "CROSS_FOV")) != 0;
    if (!val(SM, "RECORD_TRAJECTORY").empty()) record_trajectory = std::stoi(val(SM, "RECORD_TRAJECTORY")) != 0;
    if (!val(SM, "MAX_ITERATIONS").empty()) max_iterations = (uint64_t)std::stod(val(SM, "MAX_ITERATIONS"));
    if (ini.has(SM, "SCALE[0]")) scales.clear();
    for (unsigned i = 0; i < 65536 && !val(SM, idx("SCALE", i)).empty(); i++) scales.push_back((float)std::stod(val(SM, idx("SCALE", i))));
    n_scales = (uint32_t)scales.size();
    if (!val(SM, "WHAT_TO_SCALE").empty()) scale_type = std::stoi(val(SM, "WHAT_TO_SCALE"));

    // ---- TISSUE_PARAMETERS (config_reader.cpp:161-190); a child that lists any index replaces the whole vector ----
    const char *TP = "TISSUE_PARAMETERS";
    std::vector<double> D;
    for (unsigned i = 0; i < 65536 && !val(TP, idx("DIFFUSIVITY", i)).empty(); i++) D.push_back(std::stof(val(TP, idx("DIFFUSIVITY", i)))); // via FLOAT
    if (!D.empty()) diffusivity = D;
    n_substrate = (uint32_t)diffusivity.size();
    std::vector<float> t1, t2, p;
    for (unsigned i = 0; i < 65536 && !val(TP, idx("T1", i)).empty(); i++) t1.push_back(std::stof(val(TP, idx("T1", i))));
    if (!t1.empty()) T1_ms = t1;
    for (unsigned i = 0; i < 65536 && !val(TP, idx("T2", i)).empty(); i++) t2.push_back(std::stof(val(TP, idx("T2", i))));
    if (!t2.empty()) T2_ms = t2;
    for (unsigned i = 0; i < 65536 && !val(TP, idx("P_XY", i)).empty(); i++) {
        std::istringstream iss(val(TP, idx("P_XY", i)));
        for (std::istream_iterator<double> it(iss), end; it != end; ++it) p.push_back((float)*it);
    }
    if (!p.empty()) pXY = p;
    return true;
}

bool SimConfig::check(bool check_files)
{ // config_reader.cpp:195-324, same order, same messages in spirit
    auto fail = [&](const std::string &m) { error = m; return false; };
    auto sz = [](size_t a, size_t b, size_t c = (size_t)-1) { return std::to_string(a) + " vs " + std::to_string(b) + (c == (size_t)-1 ? "" : " vs " + std::to_string(c)); };
    if (RF_FA_deg.size() != RF_us.size() || RF_FA_deg.size() != RF_PH_deg.size())
        return fail("RF_FA, RF_PH and RF_us must have the same number of elements " + sz(RF_FA_deg.size(), RF_PH_deg.size(), RF_us.size()));
    if (dephasing_us.size() != dephasing_deg.size()) return fail("DEPHASING and DEPHASING_T must have the same number of elements " + sz(dephasing_deg.size(), dephasing_us.size()));
    if (gradientX_mTm.size() != gradientY_mTm.size() || gradientX_mTm.size() != gradientZ_mTm.size())
        return fail("GRADIENTS must have the same number of elements " + sz(gradientX_mTm.size(), gradientY_mTm.size(), gradientZ_mTm.size()));
    if (gradientX_mTm.size() != gradient_us.size()) return fail("GRADIENT_XYZ and GRADIENT_T must have the same number of elements " + sz(gradientX_mTm.size(), gradient_us.size()));
    if (T1_ms.size() != T2_ms.size()) return fail("T1 and T2 must have the same number of elements " + sz(T1_ms.size(), T2_ms.size()));
    if (T1_ms.size() != diffusivity.size()) return fail("T1 and diffusivity must have the same number of elements " + sz(T1_ms.size(), diffusivity.size()));
    if (T1_ms.size() * T1_ms.size() != pXY.size()) return fail("T1 and P_XY must have the same number of elements " + sz(T1_ms.size(), pXY.size()));
    if (scales.empty()) scales.push_back(1.0f); // "SCALE is not set! Using default value 1.0" (n_scales keeps the value read() gave it)
    if (diffusivity.empty() || T1_ms.empty() || T2_ms.empty()) return fail("Diffusivity, T1 and T2 must have at least one element");

    if (check_files)
        for (const auto *files : {&phantom, &xyz0, &m0})
            for (const auto &f : *files)
                if (!fs::exists(f)) return fail("File does not exist: " + f);
    xyz0.resize(phantom.size(), "");
    m0.resize(phantom.size(), "");
    if (check_files) {
        std::error_code ec;
        fs::create_directories(fs::path(output_dir), ec);
        if (ec) return fail("Creating directory " + output_dir + " failed. " + ec.message());
    }
    output_files.clear();
    for (const auto &ph : phantom) {
        fs::path f = fs::path(output_dir) / (seq_name + "_" + fs::path(ph).filename().string());
        output_files.push_back(f.replace_extension(".h5").string());
    }

    auto strictly_ascending = [](const std::vector<int32_t> &v) { return std::is_sorted(v.begin(), v.end()) && std::adjacent_find(v.begin(), v.end()) == v.end(); };
    if (TE_us.empty() || !strictly_ascending(TE_us) || TE_us[0] < 0)
        return fail("TE must exists and be in ascending order and must not have duplicates or negative values: " + join(TE_us));
    if (RF_us.empty() || !strictly_ascending(RF_us) || RF_us[0] != 0)
        return fail("RF times must be in ascending order, starts with 0 and must not have duplicates values: " + join(RF_us));
    if (!strictly_ascending(dephasing_us)) return fail("Dephasing Times must be in ascending order and must not have duplicates values: " + join(dephasing_us));
    if (!strictly_ascending(gradient_us)) return fail("Gradient times must be in a strickly ascending order and must not have duplicates values: " + join(gradient_us));
    if (TR_us < 0 || timestep_us < 0) return fail("TR and timestep must be set");
    if (timestep_us == 0) return fail("TIME_STEP must not be 0"); // the reference divides by it in timing_scale() and dies with SIGFPE
    if (scale_type != 0 && scale_type != 1 && scale_type != 2) return fail("WHAT_TO_SCALE must be 0, 1, or 2, but is " + std::to_string(scale_type));
    return true;
}

std::string SimConfig::to_json() const
{
    std::ostringstream o;
    o.precision(9);
    o << "{\"B0\": " << B0 << ", \"linear_phase_cycling\": " << linear_phase_cycling << ", \"quadratic_phase_cycling\": " << quadratic_phase_cycling
      << ", \"timestep_us\": " << timestep_us << ", \"TR_us\": " << TR_us << ", \"n_dummy_scan\": " << n_dummy_scan << ", \"n_spins\": " << n_spins
      << ", \"n_substrate\": " << n_substrate << ", \"n_scales\": " << n_scales << ", \"seed\": " << seed << ", \"max_iterations\": " << max_iterations
      << ", \"cross_fov\": " << (cross_fov ? 1 : 0) << ", \"record_trajectory\": " << (record_trajectory ? 1 : 0) << ", \"scale_type\": " << scale_type
      << ", \"diffusivity\": " << join(diffusivity) << ", \"RF_FA_deg\": " << join(RF_FA_deg) << ", \"RF_PH_deg\": " << join(RF_PH_deg)
      << ", \"dephasing_deg\": " << join(dephasing_deg) << ", \"gradientX_mTm\": " << join(gradientX_mTm) << ", \"gradientY_mTm\": " << join(gradientY_mTm)
      << ", \"gradientZ_mTm\": " << join(gradientZ_mTm) << ", \"pXY\": " << join(pXY) << ", \"T1_ms\": " << join(T1_ms) << ", \"T2_ms\": " << join(T2_ms)
      << ", \"TE\": " << join(TE_us) << ", \"RF_T\": " << join(RF_us) << ", \"dephasing_T\": " << join(dephasing_us) << ", \"gradient_T\": " << join(gradient_us)
      << ", \"scales\": " << join(scales) << ", \"seq_name\": " << jstr(seq_name) << ", \"output_dir\": " << jstr(output_dir)
      << ", \"phantom\": " << join_s(phantom) << ", \"xyz0\": " << join_s(xyz0) << ", \"m0\": " << join_s(m0) << ", \"output_files\": " << join_s(output_files) << "}";
    return o.str();
}

} // namespace swk_host
