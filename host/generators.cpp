#include "generators.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <filesystem>

#include "ini_edit.h"

namespace fs = std::filesystem;

namespace swk_host {

namespace {
constexpr double kGamma = 267515315.; // rad/s/T (src/definitions.h:20)

// value of [section] key in `config_file`, children overriding parents along PARENT_CONFIG (pgse.cpp:24-47)
bool inherited_value(const std::string &config_file, const std::string &section, const std::string &key, std::string &value, std::string &error, int depth = 0)
{
    if (depth > 64) { error = "PARENT_CONFIG chain is too deep (a cycle?)"; return false; }
    if (!fs::exists(config_file)) { error = "Config-file does not exist: " + config_file; return false; }
    IniDocument ini;
    if (!ini.load(config_file)) { error = "Failed to read config file: " + config_file; return false; }
    value.clear();
    const std::string parent = ini.get("GENERAL", "PARENT_CONFIG");
    if (!parent.empty()) {
        fs::path p(parent);
        if (p.is_relative()) p = fs::absolute(config_file).parent_path() / p;
        if (!inherited_value(p.string(), section, key, value, error, depth + 1)) return false;
    }
    const std::string own = ini.get(section, key);
    if (!own.empty()) value = own;
    return true;
}

std::string repeated(const std::string &head, const std::string &item, size_t n, const std::string &tail)
{
    std::string s = head;
    for (size_t i = 0; i < n; i++) s += item;
    return s + tail;
}
} // namespace

bool generate_dwi(const DwiArgs &a, std::string &error)
{
    std::string ts;
    if (!inherited_value(a.config, "SCAN_PARAMETERS", "TIME_STEP", ts, error)) return false;
    if (ts.empty()) { error = "TIME_STEP is not set! Create a section SCAN_PARAMETERS and set TIME_STEP"; return false; }
    int timestep_us = 0;
    try {
        timestep_us = std::stoi(ts);
    } catch (const std::exception &) {
        error = "TIME_STEP is not a number: " + ts;
        return false;
    }
    if (timestep_us <= 0) { error = "TIME_STEP must be positive"; return false; }
    if (a.DELTA_ms < a.delta_ms) { error = "\xCE\x94 must be greater than \xCE\xB4: " + std::to_string(a.DELTA_ms) + " vs " + std::to_string(a.delta_ms); return false; }
    if (a.b_value.empty() || a.dir.size() != 3) { error = "b-values and a 3-component direction are required"; return false; }

    // amplitude of the first b-value (Stejskal-Tanner, rectangular lobes): b = γ² G² δ² (Δ - δ/3), b in s/mm² (pgse.cpp:80-84)
    const double d = a.delta_ms * 1e-3, D = a.DELTA_ms * 1e-3;
    const double G2 = a.b_value[0] * 1e6 / (kGamma * kGamma * d * d * (D - d / 3.0));
    const double G = std::sqrt(G2) * 1000.; // mT/m

    std::vector<float> dir = a.dir; // normalised in place, stored back as float (pgse.cpp:87-94)
    const double norm = std::sqrt(double(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]));
    if (norm == 0) { error = "Direction vector is zero!"; return false; }
    for (float &c : dir) c = float(c / norm);

    IniDocument ini;
    ini.load(a.config);
    const uint32_t start_us = a.start_ms * 1000;
    ini.set("SCAN_PARAMETERS", "RF_FA", "90 180");
    ini.set("SCAN_PARAMETERS", "RF_PH", "0 90");
    ini.set("SCAN_PARAMETERS", "RF_T", "0 " + std::to_string(start_us + a.delta_ms * 1000 + (a.DELTA_ms - a.delta_ms) * 1000 / 2));

    // two lobes, each "0, n_points x G, 0" (pgse.cpp:104-123)
    const size_t n_points = a.delta_ms * 1000 / timestep_us;
    const char *axis_key[3] = {"GRADIENT_X", "GRADIENT_Y", "GRADIENT_Z"};
    for (int ax = 0; ax < 3; ax++) {
        const std::string lobe = repeated("0 ", std::to_string(G * dir[ax]) + " ", n_points, "0 ");
        ini.set("SCAN_PARAMETERS", axis_key[ax], lobe + lobe);
    }

    // sample times: one step before each lobe, the lobe, one step after (pgse.cpp:125-136); the reference's integer types are kept
    size_t i = 0;
    std::string t = std::to_string(start_us - timestep_us) + " ";
    for (i = 0; i < n_points; i++) t += std::to_string(start_us + i * timestep_us) + " ";
    t += std::to_string(start_us + i * timestep_us + timestep_us) + " ";
    const uint32_t gap_us = (a.DELTA_ms - a.delta_ms) * 1000;
    t += std::to_string(start_us + gap_us + i * timestep_us - timestep_us) + " ";
    for (; i < 2 * n_points; i++) t += std::to_string(start_us + gap_us + i * timestep_us) + " ";
    t += std::to_string(start_us + gap_us + i * timestep_us + timestep_us) + " ";
    ini.set("SCAN_PARAMETERS", "GRADIENT_T", t);

    // one gradient scale per b-value: G ∝ sqrt(b) (pgse.cpp:140-143)
    ini.set("SIMULATION_PARAMETERS", "WHAT_TO_SCALE", "1");
    for (size_t k = 0; k < a.b_value.size(); k++)
        ini.set("SIMULATION_PARAMETERS", "SCALE[" + std::to_string(k) + "]", std::to_string(std::sqrt(a.b_value[k] / a.b_value[0])));

    if (!ini.update_file(a.config, true)) { error = "Failed to write config file: " + a.config; return false; }
    return true;
}

namespace {
// config_default.ini as the generator emits it (config_generator.cpp:14-96)
IniDocument default_config(uint32_t TE_us, uint32_t timestep_us, const std::vector<std::string> &phantoms)
{
    IniDocument p;
    p.set("GENERAL", "PARENT_CONFIG", "");
    p.set("GENERAL", "SEQ_NAME", "noname");
    p.set("FILES", "OUTPUT_DIR", "./outputs");
    for (size_t i = 0; i < phantoms.size(); i++) p.set("FILES", "PHANTOM[" + std::to_string(i) + "]", phantoms[i]);
    for (const char *k : {"XYZ0[0]", "XYZ0[1]", "M0[0]", "M0[1]"}) p.set("FILES", k, "");
    const std::pair<const char *, const char *> tissue[] = {{"DIFFUSIVITY[0]", "1.0e-9"}, {"DIFFUSIVITY[1]", "1.0e-9"}, {"P_XY[0]", "1.0 0.0"}, {"P_XY[1]", "0.0 1.0"},
                                                            {"T1[0]", "2200"}, {"T1[1]", "2200"}, {"T2[0]", "41"}, {"T2[1]", "41"}};
    for (const auto &kv : tissue) p.set("TISSUE_PARAMETERS", kv.first, kv.second);
    p.set("SCAN_PARAMETERS", "TR", std::to_string(TE_us + timestep_us));
    p.set("SCAN_PARAMETERS", "TE", std::to_string(TE_us));
    const std::pair<const char *, const char *> scan[] = {{"RF_FA", "90.0"}, {"RF_PH", "0.0"}, {"RF_T", "0"}, {"DEPHASING", ""}, {"DEPHASING_T", ""},
                                                          {"GRADIENT_X", ""}, {"GRADIENT_Y", ""}, {"GRADIENT_Z", ""}, {"GRADIENT_T", ""}};
    for (const auto &kv : scan) p.set("SCAN_PARAMETERS", kv.first, kv.second);
    p.set("SCAN_PARAMETERS", "TIME_STEP", std::to_string(timestep_us));
    p.set("SCAN_PARAMETERS", "DUMMY_SCAN", "0");
    p.set("SCAN_PARAMETERS", "LINEAR_PHASE_CYCLING", "0");
    p.set("SCAN_PARAMETERS", "QUADRATIC_PHASE_CYCLING", "0");
    const std::pair<const char *, const char *> simp[] = {{"B0", "9.4"}, {"SEED", "0"}, {"NUMBER_OF_SPINS", "1e5"}, {"CROSS_FOV", "0"}, {"RECORD_TRAJECTORY", "0"},
                                                          {"MAX_ITERATIONS", "1e4"}, {"WHAT_TO_SCALE", "0"}, {"SCALE[0]", "1.0"}};
    for (const auto &kv : simp) p.set("SIMULATION_PARAMETERS", kv.first, kv.second);
    return p;
}
} // namespace

bool generate_config(const ConfigArgs &a, std::string &error)
{
    std::string seq = a.seq_name;
    std::transform(seq.begin(), seq.end(), seq.begin(), [](unsigned char c) { return std::tolower(c); });
    const std::string output = fs::weakly_canonical(fs::absolute(a.output)).string(); // handler.cpp:16

    // the three sequences differ in five entries (config_generator.cpp:99-169)
    struct Seq { const char *name; uint32_t TR; const char *FA, *PH; std::string RF_T; bool steady_state; };
    Seq s;
    if (seq == "gre") s = {"gre", a.TE_us + a.timestep_us, "90.0", "0", "0", false};
    else if (seq == "se") s = {"se", a.TE_us + a.timestep_us, "90.0 180.0", "0 90", "0 " + std::to_string(a.TE_us / 2), false};
    else if (seq == "bssfp") s = {"bssfp", a.TE_us * 2, "16.0", "0", "0", true};
    else { error = "Invalid sequence name!"; return false; }

    IniDocument ini;
    ini.set("GENERAL", "SEQ_NAME", s.name);
    for (size_t i = 0; i < a.phantoms.size(); i++) ini.set("FILES", "PHANTOM[" + std::to_string(i) + "]", a.phantoms[i]);
    ini.set("SCAN_PARAMETERS", "TR", std::to_string(s.TR));
    ini.set("SCAN_PARAMETERS", "TE", std::to_string(a.TE_us));
    ini.set("SCAN_PARAMETERS", "RF_FA", s.FA);
    ini.set("SCAN_PARAMETERS", "RF_PH", s.PH);
    ini.set("SCAN_PARAMETERS", "RF_T", s.RF_T);
    ini.set("SCAN_PARAMETERS", "TIME_STEP", std::to_string(a.timestep_us));
    if (s.steady_state) {
        ini.set("SCAN_PARAMETERS", "DUMMY_SCAN", "-1");
        ini.set("SCAN_PARAMETERS", "LINEAR_PHASE_CYCLING", "180");
        ini.set("SCAN_PARAMETERS", "QUADRATIC_PHASE_CYCLING", "0");
    }

    // write_ini (config_generator.cpp:172-195): parent next to the output, both created from scratch
    const fs::path out_path(output);
    std::error_code ec;
    fs::create_directories(out_path.parent_path(), ec);
    if (ec) { error = "Creating directory " + out_path.parent_path().string() + " failed. " + ec.message(); return false; }
    const fs::path parent_path = out_path.parent_path() / "default_config.ini";
    ini.set("GENERAL", "PARENT_CONFIG", parent_path.string());
    if (!ini.create_file(out_path.string(), true)) { error = "Failed to write config file: " + output; return false; }
    if (!default_config(a.TE_us, a.timestep_us, a.phantoms).create_file(parent_path.string(), true)) {
        error = "Failed to write config file: " + parent_path.string();
        return false;
    }
    return true;
}

} // namespace swk_host
