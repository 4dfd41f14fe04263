// host/main.cpp — the `spinwalk` command line of the reference (src/spinwalk.cpp:42-140) on the B200 engine.
//
//   spinwalk sim     -c a.ini [b.ini ...] [-p] [-d N[,M...]]      Monte-Carlo simulation (src/spinwalk.cpp:53-56)
//   spinwalk phantom -c|-s|-t -f FOV -z RES -o FILE [-r -n -v -d -y -e]   numerical phantom, voxel fill on the GPU (:58-72)
//   spinwalk config  -s SEQ -p PHANTOM... -e TE -t DT -o FILE     GRE / SE / bSSFP configuration (:73-78)
//   spinwalk dwi     -b B... -v X Y Z -d START δ Δ -c CONFIG      PGSE gradient table into a config (:80-84)
// Option names, defaults and mandatory flags follow the reference.  This build has no CPU path: `sim -p` is accepted with a warning and
// the simulation runs on the GPU.  Extensions: -d takes a comma-separated list (spins sharded over several GPUs, phantom
// replicated), --compat selects the reference-arithmetic kernel, --sums adds the ensemble sums to the output file, --sums-only writes
// nothing but the sums (no per-spin arrays on the host or the device: what makes 1e9-spin runs possible), --device-positions draws the
// default start positions on the GPU, -q is quiet.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "../include/spinwalk_engine.h"
#include "generators.h"
#include "sim_driver.h"

namespace {

void usage()
{
    fprintf(stderr,
            "spinwalk (B200 engine)\n"
            "Usage: spinwalk [-g] [-l LOG] SUBCOMMAND ...\n"
            "  sim      -c,--configs FILE...   config. files as many as you want\n"
            "           -p,--use_cpu           accepted with a warning: this engine has no CPU path, the run uses the GPU\n"
            "           -d,--device N[,M...]   select GPU device(s): spins are sharded over the listed devices\n"
            "           --compat               the reference CUDA build's arithmetic (minstd_rand + erfcinvf, FP64 positions): per-spin M / XYZ / T\n"
            "                                  equal the reference's bit for bit.  DEFAULT is the fast path (Philox + Box-Muller, fixed-point\n"
            "                                  positions, 20-bit field samples): another random stream, results agree with the reference within\n"
            "                                  Monte-Carlo error, not spin by spin; a spin whose step exceeds both FoV walls stays instead of being lost.\n"
            "                                  The output file records which one ran (datasets swk_mode: 0 compat / 1 fast, swk_seed).\n"
            "           --sums                 add dataset sums [scales, echoes, substrates, (sum Mx, sum My, sum Mz, count)]\n"
            "           --sums-only            write only sums / scales / TE: no per-spin arrays (needed beyond ~1e8 spins)\n"
            "           --device-positions     default start positions drawn on the GPU (Philox) instead of std::mt19937 on the host\n"
            "           --full-table           keep the [nx][ny][nz] voxel table even when the phantom does not depend on z;  -q quiet\n"
            "  phantom  -c,--cylinder | -s,--sphere | -t,--two_pools | -p,--ply -i,--ply_file MESH.ply\n"
            "           -r,--radius [50]  -n,--orientation [90]  -v,--volume_fraction [4]  -f,--fov (required)  -z,--resolution (required)\n"
            "           -d,--dchi [0.11e-6]  -y,--oxy_level [0.75]  -e,--seed [-1]  -o,--output (required)  [--device N]\n"
            "  config   -s,--seq_name GRE|SE|bSSFP  -p,--phantoms FILE...  -e,--TE us  -t,--timestep us  -o,--output FILE\n"
            "  dwi      -b,--bvalue B...  -v,--bvector X Y Z  -d,--delta START delta DELTA (ms)  -c,--config FILE\n"
            "  -g,--gpu_info  print GPU information\n");
}

// values following an option, up to the next option (negative numbers are values)
bool is_option(const char *s)
{
    if (s[0] != '-' || s[1] == '\0') return false;
    return !(isdigit((unsigned char)s[1]) || s[1] == '.');
}

struct Args {
    int argc;
    char **argv;
    int i;
    bool one(std::string &out)
    {
        if (i + 1 >= argc) return false;
        out = argv[++i];
        return true;
    }
    std::vector<std::string> many()
    {
        std::vector<std::string> v;
        while (i + 1 < argc && !is_option(argv[i + 1])) v.push_back(argv[++i]);
        return v;
    }
};

int fail_usage(const std::string &msg)
{
    fprintf(stderr, "%s\nRun with --help for more information.\n", msg.c_str());
    return 1;
}

bool to_float(const std::string &s, float &v)
{
    char *end = nullptr;
    v = strtof(s.c_str(), &end);
    return end && *end == '\0' && !s.empty();
}
bool to_double(const std::string &s, double &v)
{
    char *end = nullptr;
    v = strtod(s.c_str(), &end);
    return end && *end == '\0' && !s.empty();
}
bool to_u32(const std::string &s, uint32_t &v)
{
    char *end = nullptr;
    const long long x = strtoll(s.c_str(), &end, 10);
    if (!end || *end != '\0' || s.empty() || x < 0 || x > 0xffffffffll) return false;
    v = uint32_t(x);
    return true;
}

int run_phantom(Args &a)
{
    swk_host::PhantomArgs p;
    bool have_fov = false, have_res = false;
    for (a.i++; a.i < a.argc; a.i++) {
        const std::string o = a.argv[a.i];
        std::string v;
        float f = 0;
        uint32_t u = 0;
        if (o == "-c" || o == "--cylinder") p.cylinder = true;
        else if (o == "-s" || o == "--sphere") p.sphere = true;
        else if (o == "-t" || o == "--two_pools") p.twopools = true;
        else if (o == "-p" || o == "--ply") p.ply = true;
        else if (o == "-q") p.quiet = true;
        else if (o == "-r" || o == "--radius") { if (!a.one(v) || !to_float(v, p.radius)) return fail_usage("--radius: a number is required"); }
        else if (o == "-n" || o == "--orientation") { if (!a.one(v) || !to_float(v, p.orientation)) return fail_usage("--orientation: a number is required"); }
        else if (o == "-v" || o == "--volume_fraction") { if (!a.one(v) || !to_float(v, p.volume_fraction)) return fail_usage("--volume_fraction: a number is required"); }
        else if (o == "-f" || o == "--fov") {
            if (!a.one(v) || !to_float(v, f) || !(f > 0)) return fail_usage("--fov: Number less or equal to 0: " + v);
            p.fov = f;
            have_fov = true;
        } else if (o == "-z" || o == "--resolution") {
            if (!a.one(v) || !to_u32(v, u) || u == 0) return fail_usage("--resolution: Number less or equal to 0: " + v);
            p.resolution = u;
            have_res = true;
        } else if (o == "-d" || o == "--dchi") { if (!a.one(v) || !to_float(v, p.dchi)) return fail_usage("--dchi: a number is required"); }
        else if (o == "-y" || o == "--oxy_level") { if (!a.one(v) || !to_float(v, p.oxy_level)) return fail_usage("--oxy_level: a number is required"); }
        else if (o == "-e" || o == "--seed") { if (!a.one(v)) return fail_usage("--seed: a number is required"); p.seed = atoi(v.c_str()); }
        else if (o == "-i" || o == "--ply_file") {
            if (!a.one(p.ply_file)) return fail_usage("--ply_file: a path is required");
            if (!std::filesystem::exists(p.ply_file)) return fail_usage("--ply_file: File does not exist: " + p.ply_file);
        }
        else if (o == "-o" || o == "--output") { if (!a.one(p.output)) return fail_usage("--output: a path is required"); }
        else if (o == "--device") { if (!a.one(v)) return fail_usage("--device: a number is required"); p.device = atoi(v.c_str()); }
        else if (o == "-h" || o == "--help") { usage(); return 0; } // CLI11 gives every subcommand a help flag
        else return fail_usage("The following argument was not expected: " + o);
    }
    if (!have_fov) return fail_usage("--fov is required");
    if (!have_res) return fail_usage("--resolution is required");
    if (p.output.empty()) return fail_usage("--output is required");
    if (p.ply && p.ply_file.empty()) return fail_usage("--ply needs --ply_file");
    const int n_sel = int(p.cylinder) + int(p.sphere) + int(p.twopools) + int(p.ply);
    if (n_sel == 0 || n_sel == 4) { // src/spinwalk.cpp:91-96
        if (n_sel == 4) printf("Error! select either --cylinder or --sphere, not both!\n");
        usage();
        return 0;
    }
    std::string err;
    if (!swk_host::generate_phantom(p, err)) {
        fprintf(stderr, "Phantom generation failed.\n%s\n", err.c_str()); // src/spinwalk.cpp:124-127
        return 1;
    }
    return 0;
}

int run_config(Args &a)
{
    swk_host::ConfigArgs c;
    c.TE_us = 1000; // src/spinwalk.cpp:35
    c.timestep_us = 10;
    bool have[5] = {false, false, false, false, false};
    for (a.i++; a.i < a.argc; a.i++) {
        const std::string o = a.argv[a.i];
        std::string v;
        if (o == "-s" || o == "--seq_name") { if (!a.one(c.seq_name)) return fail_usage("--seq_name: a name is required"); have[0] = true; }
        else if (o == "-p" || o == "--phantoms") { c.phantoms = a.many(); have[1] = !c.phantoms.empty(); }
        else if (o == "-e" || o == "--TE") { if (!a.one(v) || !to_u32(v, c.TE_us) || c.TE_us == 0) return fail_usage("--TE: Number less or equal to 0: " + v); have[2] = true; }
        else if (o == "-t" || o == "--timestep") { if (!a.one(v) || !to_u32(v, c.timestep_us) || c.timestep_us == 0) return fail_usage("--timestep: Number less or equal to 0: " + v); have[3] = true; }
        else if (o == "-o" || o == "--output") { if (!a.one(c.output)) return fail_usage("--output: a path is required"); have[4] = true; }
        else if (o == "-h" || o == "--help") { usage(); return 0; } // CLI11 gives every subcommand a help flag
        else return fail_usage("The following argument was not expected: " + o);
    }
    const char *names[5] = {"--seq_name", "--phantoms", "--TE", "--timestep", "--output"};
    for (int k = 0; k < 5; k++)
        if (!have[k]) return fail_usage(std::string(names[k]) + " is required");
    std::string err;
    if (!swk_host::generate_config(c, err)) {
        printf("%s\n", err.c_str());
        fprintf(stderr, "Configuration file generation failed.\n"); // src/spinwalk.cpp:117-120
        return 1;
    }
    printf("Configuration file is generated in %s\n", std::filesystem::weakly_canonical(std::filesystem::absolute(c.output)).string().c_str());
    return 0;
}

int run_dwi(Args &a)
{
    swk_host::DwiArgs d;
    bool have_b = false, have_v = false, have_d = false;
    for (a.i++; a.i < a.argc; a.i++) {
        const std::string o = a.argv[a.i];
        if (o == "-b" || o == "--bvalue") {
            d.b_value.clear();
            for (const auto &s : a.many()) {
                double x;
                if (!to_double(s, x)) return fail_usage("--bvalue: not a number: " + s);
                d.b_value.push_back(x);
            }
            have_b = !d.b_value.empty();
        } else if (o == "-v" || o == "--bvector") {
            const auto v = a.many();
            if (v.size() != 3) return fail_usage("--bvector: 3 required");
            for (int k = 0; k < 3; k++)
                if (!to_float(v[k], d.dir[k])) return fail_usage("--bvector: not a number: " + v[k]);
            have_v = true;
        } else if (o == "-d" || o == "--delta") {
            const auto v = a.many();
            uint32_t t[3];
            if (v.size() != 3) return fail_usage("--delta: 3 required");
            for (int k = 0; k < 3; k++)
                if (!to_u32(v[k], t[k])) return fail_usage("--delta: not a non-negative integer: " + v[k]);
            d.start_ms = t[0];
            d.delta_ms = t[1];
            d.DELTA_ms = t[2];
            have_d = true;
        } else if (o == "-c" || o == "--config") {
            if (!a.one(d.config)) return fail_usage("--config: a path is required");
            if (!std::filesystem::exists(d.config)) return fail_usage("--config: File does not exist: " + d.config);
        } else if (o == "-h" || o == "--help") { usage(); return 0; } // CLI11 gives every subcommand a help flag
        else return fail_usage("The following argument was not expected: " + o);
    }
    if (!have_b) return fail_usage("--bvalue is required");
    if (!have_v) return fail_usage("--bvector is required");
    if (!have_d) return fail_usage("--delta is required");
    if (d.config.empty()) return fail_usage("--config is required");
    printf("Generating PGSE gradient table...\n");
    std::string err;
    if (!swk_host::generate_dwi(d, err)) {
        fprintf(stderr, "Diffusion gradient generation failed.\n%s\n", err.c_str()); // src/spinwalk.cpp:110-113
        return 1;
    }
    printf("Diffusion gradient table is generated successfully.\n");
    return 0;
}

int run_sim(Args &a)
{
    std::vector<std::string> configs;
    swk_host::SimOptions opt;
    bool use_cpu = false;
    for (a.i++; a.i < a.argc; a.i++) {
        const std::string o = a.argv[a.i];
        std::string v;
        if (o == "-p" || o == "--use_cpu") use_cpu = true;
        else if (o == "--compat") opt.compat = true;
        else if (o == "--sums") opt.write_sums = true;
        else if (o == "--sums-only") opt.write_sums = opt.sums_only = true;
        else if (o == "--device-positions") opt.device_positions = true;
        else if (o == "--full-table") setenv("SWK_NO_ZSLAB", "1", 1); // include/spinwalk_engine.h SWK_RUN_NO_ZSLAB
        else if (o == "--zslab") {} // round 1's opt-in: the z-slab table is the default now
        else if (o == "-q") opt.quiet = true;
        else if (o == "-d" || o == "--device") {
            if (!a.one(v)) return fail_usage("--device: a number is required");
            opt.devices.clear();
            for (size_t p = 0; p <= v.size();) {
                const size_t q = std::min(v.find(',', p), v.size());
                const std::string tok = v.substr(p, q - p);
                char *end = nullptr;
                const long id = strtol(tok.c_str(), &end, 10);
                if (tok.empty() || *end != '\0' || id < 0 || id > 1023) return fail_usage("--device: not a device id: '" + tok + "'");
                opt.devices.push_back((int)id);
                p = q + 1;
            }
        } else if (o == "-c" || o == "--configs") {
            for (const auto &s : a.many()) configs.push_back(s);
        } else if (o == "-h" || o == "--help") { usage(); return 0; } // CLI11 gives every subcommand a help flag
        else return fail_usage("The following argument was not expected: " + o);
    }
    if (configs.empty()) return fail_usage("--configs is required");
    for (const auto &c : configs)
        if (!std::filesystem::exists(c)) return fail_usage("--configs: File does not exist: " + c);
    if (use_cpu) fprintf(stderr, "warning: -p/--use_cpu ignored: this engine has no CPU path (by design); the simulation runs on the GPU\n");
    std::string err;
    if (!swk_host::run_sim(configs, opt, err)) {
        fprintf(stderr, "Simulation failed. See the log file\n%s\n", err.c_str()); // spinwalk.cpp:131-134
        return 1;
    }
    printf("Simulation completed successfully. See the log file\n");
    return 0;
}

} // namespace

int main(int argc, char **argv)
{
    Args a{argc, argv, 0};
    for (a.i = 1; a.i < argc; a.i++) {
        const std::string o = argv[a.i];
        if (o == "-g" || o == "--gpu_info") {
            char info[1024]; // ≙ callback_gpu_info: print_device_info(); exit(0) (src/spinwalk.cpp:50-51)
            swk_device_info(info, sizeof info);
            fputs(info, stdout);
            return 0;
        } else if (o == "-l" || o == "--log") { if (a.i + 1 < argc) a.i++; } // log file of the reference CLI: messages go to stdout/stderr here
        else if (o == "-h" || o == "--help") { usage(); return 0; }
        else if (o == "-v" || o == "--version") { printf("spinwalk (B200 engine) %d.%d\n", SWK_VERSION_MAJOR, SWK_VERSION_MINOR); return 0; }
        else if (o == "sim") return run_sim(a);
        else if (o == "phantom") return run_phantom(a);
        else if (o == "config") return run_config(a);
        else if (o == "dwi") return run_dwi(a);
        else return fail_usage("The following argument was not expected: " + o);
    }
    usage(); // no subcommand: the reference prints its help and returns 0 (src/spinwalk.cpp:87-90)
    return 0;
}
