// host/capi.cpp — C hooks into the host-side pieces (config reader, HDF5 codec) for the Python tests (ctypes).
#include <cstring>
#include <string>

#include "sim_config.h"

extern "C" {

// Parses `path` like `spinwalk sim -c path` would and writes the resulting configuration as JSON into buf.
// Returns 0 on success, 1 when the config is rejected (buf then holds {"ok": false, "error": "..."}), -1 if buf is too small.
int swkh_config_json(const char *path, int check_files, char *buf, size_t n)
{
    swk_host::SimConfig c;
    std::string out;
    int rc = 0;
    if (c.prepare(path, check_files != 0)) {
        out = c.to_json();
        out.insert(1, "\"ok\": true, ");
    } else {
        std::string e;
        for (char ch : c.error) {
            if (ch == '"' || ch == '\\') e += '\\';
            e += ch;
        }
        out = "{\"ok\": false, \"error\": \"" + e + "\"}";
        rc = 1;
    }
    if (out.size() + 1 > n) return -1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return rc;
}

} // extern "C"
