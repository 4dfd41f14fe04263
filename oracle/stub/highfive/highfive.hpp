/* oracle/stub/highfive/highfive.hpp — TEST INFRASTRUCTURE.
 * libhdf5 is not installed in this image, so the reference's vendored HighFive (which needs <hdf5.h>) cannot compile.
 * The phantom generators only touch HighFive inside phantom_base::save(); the harness calls run(false) and never saves,
 * so this stub just has to let src/phantom/phantom_base.cpp compile: the few names it uses, doing nothing. */
#pragma once
#include <cstddef>
#include <string>
#include <vector>
namespace HighFive {
struct DataSpace {
    DataSpace() {}
    template <class T> explicit DataSpace(const std::vector<T> &) {}
};
struct DataSet {
    template <class T> void write_raw(const T *) {}
};
struct File {
    enum Mode { Truncate = 1 };
    File(const std::string &, int) {}
    template <class T> DataSet createDataSet(const std::string &, const DataSpace &) { return DataSet(); }
};
} // namespace HighFive
