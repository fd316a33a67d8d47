// raptor_b200/csrc/h5_io.h -- minimal HDF5 reader (host code only): what the engine needs to read rl-tools' `checkpoint.h5`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200l2f {

struct H5Dataset {
    std::string path;                 // "/actor/layers/0/weights/parameters"
    std::vector<int64_t> dims;
    std::vector<float> data;          // converted to fp32 (files hold IEEE float32 or float64, either byte order)
    int elem_size = 4;                // size of one element in the file
};
struct H5Attribute {
    std::string object;               // path of the group / dataset that carries it ("/actor/layers/0")
    std::string name;                 // "type", "activation_function", "meta", ...
    std::string value;                // string attributes only (rl-tools writes nothing else); others are skipped
};
struct H5Contents {
    std::vector<std::string> groups;  // every group path, depth first, in the file's (name-sorted) order; "/" first
    std::vector<H5Dataset> datasets;
    std::vector<H5Attribute> attributes;
};

// true if the HDF5 signature sits at offset 0 or at 512, 1024, 2048, ... (a file with a user block)
bool h5_has_signature(const unsigned char* bytes, size_t length);

// Reads the whole object tree of an HDF5 file held in memory.  Returns false and sets `err` on anything malformed or outside the
// subset described in h5_io.cu; never reads outside [bytes, bytes + length).
bool h5_read(const unsigned char* bytes, size_t length, H5Contents& out, std::string& err);

}  // namespace b200l2f
