// tests/cpp/h5_fuzz.cpp -- mutilates a good HDF5 file (truncation at many lengths, random byte flips in its metadata) and feeds every variant to
// the engine's reader (raptor_b200/csrc/h5_io.cu, compiled into this program with -fsanitize=address,undefined by tests/test_h5_io.py).  Every
// variant lives in an exact-size heap block, so a single out-of-range read aborts the run.   usage: h5_fuzz file.h5 metadata_bytes flips
#include "h5_io.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>

int main(int argc, char** argv){
    if(argc < 4) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<unsigned char> good((std::istreambuf_iterator<char>(f)), {});
    const size_t meta = (size_t)std::atol(argv[2]);
    const int flips = std::atoi(argv[3]);
    b200l2f::H5Contents c; std::string err;
    if(!b200l2f::h5_read(good.data(), good.size(), c, err)){ std::printf("the unmodified file failed: %s\n", err.c_str()); return 1; }
    int failed = 0, total = 0;
    for(size_t cut = 0; cut < good.size(); cut += (cut < meta ? 5 : 1009)){
        unsigned char* b = (unsigned char*)std::malloc(cut ? cut : 1); std::memcpy(b, good.data(), cut);
        failed += !b200l2f::h5_read(b, cut, c, err); total++; std::free(b);
    }
    std::mt19937 rng(1);
    for(int t = 0; t < flips; t++){
        unsigned char* b = (unsigned char*)std::malloc(good.size()); std::memcpy(b, good.data(), good.size());
        const int k = 1 + (int)(rng() % 4);
        for(int i = 0; i < k; i++) b[8 + rng() % (meta - 8)] = (unsigned char)(rng() & 255);
        failed += !b200l2f::h5_read(b, good.size(), c, err); total++; std::free(b);
    }
    std::printf("ok %d rejected of %d variants\n", failed, total);
    return 0;
}
