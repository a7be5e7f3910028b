// fits_tool -- command-line access to the driver's output formatting, used by tests/test_output_stage.py to compare the
// bytes bin/ARTES writes with what the reference's own calls produce (write_fits_3D/4D through the vendored CFITSIO,
// src/ARTES.f90:3774-3841; list-directed text tables, :3525-3709), without needing a GPU.
//   fits_tool write out.fits n1 [n2 ...] < raw little-endian float64 values (n1 fastest)
//   fits_tool ld item ...      one list-directed record; an item is a real number, or i:<integer>, or s:<text>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fits_min.h"
#include "list_directed.h"

int main(int argc, char** argv) {
    if (argc >= 4 && std::strcmp(argv[1], "write") == 0) {
        std::vector<long> naxes;
        size_t n = 1;
        for (int i = 3; i < argc; ++i) { naxes.push_back(std::atol(argv[i])); n *= (size_t)naxes.back(); }
        std::vector<double> data(n);
        if (std::fread(data.data(), 8, n, stdin) != n) { std::fprintf(stderr, "fits_tool: short input\n"); return 2; }
        std::string err;
        if (!artes_host::fits_write_image(argv[2], naxes, data.data(), err)) { std::fprintf(stderr, "fits_tool: %s\n", err.c_str()); return 1; }
        return 0;
    }
    if (argc >= 2 && std::strcmp(argv[1], "ld") == 0) {
        std::string line;
        for (int i = 2; i < argc; ++i) {
            if (std::strncmp(argv[i], "i:", 2) == 0) line += artes_host::ld_int(std::atol(argv[i] + 2));
            else if (std::strncmp(argv[i], "s:", 2) == 0) line += std::string(" ") + (argv[i] + 2);
            else line += artes_host::ld_real(std::strtod(argv[i], nullptr));
        }
        std::printf("%s\n", line.c_str());
        return 0;
    }
    std::fprintf(stderr, "usage: fits_tool write out.fits n1 [n2 ...] < float64 | fits_tool ld item ...\n");
    return 64;
}
