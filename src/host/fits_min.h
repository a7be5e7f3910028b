// fits_min.h -- the subset of FITS the ARTES path touches, without CFITSIO.
//
// The reference links a vendored CFITSIO 3.34 binary (Makefile:21-25) and uses it in two places only:
//   * get_atmosphere reads the 9 float64 image HDUs of atmosphere.fits with ftopen/ftmrhd/ftgknj/ftgpvd
//     (src/ARTES.f90:2067-2201);
//   * write_fits_3D/4D write one primary HDU, BITPIX -64, with ftinit/ftphpr/ftpprd (:3774-3841).
// This reader/writer covers exactly that: 2880-byte blocks, 80-character cards, big-endian image data,
// BITPIX 8/16/32/64/-32/-64, no scaling keywords, no tables.
#pragma once
#include <string>
#include <vector>

namespace artes_host {

struct FitsImage {
    std::vector<long> naxes;     // NAXIS1.. (NAXIS1 varies fastest = Fortran index order)
    std::vector<double> data;    // converted to double
    std::string extname;
    size_t size() const { size_t n = naxes.empty() ? 0 : 1; for (long a : naxes) n *= (size_t)a; return n; }
};

// Reads every image HDU of `path`.  If `skip_data_from >= 0`, HDUs with that index or higher keep their
// header but their pixels are not converted (used to look at the dense matrix HDU without loading it).
bool fits_read(const std::string& path, std::vector<FitsImage>& hdus, std::string& err);

// Streams one HDU's pixels in chunks through a callback (for the 16.6 GB/wavelength matrix HDU).
bool fits_read_hdu_chunked(const std::string& path, int hdu_index, size_t chunk_elems,
                           bool (*cb)(void* user, size_t offset, const double* vals, size_t n), void* user, std::string& err);

// One primary HDU, BITPIX -64 (what write_fits_3D/4D produce).
bool fits_write_image(const std::string& path, const std::vector<long>& naxes, const double* data, std::string& err);

}  // namespace artes_host
