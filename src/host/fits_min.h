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

// Same, but only the first `n_hdus` HDUs (atmosphere.fits: HDUs 0-7 are small, HDU 8 is the dense matrix array).
bool fits_read_upto(const std::string& path, int n_hdus, std::vector<FitsImage>& hdus, std::string& err);

// Random access to the elements of one image HDU (element offsets in FITS / Fortran order), for the matrix HDU:
// a wavelength's slice is 2880 runs of `cells` contiguous values.
struct FitsSlab {
    ~FitsSlab();
    bool open(const std::string& path, int hdu_index, std::string& err);
    bool read(size_t elem_offset, size_t n, double* out, std::string& err);
    size_t elems = 0;
    std::vector<long> naxes;
  private:
    void* file_ = nullptr;
    long long data_start_ = 0;
    int bitpix_ = 0;
    std::vector<unsigned char> raw_;
};

// Streams one HDU's pixels in chunks through a callback (for the 16.6 GB/wavelength matrix HDU).
bool fits_read_hdu_chunked(const std::string& path, int hdu_index, size_t chunk_elems,
                           bool (*cb)(void* user, size_t offset, const double* vals, size_t n), void* user, std::string& err);

// One primary HDU, BITPIX -64 (what write_fits_3D/4D produce).
bool fits_write_image(const std::string& path, const std::vector<long>& naxes, const double* data, std::string& err);

}  // namespace artes_host
