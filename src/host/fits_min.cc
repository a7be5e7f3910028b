// fits_min.cc -- see fits_min.h
#include "fits_min.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace artes_host {
namespace {

const size_t BLOCK = 2880;

struct Header {
    int bitpix = 0;
    std::vector<long> naxes;
    std::string extname;
    bool ok = false;
};

bool read_header(FILE* f, Header& h) {
    char block[BLOCK];
    int naxis = -1;
    bool done = false;
    while (!done) {
        if (fread(block, 1, BLOCK, f) != BLOCK) return false;
        for (size_t i = 0; i < BLOCK; i += 80) {
            std::string card(block + i, 80);
            std::string key = card.substr(0, 8);
            while (!key.empty() && key.back() == ' ') key.pop_back();
            if (key == "END") { done = true; break; }
            if (card.size() < 10 || card[8] != '=') continue;
            std::string val = card.substr(10);
            size_t slash = val.find('/');
            if (val.find('\'') == std::string::npos && slash != std::string::npos) val = val.substr(0, slash);
            if (key == "BITPIX") h.bitpix = std::atoi(val.c_str());
            else if (key == "NAXIS") { naxis = std::atoi(val.c_str()); h.naxes.assign(naxis, 0); }
            else if (key.rfind("NAXIS", 0) == 0 && key.size() > 5) {
                int k = std::atoi(key.c_str() + 5);
                if (k >= 1 && k <= (int)h.naxes.size()) h.naxes[k - 1] = std::atol(val.c_str());
            } else if (key == "EXTNAME") {
                size_t a = val.find('\''), b = val.rfind('\'');
                if (a != std::string::npos && b > a) { h.extname = val.substr(a + 1, b - a - 1); while (!h.extname.empty() && h.extname.back() == ' ') h.extname.pop_back(); }
            }
        }
    }
    h.ok = (naxis >= 0 && h.bitpix != 0);
    return h.ok;
}

inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint16_t bswap16(uint16_t v) { return __builtin_bswap16(v); }

void convert(const unsigned char* raw, size_t n, int bitpix, double* out) {
    for (size_t i = 0; i < n; ++i) {
        switch (bitpix) {
            case -64: { uint64_t v; std::memcpy(&v, raw + 8 * i, 8); v = bswap64(v); double d; std::memcpy(&d, &v, 8); out[i] = d; break; }
            case -32: { uint32_t v; std::memcpy(&v, raw + 4 * i, 4); v = bswap32(v); float d; std::memcpy(&d, &v, 4); out[i] = d; break; }
            case 64: { uint64_t v; std::memcpy(&v, raw + 8 * i, 8); out[i] = (double)(int64_t)bswap64(v); break; }
            case 32: { uint32_t v; std::memcpy(&v, raw + 4 * i, 4); out[i] = (double)(int32_t)bswap32(v); break; }
            case 16: { uint16_t v; std::memcpy(&v, raw + 2 * i, 2); out[i] = (double)(int16_t)bswap16(v); break; }
            default: out[i] = (double)raw[i]; break;
        }
    }
}

size_t elem_bytes(int bitpix) { return (size_t)(bitpix < 0 ? -bitpix : bitpix) / 8; }

// one 80-character card: value right-justified to column 30, then " / comment" (the layout ftphpr writes)
std::string card(const char* key, const std::string& value, const char* comment = nullptr) {
    char buf[160];
    if (comment) std::snprintf(buf, sizeof(buf), "%-8s= %20s / %s", key, value.c_str(), comment);
    else std::snprintf(buf, sizeof(buf), "%-8s= %20s", key, value.c_str());
    std::string s(buf);
    s.resize(80, ' ');
    return s;
}

}  // namespace

bool fits_read(const std::string& path, std::vector<FitsImage>& hdus, std::string& err) {
    return fits_read_upto(path, 1 << 30, hdus, err);
}

FitsSlab::~FitsSlab() { if (file_) std::fclose(static_cast<FILE*>(file_)); }

bool FitsSlab::open(const std::string& path, int hdu_index, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    for (int i = 0;; ++i) {
        Header h;
        if (!read_header(f, h)) { err = "HDU not found in " + path; std::fclose(f); return false; }
        size_t n = h.naxes.empty() ? 0 : 1;
        for (long a : h.naxes) n *= (size_t)a;
        const size_t bytes = n * elem_bytes(h.bitpix);
        if (i == hdu_index) {
            file_ = f; data_start_ = ftello(f); bitpix_ = h.bitpix; elems = n; naxes = h.naxes;
            return true;
        }
        if (fseeko(f, (off_t)(bytes + (BLOCK - bytes % BLOCK) % BLOCK), SEEK_CUR) != 0) { err = "seek failed"; std::fclose(f); return false; }
    }
}

bool FitsSlab::read(size_t elem_offset, size_t n, double* out, std::string& err) {
    FILE* f = static_cast<FILE*>(file_);
    const size_t eb = elem_bytes(bitpix_);
    if (!f || elem_offset + n > elems) { err = "FitsSlab::read out of range"; return false; }
    if (fseeko(f, (off_t)(data_start_ + (long long)(elem_offset * eb)), SEEK_SET) != 0) { err = "seek failed"; return false; }
    raw_.resize(n * eb);
    if (fread(raw_.data(), eb, n, f) != n) { err = "truncated FITS data"; return false; }
    convert(raw_.data(), n, bitpix_, out);
    return true;
}

bool fits_read_upto(const std::string& path, int n_hdus, std::vector<FitsImage>& hdus, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    hdus.clear();
    for (;;) {
        if ((int)hdus.size() >= n_hdus) break;
        Header h;
        if (!read_header(f, h)) break;
        FitsImage im;
        im.naxes = h.naxes;
        im.extname = h.extname;
        const size_t n = im.size(), eb = elem_bytes(h.bitpix), bytes = n * eb;
        im.data.resize(n);
        const size_t chunk = 1 << 20;
        std::vector<unsigned char> raw(std::min(n, chunk) * eb);
        for (size_t off = 0; off < n; off += chunk) {
            const size_t m = std::min(chunk, n - off);
            if (fread(raw.data(), eb, m, f) != m) { err = "truncated FITS data in " + path; std::fclose(f); return false; }
            convert(raw.data(), m, h.bitpix, im.data.data() + off);
        }
        const size_t pad = (BLOCK - bytes % BLOCK) % BLOCK;
        if (pad) std::fseek(f, (long)pad, SEEK_CUR);
        hdus.push_back(std::move(im));
    }
    std::fclose(f);
    if (hdus.empty()) { err = "no FITS HDU found in " + path; return false; }
    return true;
}

bool fits_read_hdu_chunked(const std::string& path, int hdu_index, size_t chunk_elems,
                           bool (*cb)(void*, size_t, const double*, size_t), void* user, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    for (int i = 0;; ++i) {
        Header h;
        if (!read_header(f, h)) { err = "HDU not found"; std::fclose(f); return false; }
        size_t n = h.naxes.empty() ? 0 : 1;
        for (long a : h.naxes) n *= (size_t)a;
        const size_t eb = elem_bytes(h.bitpix), bytes = n * eb;
        if (i == hdu_index) {
            std::vector<unsigned char> raw(chunk_elems * eb);
            std::vector<double> vals(chunk_elems);
            for (size_t off = 0; off < n; off += chunk_elems) {
                const size_t m = std::min(chunk_elems, n - off);
                if (fread(raw.data(), eb, m, f) != m) { err = "truncated FITS data"; std::fclose(f); return false; }
                convert(raw.data(), m, h.bitpix, vals.data());
                if (!cb(user, off, vals.data(), m)) { std::fclose(f); return true; }
            }
            std::fclose(f);
            return true;
        }
        const size_t total = bytes + (BLOCK - bytes % BLOCK) % BLOCK;
        if (fseeko(f, (off_t)total, SEEK_CUR) != 0) { err = "seek failed"; std::fclose(f); return false; }
    }
}

bool fits_write_image(const std::string& path, const std::vector<long>& naxes, const double* data, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot create " + path; return false; }
    // byte for byte the primary header the reference's calls produce (write_fits_3D/4D, src/ARTES.f90:3774-3841:
    // ftphpr(unit, simple=T, bitpix=-64, naxis, naxes, pcount=0, gcount=1, extend=T) of the vendored CFITSIO 3.34),
    // checked against lib/libcfitsio.so.3 in tests/test_output_stage.py
    std::string hdr;
    hdr += card("SIMPLE", "T", "file does conform to FITS standard");
    hdr += card("BITPIX", "-64", "number of bits per data pixel");
    hdr += card("NAXIS", std::to_string(naxes.size()), "number of data axes");
    size_t n = naxes.empty() ? 0 : 1;
    for (size_t i = 0; i < naxes.size(); ++i) {
        hdr += card(("NAXIS" + std::to_string(i + 1)).c_str(), std::to_string(naxes[i]), ("length of data axis " + std::to_string(i + 1)).c_str());
        n *= (size_t)naxes[i];
    }
    hdr += card("EXTEND", "T", "FITS dataset may contain extensions");
    {
        std::string c1 = "COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy";
        std::string c2 = "COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H";
        c1.resize(80, ' '); c2.resize(80, ' ');
        hdr += c1 + c2;
    }
    std::string end = "END";
    end.resize(80, ' ');
    hdr += end;
    hdr.resize((hdr.size() + BLOCK - 1) / BLOCK * BLOCK, ' ');
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    std::vector<uint64_t> be(n);
    for (size_t i = 0; i < n; ++i) { uint64_t v; std::memcpy(&v, data + i, 8); be[i] = bswap64(v); }
    std::fwrite(be.data(), 8, n, f);
    const size_t pad = (BLOCK - (n * 8) % BLOCK) % BLOCK;
    std::vector<char> zeros(pad, 0);
    if (pad) std::fwrite(zeros.data(), 1, pad, f);
    std::fclose(f);
    return true;
}

}  // namespace artes_host
