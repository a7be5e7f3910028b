// list_directed.h -- Fortran list-directed output (`write (unit,*) ...`) as gfortran formats it, for the text tables of
// write_output (src/ARTES.f90:3525-3709: phase.dat, photometry.dat, spectrum.dat, normalization.dat, luminosity.dat,
// cell_depth.dat, optical_depth.dat).  Downstream scripts split on white space, but the files are kept column-compatible:
//   * every item is preceded by one blank;
//   * REAL(8): a 25-character field, 17 significant digits (G25.17E3): for 0.1 <= |x| < 1e17 an F field of width 20 with
//     17 - (digits before the point) decimals followed by 5 blanks ("   1.0000000000000000     "), otherwise
//     d.ddddddddddddddddE+ddd right-justified in the 25 characters ("   1.0000000000000001E-005");
//   * default INTEGER: an 11-character field ("           5" with the separator);
//   * character items: the text itself.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace artes_host {

inline std::string ld_real(double v) {
    char buf[96];
    std::string field;
    if (std::isnan(v)) field = "NaN";
    else if (std::isinf(v)) field = v > 0 ? "Infinity" : "-Infinity";
    else {
        const double av = std::fabs(v);
        // F editing while the value, rounded to 17 significant digits, lies in [0.1, 1e17)
        int k = (av == 0.0) ? 1 : (int)std::floor(std::log10(av)) + 1;      // digits before the decimal point
        bool fixed = (av == 0.0) || (av >= 0.1 && av < 1.e17);
        if (fixed && av != 0.0) {
            // rounding to 17 digits may carry into the next decade (9.99..9 -> 10.0..0): then one decimal less
            std::snprintf(buf, sizeof(buf), "%.16E", av);
            const int ex = std::atoi(std::strchr(buf, 'E') + 1);
            k = ex + 1;
            if (k > 17) fixed = false;
            if (k < 0) fixed = false;
        }
        if (fixed) {
            const int kk = k < 0 ? 0 : k;
            std::snprintf(buf, sizeof(buf), "%#.*f", 17 - kk, v);      // '#': Fortran keeps the point when no decimals are left ("10000000000000000.")
            std::string f(buf);
            if (f.size() < 20) f.insert(0, 20 - f.size(), ' ');
            return " " + f + "     ";
        }
        std::snprintf(buf, sizeof(buf), "%.16E", v);
        char* e = std::strchr(buf, 'E');
        const int ex = std::atoi(e + 1);
        *e = 0;
        char tail[16];
        std::snprintf(tail, sizeof(tail), "E%c%03d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        field = std::string(buf) + tail;
    }
    if (field.size() < 25) field.insert(0, 25 - field.size(), ' ');
    return " " + field;
}

inline std::string ld_int(long v) {
    char buf[32];
    std::snprintf(buf, sizeof(buf), " %11ld", v);
    return buf;
}

}  // namespace artes_host
