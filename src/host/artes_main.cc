// artes_main.cc -- ./bin/ARTES [inputDirectory] [photons] -o [outputDirectory] -k [keyWord]=[value]
//
// Host driver of the B200 transport library: the same command line, artes.in keywords, input tree and
// output files as the reference's `program artes` (src/ARTES.f90), with `call radiative_transfer`
// (:146,185,241,255) replaced by artes_gpu_run (include/artes_gpu.h).  Everything here is thin host work
// the reference does in `initialize` (:269-516), `get_atmosphere` (:2054-2235), `grid_initialize`
// (:2237-2507), `photon_package` (:2509-2539), `run` (:121-267) and `write_output` (:3472-3772).
//
// The reference driver is Fortran; this image has no Fortran compiler, so the driver that is built and
// tested here is this C++ one.  fortran/artes_driver.f90 shows the same calls through iso_c_binding.
// There is no CPU fallback: without a CUDA device artes_gpu_create fails and the driver stops.
//
// Extra keywords (optional; reference inputs run unchanged):
//   gpu:devices=N      number of GPUs of this node to shard the photons over (default 1)
//   gpu:phase_walks=single|independent   phase curves (detector:type=phase): `single` (default) walks every photon packet ONCE and
//                      peels off towards all detector azimuths (artes_gpu_run_multi: the 68 angles below 170 deg in one call, the
//                      limb-biased angles in a second); `independent` repeats the walk for every azimuth like the reference
//   gpu:seed=S         Philox key (default 1; the reference seeds from the clock, :4175-4195)
//   gpu:mode=fast|faithful   arithmetic mode of the library (default fast)
// Environment: ARTES_DRYRUN=1 parses everything, prints the run configuration and stops before touching a GPU.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/artes_gpu.h"
#include "fits_min.h"
#include "list_directed.h"

namespace fs = std::filesystem;
using artes_host::FitsImage;

namespace {

// constants src/ARTES.f90:9-16
const double PI = 4.0 * std::atan(1.0);
const double K_B = 1.3806488e-23, HH = 6.62606957e-34, CC = 2.99792458e8;
const double R_SUN = 6.95500e8, PC = 3.08572e16, AU = 1.49598e11;

struct Config {   // defaults of initialize :283-314
    bool log_file = false;
    std::string email;
    int photon_source = 1;
    double packages = 100000;
    double fstop = 1.e-5, photon_minimum = 1.e-20;
    bool thermal_weight = true, photon_scattering = true;
    int photon_emission = 1;
    double photon_bias = 0.8;
    double t_star = 5800.0, r_star = R_SUN;
    bool stellar_direction = false;
    double theta_star = PI / 2.0, phi_star = 0.0;
    double surface_albedo = 0.0, oblateness = 0.0, orbit = 5.0 * AU;
    bool ring = false;
    bool phase_curve = false, spectrum = false, imaging_mono = false, imaging_broad = false;
    double det_theta = 90.0, det_phi = 90.0;   // (sic) degrees unless the keyword is given, see SURVEY App. A.20
    int nx = 25, ny = 25;
    double distance_planet = 10.0 * PC;
    bool flow_global = false, flow_theta = false;
    // gpu:* extensions
    int devices = 1;
    uint64_t seed = 1;
    int mode = ARTES_MODE_FAST;
    bool phase_single_walk = true;   // phase curves: ONE walk per packet observed from all detector azimuths (gpu:phase_walks=independent: one walk per azimuth like the reference)
};

[[noreturn]] void die(const std::string& msg) {
    std::fprintf(stdout, "%s\n", msg.c_str());
    std::exit(0);   // the reference leaves with exit code 0 on fatal input errors (:377,:4496)
}

double fortran_real(const std::string& v) {   // '(e50.0)': accepts 1d-20 as well as 1e-20
    std::string s = v;
    for (char& c : s) if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(s.c_str(), nullptr);
}

bool on_off(const std::string& v, bool current) { return v == "on" ? true : (v == "off" ? false : current); }

// input_parameters :4361-4500
void input_parameters(Config& c, const std::string& key, const std::string& value) {
    if (key == "general:log") c.log_file = on_off(value, c.log_file);
    else if (key == "general:email") c.email = value;
    else if (key == "photon:source") { if (value == "star") c.photon_source = 1; else if (value == "planet") c.photon_source = 2; }
    else if (key == "photon:fstop") c.fstop = fortran_real(value);
    else if (key == "photon:minimum") c.photon_minimum = fortran_real(value);
    else if (key == "photon:weight") c.thermal_weight = on_off(value, c.thermal_weight);
    else if (key == "photon:scattering") c.photon_scattering = on_off(value, c.photon_scattering);
    else if (key == "photon:emission") { if (value == "isotropic") c.photon_emission = 1; else if (value == "biased") c.photon_emission = 2; }
    else if (key == "photon:bias") c.photon_bias = fortran_real(value);
    else if (key == "star:temperature") c.t_star = fortran_real(value);
    else if (key == "star:radius") c.r_star = fortran_real(value) * R_SUN;
    else if (key == "star:direction") c.stellar_direction = on_off(value, c.stellar_direction);
    else if (key == "star:theta") {
        if (c.stellar_direction) {
            c.theta_star = fortran_real(value) * PI / 180.0;
            if (c.theta_star < 1.e-3f) c.theta_star = 1.e-3;
            if (c.theta_star > PI - 1.e-3f) c.theta_star = PI - 1.e-3f;
        }
    } else if (key == "star:phi") { if (c.stellar_direction) c.phi_star = fortran_real(value) * PI / 180.0; }
    else if (key == "planet:surface_albedo") c.surface_albedo = fortran_real(value);
    else if (key == "planet:oblateness") c.oblateness = fortran_real(value);
    else if (key == "planet:orbit") c.orbit = fortran_real(value) * AU;
    else if (key == "planet:ring") c.ring = on_off(value, c.ring);
    else if (key == "detector:type") {
        if (value == "phase") c.phase_curve = true;
        else if (value == "spectrum") c.spectrum = true;
        else if (value == "imaging_mono") c.imaging_mono = true;
        else if (value == "imaging_broad") c.imaging_broad = true;
    } else if (key == "detector:theta") {
        c.det_theta = fortran_real(value) * PI / 180.0;
        if (c.det_theta < 1.e-3f) c.det_theta = 1.e-3;
        if (c.det_theta > PI - 1.e-3f) c.det_theta = PI - 1.e-3f;
    } else if (key == "detector:phi") c.det_phi = fortran_real(value) * PI / 180.0;
    else if (key == "detector:pixel") { c.nx = std::atoi(value.c_str()); c.ny = c.nx; }
    else if (key == "detector:distance") c.distance_planet = fortran_real(value) * PC;
    else if (key == "output:flow_global") c.flow_global = on_off(value, c.flow_global);
    else if (key == "output:flow_latitudinal") c.flow_theta = on_off(value, c.flow_theta);
    else if (key == "gpu:devices") c.devices = std::max(1, std::atoi(value.c_str()));
    else if (key == "gpu:seed") c.seed = std::strtoull(value.c_str(), nullptr, 10);
    else if (key == "gpu:mode") c.mode = (value == "faithful") ? ARTES_MODE_FAITHFUL : ARTES_MODE_FAST;
    else if (key == "gpu:phase_walks") c.phase_single_walk = (value != "independent");
    else die(" Wrong keyword found in input file: " + key);
}

// get_key_value :4502-4517
void get_key_value(const std::string& line_in, std::string& key, std::string& value) {
    std::string line = line_in;
    while (!line.empty() && (line.back() == ' ' || line.back() == '\r' || line.back() == '\n' || line.back() == '\t')) line.pop_back();
    const size_t eq = line.find('=');
    key = (eq == std::string::npos) ? line : line.substr(0, eq);
    value = (eq == std::string::npos) ? "" : line.substr(eq + 1);
    while (!key.empty() && key.back() == ' ') key.pop_back();
    if (!value.empty() && (value[0] == '"' || value[0] == '\'')) value = value.substr(1, value.size() >= 2 ? value.size() - 2 : 0);
}

double planck_function(double temperature, double wavelength, int photon_source) {   // :1350-1367
    const double pref = (photon_source == 1) ? 2.0 * PI : 2.0;
    return (pref * HH * CC * CC / std::pow(wavelength, 5.0)) / (std::exp(HH * CC / (wavelength * K_B * temperature)) - 1.0);
}

struct Atmosphere {
    int nr = 0, nt = 0, np = 0, nl = 0, cells = 0;
    std::vector<double> rfront, thetafront, phifront, wavelengths;   // m, rad, rad, m
    std::vector<int32_t> thetaplane;
    std::vector<double> temperature, k_sca, k_abs;                   // [cells], [nl][cells]
    std::string fits_path;
};

// get_atmosphere :2054-2169 (HDUs 0-7; the matrix HDU 8 is streamed per wavelength by load_matrices)
void get_atmosphere(const std::string& path, Atmosphere& a) {
    std::vector<FitsImage> h;
    std::string err;
    if (!artes_host::fits_read_upto(path, 8, h, err)) die("atmosphere.fits: " + err);
    if (h.size() < 8) die("atmosphere.fits: expected 9 HDUs (python/atmosphere.py:449-459)");
    a.fits_path = path;
    a.nr = (int)h[0].size() - 1; a.nt = (int)h[1].size() - 1; a.np = (int)h[2].size(); a.nl = (int)h[3].size();
    a.cells = a.nr * a.nt * a.np;
    a.rfront = h[0].data;
    a.thetafront.resize(a.nt + 1); a.thetaplane.resize(a.nt + 1);
    for (int i = 0; i <= a.nt; ++i) {   // :2096-2106: the plane test is done in degrees, before the conversion
        const double t = h[1].data[i];
        a.thetaplane[i] = (t < 90.0 - 1.e-6 || t > 90.0 + 1.e-6) ? 1 : 2;
        a.thetafront[i] = t * PI / 180.0;
    }
    a.phifront.resize(a.np);
    for (int i = 0; i < a.np; ++i) a.phifront[i] = h[2].data[i] * PI / 180.0;
    a.wavelengths.resize(a.nl);
    for (int i = 0; i < a.nl; ++i) a.wavelengths[i] = h[3].data[i] * 1.e-6;
    if ((int)h[5].size() != a.cells || (int)h[6].size() != a.cells * a.nl || (int)h[7].size() != a.cells * a.nl)
        die("atmosphere.fits: array sizes do not match the grid");
    a.temperature = h[5].data; a.k_sca = h[6].data; a.k_abs = h[7].data;
}

// The scatter-matrix HDU (nr,ntheta,nphi,nlambda,16,180) of wavelength l, de-duplicated: python/atmosphere.py
// :351-372 mixes a handful of species per layer, so the dense array (cells x 23 040 B per wavelength) holds few
// distinct 180x16 blocks.  Two streaming passes over the wavelength's slice: hash every cell's block, then gather
// one representative per distinct hash.  Host memory stays O(cells + n_uniq x 2880).
struct MatrixTable { std::vector<double> uniq; std::vector<int32_t> cell_to_uniq; int n_uniq = 0; };

bool load_matrices(const Atmosphere& a, int l, MatrixTable& m, std::string& err) {
    const size_t cells = (size_t)a.cells;
    std::vector<uint64_t> h1(cells, 1469598103934665603ull), h2(cells, 0x9E3779B97F4A7C15ull);
    std::vector<double> run(cells);
    artes_host::FitsSlab slab;
    if (!slab.open(a.fits_path, 8, err)) return false;
    if (slab.elems != cells * (size_t)a.nl * 2880) { err = "scattermatrix HDU has the wrong size"; return false; }
    auto run_offset = [&](int ang, int e) { return cells * ((size_t)l + (size_t)a.nl * ((size_t)e + 16 * (size_t)ang)); };
    for (int ang = 0; ang < 180; ++ang)
        for (int e = 0; e < 16; ++e) {
            if (!slab.read(run_offset(ang, e), cells, run.data(), err)) return false;
            for (size_t c = 0; c < cells; ++c) {
                uint64_t bits; std::memcpy(&bits, &run[c], 8);
                h1[c] = (h1[c] ^ bits) * 1099511628211ull;
                h2[c] = (h2[c] + bits) * 0xD6E8FEB86659FD93ull; h2[c] ^= h2[c] >> 32;
            }
        }
    struct Key { uint64_t a, b; bool operator==(const Key& o) const { return a == o.a && b == o.b; } };
    struct KeyHash { size_t operator()(const Key& k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ull)); } };
    std::unordered_map<Key, int, KeyHash> ids;
    std::vector<size_t> rep;
    m.cell_to_uniq.resize(cells);
    for (size_t c = 0; c < cells; ++c) {
        auto it = ids.find(Key{h1[c], h2[c]});
        if (it == ids.end()) { it = ids.emplace(Key{h1[c], h2[c]}, (int)rep.size()).first; rep.push_back(c); }
        m.cell_to_uniq[c] = it->second;
    }
    m.n_uniq = (int)rep.size();
    m.uniq.assign((size_t)m.n_uniq * 2880, 0.0);
    for (int ang = 0; ang < 180; ++ang)
        for (int e = 0; e < 16; ++e) {
            if (!slab.read(run_offset(ang, e), cells, run.data(), err)) return false;
            for (int u = 0; u < m.n_uniq; ++u) m.uniq[((size_t)u * 180 + ang) * 16 + e] = run[rep[u]];
        }
    return true;
}

// grid_initialize(2): cell_depth :2329-2393
int cell_depth(const Atmosphere& a, const Config& c, int l) {
    const double limit = (c.photon_source == 1) ? 30.0 : 5.0;
    const int grid_out = (c.photon_source == 2 && c.ring) ? 2 : 0;
    int cell_max = 1000000, depth = 0;
    for (int j = 0; j < a.nt; ++j)
        for (int k = 0; k < a.np; ++k) {
            double tot = 0.0;
            for (int i = grid_out; i < a.nr; ++i) {
                const size_t idx = (size_t)l * a.cells + (size_t)(a.nr - i - 1) + (size_t)a.nr * (j + (size_t)a.nt * k);
                const double kap = (c.photon_source == 1) ? a.k_sca[idx] + a.k_abs[idx] : a.k_abs[idx];
                tot = tot + kap * (a.rfront[a.nr - i] - a.rfront[a.nr - i - 1]);
                depth = a.nr - i - 1;
                if (tot > limit) break;
            }
            cell_max = std::min(cell_max, depth);
        }
    return cell_max;
}

// cell volumes :2272-2307 and the thermal tables :2395-2453 (planet source)
struct Thermal { std::vector<double> cell_weight, luminosity, cdf; double total = 0.0; };

void thermal_tables(const Atmosphere& a, const Config& c, int l, int depth, double ox, double oy, double oz, Thermal& t) {
    const size_t n = (size_t)a.cells;
    std::vector<double> vol(n);
    for (int k = 0; k < a.np; ++k) {
        const double dphi = (a.np == 1) ? 2.0 * PI : ((k < a.np - 1) ? a.phifront[k + 1] - a.phifront[k] : 2.0 * PI - a.phifront[k]);
        for (int j = 0; j < a.nt; ++j)
            for (int i = 0; i < a.nr; ++i)
                vol[i + (size_t)a.nr * (j + (size_t)a.nt * k)] = ox * oy * oz * (1.0 / 3.0) *
                    (std::pow(a.rfront[i + 1], 3) - std::pow(a.rfront[i], 3)) * (std::cos(a.thetafront[j]) - std::cos(a.thetafront[j + 1])) * dphi;
    }
    const double wl = a.wavelengths[l];
    auto at = [&](int i, int j, int k) { return (size_t)i + (size_t)a.nr * (j + (size_t)a.nt * k); };
    double weight_norm = 0.0;
    for (int i = depth; i < a.nr; ++i)
        for (int j = 0; j < a.nt; ++j)
            for (int k = 0; k < a.np; ++k) {
                const size_t x = at(i, j, k);
                if (a.temperature[x] > 0.0) weight_norm += a.k_abs[(size_t)l * n + x] * planck_function(a.temperature[x], wl, 2) * vol[x];
            }
    t.cell_weight.assign(n, 0.0); t.luminosity.assign(n, 0.0); t.cdf.assign(n, 0.0);
    double total = 0.0;
    for (int i = depth; i < a.nr; ++i)
        for (int j = 0; j < a.nt; ++j)
            for (int k = 0; k < a.np; ++k) {
                const size_t x = at(i, j, k);
                const double ka = a.k_abs[(size_t)l * n + x];
                if (a.temperature[x] > 0.0 && ka > 0.0) {
                    const double pf = planck_function(a.temperature[x], wl, 2);
                    t.cell_weight[x] = c.thermal_weight ? weight_norm / (vol[x] * ka * pf) : 1.0;
                    t.luminosity[x] = 4.0 * PI * vol[x] * ka * pf;
                    total = total + t.luminosity[x] * t.cell_weight[x];
                }
                t.cdf[x] = total;
            }
    t.total = total;
}

// list-directed real output of gfortran (list_directed.h): separator blank + G25.17E3 field
std::string fnum(double v) { return artes_host::ld_real(v); }

struct Run {
    Config c;
    Atmosphere a;
    std::string atmosphere, output_name, outdir;
    double x_max = 0, x_fov = 0, pixel_scale = 0, ox = 1, oy = 1, oz = 1;
    FILE* log = stdout;
    std::map<int, uint64_t> errors;
    uint64_t packets_done = 0;
    double gpu_ms = 0.0;
};

void append_line(const std::string& path, const std::string& header, const std::string& line) {
    const bool exists = fs::exists(path);
    std::ofstream f(path, std::ios::app);
    if (!exists && !header.empty()) f << header << "\n\n";
    f << line << "\n";
}

// photon_package :2509-2539
double package_energy(const Run& r, double wavelength, double det_phi, double packages, double emis_total) {
    const Config& c = r.c;
    if (c.photon_source == 1) {
        double e = PI * planck_function(c.t_star, wavelength, 1) * r.a.rfront[r.a.nr] * r.a.rfront[r.a.nr] * c.r_star * c.r_star /
                   (c.orbit * c.orbit * c.distance_planet * c.distance_planet * packages);
        if (c.phase_curve && det_phi * 180.0 / PI >= 170.0)
            e = e * (PI * c.r_star * c.r_star - 0.9 * 0.9 * PI * c.r_star * c.r_star) / (PI * c.r_star * c.r_star);
        return e;
    }
    return emis_total / (c.distance_planet * c.distance_planet * packages);
}

// ---- argument_input :4232-4258, artes.in (initialize :384-397), argument_keywords :4260-4309
void read_inputs(int argc, char** argv, Run& R, bool dryrun, std::string& atmosphere_directory) {
    Config& c = R.c;
    R.atmosphere = argv[1];
    c.packages = std::floor(fortran_real(argv[2]));
    atmosphere_directory = "input/" + R.atmosphere;
    const std::string input_file = atmosphere_directory + "/artes.in";
    if (!fs::exists(input_file)) die("Input file does not exist!");
    {
        std::ifstream f(input_file);
        std::string line;
        while (std::getline(f, line)) {
            std::string trimmed = line;
            while (!trimmed.empty() && (trimmed.back() == ' ' || trimmed.back() == '\r' || trimmed.back() == '\t')) trimmed.pop_back();
            if (trimmed.empty()) continue;
            const char first = line[0];
            if (first == '*' || first == '-' || first == '=') continue;
            std::string key, value;
            get_key_value(line, key, value);
            input_parameters(c, key, value);
        }
    }
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        if (arg == "-o" && i + 1 < argc) {
            R.output_name = argv[i + 1];
            R.outdir = "output/" + R.output_name;
            if (!dryrun) {
                std::error_code ec;
                fs::remove_all(R.outdir, ec);
                fs::create_directories(R.outdir + "/input"); fs::create_directories(R.outdir + "/output"); fs::create_directories(R.outdir + "/plot");
                for (const char* name : {"artes.in", "atmosphere.in", "atmosphere.fits", "atmosphere.dat", "pressureTemperature.dat"}) {
                    const std::string src = atmosphere_directory + "/" + name;
                    if (fs::exists(src)) fs::copy_file(src, R.outdir + "/input/" + name, fs::copy_options::overwrite_existing, ec);
                }
            }
        } else if (arg == "-k" && i + 1 < argc) {
            std::string key, value;
            get_key_value(argv[i + 1], key, value);
            input_parameters(c, key, value);
            if (!dryrun && !R.output_name.empty()) { std::ofstream f(R.outdir + "/input/artes.in", std::ios::app); f << "\n" << argv[i + 1] << "\n"; }
        }
    }
    if (R.output_name.empty()) die("No output directory given (-o [outputDirectory])");
}

// ---- detector, oblateness, field of view (initialize :451-514); returns the observer's phase angle in degrees (0 for a phase curve)
double derive_detector(Run& R) {
    Config& c = R.c;
    const Atmosphere& a = R.a;
    if (c.spectrum) { c.nx = 1; c.ny = 1; }
    else if (c.phase_curve) { c.nx = 1; c.ny = 1; c.det_theta = PI / 2.0; c.det_phi = 1.e-5; }
    R.ox = 1.0 / (1.0 - c.oblateness); R.oy = R.ox; R.oz = 1.0;
    R.x_max = (c.oblateness + 1.0) * 1.3 * a.rfront[a.nr];
    R.x_fov = 2.0 * std::atan(R.x_max / c.distance_planet) * 3600.0 * 180.0 / PI * 1000.0;
    R.pixel_scale = R.x_fov / c.nx;
    if (std::fabs(c.det_phi) < 1.e-3 || c.det_phi > 2.0 * PI - 1.e-3) c.det_phi = 1.e-3;
    if (c.det_phi > PI - 1.e-3 && c.det_phi < PI + 1.e-3) c.det_phi = PI - 1.e-3;
    if (c.phase_curve) return 0.0;
    return std::acos(std::sin(c.theta_star) * std::cos(c.phi_star) * std::sin(c.det_theta) * std::cos(c.det_phi) +
                     std::sin(c.theta_star) * std::sin(c.phi_star) * std::sin(c.det_theta) * std::sin(c.det_phi) +
                     std::cos(c.theta_star) * std::cos(c.det_theta)) * 180.0 / PI;
}

void print_dryrun(const Run& R) {
    const Config& c = R.c;
    const Atmosphere& a = R.a;
    std::printf("ARTES dry run\n atmosphere=%s output=%s photons=%.0f source=%d\n grid nr=%d ntheta=%d nphi=%d nlambda=%d\n"
                " detector type=%s theta=%.6f phi=%.6f pixels=%d\n fstop=%g minimum=%g albedo=%g oblateness=%g\n gpu devices=%d seed=%llu mode=%s\n",
                R.atmosphere.c_str(), R.output_name.c_str(), c.packages, c.photon_source, a.nr, a.nt, a.np, a.nl,
                c.phase_curve ? "phase" : c.spectrum ? "spectrum" : c.imaging_mono ? "imaging_mono" : c.imaging_broad ? "imaging_broad" : "none",
                c.det_theta, c.det_phi, c.nx, c.fstop, c.photon_minimum, c.surface_albedo, c.oblateness, c.devices,
                (unsigned long long)c.seed, c.mode == ARTES_MODE_FAST ? "fast" : "faithful");
}

}  // namespace

// detector = thread sum x package energy (:959-975), photometry (:977-1004) and the Stokes errors of write_output (:3481-3519);
// layout (ix, iy, stokes, l): sums / detector 12 planes of npx pixels, error 5 planes, photometry 11 numbers
void finish_detector(size_t npx, const std::vector<double>& sums, double energy, std::vector<double>& detector, double* photometry,
                     std::vector<double>& error) {
    for (size_t i = 0; i < 4 * npx; ++i) {
        detector[i] = sums[i] * energy; detector[4 * npx + i] = sums[4 * npx + i] * energy * energy; detector[8 * npx + i] = sums[8 * npx + i];
    }
    std::fill(photometry, photometry + 11, 0.0);
    for (int s = 0; s < 4; ++s) {
        double sum1 = 0, sum2 = 0, n = 0;
        for (size_t i = 0; i < npx; ++i) { sum1 += detector[s * npx + i]; sum2 += detector[(4 + s) * npx + i]; n += detector[(8 + s) * npx + i]; }
        photometry[2 * s] = sum1;
        if (n > 0.0) { const double d = sum2 / n - (sum1 / n) * (sum1 / n); if (d > 0.0) photometry[2 * s + 1] = std::sqrt(d) * std::sqrt(n); }
    }
    photometry[8] = std::sqrt(photometry[2] * photometry[2] + photometry[4] * photometry[4]);
    photometry[9] = photometry[0] != 0.0 ? photometry[8] / photometry[0] : 0.0;
    // Stokes errors write_output :3481-3519
    std::fill(error.begin(), error.end(), 0.0);
    for (int s = 0; s < 4; ++s)
        for (size_t i = 0; i < npx; ++i) {
            const double n = detector[(8 + s) * npx + i];
            if (n > 0.0) {
                const double d = detector[(4 + s) * npx + i] / n - std::pow(detector[s * npx + i] / n, 2);
                if (d > 0.0) error[s * npx + i] = std::sqrt(d) * std::sqrt(n);
            }
        }
    for (size_t i = 0; i < npx; ++i) {
        const double I = detector[i], Q = detector[npx + i], U = detector[2 * npx + i];
        if (Q * Q + U * U > 0.0 && I > 0.0) {
            const double pol = std::sqrt(Q * Q + U * U);
            const double dpol = std::sqrt((std::pow(Q * error[npx + i], 2) + std::pow(U * error[2 * npx + i], 2)) / (2.0 * (Q * Q + U * U)));
            error[4 * npx + i] = (pol / I) * std::sqrt(std::pow(dpol / pol, 2) + std::pow(error[i] / I, 2));
        }
    }
}

int main(int argc, char** argv) {
    if (argc <= 2) {
        std::printf("How to run ARTES:\n./bin/ARTES [inputDirectory] [photons] -o [outputDirectory] -k [keyWord]=[value]\n");
        return 0;
    }
    Run R;
    Config& c = R.c;
    const bool dryrun = std::getenv("ARTES_DRYRUN") != nullptr;
    std::string atmosphere_directory;
    read_inputs(argc, argv, R, dryrun, atmosphere_directory);

    // ---- get_atmosphere
    get_atmosphere(atmosphere_directory + "/atmosphere.fits", R.a);
    const Atmosphere& a = R.a;
    const double phase_observer = derive_detector(R);
    if (dryrun) { print_dryrun(R); return 0; }

    { std::ofstream f(R.outdir + "/error.log"); }
    if (c.log_file) R.log = std::fopen((R.outdir + "/output.log").c_str(), "w");
    FILE* L = R.log ? R.log : stdout;

    // ---- python :1328-1348
    {
        FILE* f = std::fopen((R.outdir + "/plot.dat").c_str(), "w");
        std::fprintf(f, "[plot]\nphoton_source=%d\ndistance=%.7E\nplanet_radius=%.7E\nntheta=%d\nfov=%.7E\n", c.photon_source, c.distance_planet,
                     a.rfront[0], a.nt, R.x_fov);
        std::fclose(f);
    }

    // ---- GPU context
    artes_gpu_ctx* ctx = nullptr;
    if (artes_gpu_create(&ctx, c.devices, nullptr) != 0) {
        std::fprintf(stderr, "ARTES: %s\n", artes_gpu_last_error(nullptr));
        return 1;
    }
#define GPU(call) do { if ((call) != 0) { std::fprintf(stderr, "ARTES: %s: %s\n", #call, artes_gpu_last_error(ctx)); return 1; } } while (0)
    GPU(artes_gpu_set_grid(ctx, a.nr, a.nt, a.np, a.rfront.data(), a.thetafront.data(), a.thetaplane.data(), a.phifront.data(), R.ox, R.oy, R.oz));
    char devname[128] = "";
    int sms = 0, ccM = 0, ccm = 0;
    artes_gpu_device_info(ctx, &sms, &ccM, &ccm, devname, sizeof(devname));

    std::fprintf(L, "--------------------------------------------------------------\n");
    std::fprintf(L, "ARTES  --  photon-packet transport on %d x %s (sm_%d%d, %d SMs)\n", c.devices, devname, ccM, ccm, sms);
    std::fprintf(L, "--------------------------------------------------------------\n");
    std::fprintf(L, "Atmosphere: %s   photons: %.0f   source: %s\n", R.atmosphere.c_str(), c.packages, c.photon_source == 1 ? "star" : "planet");
    std::fprintf(L, "Grid: nr=%d ntheta=%d nphi=%d   wavelengths: %d\n", a.nr, a.nt, a.np, a.nl);
    if (!c.phase_curve) std::fprintf(L, "Observer phase angle [deg]: %.3f\n", phase_observer);
    std::fflush(L);

    const size_t npx = (size_t)c.nx * c.ny;
    const uint64_t packages = (uint64_t)c.packages;
    std::vector<double> det_sum(12 * npx), det_acc(12 * npx, 0.0), flow4, flow3;
    if (c.flow_theta) flow4.assign((size_t)4 * a.cells, 0.0);
    if (c.flow_global) flow3.assign((size_t)3 * a.cells, 0.0);
    double flux[2] = {0, 0};
    uint64_t err_hist[ARTES_ERR_SLOTS];
    Thermal thermal;
    int depth = 0;
    double photometry[11];
    const auto t_start = std::chrono::steady_clock::now();

    auto write_optical_depth = [&](int l) {
        if (c.imaging_broad || c.spectrum) {   // optical_depth.dat :2459-2491
            double tt = 0, ts = 0, ta = 0;
            for (int i = 0; i < a.nr; ++i) {
                const double dr = a.rfront[i + 1] - a.rfront[i];
                const double ks = a.k_sca[(size_t)l * a.cells + i], ka = a.k_abs[(size_t)l * a.cells + i];
                tt += dr * (ks + ka); ts += dr * ks; ta += dr * ka;
            }
            append_line(R.outdir + "/output/optical_depth.dat",
                        " # Wavelength [micron] - Total optical depth - Absorption optical depth - Scattering optical depth",
                        fnum(a.wavelengths[l] * 1.e6) + fnum(tt) + fnum(ta) + fnum(ts));
        }
    };

    // Every call of radiative_transfer walks its own photon-id range (call k: k*packages + [0, packages)): the reference
    // carries its generator state from one call to the next (:4197-4230), so successive launches are statistically
    // independent; replaying ids 0.. for every wavelength / phase angle would correlate them and understate error.fits.
    // The batched paths number their launches the same way, so both paths produce the same streams.
    uint64_t launch_index = 0;
    auto make_launch = [&](double det_phi) -> artes_launch_t {
        artes_launch_t Ln;
        std::memset(&Ln, 0, sizeof(Ln));
        Ln.struct_size = sizeof(Ln); Ln.mode = c.mode; Ln.n_photons = packages; Ln.photon_id_base = launch_index * packages; Ln.seed = c.seed;
        Ln.photon_source = c.photon_source; Ln.photon_scattering = c.photon_scattering ? 1 : 0; Ln.photon_emission = c.photon_emission;
        Ln.stellar_direction = c.stellar_direction ? 1 : 0;
        Ln.limb_emission = (c.phase_curve && det_phi * 180.0 / PI >= 170.0) ? 1 : 0;
        Ln.flow_global = c.flow_global ? 1 : 0; Ln.flow_theta = c.flow_theta ? 1 : 0; Ln.nx = c.nx; Ln.ny = c.ny;
        Ln.fstop = c.fstop; Ln.photon_minimum = c.photon_minimum; Ln.photon_bias = c.photon_bias; Ln.surface_albedo = c.surface_albedo;
        Ln.theta_star = c.theta_star; Ln.phi_star = c.phi_star; Ln.det_theta = c.det_theta; Ln.det_phi = det_phi;
        Ln.x_max = R.x_max; Ln.y_max = R.x_max;
        return Ln;
    };

    // grid_initialize(2) + table upload for wavelength index l
    auto prepare_wavelength = [&](int l) -> int {
        depth = cell_depth(a, c, l);
        MatrixTable mt;
        std::string err;
        if (!load_matrices(a, l, mt, err)) { std::fprintf(stderr, "ARTES: %s\n", err.c_str()); return 1; }
        const double* cw = nullptr; const double* cdf = nullptr;
        if (c.photon_source == 2) { thermal_tables(a, c, l, depth, R.ox, R.oy, R.oz, thermal); cw = thermal.cell_weight.data(); cdf = thermal.cdf.data(); }
        if (artes_gpu_set_wavelength(ctx, a.k_sca.data() + (size_t)l * a.cells, a.k_abs.data() + (size_t)l * a.cells, mt.n_uniq, mt.uniq.data(),
                                     mt.cell_to_uniq.data(), depth, cw, cdf) != 0) {
            std::fprintf(stderr, "ARTES: set_wavelength: %s\n", artes_gpu_last_error(ctx));
            return 1;
        }
        write_optical_depth(l);
        return 0;
    };
    // all wavelengths at once (star source): stacked tables for artes_gpu_set_wavelengths, one common matrix list
    std::vector<int32_t> wl_depths;
    std::vector<Thermal> wl_thermal;     // thermal source: emissivity tables of every wavelength
    auto prepare_all_wavelengths = [&]() -> int {
        std::vector<double> uniq, cw_all, cdf_all;
        if (c.photon_source == 2) { wl_thermal.assign(a.nl, Thermal()); cw_all.resize((size_t)a.nl * a.cells); cdf_all.resize((size_t)a.nl * a.cells); }
        std::vector<int32_t> c2u((size_t)a.nl * a.cells);
        int n_uniq = 0;
        wl_depths.assign(a.nl, 0);
        for (int l = 0; l < a.nl; ++l) {
            wl_depths[l] = cell_depth(a, c, l);
            MatrixTable mt;
            std::string err;
            if (!load_matrices(a, l, mt, err)) { std::fprintf(stderr, "ARTES: %s\n", err.c_str()); return 1; }
            uniq.insert(uniq.end(), mt.uniq.begin(), mt.uniq.end());
            for (size_t i = 0; i < (size_t)a.cells; ++i) c2u[(size_t)l * a.cells + i] = mt.cell_to_uniq[i] + n_uniq;
            n_uniq += mt.n_uniq;
            if (c.photon_source == 2) {
                thermal_tables(a, c, l, wl_depths[l], R.ox, R.oy, R.oz, wl_thermal[l]);
                std::copy(wl_thermal[l].cell_weight.begin(), wl_thermal[l].cell_weight.end(), cw_all.begin() + (size_t)l * a.cells);
                std::copy(wl_thermal[l].cdf.begin(), wl_thermal[l].cdf.end(), cdf_all.begin() + (size_t)l * a.cells);
            }
            write_optical_depth(l);
        }
        if (artes_gpu_set_wavelengths(ctx, a.nl, a.k_sca.data(), a.k_abs.data(), n_uniq, uniq.data(), c2u.data(), wl_depths.data(),
                                      c.photon_source == 2 ? cw_all.data() : nullptr, c.photon_source == 2 ? cdf_all.data() : nullptr) != 0) {
            std::fprintf(stderr, "ARTES: set_wavelengths: %s\n", artes_gpu_last_error(ctx));
            return 1;
        }
        return 0;
    };
    // the wl_count loop (:132-204) as ONE batched launch: launch l = wavelength l, photon ids l*packages + [0, packages)
    std::vector<double> det_all, flux_all;
    auto run_all_wavelengths = [&]() -> int {
        std::vector<artes_launch_t> Ls;
        for (int l = 0; l < a.nl; ++l) { artes_launch_t Ln = make_launch(c.det_phi); Ln.wl_index = l; Ls.push_back(Ln); }
        det_all.assign((size_t)a.nl * 12 * npx, 0.0); flux_all.assign((size_t)2 * a.nl, 0.0);
        const int n_chunks = (a.nl + ARTES_MAX_BATCH - 1) / ARTES_MAX_BATCH;
        std::fprintf(stdout, "Wavelengths: %d, %s\n", a.nl, n_chunks == 1 ? "one batched launch" : "batched launches of 256 wavelengths"); std::fflush(stdout);
        for (int l0 = 0; l0 < a.nl; l0 += ARTES_MAX_BATCH) {      // a batch holds at most ARTES_MAX_BATCH launches
            const int nb = std::min(ARTES_MAX_BATCH, a.nl - l0);
            for (int l = l0; l < l0 + nb; ++l) Ls[l].photon_id_base = (uint64_t)l0 * packages;   // launch l walks ids l*packages + [0, packages)
            artes_stats_t st;
            if (artes_gpu_run_batch(ctx, Ls.data() + l0, nb, det_all.data() + (size_t)l0 * 12 * npx, flux_all.data() + 2 * l0, err_hist, &st) != 0) {
                std::fprintf(stderr, "ARTES: artes_gpu_run_batch: %s\n", artes_gpu_last_error(ctx));
                return 1;
            }
            for (int k = 0; k < ARTES_ERR_SLOTS; ++k) if (err_hist[k]) R.errors[k] += err_hist[k];
            R.packets_done += (uint64_t)nb * packages; R.gpu_ms += st.kernel_ms + st.reduce_ms;
        }
        return 0;
    };
    // `call radiative_transfer`
    auto radiative_transfer = [&](double det_phi) -> int {
        const artes_launch_t Ln = make_launch(det_phi);
        ++launch_index;
        artes_stats_t st;
        if (artes_gpu_run(ctx, &Ln, det_sum.data(), flux, flow4.empty() ? nullptr : flow4.data(), flow3.empty() ? nullptr : flow3.data(),
                          err_hist, &st) != 0) {
            std::fprintf(stderr, "ARTES: artes_gpu_run: %s\n", artes_gpu_last_error(ctx));
            return 1;
        }
        for (int k = 0; k < ARTES_ERR_SLOTS; ++k) if (err_hist[k]) R.errors[k] += err_hist[k];
        R.packets_done += packages; R.gpu_ms += st.kernel_ms + st.reduce_ms;
        return 0;
    };

    // detector = thread sum x package energy (:959-975) and photometry (:977-1004); det layout (ix, iy, stokes, l)
    std::vector<double> detector(12 * npx), error(5 * npx);
    auto finish_detector = [&](const std::vector<double>& sums, double energy) { ::finish_detector(npx, sums, energy, detector, photometry, error); };

    // write_output :3472-3772
    auto write_output = [&](int l, double det_phi) {
        const double wavelength = a.wavelengths[l];
        const std::string out = R.outdir + "/output/";
        std::string err;
        const std::string hdr = " # Wavelength [micron] - Stokes I, Q, U, V [W m-2 micron-1]";
        if (c.phase_curve) {
            const double deg = det_phi * 180.0 / PI;
            const double ph = deg < 1.0 ? 0.0 : (deg > 179.0 ? 180.0 : deg);
            std::string line = fnum(ph);
            for (int s = 0; s < 4; ++s) line += fnum(detector[s * npx] * 1.e-6) + fnum(error[s * npx] * 1.e-6);
            append_line(out + "phase.dat", hdr, line);
        } else if (c.imaging_mono || c.imaging_broad) {
            std::vector<double> img(4 * npx);
            for (size_t i = 0; i < 4 * npx; ++i) img[i] = detector[i] * 1.e-6 / (R.pixel_scale * R.pixel_scale);
            artes_host::fits_write_image(out + "stokes.fits", {c.nx, c.ny, 4}, img.data(), err);
            artes_host::fits_write_image(out + "error.fits", {c.nx, c.ny, 5}, error.data(), err);
            if (c.imaging_mono) {
                std::string line = fnum(wavelength * 1.e6);
                for (int k = 0; k < 8; ++k) line += fnum(1.e-6 * photometry[k]);
                append_line(out + "photometry.dat", hdr, line);
            }
        } else if (c.spectrum) {
            std::string line = fnum(wavelength * 1.e6);
            for (int s = 0; s < 4; ++s) line += fnum(1.e-6 * detector[s * npx]);
            append_line(out + "spectrum.dat", hdr, line);
        }
        if (c.photon_source == 1) {
            if ((c.phase_curve && det_phi < PI / 180.0) || !c.phase_curve) {
                const double pf = planck_function(c.t_star, wavelength, 1);
                append_line(out + "normalization.dat", "",
                            fnum(wavelength * 1.e6) + fnum(1.e-6 * pf * c.r_star * c.r_star / (c.distance_planet * c.distance_planet)) +
                            fnum(1.e-6 * pf * a.rfront[a.nr] * a.rfront[a.nr] * c.r_star * c.r_star /
                                 (c.orbit * c.orbit * c.distance_planet * c.distance_planet)));
            }
        } else {
            if (c.imaging_mono) artes_host::fits_write_image(out + "cell_luminosity.fits", {a.nr, a.nt, a.np}, thermal.luminosity.data(), err);
            const double e_pack = thermal.total / (double)packages;
            append_line(out + "luminosity.dat",
                        " # Wavelength [deg] - Emitted luminosity [W micron-1] -  Emergent luminosity [W micron-1] - Emergent luminosity [a.u.]",
                        fnum(wavelength) + fnum(flux[0] * e_pack * 1.e-6) + fnum(flux[1] * e_pack * 1.e-6) + fnum(flux[1]));
        }
        if (c.imaging_mono || c.spectrum) {
            append_line(out + "cell_depth.dat", " # Wavelength [micron] - Cell depth", fnum(wavelength * 1.e6) + artes_host::ld_int(depth));
        }
        if (c.flow_global) {   // unit vectors :3715-3738
            std::vector<double> tr(flow3.size(), 0.0);
            for (int k = 0; k < a.np; ++k) for (int j = 0; j < a.nt; ++j) for (int i = depth; i < a.nr; ++i) {
                const size_t x = 3 * ((size_t)i + (size_t)a.nr * (j + (size_t)a.nt * k));
                const double n = std::sqrt(flow3[x] * flow3[x] + flow3[x + 1] * flow3[x + 1] + flow3[x + 2] * flow3[x + 2]);
                for (int m = 0; m < 3; ++m) tr[x + m] = n > 0.0 ? flow3[x + m] / n : flow3[x + m];
            }
            artes_host::fits_write_image(out + "flow_global.fits", {3, a.nr, a.nt, a.np}, tr.data(), err);
        }
        if (c.flow_theta) {    // normalised to the emergent flux :3744-3766
            std::vector<double> tr(flow4.size(), 0.0);
            for (int k = 0; k < a.np; ++k) for (int j = 0; j < a.nt; ++j) for (int i = depth; i < a.nr; ++i) {
                const size_t x = 4 * ((size_t)i + (size_t)a.nr * (j + (size_t)a.nt * k));
                for (int m = 0; m < 4; ++m) tr[x + m] = flow4[x + m] / flux[1];
            }
            artes_host::fits_write_image(out + "flow_latitudinal.fits", {4, a.nr, a.nt, a.np}, tr.data(), err);
        }
    };

    // ---- run :121-267
    int rc = 0;
    // without flow counters: every wavelength of a spectrum / broadband image in one batched launch
    const bool batch_wl = (c.spectrum || c.imaging_broad) && !c.flow_global && !c.flow_theta && a.nl > 1;
    if (batch_wl) {
        if (!(rc = prepare_all_wavelengths()) && !(rc = run_all_wavelengths())) {
            for (int l = 0; l < a.nl; ++l) {
                depth = wl_depths[l];
                if (c.photon_source == 2) thermal = wl_thermal[l];
                flux[0] = flux_all[2 * l]; flux[1] = flux_all[2 * l + 1];
                if (c.spectrum) {
                    std::copy(det_all.begin() + (size_t)l * 12 * npx, det_all.begin() + (size_t)(l + 1) * 12 * npx, det_sum.begin());
                    finish_detector(det_sum, package_energy(R, a.wavelengths[l], c.det_phi, (double)packages, thermal.total));
                    write_output(l, c.det_phi);
                } else {
                    for (size_t i = 0; i < det_acc.size(); ++i) det_acc[i] += det_all[(size_t)l * 12 * npx + i];
                }
            }
            if (!c.spectrum && c.imaging_broad) {   // (`if (spectrum) ... else if (imaging_broad)`, :132/167) scaled with the LAST wavelength's package energy (:175-200, 959-975)
                finish_detector(det_acc, package_energy(R, a.wavelengths[a.nl - 1], c.det_phi, (double)packages, thermal.total));
                write_output(a.nl - 1, c.det_phi);
            }
        }
    } else if (c.spectrum) {
        for (int l = 0; l < a.nl && !rc; ++l) {
            if ((rc = prepare_wavelength(l))) break;
            std::fprintf(stdout, "\rWavelength: %7.3f micron", a.wavelengths[l] * 1.e6); std::fflush(stdout);
            if ((rc = radiative_transfer(c.det_phi))) break;
            finish_detector(det_sum, package_energy(R, a.wavelengths[l], c.det_phi, (double)packages, thermal.total));
            write_output(l, c.det_phi);
        }
        std::fprintf(stdout, "\n");
    } else if (c.imaging_broad) {
        // the detector accumulates over the wavelengths and is scaled with the LAST wavelength's package energy (:175-200, 959-975)
        for (int l = 0; l < a.nl && !rc; ++l) {
            if ((rc = prepare_wavelength(l))) break;
            std::fprintf(stdout, "\rWavelength: %6.3f micron", a.wavelengths[l] * 1.e6); std::fflush(stdout);
            if ((rc = radiative_transfer(c.det_phi))) break;
            for (size_t i = 0; i < det_acc.size(); ++i) det_acc[i] += det_sum[i];
        }
        std::fprintf(stdout, "\n");
        if (!rc) {
            finish_detector(det_acc, package_energy(R, a.wavelengths[a.nl - 1], c.det_phi, (double)packages, thermal.total));
            write_output(a.nl - 1, c.det_phi);
        }
    } else if (c.phase_curve || c.imaging_mono) {
        if (!(rc = prepare_wavelength(0))) {
            if (c.phase_curve) {
                std::vector<double> phis(73);                // :215-245
                for (int i = 1; i <= 73; ++i) {
                    if (i == 1) phis[0] = 1.e-5 * PI / 180.0;
                    else if (i == 2) phis[1] = 2.5 * PI / 180.0;
                    else if (i == 73) phis[72] = (180.0 - 1e-5) * PI / 180.0;
                    else phis[i - 1] = phis[i - 2] + 2.5 * PI / 180.0;
                }
                if (!c.flow_global && !c.flow_theta) {
                    // the 73 calls of radiative_transfer (:215-245) on the GPU:
                    //  - gpu:phase_walks=single (default): every packet is walked ONCE and observed from all azimuths (peel-off does
                    //    not disturb the walk): one artes_gpu_run_multi call for the angles below 170 deg, one for the limb-biased
                    //    angles (a different emission law, :1041-1055).  Each angle has the expectation value and the noise of the
                    //    reference's run with `packages` packets; the angles share their walks, so their noise is correlated.
                    //  - gpu:phase_walks=independent: ONE batched launch of 73 independent walks (artes_gpu_run_batch): angle k walks
                    //    the photon ids k*packages + [0, packages), statistically independent like the reference's runs.
                    std::vector<artes_launch_t> Ls;
                    for (double phi : phis) Ls.push_back(make_launch(phi));
                    std::vector<double> det_all((size_t)73 * 12 * npx), flux_all(2 * 73);
                    artes_stats_t st;
                    std::memset(&st, 0, sizeof(st));
                    uint64_t walked = 0;
                    bool failed = false;
                    if (c.phase_single_walk) {
                        std::fprintf(stdout, "Phase angles: 73, one walk per packet observed from all azimuths (two launches)\n"); std::fflush(stdout);
                        for (int flag = 0; flag < 2 && !failed; ++flag) {
                            std::vector<artes_launch_t> grp;
                            std::vector<int> idx;
                            for (int i = 0; i < 73; ++i) if (Ls[i].limb_emission == flag) { grp.push_back(Ls[i]); idx.push_back(i); }
                            if (grp.empty()) continue;
                            for (auto& g : grp) g.photon_id_base = (uint64_t)flag * packages;
                            std::vector<double> det_g(grp.size() * 12 * npx), flux_g(2 * grp.size());
                            artes_stats_t sg;
                            uint64_t eh[ARTES_ERR_SLOTS];
                            if (artes_gpu_run_multi(ctx, grp.data(), (int)grp.size(), det_g.data(), flux_g.data(), eh, &sg) != 0) {
                                std::fprintf(stderr, "ARTES: artes_gpu_run_multi: %s\n", artes_gpu_last_error(ctx));
                                failed = true;
                                break;
                            }
                            for (size_t j = 0; j < idx.size(); ++j) {
                                std::copy(det_g.begin() + j * 12 * npx, det_g.begin() + (j + 1) * 12 * npx, det_all.begin() + (size_t)idx[j] * 12 * npx);
                                flux_all[2 * idx[j]] = flux_g[2 * j]; flux_all[2 * idx[j] + 1] = flux_g[2 * j + 1];
                            }
                            for (int k = 0; k < ARTES_ERR_SLOTS; ++k) err_hist[k] = (flag == 0 ? 0 : err_hist[k]) + eh[k];
                            st.kernel_ms += sg.kernel_ms; st.reduce_ms += sg.reduce_ms;
                            walked += sg.n_emit;
                        }
                    } else {
                        std::fprintf(stdout, "Phase angles: 73, one batched launch\n"); std::fflush(stdout);
                        failed = artes_gpu_run_batch(ctx, Ls.data(), 73, det_all.data(), flux_all.data(), err_hist, &st) != 0;
                        if (failed) std::fprintf(stderr, "ARTES: artes_gpu_run_batch: %s\n", artes_gpu_last_error(ctx));
                        walked = 73 * packages;
                    }
                    if (failed) {
                        rc = 1;
                    } else {
                        for (int k = 0; k < ARTES_ERR_SLOTS; ++k) if (err_hist[k]) R.errors[k] += err_hist[k];
                        R.packets_done += walked; R.gpu_ms += st.kernel_ms + st.reduce_ms;
                        for (int i = 0; i < 73; ++i) {
                            std::copy(det_all.begin() + (size_t)i * 12 * npx, det_all.begin() + (size_t)(i + 1) * 12 * npx, det_sum.begin());
                            flux[0] = flux_all[2 * i]; flux[1] = flux_all[2 * i + 1];
                            finish_detector(det_sum, package_energy(R, a.wavelengths[0], phis[i], (double)packages, thermal.total));
                            write_output(0, phis[i]);
                        }
                    }
                } else {
                    for (int i = 0; i < 73 && !rc; ++i) {
                        std::fprintf(stdout, "\rPhase angle: %6.1f degrees", phis[i] * 180.0 / PI); std::fflush(stdout);
                        if ((rc = radiative_transfer(phis[i]))) break;
                        finish_detector(det_sum, package_energy(R, a.wavelengths[0], phis[i], (double)packages, thermal.total));
                        write_output(0, phis[i]);
                    }
                    std::fprintf(stdout, "\n");
                }
            } else {
                if (!(rc = radiative_transfer(c.det_phi))) {
                    finish_detector(det_sum, package_energy(R, a.wavelengths[0], c.det_phi, (double)packages, thermal.total));
                    write_output(0, c.det_phi);
                }
            }
        }
    } else {
        std::fprintf(stderr, "ARTES: no detector:type given (phase, spectrum, imaging_mono, imaging_broad)\n");
        rc = 1;
    }

    // ---- error.log (one line per code with its count; the reference appends one line per occurrence) and the closing log
    {
        std::ofstream f(R.outdir + "/error.log", std::ios::app);
        for (const auto& kv : R.errors) { char b[64]; std::snprintf(b, sizeof(b), " error %03d x %llu", kv.first, (unsigned long long)kv.second); f << b << "\n"; }
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::fprintf(L, "Photon packets: %llu   GPU time: %.3f s   wall time: %.3f s   rate: %.4e packets/s\n", (unsigned long long)R.packets_done,
                 R.gpu_ms * 1e-3, wall, R.gpu_ms > 0 ? R.packets_done / (R.gpu_ms * 1e-3) : 0.0);
    if (!R.errors.empty()) std::fprintf(L, "WARNING: check error log!\n");
    {   // perf.json next to the reference's outputs
        FILE* f = std::fopen((R.outdir + "/perf.json").c_str(), "w");
        if (f) {
            std::fprintf(f, "{\"packets\": %llu, \"gpu_s\": %.6f, \"wall_s\": %.6f, \"packets_per_s\": %.6e, \"devices\": %d}\n",
                         (unsigned long long)R.packets_done, R.gpu_ms * 1e-3, wall, R.gpu_ms > 0 ? R.packets_done / (R.gpu_ms * 1e-3) : 0.0, c.devices);
            std::fclose(f);
        }
    }
    if (R.log && R.log != stdout) std::fclose(R.log);
    artes_gpu_destroy(ctx);
    return rc;
}
