// artes_gpu.cu -- host side of the C-ABI declared in include/artes_gpu.h.
//
// Owns the device contexts, turns the reference's program-scope arrays into HBM-resident tables
// (SoA, de-duplicated matrices), launches the sm_100a transport kernels and reduces the per-device
// accumulators with NCCL (loaded lazily with dlopen so that a single-GPU run needs no NCCL at all).
// There is no CPU fallback: every entry point fails if no CUDA device is usable.

#include "../../include/artes_gpu.h"
#include "kernel_args.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace artes {
namespace faithful {
size_t smem_bytes(int nr, int nt, int np);
cudaError_t launch_transport(const KernelArgs& a, bool trace, int sm_count, cudaStream_t stream);
cudaError_t launch_transport4(const KernelArgs& a, bool trace, int sm_count, cudaStream_t stream);
size_t engine3_scratch_bytes(int sm_count);
cudaError_t launch_cell_face(const DevTables& T, unsigned long long n, const double* pos, const double* dir,
                             const int* face, const int* cell, int* out_i, double* out_d, cudaStream_t stream);
}  // namespace faithful
namespace fast {
size_t smem_bytes(int nr, int nt, int np);
cudaError_t launch_transport(const KernelArgs& a, bool trace, int sm_count, cudaStream_t stream);
cudaError_t launch_transport4(const KernelArgs& a, bool trace, int sm_count, cudaStream_t stream);
size_t engine3_scratch_bytes(int sm_count);
cudaError_t launch_cell_face(const DevTables& T, unsigned long long n, const double* pos, const double* dir,
                             const int* face, const int* cell, int* out_i, double* out_d, cudaStream_t stream);
bool engine2_supports(const KernelArgs& a);
bool engine2_batch_fits(const KernelArgs& a, int n);
int engine2_sdet_doubles(const KernelArgs& a, int n);
size_t engine2_scratch_bytes(int sm_count);
cudaError_t launch_transport2(const KernelArgs& a, int sm_count, cudaStream_t stream);
cudaError_t launch_transport2_trace(const KernelArgs& a, int sm_count, cudaStream_t stream);
}  // namespace fast
cudaError_t fma_peak(double* fp64_tflops, double* fp32_tflops, int sm_count, cudaStream_t stream);
cudaError_t ingest_cell_tables(const double* k_sca, const double* k_abs, const int* c2u, size_t n, double* kext, double* albedo,
                               double* cellrec, cudaStream_t stream);
void ingest_set_chunk_override(int planes);
cudaError_t ingest_dedup_device(const double* src, size_t cells, size_t plane_stride, cudaStream_t stream,
                                std::vector<double>& uniq, std::vector<int32_t>& c2u, int* exact, double* copy_ms, double* kernel_ms);
}  // namespace artes

using namespace artes;

namespace {

const double PI = 4.0 * std::atan(1.0);
std::string g_create_error;

// ---- lazily bound NCCL -------------------------------------------------------------------------
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string err;
    bool load() {
        if (ok) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (handle) break; }
        if (!handle) for (const char* n : names) { handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define BIND(sym) sym = reinterpret_cast<decltype(sym)>(dlsym(handle, "nccl" #sym)); if (!sym) { err = "missing symbol nccl" #sym; return false; }
        BIND(GetUniqueId) BIND(CommInitRank) BIND(CommInitAll) BIND(CommDestroy) BIND(AllReduce) BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
        ok = true;
        return true;
    }
};
Nccl g_nccl;

struct DeviceState {
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // kernel start/stop, reduce stop, spare
    std::vector<void*> grid_allocs, wl_allocs;
    std::vector<size_t> wl_caps;           // bytes of wl_allocs[i]: the wavelength tables are re-used when the next set fits
    size_t wl_next = 0;
    DevTables T{};
    double* out_d = nullptr;               // [det 10*npx | flux 2 | flow4 4*cells | flow3 3*cells]
    size_t out_d_cap = 0;
    unsigned long long* out_u = nullptr;   // [err 64 | stats 8 | counter 1]
    double* scratch = nullptr;             // ray/event engine: cold photon records (L2-resident working set)
    double* geo = nullptr;                 // batched launches: [n][8] detector geometry
    size_t geo_cap = 0;
    ncclComm_t comm = nullptr;
    unsigned long long n_photons = 0;
    int launches = 0;
};

}  // namespace

struct artes_gpu_ctx {
    std::vector<DeviceState> devs;
    std::string error;
    bool have_grid = false, have_wl = false, thermal = false;
    int nr = 0, nt = 0, np = 0, cells = 0;
    // host copies needed to rebuild per-wavelength tables
    std::vector<double> sinbeta, cos2beta, sin2beta;
    // NCCL across processes
    int nranks = 1, rank = 0;
    bool rank_comm = false;
    // pending async launch
    bool pending = false;
    artes_launch_t pending_launch{};
    size_t n_out_d = 0;
    double last_h2d_ms = 0.0;
    int last_engine = 0;
    // tables of several wavelengths at once (artes_gpu_set_wavelengths): per-cell tables are [n_wl][cells]
    int n_wl = 1;
    std::vector<int> wl_depth;
};

namespace {

int fail(artes_gpu_ctx* c, int code, const std::string& msg) {
    if (c) c->error = msg; else g_create_error = msg;
    return code;
}

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(ctx, -2, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define NC(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return fail(ctx, -4, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); } while (0)

template <typename Tp>
int upload(artes_gpu_ctx* ctx, DeviceState& d, std::vector<void*>& pool, const Tp* host, size_t n, const Tp** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(Tp);
    CU(cudaMalloc(&p, bytes));
    pool.push_back(p);
    if (n) CU(cudaMemcpyAsync(p, host, n * sizeof(Tp), cudaMemcpyHostToDevice, d.stream));
    *out = static_cast<const Tp*>(p);
    return 0;
}

// upload into the next buffer of the wavelength pool, (re)allocating only when it is too small: a spectrum calls
// set_wavelength once per wavelength with tables of the same size, and cudaMalloc / cudaFree cost more than the copy
template <typename Tp>
int upload_wl(artes_gpu_ctx* ctx, DeviceState& d, const Tp* host, size_t n, const Tp** out) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(Tp);
    const size_t i = d.wl_next++;
    if (i >= d.wl_allocs.size()) { d.wl_allocs.push_back(nullptr); d.wl_caps.push_back(0); }
    if (d.wl_caps[i] < bytes) {
        if (d.wl_allocs[i]) cudaFree(d.wl_allocs[i]);
        d.wl_allocs[i] = nullptr; d.wl_caps[i] = 0;
        CU(cudaMalloc(&d.wl_allocs[i], bytes));
        d.wl_caps[i] = bytes;
    }
    if (n && host) CU(cudaMemcpyAsync(d.wl_allocs[i], host, n * sizeof(Tp), cudaMemcpyHostToDevice, d.stream));      // host == null: device-built table
    *out = static_cast<const Tp*>(d.wl_allocs[i]);
    return 0;
}

void free_pool(std::vector<void*>& pool) {
    for (void* p : pool) cudaFree(p);
    pool.clear();
}

int ensure_outputs(artes_gpu_ctx* ctx, DeviceState& d, size_t n_d) {
    if (!d.out_u) CU(cudaMalloc(&d.out_u, (ARTES_ERR_SLOTS + 8 + 8) * sizeof(unsigned long long)));
    if (n_d > d.out_d_cap) {
        if (d.out_d) cudaFree(d.out_d);
        d.out_d = nullptr;
        CU(cudaMalloc(&d.out_d, n_d * sizeof(double)));
        d.out_d_cap = n_d;
    }
    return 0;
}

// Scheduling parameters are compile-time choices of the product library; the environment overrides exist only in a
// tuning build (make TUNING=1 -> -DARTES_TUNING), which tools/gpu_tune.py uses for the measurements quoted in DESIGN.md.
// photon records of the event-list engines (L2-resident working set): large enough for every engine
size_t scratch_bytes(int sm_count) {
    return std::max(fast::engine2_scratch_bytes(sm_count), std::max(fast::engine3_scratch_bytes(sm_count), faithful::engine3_scratch_bytes(sm_count)));
}

int env_int(const char* name, int dflt) {
#ifdef ARTES_TUNING
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    int x = std::atoi(v);
    return x < 1 ? 1 : (x > 4096 ? 4096 : x);
#else
    (void)name;
    return dflt;
#endif
}

// LaunchArgs from the ABI struct: host-evaluated detector / star geometry (src/ARTES.f90:495-502, 1080-1109, 4628, 4871)
void fill_launch(const artes_launch_t& L, LaunchArgs& a) {
    a.n_photons = L.n_photons; a.id_base = L.photon_id_base; a.seed = L.seed;
    a.photon_source = L.photon_source; a.photon_scattering = L.photon_scattering; a.photon_emission = L.photon_emission;
    a.stellar_direction = L.stellar_direction; a.limb_emission = L.limb_emission;
    a.flow_global = L.flow_global; a.flow_theta = L.flow_theta; a.nx = L.nx; a.ny = L.ny;
    // warp regrouping thresholds; tunable for experiments through the environment
    static const int defer_events = env_int("ARTES_DEFER_EVENTS", 12), defer_refill = env_int("ARTES_DEFER_REFILL", 4);
    a.defer_events = defer_events; a.defer_refill = defer_refill;
    static const int e2_trips = env_int("ARTES_E2_TRIPS", 0);
    static const int e2_cfg = env_int("ARTES_E2_CFG", 32) & 31;      // block shape, see launch_transport2
    static const int e2_inner = env_int("ARTES_E2_INNER", 0);
    static const int e2_per_sm = env_int("ARTES_E2_PER_SM", 8) & 7;
    a.e2_trips = e2_trips; a.e2_pad = e2_cfg; a.e2_inner = e2_inner; a.e2_pad2 = e2_per_sm;
    a.fstop = L.fstop; a.photon_minimum = L.photon_minimum; a.photon_bias = L.photon_bias;
    a.surface_albedo = L.surface_albedo; a.theta_star = L.theta_star; a.phi_star = L.phi_star;
    a.x_max = L.x_max; a.y_max = L.y_max;
    a.det[0] = 1.0 * std::sin(L.det_theta) * std::cos(L.det_phi);
    a.det[1] = 1.0 * std::sin(L.det_theta) * std::sin(L.det_phi);
    a.det[2] = 1.0 * std::cos(L.det_theta);
    a.sin_dt = std::sin(L.det_theta); a.cos_dt = std::cos(L.det_theta);
    a.sin_dp = std::sin(L.det_phi); a.cos_dp = std::cos(L.det_phi);
    double pn = std::atan2(a.det[1], a.det[0]);
    if (pn < 0.0) pn = pn + 2.0 * PI;
    if (pn > 2.0 * PI) pn = pn - 2.0 * PI;
    a.det_atan2 = pn;
    double r = std::sqrt(a.det[0] * a.det[0] + a.det[1] * a.det[1] + a.det[2] * a.det[2]);
    a.det_sph_theta = std::acos(a.det[2] / r);
    double ph = std::atan2(a.det[1], a.det[0]);
    if (ph < 0.0) ph = ph + 2.0 * PI;
    a.det_sph_phi = ph;
    const double ay = -(PI / 2.0 - L.theta_star);
    a.rot_y_cos = std::cos(ay); a.rot_y_sin = std::sin(ay);
    a.rot_z_cos = std::cos(L.phi_star); a.rot_z_sin = std::sin(L.phi_star);
    double td = PI - L.theta_star, pd = PI + L.phi_star;
    if (td < 0.0) td = td + 2.0 * PI;
    if (td > 2.0 * PI) td = td - 2.0 * PI;
    if (pd < 0.0) pd = pd + 2.0 * PI;
    if (pd > 2.0 * PI) pd = pd - 2.0 * PI;
    a.star_dir[0] = 1.0 * std::sin(td) * std::cos(pd);
    a.star_dir[1] = 1.0 * std::sin(td) * std::sin(pd);
    a.star_dir[2] = 1.0 * std::cos(td);
}

int check_launch(artes_gpu_ctx* ctx, const artes_launch_t* L) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!L || L->struct_size != sizeof(artes_launch_t)) return fail(ctx, -1, "artes_launch_t: struct_size mismatch (ABI)");
    if (!ctx->have_grid || !ctx->have_wl) return fail(ctx, -1, "set_grid / set_wavelength not called");
    if (L->mode != ARTES_MODE_FAITHFUL && L->mode != ARTES_MODE_FAST) return fail(ctx, -1, "unknown mode");
    if (L->photon_source != 1 && L->photon_source != 2) return fail(ctx, -1, "photon_source must be 1 or 2");
    if (L->photon_source == 2 && !ctx->thermal) return fail(ctx, -1, "photon_source=2 needs cell_weight and emis_cdf");
    if (L->nx < 1 || L->ny < 1 || !(L->x_max > 0.0) || !(L->y_max > 0.0)) return fail(ctx, -1, "bad detector geometry");
    if (L->wl_index < 0 || L->wl_index >= ctx->n_wl) return fail(ctx, -1, "wl_index outside the tables of set_wavelength(s)");
    return 0;
}

// device tables of wavelength wl_index out of the stacked set of artes_gpu_set_wavelengths (one place for run, trace, ...)
DevTables tables_of(const artes_gpu_ctx* c, const DeviceState& d, int wl_index) {
    DevTables T = d.T;
    if (wl_index > 0) {
        const size_t o = (size_t)wl_index * c->cells;
        T.kext += o; T.albedo += o; T.c2u += o; T.cellrec += 4 * o;
        if (T.cell_weight) { T.cell_weight += o; T.emis_cdf += o; }
    }
    T.cell_depth = c->wl_depth[wl_index];
    return T;
}

size_t out_doubles(const artes_gpu_ctx* c, const artes_launch_t& L) {
    return (size_t)10 * L.nx * L.ny + 2 + (size_t)7 * c->cells;
}

}  // namespace

extern "C" {

int artes_gpu_abi_version(void) { return ARTES_GPU_ABI_VERSION; }

const char* artes_gpu_last_error(const artes_gpu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int artes_gpu_create(artes_gpu_ctx** out, int ndev, const int* dev_ids) {
    artes_gpu_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, -1, "null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1)
        return fail(nullptr, -2, std::string("no CUDA device (libartes_gpu has no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
    if (ndev < 1) ndev = 1;
    if (ndev > count && !dev_ids) return fail(nullptr, -1, "more devices requested than present");
    ctx = new artes_gpu_ctx();
    ctx->devs.resize(ndev);
    for (int i = 0; i < ndev; ++i) {
        DeviceState& d = ctx->devs[i];
        d.dev = dev_ids ? dev_ids[i] : i;
        cudaDeviceProp prop;
        if ((e = cudaSetDevice(d.dev)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, d.dev)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess) {
            std::string msg = std::string("device init: ") + cudaGetErrorString(e);
            delete ctx;
            return fail(nullptr, -2, msg);
        }
        d.sm_count = prop.multiProcessorCount;
        for (auto& ev : d.ev) cudaEventCreate(&ev);
    }
    if (ndev > 1) {  // single-process multi-GPU: one communicator per device
        if (!g_nccl.load()) { std::string m = g_nccl.err; delete ctx; return fail(nullptr, -4, m); }
        std::vector<ncclComm_t> comms(ndev);
        std::vector<int> ids(ndev);
        for (int i = 0; i < ndev; ++i) ids[i] = ctx->devs[i].dev;
        ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, ids.data());
        if (r != ncclSuccess) { std::string m = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r); delete ctx; return fail(nullptr, -4, m); }
        for (int i = 0; i < ndev; ++i) ctx->devs[i].comm = comms[i];
    }
    *out = ctx;
    return 0;
}

int artes_gpu_destroy(artes_gpu_ctx* ctx) {
    if (!ctx) return 0;
    for (auto& d : ctx->devs) {
        cudaSetDevice(d.dev);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.comm && g_nccl.ok) g_nccl.CommDestroy(d.comm);
        free_pool(d.grid_allocs);
        free_pool(d.wl_allocs); d.wl_caps.clear(); d.wl_next = 0;
        if (d.out_d) cudaFree(d.out_d);
        if (d.out_u) cudaFree(d.out_u);
        if (d.scratch) cudaFree(d.scratch);
        if (d.geo) cudaFree(d.geo);
        for (auto& ev : d.ev) if (ev) cudaEventDestroy(ev);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
    return 0;
}

int artes_gpu_device_info(const artes_gpu_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    if (!ctx || ctx->devs.empty()) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->devs[0].dev) != cudaSuccess) return -2;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) { std::strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return 0;
}

int artes_gpu_set_grid(artes_gpu_ctx* ctx, int nr, int ntheta, int nphi, const double* rfront, const double* thetafront,
                       const int32_t* thetaplane, const double* phifront, double ox, double oy, double oz) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (nr < 1 || ntheta < 1 || nphi < 1 || !rfront || !thetafront || !thetaplane || !phifront) return fail(ctx, -1, "bad grid");
    ctx->nr = nr; ctx->nt = ntheta; ctx->np = nphi; ctx->cells = nr * ntheta * nphi;
    std::vector<double> tcos(ntheta + 1), ttan(ntheta + 1), pcos(nphi), psin(nphi), trig(540);
    for (int i = 0; i <= ntheta; ++i) { tcos[i] = std::cos(thetafront[i]); ttan[i] = std::tan(thetafront[i]); }  // :2261-2264
    for (int i = 0; i < nphi; ++i) { pcos[i] = std::cos(phifront[i]); psin[i] = std::sin(phifront[i]); }          // :2267-2270
    ctx->sinbeta.resize(180); ctx->cos2beta.resize(180); ctx->sin2beta.resize(180);
    for (int i = 1; i <= 180; ++i) {  // :409-420
        ctx->sinbeta[i - 1] = (std::sin((double)i * PI / 180.0) + std::sin((double)(i - 1) * PI / 180.0)) / 2.0;
        ctx->cos2beta[i - 1] = (std::cos(2.0 * (double)i * PI / 180.0) + std::cos(2.0 * (double)(i - 1) * PI / 180.0)) / 2.0;
        ctx->sin2beta[i - 1] = (std::sin(2.0 * (double)i * PI / 180.0) + std::sin(2.0 * (double)(i - 1) * PI / 180.0)) / 2.0;
        trig[i - 1] = ctx->sinbeta[i - 1]; trig[180 + i - 1] = ctx->cos2beta[i - 1]; trig[360 + i - 1] = ctx->sin2beta[i - 1];
    }
    std::vector<double> cdfA(2 * 181, 0.0);  // fast-mode prefix sums of cos2beta / sin2beta
    for (int i = 1; i <= 180; ++i) { cdfA[i] = cdfA[i - 1] + ctx->cos2beta[i - 1]; cdfA[181 + i] = cdfA[181 + i - 1] + ctx->sin2beta[i - 1]; }
    std::vector<double> cdfA2(2 * 181);
    for (int i = 0; i <= 180; ++i) { cdfA2[2 * i] = cdfA[i]; cdfA2[2 * i + 1] = cdfA[181 + i]; }
    std::vector<int> tp(thetaplane, thetaplane + ntheta + 1);
    for (auto& d : ctx->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaStreamSynchronize(d.stream));
        free_pool(d.grid_allocs);
        DevTables& T = d.T;
        T.nr = nr; T.nt = ntheta; T.np = nphi; T.cells = ctx->cells; T.ox = ox; T.oy = oy; T.oz = oz;
        T.inv_ox = 1.0 / ox; T.inv_oy = 1.0 / oy; T.inv_oz = 1.0 / oz;
        int rc = 0;
        rc |= upload(ctx, d, d.grid_allocs, rfront, (size_t)nr + 1, &T.rfront);
        rc |= upload(ctx, d, d.grid_allocs, thetafront, (size_t)ntheta + 1, &T.thetafront);
        rc |= upload(ctx, d, d.grid_allocs, ttan.data(), ttan.size(), &T.ttan);
        rc |= upload(ctx, d, d.grid_allocs, tcos.data(), tcos.size(), &T.tcos);
        rc |= upload(ctx, d, d.grid_allocs, tp.data(), tp.size(), &T.tplane);
        rc |= upload(ctx, d, d.grid_allocs, phifront, (size_t)nphi, &T.phifront);
        rc |= upload(ctx, d, d.grid_allocs, psin.data(), psin.size(), &T.psin);
        rc |= upload(ctx, d, d.grid_allocs, pcos.data(), pcos.size(), &T.pcos);
        rc |= upload(ctx, d, d.grid_allocs, trig.data(), trig.size(), &T.trig);
        rc |= upload(ctx, d, d.grid_allocs, cdfA.data(), cdfA.size(), &T.cdfA);
        rc |= upload(ctx, d, d.grid_allocs, cdfA2.data(), cdfA2.size(), &T.cdfA2);
        if (rc) return rc;
        CU(cudaStreamSynchronize(d.stream));
    }
    ctx->have_grid = true;
    ctx->have_wl = false;
    return 0;
}

}  // extern "C"

namespace {
bool g_full_matrix = false;      // test hook, see artes_gpu_test_full_matrix
// tables of n_wl wavelengths: the per-cell arrays are [n_wl][cells], cell_to_uniq indexes one common list of matrices
int set_wavelength_tables(artes_gpu_ctx* ctx, int n_wl, const double* k_sca, const double* k_abs, int n_uniq,
                          const double* uniq_matrix, const int32_t* cell_to_uniq, const int* cell_depths,
                          const double* cell_weight, const double* emis_cdf) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!ctx->have_grid) return fail(ctx, -1, "set_grid first");
    if (ctx->pending) return fail(ctx, -1, "a launch is pending (call artes_gpu_wait before changing the tables)");
    ctx->have_wl = false;      // a failed upload leaves the context without wavelength tables (a later run is refused) instead of half-updated ones
    if (!k_sca || !k_abs || !uniq_matrix || !cell_to_uniq || n_uniq < 1 || n_wl < 1 || !cell_depths) return fail(ctx, -1, "bad wavelength tables");
    for (int l = 0; l < n_wl; ++l) if (cell_depths[l] < 0 || cell_depths[l] >= ctx->nr) return fail(ctx, -1, "cell_depth out of range");
    const int cell_depth = cell_depths[0];
    if ((long long)ctx->cells * n_wl > 2000000000LL) return fail(ctx, -1, "too many cells x wavelengths for one table set");
    const int n = ctx->cells * n_wl;
    for (int i = 0; i < n; ++i) if (cell_to_uniq[i] < 0 || cell_to_uniq[i] >= n_uniq) return fail(ctx, -1, "cell_to_uniq out of range");
    // cell_opacity, cell_albedo (:2178-2188) and the per-cell record are derived ON THE DEVICE from the uploaded k_sca, k_abs and
    // cell_to_uniq (ingest.cu): a third of the host-to-device bytes and no host pass over the cells
    std::vector<double> p1k((size_t)n_uniq * 4, 0.0), mrow((size_t)n_uniq * 720), cdfP((size_t)n_uniq * 181 * 4, 0.0);
    for (int u = 0; u < n_uniq; ++u) {
        for (int a = 0; a < 180; ++a)
            for (int k = 0; k < 4; ++k) {
                const double m = uniq_matrix[((size_t)u * 180 + a) * 16 + k];
                mrow[((size_t)u * 180 + a) * 4 + k] = m;
                p1k[(size_t)u * 4 + k] = p1k[(size_t)u * 4 + k] + m * ctx->sinbeta[a] * PI / 180.0;  // :2221-2224
                cdfP[((size_t)u * 181 + a + 1) * 4 + k] = cdfP[((size_t)u * 181 + a) * 4 + k] + m * ctx->sinbeta[a] * PI / 180.0;
            }
    }
    // Block-diagonal scattering matrices (F13 = F14 = F23 = F24 = F31 = F32 = F41 = F42 = 0 exactly, in every block and at every angle) get a
    // second, half-size copy with the eight elements that can be non-zero: the interaction event then reads 2 x 64 bytes per angle
    // instead of 2 x 128, and products with exact zeros add nothing, so the Stokes vectors are the same.
    std::vector<double> mc;
    {
        static const int zero_at[8] = {2, 3, 6, 7, 8, 9, 12, 13}, keep[8] = {0, 1, 4, 5, 10, 11, 14, 15};
        bool compact = !g_full_matrix;       // (artes_gpu_test_full_matrix: the test hook that forces the 16-element path)
        for (size_t r = 0; compact && r < (size_t)n_uniq * 180; ++r)
            for (int k = 0; k < 8; ++k) if (uniq_matrix[r * 16 + zero_at[k]] != 0.0) { compact = false; break; }
        if (compact) {
            mc.resize((size_t)n_uniq * 180 * 8);
            for (size_t r = 0; r < (size_t)n_uniq * 180; ++r)
                for (int k = 0; k < 8; ++k) mc[r * 8 + k] = uniq_matrix[r * 16 + keep[k]];
        }
    }
    std::vector<double> cdf_lin;
    ctx->thermal = (cell_weight && emis_cdf);
    if (ctx->thermal) {  // reorder emissivity_cumulative into its (i,j,k) construction order (:2425-2427); wavelength l at l * cells
        const int nr = ctx->nr, nt = ctx->nt, np = ctx->np;
        cdf_lin.assign((size_t)n, 0.0);
        for (int l = 0; l < n_wl; ++l) {
            size_t p = (size_t)l * ctx->cells;
            const double* src = emis_cdf + (size_t)l * ctx->cells;
            for (int i = cell_depths[l]; i < nr; ++i)
                for (int j = 0; j < nt; ++j)
                    for (int k = 0; k < np; ++k) cdf_lin[p++] = src[i + nr * (j + nt * k)];
        }
    }
    for (auto& d : ctx->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaStreamSynchronize(d.stream));
        d.wl_next = 0;
        DevTables& T = d.T;
        T.cell_depth = cell_depth; T.n_uniq = n_uniq;
        const cudaEvent_t e0 = d.ev[0], e1 = d.ev[1];      // the device's own timing events (no launch is pending: the stream was just drained)
        CU(cudaEventRecord(e0, d.stream));
        int rc = 0;
        const double *d_sca = nullptr, *d_abs = nullptr;
        rc |= upload_wl(ctx, d, (const double*)nullptr, (size_t)n, &T.kext);
        rc |= upload_wl(ctx, d, (const double*)nullptr, (size_t)4 * n, &T.cellrec);
        rc |= upload_wl(ctx, d, (const double*)nullptr, (size_t)n, &T.albedo);
        rc |= upload_wl(ctx, d, cell_to_uniq, (size_t)n, &T.c2u);
        rc |= upload_wl(ctx, d, k_sca, (size_t)n, &d_sca);
        rc |= upload_wl(ctx, d, k_abs, (size_t)n, &d_abs);
        if (rc) return rc;
        {
            cudaError_t ce = ingest_cell_tables(d_sca, d_abs, T.c2u, (size_t)n, const_cast<double*>(T.kext), const_cast<double*>(T.albedo),
                                                const_cast<double*>(T.cellrec), d.stream);
            if (ce != cudaSuccess) return fail(ctx, -2, std::string("cell tables: ") + cudaGetErrorString(ce));
        }
        rc |= upload_wl(ctx, d, uniq_matrix, (size_t)n_uniq * 2880, &T.M);
        rc |= upload_wl(ctx, d, mrow.data(), mrow.size(), &T.Mrow);
        rc |= upload_wl(ctx, d, p1k.data(), p1k.size(), &T.p1k);
        rc |= upload_wl(ctx, d, cdfP.data(), cdfP.size(), &T.cdfP);
        rc |= upload_wl(ctx, d, mc.data(), mc.size(), &T.Mc);      // (always takes its pool slot, so that the slots of the tables after it do not move)
        if (mc.empty()) T.Mc = nullptr;
        T.cell_weight = nullptr; T.emis_cdf = nullptr;
        if (ctx->thermal) {
            rc |= upload_wl(ctx, d, cell_weight, (size_t)n, &T.cell_weight);
            rc |= upload_wl(ctx, d, cdf_lin.data(), cdf_lin.size(), &T.emis_cdf);
        }
        if (rc) return rc;
        CU(cudaEventRecord(e1, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        ctx->last_h2d_ms = ms;
    }
    ctx->have_wl = true;
    ctx->n_wl = n_wl;
    ctx->wl_depth.assign(cell_depths, cell_depths + n_wl);
    return 0;
}
}  // namespace

extern "C" {

int artes_gpu_set_wavelength(artes_gpu_ctx* ctx, const double* k_sca, const double* k_abs, int n_uniq,
                             const double* uniq_matrix, const int32_t* cell_to_uniq, int cell_depth,
                             const double* cell_weight, const double* emis_cdf) {
    return set_wavelength_tables(ctx, 1, k_sca, k_abs, n_uniq, uniq_matrix, cell_to_uniq, &cell_depth, cell_weight, emis_cdf);
}

int artes_gpu_set_wavelengths(artes_gpu_ctx* ctx, int n_wl, const double* k_sca, const double* k_abs, int n_uniq,
                              const double* uniq_matrix, const int32_t* cell_to_uniq, const int32_t* cell_depths,
                              const double* cell_weight, const double* emis_cdf) {
    return set_wavelength_tables(ctx, n_wl, k_sca, k_abs, n_uniq, uniq_matrix, cell_to_uniq, cell_depths, cell_weight, emis_cdf);
}

// host de-duplication of the per-cell 180x16 blocks of one wavelength; element (cell, e, a) at cell + plane_stride*(e + 16*a)
static int dedup_host(artes_gpu_ctx* ctx, const double* dense, size_t plane_stride, std::vector<double>& uniq, std::vector<int32_t>& c2u) {
    // python/atmosphere.py:351-372 mixes a handful of species, so the dense HDU (cells x 23 040 B) holds few distinct blocks.
    const size_t n = (size_t)ctx->cells;
    std::vector<double> block(2880);
    c2u.assign(n, 0);
    uniq.clear();
    std::unordered_multimap<uint64_t, int> seen;
    for (size_t cidx = 0; cidx < n; ++cidx) {
        uint64_t h = 1469598103934665603ull;
        for (int a = 0; a < 180; ++a)
            for (int e = 0; e < 16; ++e) {
                double v = dense[cidx + plane_stride * ((size_t)e + 16 * (size_t)a)];
                block[a * 16 + e] = v;
                uint64_t bits; std::memcpy(&bits, &v, 8);
                h ^= bits; h *= 1099511628211ull;
            }
        int found = -1;
        auto range = seen.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (std::memcmp(&uniq[(size_t)it->second * 2880], block.data(), 2880 * 8) == 0) { found = it->second; break; }
        if (found < 0) {
            found = (int)(uniq.size() / 2880);
            uniq.insert(uniq.end(), block.begin(), block.end());
            seen.emplace(h, found);
        }
        c2u[cidx] = found;
    }
    return 0;
}

// the same on the device (ingest.cu); falls back to the host path if a hash group fails the element-wise check
static int dedup_device(artes_gpu_ctx* ctx, const double* dense, size_t plane_stride, std::vector<double>& uniq, std::vector<int32_t>& c2u) {
    DeviceState& d = ctx->devs[0];
    CU(cudaSetDevice(d.dev));
    int exact = 1;
    cudaError_t e = ingest_dedup_device(dense, (size_t)ctx->cells, plane_stride, d.stream, uniq, c2u, &exact, nullptr, nullptr);
    if (e != cudaSuccess) return fail(ctx, -2, std::string("matrix ingest: ") + cudaGetErrorString(e));
    if (!exact) return dedup_host(ctx, dense, plane_stride, uniq, c2u);
    return 0;
}

int artes_gpu_set_wavelength_dense(artes_gpu_ctx* ctx, const double* k_sca, const double* k_abs, const double* dense,
                                   int cell_depth, const double* cell_weight, const double* emis_cdf) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!ctx->have_grid) return fail(ctx, -1, "set_grid first");
    if (!dense) return fail(ctx, -1, "null matrix");
    std::vector<double> uniq;
    std::vector<int32_t> c2u;
    const int rc = dedup_device(ctx, dense, (size_t)ctx->cells, uniq, c2u);
    if (rc) return rc;
    return artes_gpu_set_wavelength(ctx, k_sca, k_abs, (int)(uniq.size() / 2880), uniq.data(), c2u.data(), cell_depth, cell_weight, emis_cdf);
}

int artes_gpu_set_wavelength_dense_wl(artes_gpu_ctx* ctx, int n_wl, int wl_index, const double* k_sca_all, const double* k_abs_all,
                                      const double* matrix_all, int cell_depth, const double* cell_weight, const double* emis_cdf) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!ctx->have_grid) return fail(ctx, -1, "set_grid first");
    if (!k_sca_all || !k_abs_all || !matrix_all || n_wl < 1 || wl_index < 0 || wl_index >= n_wl) return fail(ctx, -1, "bad dense wavelength arguments");
    const size_t cells = (size_t)ctx->cells;
    std::vector<double> uniq;
    std::vector<int32_t> c2u;
    int rc = dedup_device(ctx, matrix_all + cells * (size_t)wl_index, cells * (size_t)n_wl, uniq, c2u);
    if (rc) return rc;
    return artes_gpu_set_wavelength(ctx, k_sca_all + cells * (size_t)wl_index, k_abs_all + cells * (size_t)wl_index, (int)(uniq.size() / 2880),
                                    uniq.data(), c2u.data(), cell_depth, cell_weight, emis_cdf);
}

int artes_gpu_run_async(artes_gpu_ctx* ctx, const artes_launch_t* L) {
    int rc = check_launch(ctx, L);
    if (rc) return rc;
    if (ctx->pending) return fail(ctx, -1, "a launch is already pending (call artes_gpu_wait)");
    const int ndev = (int)ctx->devs.size();
    const size_t npx = (size_t)L->nx * L->ny;
    const size_t n_d = out_doubles(ctx, *L);
    ctx->n_out_d = n_d;
    // contiguous photon-id ranges: first over ranks (done by the caller through photon_id_base), then over devices
    unsigned long long per = L->n_photons / ndev, rem = L->n_photons % ndev, off = 0;
    for (int i = 0; i < ndev; ++i) {
        DeviceState& d = ctx->devs[i];
        CU(cudaSetDevice(d.dev));
        rc = ensure_outputs(ctx, d, n_d);
        if (rc) return rc;
        CU(cudaMemsetAsync(d.out_d, 0, n_d * sizeof(double), d.stream));
        CU(cudaMemsetAsync(d.out_u, 0, (ARTES_ERR_SLOTS + 16) * sizeof(unsigned long long), d.stream));
        KernelArgs a{};
        a.T = tables_of(ctx, d, L->wl_index);
        fill_launch(*L, a.L);
        a.L.n_photons = per + ((unsigned long long)i < rem ? 1 : 0);
        a.L.id_base = L->photon_id_base + off;
        off += a.L.n_photons;
        d.n_photons = a.L.n_photons;
        a.O.det = d.out_d;
        a.O.flux = d.out_d + 10 * npx;
        a.O.flow4 = d.out_d + 10 * npx + 2;
        a.O.flow3 = d.out_d + 10 * npx + 2 + (size_t)4 * ctx->cells;
        a.O.err = d.out_u;
        a.O.stats = d.out_u + ARTES_ERR_SLOTS;
        a.O.counter = d.out_u + ARTES_ERR_SLOTS + 8;
        CU(cudaEventRecord(d.ev[0], d.stream));
        d.launches = 0;
        if (a.L.n_photons > 0) {
            // fast mode: the ray/event engine wherever it applies (ARTES_ENGINE=1 forces the persistent-lane engine)
            static const int engine = env_int("ARTES_ENGINE", 2);
            cudaError_t e;
            ctx->last_engine = 1;
            if (engine != 1 && !(engine == 2 && L->mode == ARTES_MODE_FAST && fast::engine2_supports(a))) {
                // the event-list engine around the reference-order event bodies (engine3.cuh): the faithful mode, and what the
                // ray/event engine leaves out in fast mode (flow_global, oblate planets)
                ctx->last_engine = 3;
                if (!d.scratch) CU(cudaMalloc(&d.scratch, scratch_bytes(d.sm_count)));
                a.O.scratch = d.scratch;
                e = (L->mode == ARTES_MODE_FAITHFUL) ? faithful::launch_transport4(a, false, d.sm_count, d.stream)
                                                     : fast::launch_transport4(a, false, d.sm_count, d.stream);
            }
            else if (L->mode == ARTES_MODE_FAITHFUL) e = faithful::launch_transport(a, false, d.sm_count, d.stream);
            else if (engine == 2 && fast::engine2_supports(a)) {
                ctx->last_engine = 2;
                if (!d.scratch) CU(cudaMalloc(&d.scratch, scratch_bytes(d.sm_count)));
                a.O.scratch = d.scratch;
                e = fast::launch_transport2(a, d.sm_count, d.stream);
            }
            else e = fast::launch_transport(a, false, d.sm_count, d.stream);
            if (e != cudaSuccess) return fail(ctx, -2, std::string("transport launch: ") + cudaGetErrorString(e));
            d.launches = 1;
        }
        CU(cudaEventRecord(d.ev[1], d.stream));
    }
    // NCCL sum of the packed accumulators (the thread sum of :959-975 across devices / ranks)
    const bool reduce = (ndev > 1) || ctx->rank_comm;
    if (reduce) {
        NC(g_nccl.GroupStart());
        for (auto& d : ctx->devs) {
            NC(g_nccl.AllReduce(d.out_d, d.out_d, n_d, ncclDouble, ncclSum, d.comm, d.stream));
            NC(g_nccl.AllReduce(d.out_u, d.out_u, ARTES_ERR_SLOTS + 8, ncclUint64, ncclSum, d.comm, d.stream));
        }
        NC(g_nccl.GroupEnd());
    }
    for (auto& d : ctx->devs) { CU(cudaSetDevice(d.dev)); CU(cudaEventRecord(d.ev[2], d.stream)); }
    ctx->pending = true;
    ctx->pending_launch = *L;
    return 0;
}

int artes_gpu_wait(artes_gpu_ctx* ctx, double* det_sum, double* flux, double* flow4, double* flow3,
                   uint64_t* err_hist, artes_stats_t* stats) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!ctx->pending) return fail(ctx, -1, "no pending launch");
    ctx->pending = false;
    const artes_launch_t& L = ctx->pending_launch;
    const size_t npx = (size_t)L.nx * L.ny;
    const size_t n_d = ctx->n_out_d;
    double kernel_ms = 0.0, reduce_ms = 0.0;
    for (auto& d : ctx->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaStreamSynchronize(d.stream));
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, d.ev[0], d.ev[1]);
        cudaEventElapsedTime(&b, d.ev[1], d.ev[2]);
        kernel_ms = std::max(kernel_ms, (double)a);
        reduce_ms = std::max(reduce_ms, (double)b);
    }
    // every device holds the reduced result; read device 0
    DeviceState& d0 = ctx->devs[0];
    CU(cudaSetDevice(d0.dev));
    std::vector<double> h(n_d);
    std::vector<unsigned long long> hu(ARTES_ERR_SLOTS + 8);
    cudaEvent_t e0 = d0.ev[3];
    CU(cudaEventRecord(e0, d0.stream));
    CU(cudaMemcpyAsync(h.data(), d0.out_d, n_d * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    CU(cudaMemcpyAsync(hu.data(), d0.out_u, hu.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d0.stream));
    CU(cudaEventRecord(d0.ev[2], d0.stream));
    CU(cudaStreamSynchronize(d0.stream));
    float d2h = 0.f;
    cudaEventElapsedTime(&d2h, e0, d0.ev[2]);
    if (det_sum) {  // expand to detector(nx,ny,4,3): the count plane is shared by Q,U,V (:4969-4972)
        std::memcpy(det_sum, h.data(), 8 * npx * sizeof(double));
        std::memcpy(det_sum + 8 * npx, h.data() + 8 * npx, npx * sizeof(double));
        for (int k = 1; k < 4; ++k) std::memcpy(det_sum + (8 + k) * npx, h.data() + 9 * npx, npx * sizeof(double));
    }
    if (flux) { flux[0] = h[10 * npx]; flux[1] = h[10 * npx + 1]; }
    if (flow4) std::memcpy(flow4, h.data() + 10 * npx + 2, (size_t)4 * ctx->cells * sizeof(double));
    if (flow3) std::memcpy(flow3, h.data() + 10 * npx + 2 + (size_t)4 * ctx->cells, (size_t)3 * ctx->cells * sizeof(double));
    if (hu[ARTES_ERR_SLOTS - 1]) return fail(ctx, -5, "transport kernel stopped by its watchdog (no scheduling progress): results discarded");
    if (err_hist) for (int k = 0; k < ARTES_ERR_SLOTS; ++k) err_hist[k] = hu[k];
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        const unsigned long long* s = hu.data() + ARTES_ERR_SLOTS;
        stats->n_emit = s[0]; stats->n_cell_face = s[1]; stats->n_scatter = s[2]; stats->n_peel = s[3];
        stats->n_surface = s[4]; stats->n_draws = s[5]; stats->n_error = s[6];
        unsigned long long launches = 0;
        for (auto& d : ctx->devs) launches += (unsigned long long)d.launches;
        stats->reserved = launches;   // kernels launched by this call (all devices)
        stats->kernel_ms = kernel_ms; stats->reduce_ms = reduce_ms; stats->h2d_ms = ctx->last_h2d_ms; stats->d2h_ms = d2h;
    }
    return 0;
}

int artes_gpu_run(artes_gpu_ctx* ctx, const artes_launch_t* L, double* det_sum, double* flux, double* flow4,
                  double* flow3, uint64_t* err_hist, artes_stats_t* stats) {
    int rc = artes_gpu_run_async(ctx, L);
    if (rc) return rc;
    return artes_gpu_wait(ctx, det_sum, flux, flow4, flow3, err_hist, stats);
}

// Several launches that differ only in the detector direction, as ONE kernel launch (the phase-curve loop
// src/ARTES.f90:215-245 calls radiative_transfer 73 times with a new det_phi).  A transport launch ends with a drain
// phase in which the last photons of every block finish one scattering after the other while the SMs idle; at the
// reference's 1e6 packets per launch that is ~40 % of the launch on B200.  Batched, the work items of all launches
// come out of one counter, every photon carries the index of its launch, and the drain is paid once per batch.
// Launch k uses the photon ids photon_id_base(launch 0) + k * n_photons + [0, n_photons): its result equals the single
// launch with that photon_id_base up to the order of the floating-point sums.
static int run_batch_impl(artes_gpu_ctx* ctx, const artes_launch_t* Ls, int n, double* det_sum, double* flux,
                          uint64_t* err_hist, artes_stats_t* stats, bool multi) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!Ls || n < 1 || n > ARTES_MAX_BATCH) return fail(ctx, -1, "run_batch: 1 <= n <= ARTES_MAX_BATCH launches");
    if (ctx->pending) return fail(ctx, -1, "a launch is already pending (call artes_gpu_wait)");
    for (int k = 0; k < n; ++k) {
        int rc = check_launch(ctx, Ls + k);
        if (rc) return rc;
        artes_launch_t a = Ls[k], b = Ls[0];
        a.det_theta = b.det_theta; a.det_phi = b.det_phi; a.limb_emission = b.limb_emission; a.wl_index = b.wl_index;
        if (std::memcmp(&a, &b, sizeof(a)) != 0) return fail(ctx, -1, "run_batch: launches may differ in det_theta, det_phi, limb_emission and wl_index only");
    }
    const artes_launch_t& L0 = Ls[0];
    if (L0.flow_global || L0.flow_theta) return fail(ctx, -1, "run_batch: flow counters are per launch; use artes_gpu_run");
    const size_t npx = (size_t)L0.nx * L0.ny;
    const unsigned long long P = L0.n_photons;
    static const int engine = env_int("ARTES_ENGINE", 2);
    bool batched = (n > 1 || multi) && P > 0 && L0.mode == ARTES_MODE_FAST && engine == 2;
    int sdet_doubles = 0;
    if (batched) {
        KernelArgs a{};
        a.T = ctx->devs[0].T;
        fill_launch(L0, a.L);
        // a launch table too large for shared memory or an image stack whose pixel offsets leave 32 bits: one launch after the other
        batched = fast::engine2_supports(a) && fast::engine2_batch_fits(a, n);
        // block-private detector images in shared memory: multi-detector walks only.  Measured (profiles/r02_e_*): behind the
        // warp-aggregated deposits of the single-detector kernels a private image gains nothing (1 x 1 detectors: -0.3 .. -1.7 %)
        // and a 50 KB image (25 x 25) costs 18 % through the L1 it takes away.
        if (multi) sdet_doubles = fast::engine2_sdet_doubles(a, n);
    }
    if (multi) {
        // ONE walk for all detectors needs one emission law and one set of tables: same limb flag and wavelength everywhere,
        // star source over a black surface (the thermal / surface peel-offs keep their per-detector launches)
        bool same = true;
        for (int k = 1; k < n; ++k) same = same && Ls[k].limb_emission == L0.limb_emission && Ls[k].wl_index == L0.wl_index;
        if (!same) return fail(ctx, -1, "run_multi: the detectors of one walk must share limb_emission and wl_index");
        if (!(batched && L0.photon_source == 1 && !(L0.surface_albedo > 0.0)))
            return run_batch_impl(ctx, Ls, n, det_sum, flux, err_hist, stats, false);      // independent walks: same expectation values
    }
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (err_hist) std::memset(err_hist, 0, ARTES_ERR_SLOTS * sizeof(uint64_t));
    if (!batched) {   // one launch after the other: faithful mode, oblate planets, the persistent-lane engine
        for (int k = 0; k < n; ++k) {
            artes_launch_t Lk = Ls[k];
            Lk.photon_id_base = L0.photon_id_base + (unsigned long long)k * P;
            uint64_t eh[ARTES_ERR_SLOTS];
            artes_stats_t st;
            int rc = artes_gpu_run(ctx, &Lk, det_sum ? det_sum + (size_t)k * 12 * npx : nullptr, flux ? flux + 2 * k : nullptr, nullptr, nullptr, eh, &st);
            if (rc) return rc;
            if (err_hist) for (int c = 0; c < ARTES_ERR_SLOTS; ++c) err_hist[c] += eh[c];
            if (stats) {
                stats->n_emit += st.n_emit; stats->n_cell_face += st.n_cell_face; stats->n_scatter += st.n_scatter; stats->n_peel += st.n_peel;
                stats->n_surface += st.n_surface; stats->n_draws += st.n_draws; stats->n_error += st.n_error; stats->reserved += st.reserved;
                stats->kernel_ms += st.kernel_ms; stats->reduce_ms += st.reduce_ms; stats->d2h_ms += st.d2h_ms; stats->h2d_ms = st.h2d_ms;
            }
        }
        return 0;
    }
    const int ndev = (int)ctx->devs.size();
    const size_t n_d = (size_t)n * (10 * npx + 2);
    constexpr int GEO = 12;   // = e2::GEO (engine2.cuh): det(3), sin_dt, cos_dt, sin_dp, cos_dp, limb_emission, det_sph_theta, det_sph_phi, cell_depth, wavelength index
    bool wl_batch = false;
    for (int k = 0; k < n; ++k) wl_batch = wl_batch || Ls[k].wl_index != 0;
    std::vector<double> geo((size_t)n * GEO);
    for (int k = 0; k < n; ++k) {
        LaunchArgs t{};
        fill_launch(Ls[k], t);
        double* g = geo.data() + (size_t)k * GEO;
        g[0] = t.det[0]; g[1] = t.det[1]; g[2] = t.det[2]; g[3] = t.sin_dt; g[4] = t.cos_dt; g[5] = t.sin_dp; g[6] = t.cos_dp;
        g[7] = Ls[k].limb_emission ? 1.0 : 0.0; g[8] = t.det_sph_theta; g[9] = t.det_sph_phi;
        g[10] = (double)ctx->wl_depth[Ls[k].wl_index]; g[11] = (double)Ls[k].wl_index;
    }
    const unsigned long long total = multi ? P : P * (unsigned long long)n;
    unsigned long long per = total / ndev, rem = total % ndev, off = 0;
    for (int i = 0; i < ndev; ++i) {
        DeviceState& d = ctx->devs[i];
        CU(cudaSetDevice(d.dev));
        int rc = ensure_outputs(ctx, d, n_d);
        if (rc) return rc;
        if (d.geo_cap < geo.size()) {
            if (d.geo) cudaFree(d.geo);
            d.geo = nullptr; d.geo_cap = 0;
            CU(cudaMalloc(&d.geo, geo.size() * sizeof(double)));
            d.geo_cap = geo.size();
        }
        CU(cudaMemcpyAsync(d.geo, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemsetAsync(d.out_d, 0, n_d * sizeof(double), d.stream));
        CU(cudaMemsetAsync(d.out_u, 0, (ARTES_ERR_SLOTS + 16) * sizeof(unsigned long long), d.stream));
        KernelArgs a{};
        a.T = d.T;
        fill_launch(L0, a.L);
        a.L.n_photons = per + ((unsigned long long)i < rem ? 1 : 0);
        a.L.id_base = L0.photon_id_base + off;
        a.L.n_batch = n; a.L.per_launch = P; a.L.batch_base = L0.photon_id_base; a.L.geo = d.geo; a.L.wl_batch = (wl_batch && !multi) ? 1 : 0;
        a.L.multi = multi ? 1 : 0; a.L.sdet_doubles = sdet_doubles;
        if (multi) a.T = tables_of(ctx, d, L0.wl_index);
        off += a.L.n_photons;
        d.n_photons = a.L.n_photons;
        a.O.det = d.out_d;
        a.O.flux = d.out_d + (size_t)n * 10 * npx;
        a.O.flow4 = nullptr; a.O.flow3 = nullptr;
        a.O.err = d.out_u;
        a.O.stats = d.out_u + ARTES_ERR_SLOTS;
        a.O.counter = d.out_u + ARTES_ERR_SLOTS + 8;
        CU(cudaEventRecord(d.ev[0], d.stream));
        d.launches = 0;
        if (a.L.n_photons > 0) {
            if (!d.scratch) CU(cudaMalloc(&d.scratch, scratch_bytes(d.sm_count)));
            a.O.scratch = d.scratch;
            cudaError_t e = fast::launch_transport2(a, d.sm_count, d.stream);
            if (e != cudaSuccess) return fail(ctx, -2, std::string("transport launch: ") + cudaGetErrorString(e));
            d.launches = 1;
        }
        CU(cudaEventRecord(d.ev[1], d.stream));
    }
    ctx->last_engine = 2;
    if ((ndev > 1) || ctx->rank_comm) {
        NC(g_nccl.GroupStart());
        for (auto& d : ctx->devs) {
            NC(g_nccl.AllReduce(d.out_d, d.out_d, n_d, ncclDouble, ncclSum, d.comm, d.stream));
            NC(g_nccl.AllReduce(d.out_u, d.out_u, ARTES_ERR_SLOTS + 8, ncclUint64, ncclSum, d.comm, d.stream));
        }
        NC(g_nccl.GroupEnd());
    }
    double kernel_ms = 0.0, reduce_ms = 0.0;
    for (auto& d : ctx->devs) {
        CU(cudaSetDevice(d.dev));
        CU(cudaEventRecord(d.ev[2], d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, d.ev[0], d.ev[1]);
        cudaEventElapsedTime(&b, d.ev[1], d.ev[2]);
        kernel_ms = std::max(kernel_ms, (double)a);
        reduce_ms = std::max(reduce_ms, (double)b);
    }
    DeviceState& d0 = ctx->devs[0];
    CU(cudaSetDevice(d0.dev));
    std::vector<double> h(n_d);
    std::vector<unsigned long long> hu(ARTES_ERR_SLOTS + 8);
    CU(cudaEventRecord(d0.ev[3], d0.stream));
    CU(cudaMemcpyAsync(h.data(), d0.out_d, n_d * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    CU(cudaMemcpyAsync(hu.data(), d0.out_u, hu.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d0.stream));
    CU(cudaEventRecord(d0.ev[2], d0.stream));
    CU(cudaStreamSynchronize(d0.stream));
    float d2h = 0.f;
    cudaEventElapsedTime(&d2h, d0.ev[3], d0.ev[2]);
    for (int k = 0; k < n; ++k) {
        const double* hk = h.data() + (size_t)k * 10 * npx;
        if (det_sum) {
            double* o = det_sum + (size_t)k * 12 * npx;
            std::memcpy(o, hk, 9 * npx * sizeof(double));
            for (int q = 1; q < 4; ++q) std::memcpy(o + (8 + q) * npx, hk + 9 * npx, npx * sizeof(double));
        }
        if (flux) { flux[2 * k] = h[(size_t)n * 10 * npx + 2 * k]; flux[2 * k + 1] = h[(size_t)n * 10 * npx + 2 * k + 1]; }
    }
    if (hu[ARTES_ERR_SLOTS - 1]) return fail(ctx, -5, "transport kernel stopped by its watchdog (no scheduling progress): results discarded");
    if (err_hist) for (int c = 0; c < ARTES_ERR_SLOTS; ++c) err_hist[c] = hu[c];
    if (stats) {
        const unsigned long long* u = hu.data() + ARTES_ERR_SLOTS;
        stats->n_emit = u[0]; stats->n_cell_face = u[1]; stats->n_scatter = u[2]; stats->n_peel = u[3];
        stats->n_surface = u[4]; stats->n_draws = u[5]; stats->n_error = u[6];
        unsigned long long launches = 0;
        for (auto& d : ctx->devs) launches += (unsigned long long)d.launches;
        stats->reserved = launches;
        stats->kernel_ms = kernel_ms; stats->reduce_ms = reduce_ms; stats->h2d_ms = ctx->last_h2d_ms; stats->d2h_ms = d2h;
    }
    return 0;
}

int artes_gpu_run_batch(artes_gpu_ctx* ctx, const artes_launch_t* Ls, int n, double* det_sum, double* flux,
                        uint64_t* err_hist, artes_stats_t* stats) {
    return run_batch_impl(ctx, Ls, n, det_sum, flux, err_hist, stats, false);
}

// ONE random walk of n_photons packets observed by all n detectors (see include/artes_gpu.h)
int artes_gpu_run_multi(artes_gpu_ctx* ctx, const artes_launch_t* Ls, int n, double* det_sum, double* flux,
                        uint64_t* err_hist, artes_stats_t* stats) {
    return run_batch_impl(ctx, Ls, n, det_sum, flux, err_hist, stats, true);
}

int artes_gpu_nccl_unique_id(void* id_out) {
    if (!id_out) return -1;
    if (!g_nccl.load()) { g_create_error = g_nccl.err; return -4; }
    static_assert(sizeof(ncclUniqueId) <= ARTES_NCCL_ID_BYTES, "ncclUniqueId larger than ARTES_NCCL_ID_BYTES");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_error = g_nccl.GetErrorString(r); return -4; }
    std::memset(id_out, 0, ARTES_NCCL_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return 0;
}

int artes_gpu_nccl_init_rank(artes_gpu_ctx* ctx, int nranks, int rank, const void* id_bytes) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (ctx->devs.size() != 1) return fail(ctx, -1, "nccl_init_rank needs a single-device context (one process per GPU)");
    if (nranks < 1 || rank < 0 || rank >= nranks || !id_bytes) return fail(ctx, -1, "bad rank arguments");
    if (!g_nccl.load()) return fail(ctx, -4, g_nccl.err);
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    DeviceState& d = ctx->devs[0];
    CU(cudaSetDevice(d.dev));
    NC(g_nccl.CommInitRank(&d.comm, nranks, id, rank));
    ctx->nranks = nranks; ctx->rank = rank; ctx->rank_comm = true;
    return 0;
}

int artes_gpu_trace(artes_gpu_ctx* ctx, const artes_launch_t* L, const double* xi, uint64_t n, int max_draws,
                    int32_t* seq_len, uint64_t* seq_hash, int32_t* seq_head, int max_rec, double* fstate) {
    int rc = check_launch(ctx, L);
    if (rc) return rc;
    if (!xi || !seq_len || !seq_hash || n == 0 || max_draws < 1) return fail(ctx, -1, "bad trace arguments");
    DeviceState& d = ctx->devs[0];
    CU(cudaSetDevice(d.dev));
    const size_t npx = (size_t)L->nx * L->ny;
    const size_t n_d = out_doubles(ctx, *L);
    rc = ensure_outputs(ctx, d, n_d);
    if (rc) return rc;
    CU(cudaMemsetAsync(d.out_d, 0, n_d * sizeof(double), d.stream));
    CU(cudaMemsetAsync(d.out_u, 0, (ARTES_ERR_SLOTS + 16) * sizeof(unsigned long long), d.stream));
    double *d_xi = nullptr, *d_f = nullptr;
    int *d_len = nullptr, *d_head = nullptr;
    unsigned long long* d_hash = nullptr;
    CU(cudaMalloc(&d_xi, (size_t)n * max_draws * sizeof(double)));
    CU(cudaMalloc(&d_len, n * sizeof(int)));
    CU(cudaMalloc(&d_hash, n * sizeof(unsigned long long)));
    if (seq_head && max_rec > 0) { CU(cudaMalloc(&d_head, (size_t)n * max_rec * 5 * sizeof(int))); CU(cudaMemsetAsync(d_head, 0xff, (size_t)n * max_rec * 5 * sizeof(int), d.stream)); }
    if (fstate) CU(cudaMalloc(&d_f, n * 8 * sizeof(double)));
    CU(cudaMemcpyAsync(d_xi, xi, (size_t)n * max_draws * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    KernelArgs a{};
    a.T = tables_of(ctx, d, L->wl_index);      // the trace hook walks the tables of the launch's wavelength, like artes_gpu_run
    fill_launch(*L, a.L);
    a.L.n_photons = n;
    a.O.det = d.out_d; a.O.flux = d.out_d + 10 * npx; a.O.flow4 = d.out_d + 10 * npx + 2;
    a.O.flow3 = d.out_d + 10 * npx + 2 + (size_t)4 * ctx->cells;
    a.O.err = d.out_u; a.O.stats = d.out_u + ARTES_ERR_SLOTS; a.O.counter = d.out_u + ARTES_ERR_SLOTS + 8;
    a.R.xi = d_xi; a.R.max_draws = max_draws; a.R.max_rec = d_head ? max_rec : 0;
    a.R.seq_len = d_len; a.R.seq_hash = d_hash; a.R.seq_head = d_head; a.R.fstate = d_f;
    // fast mode: the ray/event engine wherever it applies, so that the trace hook walks the production path
    static const int engine = env_int("ARTES_ENGINE", 2);
    cudaError_t e;
    ctx->last_engine = 1;
    if (engine != 1 && !(engine == 2 && L->mode == ARTES_MODE_FAST && fast::engine2_supports(a))) {
        ctx->last_engine = 3;
        if (!d.scratch) CU(cudaMalloc(&d.scratch, scratch_bytes(d.sm_count)));
        a.O.scratch = d.scratch;
        e = (L->mode == ARTES_MODE_FAITHFUL) ? faithful::launch_transport4(a, true, d.sm_count, d.stream)
                                             : fast::launch_transport4(a, true, d.sm_count, d.stream);
    }
    else if (L->mode == ARTES_MODE_FAITHFUL) e = faithful::launch_transport(a, true, d.sm_count, d.stream);
    else if (engine == 2 && fast::engine2_supports(a)) {
        ctx->last_engine = 2;
        if (!d.scratch) CU(cudaMalloc(&d.scratch, scratch_bytes(d.sm_count)));
        a.O.scratch = d.scratch;
        e = fast::launch_transport2_trace(a, d.sm_count, d.stream);
    } else e = fast::launch_transport(a, true, d.sm_count, d.stream);
    if (e != cudaSuccess) return fail(ctx, -2, std::string("trace launch: ") + cudaGetErrorString(e));
    CU(cudaMemcpyAsync(seq_len, d_len, n * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaMemcpyAsync(seq_hash, d_hash, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
    if (d_head) CU(cudaMemcpyAsync(seq_head, d_head, (size_t)n * max_rec * 5 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    if (d_f) CU(cudaMemcpyAsync(fstate, d_f, n * 8 * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    cudaFree(d_xi); cudaFree(d_len); cudaFree(d_hash); cudaFree(d_head); cudaFree(d_f);
    return 0;
}

int artes_gpu_cell_face(artes_gpu_ctx* ctx, int mode, uint64_t n, const double* pos, const double* dir,
                        const int32_t* face, const int32_t* cell, int32_t* out_i, double* out_d) {
    if (!ctx) return fail(nullptr, -1, "null context");
    if (!ctx->have_grid || !ctx->have_wl) return fail(ctx, -1, "set_grid / set_wavelength not called");
    if (n == 0) return 0;
    DeviceState& d = ctx->devs[0];
    CU(cudaSetDevice(d.dev));
    double *dp = nullptr, *dd = nullptr, *dod = nullptr;
    int *df = nullptr, *dc = nullptr, *doi = nullptr;
    CU(cudaMalloc(&dp, n * 3 * sizeof(double))); CU(cudaMalloc(&dd, n * 3 * sizeof(double))); CU(cudaMalloc(&dod, n * sizeof(double)));
    CU(cudaMalloc(&df, n * 2 * sizeof(int))); CU(cudaMalloc(&dc, n * 3 * sizeof(int))); CU(cudaMalloc(&doi, n * 7 * sizeof(int)));
    CU(cudaMemcpyAsync(dp, pos, n * 3 * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(dd, dir, n * 3 * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(df, face, n * 2 * sizeof(int), cudaMemcpyHostToDevice, d.stream));
    CU(cudaMemcpyAsync(dc, cell, n * 3 * sizeof(int), cudaMemcpyHostToDevice, d.stream));
    cudaError_t e = (mode == ARTES_MODE_FAITHFUL) ? faithful::launch_cell_face(d.T, n, dp, dd, df, dc, doi, dod, d.stream)
                                                  : fast::launch_cell_face(d.T, n, dp, dd, df, dc, doi, dod, d.stream);
    if (e != cudaSuccess) return fail(ctx, -2, std::string("cell_face launch: ") + cudaGetErrorString(e));
    CU(cudaMemcpyAsync(out_i, doi, n * 7 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaMemcpyAsync(out_d, dod, n * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    cudaFree(dp); cudaFree(dd); cudaFree(dod); cudaFree(df); cudaFree(dc); cudaFree(doi);
    return 0;
}

int artes_gpu_last_engine(const artes_gpu_ctx* ctx) { return ctx ? ctx->last_engine : 0; }

int artes_gpu_test_ingest_chunk(int planes) { ingest_set_chunk_override(planes); return 0; }

int artes_gpu_test_full_matrix(int on) { g_full_matrix = on != 0; return 0; }

int artes_gpu_fma_peak(artes_gpu_ctx* ctx, double* fp64_tflops, double* fp32_tflops) {
    if (!ctx) return fail(nullptr, -1, "null context");
    DeviceState& d = ctx->devs[0];
    CU(cudaSetDevice(d.dev));
    cudaError_t e = fma_peak(fp64_tflops, fp32_tflops, d.sm_count, d.stream);
    if (e != cudaSuccess) return fail(ctx, -2, std::string("fma_peak: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
