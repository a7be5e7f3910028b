// ingest.cu -- device-side de-duplication of the reference's dense scattering-matrix array.
//
// get_atmosphere (src/ARTES.f90:2054-2235) reads cell_scatter_matrix(cells, n_wl, 16, 180) -- 23 040 B per cell and
// wavelength, 16.6 GB per wavelength at the scale configuration -- although python/atmosphere.py:351-372 mixes only a
// handful of species per layer, so that the array holds few distinct 180x16 blocks.  The transport kernels work on the
// de-duplicated table (DevTables::M).  Here the 2880 (element, angle) planes of one wavelength -- each a contiguous run of
// `cells` doubles in the reference's layout -- are streamed to HBM, every cell's block is hashed by one thread (plane by
// plane: the loads of a warp are 32 consecutive doubles), the hashes are grouped on the host (cells x 16 B), and every
// cell is then compared element by element with the representative of its group on the device, so that two cells share a
// block only if all 2880 values are bit-identical.  The planes stay resident when they fit (180 GB of HBM3e); otherwise
// they are streamed twice in chunks.

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace artes {

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {      // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// one thread per cell: fold planes [p0, p0 + np) (buffer row q holds plane p0 + q) into the cell's two running hashes
__global__ void ingest_hash_kernel(const double* __restrict__ planes, size_t cells, int np, int p0,
                                   unsigned long long* __restrict__ h1, unsigned long long* __restrict__ h2) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    unsigned long long a = h1[c], b = h2[c];
    for (int q = 0; q < np; ++q) {
        const unsigned long long v = (unsigned long long)__double_as_longlong(planes[(size_t)q * cells + c]);
        a = (a ^ v) * 1099511628211ull;                                   // FNV-1a over the 64-bit words
        b = mix64(b + v + 0x9e3779b97f4a7c15ull * (unsigned long long)(p0 + q + 1));
    }
    h1[c] = a; h2[c] = b;
}

// every cell against the representative of its group, same planes
__global__ void ingest_verify_kernel(const double* __restrict__ planes, size_t cells, int np, const int* __restrict__ rep, int* mismatch) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const size_t r = (size_t)rep[c];
    if (r == c) return;
    int bad = 0;
    for (int q = 0; q < np; ++q) {
        const double* row = planes + (size_t)q * cells;
        bad |= (__double_as_longlong(row[c]) != __double_as_longlong(row[r])) ? 1 : 0;
    }
    if (bad) atomicOr(mismatch, 1);
}

struct Free { std::vector<void*> p; ~Free() { for (void* q : p) cudaFree(q); } };

// derived per-cell tables of grid_initialize / get_atmosphere (src/ARTES.f90:2172-2188): cell_opacity = k_sca + k_abs,
// cell_albedo = k_sca / cell_opacity floored at 1e-20, and the 32-byte record {opacity, albedo, matrix block, 0} the
// interaction event reads with one 256-bit load
__global__ void ingest_cell_tables_kernel(const double* __restrict__ k_sca, const double* __restrict__ k_abs, const int* __restrict__ c2u,
                                          size_t n, double* __restrict__ kext, double* __restrict__ albedo, double* __restrict__ cellrec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ks = k_sca[i], ke = ks + k_abs[i];
    double al = 0.0;
    if (ke > 0.0) al = ks / ke;
    if (al < 1.e-20) al = 1.e-20;
    kext[i] = ke;
    albedo[i] = al;
    double4 r;
    r.x = ke; r.y = al; r.z = __longlong_as_double((long long)c2u[i]); r.w = 0.0;
    reinterpret_cast<double4*>(cellrec)[i] = r;
}

}  // namespace

cudaError_t ingest_cell_tables(const double* k_sca, const double* k_abs, const int* c2u, size_t n, double* kext, double* albedo,
                               double* cellrec, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ingest_cell_tables_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(k_sca, k_abs, c2u, n, kext, albedo, cellrec);
    return cudaGetLastError();
}

// test hook: > 0 forces the chunked two-pass path with that many planes per chunk (the resident path needs no more than HBM offers)
static int g_chunk_override = 0;
void ingest_set_chunk_override(int planes) { g_chunk_override = planes; }

// src: plane (e, a) of the wavelength = `cells` contiguous doubles at src + plane_stride * (e + 16 a).
// Returns the blocks in order of first appearance (cell index), like the host path.  *exact = 0 if a hash group failed the
// element-wise check (the caller then de-duplicates on the host).
cudaError_t ingest_dedup_device(const double* src, size_t cells, size_t plane_stride, cudaStream_t stream,
                                std::vector<double>& uniq, std::vector<int32_t>& c2u, int* exact, double* copy_ms, double* kernel_ms) {
    constexpr int NPL = 2880;
    *exact = 1;
    Free pool;
    size_t free_b = 0, total_b = 0;
    cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    const size_t plane_bytes = cells * sizeof(double);
    size_t budget = free_b / 2;
    int chunk = (int)std::min<size_t>(NPL, budget / (plane_bytes ? plane_bytes : 1));
    if (g_chunk_override > 0) chunk = std::min(chunk, g_chunk_override);
    if (chunk < 1) return cudaErrorMemoryAllocation;
    const bool resident = chunk == NPL;
    double* buf = nullptr;
    unsigned long long *h1 = nullptr, *h2 = nullptr;
    int *rep = nullptr, *mis = nullptr;
    if ((e = cudaMalloc(&buf, (size_t)chunk * plane_bytes)) != cudaSuccess) return e;
    pool.p.push_back(buf);
    if ((e = cudaMalloc(&h1, cells * 8)) != cudaSuccess) return e;
    pool.p.push_back(h1);
    if ((e = cudaMalloc(&h2, cells * 8)) != cudaSuccess) return e;
    pool.p.push_back(h2);
    if ((e = cudaMalloc(&rep, cells * 4 + 4)) != cudaSuccess) return e;
    pool.p.push_back(rep);
    mis = rep + cells;
    cudaEvent_t ev[4];
    for (auto& x : ev) cudaEventCreate(&x);
    float t_copy = 0.f, t_kern = 0.f;
    auto timed = [&](float& acc, auto&& fn) -> cudaError_t {
        cudaEventRecord(ev[0], stream);
        cudaError_t r = fn();
        cudaEventRecord(ev[1], stream);
        if (r == cudaSuccess) r = cudaStreamSynchronize(stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[0], ev[1]);
        acc += ms;
        return r;
    };
    const unsigned blocks = (unsigned)((cells + 255) / 256);
    auto upload = [&](int p0, int np) {
        return cudaMemcpy2DAsync(buf, plane_bytes, src + plane_stride * (size_t)p0, plane_stride * sizeof(double), plane_bytes, (size_t)np,
                                 cudaMemcpyHostToDevice, stream);
    };
    // ---- pass 1: hashes
    e = cudaMemsetAsync(h1, 0x5a, cells * 8, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h2, 0xc3, cells * 8, stream);
    for (int p0 = 0; p0 < NPL && e == cudaSuccess; p0 += chunk) {
        const int np = std::min(chunk, NPL - p0);
        e = timed(t_copy, [&] { return upload(p0, np); });
        if (e != cudaSuccess) break;
        e = timed(t_kern, [&] { ingest_hash_kernel<<<blocks, 256, 0, stream>>>(buf, cells, np, p0, h1, h2); return cudaGetLastError(); });
    }
    std::vector<unsigned long long> a(cells), b(cells);
    if (e == cudaSuccess) e = cudaMemcpyAsync(a.data(), h1, cells * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b.data(), h2, cells * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    std::vector<int> rep_cell;          // first cell of every group
    if (e == cudaSuccess) {
        struct KeyHash { size_t operator()(const std::pair<unsigned long long, unsigned long long>& k) const { return (size_t)(k.first ^ (k.second * 0x9e3779b97f4a7c15ull)); } };
        std::unordered_map<std::pair<unsigned long long, unsigned long long>, int, KeyHash> groups;
        c2u.assign(cells, 0);
        std::vector<int> rep_h(cells);
        for (size_t c = 0; c < cells; ++c) {
            auto it = groups.find({a[c], b[c]});
            if (it == groups.end()) { it = groups.emplace(std::make_pair(a[c], b[c]), (int)rep_cell.size()).first; rep_cell.push_back((int)c); }
            c2u[c] = it->second;
            rep_h[c] = rep_cell[it->second];
        }
        e = cudaMemcpyAsync(rep, rep_h.data(), cells * 4, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(mis, 0, 4, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);     // rep_h goes out of scope below
    }
    // ---- pass 2: element-wise check of every group (no second upload when the planes are resident)
    for (int p0 = 0; p0 < NPL && e == cudaSuccess; p0 += chunk) {
        const int np = std::min(chunk, NPL - p0);
        if (!resident) { e = timed(t_copy, [&] { return upload(p0, np); }); if (e != cudaSuccess) break; }
        e = timed(t_kern, [&] { ingest_verify_kernel<<<blocks, 256, 0, stream>>>(buf, cells, np, rep, mis); return cudaGetLastError(); });
    }
    int mismatch = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&mismatch, mis, 4, cudaMemcpyDeviceToHost);
    for (auto& x : ev) cudaEventDestroy(x);
    if (e != cudaSuccess) return e;
    if (mismatch) { *exact = 0; return cudaSuccess; }
    // ---- the representatives' blocks, [n_uniq][180][16] (a few thousand strided host reads per block)
    uniq.resize(rep_cell.size() * (size_t)NPL);
    for (size_t u = 0; u < rep_cell.size(); ++u)
        for (int an = 0; an < 180; ++an)
            for (int el = 0; el < 16; ++el)
                uniq[u * NPL + (size_t)an * 16 + el] = src[(size_t)rep_cell[u] + plane_stride * ((size_t)el + 16 * (size_t)an)];
    if (copy_ms) *copy_ms = t_copy;
    if (kernel_ms) *kernel_ms = t_kern;
    return cudaSuccess;
}

}  // namespace artes
