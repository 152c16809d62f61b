// Final reduction of the path (kernel K5): cross-chain sufficient statistics for posterior moments
// and ESS in one kernel, and their sum over the GPUs of a job through NCCL -- both behind the C ABI, so a Julia host
// reaches them with `ccall` exactly like the Python harness does.  The reference has no equivalent (single chain, no
// ESS estimator); the statistics are the ones SURVEY.md 8d defines.
//
// NCCL is bound at run time (dlopen of libnccl.so.2; a copy already loaded by the host process -- e.g. PyTorch's -- is
// reused), so libpdmpflux_cuda.so carries no link-time dependency on it and single-GPU hosts never load it.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

int pdmpflux_fail_(int code, const std::string& msg);  // api.cu: sets the thread-local error message
void pdmpflux_count_launch_();                          // api.cu: gpu_launches counter

namespace {

constexpr int kTile = 32;    // coordinates per block
constexpr int kRows = 8;     // chain lanes per block
constexpr int kMaxSplit = 128;

// partial[s][r][i], r = 0: sum_c m_c, 1: sum_c m_c^2, 2: sum_c s_c  with m_c = m1[c][i] / T[c], s_c = m2[c][i] / T[c]
__global__ void __launch_bounds__(kTile* kRows) moments_partial_kernel(int d, int64_t n_chains, const double* __restrict__ m1,
                                                                       const double* __restrict__ m2,
                                                                       const double* __restrict__ T, double* __restrict__ partial) {
    __shared__ double sh[3][kRows][kTile + 1];
    const int i = blockIdx.x * kTile + threadIdx.x;
    const int64_t per = (n_chains + gridDim.y - 1) / gridDim.y;
    const int64_t c0 = (int64_t)blockIdx.y * per, c1 = min(n_chains, c0 + per);
    double a = 0.0, b = 0.0, s = 0.0;
    if (i < d)
        for (int64_t c = c0 + threadIdx.y; c < c1; c += kRows) {  // fixed order: bitwise reproducible
            const double inv = T ? 1.0 / T[c] : 1.0;
            const double m = m1[c * d + i] * inv;
            a += m; b += m * m; s += m2[c * d + i] * inv;
        }
    sh[0][threadIdx.y][threadIdx.x] = a; sh[1][threadIdx.y][threadIdx.x] = b; sh[2][threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y < 3 && i < d) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) t += sh[threadIdx.y][r][threadIdx.x];
        partial[((size_t)blockIdx.y * 3 + threadIdx.y) * d + i] = t;
    }
}

// sums[r][i] = sum_s partial[s][r][i] (r < 3), sums[3][i] = n_chains
__global__ void __launch_bounds__(256) moments_final_kernel(int d, int n_split, double count, const double* __restrict__ partial,
                                                            double* __restrict__ sums) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 4 * d) return;
    const int r = e / d, i = e - r * d;
    if (r == 3) { sums[e] = count; return; }
    double t = 0.0;
    for (int s = 0; s < n_split; ++s) t += partial[((size_t)s * 3 + r) * d + i];
    sums[e] = t;
}

// ---- NCCL, bound at run time ------------------------------------------------------------------------------
struct Nccl {
    struct Id { char b[PDMPFLUX_COMM_ID_BYTES]; };  // ncclUniqueId (passed by value)
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {  // a copy the host process already holds (PyTorch bundles one) wins
            n.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (n.lib) break;
        }
        for (int k = 0; !n.lib && k < 2; ++k) n.lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
        if (!n.lib) { n.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char* s) { void* p = dlsym(n.lib, s); if (!p) n.why = std::string("NCCL symbol missing: ") + s; return p; };
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return n;
}

int nccl_fail(const char* what, int rc) {
    Nccl& n = nccl();
    return pdmpflux_fail_(PDMPFLUX_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}

}  // namespace

struct pdmpflux_comm_s {
    void* comm = nullptr;
    int nranks = 1, rank = 0;
};

extern "C" {
#pragma GCC visibility push(default)

int pdmpflux_moments_reduce(int dim, int64_t n_chains, const double* m1, const double* m2, const double* T, double* sums,
                            int32_t on_device, void* stream_) {
    if (!m1 || !m2 || !sums || dim <= 0 || n_chains <= 0) return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "invalid argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return pdmpflux_fail_(PDMPFLUX_ERR_CUDA, "no CUDA device: no CPU fallback");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int n_split = (int)std::min<int64_t>(kMaxSplit, (n_chains + 255) / 256);
    const size_t nd = sizeof(double) * (size_t)dim * n_chains;
    double *d1 = nullptr, *d2 = nullptr, *dT = nullptr, *dS = nullptr, *part = nullptr;
    // Partial sums (<= 3 MB): a grow-only device buffer cached per host thread and device, so the per-step call issues
    // no allocation at all (a stream-ordered allocation here costs a pool round trip every step).  Calls made by one
    // host thread on the same device must therefore be issued on one stream at a time.
    struct Scratch { double* p = nullptr; size_t bytes = 0; int dev = -1; };
    thread_local Scratch scratch;
    auto cleanup = [&] {
        if (!on_device) { if (d1) cudaFreeAsync(d1, stream); if (d2) cudaFreeAsync(d2, stream); if (dT) cudaFreeAsync(dT, stream); if (dS) cudaFreeAsync(dS, stream); }
    };
#define TRY_(expr)                                                                                                   \
    do {                                                                                                             \
        cudaError_t e_ = (expr);                                                                                     \
        if (e_ != cudaSuccess) { cleanup(); return pdmpflux_fail_(PDMPFLUX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } \
    } while (0)
    {
        int dev = 0;
        TRY_(cudaGetDevice(&dev));
        const size_t need = sizeof(double) * 3 * (size_t)dim * n_split;
        if (scratch.dev != dev || scratch.bytes < need) {
            if (scratch.p && scratch.dev == dev) cudaFree(scratch.p);   // (a buffer of another device is left to its context)
            scratch = Scratch{};
            TRY_(cudaMalloc(reinterpret_cast<void**>(&scratch.p), need));
            scratch.bytes = need; scratch.dev = dev;
        }
        part = scratch.p;
    }
    const double *p1 = m1, *p2 = m2, *pT = T;
    double* pS = sums;
    if (!on_device) {
        TRY_(cudaMallocAsync(&d1, nd, stream)); TRY_(cudaMallocAsync(&d2, nd, stream));
        TRY_(cudaMallocAsync(&dS, sizeof(double) * 4 * dim, stream));
        TRY_(cudaMemcpyAsync(d1, m1, nd, cudaMemcpyHostToDevice, stream));
        TRY_(cudaMemcpyAsync(d2, m2, nd, cudaMemcpyHostToDevice, stream));
        if (T) {
            TRY_(cudaMallocAsync(&dT, sizeof(double) * n_chains, stream));
            TRY_(cudaMemcpyAsync(dT, T, sizeof(double) * n_chains, cudaMemcpyHostToDevice, stream));
        }
        p1 = d1; p2 = d2; pT = dT; pS = dS;
    }
    dim3 grid((unsigned)((dim + kTile - 1) / kTile), (unsigned)n_split), block(kTile, kRows);
    moments_partial_kernel<<<grid, block, 0, stream>>>(dim, n_chains, p1, p2, pT, part);
    TRY_(cudaGetLastError());
    moments_final_kernel<<<(4 * dim + 255) / 256, 256, 0, stream>>>(dim, n_split, (double)n_chains, part, pS);
    TRY_(cudaGetLastError());
    pdmpflux_count_launch_(); pdmpflux_count_launch_();
    if (!on_device) {
        TRY_(cudaMemcpyAsync(sums, pS, sizeof(double) * 4 * dim, cudaMemcpyDeviceToHost, stream));
        TRY_(cudaStreamSynchronize(stream));
    }
    cleanup();
#undef TRY_
    return PDMPFLUX_OK;
}

int pdmpflux_comm_unique_id(void* id_out, size_t bytes) {
    if (!id_out || bytes < PDMPFLUX_COMM_ID_BYTES) return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "id_out must hold PDMPFLUX_COMM_ID_BYTES bytes");
    Nccl& n = nccl();
    if (!n.GetUniqueId) return pdmpflux_fail_(PDMPFLUX_ERR_UNSUPPORTED, "NCCL is not available: " + n.why);
    const int rc = n.GetUniqueId(id_out);
    return rc == 0 ? PDMPFLUX_OK : nccl_fail("ncclGetUniqueId", rc);
}

int pdmpflux_comm_create(const void* unique_id, int n_ranks, int rank, pdmpflux_comm_t* out) {
    if (!out) return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (n_ranks <= 0 || rank < 0 || rank >= n_ranks) return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "need 0 <= rank < n_ranks");
    auto c = new pdmpflux_comm_s();
    c->nranks = n_ranks; c->rank = rank;
    if (n_ranks > 1) {
        if (!unique_id) { delete c; return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "unique_id is NULL"); }
        Nccl& n = nccl();
        if (!n.CommInitRank) { delete c; return pdmpflux_fail_(PDMPFLUX_ERR_UNSUPPORTED, "NCCL is not available: " + n.why); }
        Nccl::Id id;
        std::memcpy(id.b, unique_id, sizeof(id.b));
        const int rc = n.CommInitRank(&c->comm, n_ranks, id, rank);
        if (rc != 0) { delete c; return nccl_fail("ncclCommInitRank", rc); }
    }
    *out = c;
    return PDMPFLUX_OK;
}

int pdmpflux_comm_destroy(pdmpflux_comm_t c) {
    if (c && c->comm) nccl().CommDestroy(c->comm);
    delete c;
    return PDMPFLUX_OK;
}

int pdmpflux_moments_allreduce(pdmpflux_comm_t c, double* sums, int64_t n, void* stream_) {
    if (!c || !sums || n <= 0) return pdmpflux_fail_(PDMPFLUX_ERR_ARGUMENT, "invalid argument");
    if (c->nranks == 1) return PDMPFLUX_OK;  // a single rank owns every chain: nothing to add
    const int rc = nccl().AllReduce(sums, sums, (size_t)n, /* ncclFloat64 */ 8, /* ncclSum */ 0, c->comm, static_cast<cudaStream_t>(stream_));
    return rc == 0 ? PDMPFLUX_OK : nccl_fail("ncclAllReduce", rc);
}

#pragma GCC visibility pop
}
