// Random coefficient field on the device: the recipe of tools/generate_st1_field.jl:86-120 --
// white Gaussian noise, real FFT, division by (1 + |k|)^p in Fourier space (:41-84), inverse real FFT, exp(alpha |G|) --
// with cuFFT (D2Z / Z2D, a plain library call, bound at run time like NCCL) around hand-written kernels for the noise,
// the spectral filter, the statistics and the final map.  Optionally the filtered field is normalised to unit variance
// before the exponential (what homogenization.jl_b200/inputs.py does for BASELINE.json configs[4]; the tool's own
// alpha = 100 on the unnormalised field overflows any solver).
//
// Noise: either handed over by the caller (parity tests feed numpy's stream) or generated here with Philox4x32-10
// (key = seed, counter = cell index) and Box-Muller, a generator tests/ restate in numpy.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <cufft.h>
#include <dlfcn.h>

#include "../../include/hmg.h"
#include "hmg_host.hpp"

namespace hmg {
void set_last_error(const std::string& msg);     // api.cu
}

namespace {

using hmg::Error;

#define FCUDA_OK(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            throw Error(std::string("hmg: CUDA error: ") + cudaGetErrorString(e_) + " in " #call); \
    } while (0)

struct CufftApi {
    cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, cufftType, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal*, cufftDoubleComplex*) = nullptr;
    cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex*, cufftDoubleReal*) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
};
const CufftApi& cufft() {
    static CufftApi api;
    static bool loaded = false;
    if (loaded) return api;
    void* h = nullptr;
    for (const char* name : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"})
        if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) throw Error("hmg: cannot load libcufft (the field generator needs cuFFT)");
    auto sym = [&](const char* name) {
        void* f = dlsym(h, name);
        if (!f) throw Error(std::string("hmg: cuFFT symbol missing: ") + name);
        return f;
    };
    api.PlanMany = reinterpret_cast<decltype(api.PlanMany)>(sym("cufftPlanMany"));
    api.SetStream = reinterpret_cast<decltype(api.SetStream)>(sym("cufftSetStream"));
    api.ExecD2Z = reinterpret_cast<decltype(api.ExecD2Z)>(sym("cufftExecD2Z"));
    api.ExecZ2D = reinterpret_cast<decltype(api.ExecZ2D)>(sym("cufftExecZ2D"));
    api.Destroy = reinterpret_cast<decltype(api.Destroy)>(sym("cufftDestroy"));
    loaded = true;
    return api;
}
#define CUFFT_OK(call)                                                                       \
    do {                                                                                     \
        cufftResult r_ = (call);                                                             \
        if (r_ != CUFFT_SUCCESS) throw Error(std::string("hmg: cuFFT error ") + std::to_string((int)r_) + " in " #call); \
    } while (0)

// ---- Philox4x32-10 (Salmon et al. 2011), key = (seed lo, seed hi), counter = (cell lo, cell hi, 0, 0) ----------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// one standard normal per cell: Box-Muller on two 53-bit uniforms in (0, 1)
__global__ void __launch_bounds__(256) noise_kernel(double* __restrict__ g, int64_t n, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)((uint64_t)i >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = ((double)((((uint64_t)c[0] << 32) | c[1]) >> 11) + 0.5) * 0x1.0p-53;
        const double u2 = ((double)((((uint64_t)c[2] << 32) | c[3]) >> 11) + 0.5) * 0x1.0p-53;
        g[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

// B[i, j(, k)] /= (1 + sqrt(x^2 + y^2 + z^2))^p with the wrapped frequency index coord(m, i) = | |i - m - 1| - m |
// of tools/generate_st1_field.jl:41,53-66 (1-based i there); the half-spectrum dimension (cuFFT: the LAST, Julia: the
// first -- the array is the transpose, the physical field the same) runs over 0 .. n/2.  `scale` folds in the 1/N of
// the inverse transform (:38-39).
struct FilterShape { int dim; int n[3]; };
__device__ __forceinline__ double wrapped(int m, int i0) { return fabs(fabs((double)(i0 - m)) - (double)m); }   // i0 = i - 1
__global__ void __launch_bounds__(256) filter_kernel(cufftDoubleComplex* __restrict__ B, FilterShape S, double p, double scale) {
    const int nh = S.n[S.dim - 1] / 2 + 1;
    const int64_t total = (S.dim == 3 ? (int64_t)S.n[0] * S.n[1] : (int64_t)S.n[0]) * nh;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(t % nh);
        int64_t rest = t / nh;
        double k2 = (double)h * (double)h;
        if (S.dim == 3) {
            const int b = (int)(rest % S.n[1]);
            const int a = (int)(rest / S.n[1]);
            const double ya = wrapped(S.n[1] / 2, b), za = wrapped(S.n[0] / 2, a);
            k2 += ya * ya + za * za;
        } else {
            const double ya = wrapped(S.n[0] / 2, (int)rest);
            k2 += ya * ya;
        }
        const double f = scale / pow(1.0 + sqrt(k2), p);
        cufftDoubleComplex v = B[t];
        v.x *= f; v.y *= f;
        B[t] = v;
    }
}

// deterministic two-stage sums (fixed grid, partials added in block order by the host)
__global__ void __launch_bounds__(256) moments_kernel(const double* __restrict__ g, int64_t n, double shift, double* __restrict__ part) {
    __shared__ double s1[256], s2[256];
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = g[i] - shift;
        a += v;
        b = fma(v, v, b);
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s1[0]; part[2 * blockIdx.x + 1] = s2[0]; }
}
__global__ void __launch_bounds__(256) exp_abs_kernel(double* __restrict__ g, int64_t n, double alpha_over_std) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] = exp(alpha_over_std * fabs(g[i]));
}

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { FCUDA_OK(cudaMalloc(&p, bytes ? bytes : 1)); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

void generate_field(int dim, const int* ns, uint64_t seed, double alpha, double p, int normalize, const double* noise,
                    double* out, int device) {
    HMG_CHECK(dim == 2 || dim == 3, "the field generator takes 2 or 3 dimensions");
    HMG_CHECK(ns && out, "null argument");
    int64_t n = 1;
    for (int d = 0; d < dim; ++d) {
        HMG_CHECK(ns[d] >= 2 && ns[d] % 2 == 0, "every extent must be even (tools/generate_st1_field.jl:89)");
        n *= ns[d];
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Error("hmg: no CUDA device available -- this library has no CPU fallback");
    HMG_CHECK(device >= 0 && device < ndev, "device index out of range");
    FCUDA_OK(cudaSetDevice(device));
    FilterShape S{dim, {ns[0], ns[1], dim == 3 ? ns[2] : 1}};
    const int nh = ns[dim - 1] / 2 + 1;
    const int64_t nc = n / ns[dim - 1] * nh;
    cudaStream_t st = nullptr;
    FCUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cufftHandle fwd = 0, inv = 0;
    bool have_fwd = false, have_inv = false;
    try {
        DevBuf real((size_t)n * sizeof(double)), cplx((size_t)nc * sizeof(cufftDoubleComplex));
        double* g = static_cast<double*>(real.p);
        auto* B = static_cast<cufftDoubleComplex*>(cplx.p);
        const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
        if (noise) FCUDA_OK(cudaMemcpyAsync(g, noise, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        else noise_kernel<<<grid, 256, 0, st>>>(g, n, seed);
        int dims[3] = {ns[0], ns[1], dim == 3 ? ns[2] : 0};
        CUFFT_OK(cufft().PlanMany(&fwd, dim, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 1));
        have_fwd = true;
        CUFFT_OK(cufft().PlanMany(&inv, dim, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 1));
        have_inv = true;
        CUFFT_OK(cufft().SetStream(fwd, st));
        CUFFT_OK(cufft().SetStream(inv, st));
        CUFFT_OK(cufft().ExecD2Z(fwd, g, B));
        filter_kernel<<<(unsigned)std::min<int64_t>((nc + 255) / 256, 148 * 8), 256, 0, st>>>(B, S, p, 1.0 / (double)n);
        CUFFT_OK(cufft().ExecZ2D(inv, B, g));
        double scale = alpha;
        if (normalize) {
            // population standard deviation, two passes (mean first), as numpy's std
            const int nb = 148 * 4;
            DevBuf part((size_t)2 * nb * sizeof(double));
            std::vector<double> h((size_t)2 * nb);
            auto pass = [&](double shift, double& s1, double& s2) {
                moments_kernel<<<nb, 256, 0, st>>>(g, n, shift, static_cast<double*>(part.p));
                FCUDA_OK(cudaMemcpyAsync(h.data(), part.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
                FCUDA_OK(cudaStreamSynchronize(st));
                s1 = s2 = 0.0;
                for (int b = 0; b < nb; ++b) { s1 += h[2 * b]; s2 += h[2 * b + 1]; }
            };
            double s1, s2;
            pass(0.0, s1, s2);
            const double mean = s1 / (double)n;
            pass(mean, s1, s2);
            const double sd = std::sqrt(s2 / (double)n);
            HMG_CHECK(sd > 0.0, "the filtered field is constant");
            scale = alpha / sd;
        }
        exp_abs_kernel<<<grid, 256, 0, st>>>(g, n, scale);
        FCUDA_OK(cudaGetLastError());
        FCUDA_OK(cudaMemcpyAsync(out, g, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        FCUDA_OK(cudaStreamSynchronize(st));
    } catch (...) {
        if (have_fwd) cufft().Destroy(fwd);
        if (have_inv) cufft().Destroy(inv);
        cudaStreamDestroy(st);
        throw;
    }
    cufft().Destroy(fwd);
    cufft().Destroy(inv);
    cudaStreamDestroy(st);
}

}  // namespace

extern "C" int hmg_generate_field(int dim, const int* n, uint64_t seed, double alpha, double p, int normalize,
                                  const double* noise_or_null, double* out_host, int device) {
    try {
        generate_field(dim, n, seed, alpha, p, normalize, noise_or_null, out_host, device);
        return 0;
    } catch (const std::exception& ex) {
        hmg::set_last_error(ex.what());
        return 1;
    }
}
