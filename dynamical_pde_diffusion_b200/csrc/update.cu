// Euler predictor / Heun corrector / guidance update of the guided EDM sampler step (sm_100a).
//
// Reference math: src/diffusion_pde/sampling/sample.py:316 (state init), :327-328 (Euler), :330-334 (Heun),
// :354-355 (guidance update).  The reference keeps the state in fp64 and evaluates the denoiser on an fp32 copy;
// these kernels keep that split: fp64 state in HBM, fp32 copies emitted by the same pass that produces the state.
// Arithmetic follows the reference's operation order with explicit round-to-nearest intrinsics (no FMA
// contraction), so the fp64 state is reproduced to the last bit for identical denoiser outputs.
//
// All kernels are pure streaming: 16-byte vector loads/stores, 4 elements per thread per iteration, grid-stride
// over a grid sized to the SM count.  Roofline: HBM bandwidth (bytes per element listed per kernel).

#include "common.cuh"

namespace dpde {

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {

struct D4 {
    double v[4];
};
struct F4 {
    float v[4];
};

__device__ __forceinline__ D4 load_d4(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return D4{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void store_d4(double* p, const D4& d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(d.v[0], d.v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(d.v[2], d.v[3]);
}
__device__ __forceinline__ F4 load_f4(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    return F4{{a.x, a.y, a.z, a.w}};
}
__device__ __forceinline__ void store_f4(float* p, const F4& f) {
    *reinterpret_cast<float4*>(p) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
}

// Correctly rounded a / b for a divisor that is the same for every element of a launch (sigma_cur, sigma_next).
// `Div` carries b and r = RN(1 / b) computed once on the host.  q0 = RN(a r) is within 1.5 ulp of a / b; one residual
// correction makes it faithful, and by Markstein's theorem (r correctly rounded, q1 faithful, exact residual from the
// FMA) the second correction RN(q1 + rem r) IS RN(a / b).  5 fp64 instructions instead of the ~15 + MUFU.RCP64H of the
// general division: the streaming kernels carry 1-3 divisions per element and were fp64 / XU co-limited
// (dpde_euler_predict_bwd at 0.68 of the HBM peak).  Zero, infinite / NaN and results outside the normal range fall
// back to the IEEE division, so the result is bit-identical to __ddiv_rn for every input.
struct Div {
    double b, r;
};
__device__ __forceinline__ double div_rn(double a, const Div& d) {
    const double q0 = __dmul_rn(a, d.r);
    const double q1 = fma(fma(-q0, d.b, a), d.r, q0);
    const double q = fma(fma(-q1, d.b, a), d.r, q1);
    const double m = fabs(q);
    // 2^-960 < |q| < 2^960 keeps a, q b and the residuals normal and exact; zero keeps its sign through q0 (b > 0);
    // everything else (infinities, NaN, results near the ends of the range) takes the IEEE division
    if (!(m > 1.0261342003245941e-289 && m < 9.7453140114e+288)) return a == 0.0 ? q0 : __ddiv_rn(a, d.b);
    return q;
}

// d_cur = (x - x0)/s_cur ; x_eu = x + h d_cur                     (sample.py:327-328)
__device__ __forceinline__ double euler_point(double x, double x0, const Div& s_cur, double h, double& d_cur) {
    d_cur = div_rn(__dsub_rn(x, x0), s_cur);
    return __dadd_rn(x, __dmul_rn(h, d_cur));
}

// ---- x = latents * sigma0 (16 B read, 12 B written per element) -------------------------------------------
__global__ void __launch_bounds__(kThreads) init_kernel(const double* __restrict__ lat, double s0,
                                                         double* __restrict__ x64, float* __restrict__ x32,
                                                         int64_t n, bool vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
#pragma unroll 2
    for (int64_t i = tid; i < n4; i += nth) {
        D4 l = load_d4(lat + 4 * i), o;
        F4 f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o.v[k] = __dmul_rn(l.v[k], s0);
            f.v[k] = (float)o.v[k];
        }
        store_d4(x64 + 4 * i, o);
        store_f4(x32 + 4 * i, f);
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) {
        const double o = __dmul_rn(lat[i], s0);
        x64[i] = o;
        x32[i] = (float)o;
    }
}

// ---- Euler predictor: reads x_cur (8) + x0_cur (4), writes x_eu32 (4) --------------------------------------
__global__ void __launch_bounds__(kThreads) euler_predict_kernel(const double* __restrict__ x, const float* __restrict__ x0,
                                                                  Div s_cur, double h, float* __restrict__ out,
                                                                  int64_t n, bool vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
#pragma unroll 2
    for (int64_t i = tid; i < n4; i += nth) {
        const D4 xv = load_d4(x + 4 * i);
        const F4 dv = load_f4(x0 + 4 * i);
        F4 f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double d;
            f.v[k] = (float)euler_point(xv.v[k], (double)dv.v[k], s_cur, h, d);
        }
        store_f4(out + 4 * i, f);
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) {
        double d;
        out[i] = (float)euler_point(x[i], (double)x0[i], s_cur, h, d);
    }
}

// ---- predictor backward: seed = fp32(-((h g)/s_cur)); 4 B read + 4 B written -------------------------------
__global__ void __launch_bounds__(kThreads) euler_bwd_kernel(const float* __restrict__ g, Div s_cur, double h,
                                                              float* __restrict__ out, int64_t n, bool vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
    // four independent 16-byte loads per thread before the first use: 8 B of traffic per element is too little to hide the HBM
    // latency with one load per thread in flight (2048 threads x 16 B = 32 KB per SM; measured 0.68 of the HBM peak)
    int64_t i = tid;
    for (; i + 3 * nth < n4; i += 4 * nth) {
        F4 gv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) gv[u] = load_f4(g + 4 * (i + u * nth));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            F4 f;
#pragma unroll
            for (int k = 0; k < 4; ++k) f.v[k] = (float)(-div_rn(__dmul_rn(h, (double)gv[u].v[k]), s_cur));
            store_f4(out + 4 * (i + u * nth), f);
        }
    }
    for (; i < n4; i += nth) {
        const F4 gv = load_f4(g + 4 * i);
        F4 f;
#pragma unroll
        for (int k = 0; k < 4; ++k) f.v[k] = (float)(-div_rn(__dmul_rn(h, (double)gv.v[k]), s_cur));
        store_f4(out + 4 * i, f);
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) out[i] = (float)(-div_rn(__dmul_rn(h, (double)g[i]), s_cur));
}

// ---- Heun + guidance update -----------------------------------------------------------------------------------
// reads x_cur (8) + x0_cur (4) + x0_next (4) + g_eu (4) + g_cur (4), writes x_next64 (8) + x_next32 (4): 36 B / element
// (last step: 8 + 4 + 4 read, 12 written).
template <bool LAST, bool HAS_GEU, bool HAS_GCUR>
__device__ __forceinline__ double heun_point(double x, double x0c, double x0n, double geu, double gcur, const Div& s_cur,
                                             const Div& s_next, double h) {
    double d_cur;
    const double x_eu = euler_point(x, x0c, s_cur, h, d_cur);
    double x_new = x_eu;
    if (!LAST) {
        const double d_prime = div_rn(__dsub_rn(x_eu, x0n), s_next);                  // sample.py:333
        const double mix = __dadd_rn(__dmul_rn(0.5, d_cur), __dmul_rn(0.5, d_prime));  // sample.py:334
        x_new = __dadd_rn(x, __dmul_rn(h, mix));
    }
    // gradient w.r.t. x_cur assembled as autograd does: direct path through x_eu, path through d_cur, denoiser path
    double grad = 0.0;
    if (!LAST && HAS_GEU) grad = __dadd_rn(geu, div_rn(__dmul_rn(h, geu), s_cur));
    if (HAS_GCUR) grad = __dadd_rn(grad, gcur);
    return __dsub_rn(x_new, grad);                                                      // sample.py:355
}

// ROWS: the update covers elements [first, first + span) of every plane of `plane` elements (owned rows of a row
// slab); n then counts planes * span.  Flat launches (ROWS == false) index 0..n directly.
template <bool ROWS>
__device__ __forceinline__ int64_t elem_index(int64_t i, int64_t span, int64_t plane, int64_t first) {
    if (!ROWS) return i;
    const int64_t p = i / span;
    return p * plane + first + (i - p * span);
}

template <bool LAST, bool HAS_GEU, bool HAS_GCUR, bool ROWS>
__global__ void __launch_bounds__(kThreads)
heun_update_kernel(const double* __restrict__ x, const float* __restrict__ x0c, const float* __restrict__ x0n,
                   const float* __restrict__ geu, const float* __restrict__ gcur, Div s_cur, Div s_next, double h,
                   double* __restrict__ o64, float* __restrict__ o32, int64_t n, bool vec, int64_t span, int64_t plane,
                   int64_t first) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
#pragma unroll 2
    for (int64_t q = tid; q < n4; q += nth) {
        const int64_t i = elem_index<ROWS>(4 * q, span, plane, first) / 4;   // vec implies span, plane, first % 4 == 0
        const D4 xv = load_d4(x + 4 * i);
        const F4 a = load_f4(x0c + 4 * i);
        F4 b{}, ge{}, gc{};
        if (!LAST) b = load_f4(x0n + 4 * i);
        if (!LAST && HAS_GEU) ge = load_f4(geu + 4 * i);
        if (HAS_GCUR) gc = load_f4(gcur + 4 * i);
        D4 o;
        F4 f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o.v[k] = heun_point<LAST, HAS_GEU, HAS_GCUR>(xv.v[k], (double)a.v[k], (double)b.v[k], (double)ge.v[k],
                                                         (double)gc.v[k], s_cur, s_next, h);
            f.v[k] = (float)o.v[k];
        }
        store_d4(o64 + 4 * i, o);
        store_f4(o32 + 4 * i, f);
    }
    for (int64_t q = 4 * n4 + tid; q < n; q += nth) {
        const int64_t i = elem_index<ROWS>(q, span, plane, first);
        const double o = heun_point<LAST, HAS_GEU, HAS_GCUR>(
            x[i], (double)x0c[i], LAST ? 0.0 : (double)x0n[i], (!LAST && HAS_GEU) ? (double)geu[i] : 0.0,
            HAS_GCUR ? (double)gcur[i] : 0.0, s_cur, s_next, h);
        o64[i] = o;
        o32[i] = (float)o;
    }
}

// ---- Heun + guidance update of a row slab FUSED with the halo exchange (one kernel: compute + transfer over NVLink) ----
// The owned boundary rows (2 x halo per plane) come first in the index space; every unit of them is stored to the local
// next-state buffers AND through the mapped peer pointers into the neighbours' ghost rows.  CTAs that held boundary units
// fence (system scope) and take a ticket; the last of them publishes the epoch in the neighbours' flag words with
// st.release.sys.  All CTAs then continue with the interior rows, so the NVLink transfer and the neighbours' wake-up
// overlap the bulk of the update.  Safe against the neighbours' reads of the same ghost rows: see slab.py ("ordering").
struct PushArgs {
    double *up64, *dn64;
    float *up32, *dn32;
    unsigned long long *flag_up, *flag_dn;
    unsigned int* ticket;
    unsigned long long epoch;
    int H_up, H_dn;
};

template <bool LAST, bool HAS_GEU, bool HAS_GCUR, bool VEC>
__device__ __forceinline__ void update_unit(const double* __restrict__ x, const float* __restrict__ x0c, const float* __restrict__ x0n,
                                            const float* __restrict__ geu, const float* __restrict__ gcur, const Div& s_cur,
                                            const Div& s_next, double h, int64_t i, D4& o, F4& f) {
    if (VEC) {
        const D4 xv = load_d4(x + i);
        const F4 a = load_f4(x0c + i);
        F4 b{}, ge{}, gc{};
        if (!LAST) b = load_f4(x0n + i);
        if (!LAST && HAS_GEU) ge = load_f4(geu + i);
        if (HAS_GCUR) gc = load_f4(gcur + i);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o.v[k] = heun_point<LAST, HAS_GEU, HAS_GCUR>(xv.v[k], (double)a.v[k], (double)b.v[k], (double)ge.v[k], (double)gc.v[k],
                                                         s_cur, s_next, h);
            f.v[k] = (float)o.v[k];
        }
    } else {
        o.v[0] = heun_point<LAST, HAS_GEU, HAS_GCUR>(x[i], (double)x0c[i], LAST ? 0.0 : (double)x0n[i],
                                                     (!LAST && HAS_GEU) ? (double)geu[i] : 0.0, HAS_GCUR ? (double)gcur[i] : 0.0,
                                                     s_cur, s_next, h);
        f.v[0] = (float)o.v[0];
    }
}

template <bool VEC>
__device__ __forceinline__ void store_unit(double* __restrict__ p64, float* __restrict__ p32, int64_t i, const D4& o, const F4& f) {
    if (VEC) {
        store_d4(p64 + i, o);
        store_f4(p32 + i, f);
    } else {
        p64[i] = o.v[0];
        p32[i] = f.v[0];
    }
}

template <bool LAST, bool HAS_GEU, bool HAS_GCUR, bool VEC>
__global__ void __launch_bounds__(kThreads)
heun_update_slab_kernel(const double* __restrict__ x, const float* __restrict__ x0c, const float* __restrict__ x0n,
                        const float* __restrict__ geu, const float* __restrict__ gcur, Div s_cur, Div s_next, double h,
                        double* __restrict__ o64, float* __restrict__ o32, int64_t planes, int H, int W, int halo,
                        const __grid_constant__ PushArgs pa, unsigned int n_boundary_ctas) {
    constexpr int U = VEC ? 4 : 1;
    const int rowu = W / U;                                          // units per row
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t side_units = (int64_t)halo * rowu, nb = planes * 2 * side_units;
    // ---- boundary rows: local store + peer store
    for (int64_t q = tid; q < nb; q += nth) {
        const int64_t pl = q / (2 * side_units), rem = q - pl * 2 * side_units;
        const int side = rem >= side_units, rr = (int)(rem - side * side_units);
        const int row = rr / rowu, col = (rr - row * rowu) * U;
        const int lr = side ? H - 2 * halo + row : halo + row;
        const int64_t i = (pl * H + lr) * (int64_t)W + col;
        D4 o;
        F4 f;
        update_unit<LAST, HAS_GEU, HAS_GCUR, VEC>(x, x0c, x0n, geu, gcur, s_cur, s_next, h, i, o, f);
        store_unit<VEC>(o64, o32, i, o, f);
        if (!side && pa.up64) store_unit<VEC>(pa.up64, pa.up32, (pl * pa.H_up + (pa.H_up - halo + row)) * (int64_t)W + col, o, f);
        if (side && pa.dn64) store_unit<VEC>(pa.dn64, pa.dn32, (pl * pa.H_dn + row) * (int64_t)W + col, o, f);
    }
    if (blockIdx.x < n_boundary_ctas) {
        __threadfence_system();                                      // this thread's peer stores are visible system-wide ...
        __syncthreads();                                             // ... for every thread of the CTA, before its ticket
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(pa.ticket, 1u) == n_boundary_ctas - 1) {
                *pa.ticket = 0u;
                __threadfence_system();
                if (pa.flag_up) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pa.flag_up), "l"(pa.epoch) : "memory");
                if (pa.flag_dn) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pa.flag_dn), "l"(pa.epoch) : "memory");
            }
        }
    }
    // ---- interior rows [2 halo, H - 2 halo)
    const int irows = H - 4 * halo;
    const int64_t plane_units = (int64_t)irows * rowu, ni = planes * plane_units;
#pragma unroll 2
    for (int64_t q = tid; q < ni; q += nth) {
        const int64_t pl = q / plane_units, rem = q - pl * plane_units;
        const int64_t i = (pl * H + 2 * halo) * (int64_t)W + rem * U;
        D4 o;
        F4 f;
        update_unit<LAST, HAS_GEU, HAS_GCUR, VEC>(x, x0c, x0n, geu, gcur, s_cur, s_next, h, i, o, f);
        store_unit<VEC>(o64, o32, i, o, f);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int stream_grid(int64_t n) {
    int64_t blocks = (n / 4 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace dpde

using namespace dpde;

extern "C" {

int dpde_abi_version(void) { return DPDE_ABI_VERSION; }
const char* dpde_last_error(void) { return error_buffer(); }

int dpde_sampler_init(const double* latents, double sigma0, double* x64, float* x32, int64_t n, dpde_stream_t stream) {
    if (!latents || !x64 || !x32 || n < 0) return fail(DPDE_ERR_INVALID, "dpde_sampler_init: null pointer or n < 0");
    if (n == 0) return DPDE_OK;
    const bool vec = aligned16(latents) && aligned16(x64) && aligned16(x32);
    init_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(latents, sigma0, x64, x32, n, vec);
    return check_launch("dpde_sampler_init");
}

int dpde_euler_predict(const double* x_cur, const float* x0_cur, double sigma_cur, double sigma_next, float* x_eu32,
                       int64_t n, dpde_stream_t stream) {
    if (!x_cur || !x0_cur || !x_eu32 || n < 0) return fail(DPDE_ERR_INVALID, "dpde_euler_predict: null pointer or n < 0");
    if (!(sigma_cur > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_euler_predict: sigma_cur must be > 0");
    if (n == 0) return DPDE_OK;
    const bool vec = aligned16(x_cur) && aligned16(x0_cur) && aligned16(x_eu32);
    euler_predict_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(x_cur, x0_cur, Div{sigma_cur, 1.0 / sigma_cur},
                                                                               sigma_next - sigma_cur, x_eu32, n, vec);
    return check_launch("dpde_euler_predict");
}

int dpde_euler_predict_bwd(const float* g_eu32, double sigma_cur, double sigma_next, float* seed32, int64_t n,
                           dpde_stream_t stream) {
    if (!g_eu32 || !seed32 || n < 0) return fail(DPDE_ERR_INVALID, "dpde_euler_predict_bwd: null pointer or n < 0");
    if (!(sigma_cur > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_euler_predict_bwd: sigma_cur must be > 0");
    if (n == 0) return DPDE_OK;
    const bool vec = aligned16(g_eu32) && aligned16(seed32);
    euler_bwd_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(g_eu32, Div{sigma_cur, 1.0 / sigma_cur},
                                                                           sigma_next - sigma_cur, seed32, n, vec);
    return check_launch("dpde_euler_predict_bwd");
}

namespace {
int heun_launch(const char* who, const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                const float* g_cur, double sigma_cur, double sigma_next, double* x_next64, float* x_next32, int64_t n,
                bool rows, int64_t span, int64_t plane, int64_t first, dpde_stream_t stream) {
    if (!x_cur || !x0_cur || !x_next64 || !x_next32 || n < 0) return fail(DPDE_ERR_INVALID, "%s: null pointer or n < 0", who);
    if (!(sigma_cur > 0.0)) return fail(DPDE_ERR_INVALID, "%s: sigma_cur must be > 0", who);
    const bool last = (x0_next == nullptr);
    if (!last && !(sigma_next > 0.0)) return fail(DPDE_ERR_INVALID, "%s: Heun correction needs sigma_next > 0", who);
    if (n == 0) return DPDE_OK;
    bool vec = aligned16(x_cur) && aligned16(x0_cur) && aligned16(x_next64) && aligned16(x_next32) &&
               (last || aligned16(x0_next)) && (!g_eu || aligned16(g_eu)) && (!g_cur || aligned16(g_cur));
    if (rows) vec = vec && span % 4 == 0 && plane % 4 == 0 && first % 4 == 0;
    const double h = sigma_next - sigma_cur;
    const Div d_cur{sigma_cur, 1.0 / sigma_cur}, d_next{sigma_next, last ? 0.0 : 1.0 / sigma_next};
    const int grid = stream_grid(n);
    cudaStream_t s = (cudaStream_t)stream;
#define DPDE_LAUNCH2(L, GE, GC, R)                                                                                    \
    heun_update_kernel<L, GE, GC, R><<<grid, kThreads, 0, s>>>(x_cur, x0_cur, x0_next, g_eu, g_cur, d_cur, d_next, \
                                                               h, x_next64, x_next32, n, vec, span, plane, first)
#define DPDE_LAUNCH(L, GE, GC) \
    do { if (rows) DPDE_LAUNCH2(L, GE, GC, true); else DPDE_LAUNCH2(L, GE, GC, false); } while (0)
    if (last) {
        if (g_cur) DPDE_LAUNCH(true, false, true); else DPDE_LAUNCH(true, false, false);
    } else if (g_eu) {
        if (g_cur) DPDE_LAUNCH(false, true, true); else DPDE_LAUNCH(false, true, false);
    } else {
        if (g_cur) DPDE_LAUNCH(false, false, true); else DPDE_LAUNCH(false, false, false);
    }
#undef DPDE_LAUNCH
#undef DPDE_LAUNCH2
    return check_launch(who);
}
}  // namespace

int dpde_heun_guided_update(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                            const float* g_cur, double sigma_cur, double sigma_next, double* x_next64, float* x_next32,
                            int64_t n, dpde_stream_t stream) {
    return heun_launch("dpde_heun_guided_update", x_cur, x0_cur, x0_next, g_eu, g_cur, sigma_cur, sigma_next, x_next64,
                       x_next32, n, false, n, n, 0, stream);
}

int dpde_heun_guided_update_rows(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                                 const float* g_cur, double sigma_cur, double sigma_next, double* x_next64,
                                 float* x_next32, int64_t planes, int64_t plane_elems, int64_t first, int64_t count,
                                 dpde_stream_t stream) {
    if (planes < 0 || plane_elems < 0 || first < 0 || count < 0 || first + count > plane_elems)
        return fail(DPDE_ERR_INVALID, "dpde_heun_guided_update_rows: need 0 <= first, first + count <= plane_elems");
    return heun_launch("dpde_heun_guided_update_rows", x_cur, x0_cur, x0_next, g_eu, g_cur, sigma_cur, sigma_next,
                       x_next64, x_next32, planes * count, true, count, plane_elems, first, stream);
}

int dpde_heun_guided_update_rows_push(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                                      const float* g_cur, double sigma_cur, double sigma_next, double* x_next64,
                                      float* x_next32, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                                      const dpde_halo_peers* peers, dpde_stream_t stream) {
    const char* who = "dpde_heun_guided_update_rows_push";
    if (!x_cur || !x0_cur || !x_next64 || !x_next32 || !peers) return fail(DPDE_ERR_INVALID, "%s: null pointer", who);
    if (planes < 0 || W < 1 || halo < 1 || H_local < 4 * halo) return fail(DPDE_ERR_INVALID, "%s: need halo >= 1 and H_local >= 4 halo", who);
    if (!(sigma_cur > 0.0)) return fail(DPDE_ERR_INVALID, "%s: sigma_cur must be > 0", who);
    const bool last = (x0_next == nullptr);
    if (!last && !(sigma_next > 0.0)) return fail(DPDE_ERR_INVALID, "%s: Heun correction needs sigma_next > 0", who);
    if (!peers->ticket) return fail(DPDE_ERR_INVALID, "%s: ticket is NULL", who);
    if (peers->epoch == 0) return fail(DPDE_ERR_INVALID, "%s: epoch must be >= 1 (flags start zeroed)", who);
    if ((peers->up64 != nullptr) != (peers->up32 != nullptr) || (peers->down64 != nullptr) != (peers->down32 != nullptr))
        return fail(DPDE_ERR_INVALID, "%s: a neighbour needs both its fp64 and its fp32 buffer", who);
    if ((peers->up64 && (!peers->flag_up || peers->H_up < 3 * halo)) || (peers->down64 && (!peers->flag_down || peers->H_down < 3 * halo)))
        return fail(DPDE_ERR_INVALID, "%s: neighbour without flag or shorter than 3 halo", who);
    if (planes == 0) return DPDE_OK;
    bool vec = W % 4 == 0 && aligned16(x_cur) && aligned16(x0_cur) && aligned16(x_next64) && aligned16(x_next32) &&
               (last || aligned16(x0_next)) && (!g_eu || aligned16(g_eu)) && (!g_cur || aligned16(g_cur)) &&
               (!peers->up64 || (aligned16(peers->up64) && aligned16(peers->up32))) &&
               (!peers->down64 || (aligned16(peers->down64) && aligned16(peers->down32)));
    const int unit = vec ? 4 : 1;
    const int64_t n_units = planes * (int64_t)(H_local - 2 * halo) * (W / unit);
    const int64_t nb_units = planes * 2 * (int64_t)halo * (W / unit);
    int64_t blocks = (n_units + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int64_t nbc = (nb_units + kThreads - 1) / kThreads;              // CTAs whose first grid-stride pass holds boundary units
    if (nbc > blocks) nbc = blocks;
    PushArgs pa{peers->up64, peers->down64, peers->up32, peers->down32, (unsigned long long*)peers->flag_up,
                (unsigned long long*)peers->flag_down, (unsigned int*)peers->ticket, peers->epoch, peers->H_up, peers->H_down};
    const Div d_cur{sigma_cur, 1.0 / sigma_cur}, d_next{sigma_next, last ? 0.0 : 1.0 / sigma_next};
    const double h = sigma_next - sigma_cur;
    cudaStream_t s = (cudaStream_t)stream;
#define DPDE_SLAB2(L, GE, GC, V)                                                                                          \
    heun_update_slab_kernel<L, GE, GC, V><<<(int)blocks, kThreads, 0, s>>>(x_cur, x0_cur, x0_next, g_eu, g_cur, d_cur, d_next, h, \
                                                                           x_next64, x_next32, planes, H_local, W, halo, pa, (unsigned)nbc)
#define DPDE_SLAB(L, GE, GC) \
    do { if (vec) DPDE_SLAB2(L, GE, GC, true); else DPDE_SLAB2(L, GE, GC, false); } while (0)
    if (last) {
        if (g_cur) DPDE_SLAB(true, false, true); else DPDE_SLAB(true, false, false);
    } else if (g_eu) {
        if (g_cur) DPDE_SLAB(false, true, true); else DPDE_SLAB(false, true, false);
    } else {
        if (g_cur) DPDE_SLAB(false, false, true); else DPDE_SLAB(false, false, false);
    }
#undef DPDE_SLAB
#undef DPDE_SLAB2
    return check_launch(who);
}

}  // extern "C"

// ---- row-slab halo staging (config 5: 4096^2 grids split over GPUs by rows) -----------------------------------
namespace dpde {
namespace {
// DIR 0: field -> staging (pack owned boundary rows); DIR 1: staging -> field ghost rows (unpack).
template <typename T, int DIR>
__global__ void __launch_bounds__(kThreads)
halo_rows_kernel(T* __restrict__ field, int64_t planes, int H, int W, int halo, T* __restrict__ up, T* __restrict__ down) {
    const int64_t per_plane = (int64_t)halo * W, total = planes * per_plane;
    const int up_row = DIR == 0 ? halo : 0, down_row = DIR == 0 ? H - 2 * halo : H - halo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pl = i / per_plane, r = i - pl * per_plane;
        T* fu = field + pl * (int64_t)H * W + (int64_t)up_row * W + r;
        T* fd = field + pl * (int64_t)H * W + (int64_t)down_row * W + r;
        if (DIR == 0) {
            if (up) up[i] = *fu;
            if (down) down[i] = *fd;
        } else {
            if (up) *fu = up[i];
            if (down) *fd = down[i];
        }
    }
}

template <int DIR>
int halo_launch(void* field, int32_t dtype, int64_t planes, int32_t H, int32_t W, int32_t halo, void* up, void* down,
                dpde_stream_t stream, const char* who) {
    if (!field) return fail(DPDE_ERR_INVALID, "%s: field is NULL", who);
    if (planes < 0 || W < 1 || halo < 1 || H < 3 * halo) return fail(DPDE_ERR_INVALID, "%s: need halo >= 1 and H_local >= 3 halo", who);
    if (dtype != DPDE_F32 && dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "%s: dtype must be f32/f64", who);
    if (planes == 0 || (!up && !down)) return DPDE_OK;
    const int64_t total = planes * (int64_t)halo * W;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == DPDE_F32)
        halo_rows_kernel<float, DIR><<<(int)blocks, kThreads, 0, s>>>((float*)field, planes, H, W, halo, (float*)up, (float*)down);
    else
        halo_rows_kernel<double, DIR><<<(int)blocks, kThreads, 0, s>>>((double*)field, planes, H, W, halo, (double*)up, (double*)down);
    return check_launch(who);
}
}  // namespace
}  // namespace dpde

extern "C" {
int dpde_halo_pack(const void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                   void* send_up, void* send_down, dpde_stream_t stream) {
    return dpde::halo_launch<0>(const_cast<void*>(field), dtype, planes, H_local, W, halo, send_up, send_down, stream,
                                "dpde_halo_pack");
}
int dpde_halo_unpack(void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                     const void* recv_up, const void* recv_down, dpde_stream_t stream) {
    return dpde::halo_launch<1>(field, dtype, planes, H_local, W, halo, const_cast<void*>(recv_up),
                                const_cast<void*>(recv_down), stream, "dpde_halo_unpack");
}
}
