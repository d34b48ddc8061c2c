// PDE-residual + masked-observation losses and their analytic VJP w.r.t. the denoised estimate (sm_100a).
//
// Replaces, per guided step, the reference's ~40 ATen ops and their autograd mirror:
//   sample.py:336-353  masked observation losses, loss_fn call, weighted combination
//   sample.py:106-134  laplacian (reflect pad + 3x3 conv, fp64)
//   pde_losses.py:71-117  heat_loss2 / llg_loss2
//   tests/test_llg_pde_loss.py:70-117  m x H_eff residual (exchange + applied field; uniaxial anisotropy added)
// by two passes over the fields:
//   pass 1 (guidance_reduce_kernel): three global sums, deterministic two-level reduction, scalars finalised by the
//           last CTA (no host round trip: the reference syncs 6 times per step for .item() / mask.sum());
//   pass 2 (guidance_vjp_kernel): seed gradient d loss_comb / d x0-hat written once, in the dtype the denoiser
//           backward consumes.  The residual r (heat) / the field-gradient G_H (LLG) is staged in shared memory
//           with a one-pixel ring so the transposed stencil K^T needs no second global pass.
//
// Numerics: inputs are the denoiser's fp32 output (exactly representable in the reference's fp64 copy,
// sample.py:325); all arithmetic and all accumulation here is fp64, the result is rounded once on store --
// the same place the reference's `.to(fp64)` backward rounds (sample.py:325,332).
//
// Memory-bound stencil: no tensor cores.  Work is cut into tiles of TILE_PIX pixels (tile_w in {16,32,64,128}
// chosen from W so narrow grids such as the reference's 64x16 LLG film keep all lanes busy); a persistent grid
// (SM count x occupancy) strides over tiles; consecutive lanes touch consecutive columns (coalesced 128 B rows).

#include <atomic>
#include <type_traits>

#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <cudaTypedefs.h>

#include "common.cuh"

namespace dpde {
namespace {

struct View {
    const void* p;
    int dtype;
    int64_t sb, sc;
};

struct Params {
    int B, C, ch_a, H, W, kind, has_a, has_u;
    int Hg, yg0, ylo, yhi;  // row slab: global height, global row of local row 0, owned local rows [ylo, yhi)
    int n_u_units, units_per_sample;
    int tile_w, tile_h, tw_log2, tiles_x, tiles_y;
    int64_t tiles_per_plane, n_tiles;
    View x0, dxdt, obs_a, mask_a, obs_u, mask_u;
    const double* coef;
    double inv_dx2, w_a, w_u, w_pde, hw;
    double gamma, alpha, c_ex, c_an, tau, e[3];
    // cross-rank exchange of the partial sums (row slabs / coupled batch shards): mb_world > 0 makes the last CTA of a
    // reduce pass post this rank's three sums into every rank's mailbox (peer stores + st.release.sys), see peer.cu
    int mb_world, mb_rank;
    unsigned long long mb_epoch;
    void* mb_box[DPDE_MAX_RANKS];
};

// One mailbox = 2 parities x DPDE_MAX_RANKS source ranks x {sums[3], flag}: 32 bytes per slot
struct MailSlot {
    double s[3];
    unsigned long long flag;
};
static_assert(sizeof(MailSlot) == 32 && 2 * DPDE_MAX_RANKS * sizeof(MailSlot) == DPDE_MAILBOX_BYTES, "mailbox layout");

__device__ __forceinline__ MailSlot* mail_slot(void* box, unsigned long long epoch, int src) {
    return reinterpret_cast<MailSlot*>(box) + (epoch & 1ull) * DPDE_MAX_RANKS + src;
}

struct TileCoord {
    int b, unit, y0, x0;
};

__device__ __forceinline__ TileCoord decode_tile(const Params& p, int64_t t) {
    const int64_t pu = t / p.tiles_per_plane;
    const int r = (int)(t - pu * p.tiles_per_plane);
    TileCoord c;
    c.b = (int)(pu / p.units_per_sample);
    c.unit = (int)(pu - (int64_t)c.b * p.units_per_sample);
    const int ty = r / p.tiles_x;
    c.y0 = p.ylo + ty * p.tile_h;
    c.x0 = (r - ty * p.tiles_x) * p.tile_w;
    return c;
}

template <typename T>
__device__ __forceinline__ double ldg_d(const T* p) {
    return (double)__ldg(p);
}

// unscaled 5-point sum with reflect boundaries (sample.py:126-133): u[-1] = u[1], u[H] = u[H-2]
// `y` indexes the local buffer; `gy` is the same row in the global grid of height Hg (gy == y without slabs).
template <typename T>
__device__ __forceinline__ double lap5(const T* __restrict__ u, int y, int x, int gy, int Hg, int W, double& centre) {
    const int yu = (gy == 0) ? y + 1 : y - 1, yd = (gy == Hg - 1) ? y - 1 : y + 1;
    const int xl = (x == 0) ? 1 : x - 1, xr = (x == W - 1) ? W - 2 : x + 1;
    const T* row = u + (int64_t)y * W;
    centre = ldg_d(row + x);
    return ((ldg_d(u + (int64_t)yu * W + x) + ldg_d(u + (int64_t)yd * W + x)) + (ldg_d(row + xl) + ldg_d(row + xr))) -
           4.0 * centre;
}

// weights of the transposed stencil: neighbour q contributes twice when it sits on the boundary line (row/col 0 or
// n-1) -- its reflect padding read the target twice.
__device__ __forceinline__ double adj_w(int q, int n) { return (q == 0 || q == n - 1) ? 2.0 : 1.0; }

__device__ __forceinline__ double masked_diff(const View& obs, const View& mask, int b, int ch, int64_t pix, double x,
                                              double& m) {
    m = ld_any(mask.p, mask.dtype, (int64_t)b * mask.sb + (int64_t)ch * mask.sc + pix);
    const double o = ld_any(obs.p, obs.dtype, (int64_t)b * obs.sb + (int64_t)ch * obs.sc + pix);
    return m * (x - o);  // (mask * (x - obs)), sample.py:340,342
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Is the residual at local row y / column x defined and needed?  Rows ylo-1 .. yhi feed the transposed stencil of
// the owned rows [ylo, yhi); the row must exist in the global grid (and therefore, with >= 2 ghost rows, its
// vertical neighbours exist in the local buffer).
__device__ __forceinline__ bool residual_needed(const Params& p, int y, int x) {
    const int gy = y + p.yg0;
    return y >= p.ylo - 1 && y <= p.yhi && gy >= 0 && gy < p.Hg && x >= 0 && x < p.W;
}

// heat residual r = dudt - alpha * lap(u)   (pde_losses.py:91-94)
template <typename T>
__device__ __forceinline__ double heat_residual(const Params& p, const T* __restrict__ u, const T* __restrict__ dudt,
                                                double alpha, int y, int x, double& centre) {
    const double lap = lap5(u, y, x, y + p.yg0, p.Hg, p.W, centre) * p.inv_dx2;
    const double dt = dudt ? ldg_d(dudt + (int64_t)y * p.W + x) : 0.0;
    return dt - alpha * lap;
}

struct LLGPoint {
    double m[3], Hf[3], a[3], r[3];
};

// r = dmdt - tau (-gamma m x H - alpha m x (m x H)),  H = h_ext + c_ex lap(m) + c_an (m.e) e
// (tests/test_llg_pde_loss.py:70-117, pde_losses.py:246-250)
template <typename T>
__device__ __forceinline__ void llg_point(const Params& p, const T* __restrict__ m0, int64_t msc,
                                          const T* __restrict__ d0, int64_t dsc, const double* hext, int y, int x,
                                          LLGPoint& o) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double lap = lap5(m0 + k * msc, y, x, y + p.yg0, p.Hg, p.W, o.m[k]) * p.inv_dx2;
        o.Hf[k] = hext[k] + p.c_ex * lap;
    }
    if (p.c_an != 0.0) {
        const double me = p.c_an * (o.m[0] * p.e[0] + o.m[1] * p.e[1] + o.m[2] * p.e[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) o.Hf[k] += me * p.e[k];
    }
    cross3(o.m, o.Hf, o.a);
    double ma[3];
    cross3(o.m, o.a, ma);
    const int64_t pix = (int64_t)y * p.W + x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double rhs = -p.gamma * o.a[k] - p.alpha * ma[k];
        const double dt = d0 ? ldg_d(d0 + k * dsc + pix) : 0.0;
        o.r[k] = dt - rhs * p.tau;
    }
}

// sums -> losses and seed coefficients (sample.py:340-353; pde_losses.py:94,116)
__device__ __forceinline__ void finalize_scalars(const Params& p, const double* sums, double* scal, float* trace) {
    const double la = p.has_a ? sqrt(sums[0]) : 0.0;
    const double lu = p.has_u ? sqrt(sums[1]) : 0.0;
    double lp = 0.0, cp = 0.0;
    if (p.kind == DPDE_PDE_HEAT) {
        lp = sqrt(sums[2] / p.hw);
        cp = p.w_pde / (p.hw * lp);
    } else if (p.kind == DPDE_PDE_LLG_NORM || p.kind == DPDE_PDE_LLG_RESIDUAL) {
        const double root = sqrt(sums[2]);
        lp = root / p.hw;
        cp = p.w_pde / (root * p.hw);
    }
    const double comb = (p.w_a * la + p.w_u * lu) + p.w_pde * lp;
    scal[0] = la;
    scal[1] = lu;
    scal[2] = lp;
    scal[3] = comb;
    scal[4] = p.has_a ? p.w_a / la : 0.0;
    scal[5] = p.has_u ? p.w_u / lu : 0.0;
    scal[6] = cp;
    scal[7] = 0.0;
    if (trace) {
        trace[0] = (float)la;
        trace[1] = (float)lu;
        trace[2] = (float)lp;
        trace[3] = (float)comb;
    }
}

// ---- end of every reduce pass: deterministic CTA partial -> the last CTA (ticket counter) combines all partials in index
//      order, optionally finalises the scalars, and -- fused compute + collective -- posts this rank's sums to every
//      rank's mailbox over NVLink peer memory: thread r stores the three doubles into slot [epoch & 1][my rank] of rank
//      r's mailbox and publishes them with st.release.sys on the slot's flag word.  dpde_mailbox_wait_finalize on each
//      rank acquires the `world` flags and adds the slots in rank order, so every rank forms the same total.
template <int NT>
__device__ __forceinline__ void reduce_epilogue_n(const Params& p, double s_a, double s_u, double s_p, double* scratch, bool* is_last,
                                                  double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                  double* __restrict__ sums, int finalize, double* __restrict__ scal,
                                                  float* __restrict__ trace, int part_base = 0) {
    // part_base > 0: an earlier kernel of the same stream (the LEAN kernel of a pass) already filled slots [0, part_base)
    const int tid = threadIdx.x;
    block_sum3<NT>(s_a, s_u, s_p, scratch);
    if (tid == 0) {
        partials[3 * (part_base + blockIdx.x) + 0] = s_a;
        partials[3 * (part_base + blockIdx.x) + 1] = s_u;
        partials[3 * (part_base + blockIdx.x) + 2] = s_p;
        __threadfence();
        *is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = tid; i < part_base + (int)gridDim.x; i += NT) {
        a += __ldcg(partials + 3 * i);
        b += __ldcg(partials + 3 * i + 1);
        c += __ldcg(partials + 3 * i + 2);
    }
    block_sum3<NT>(a, b, c, scratch);
    if (tid == 0) {
        sums[0] = a;
        sums[1] = b;
        sums[2] = c;
        if (finalize) finalize_scalars(p, sums, scal, trace);
        *ticket = 0u;
        scratch[0] = a;
        scratch[1] = b;
        scratch[2] = c;
    }
    if (p.mb_world > 0) {
        __syncthreads();
        if (tid < p.mb_world) {
            MailSlot* slot = mail_slot(p.mb_box[tid], p.mb_epoch, p.mb_rank);
            slot->s[0] = scratch[0];
            slot->s[1] = scratch[1];
            slot->s[2] = scratch[2];
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&slot->flag), "l"(p.mb_epoch) : "memory");
        }
    }
}

__device__ __forceinline__ void reduce_epilogue(const Params& p, double s_a, double s_u, double s_p, double* scratch, bool* is_last,
                                                double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                double* __restrict__ sums, int finalize, double* __restrict__ scal,
                                                float* __restrict__ trace) {
    reduce_epilogue_n<kThreads>(p, s_a, s_u, s_p, scratch, is_last, partials, ticket, sums, finalize, scal, trace);
}

// =========================================================================================================
// pass 1: global sums
// =========================================================================================================
template <typename T, int KIND, int TILE_PIX>
__global__ void __launch_bounds__(kThreads)
guidance_reduce_kernel(const __grid_constant__ Params p, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                       double* __restrict__ sums, int finalize, double* __restrict__ scal, float* __restrict__ trace) {
    constexpr int PPT = TILE_PIX / kThreads;
    __shared__ double scratch[3 * (kThreads / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    double s_a = 0.0, s_u = 0.0, s_p = 0.0;
    const T* x0 = reinterpret_cast<const T*>(p.x0.p);
    const T* dx = reinterpret_cast<const T*>(p.dxdt.p);

    for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        if (tc.unit < p.ch_a) {  // ---- initial-condition channel: observation loss only
            if (!p.has_a) continue;
            const T* a = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)tc.unit * p.x0.sc;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, y = tc.y0 + (e >> p.tw_log2), x = tc.x0 + (e & (p.tile_w - 1));
                if (y < p.yhi && x < p.W) {
                    double m;
                    const int64_t pix = (int64_t)y * p.W + x;
                    const double d = masked_diff(p.obs_a, p.mask_a, tc.b, tc.unit, pix, ldg_d(a + pix), m);
                    s_a += d * d;
                }
            }
            continue;
        }
        const int cu = tc.unit - p.ch_a;  // u-unit index
        if (KIND == DPDE_PDE_HEAT || KIND == DPDE_PDE_NONE) {
            const int ch = p.ch_a + cu;
            const T* u = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)ch * p.x0.sc;
            const T* du = dx ? dx + (int64_t)tc.b * p.dxdt.sb + (int64_t)ch * p.dxdt.sc : nullptr;
            const double alpha = (KIND == DPDE_PDE_HEAT) ? __ldg(p.coef + tc.b) : 0.0;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, y = tc.y0 + (e >> p.tw_log2), x = tc.x0 + (e & (p.tile_w - 1));
                if (y < p.yhi && x < p.W) {
                    double uc;
                    if (KIND == DPDE_PDE_HEAT) {
                        const double r = heat_residual(p, u, du, alpha, y, x, uc);
                        s_p += r * r;
                    } else {
                        uc = ldg_d(u + (int64_t)y * p.W + x);
                    }
                    if (p.has_u) {
                        double m;
                        const double d = masked_diff(p.obs_u, p.mask_u, tc.b, cu, (int64_t)y * p.W + x, uc, m);
                        s_u += d * d;
                    }
                }
            }
        } else {  // LLG kinds: one unit = the 3-component magnetisation
            const T* m0 = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc;
            const T* d0 = dx ? dx + (int64_t)tc.b * p.dxdt.sb + (int64_t)p.ch_a * p.dxdt.sc : nullptr;
            double hext[3] = {0.0, 0.0, 0.0};
            if (KIND == DPDE_PDE_LLG_RESIDUAL) {
                hext[0] = __ldg(p.coef + 3 * tc.b);
                hext[1] = __ldg(p.coef + 3 * tc.b + 1);
                hext[2] = __ldg(p.coef + 3 * tc.b + 2);
            }
#pragma unroll 1
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, y = tc.y0 + (e >> p.tw_log2), x = tc.x0 + (e & (p.tile_w - 1));
                if (y < p.yhi && x < p.W) {
                    const int64_t pix = (int64_t)y * p.W + x;
                    double mv[3];
                    if (KIND == DPDE_PDE_LLG_RESIDUAL) {
                        LLGPoint q;
                        llg_point(p, m0, p.x0.sc, d0, p.dxdt.sc, hext, y, x, q);
                        s_p += (q.r[0] * q.r[0] + q.r[1] * q.r[1]) + q.r[2] * q.r[2];
                        mv[0] = q.m[0]; mv[1] = q.m[1]; mv[2] = q.m[2];
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) mv[c] = ldg_d(m0 + c * p.x0.sc + pix);
                        const double n = sqrt((mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2]);
                        s_p += (1.0 - n) * (1.0 - n);  // pde_losses.py:115-116
                    }
                    if (p.has_u) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            double m;
                            const double d = masked_diff(p.obs_u, p.mask_u, tc.b, c, pix, mv[c], m);
                            s_u += d * d;
                        }
                    }
                }
            }
        }
    }

    reduce_epilogue(p, s_a, s_u, s_p, scratch, &is_last, partials, ticket, sums, finalize, scal, trace);
}

__global__ void finalize_kernel(const __grid_constant__ Params p, const double* __restrict__ sums,
                                double* __restrict__ scal, float* __restrict__ trace) {
    if (threadIdx.x == 0 && blockIdx.x == 0) finalize_scalars(p, sums, scal, trace);
}

// Other half of the mailbox exchange: one warp.  Lane r < world spins (ld.acquire.sys, bounded by a timeout) on the flag
// of slot [epoch & 1][r] of THIS rank's mailbox, i.e. on rank r's post of the current step; the slots are then added in
// rank order -- the same order on every rank, so all ranks finalise identical totals -- and the scalars are finalised.
__global__ void mailbox_wait_finalize_kernel(const __grid_constant__ Params p, unsigned long long timeout_ns, int* __restrict__ status,
                                             double* __restrict__ sums, double* __restrict__ scal, float* __restrict__ trace) {
    const int lane = threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    if (lane < p.mb_world) {
        const MailSlot* slot = mail_slot(p.mb_box[p.mb_rank], p.mb_epoch, lane);
        unsigned long long t0, now, f;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&slot->flag) : "memory");
            if (f >= p.mb_epoch) break;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) {
                ok = false;
                break;
            }
            __nanosleep(100);
        }
        if (ok) {
            v[0] = slot->s[0];
            v[1] = slot->s[1];
            v[2] = slot->s[2];
        }
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    double tot[3] = {0.0, 0.0, 0.0};
    for (int r = 0; r < p.mb_world; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) tot[k] += __shfl_sync(0xffffffffu, v[k], r);
    }
    if (lane == 0) {
        if (!all_ok && status) *status = 1;
        sums[0] = tot[0];
        sums[1] = tot[1];
        sums[2] = tot[2];
        finalize_scalars(p, sums, scal, trace);
    }
}

// =========================================================================================================
// pass 2: seed gradient
// =========================================================================================================
// ring position e in [0, 2*(tw+2) + 2*th) -> tile-relative (ry, rx) in [-1, th] x [-1, tw]
__device__ __forceinline__ void ring_coord(int e, int th, int tw, int& ry, int& rx) {
    const int SW = tw + 2;
    if (e < SW) {
        ry = -1;
        rx = e - 1;
    } else if (e < 2 * SW) {
        ry = th;
        rx = e - SW - 1;
    } else {
        const int j = e - 2 * SW;
        ry = j >> 1;
        rx = (j & 1) ? tw : -1;
    }
}

template <typename T, int KIND, int TILE_PIX>
__global__ void __launch_bounds__(kThreads)
guidance_vjp_kernel(const __grid_constant__ Params p, const double* __restrict__ scal, const double* __restrict__ upstream,
                    T* __restrict__ g_x0, T* __restrict__ g_dxdt) {
    constexpr int PPT = TILE_PIX / kThreads;
    constexpr bool STENCIL = (KIND == DPDE_PDE_HEAT || KIND == DPDE_PDE_LLG_RESIDUAL);
    // widest staged tile: (TILE_PIX/128 + 2) x (128 + 2) pixels, 1 (heat) or 3 (LLG) doubles each
    constexpr int STAGE = (TILE_PIX / 128 + 2) * 130;
    constexpr int NCOMP = (KIND == DPDE_PDE_LLG_RESIDUAL) ? 3 : 1;
    __shared__ double stage[STENCIL ? NCOMP * STAGE : 1];

    const int tid = threadIdx.x;
    const double up = upstream ? __ldg(upstream) : 1.0;
    const double c_a = __ldg(scal + 4) * up, c_u = __ldg(scal + 5) * up, c_p = __ldg(scal + 6) * up;
    const T* x0 = reinterpret_cast<const T*>(p.x0.p);
    const T* dx = reinterpret_cast<const T*>(p.dxdt.p);
    const int64_t plane = (int64_t)p.H * p.W;
    const int SW = p.tile_w + 2;

    for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        if (tc.unit < p.ch_a) {  // ---- a-channel: g = c_a * mask * (mask * (a - obs))
            const T* a = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)tc.unit * p.x0.sc;
            T* g = g_x0 + ((int64_t)tc.b * p.C + tc.unit) * plane;
            T* gd = g_dxdt ? g_dxdt + ((int64_t)tc.b * p.C + tc.unit) * plane : nullptr;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, y = tc.y0 + (e >> p.tw_log2), x = tc.x0 + (e & (p.tile_w - 1));
                if (y < p.yhi && x < p.W) {
                    const int64_t pix = (int64_t)y * p.W + x;
                    double v = 0.0;
                    if (p.has_a) {
                        double m;
                        const double d = masked_diff(p.obs_a, p.mask_a, tc.b, tc.unit, pix, ldg_d(a + pix), m);
                        v = c_a * (m * d);
                    }
                    g[pix] = (T)v;
                    if (gd) gd[pix] = (T)0;
                }
            }
            continue;
        }
        const int cu = tc.unit - p.ch_a;

        if (KIND == DPDE_PDE_HEAT || KIND == DPDE_PDE_NONE) {
            const int ch = p.ch_a + cu;
            const T* u = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)ch * p.x0.sc;
            const T* du = dx ? dx + (int64_t)tc.b * p.dxdt.sb + (int64_t)ch * p.dxdt.sc : nullptr;
            T* g = g_x0 + ((int64_t)tc.b * p.C + ch) * plane;
            T* gd = g_dxdt ? g_dxdt + ((int64_t)tc.b * p.C + ch) * plane : nullptr;
            const double alpha = (KIND == DPDE_PDE_HEAT) ? __ldg(p.coef + tc.b) : 0.0;
            double uc[PPT];
            if (KIND == DPDE_PDE_HEAT) {
                // phase 1: residual on the tile's own pixels; phase 2: on its one-pixel ring
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const int e = k * kThreads + tid, ly = e >> p.tw_log2, lx = e & (p.tile_w - 1);
                    const int y = tc.y0 + ly, x = tc.x0 + lx;
                    double r = 0.0;
                    uc[k] = 0.0;
                    if (residual_needed(p, y, x)) r = heat_residual(p, u, du, alpha, y, x, uc[k]);
                    stage[(ly + 1) * SW + lx + 1] = r;
                }
                const int ring = 2 * SW + 2 * p.tile_h;
                for (int e = tid; e < ring; e += kThreads) {
                    int ry, rx;
                    ring_coord(e, p.tile_h, p.tile_w, ry, rx);
                    const int y = tc.y0 + ry, x = tc.x0 + rx;
                    double r = 0.0, dummy;
                    if (residual_needed(p, y, x)) r = heat_residual(p, u, du, alpha, y, x, dummy);
                    stage[(ry + 1) * SW + rx + 1] = r;
                }
                __syncthreads();
            }
            // phase 3: g_u = c_u mask (mask (u - obs)) - c_p (alpha/dx^2) K^T r ;  g_dudt = c_p r
            const double kp = -c_p * alpha * p.inv_dx2;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, ly = e >> p.tw_log2, lx = e & (p.tile_w - 1);
                const int y = tc.y0 + ly, x = tc.x0 + lx;
                if (y < p.yhi && x < p.W) {
                    const int64_t pix = (int64_t)y * p.W + x;
                    double v = 0.0;
                    if (KIND == DPDE_PDE_HEAT) {
                        const double* c = &stage[(ly + 1) * SW + lx + 1];
                        const double acc = ((adj_w(y + p.yg0 - 1, p.Hg) * c[-SW] + adj_w(y + p.yg0 + 1, p.Hg) * c[SW]) +
                                            (adj_w(x - 1, p.W) * c[-1] + adj_w(x + 1, p.W) * c[1])) - 4.0 * c[0];
                        v = kp * acc;
                        if (gd) gd[pix] = (T)(c_p * c[0]);
                    } else {
                        uc[k] = ldg_d(u + pix);
                        if (gd) gd[pix] = (T)0;
                    }
                    if (p.has_u) {
                        double m;
                        const double d = masked_diff(p.obs_u, p.mask_u, tc.b, cu, pix, uc[k], m);
                        v += c_u * (m * d);
                    }
                    g[pix] = (T)v;
                }
            }
            if (KIND == DPDE_PDE_HEAT) __syncthreads();  // stage is reused by the next tile
        } else if (KIND == DPDE_PDE_LLG_NORM) {
            // g_m = -c_p (1 - n) m / n  (0 where n == 0, as torch.linalg.norm's backward)  + observation term
            const T* m0 = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc;
            T* g = g_x0 + ((int64_t)tc.b * p.C + p.ch_a) * plane;
            T* gd = g_dxdt ? g_dxdt + ((int64_t)tc.b * p.C + p.ch_a) * plane : nullptr;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, y = tc.y0 + (e >> p.tw_log2), x = tc.x0 + (e & (p.tile_w - 1));
                if (y < p.yhi && x < p.W) {
                    const int64_t pix = (int64_t)y * p.W + x;
                    double mv[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) mv[c] = ldg_d(m0 + c * p.x0.sc + pix);
                    const double n = sqrt((mv[0] * mv[0] + mv[1] * mv[1]) + mv[2] * mv[2]);
                    const double f = (n > 0.0) ? -c_p * (1.0 - n) / n : 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double v = f * mv[c];
                        if (p.has_u) {
                            double m;
                            const double d = masked_diff(p.obs_u, p.mask_u, tc.b, c, pix, mv[c], m);
                            v += c_u * (m * d);
                        }
                        g[c * plane + pix] = (T)v;
                        if (gd) gd[c * plane + pix] = (T)0;
                    }
                }
            }
        } else {  // ---- LLG m x H_eff residual
            const T* m0 = x0 + (int64_t)tc.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc;
            const T* d0 = dx ? dx + (int64_t)tc.b * p.dxdt.sb + (int64_t)p.ch_a * p.dxdt.sc : nullptr;
            T* g = g_x0 + ((int64_t)tc.b * p.C + p.ch_a) * plane;
            T* gd = g_dxdt ? g_dxdt + ((int64_t)tc.b * p.C + p.ch_a) * plane : nullptr;
            const double hext[3] = {__ldg(p.coef + 3 * tc.b), __ldg(p.coef + 3 * tc.b + 1), __ldg(p.coef + 3 * tc.b + 2)};
            double local[PPT][3];  // -tau [G_m + c_an (e.G_H) e] + observation term, per own pixel
            // With seed s = c_p r, q = s x m:  G_H = -gamma q - alpha (q x m),
            //                                   G_m = -gamma (H x s) - alpha (a x s + H x q)
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, ly = e >> p.tw_log2, lx = e & (p.tile_w - 1);
                const int y = tc.y0 + ly, x = tc.x0 + lx;
                double GH[3] = {0.0, 0.0, 0.0};
                local[k][0] = local[k][1] = local[k][2] = 0.0;
                if (residual_needed(p, y, x)) {
                    const bool owned = y < p.yhi;
                    const int64_t pix = (int64_t)y * p.W + x;
                    LLGPoint q;
                    llg_point(p, m0, p.x0.sc, d0, p.dxdt.sc, hext, y, x, q);
                    double s[3] = {c_p * q.r[0], c_p * q.r[1], c_p * q.r[2]}, qq[3], qm[3], Hs[3], as[3], Hq[3];
                    cross3(s, q.m, qq);
                    cross3(qq, q.m, qm);
                    cross3(q.Hf, s, Hs);
                    cross3(q.a, s, as);
                    cross3(q.Hf, qq, Hq);
                    double eG = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        GH[c] = -p.gamma * qq[c] - p.alpha * qm[c];
                        eG += p.e[c] * GH[c];
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double Gm = -p.gamma * Hs[c] - p.alpha * (as[c] + Hq[c]);
                        double v = -p.tau * (Gm + p.c_an * eG * p.e[c]);
                        if (p.has_u && owned) {
                            double m;
                            const double d = masked_diff(p.obs_u, p.mask_u, tc.b, c, pix, q.m[c], m);
                            v += c_u * (m * d);
                        }
                        local[k][c] = v;
                        if (gd && owned) gd[c * plane + pix] = (T)s[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) stage[c * STAGE + (ly + 1) * SW + lx + 1] = GH[c];
            }
            const int ring = 2 * SW + 2 * p.tile_h;
            for (int e = tid; e < ring; e += kThreads) {
                int ry, rx;
                ring_coord(e, p.tile_h, p.tile_w, ry, rx);
                const int y = tc.y0 + ry, x = tc.x0 + rx;
                double GH[3] = {0.0, 0.0, 0.0};
                if (residual_needed(p, y, x)) {
                    LLGPoint q;
                    llg_point(p, m0, p.x0.sc, d0, p.dxdt.sc, hext, y, x, q);
                    double s[3] = {c_p * q.r[0], c_p * q.r[1], c_p * q.r[2]}, qq[3], qm[3];
                    cross3(s, q.m, qq);
                    cross3(qq, q.m, qm);
#pragma unroll
                    for (int c = 0; c < 3; ++c) GH[c] = -p.gamma * qq[c] - p.alpha * qm[c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) stage[c * STAGE + (ry + 1) * SW + rx + 1] = GH[c];
            }
            __syncthreads();
            const double kx = -p.tau * p.c_ex * p.inv_dx2;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const int e = k * kThreads + tid, ly = e >> p.tw_log2, lx = e & (p.tile_w - 1);
                const int y = tc.y0 + ly, x = tc.x0 + lx;
                if (y < p.yhi && x < p.W) {
                    const int64_t pix = (int64_t)y * p.W + x;
                    const double wu = adj_w(y + p.yg0 - 1, p.Hg), wd = adj_w(y + p.yg0 + 1, p.Hg), wl = adj_w(x - 1, p.W), wr = adj_w(x + 1, p.W);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double* s = &stage[c * STAGE + (ly + 1) * SW + lx + 1];
                        const double acc = ((wu * s[-SW] + wd * s[SW]) + (wl * s[-1] + wr * s[1])) - 4.0 * s[0];
                        g[c * plane + pix] = (T)(local[k][c] + kx * acc);
                    }
                }
            }
            __syncthreads();
        }
    }
}

// =========================================================================================================
// stand-alone laplacian(u, dx) and its transpose (Level-1 op, sample.py:106-134)
// =========================================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
laplacian_kernel(const T* __restrict__ u, T* __restrict__ out, int64_t planes, int H, int W, int64_t stride_in,
                 double inv_dx2, int adjoint) {
    const int64_t hw = (int64_t)H * W, total = planes * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pl = i / hw;
        const int r = (int)(i - pl * hw), y = r / W, x = r - y * W;
        const T* src = u + pl * stride_in;
        double v;
        if (!adjoint) {
            double c;
            v = lap5(src, y, x, y, H, W, c);
        } else {
            const T* row = src + (int64_t)y * W;
            v = -4.0 * ldg_d(row + x);
            if (y > 0) v += adj_w(y - 1, H) * ldg_d(row - W + x);
            if (y < H - 1) v += adj_w(y + 1, H) * ldg_d(row + W + x);
            if (x > 0) v += adj_w(x - 1, W) * ldg_d(row + x - 1);
            if (x < W - 1) v += adj_w(x + 1, W) * ldg_d(row + x + 1);
        }
        out[i] = (T)(v * inv_dx2);
    }
}

// =========================================================================================================
// per-sample heat residual of the training-time physics loss (EDMHeatLoss, models/loss.py:143):
//   out[b] = sum_{c,h,w} (dudt - alpha_b lap(u))^2          (the caller applies 1/(H W), mean / sum and coeff / sigma^2)
// and its VJP   g_u = up_b 2 (-alpha_b / dx^2) K^T r,   g_dudt = up_b 2 r.
// Two-level deterministic reduction: kPerSampleBlocks partial sums per sample, combined in index order.
// =========================================================================================================
constexpr int kPerSampleBlocks = 1024;   // upper bound on partial sums per sample (workspace = B x this)

template <typename T>
__global__ void __launch_bounds__(kThreads)
heat_residual_sq_kernel(const T* __restrict__ u, const T* __restrict__ dudt, int64_t sb_u, int64_t sc_u, int64_t sb_d, int64_t sc_d,
                        const double* __restrict__ alpha, int Cu, int H, int W, double inv_dx2, double* __restrict__ partials) {
    __shared__ double scratch[3 * (kThreads / 32)];
    const int b = blockIdx.y, hw = H * W;
    const int64_t total = (int64_t)Cu * hw;
    const double a_s = __ldg(alpha + b) * inv_dx2;
    double acc = 0.0, z0 = 0.0, z1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int c = (int)(i / hw), r = (int)(i - (int64_t)c * hw), y = r / W, x = r - y * W;
        double centre;
        const double s = lap5(u + b * sb_u + c * sc_u, y, x, y, H, W, centre);
        const double dt = dudt ? ldg_d(dudt + b * sb_d + c * sc_d + r) : 0.0;
        const double res = dt - a_s * s;
        acc += res * res;
    }
    block_sum3(acc, z0, z1, scratch);
    if (threadIdx.x == 0) partials[(int64_t)b * gridDim.x + blockIdx.x] = acc;
}

__global__ void per_sample_combine_kernel(const double* __restrict__ partials, int nblk, int B, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int k = 0; k < nblk; ++k) s += partials[(int64_t)b * nblk + k];
    out[b] = s;
}

// marching path: partials[item], the items of sample b are b, b + B, ... (n_per of them).  One warp per sample: lane l
// adds items l, l + 32, ... in order, then a fixed shuffle tree -- deterministic.
__global__ void per_sample_items_kernel(const double* __restrict__ partials, int n_per, int B, double* __restrict__ out) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    double s = 0.0;
    for (int k = lane; k < n_per; k += 32) s += partials[(int64_t)k * B + b];
    s = warp_sum(s);
    if (lane == 0) out[b] = s;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
heat_residual_sq_vjp_kernel(const T* __restrict__ u, const T* __restrict__ dudt, int64_t sb_u, int64_t sc_u, int64_t sb_d, int64_t sc_d,
                            const double* __restrict__ alpha, const double* __restrict__ upstream, int B, int Cu, int H, int W,
                            double inv_dx2, T* __restrict__ g_u, T* __restrict__ g_dudt) {
    const int hw = H * W;
    const int64_t total = (int64_t)B * Cu * hw;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int64_t pl = i / hw;
        const int b = (int)(pl / Cu), c = (int)(pl - (int64_t)b * Cu), r = (int)(i - pl * hw), y = r / W, x = r - y * W;
        const T* up = u + b * sb_u + c * sc_u;
        const T* dp = dudt ? dudt + b * sb_d + c * sc_d : nullptr;
        const double a_s = __ldg(alpha + b) * inv_dx2, seed = 2.0 * __ldg(upstream + b);
        auto res = [&](int yy, int xx) {
            double centre;
            const double s = lap5(up, yy, xx, yy, H, W, centre);
            return (dp ? ldg_d(dp + (int64_t)yy * W + xx) : 0.0) - a_s * s;
        };
        const double rc = res(y, x);
        double acc = -4.0 * rc;
        if (y > 0) acc += adj_w(y - 1, H) * res(y - 1, x);
        if (y < H - 1) acc += adj_w(y + 1, H) * res(y + 1, x);
        if (x > 0) acc += adj_w(x - 1, W) * res(y, x - 1);
        if (x < W - 1) acc += adj_w(x + 1, W) * res(y, x + 1);
        g_u[i] = (T)(seed * (-a_s) * acc);
        if (g_dudt) g_dudt[i] = (T)(seed * rc);
    }
}

#include "heat_march.cuh"
#include "llg_tile.cuh"
#include "llg_march.cuh"

// =========================================================================================================
// host side
// =========================================================================================================
// Test / tuning hooks (dpde_set_fast_path, dpde_set_tuning): process-wide, read once per launch.  They are atomics so a
// concurrent setter is not a data race, but a launch in flight on another thread may see either value: set them
// before the threads that launch start (the tests and scripts/kernel_probe.py are their only callers).
std::atomic<bool> g_fast_path{true};
// experiment knobs (dpde_set_tuning): [0] strip layout 0 = per pass (reduce 120 columns + 1 halo lane, VJP 112 + 2,
// sector aligned), 1 = 120 + 1 in both, 2 = 112 + 2 in both;
// [1] unused; [2] rows per chunk (0 = automatic); [3] / [4] 1 = pair a-planes with u-planes in the
// reduce / VJP pass (measured slower than separate streaming items on 8x2x4096^2: 3.0 vs 3.8 TB/s);
// [5] 1 = the LLG marching kernels send interior work items through their general loop too (A/B runs of the lean loop);
// [6] 1 = the LLG m x H_eff residual runs the tile kernels on every size, 2 = the marching kernels on every size with W >= 128
//     (default: marching on large grids, tiles on small ones)
// [7] 1 = the LLG marching kernels without TMA: reduce pass with the cp.async feed, VJP as the two-CTA kernel with register windows
std::atomic<int> g_tuning[8] = {};

inline bool al(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
inline bool s4(const View& v) { return v.sb % 4 == 0 && v.sc % 4 == 0; }

// Can the row-marching kernels take this problem?  (else the generic tile kernels run)
// float4 per streaming work item: kABlock on large problems, halved until there is about one item per resident warp.  Not
// further: an item of 32 k float4 keeps k loads per lane in flight, and on L2-resident problems (the bench workload:
// 64 x 128^2 planes) a warp that walks through four 32-float4 items pays four dependent memory latencies instead of one.
inline int stream_block4(int64_t plane4, int64_t planes) {
    int blk = kABlock;
    const int64_t want = (int64_t)sm_count() * 8;
    while (blk > 32 && ((plane4 + blk - 1) / blk) * planes < want) blk >>= 1;
    return blk;
}

bool march_eligible(const Params& p, const void* g_x0, const void* g_dxdt) {
    if (!g_fast_path || p.kind != DPDE_PDE_HEAT || p.x0.dtype != DPDE_F32 || p.W % 4 != 0) return false;
    if ((int64_t)p.H * p.W >= (1ll << 30) || (int64_t)p.B * p.C * ((p.H + 3) / 4) * ((p.W + 111) / 112) >= (1ll << 30)) return false;
    if (!al(p.x0.p, 16) || !s4(p.x0)) return false;
    if (p.dxdt.p && (!al(p.dxdt.p, 16) || !s4(p.dxdt))) return false;
    if (g_x0 && !al(g_x0, 16)) return false;
    if (g_dxdt && !al(g_dxdt, 16)) return false;
    if (p.has_a && (p.obs_a.dtype != DPDE_F32 || p.mask_a.dtype != DPDE_U8 || !al(p.obs_a.p, 16) || !al(p.mask_a.p, 4) ||
                    !s4(p.obs_a) || !s4(p.mask_a))) return false;
    if (p.has_u && (p.obs_u.dtype != DPDE_F32 || p.mask_u.dtype != DPDE_U8 || !al(p.obs_u.p, 16) || !al(p.mask_u.p, 4) ||
                    !s4(p.obs_u) || !s4(p.mask_u))) return false;
    return true;
}

// 0: a-planes are separate streaming items; 1: paired with the u-plane of the same index; 2: paired, mask_a empty
inline int pairing(const Params& p, bool vjp) {
    return ((vjp ? g_tuning[4] != 0 : g_tuning[3] != 0) && p.ch_a >= 1 && p.ch_a == p.n_u_units) ? (p.has_a ? 1 : 2) : 0;
}

MarchGeom march_geometry(const Params& p, bool vjp) {
    MarchGeom g{};
    const int rows = p.yhi - p.ylo;
    if (p.W <= 128) {
        int lw = 1, l2 = 0;
        while (lw * 4 < p.W) { lw <<= 1; ++l2; }
        g.lw_log2 = l2; g.segs_per_warp = 32 / lw; g.strips = 1; g.strip_w = p.W; g.halo_lane = 0;
    } else {
        g.lw_log2 = 5; g.segs_per_warp = 1;
        // strip layout per pass (measured, 8x2x4096^2): the reduce pass needs one halo column and is issue-bound, so the
        // 120 + 1 layout (6 % idle lanes) beats the sector-aligned 112 + 2 (12.5 %): 4.48 vs 4.11 TB/s; the VJP needs two
        // halo columns and is closer to the memory system's limits: 4.75 (112 + 2) vs 4.59 TB/s (120 + 1)
        const bool narrow_halo = g_tuning[0] == 0 ? !vjp : g_tuning[0] == 1;
        if (narrow_halo) { g.strip_w = 120; g.halo_lane = 1; } else { g.strip_w = 112; g.halo_lane = 2; }
        g.strips = (p.W + g.strip_w - 1) / g.strip_w;
    }
    // rows per chunk: long chunks amortise the warm-up rows (2 of R + 2 in the reduce pass, 4 of R + 4 in the VJP:
    // measured on 8x2x4096^2, VJP 4.41 / 4.53 / 4.73 TB/s at R = 32 / 64 / 128, reduce best at 64), short ones expose
    // more warps on small problems: take the longest chunk that still leaves two items per resident warp.
    const int64_t per_row_items = (int64_t)g.strips * p.n_u_units * p.B;
    const int64_t want_warps = (int64_t)sm_count() * (vjp ? 16 : 24);
    int R = vjp ? 128 : 64;
    while (R > 32 && ((rows + R - 1) / R) * per_row_items / g.segs_per_warp < 2 * want_warps) R >>= 1;
    while (R > 4 && ((rows + R - 1) / R) * per_row_items / g.segs_per_warp < want_warps) R >>= 1;
    if (g_tuning[2] > 0) R = g_tuning[2];
    g.R = R;
    g.chunks = (rows + R - 1) / R;
    g.n_seg_items = (int)((int64_t)g.chunks * per_row_items);
    g.n_warp_items = (g.n_seg_items + g.segs_per_warp - 1) / g.segs_per_warp;
    g.a_plane4 = (int)((int64_t)rows * p.W / 4);
    g.a_block4 = stream_block4(g.a_plane4, (int64_t)(p.ch_a > 0 ? p.ch_a : 1) * p.B);
    g.a_blocks_per_plane = (g.a_plane4 + g.a_block4 - 1) / g.a_block4;
    g.n_a_items = g.a_blocks_per_plane * p.ch_a * p.B;
    return g;
}

// Opt a kernel instantiation in to > 48 KB of dynamic shared memory (once per instantiation AND device: the attribute is
// per device and a process may drive several GPUs) and return its occupancy.
template <typename K>
int march_occupancy(K kernel, int smem) {
    static thread_local const void* configured[16][64] = {{nullptr}};
    int dev = 0;
    cudaGetDevice(&dev);
    auto& mine = configured[dev & 15];
    bool seen = false;
    for (const void* k : mine) seen = seen || k == (const void*)kernel;
    if (!seen || dev > 15) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (auto& k : mine)
            if (!k) { k = (const void*)kernel; break; }
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    return occ;
}

inline int clamp_grid(int64_t grid, int64_t need, int64_t cap) {
    if (grid > need) grid = need;
    if (grid > cap) grid = cap;
    return (int)(grid < 1 ? 1 : grid);
}

template <typename K>
int march_grid(K kernel, const MarchGeom& g, int smem, bool paired) {
    const int occ = march_occupancy(kernel, smem);
    const int64_t need = ((int64_t)g.n_warp_items + (paired ? 0 : g.n_a_items) + kThreads / 32 - 1) / (kThreads / 32);
    return clamp_grid((int64_t)sm_count() * occ, need, kMaxPartials);
}

template <bool HAS_D, bool HAS_O, int PA, bool PS>
int launch_march_reduce_pa(const Params& p, const MarchGeom& g, double* partials, unsigned int* ticket, double* sums, int finalize,
                           double* scal, float* trace, cudaStream_t s) {
    constexpr int smem = ring_bytes(PA, false);
    auto k = heat_march_reduce_kernel<HAS_D, HAS_O, PA, PS>;
    k<<<march_grid(k, g, smem, PA != 0), kThreads, smem, s>>>(p, g, partials, ticket, sums, finalize, scal, trace);
    return check_launch("dpde_guidance_reduce (march)");
}

template <bool HAS_D, bool HAS_O>
int launch_march_reduce(const Params& p, double* partials, unsigned int* ticket, double* sums, int finalize, double* scal,
                        float* trace, cudaStream_t s) {
    const MarchGeom g = march_geometry(p, false);
    switch (pairing(p, false)) {
        case 1: return launch_march_reduce_pa<HAS_D, HAS_O, 1, false>(p, g, partials, ticket, sums, finalize, scal, trace, s);
        case 2: return launch_march_reduce_pa<HAS_D, HAS_O, 2, false>(p, g, partials, ticket, sums, finalize, scal, trace, s);
        default: return launch_march_reduce_pa<HAS_D, HAS_O, 0, false>(p, g, partials, ticket, sums, finalize, scal, trace, s);
    }
}

template <bool HAS_D, bool HAS_O, int PA, bool PS>
int launch_march_vjp_pa(const Params& p, const MarchGeom& g, const double* scal, const double* upstream, float* g_x0, float* g_dxdt,
                        cudaStream_t s) {
    constexpr int smem = ring_bytes(PA, true);
    auto k = heat_march_vjp_kernel<HAS_D, HAS_O, PA, PS>;
    k<<<march_grid(k, g, smem, PA != 0), kThreads, smem, s>>>(p, g, scal, upstream, g_x0, g_dxdt);
    return check_launch("dpde_guidance_vjp (march)");
}

template <bool HAS_D, bool HAS_O>
int launch_march_vjp(const Params& p, const double* scal, const double* upstream, float* g_x0, float* g_dxdt, cudaStream_t s) {
    const MarchGeom g = march_geometry(p, true);
    switch (pairing(p, true)) {
        case 1: return launch_march_vjp_pa<HAS_D, HAS_O, 1, false>(p, g, scal, upstream, g_x0, g_dxdt, s);
        case 2: return launch_march_vjp_pa<HAS_D, HAS_O, 2, false>(p, g, scal, upstream, g_x0, g_dxdt, s);
        default: return launch_march_vjp_pa<HAS_D, HAS_O, 0, false>(p, g, scal, upstream, g_x0, g_dxdt, s);
    }
}

// ---- LLG fast path (llg_tile.cuh) ----------------------------------------------------------------------
bool llg_eligible(const Params& p, const void* g_x0, const void* g_dxdt) {
    if (!g_fast_path || (p.kind != DPDE_PDE_LLG_RESIDUAL && p.kind != DPDE_PDE_LLG_NORM)) return false;
    if (p.x0.dtype != DPDE_F32 || p.W % 4 != 0 || p.W < 4) return false;
    if ((int64_t)p.H * p.W >= (1ll << 30) || (int64_t)p.B * p.C * (((int64_t)p.H * p.W + 1023) / 1024 + p.H + p.W) >= (1ll << 30)) return false;
    if (!al(p.x0.p, 16) || !s4(p.x0)) return false;
    if (p.dxdt.p && (!al(p.dxdt.p, 16) || !s4(p.dxdt))) return false;
    if (g_x0 && !al(g_x0, 16)) return false;
    if (g_dxdt && !al(g_dxdt, 16)) return false;
    if (p.has_a && (p.obs_a.dtype != DPDE_F32 || p.mask_a.dtype != DPDE_U8 || !al(p.obs_a.p, 16) || !al(p.mask_a.p, 4) ||
                    !s4(p.obs_a) || !s4(p.mask_a))) return false;
    if (p.has_u && (p.obs_u.dtype != DPDE_F32 || p.mask_u.dtype != DPDE_U8 || !al(p.obs_u.p, 16) || !al(p.mask_u.p, 4) ||
                    !s4(p.obs_u) || !s4(p.mask_u))) return false;
    return true;
}

struct LlgGeom {
    MarchGeom g;  // only the a-plane fields are used
    int tw, tiles_x, n_tiles, n_norm_items;
};

LlgGeom llg_geometry(const Params& p) {
    LlgGeom L{};
    const int rows = p.yhi - p.ylo;
    L.tw = p.W >= 64 ? 64 : p.W >= 32 ? 32 : 16;
    const int th = 1024 / L.tw;
    L.tiles_x = (p.W + L.tw - 1) / L.tw;
    L.n_tiles = p.B * ((rows + th - 1) / th) * L.tiles_x;
    L.g.a_plane4 = (int)((int64_t)rows * p.W / 4);
    L.g.a_block4 = stream_block4(L.g.a_plane4, p.B);
    L.g.a_blocks_per_plane = (L.g.a_plane4 + L.g.a_block4 - 1) / L.g.a_block4;
    L.g.n_a_items = L.g.a_blocks_per_plane * p.ch_a * p.B;
    L.n_norm_items = L.g.a_blocks_per_plane * p.B;
    return L;
}

template <typename K>
int llg_grid(K kernel, int smem, int64_t cta_items) {
    static thread_local const void* configured[16][16] = {{nullptr}};   // per device, as in march_grid
    int dev = 0;
    cudaGetDevice(&dev);
    auto& mine = configured[dev & 15];
    bool seen = false;
    for (const void* k : mine) seen = seen || k == (const void*)kernel;
    if (!seen || dev > 15) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (auto& k : mine)
            if (!k) { k = (const void*)kernel; break; }
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    int64_t grid = (int64_t)sm_count() * occ;
    if (grid > cta_items) grid = cta_items;
    if (grid > kMaxPartials) grid = kMaxPartials;
    return (int)(grid < 1 ? 1 : grid);
}

template <int TW>
int launch_llg_tile_reduce(const Params& p, const LlgGeom& L, double* partials, unsigned int* ticket, double* sums, int finalize,
                           double* scal, float* trace, cudaStream_t s) {
    constexpr int smem = LlgTile<TW>::smem_bytes(false);
    auto k = llg_tile_reduce_kernel<TW>;
    const int64_t items = (int64_t)L.n_tiles + (L.g.n_a_items + kThreads / 32 - 1) / (kThreads / 32);
    k<<<llg_grid(k, smem, items), kThreads, smem, s>>>(p, L.g, L.tiles_x, L.n_tiles, partials, ticket, sums, finalize, scal, trace);
    return check_launch("dpde_guidance_reduce (llg tile)");
}

template <int TW>
int launch_llg_tile_vjp(const Params& p, const LlgGeom& L, const double* scal, const double* upstream, float* g_x0, float* g_dxdt,
                        cudaStream_t s) {
    constexpr int smem = LlgTile<TW>::smem_bytes(true);
    auto k = llg_tile_vjp_kernel<TW>;
    const int64_t items = (int64_t)L.n_tiles + (L.g.n_a_items + kThreads / 32 - 1) / (kThreads / 32);
    k<<<llg_grid(k, smem, items), kThreads, smem, s>>>(p, L.g, L.tiles_x, L.n_tiles, scal, upstream, g_x0, g_dxdt);
    return check_launch("dpde_guidance_vjp (llg tile)");
}

// ---- LLG m x H_eff residual, marching kernels (llg_march.cuh) ---------------------------------------------
// Large grids only.  Measured (profiles/r2y_probe_llg_crossover*.log, reduce / VJP per launch): 8 x 6 x 2048^2 marching 0.30 / 0.55 ms against
// 0.52 / 0.94 ms of the tile kernels; 2 Mi pixels (8 x 512^2, 32 x 256^2) 32 / 56-61 us against 42 / 65-67 us; 1 Mi pixels 25-28 / 40 us
// against 30 / 40 us (a tie); config 3's shard (32 x 6 x 128^2, L2 resident) 26 / 35 us against ~20 / 23 us -- marching from 2 Mi pixels up.
inline bool llg_march_wanted(const Params& p) {
    if (p.kind != DPDE_PDE_LLG_RESIDUAL || p.W < 128 || g_tuning[6] == 1) return false;
    return g_tuning[6] == 2 || (int64_t)p.B * (p.yhi - p.ylo) * p.W >= (2ll << 20);          // key 6 = 2: marching on every size (tests)
}

LlgMarchGeom llg_march_geometry(const Params& p, bool vjp, bool lean_ok) {
    LlgMarchGeom g{};
    const int rows = p.yhi - p.ylo;
    g.strips = (p.W + kLlgStrip - 1) / kLlgStrip;
    const int64_t per_row_items = (int64_t)g.strips * p.B;
    // rows per chunk: long chunks amortise the warm-up rows, short ones give every resident warp a few items
    const int64_t want_warps = (int64_t)sm_count() * 12;
    int R = 64;
    while (R > 8 && ((rows + R - 1) / R) * per_row_items < 2 * want_warps) R >>= 1;
    if (vjp) R -= 2;                         // R + 2 row iterations in groups of the ring depth (4): 62, 30, 14, 6
    if (g_tuning[2] > 0) R = g_tuning[2];
    g.R = R;
    g.chunks = (rows + R - 1) / R;
    g.n_items = (int)((int64_t)g.chunks * per_row_items);
    g.a.a_plane4 = (int)((int64_t)rows * p.W / 4);
    g.a.a_block4 = stream_block4(g.a.a_plane4, p.B);
    g.a.a_blocks_per_plane = (g.a.a_plane4 + g.a.a_block4 - 1) / g.a.a_block4;
    g.a.n_a_items = g.a.a_blocks_per_plane * p.ch_a * p.B;
    // interior rectangle (lean loop): strips whose 32 lanes all hold grid columns away from the edge columns, full chunks whose
    // rows + stencil margin need neither reflection nor clamping, row iterations in whole groups of the ring depth
    const int n_it = vjp ? R + 2 : R, margin = vjp ? 2 : 1;
    if (lean_ok && g_tuning[5] == 0 && n_it % kLR == 0 && n_it >= (vjp ? 2 : 1) * kLR) {
        int s_lo = -1, s_hi = -2, c_lo = -1, c_hi = -2;
        for (int st = 0; st < g.strips; ++st) {
            const int first = st * kLlgStrip - 2, last = first + 2 * 31;
            if (first >= 2 && last + 2 < p.W) { if (s_lo < 0) s_lo = st; s_hi = st; }
        }
        for (int c = 0; c < g.chunks; ++c) {
            const int ys = p.ylo + c * R;
            if (ys + R <= p.yhi && rows_inside(p, ys - margin, ys + R + margin - 1)) { if (c_lo < 0) c_lo = c; c_hi = c; }
        }
        if (s_lo >= 0 && c_lo >= 0) {
            g.s_lo = s_lo; g.s_hi = s_hi; g.c_lo = c_lo; g.c_hi = c_hi;
            g.n_int_items = p.B * (s_hi - s_lo + 1) * (c_hi - c_lo + 1);
        }
    }
    return g;
}

template <typename K>
int llg_march_grid(K kernel, int64_t warp_items, int smem) {
    static thread_local const void* configured[16][8] = {{nullptr}};   // > 48 KB opt-in, per device, as in march_occupancy
    int dev = 0;
    cudaGetDevice(&dev);
    auto& mine = configured[dev & 15];
    bool seen = false;
    for (const void* k : mine) seen = seen || k == (const void*)kernel;
    if (!seen || dev > 15) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (auto& k : mine)
            if (!k) { k = (const void*)kernel; break; }
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kLlgThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    const int64_t need = (warp_items + kLlgThreads / 32 - 1) / (kLlgThreads / 32);
    return clamp_grid((int64_t)sm_count() * occ, need, kMaxPartials);
}

// ---- TMA tensor maps for the lean interior items of the LLG marching kernels ---------------------------------
// One map per tensor: dims (W, H, planes, batch) of fp32, box 68 columns x 1 row x 3 planes = one warp's row of the three
// components (+ the four columns that align the box start to 16 bytes), 816 bytes per copy.  The encoder is a driver entry point (no link against libcuda).
inline PFN_cuTensorMapEncodeTiled tma_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(f);
    }();
    return fn;
}

// planes = channels of the tensor, first_plane = channel the 3-plane box starts at (coordinate passed by the kernel)
inline bool tma_map_f32(CUtensorMap* map, const View& v, int B, int planes, int H, int W) {
    auto enc = tma_encoder();
    if (!enc || !v.p || v.dtype != DPDE_F32 || W < 64 || planes < 3 || v.sc <= 0) return false;
    if (!al(v.p, 16) || (W % 4) || (v.sc % 4) || (v.sb % 4)) return false;
    const bool bcast = v.sb == 0;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)(bcast ? 1 : B)};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)v.sc * 4, (cuuint64_t)(bcast ? (int64_t)planes * v.sc : v.sb) * 4};
    if (strides[2] == 0 || strides[2] % 16) return false;
    cuuint32_t box[4] = {kLlgBoxCols, kLlgBoxRows, 3, 1}, estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(v.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// One zeroed work-queue counter for a launch on stream `s`: a slot of a small device array, cleared by a stream-ordered memset.  Slots are
// handed out round-robin; a slot is reused after kQueueSlots further launches of this process.
constexpr int kQueueSlots = 256;
__device__ unsigned int g_queue_slots[kQueueSlots];
inline unsigned int* work_queue_counter(cudaStream_t s) {
    static std::atomic<unsigned> next{0};
    static thread_local unsigned int* base[16] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    unsigned int*& b = base[dev & 15];
    if (!b || dev > 15) {
        void* ptr = nullptr;
        if (cudaGetSymbolAddress(&ptr, g_queue_slots) != cudaSuccess) return nullptr;
        b = reinterpret_cast<unsigned int*>(ptr);
    }
    unsigned int* slot = b + (next.fetch_add(1) % kQueueSlots);
    if (cudaMemsetAsync(slot, 0, sizeof(unsigned int), s) != cudaSuccess) return nullptr;
    return slot;
}

template <bool HAS_D, bool HAS_O>
bool llg_tma_maps(const Params& p, LlgTmaMaps* maps) {
    bool ok = tma_map_f32(&maps->m, p.x0, p.B, p.C, p.H, p.W);
    if (HAS_D) ok = ok && tma_map_f32(&maps->d, p.dxdt, p.B, p.C, p.H, p.W);
    if (HAS_O) ok = ok && tma_map_f32(&maps->o, p.obs_u, p.B, 3, p.H, p.W);
    maps->o_bcast = p.obs_u.sb == 0;
    return ok;
}

template <bool HAS_D, bool HAS_O>
int launch_llg_march_reduce(const Params& p, double* partials, unsigned int* ticket, double* sums, int finalize, double* scal, float* trace,
                            cudaStream_t s) {
    const LlgMarchGeom g = llg_march_geometry(p, false, true);
    LlgTmaMaps none{};
    // lean interior items fed by TMA (tuning key 7 = 1 switches it off: tests compare the two feeds bit for bit); falls back to the
    // cp.async ring when a tensor map cannot be encoded (driver entry point missing, strides not multiples of 16 bytes)
    if (g_tuning[7] != 1 && g.n_int_items > 0 && llg_tma_maps<HAS_D, HAS_O>(p, &none)) {
        auto k = llg_march_reduce_kernel<HAS_D, HAS_O, true>;
        const int smem = llg_tma_smem_bytes();
        k<<<llg_march_grid(k, (int64_t)g.n_items + g.a.n_a_items, smem), kLlgThreads, smem, s>>>(p, g, none, partials, ticket, sums, finalize, scal, trace);
        return check_launch("dpde_guidance_reduce (llg march, tma)");
    }
    auto k = llg_march_reduce_kernel<HAS_D, HAS_O, false>;
    k<<<llg_march_grid(k, (int64_t)g.n_items + g.a.n_a_items, llg_smem_bytes()), kLlgThreads, llg_smem_bytes(), s>>>(p, g, none, partials, ticket, sums, finalize, scal, trace);
    return check_launch("dpde_guidance_reduce (llg march)");
}

template <bool HAS_D, bool HAS_O>
int launch_llg_march_vjp(const Params& p, const double* scal, const double* upstream, float* g_x0, float* g_dxdt, cudaStream_t s) {
    const LlgMarchGeom g = llg_march_geometry(p, true, g_dxdt == nullptr);
    auto k = llg_march_vjp_kernel<HAS_D, HAS_O>;
    LlgTmaMaps maps{};
    // three CTAs per SM: TMA-fed lean items without register windows + light general items, dynamic longest-first queue (llg_vjp_lean3_kernel);
    // tuning key 7 = 1 keeps the two-CTA kernel with the cp.async ring (tests compare the two bit for bit), which is also the fallback when
    // d / d dmdt is wanted (no interior rectangle) or a tensor map cannot be encoded
    if (g_tuning[7] != 1 && g.n_int_items > 0 && (g.R + 2) >= 8 && llg_tma_maps<HAS_D, HAS_O>(p, &maps)) {
        auto k3 = llg_vjp_lean3_kernel<HAS_D, HAS_O>;
        unsigned int* queue = work_queue_counter(s);
        if (!queue) return fail(DPDE_ERR_CUDA, "dpde_guidance_vjp: work-queue counter unavailable");
        k3<<<llg_march_grid(k3, (int64_t)g.n_items + g.a.n_a_items, llg_v3_smem_bytes()), kLlgThreads, llg_v3_smem_bytes(), s>>>(p, g, maps, scal, upstream, g_x0, g_dxdt, queue);
        return check_launch("dpde_guidance_vjp (llg march, three CTAs)");
    }
    k<<<llg_march_grid(k, (int64_t)g.n_items + g.a.n_a_items, llg_smem_bytes()), kLlgThreads, llg_smem_bytes(), s>>>(p, g, scal, upstream, g_x0, g_dxdt);
    return check_launch("dpde_guidance_vjp (llg march)");
}

int launch_llg_reduce(const Params& p, double* partials, unsigned int* ticket, double* sums, int finalize, double* scal,
                      float* trace, cudaStream_t s) {
    if (llg_march_wanted(p)) {
        const bool d = p.dxdt.p != nullptr, o = p.has_u != 0;
        return d ? (o ? launch_llg_march_reduce<true, true>(p, partials, ticket, sums, finalize, scal, trace, s)
                      : launch_llg_march_reduce<true, false>(p, partials, ticket, sums, finalize, scal, trace, s))
                 : (o ? launch_llg_march_reduce<false, true>(p, partials, ticket, sums, finalize, scal, trace, s)
                      : launch_llg_march_reduce<false, false>(p, partials, ticket, sums, finalize, scal, trace, s));
    }
    const LlgGeom L = llg_geometry(p);
    if (p.kind == DPDE_PDE_LLG_NORM) {
        const int64_t items = ((int64_t)L.n_norm_items + L.g.n_a_items + kThreads / 32 - 1) / (kThreads / 32);
        auto go = [&](auto k) { k<<<llg_grid(k, 0, items), kThreads, 0, s>>>(p, L.g, L.n_norm_items, partials, ticket, sums, finalize, scal, trace); };
        if (p.has_u) go(llg_norm_reduce_kernel<true, 1>); else go(llg_norm_reduce_kernel<false, 1>);
        return check_launch("dpde_guidance_reduce (llg norm)");
    }
    switch (L.tw) {
        case 64: return launch_llg_tile_reduce<64>(p, L, partials, ticket, sums, finalize, scal, trace, s);
        case 32: return launch_llg_tile_reduce<32>(p, L, partials, ticket, sums, finalize, scal, trace, s);
        default: return launch_llg_tile_reduce<16>(p, L, partials, ticket, sums, finalize, scal, trace, s);
    }
}

int launch_llg_vjp(const Params& p, const double* scal, const double* upstream, float* g_x0, float* g_dxdt, cudaStream_t s) {
    if (llg_march_wanted(p)) {
        const bool d = p.dxdt.p != nullptr, o = p.has_u != 0;
        return d ? (o ? launch_llg_march_vjp<true, true>(p, scal, upstream, g_x0, g_dxdt, s) : launch_llg_march_vjp<true, false>(p, scal, upstream, g_x0, g_dxdt, s))
                 : (o ? launch_llg_march_vjp<false, true>(p, scal, upstream, g_x0, g_dxdt, s) : launch_llg_march_vjp<false, false>(p, scal, upstream, g_x0, g_dxdt, s));
    }
    const LlgGeom L = llg_geometry(p);
    if (p.kind == DPDE_PDE_LLG_NORM) {
        const int64_t items = ((int64_t)L.n_norm_items + L.g.n_a_items + kThreads / 32 - 1) / (kThreads / 32);
        auto go = [&](auto k) { k<<<llg_grid(k, norm_ring_bytes(), items), kThreads, norm_ring_bytes(), s>>>(p, L.g, L.n_norm_items, scal, upstream, g_x0, g_dxdt); };
        if (p.has_u) go(llg_norm_vjp_kernel<true>); else go(llg_norm_vjp_kernel<false>);
        return check_launch("dpde_guidance_vjp (llg norm)");
    }
    switch (L.tw) {
        case 64: return launch_llg_tile_vjp<64>(p, L, scal, upstream, g_x0, g_dxdt, s);
        case 32: return launch_llg_tile_vjp<32>(p, L, scal, upstream, g_x0, g_dxdt, s);
        default: return launch_llg_tile_vjp<16>(p, L, scal, upstream, g_x0, g_dxdt, s);
    }
}

// ---- per-sample heat residual (training loss) on the marching kernels ----------------------------------
Params per_sample_params(const void* u, const void* dudt, int dtype, int B, int Cu, int H, int W, int64_t sb_u, int64_t sc_u,
                         int64_t sb_d, int64_t sc_d, const double* alpha, double dx) {
    Params p{};
    p.B = B; p.C = Cu; p.ch_a = 0; p.H = H; p.W = W; p.kind = DPDE_PDE_HEAT; p.has_a = 0; p.has_u = 0;
    p.Hg = H; p.yg0 = 0; p.ylo = 0; p.yhi = H; p.n_u_units = Cu; p.units_per_sample = Cu;
    p.x0 = View{u, dtype, sb_u, sc_u};
    p.dxdt = View{dudt, dtype, sb_d, sc_d};
    p.coef = alpha;
    p.inv_dx2 = 1.0 / (dx * dx);
    p.hw = (double)H * (double)W;
    return p;
}

// upper bound on the row-segment items of march_geometry for (B, Cu, H, W): chunks <= ceil(H/4), strips <= ceil(W/112)
inline int64_t per_sample_item_bound(int B, int Cu, int H, int W) {
    return (int64_t)B * Cu * ((H + 3) / 4) * (W <= 128 ? 1 : (W + 111) / 112);
}

template <bool HAS_D>
int launch_march_per_sample_reduce(const Params& p, double* partials, double* out, cudaStream_t s) {
    const MarchGeom g = march_geometry(p, false);
    if (int rc = launch_march_reduce_pa<HAS_D, false, 0, true>(p, g, partials, nullptr, nullptr, 0, nullptr, nullptr, s)) return rc;
    per_sample_items_kernel<<<(p.B + 3) / 4, 128, 0, s>>>(partials, g.n_seg_items / p.B, p.B, out);
    return check_launch("dpde_heat_residual_sq (march)");
}

template <bool HAS_D>
int launch_march_per_sample_vjp(const Params& p, const double* upstream, float* g_u, float* g_dudt, cudaStream_t s) {
    const MarchGeom g = march_geometry(p, true);
    return launch_march_vjp_pa<HAS_D, false, 0, true>(p, g, nullptr, upstream, g_u, g_dudt, s);
}

View to_view(const dpde_view& v) { return View{v.ptr, v.dtype, v.stride_b, v.stride_c}; }

int validate_and_fill(const dpde_guidance_desc* d, int tile_pix, Params& p, const char* who) {
    if (!d) return fail(DPDE_ERR_INVALID, "%s: desc is NULL", who);
    if (d->B < 1 || d->C < 1 || d->ch_a < 0 || d->ch_a > d->C) return fail(DPDE_ERR_INVALID, "%s: bad B/C/ch_a", who);
    if (d->H < 2 || d->W < 2) return fail(DPDE_ERR_INVALID, "%s: H and W must be >= 2 (reflect padding)", who);
    if (!d->x0.ptr) return fail(DPDE_ERR_INVALID, "%s: x0 is NULL", who);
    if (d->x0.dtype != DPDE_F32 && d->x0.dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "%s: x0 must be f32/f64", who);
    if (d->dxdt.ptr && d->dxdt.dtype != d->x0.dtype) return fail(DPDE_ERR_INVALID, "%s: dxdt dtype != x0 dtype", who);
    const int cu = d->C - d->ch_a;
    switch (d->pde_kind) {
        case DPDE_PDE_NONE: break;
        case DPDE_PDE_HEAT:
            if (cu < 1) return fail(DPDE_ERR_INVALID, "%s: heat residual needs at least one u channel", who);
            if (!d->sample_coef) return fail(DPDE_ERR_INVALID, "%s: heat residual needs sample_coef = alpha (B,)", who);
            break;
        case DPDE_PDE_LLG_RESIDUAL:
            if (!d->sample_coef) return fail(DPDE_ERR_INVALID, "%s: LLG residual needs sample_coef = h_ext (B,3)", who);
            /* fallthrough */
        case DPDE_PDE_LLG_NORM:
            if (cu != 3) return fail(DPDE_ERR_INVALID, "%s: LLG kinds need exactly 3 magnetisation channels, got %d", who, cu);
            break;
        default: return fail(DPDE_ERR_INVALID, "%s: unknown pde_kind %d", who, d->pde_kind);
    }
    if (d->pde_kind != DPDE_PDE_NONE && d->pde_kind != DPDE_PDE_LLG_NORM && !(d->dx > 0.0))
        return fail(DPDE_ERR_INVALID, "%s: dx must be > 0", who);
    if (d->has_a && (d->ch_a < 1 || !d->obs_a.ptr || !d->mask_a.ptr)) return fail(DPDE_ERR_INVALID, "%s: has_a without obs_a/mask_a", who);
    if (d->has_u && (cu < 1 || !d->obs_u.ptr || !d->mask_u.ptr)) return fail(DPDE_ERR_INVALID, "%s: has_u without obs_u/mask_u", who);

    p.B = d->B; p.C = d->C; p.ch_a = d->ch_a; p.H = d->H; p.W = d->W; p.kind = d->pde_kind;
    p.mb_world = 0; p.mb_rank = 0; p.mb_epoch = 0;
    for (auto& b : p.mb_box) b = nullptr;
    if (d->slab_H_global > 0) {
        const int need = (d->pde_kind == DPDE_PDE_HEAT || d->pde_kind == DPDE_PDE_LLG_RESIDUAL) ? 2 : 0;
        if (d->slab_halo < need) return fail(DPDE_ERR_INVALID, "%s: slab_halo %d < %d needed by this residual", who, d->slab_halo, need);
        if (d->H - 2 * d->slab_halo < 1 || d->slab_row0 < 0 || d->slab_row0 + (d->H - 2 * d->slab_halo) > d->slab_H_global)
            return fail(DPDE_ERR_INVALID, "%s: slab rows [%d, %d) outside the global grid of %d rows", who, d->slab_row0,
                        d->slab_row0 + d->H - 2 * d->slab_halo, d->slab_H_global);
        p.Hg = d->slab_H_global; p.ylo = d->slab_halo; p.yhi = d->H - d->slab_halo; p.yg0 = d->slab_row0 - d->slab_halo;
    } else {
        p.Hg = d->H; p.ylo = 0; p.yhi = d->H; p.yg0 = 0;
    }
    p.has_a = d->has_a != 0; p.has_u = d->has_u != 0;
    const bool llg = d->pde_kind == DPDE_PDE_LLG_NORM || d->pde_kind == DPDE_PDE_LLG_RESIDUAL;
    p.n_u_units = llg ? 1 : cu;
    p.units_per_sample = d->ch_a + p.n_u_units;
    p.tile_w = d->W > 64 ? 128 : d->W > 32 ? 64 : d->W > 16 ? 32 : 16;
    p.tile_h = tile_pix / p.tile_w;
    p.tw_log2 = p.tile_w == 128 ? 7 : p.tile_w == 64 ? 6 : p.tile_w == 32 ? 5 : 4;
    p.tiles_x = (d->W + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (p.yhi - p.ylo + p.tile_h - 1) / p.tile_h;
    p.tiles_per_plane = (int64_t)p.tiles_x * p.tiles_y;
    p.n_tiles = p.tiles_per_plane * p.units_per_sample * d->B;
    p.x0 = to_view(d->x0); p.dxdt = to_view(d->dxdt);
    p.obs_a = to_view(d->obs_a); p.mask_a = to_view(d->mask_a);
    p.obs_u = to_view(d->obs_u); p.mask_u = to_view(d->mask_u);
    p.coef = d->sample_coef;
    p.inv_dx2 = d->dx > 0.0 ? 1.0 / (d->dx * d->dx) : 0.0;
    p.w_a = d->w_a; p.w_u = d->w_u; p.w_pde = d->w_pde;
    p.hw = (double)p.Hg * (double)d->W;
    p.gamma = d->gamma; p.alpha = d->alpha; p.c_ex = d->c_ex; p.c_an = d->c_an; p.tau = d->tau;
    p.e[0] = d->easy_axis[0]; p.e[1] = d->easy_axis[1]; p.e[2] = d->easy_axis[2];
    return DPDE_OK;
}

template <typename K>
int persistent_grid(K kernel, int64_t n_tiles) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ < 1) occ = 1;
    int64_t g = (int64_t)sm_count() * occ;
    if (g > n_tiles) g = n_tiles;
    if (g > kMaxPartials) g = kMaxPartials;
    return (int)(g < 1 ? 1 : g);
}

constexpr int kTileHeat = 2048, kTileLLG = 1024;
inline int tile_pix_for(int kind) { return kind == DPDE_PDE_LLG_RESIDUAL ? kTileLLG : kTileHeat; }

template <typename T>
int launch_reduce(const Params& p, double* partials, unsigned int* ticket, double* sums, int finalize, double* scal,
                  float* trace, cudaStream_t s) {
#define DPDE_RED(KIND, TP)                                                                              \
    {                                                                                                   \
        auto k = guidance_reduce_kernel<T, KIND, TP>;                                                   \
        k<<<persistent_grid(k, p.n_tiles), kThreads, 0, s>>>(p, partials, ticket, sums, finalize, scal, trace); \
    }
    switch (p.kind) {
        case DPDE_PDE_NONE: DPDE_RED(DPDE_PDE_NONE, kTileHeat) break;
        case DPDE_PDE_HEAT: DPDE_RED(DPDE_PDE_HEAT, kTileHeat) break;
        case DPDE_PDE_LLG_NORM: DPDE_RED(DPDE_PDE_LLG_NORM, kTileHeat) break;
        default: DPDE_RED(DPDE_PDE_LLG_RESIDUAL, kTileLLG) break;
    }
#undef DPDE_RED
    return check_launch("dpde_guidance_reduce");
}

template <typename T>
int launch_vjp(const Params& p, const double* scal, const double* upstream, void* g, void* gd, cudaStream_t s) {
#define DPDE_VJP(KIND, TP)                                                                   \
    {                                                                                        \
        auto k = guidance_vjp_kernel<T, KIND, TP>;                                           \
        k<<<persistent_grid(k, p.n_tiles), kThreads, 0, s>>>(p, scal, upstream, (T*)g, (T*)gd); \
    }
    switch (p.kind) {
        case DPDE_PDE_NONE: DPDE_VJP(DPDE_PDE_NONE, kTileHeat) break;
        case DPDE_PDE_HEAT: DPDE_VJP(DPDE_PDE_HEAT, kTileHeat) break;
        case DPDE_PDE_LLG_NORM: DPDE_VJP(DPDE_PDE_LLG_NORM, kTileHeat) break;
        default: DPDE_VJP(DPDE_PDE_LLG_RESIDUAL, kTileLLG) break;
    }
#undef DPDE_VJP
    return check_launch("dpde_guidance_vjp");
}

}  // namespace
}  // namespace dpde

using namespace dpde;

extern "C" {

int dpde_set_fast_path(int enable) {
    return g_fast_path.exchange(enable != 0) ? 1 : 0;
}

int dpde_set_tuning(int key, int value) {
    if (key < 0 || key >= 8) return fail(DPDE_ERR_INVALID, "dpde_set_tuning: key must be in [0, 8)");
    g_tuning[key] = value;
    return DPDE_OK;
}

size_t dpde_guidance_workspace_bytes(void) { return (size_t)(3 * kMaxPartials + 2) * sizeof(double); }

namespace {
int fill_mailbox(const dpde_mailbox* mbox, Params& p, const char* who) {
    if (!mbox) return fail(DPDE_ERR_INVALID, "%s: mailbox is NULL", who);
    if (mbox->world < 1 || mbox->world > DPDE_MAX_RANKS || mbox->rank < 0 || mbox->rank >= mbox->world)
        return fail(DPDE_ERR_INVALID, "%s: need 1 <= world <= %d and 0 <= rank < world", who, DPDE_MAX_RANKS);
    if (mbox->epoch == 0) return fail(DPDE_ERR_INVALID, "%s: epoch must be >= 1 (mailboxes start zeroed)", who);
    for (int r = 0; r < mbox->world; ++r)
        if (!mbox->boxes[r]) return fail(DPDE_ERR_INVALID, "%s: boxes[%d] is NULL", who, r);
    p.mb_world = mbox->world; p.mb_rank = mbox->rank; p.mb_epoch = mbox->epoch;
    for (int r = 0; r < mbox->world; ++r) p.mb_box[r] = mbox->boxes[r];
    return DPDE_OK;
}

int reduce_impl(const dpde_guidance_desc* desc, void* workspace, double* sums, int finalize, double* scalars, float* trace_row,
                const dpde_mailbox* mbox, dpde_stream_t stream, const char* who) {
    Params p;
    if (int rc = validate_and_fill(desc, desc ? tile_pix_for(desc->pde_kind) : kTileHeat, p, who)) return rc;
    if (!workspace || !sums) return fail(DPDE_ERR_INVALID, "%s: workspace/sums is NULL", who);
    if (finalize && !scalars) return fail(DPDE_ERR_INVALID, "%s: finalize needs scalars", who);
    if (mbox)
        if (int rc = fill_mailbox(mbox, p, who)) return rc;
    double* partials = reinterpret_cast<double*>(workspace);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + 3 * kMaxPartials);
    cudaStream_t s = (cudaStream_t)stream;
    if (march_eligible(p, nullptr, nullptr)) {
        const bool d = p.dxdt.p != nullptr, o = p.has_u != 0;
        return d ? (o ? launch_march_reduce<true, true>(p, partials, ticket, sums, finalize, scalars, trace_row, s)
                      : launch_march_reduce<true, false>(p, partials, ticket, sums, finalize, scalars, trace_row, s))
                 : (o ? launch_march_reduce<false, true>(p, partials, ticket, sums, finalize, scalars, trace_row, s)
                      : launch_march_reduce<false, false>(p, partials, ticket, sums, finalize, scalars, trace_row, s));
    }
    if (llg_eligible(p, nullptr, nullptr)) return launch_llg_reduce(p, partials, ticket, sums, finalize, scalars, trace_row, s);
    return p.x0.dtype == DPDE_F32 ? launch_reduce<float>(p, partials, ticket, sums, finalize, scalars, trace_row, s)
                                  : launch_reduce<double>(p, partials, ticket, sums, finalize, scalars, trace_row, s);
}
}  // namespace

int dpde_guidance_reduce(const dpde_guidance_desc* desc, void* workspace, double* sums, int finalize, double* scalars,
                         float* trace_row, dpde_stream_t stream) {
    return reduce_impl(desc, workspace, sums, finalize, scalars, trace_row, nullptr, stream, "dpde_guidance_reduce");
}

int dpde_guidance_reduce_post(const dpde_guidance_desc* desc, void* workspace, double* sums, const dpde_mailbox* mbox,
                              dpde_stream_t stream) {
    if (!mbox) return fail(DPDE_ERR_INVALID, "dpde_guidance_reduce_post: mailbox is NULL");
    return reduce_impl(desc, workspace, sums, 0, nullptr, nullptr, mbox, stream, "dpde_guidance_reduce_post");
}

int dpde_mailbox_wait_finalize(const dpde_guidance_desc* desc, const dpde_mailbox* mbox, double timeout_s, int32_t* status,
                               double* sums, double* scalars, float* trace_row, dpde_stream_t stream) {
    const char* who = "dpde_mailbox_wait_finalize";
    Params p;
    if (int rc = validate_and_fill(desc, desc ? tile_pix_for(desc->pde_kind) : kTileHeat, p, who)) return rc;
    if (!sums || !scalars) return fail(DPDE_ERR_INVALID, "%s: sums/scalars is NULL", who);
    if (!(timeout_s > 0.0)) return fail(DPDE_ERR_INVALID, "%s: timeout_s must be > 0", who);
    if (int rc = fill_mailbox(mbox, p, who)) return rc;
    mailbox_wait_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, (unsigned long long)(timeout_s * 1e9), status, sums, scalars, trace_row);
    return check_launch(who);
}

int dpde_guidance_finalize(const dpde_guidance_desc* desc, const double* sums, double* scalars, float* trace_row,
                           dpde_stream_t stream) {
    Params p;
    if (int rc = validate_and_fill(desc, desc ? tile_pix_for(desc->pde_kind) : kTileHeat, p, "dpde_guidance_finalize")) return rc;
    if (!sums || !scalars) return fail(DPDE_ERR_INVALID, "dpde_guidance_finalize: sums/scalars is NULL");
    finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, sums, scalars, trace_row);
    return check_launch("dpde_guidance_finalize");
}

int dpde_guidance_vjp(const dpde_guidance_desc* desc, const double* scalars, const double* upstream, void* g_x0,
                      void* g_dxdt, dpde_stream_t stream) {
    Params p;
    if (int rc = validate_and_fill(desc, desc ? tile_pix_for(desc->pde_kind) : kTileHeat, p, "dpde_guidance_vjp")) return rc;
    if (!scalars || !g_x0) return fail(DPDE_ERR_INVALID, "dpde_guidance_vjp: scalars/g_x0 is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (march_eligible(p, g_x0, g_dxdt)) {
        const bool d = p.dxdt.p != nullptr, o = p.has_u != 0;
        float *gx = (float*)g_x0, *gd = (float*)g_dxdt;
        return d ? (o ? launch_march_vjp<true, true>(p, scalars, upstream, gx, gd, s) : launch_march_vjp<true, false>(p, scalars, upstream, gx, gd, s))
                 : (o ? launch_march_vjp<false, true>(p, scalars, upstream, gx, gd, s) : launch_march_vjp<false, false>(p, scalars, upstream, gx, gd, s));
    }
    if (llg_eligible(p, g_x0, g_dxdt)) return launch_llg_vjp(p, scalars, upstream, (float*)g_x0, (float*)g_dxdt, s);
    return p.x0.dtype == DPDE_F32 ? launch_vjp<float>(p, scalars, upstream, g_x0, g_dxdt, s)
                                  : launch_vjp<double>(p, scalars, upstream, g_x0, g_dxdt, s);
}

int dpde_laplacian(const void* u, void* out, int32_t dtype, int64_t planes, int32_t H, int32_t W, int64_t plane_stride_in,
                   double dx, int32_t adjoint, dpde_stream_t stream) {
    if (!u || !out) return fail(DPDE_ERR_INVALID, "dpde_laplacian: null pointer");
    if (planes < 0 || H < 2 || W < 2) return fail(DPDE_ERR_INVALID, "dpde_laplacian: need planes >= 0, H, W >= 2");
    if (!(dx > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_laplacian: dx must be > 0");
    if (dtype != DPDE_F32 && dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "dpde_laplacian: dtype must be f32/f64");
    if (planes == 0) return DPDE_OK;
    const int64_t total = planes * (int64_t)H * W;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const double inv = 1.0 / (dx * dx);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == DPDE_F32)
        laplacian_kernel<float><<<(int)blocks, kThreads, 0, s>>>((const float*)u, (float*)out, planes, H, W, plane_stride_in, inv, adjoint);
    else
        laplacian_kernel<double><<<(int)blocks, kThreads, 0, s>>>((const double*)u, (double*)out, planes, H, W, plane_stride_in, inv, adjoint);
    return check_launch("dpde_laplacian");
}

size_t dpde_heat_residual_sq_workspace_bytes(int32_t B, int32_t Cu, int32_t H, int32_t W) {
    if (B < 1 || Cu < 1 || H < 1 || W < 1) return 0;
    const int64_t a = (int64_t)B * kPerSampleBlocks, b = per_sample_item_bound(B, Cu, H, W);
    return (size_t)(a > b ? a : b) * sizeof(double);
}

int dpde_heat_residual_sq(const void* u, const void* dudt, int32_t dtype, int32_t B, int32_t Cu, int32_t H, int32_t W, int64_t sb_u,
                          int64_t sc_u, int64_t sb_d, int64_t sc_d, const double* alpha, double dx, void* workspace, double* out,
                          dpde_stream_t stream) {
    if (!u || !alpha || !workspace || !out) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq: null pointer");
    if (B < 0 || Cu < 1 || H < 2 || W < 2) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq: need B >= 0, Cu >= 1, H, W >= 2");
    if (!(dx > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq: dx must be > 0");
    if (dtype != DPDE_F32 && dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "dpde_heat_residual_sq: dtype must be f32/f64");
    if (B == 0) return DPDE_OK;
    if (B > 65535) return fail(DPDE_ERR_UNSUPPORTED, "dpde_heat_residual_sq: at most 65535 samples per call");
    const double inv = 1.0 / (dx * dx);
    cudaStream_t s = (cudaStream_t)stream;
    {
        const Params p = per_sample_params(u, dudt, dtype, B, Cu, H, W, sb_u, sc_u, sb_d, sc_d, alpha, dx);
        if (march_eligible(p, nullptr, nullptr))
            return dudt ? launch_march_per_sample_reduce<true>(p, reinterpret_cast<double*>(workspace), out, s)
                        : launch_march_per_sample_reduce<false>(p, reinterpret_cast<double*>(workspace), out, s);
    }
    int64_t nblk = ((int64_t)Cu * H * W + 8 * kThreads - 1) / (8 * kThreads);   // ~8 pixels per thread
    if (nblk > kPerSampleBlocks) nblk = kPerSampleBlocks;
    const dim3 grid((unsigned)nblk, (unsigned)B);
    double* partials = reinterpret_cast<double*>(workspace);
    if (dtype == DPDE_F32)
        heat_residual_sq_kernel<float><<<grid, kThreads, 0, s>>>((const float*)u, (const float*)dudt, sb_u, sc_u, sb_d, sc_d, alpha, Cu, H, W, inv, partials);
    else
        heat_residual_sq_kernel<double><<<grid, kThreads, 0, s>>>((const double*)u, (const double*)dudt, sb_u, sc_u, sb_d, sc_d, alpha, Cu, H, W, inv, partials);
    per_sample_combine_kernel<<<(B + 127) / 128, 128, 0, s>>>(partials, (int)nblk, B, out);
    return check_launch("dpde_heat_residual_sq");
}

int dpde_heat_residual_sq_vjp(const void* u, const void* dudt, int32_t dtype, int32_t B, int32_t Cu, int32_t H, int32_t W, int64_t sb_u,
                              int64_t sc_u, int64_t sb_d, int64_t sc_d, const double* alpha, double dx, const double* upstream,
                              void* g_u, void* g_dudt, dpde_stream_t stream) {
    if (!u || !alpha || !upstream || !g_u) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq_vjp: null pointer");
    if (B < 0 || Cu < 1 || H < 2 || W < 2) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq_vjp: need B >= 0, Cu >= 1, H, W >= 2");
    if (!(dx > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_heat_residual_sq_vjp: dx must be > 0");
    if (dtype != DPDE_F32 && dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "dpde_heat_residual_sq_vjp: dtype must be f32/f64");
    if (B == 0) return DPDE_OK;
    const int64_t total = (int64_t)B * Cu * H * W;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const double inv = 1.0 / (dx * dx);
    cudaStream_t s = (cudaStream_t)stream;
    {
        const Params p = per_sample_params(u, dudt, dtype, B, Cu, H, W, sb_u, sc_u, sb_d, sc_d, alpha, dx);
        if (march_eligible(p, g_u, g_dudt))
            return dudt ? launch_march_per_sample_vjp<true>(p, upstream, (float*)g_u, (float*)g_dudt, s)
                        : launch_march_per_sample_vjp<false>(p, upstream, (float*)g_u, (float*)g_dudt, s);
    }
    if (dtype == DPDE_F32)
        heat_residual_sq_vjp_kernel<float><<<(int)blocks, kThreads, 0, s>>>((const float*)u, (const float*)dudt, sb_u, sc_u, sb_d, sc_d, alpha, upstream, B, Cu, H, W, inv, (float*)g_u, (float*)g_dudt);
    else
        heat_residual_sq_vjp_kernel<double><<<(int)blocks, kThreads, 0, s>>>((const double*)u, (const double*)dudt, sb_u, sc_u, sb_d, sc_d, alpha, upstream, B, Cu, H, W, inv, (double*)g_u, (double*)g_dudt);
    return check_launch("dpde_heat_residual_sq_vjp");
}

}  // extern "C"
