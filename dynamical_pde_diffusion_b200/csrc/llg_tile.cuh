// LLG guidance, fast path: convert-once shared-memory tiles (m x H_eff residual) and 128-bit streaming (soft norm).
//
// Included by guidance.cu after heat_march.cuh (inside namespace dpde::{anonymous}); uses Params / adj_w / cross3 /
// block_sum3 / finalize_scalars / row_offset / ldg4 / u8_to_double / MarchGeom's a-plane fields.
//
// Why this shape (ncu of the generic tile kernel, profiles/r1g_*): per pixel it issued 15 scalar global loads and
// 15 F2F.F64.F32 for the three 5-point stencils alone (every magnetisation value was fetched and widened five
// times), 26 % issue utilisation, 9 % of the fp64 pipe, 9-12 % of HBM bandwidth.  The m x H_eff VJP needs ~160
// fp64 instructions per pixel: at 64 lanes/clk/SM that is the same time as its 60 B/pixel of HBM traffic, so the
// kernel has to keep BOTH the fp64 pipe and the memory system busy and must not waste issue slots on loads.
//   * A CTA owns a TW x TH = 1024-pixel tile (TW in {64, 32, 16} chosen from W, so the reference's 64x16 film keeps
//     all lanes busy).  The three magnetisation planes of the tile plus a 2-pixel halo are fetched ONCE with
//     128-bit loads, widened ONCE to fp64 and parked in shared memory (32 KB); rows outside the grid are REFLECTED
//     by the loader (the reference's padding, sample.py:126-133), column reflection is a register select.
//   * Phase A: a thread evaluates TWO adjacent pixels at a time (LDS.128 centre / up / down, LDS.64 left / right):
//     H_eff, a = m x H, the residual r, and for the VJP the field gradient G_H, which goes to a second fp64 tile
//     (29 KB, 1-pixel ring computed by the first warps of the CTA), while the pointwise part of the gradient stays
//     in registers.
//   * Phase B (VJP): the transposed 5-point stencil of G_H from shared memory, + the pointwise part, rounded once
//     to fp32 and stored with 64-bit coalesced stores.  Nothing but g is written to HBM.
//   * a-planes (masked difference only) are streamed by the same CTAs with 128-bit loads after their tiles.
// Arithmetic and accumulation are fp64 throughout, in the operation order of the generic kernel (llg_point), so the
// two paths agree to rounding.
//
// Eligibility (host): fp32 fields, W % 4 == 0, 16-byte aligned bases, strides % 4 == 0, fp32 observations, uint8 masks.

template <int TW>
struct LlgTile {
    static constexpr int TH = 1024 / TW;
    static constexpr int PITCH = TW + 4;        // staged columns x0-2 .. x0+TW+1 (doubles)
    static constexpr int MROWS = TH + 4;        // staged rows y0-2 .. y0+TH+1
    static constexpr int GROWS = TH + 2;        // G_H rows y0-1 .. y0+TH
    static constexpr int NG = TW / 4 + 2;       // 16-byte chunks per raw magnetisation row (x0-4 .. x0+TW+3)
    static constexpr int NC = TW / 4;           // 16-byte chunks per tile row
    static constexpr int M_DOUBLES = 3 * MROWS * PITCH;
    static constexpr int G_DOUBLES = 3 * GROWS * PITCH;
    static constexpr int NPAIR = TH * TW / 2;   // pixel pairs per tile (2 per thread)
    static constexpr int NRING = TW + 2 * GROWS;  // ring work items: top pairs, bottom pairs, side singles
    // raw (fp32 / uint8) landing zones of the cp.async prefetch
    static constexpr int RAW_M_BYTES = 3 * MROWS * NG * 16;
    static constexpr int RAW_D_BYTES = 3 * GROWS * NC * 16;   // dmdt rows y0-1 .. y0+TH, tile columns
    static constexpr int RAW_O_BYTES = 3 * TH * NC * 16;      // observations, tile pixels
    static constexpr int RAW_K_BYTES = 3 * TH * TW;           // masks (uint8), tile pixels
    static constexpr int AUX_BYTES = RAW_D_BYTES + RAW_O_BYTES + RAW_K_BYTES;
    static constexpr int smem_bytes(bool vjp) { return (M_DOUBLES + (vjp ? G_DOUBLES : 0)) * 8 + RAW_M_BYTES + (vjp ? 1 : 2) * AUX_BYTES; }
};

struct LlgTileCoord {
    int b, y0, x0;
};

__device__ __forceinline__ LlgTileCoord llg_decode_tile(const Params& p, int t, int tiles_x, int TH, int TW) {
    // batch index innermost: observation / mask tiles that broadcast over the batch are fetched from HBM once
    const unsigned q = (unsigned)t / (unsigned)p.B;
    LlgTileCoord c;
    c.b = (int)((unsigned)t - q * p.B);
    const unsigned ty = q / (unsigned)tiles_x;
    c.x0 = (int)(q - ty * tiles_x) * TW;
    c.y0 = p.ylo + (int)ty * TH;
    return c;
}

// Start the asynchronous copies of one tile's magnetisation rows (+ 2-pixel halo, rows reflected) into `raw`.
template <int TW>
__device__ __forceinline__ void llg_prefetch_m(const Params& p, const LlgTileCoord& tc, unsigned raw, int tid) {
    using T = LlgTile<TW>;
    const float* m0 = reinterpret_cast<const float*>(p.x0.p) + (int64_t)tc.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc;
    constexpr int TOTAL = 3 * T::MROWS * T::NG;
#pragma unroll
    for (int k = 0; k < (TOTAL + kThreads - 1) / kThreads; ++k) {
        const int i = k * kThreads + tid;
        if (i >= TOTAL) break;
        const int g = i % T::NG, rc = i / T::NG, sr = rc % T::MROWS, c = rc / T::MROWS;
        const int x = tc.x0 - 4 + 4 * g;
        if (x < 0 || x >= p.W) continue;  // out-of-grid columns are never consumed (edge selects in llg_eval)
        cp_async16(raw + i * 16, m0 + (int64_t)c * p.x0.sc + row_offset(p, tc.y0 - 2 + sr) + x);
    }
}

// ... and of its time derivative (rows y0-1 .. y0+TH), observations and masks (tile pixels) into `aux`.
template <int TW>
__device__ __forceinline__ void llg_prefetch_aux(const Params& p, const LlgTileCoord& tc, unsigned aux, int tid) {
    using T = LlgTile<TW>;
    if (p.dxdt.p) {
        const float* d0 = reinterpret_cast<const float*>(p.dxdt.p) + (int64_t)tc.b * p.dxdt.sb + (int64_t)p.ch_a * p.dxdt.sc;
        constexpr int TOTAL = 3 * T::GROWS * T::NC;
#pragma unroll
        for (int k = 0; k < (TOTAL + kThreads - 1) / kThreads; ++k) {
            const int i = k * kThreads + tid;
            if (i >= TOTAL) break;
            const int g = i % T::NC, rc = i / T::NC, gr = rc % T::GROWS, c = rc / T::GROWS;
            const int x = tc.x0 + 4 * g, y = tc.y0 - 1 + gr;
            if (x >= p.W || y < 0 || y >= p.H) continue;
            cp_async16(aux + i * 16, d0 + (int64_t)c * p.dxdt.sc + (int64_t)y * p.W + x);
        }
    }
    if (p.has_u) {
        const float* ob = reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)tc.b * p.obs_u.sb;
        const unsigned char* mk = reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)tc.b * p.mask_u.sb;
        constexpr int TOTAL = 3 * T::TH * T::NC;
#pragma unroll
        for (int k = 0; k < (TOTAL + kThreads - 1) / kThreads; ++k) {
            const int i = k * kThreads + tid;
            if (i >= TOTAL) break;
            const int g = i % T::NC, rc = i / T::NC, ly = rc % T::TH, c = rc / T::TH;
            const int x = tc.x0 + 4 * g, y = tc.y0 + ly;
            if (x >= p.W || y >= p.yhi) continue;
            const int64_t pix = (int64_t)y * p.W + x;
            cp_async16(aux + T::RAW_D_BYTES + i * 16, ob + (int64_t)c * p.obs_u.sc + pix);
            cp_async4(aux + T::RAW_D_BYTES + T::RAW_O_BYTES + i * 4, mk + (int64_t)c * p.mask_u.sc + pix);
        }
    }
}

// Widen the landed magnetisation rows once and park them in the fp64 stage.
template <int TW>
__device__ __forceinline__ void llg_convert_m(const Params& p, int x0, const unsigned char* __restrict__ raw, double* __restrict__ ms, int tid) {
    using T = LlgTile<TW>;
    constexpr int TOTAL = 3 * T::MROWS * T::NG;
#pragma unroll
    for (int k = 0; k < (TOTAL + kThreads - 1) / kThreads; ++k) {
        const int i = k * kThreads + tid;
        if (i >= TOTAL) break;
        const int g = i % T::NG, rc = i / T::NG;   // rc = c * MROWS + sr
        const int x = x0 - 4 + 4 * g;
        if (x < 0 || x >= p.W) continue;
        const float4 v = *reinterpret_cast<const float4*>(raw + i * 16);
        double* dst = ms + rc * T::PITCH + (4 * g - 2);
        if (g == 0) {
            *reinterpret_cast<double2*>(dst + 2) = make_double2((double)v.z, (double)v.w);
        } else if (g == T::NG - 1) {
            *reinterpret_cast<double2*>(dst) = make_double2((double)v.x, (double)v.y);
        } else {
            *reinterpret_cast<double2*>(dst) = make_double2((double)v.x, (double)v.y);
            *reinterpret_cast<double2*>(dst + 2) = make_double2((double)v.z, (double)v.w);
        }
    }
}

template <int NP>
struct LlgPt {
    double m[NP][3], Hf[NP][3], a[NP][3], r[NP][3];
};

// H_eff, a = m x H and the residual of NP (1 or 2, horizontally adjacent) pixels whose first one sits at staged
// position (sr, sc) = global column x;  dtv = dmdt of those pixels.
template <int NP, int TW>
__device__ __forceinline__ void llg_eval(const Params& p, const double* __restrict__ ms, int sr, int sc, int x,
                                         const double* hext, const double (*dtv)[3], LlgPt<NP>& o) {
    using T = LlgTile<TW>;
    double lap[NP][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double* q = ms + (c * T::MROWS + sr) * T::PITCH + sc;
        if (NP == 2) {
            const double2 ct = *reinterpret_cast<const double2*>(q);
            const double2 up = *reinterpret_cast<const double2*>(q - T::PITCH), dn = *reinterpret_cast<const double2*>(q + T::PITCH);
            double lf = q[-1], rt = q[2];
            if (x == 0) lf = ct.y;             // reflect: m[-1] = m[1]
            if (x + 2 == p.W) rt = ct.x;       // m[W] = m[W-2]
            o.m[0][c] = ct.x;
            o.m[NP - 1][c] = ct.y;
            lap[0][c] = ((up.x + dn.x) + (lf + ct.y)) - 4.0 * ct.x;
            lap[NP - 1][c] = ((up.y + dn.y) + (ct.x + rt)) - 4.0 * ct.y;
        } else {
            const double ct = q[0];
            const double lf = (x == 0) ? q[1] : q[-1], rt = (x == p.W - 1) ? q[-1] : q[1];
            o.m[0][c] = ct;
            lap[0][c] = ((q[-T::PITCH] + q[T::PITCH]) + (lf + rt)) - 4.0 * ct;
        }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o.Hf[k][c] = hext[c] + p.c_ex * (lap[k][c] * p.inv_dx2);
        if (p.c_an != 0.0) {
            const double me = p.c_an * (o.m[k][0] * p.e[0] + o.m[k][1] * p.e[1] + o.m[k][2] * p.e[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) o.Hf[k][c] += me * p.e[c];
        }
        cross3(o.m[k], o.Hf[k], o.a[k]);
        double ma[3];
        cross3(o.m[k], o.a[k], ma);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double rhs = -p.gamma * o.a[k][c] - p.alpha * ma[c];
            o.r[k][c] = dtv[k][c] - rhs * p.tau;
        }
    }
}

// dmdt of a pixel pair from the landed aux buffer (gr = row - (y0 - 1), lx = column - x0); zeros when absent
template <int TW>
__device__ __forceinline__ void llg_dt_pair(const Params& p, const unsigned char* __restrict__ aux, int gr, int lx, double (*dtv)[3]) {
    using T = LlgTile<TW>;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (p.dxdt.p) {
            const float2 f = *reinterpret_cast<const float2*>(aux + ((c * T::GROWS + gr) * TW + lx) * 4);
            dtv[0][c] = (double)f.x;
            dtv[1][c] = (double)f.y;
        } else {
            dtv[0][c] = 0.0;
            dtv[1][c] = 0.0;
        }
    }
}

// observation term of a pixel pair of component c from the landed aux buffer: mask and mask (x - obs) per pixel
template <int TW>
__device__ __forceinline__ void llg_obs_pair(const unsigned char* __restrict__ aux, int c, int ly, int lx, double x0v, double x1v,
                                             double& d0, double& d1, double& k0, double& k1) {
    using T = LlgTile<TW>;
    const int e = (c * T::TH + ly) * TW + lx;
    const float2 o = *reinterpret_cast<const float2*>(aux + T::RAW_D_BYTES + e * 4);
    const uchar2 k = *reinterpret_cast<const uchar2*>(aux + T::RAW_D_BYTES + T::RAW_O_BYTES + e);
    k0 = u8_to_double(k.x);
    k1 = u8_to_double(k.y);
    d0 = k0 * (x0v - (double)o.x);
    d1 = k1 * (x1v - (double)o.y);
}

// ---------------------------------------------------------------------------------------------------------
// m x H_eff residual, pass 1: S_a, S_u, S_pde
//   Pipeline per CTA: while tile t is evaluated, the cp.async copies of tile t + grid (magnetisation rows and the
//   second aux buffer) are in flight; the only exposed wait is at the top of the loop.
// ---------------------------------------------------------------------------------------------------------
template <int TW>
__global__ void __launch_bounds__(kThreads, 2)
llg_tile_reduce_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, int tiles_x, int n_tiles,
                       double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ sums,
                       int finalize, double* __restrict__ scal, float* __restrict__ trace) {
    using T = LlgTile<TW>;
    extern __shared__ __align__(16) unsigned char llg_smem[];
    double* ms = reinterpret_cast<double*>(llg_smem);
    unsigned char* raw_m = llg_smem + T::M_DOUBLES * 8;
    unsigned char* aux0 = raw_m + T::RAW_M_BYTES;
    const unsigned raw_m_s = (unsigned)__cvta_generic_to_shared(raw_m), aux_s = (unsigned)__cvta_generic_to_shared(aux0);
    __shared__ double scratch[3 * (kThreads / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    double s_a = 0.0, s_u = 0.0, s_p = 0.0;

    int t = blockIdx.x, buf = 0;
    if (t < n_tiles) {
        const LlgTileCoord tc = llg_decode_tile(p, t, tiles_x, T::TH, TW);
        llg_prefetch_m<TW>(p, tc, raw_m_s, tid);
        llg_prefetch_aux<TW>(p, tc, aux_s, tid);
    }
    cp_async_commit();
    for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
        const LlgTileCoord tc = llg_decode_tile(p, t, tiles_x, T::TH, TW);
        const double hext[3] = {__ldg(p.coef + 3 * tc.b), __ldg(p.coef + 3 * tc.b + 1), __ldg(p.coef + 3 * tc.b + 2)};
        cp_async_wait<0>();
        __syncthreads();   // this tile's copies have landed; every thread is done with the previous tile's stage
        llg_convert_m<TW>(p, tc.x0, raw_m, ms, tid);
        __syncthreads();   // stage complete, raw_m free again
        if (t + (int)gridDim.x < n_tiles) {
            const LlgTileCoord tn = llg_decode_tile(p, t + gridDim.x, tiles_x, T::TH, TW);
            llg_prefetch_m<TW>(p, tn, raw_m_s, tid);
            llg_prefetch_aux<TW>(p, tn, aux_s + (buf ^ 1) * T::AUX_BYTES, tid);
        }
        cp_async_commit();
        const unsigned char* aux = aux0 + buf * T::AUX_BYTES;
#pragma unroll 1
        for (int k = 0; k < T::NPAIR / kThreads; ++k) {
            const int e = k * kThreads + tid, ly = e / (TW / 2), lx = 2 * (e % (TW / 2));
            const int y = tc.y0 + ly, x = tc.x0 + lx;
            if (y < p.yhi && x < p.W) {
                LlgPt<2> q;
                double dtv[2][3];
                llg_dt_pair<TW>(p, aux, ly + 1, lx, dtv);
                llg_eval<2, TW>(p, ms, ly + 2, lx + 2, x, hext, dtv, q);
                s_p += ((q.r[0][0] * q.r[0][0] + q.r[0][1] * q.r[0][1]) + q.r[0][2] * q.r[0][2]) +
                       ((q.r[1][0] * q.r[1][0] + q.r[1][1] * q.r[1][1]) + q.r[1][2] * q.r[1][2]);
                if (p.has_u) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double e0, e1, k0, k1;
                        llg_obs_pair<TW>(aux, c, ly, lx, q.m[0][c], q.m[1][c], e0, e1, k0, k1);
                        s_u += e0 * e0 + e1 * e1;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    if (p.has_a) {
        const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);
        for (int item = warp0; item < g.n_a_items; item += nwarps) a_item_reduce(p, g, item, lane, s_a);
    }
    reduce_epilogue(p, s_a, s_u, s_p, scratch, &is_last, partials, ticket, sums, finalize, scal, trace);
}

// ---------------------------------------------------------------------------------------------------------
// m x H_eff residual, pass 2: seed gradient
//   With seed s = c_p r, q = s x m:  G_H = -gamma q - alpha (q x m),  G_m = -gamma (H x s) - alpha (a x s + H x q),
//   d loss / d m = -tau [ G_m + (c_ex/dx^2) K^T G_H + c_an (e . G_H) e ]  + observation term.
// ---------------------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void llg_field_grad(const Params& p, double c_p, const LlgPt<NP>& q, int k, double* s, double* qq, double* GH) {
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] = c_p * q.r[k][c];
    double qm[3];
    cross3(s, q.m[k], qq);
    cross3(qq, q.m[k], qm);
#pragma unroll
    for (int c = 0; c < 3; ++c) GH[c] = -p.gamma * qq[c] - p.alpha * qm[c];
}

template <int TW>
__global__ void __launch_bounds__(kThreads, 2)
llg_tile_vjp_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, int tiles_x, int n_tiles,
                    const double* __restrict__ scal, const double* __restrict__ upstream, float* __restrict__ g_x0,
                    float* __restrict__ g_dxdt) {
    using T = LlgTile<TW>;
    extern __shared__ __align__(16) unsigned char llg_smem[];
    double* ms = reinterpret_cast<double*>(llg_smem);
    double* gs = ms + T::M_DOUBLES;
    unsigned char* raw_m = llg_smem + (T::M_DOUBLES + T::G_DOUBLES) * 8;
    unsigned char* aux = raw_m + T::RAW_M_BYTES;
    const unsigned raw_m_s = (unsigned)__cvta_generic_to_shared(raw_m), aux_s = (unsigned)__cvta_generic_to_shared(aux);
    const int tid = threadIdx.x, lane = tid & 31;
    const double up = upstream ? __ldg(upstream) : 1.0;
    const double c_a = __ldg(scal + 4) * up, c_u = __ldg(scal + 5) * up, c_p = __ldg(scal + 6) * up;
    const float* dxp = reinterpret_cast<const float*>(p.dxdt.p);
    const int64_t plane = (int64_t)p.H * p.W;
    const double kx = -p.tau * p.c_ex * p.inv_dx2;
    constexpr int NK = T::NPAIR / kThreads;  // 2

    int t = blockIdx.x;
    if (t < n_tiles) {
        const LlgTileCoord tc = llg_decode_tile(p, t, tiles_x, T::TH, TW);
        llg_prefetch_m<TW>(p, tc, raw_m_s, tid);
        llg_prefetch_aux<TW>(p, tc, aux_s, tid);
    }
    cp_async_commit();
    for (; t < n_tiles; t += gridDim.x) {
        const LlgTileCoord tc = llg_decode_tile(p, t, tiles_x, T::TH, TW);
        const bool more = t + (int)gridDim.x < n_tiles;
        const LlgTileCoord tn = llg_decode_tile(p, more ? t + gridDim.x : t, tiles_x, T::TH, TW);
        float* gm = g_x0 + ((int64_t)tc.b * p.C + p.ch_a) * plane;
        float* gd = g_dxdt ? g_dxdt + ((int64_t)tc.b * p.C + p.ch_a) * plane : nullptr;
        const double hext[3] = {__ldg(p.coef + 3 * tc.b), __ldg(p.coef + 3 * tc.b + 1), __ldg(p.coef + 3 * tc.b + 2)};
        cp_async_wait<0>();
        __syncthreads();   // this tile's copies have landed; every thread is done with phase B of the previous tile
        llg_convert_m<TW>(p, tc.x0, raw_m, ms, tid);
        __syncthreads();   // m stage complete, raw_m free again
        if (more) llg_prefetch_m<TW>(p, tn, raw_m_s, tid);
        cp_async_commit();

        // ---- phase A, own pixels: G_H -> shared tile, pointwise gradient part -> registers
        double local[NK][2][3];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int e = k * kThreads + tid, ly = e / (TW / 2), lx = 2 * (e % (TW / 2));
            const int y = tc.y0 + ly, x = tc.x0 + lx;
            double GH[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) local[k][j][c] = 0.0;
            // rows ylo-1 .. yhi of the global grid feed the transposed stencil; own rows are y < yhi
            if (y <= p.yhi && y + p.yg0 < p.Hg && x < p.W) {
                LlgPt<2> q;
                double dtv[2][3];
                llg_dt_pair<TW>(p, aux, ly + 1, lx, dtv);
                llg_eval<2, TW>(p, ms, ly + 2, lx + 2, x, hext, dtv, q);
                double sv[2][3];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double qq[3], Hs[3], as[3], Hq[3];
                    llg_field_grad<2>(p, c_p, q, j, sv[j], qq, GH[j]);
                    cross3(q.Hf[j], sv[j], Hs);
                    cross3(q.a[j], sv[j], as);
                    cross3(q.Hf[j], qq, Hq);
                    const double eG = (p.e[0] * GH[j][0] + p.e[1] * GH[j][1]) + p.e[2] * GH[j][2];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const double Gm = -p.gamma * Hs[c] - p.alpha * (as[c] + Hq[c]);
                        local[k][j][c] = -p.tau * (Gm + p.c_an * eG * p.e[c]);
                    }
                }
                if (gd && y < p.yhi) {
                    const int64_t pix = (int64_t)y * p.W + x;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        *reinterpret_cast<float2*>(gd + c * plane + pix) = make_float2((float)sv[0][c], (float)sv[1][c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                *reinterpret_cast<double2*>(gs + (c * T::GROWS + ly + 1) * T::PITCH + lx + 2) = make_double2(GH[0][c], GH[1][c]);
        }
        // ---- phase A, 1-pixel ring: G_H only (zero outside the grid / outside the rows that feed owned pixels)
        if (tid < T::NRING) {
            if (tid < TW) {  // top / bottom rows, as pairs
                const int bottom = tid >= TW / 2, lx = 2 * (tid - bottom * (TW / 2));
                const int ly = bottom ? T::TH : -1, y = tc.y0 + ly, x = tc.x0 + lx, gy = y + p.yg0;
                double GH[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
                if (y >= p.ylo - 1 && y <= p.yhi && gy >= 0 && gy < p.Hg && x < p.W) {
                    LlgPt<2> q;
                    double dtv[2][3];
                    llg_dt_pair<TW>(p, aux, ly + 1, lx, dtv);
                    llg_eval<2, TW>(p, ms, ly + 2, lx + 2, x, hext, dtv, q);
                    double s[3], qq[3];
                    llg_field_grad<2>(p, c_p, q, 0, s, qq, GH[0]);
                    llg_field_grad<2>(p, c_p, q, 1, s, qq, GH[1]);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    *reinterpret_cast<double2*>(gs + (c * T::GROWS + ly + 1) * T::PITCH + lx + 2) = make_double2(GH[0][c], GH[1][c]);
            } else {  // left / right columns, single pixels, rows y0-1 .. y0+TH (their dmdt is not staged: direct loads)
                const int j = tid - TW, gr = j >> 1, right = j & 1;
                const int ly = gr - 1, lx = right ? TW : -1, y = tc.y0 + ly, x = tc.x0 + lx, gy = y + p.yg0;
                double GH[3] = {0.0, 0.0, 0.0};
                if (y >= p.ylo - 1 && y <= p.yhi && gy >= 0 && gy < p.Hg && x >= 0 && x < p.W) {
                    LlgPt<1> q;
                    double dtv[1][3] = {{0.0, 0.0, 0.0}};
                    if (dxp) {
                        const float* d0 = dxp + (int64_t)tc.b * p.dxdt.sb + (int64_t)p.ch_a * p.dxdt.sc + (int64_t)y * p.W + x;
#pragma unroll
                        for (int c = 0; c < 3; ++c) dtv[0][c] = (double)__ldg(d0 + c * p.dxdt.sc);
                    }
                    llg_eval<1, TW>(p, ms, ly + 2, lx + 2, x, hext, dtv, q);
                    double s[3], qq[3];
                    llg_field_grad<1>(p, c_p, q, 0, s, qq, GH);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) gs[(c * T::GROWS + gr) * T::PITCH + lx + 2] = GH[c];
            }
        }
        __syncthreads();   // G_H tile complete

        // ---- phase B: transposed stencil of G_H + pointwise part + observation term -> g
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int e = k * kThreads + tid, ly = e / (TW / 2), lx = 2 * (e % (TW / 2));
            const int y = tc.y0 + ly, x = tc.x0 + lx;
            if (y < p.yhi && x < p.W) {
                const int64_t pix = (int64_t)y * p.W + x;
                const int gy = y + p.yg0;
                const double wu = adj_w(gy - 1, p.Hg), wd = adj_w(gy + 1, p.Hg);
                const double wl0 = adj_w(x - 1, p.W), wl1 = adj_w(x, p.W), wr0 = adj_w(x + 1, p.W), wr1 = adj_w(x + 2, p.W);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double* q = gs + (c * T::GROWS + ly + 1) * T::PITCH + lx + 2;
                    const double2 ct = *reinterpret_cast<const double2*>(q);
                    const double2 upv = *reinterpret_cast<const double2*>(q - T::PITCH), dnv = *reinterpret_cast<const double2*>(q + T::PITCH);
                    const double lf = q[-1], rt = q[2];
                    const double a0 = ((wu * upv.x + wd * dnv.x) + (wl0 * lf + wr0 * ct.y)) - 4.0 * ct.x;
                    const double a1 = ((wu * upv.y + wd * dnv.y) + (wl1 * ct.x + wr1 * rt)) - 4.0 * ct.y;
                    double v0 = local[k][0][c], v1 = local[k][1][c];
                    if (p.has_u) {
                        const double2 mc = *reinterpret_cast<const double2*>(ms + (c * T::MROWS + ly + 2) * T::PITCH + lx + 2);
                        double e0, e1, k0, k1;
                        llg_obs_pair<TW>(aux, c, ly, lx, mc.x, mc.y, e0, e1, k0, k1);
                        v0 += c_u * (k0 * e0);
                        v1 += c_u * (k1 * e1);
                    }
                    *reinterpret_cast<float2*>(gm + c * plane + pix) = make_float2((float)(v0 + kx * a0), (float)(v1 + kx * a1));
                }
            }
        }
        __syncthreads();   // aux consumed: start the next tile's copies (they overlap its stage conversion ... phase A)
        if (more) llg_prefetch_aux<TW>(p, tn, aux_s, tid);
        cp_async_commit();
    }
    cp_async_wait<0>();
    {
        const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);
        for (int item = warp0; item < g.n_a_items; item += nwarps) a_item_vjp(p, g, item, lane, c_a, g_x0, g_dxdt);
    }
}

// ---------------------------------------------------------------------------------------------------------
// soft unit-norm loss (llg_loss2, pde_losses.py:99-117): pointwise, streamed with 128-bit accesses.
// One warp item = kABlock float4 of the three magnetisation planes of one sample.
// ---------------------------------------------------------------------------------------------------------
struct NormItem {
    int b, first4, n4;
};
__device__ __forceinline__ NormItem norm_decode(const Params& p, const MarchGeom& g, int item) {
    const unsigned blk = (unsigned)item / (unsigned)p.B;
    NormItem n;
    n.b = (int)((unsigned)item - blk * p.B);
    n.first4 = (int)blk * g.a_block4;
    n.n4 = min(g.a_block4, g.a_plane4 - n.first4);
    return n;
}

__device__ __forceinline__ void widen4(const float4& f, double* d) {
    d[0] = (double)f.x; d[1] = (double)f.y; d[2] = (double)f.z; d[3] = (double)f.w;
}

// The three magnetisation planes (+ their observations and mask words) of a norm item stream through a per-thread
// cp.async ring: D elements (one float4 per plane and array) of every lane are in flight all the time, the slot an iteration
// drains is refilled with element t + D at once (VJP pass; round 1 loaded at the point of use).  Region: D x kThreads x 108 bytes (m | obs | masks); a-plane items use the same bytes as an
// ARing<3 D, kThreads, 2> (3 D x 36 bytes per thread = the same 108 D), a warp runs one kind of item at a time and a
// thread only ever touches its own bytes.
constexpr int kNormD = 3;
__host__ __device__ constexpr int norm_ring_bytes() { return kNormD * kThreads * 108; }

template <int D, bool HAS_O, typename F>
__device__ __forceinline__ void norm_item_stream(int n4, int lane, unsigned sbase, const float* m0, int64_t scm, const float* po, int64_t sco,
                                                 const unsigned char* pm, int64_t sck, F&& consume) {
    const unsigned m_base = sbase + threadIdx.x * 16, o_base = m_base + D * 3 * kThreads * 16;
    const unsigned k_base = sbase + 2 * D * 3 * kThreads * 16 + threadIdx.x * 4;
    auto M = [&](int slot, int c) { return m_base + (unsigned)((slot * 3 + c) * kThreads * 16); };
    auto O = [&](int slot, int c) { return o_base + (unsigned)((slot * 3 + c) * kThreads * 16); };
    auto K = [&](int slot, int c) { return k_base + (unsigned)((slot * 3 + c) * kThreads * 4); };
    const int T = (n4 + 31) >> 5;
    auto issue = [&](int slot, int t) {
        const int i = lane + 32 * t;
        if (i < n4) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cp_async16(M(slot, c), m0 + c * scm + 4 * i);
                if (HAS_O) {
                    cp_async16(O(slot, c), po + c * sco + 4 * i);
                    cp_async4(K(slot, c), pm + c * sck + 4 * i);
                }
            }
        }
        cp_async_commit();
    };
    static_for<D>([&](auto J) { issue(decltype(J)::value, decltype(J)::value); });
    const int groups = (T + D - 1) / D;
#pragma unroll 1
    for (int gi = 0; gi < groups; ++gi) {
        static_for<D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            const int t = gi * D + j, i = lane + 32 * t;
            cp_async_wait<D - 1>();                                   // element t has landed
            float4 mf[3], of[3];
            unsigned kk[3] = {0u, 0u, 0u};
            if (i < n4) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    mf[c] = lds128(M(j, c));
                    if (HAS_O) {
                        of[c] = lds128(O(j, c));
                        kk[c] = lds32(K(j, c));
                    }
                }
            }
            issue(j, t + D);
            if (i < n4) consume(i, mf, of, kk);
        });
    }
    cp_async_wait<0>();
}

// The reduce pass keeps the LDG form with one element of the magnetisation prefetched in registers (PF = 1) at three CTAs per SM.
// Measured on 8 x 6 x 2048^2 (profiles/r2w_probe_norm_reduce_pf.log, r2v_*, r2s_*): PF = 0 / 1 0.212 / 0.199 ms; PF = 2 / 3 need 128
// registers, i.e. two CTAs per SM: 0.253 / 0.229 ms -- the pass wants resident warps more than loads in flight per warp; m (or m,
// observations and masks) streamed through a cp.async ring: 0.248 (0.263) ms, slower, as in the heat reduce pass.  Without
// a-planes the norm items alone take 0.129 ms (3.1 TB/s), the a-plane items the other 0.07 ms (~6 TB/s).
template <bool HAS_O, int PF>
__global__ void __launch_bounds__(kThreads, PF >= 2 ? 2 : 3)   // measured 8x6x2048^2: 0.277 / 0.215 / 0.233 ms at 2 / 3 / 4 CTAs per SM
llg_norm_reduce_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, int n_items,
                       double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ sums,
                       int finalize, double* __restrict__ scal, float* __restrict__ trace) {
    __shared__ double scratch[3 * (kThreads / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);
    double s_a = 0.0, s_u = 0.0, s_p = 0.0;
    auto do_u = [&](int item) {
        const NormItem it = norm_decode(p, g, item);
        const int base = p.ylo * p.W + 4 * it.first4;
        const float* m0 = reinterpret_cast<const float*>(p.x0.p) + (int64_t)it.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc + base;
        const float* po = HAS_O ? reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)it.b * p.obs_u.sb + base : nullptr;
        const unsigned char* pm = HAS_O ? reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)it.b * p.mask_u.sb + base : nullptr;
        double sp0 = 0.0, sp1 = 0.0, su0 = 0.0, su1 = 0.0;
        auto body = [&](int i, const float4* mf, const float4* of, const unsigned* kk) {
            double mv[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c) widen4(mf[c], mv[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // |m| = s rsqrt(s): one reciprocal square root (1 ulp) instead of the IEEE square root's longer sequence
                const double s2 = (mv[0][j] * mv[0][j] + mv[1][j] * mv[1][j]) + mv[2][j] * mv[2][j];
                const double n = s2 > 0.0 ? s2 * rsqrt(s2) : 0.0;
                double& acc = (j & 1) ? sp1 : sp0;
                acc = fma(1.0 - n, 1.0 - n, acc);
            }
            if (HAS_O) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // 0 / 1 masks: an unobserved pixel selects its own value as the observation (difference exactly 0)
                    const unsigned k = kk[c];
                    const double d0 = mv[c][0] - (double)sel_obs(k & 0xffu, of[c].x, mf[c].x), d1 = mv[c][1] - (double)sel_obs(k & 0xff00u, of[c].y, mf[c].y);
                    const double d2 = mv[c][2] - (double)sel_obs(k & 0xff0000u, of[c].z, mf[c].z), d3 = mv[c][3] - (double)sel_obs(k & 0xff000000u, of[c].w, mf[c].w);
                    su0 = fma(d0, d0, su0);
                    su1 = fma(d1, d1, su1);
                    su0 = fma(d2, d2, su0);
                    su1 = fma(d3, d3, su1);
                }
            }
        };
        // PF elements of the magnetisation (the HBM stream) are prefetched into registers ahead of the one being processed;
        // observations / masks (L2 hits: they broadcast over the batch and the batch index is innermost) load at the top.
        float4 nx[PF > 0 ? PF : 1][3];
#pragma unroll
        for (int q = 0; q < PF; ++q)
            if (lane + 32 * q < it.n4) {
#pragma unroll
                for (int c = 0; c < 3; ++c) nx[q][c] = ldg4(m0 + c * p.x0.sc + 4 * (lane + 32 * q));
            }
#pragma unroll 1
        for (int i0 = lane; i0 < it.n4; i0 += 32 * (PF > 0 ? PF : 1)) {
#pragma unroll
            for (int q = 0; q < (PF > 0 ? PF : 1); ++q) {
                const int i = i0 + 32 * q;
                if (i < it.n4) {
                    float4 mf[3], of[3];
                    unsigned kk[3];
                    if (HAS_O) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            of[c] = ldg4(po + c * p.obs_u.sc + 4 * i);
                            kk[c] = __ldg(reinterpret_cast<const unsigned*>(pm + c * p.mask_u.sc + 4 * i));
                        }
                    }
                    if (PF > 0) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) mf[c] = nx[q][c];
                        if (i + 32 * PF < it.n4) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) nx[q][c] = ldg4(m0 + c * p.x0.sc + 4 * (i + 32 * PF));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) mf[c] = ldg4(m0 + c * p.x0.sc + 4 * i);
                    }
                    body(i, mf, of, kk);
                }
            }
        }
        s_p += sp0 + sp1;
        s_u += su0 + su1;
    };
    auto do_a = [&](int item) { a_item_reduce(p, g, item, lane, s_a); };
    run_interleaved(warp0, nwarps, n_items, p.has_a ? g.n_a_items : 0, (tid >> 5) & 1, do_u, do_a);
    reduce_epilogue(p, s_a, s_u, s_p, scratch, &is_last, partials, ticket, sums, finalize, scal, trace);
}

// three elements per lane in flight, two CTAs per SM (measured 8 x 6 x 2048^2: 0.322 ms; D = 2 at three CTAs 0.334, at four 0.364;
// round 1's load-at-use form with sqrt + division 0.393)
template <bool HAS_O>
__global__ void __launch_bounds__(kThreads, 2)
llg_norm_vjp_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, int n_items,
                    const double* __restrict__ scal, const double* __restrict__ upstream, float* __restrict__ g_x0,
                    float* __restrict__ g_dxdt) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(ring_mem);
    const double up = upstream ? __ldg(upstream) : 1.0;
    const double c_a = __ldg(scal + 4) * up, c_u = __ldg(scal + 5) * up, c_p = __ldg(scal + 6) * up;
    const int64_t plane = (int64_t)p.H * p.W;
    auto do_u = [&](int item) {
        const NormItem it = norm_decode(p, g, item);
        const int base = p.ylo * p.W + 4 * it.first4;
        const float* m0 = reinterpret_cast<const float*>(p.x0.p) + (int64_t)it.b * p.x0.sb + (int64_t)p.ch_a * p.x0.sc + base;
        const float* po = HAS_O ? reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)it.b * p.obs_u.sb + base : nullptr;
        const unsigned char* pm = HAS_O ? reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)it.b * p.mask_u.sb + base : nullptr;
        float* gm = g_x0 + ((int64_t)it.b * p.C + p.ch_a) * plane + base;
        float* gd = g_dxdt ? g_dxdt + ((int64_t)it.b * p.C + p.ch_a) * plane + base : nullptr;
        norm_item_stream<kNormD, HAS_O>(it.n4, lane, sbase, m0, p.x0.sc, po, p.obs_u.sc, pm, p.mask_u.sc,
                                        [&](int i, const float4* mf, const float4* of, const unsigned* kk) {
            double mv[3][4], f[4];
#pragma unroll
            for (int c = 0; c < 3; ++c) widen4(mf[c], mv[c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // g_m = -c_p (1 - n) m / n = c_p (1 - 1 / n) m  (0 where n == 0, as torch.linalg.norm's backward): one rsqrt, no division
                const double s2 = (mv[0][j] * mv[0][j] + mv[1][j] * mv[1][j]) + mv[2][j] * mv[2][j];
                f[j] = s2 > 0.0 ? fma(-c_p, rsqrt(s2), c_p) : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = f[j] * mv[c][j];
                if (HAS_O) {
                    const unsigned k = kk[c];
                    v[0] = fma(c_u, mv[c][0] - (double)sel_obs(k & 0xffu, of[c].x, mf[c].x), v[0]);
                    v[1] = fma(c_u, mv[c][1] - (double)sel_obs(k & 0xff00u, of[c].y, mf[c].y), v[1]);
                    v[2] = fma(c_u, mv[c][2] - (double)sel_obs(k & 0xff0000u, of[c].z, mf[c].z), v[2]);
                    v[3] = fma(c_u, mv[c][3] - (double)sel_obs(k & 0xff000000u, of[c].w, mf[c].w), v[3]);
                }
                *reinterpret_cast<float4*>(gm + c * plane + 4 * i) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
                if (gd) *reinterpret_cast<float4*>(gd + c * plane + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        });
    };
    auto do_a = [&](int item) { a_item_vjp_ring<3 * kNormD, kThreads, 2>(p, g, item, lane, sbase, c_a, g_x0, g_dxdt); };
    run_interleaved(warp0, nwarps, n_items, g.n_a_items, (tid >> 5) & 1, do_u, do_a);
}
