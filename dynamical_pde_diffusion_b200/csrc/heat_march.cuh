// Heat-equation guidance, fast path: register row-marching with 128-bit global accesses and warp shuffles.
//
// Included by guidance.cu (inside namespace dpde::{anonymous}); uses Params / adj_w / block_sum3 / finalize_scalars.
//
// Why this shape (ncu, profiles/r1a_*): the tile-per-CTA kernel issued ~100 instructions and 6-8 scalar loads per
// pixel and stalled on every load (0.8 eligible warps per scheduler, 21 % of DRAM bandwidth).  Here one lane owns
// FOUR adjacent columns (one float4) and marches down the rows of its chunk:
//   * every field row is read once from HBM as a coalesced 512-byte warp access (LDG.128), the vertical stencil
//     neighbours live in registers (three-row window of u, three-row window of the residual r);
//   * horizontal neighbours come from the adjacent lane by __shfl (u: 2 shuffles per row, r: 2 double shuffles);
//   * the next row's loads are issued before the current row's arithmetic (software prefetch), all loads are
//     unconditional (row / column indices are clamped, results of invalid lanes are zeroed afterwards);
//   * grids wider than 128 columns are cut into strips of 120 output columns + one halo lane on each side
//     (halo reads hit L2); narrow grids pack several row segments into one warp (W = 64: two, W = 16: eight).
// A lane spends ~30 instructions per pixel; arithmetic and accumulation stay fp64 (see guidance.cu header).
//
// Eligibility (checked on the host, else the generic tile kernel runs): fp32 fields, W % 4 == 0, 16-byte aligned
// base pointers and strides % 4 == 0, observations fp32, masks uint8 (bool).

struct MarchGeom {
    int lw_log2, segs_per_warp, strips, strip_w, halo_lane, R, chunks;
    int64_t n_seg_items, n_warp_items;
    int64_t a_total4, a_plane4;  // float4 count over all a-planes (owned rows) and per plane
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uchar4 ldg4(const unsigned char* p) { return __ldg(reinterpret_cast<const uchar4*>(p)); }

struct MarchLane {
    int b, cu, col0, ys, ye;
    bool lane_ok, out_ok, left_edge, right_edge;
};

__device__ __forceinline__ MarchLane march_decode(const Params& p, const MarchGeom& g, int64_t wi, int lane) {
    const int LW = 1 << g.lw_log2, seg = lane >> g.lw_log2, l = lane & (LW - 1);
    int64_t si = wi * g.segs_per_warp + seg;
    const bool seg_ok = si < g.n_seg_items;
    if (!seg_ok) si = g.n_seg_items - 1;
    MarchLane m;
    m.b = (int)(si % p.B);
    int64_t t = si / p.B;
    m.cu = (int)(t % p.n_u_units);
    t /= p.n_u_units;
    const int strip = (int)(t % g.strips), chunk = (int)(t / g.strips);
    m.col0 = strip * g.strip_w - 4 * g.halo_lane + 4 * l;
    m.lane_ok = seg_ok && m.col0 >= 0 && m.col0 < p.W;
    m.out_ok = m.lane_ok && m.col0 >= strip * g.strip_w && m.col0 < (strip + 1) * g.strip_w;
    m.left_edge = (m.col0 == 0);
    m.right_edge = (m.col0 + 4 == p.W);
    m.ys = p.ylo + chunk * g.R;
    m.ye = min(m.ys + g.R, p.yhi);
    return m;
}

// residual of one row for the lane's four columns: r = dudt - a_s * (up + down + left + right - 4 centre)
__device__ __forceinline__ void heat_row_residual(const float4& up, const float4& c, const float4& dn, float lf, float rt,
                                                  const float4& dt, double a_s, bool ok, double* r) {
    const double c0 = c.x, c1 = c.y, c2 = c.z, c3 = c.w;
    const double s0 = (((double)up.x + (double)dn.x) + ((double)lf + c1)) - 4.0 * c0;
    const double s1 = (((double)up.y + (double)dn.y) + (c0 + c2)) - 4.0 * c1;
    const double s2 = (((double)up.z + (double)dn.z) + (c1 + c3)) - 4.0 * c2;
    const double s3 = (((double)up.w + (double)dn.w) + (c2 + (double)rt)) - 4.0 * c3;
    r[0] = ok ? (double)dt.x - a_s * s0 : 0.0;
    r[1] = ok ? (double)dt.y - a_s * s1 : 0.0;
    r[2] = ok ? (double)dt.z - a_s * s2 : 0.0;
    r[3] = ok ? (double)dt.w - a_s * s3 : 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// pass 1 (fast): S_a, S_u, S_pde
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
heat_march_reduce_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g,
                         double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ sums,
                         int finalize, double* __restrict__ scal, float* __restrict__ trace) {
    __shared__ double scratch[3 * (kThreads / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const float* x0 = reinterpret_cast<const float*>(p.x0.p);
    const float* dxp = reinterpret_cast<const float*>(p.dxdt.p);
    double s_a = 0.0, s_u = 0.0, s_p = 0.0;

    // ---- a-planes: sum (mask (a - obs))^2, pure float4 streaming over owned rows
    if (p.has_a) {
        const float* ob = reinterpret_cast<const float*>(p.obs_a.p);
        const unsigned char* mk = reinterpret_cast<const unsigned char*>(p.mask_a.p);
        for (int64_t i = (int64_t)blockIdx.x * kThreads + tid; i < g.a_total4; i += (int64_t)gridDim.x * kThreads) {
            const int64_t pl = i / g.a_plane4, pix = (int64_t)p.ylo * p.W + 4 * (i - pl * g.a_plane4);
            const int b = (int)(pl / p.ch_a), ch = (int)(pl - (int64_t)b * p.ch_a);
            const float4 a = ldg4(x0 + (int64_t)b * p.x0.sb + (int64_t)ch * p.x0.sc + pix);
            const float4 o = ldg4(ob + (int64_t)b * p.obs_a.sb + (int64_t)ch * p.obs_a.sc + pix);
            const uchar4 m = ldg4(mk + (int64_t)b * p.mask_a.sb + (int64_t)ch * p.mask_a.sc + pix);
            const double d0 = (double)m.x * ((double)a.x - (double)o.x), d1 = (double)m.y * ((double)a.y - (double)o.y);
            const double d2 = (double)m.z * ((double)a.z - (double)o.z), d3 = (double)m.w * ((double)a.w - (double)o.w);
            s_a += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
    }

    // ---- u-planes: march down the rows
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int LW = 1 << g.lw_log2;
    for (int64_t wi = warp0; wi < g.n_warp_items; wi += nwarps) {
        const MarchLane m = march_decode(p, g, wi, lane);
        const int colc = m.lane_ok ? m.col0 : 0, ch = p.ch_a + m.cu;
        const float* u = x0 + (int64_t)m.b * p.x0.sb + (int64_t)ch * p.x0.sc + colc;
        const float* du = dxp ? dxp + (int64_t)m.b * p.dxdt.sb + (int64_t)ch * p.dxdt.sc + colc : nullptr;
        const float* ob = p.has_u ? reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)m.b * p.obs_u.sb + (int64_t)m.cu * p.obs_u.sc + colc : nullptr;
        const unsigned char* mk = p.has_u ? reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)m.b * p.mask_u.sb + (int64_t)m.cu * p.mask_u.sc + colc : nullptr;
        const double a_s = __ldg(p.coef + m.b) * p.inv_dx2;
        const int rmax = p.H - 1;
        auto row = [&](int y) { return (int64_t)min(max(y, 0), rmax) * p.W; };
        float4 ua = ldg4(u + row(m.ys - 1)), ub = ldg4(u + row(m.ys));
        float4 n_uc = ldg4(u + row(m.ys + 1));
        float4 n_dt = du ? ldg4(du + row(m.ys)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 n_ob = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 n_mk = make_uchar4(0, 0, 0, 0);
        if (p.has_u) {
            n_ob = ldg4(ob + row(m.ys));
            n_mk = ldg4(mk + row(m.ys));
        }
#pragma unroll 1
        for (int it = 0; it < g.R; ++it) {
            const int j = m.ys + it;
            const float4 uc = n_uc, dt = n_dt, o = n_ob;
            const uchar4 k = n_mk;
            n_uc = ldg4(u + row(j + 2));                         // prefetch the next row's operands
            if (du) n_dt = ldg4(du + row(j + 1));
            if (p.has_u) {
                n_ob = ldg4(ob + row(j + 1));
                n_mk = ldg4(mk + row(j + 1));
            }
            const int gj = j + p.yg0;
            float lf = __shfl_up_sync(0xffffffffu, ub.w, 1, LW), rt = __shfl_down_sync(0xffffffffu, ub.x, 1, LW);
            if (m.left_edge) lf = ub.y;                          // reflect: u[-1] = u[1]
            if (m.right_edge) rt = ub.z;
            const bool ok = m.out_ok && j < m.ye;
            double r[4];
            heat_row_residual(gj == 0 ? uc : ua, ub, gj == p.Hg - 1 ? ua : uc, lf, rt, dt, a_s, ok, r);
            s_p += (r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]);
            if (p.has_u && ok) {
                const double d0 = (double)k.x * ((double)ub.x - (double)o.x), d1 = (double)k.y * ((double)ub.y - (double)o.y);
                const double d2 = (double)k.z * ((double)ub.z - (double)o.z), d3 = (double)k.w * ((double)ub.w - (double)o.w);
                s_u += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            ua = ub;
            ub = uc;
        }
    }

    block_sum3(s_a, s_u, s_p, scratch);
    if (tid == 0) {
        partials[3 * blockIdx.x + 0] = s_a;
        partials[3 * blockIdx.x + 1] = s_u;
        partials[3 * blockIdx.x + 2] = s_p;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += kThreads) {
        a += __ldcg(partials + 3 * i);
        b += __ldcg(partials + 3 * i + 1);
        c += __ldcg(partials + 3 * i + 2);
    }
    block_sum3(a, b, c, scratch);
    if (tid == 0) {
        sums[0] = a;
        sums[1] = b;
        sums[2] = c;
        if (finalize) finalize_scalars(p, sums, scal, trace);
        *ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass 2 (fast): seed gradient
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
heat_march_vjp_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, const double* __restrict__ scal,
                      const double* __restrict__ upstream, float* __restrict__ g_x0, float* __restrict__ g_dxdt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double up = upstream ? __ldg(upstream) : 1.0;
    const double c_a = __ldg(scal + 4) * up, c_u = __ldg(scal + 5) * up, c_p = __ldg(scal + 6) * up;
    const float* x0 = reinterpret_cast<const float*>(p.x0.p);
    const float* dxp = reinterpret_cast<const float*>(p.dxdt.p);
    const int64_t plane = (int64_t)p.H * p.W;

    // ---- a-planes: g = c_a mask (mask (a - obs)), zeros when the mask is empty (sample.py:337-342)
    {
        const float* ob = reinterpret_cast<const float*>(p.obs_a.p);
        const unsigned char* mk = reinterpret_cast<const unsigned char*>(p.mask_a.p);
        for (int64_t i = (int64_t)blockIdx.x * kThreads + tid; i < g.a_total4; i += (int64_t)gridDim.x * kThreads) {
            const int64_t pl = i / g.a_plane4, pix = (int64_t)p.ylo * p.W + 4 * (i - pl * g.a_plane4);
            const int b = (int)(pl / p.ch_a), ch = (int)(pl - (int64_t)b * p.ch_a);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.has_a) {
                const float4 a = ldg4(x0 + (int64_t)b * p.x0.sb + (int64_t)ch * p.x0.sc + pix);
                const float4 o = ldg4(ob + (int64_t)b * p.obs_a.sb + (int64_t)ch * p.obs_a.sc + pix);
                const uchar4 m = ldg4(mk + (int64_t)b * p.mask_a.sb + (int64_t)ch * p.mask_a.sc + pix);
                v.x = (float)(c_a * ((double)m.x * ((double)m.x * ((double)a.x - (double)o.x))));
                v.y = (float)(c_a * ((double)m.y * ((double)m.y * ((double)a.y - (double)o.y))));
                v.z = (float)(c_a * ((double)m.z * ((double)m.z * ((double)a.z - (double)o.z))));
                v.w = (float)(c_a * ((double)m.w * ((double)m.w * ((double)a.w - (double)o.w))));
            }
            const int64_t off = ((int64_t)b * p.C + ch) * plane + pix;
            *reinterpret_cast<float4*>(g_x0 + off) = v;
            if (g_dxdt) *reinterpret_cast<float4*>(g_dxdt + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    // ---- u-planes
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int LW = 1 << g.lw_log2;
    for (int64_t wi = warp0; wi < g.n_warp_items; wi += nwarps) {
        const MarchLane m = march_decode(p, g, wi, lane);
        const int colc = m.lane_ok ? m.col0 : 0, ch = p.ch_a + m.cu;
        const float* u = x0 + (int64_t)m.b * p.x0.sb + (int64_t)ch * p.x0.sc + colc;
        const float* du = dxp ? dxp + (int64_t)m.b * p.dxdt.sb + (int64_t)ch * p.dxdt.sc + colc : nullptr;
        const float* ob = p.has_u ? reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)m.b * p.obs_u.sb + (int64_t)m.cu * p.obs_u.sc + colc : nullptr;
        const unsigned char* mk = p.has_u ? reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)m.b * p.mask_u.sb + (int64_t)m.cu * p.mask_u.sc + colc : nullptr;
        float* gout = g_x0 + ((int64_t)m.b * p.C + ch) * plane + colc;
        float* gdout = g_dxdt ? g_dxdt + ((int64_t)m.b * p.C + ch) * plane + colc : nullptr;
        const double alpha = __ldg(p.coef + m.b);
        const double a_s = alpha * p.inv_dx2, kp = -c_p * a_s;
        const double wl1 = m.left_edge ? 2.0 : 1.0, wr2 = m.right_edge ? 2.0 : 1.0;   // transposed-stencil edge weights
        const int rmax = p.H - 1;
        auto row = [&](int y) { return (int64_t)min(max(y, 0), rmax) * p.W; };

        // window: ua = u[j-1], ub = u[j], uc = u[j+1]; r2 = r[j-2], r1 = r[j-1]; j starts at ys-1
        float4 ua = ldg4(u + row(m.ys - 2)), ub = ldg4(u + row(m.ys - 1));
        float4 n_uc = ldg4(u + row(m.ys));
        float4 n_dt = du ? ldg4(du + row(m.ys - 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 n_ob = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 n_mk = make_uchar4(0, 0, 0, 0);
        if (p.has_u) {
            n_ob = ldg4(ob + row(m.ys - 2));
            n_mk = ldg4(mk + row(m.ys - 2));
        }
        double r2[4] = {0.0, 0.0, 0.0, 0.0}, r1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int it = 0; it < g.R + 2; ++it) {
            const int j = m.ys - 1 + it;                          // residual row computed in this iteration
            const float4 uc = n_uc, dt = n_dt, o = n_ob;
            const uchar4 k = n_mk;
            n_uc = ldg4(u + row(j + 2));                          // prefetch for the next iteration
            if (du) n_dt = ldg4(du + row(j + 1));
            if (p.has_u) {
                n_ob = ldg4(ob + row(j));
                n_mk = ldg4(mk + row(j));
            }
            const int gj = j + p.yg0;
            float lf = __shfl_up_sync(0xffffffffu, ub.w, 1, LW), rt = __shfl_down_sync(0xffffffffu, ub.x, 1, LW);
            if (m.left_edge) lf = ub.y;
            if (m.right_edge) rt = ub.z;
            const bool rok = m.lane_ok && gj >= 0 && gj < p.Hg && j <= p.yhi;
            double r0[4];
            heat_row_residual(gj == 0 ? uc : ua, ub, gj == p.Hg - 1 ? ua : uc, lf, rt, dt, a_s, rok, r0);

            // output row jo = j - 1: K^T r needs r[jo-1] (r2), r[jo] (r1) with its lane neighbours, r[jo+1] (r0)
            double l1 = __shfl_up_sync(0xffffffffu, r1[3], 1, LW), q1 = __shfl_down_sync(0xffffffffu, r1[0], 1, LW);
            if (m.left_edge) l1 = 0.0;
            if (m.right_edge) q1 = 0.0;
            const int jo = j - 1;
            if (it >= 2 && jo < m.ye && m.out_ok) {
                const int gjo = jo + p.yg0;
                const double wu = adj_w(gjo - 1, p.Hg), wd = adj_w(gjo + 1, p.Hg);
                const double a0 = ((wu * r2[0] + wd * r0[0]) + (l1 + r1[1])) - 4.0 * r1[0];
                const double a1 = ((wu * r2[1] + wd * r0[1]) + (wl1 * r1[0] + r1[2])) - 4.0 * r1[1];
                const double a2 = ((wu * r2[2] + wd * r0[2]) + (r1[1] + wr2 * r1[3])) - 4.0 * r1[2];
                const double a3 = ((wu * r2[3] + wd * r0[3]) + (r1[2] + q1)) - 4.0 * r1[3];
                double v0 = kp * a0, v1 = kp * a1, v2 = kp * a2, v3 = kp * a3;
                if (p.has_u) {   // ua is u[jo]
                    v0 += c_u * ((double)k.x * ((double)k.x * ((double)ua.x - (double)o.x)));
                    v1 += c_u * ((double)k.y * ((double)k.y * ((double)ua.y - (double)o.y)));
                    v2 += c_u * ((double)k.z * ((double)k.z * ((double)ua.z - (double)o.z)));
                    v3 += c_u * ((double)k.w * ((double)k.w * ((double)ua.w - (double)o.w)));
                }
                *reinterpret_cast<float4*>(gout + (int64_t)jo * p.W) = make_float4((float)v0, (float)v1, (float)v2, (float)v3);
                if (gdout)
                    *reinterpret_cast<float4*>(gdout + (int64_t)jo * p.W) =
                        make_float4((float)(c_p * r1[0]), (float)(c_p * r1[1]), (float)(c_p * r1[2]), (float)(c_p * r1[3]));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r2[i] = r1[i];
                r1[i] = r0[i];
            }
            ua = ub;
            ub = uc;
        }
    }
}
