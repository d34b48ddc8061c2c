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
//   * rows are prefetched with cp.async (LDGSTS) into a per-lane shared-memory ring of kRing (8) row elements:
//     each lane copies its own 16 bytes and reads only its own slots back, so no barrier is needed and prefetch
//     depth costs no registers (Little's law: ~50 KB must be in flight per SM to cover HBM latency at 6.5 TB/s;
//     16 warps x 3-4 rows x 52 B x 32 lanes = 80-106 KB).  One ring element holds u, dudt, obs and mask of the SAME
//     row, so one row-offset computation serves four copies; consumers read the three fields from elements
//     it+2 / it+1 / it.  Row indices outside the grid are REFLECTED (u[-1] = u[1]), which is exactly the
//     reference's padding, so the stencil needs no boundary selects; column indices of idle lanes are clamped and
//     their results multiplied by zero;
//   * the three-row windows are kept in fp64, so every loaded value is converted once (F2F.F64.F32 issues at a
//     quarter of the DFMA rate) and the uint8 mask is widened with the 2^52 trick (one DADD, no I2F);
//   * grids wider than 128 columns are cut into strips of 112 output columns + two halo lanes on each side, so that
//     every warp access starts on a 32-byte sector boundary (one halo lane would do for the stencil, but a strip
//     starting at column 120 k - 4 splits every 128-byte wavefront over two lines: +25 % sectors on the L2
//     crossbar, measured); narrow grids pack several row segments into one warp (W = 64: two, W = 16: eight).
// A lane spends ~30 instructions per pixel; arithmetic and accumulation stay fp64 (see guidance.cu header).
//
// Eligibility (checked on the host, else the generic tile kernel runs): fp32 fields, W % 4 == 0, 16-byte aligned
// base pointers and strides % 4 == 0, observations fp32, masks uint8 (bool).

struct MarchGeom {
    int lw_log2, segs_per_warp, strips, strip_w, halo_lane, R, chunks;
    int n_seg_items, n_warp_items;       // u-plane work: row segments, and warps' worth of them
    int a_blocks_per_plane, n_a_items;   // a-plane work: blocks of kABlock float4 within one plane's owned rows
    int a_plane4;                        // float4 per a-plane (owned rows)
    int a_block4;                        // float4 per a-plane work item (kABlock, smaller on small problems)
};

constexpr int kABlock = 1024;            // largest a-plane work item, in float4 (4096 pixels)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uchar4 ldg4(const unsigned char* p) { return __ldg(reinterpret_cast<const uchar4*>(p)); }

// Ring depth (rows per lane, power of two) and dynamic shared memory per CTA.  PA = 0: a-planes are separate
// streaming work items (8-deep ring of u / dudt / obs_u / mask_u: 106496 B);  PA = 1 / 2: every marching item also
// carries the a-plane of the same index (requires ch_a == number of u-planes), whose row, observation and mask
// (PA == 1; PA == 2: mask_a empty, nothing to read, ring as PA = 0) ride in the same ring element -- 4 deep, 90112 B.
// The reduce pass carries no residual window and no output pointers: it fits 80 registers, so with a 4-deep ring
// (53 KB) THREE CTAs share an SM (24 warps instead of 16) -- it is issue-bound, not bandwidth-bound.
__host__ __device__ constexpr int ring_depth(int PA, bool vjp) { return (PA == 1 || !vjp) ? 4 : 8; }
__host__ __device__ constexpr int ring_bytes(int PA, bool vjp) { return ring_depth(PA, vjp) * kThreads * (PA == 1 ? 5 * 16 + 2 * 4 : 3 * 16 + 4); }

// shared-memory operands are 32-bit shared-window addresses computed once per thread (the generic->shared
// conversion otherwise costs ~6 instructions per copy)
__device__ __forceinline__ void cp_async16(unsigned smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned smem) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned smem) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// exact uint8 -> double without the (quarter-rate) I2F.F64: 2^52 + k has k in its low mantissa bits
__device__ __forceinline__ double u8_to_double(unsigned k) { return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0; }

struct D4v {
    double v[4];
};
__device__ __forceinline__ D4v widen(const float4& f) { return D4v{{(double)f.x, (double)f.y, (double)f.z, (double)f.w}}; }

struct MarchLane {
    int b, cu, col0, ys, ye;
    bool lane_ok, out_ok, left_edge, right_edge;
};

__device__ __forceinline__ MarchLane march_decode(const Params& p, const MarchGeom& g, int wi, int lane) {
    const int LW = 1 << g.lw_log2, seg = lane >> g.lw_log2, l = lane & (LW - 1);
    unsigned si = (unsigned)wi * g.segs_per_warp + seg;
    const bool seg_ok = si < (unsigned)g.n_seg_items;
    if (!seg_ok) si = g.n_seg_items - 1;
    MarchLane m;
    unsigned t = si / (unsigned)p.B;
    m.b = (int)(si - t * p.B);
    unsigned t2 = t / (unsigned)p.n_u_units;
    m.cu = (int)(t - t2 * p.n_u_units);
    const unsigned chunk = t2 / (unsigned)g.strips;
    const int strip = (int)(t2 - chunk * g.strips);
    m.col0 = strip * g.strip_w - 4 * g.halo_lane + 4 * l;
    m.lane_ok = seg_ok && m.col0 >= 0 && m.col0 < p.W;
    m.out_ok = m.lane_ok && m.col0 >= strip * g.strip_w && m.col0 < (strip + 1) * g.strip_w;
    m.left_edge = (m.col0 == 0);
    m.right_edge = (m.col0 + 4 == p.W);
    m.ys = p.ylo + (int)chunk * g.R;
    m.ye = min(m.ys + g.R, p.yhi);
    return m;
}

// Local buffer offset (in elements) of row y, with the reference's reflect padding applied in GLOBAL row numbers:
// rows above the grid mirror about row 0, rows below about row Hg-1 (sample.py:132).  Clamped into the buffer.
__device__ __forceinline__ int row_offset(const Params& p, int y) {
    int gy = y + p.yg0;
    gy = gy < 0 ? -gy : gy;
    gy = gy > p.Hg - 1 ? 2 * (p.Hg - 1) - gy : gy;
    return min(max(gy - p.yg0, 0), p.H - 1) * p.W;
}

// Per-lane prefetch ring.  Element s holds u, dudt, obs and mask (and, paired, a / obs_a / mask_a) of row (row0 + s)
// for the lane's four columns.  HAS_D / HAS_O / PA are compile-time, so the loop carries no pointer tests.
template <bool HAS_D, bool HAS_O, int PA, bool VJP>
struct RowRing {
    static constexpr int RD = ring_depth(PA, VJP);
    unsigned su, sd, so, sm, sa, soa, sma;   // shared-window byte addresses of this lane's slot 0 in each field ring
    const float *u, *du, *ob, *a, *oa;
    const unsigned char *mk, *ma;
    int row0;
    int off_run;                             // fast items: element offset of the next ring element to issue
    bool fast;                               // every row this item touches is inside the grid and the buffer: no reflection

    static __device__ __forceinline__ unsigned slot16(int s) { return (unsigned)(s & (RD - 1)) * (kThreads * 16); }
    static __device__ __forceinline__ unsigned slot4(int s) { return (unsigned)(s & (RD - 1)) * (kThreads * 4); }

    // start the copies of element s (fields selected by warp-uniform flags); always commits exactly one group.
    // fo also selects the a-plane fields: they are consumed together with the observation of the same row.
    int colc;                                // the lane's first column (clamped to 0 for idle lanes)
    // Start of a work item whose ring elements 0 .. last cover rows first .. first + last.  Interior items (all but the
    // first and last chunk of a plane) need no reflection: the row offset just advances by W per element.
    __device__ __forceinline__ void begin_item(const Params& p, int first, int last) {
        row0 = first;
        const int g0 = first + p.yg0, g1 = first + last + p.yg0;
        fast = __all_sync(0xffffffffu, g0 >= 0 && g1 <= p.Hg - 1 && first >= 0 && first + last <= p.H - 1);
        off_run = first * p.W;
    }
    __device__ __forceinline__ void issue(const Params& p, int s, bool fu, bool fd, bool fo) {
        const int off = fast ? off_run : row_offset(p, row0 + s);
        off_run += p.W;
        if (fu) cp_async16(su + slot16(s), u + off);
        if (HAS_D && fd) cp_async16(sd + slot16(s), du + off);
        if (HAS_O && fo) {
            cp_async16(so + slot16(s), ob + off);
            cp_async4(sm + slot4(s), mk + off);
        }
        if (PA == 1 && fo) {
            cp_async16(sa + slot16(s), a + off);
            cp_async16(soa + slot16(s), oa + off);
            cp_async4(sma + slot4(s), ma + off);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ float4 direct_u(const Params& p, int y) const { return ldg4(u + row_offset(p, y)); }
    __device__ __forceinline__ float4 get_u(int s) const { return lds128(su + slot16(s)); }
    __device__ __forceinline__ float4 get_d(int s) const { return HAS_D ? lds128(sd + slot16(s)) : make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float4 get_o(int s) const { return HAS_O ? lds128(so + slot16(s)) : make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ unsigned get_m(int s) const { return HAS_O ? lds32(sm + slot4(s)) : 0u; }
    __device__ __forceinline__ float4 get_a(int s) const { return lds128(sa + slot16(s)); }
    __device__ __forceinline__ float4 get_oa(int s) const { return lds128(soa + slot16(s)); }
    __device__ __forceinline__ unsigned get_ma(int s) const { return lds32(sma + slot4(s)); }

    __device__ __forceinline__ void init(unsigned char* smem, int tid) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
        constexpr unsigned F16 = RD * kThreads * 16, F4 = RD * kThreads * 4;
        su = base + tid * 16;
        sd = su + F16;
        so = sd + F16;
        sa = so + F16;
        soa = sa + F16;
        const unsigned small = base + (PA == 1 ? 5 : 3) * F16 + tid * 4;
        sm = small;
        sma = small + F4;
    }
    // Bind the lane's column pointers of one work item.
    __device__ __forceinline__ void bind(const Params& p, const MarchLane& m, const float* x0, const float* dxp) {
        const int ch = p.ch_a + m.cu;
        colc = m.lane_ok ? m.col0 : 0;
        // every pointer already includes the lane's column, so one row offset serves all copies of an element
        u = x0 + (int64_t)m.b * p.x0.sb + (int64_t)ch * p.x0.sc + colc;
        du = HAS_D ? dxp + (int64_t)m.b * p.dxdt.sb + (int64_t)ch * p.dxdt.sc + colc : nullptr;
        ob = HAS_O ? reinterpret_cast<const float*>(p.obs_u.p) + (int64_t)m.b * p.obs_u.sb + (int64_t)m.cu * p.obs_u.sc + colc : nullptr;
        mk = HAS_O ? reinterpret_cast<const unsigned char*>(p.mask_u.p) + (int64_t)m.b * p.mask_u.sb + (int64_t)m.cu * p.mask_u.sc + colc : nullptr;
        if (PA == 1) {   // the a-plane paired with u-plane cu is a-channel cu
            a = x0 + (int64_t)m.b * p.x0.sb + (int64_t)m.cu * p.x0.sc + colc;
            oa = reinterpret_cast<const float*>(p.obs_a.p) + (int64_t)m.b * p.obs_a.sb + (int64_t)m.cu * p.obs_a.sc + colc;
            ma = reinterpret_cast<const unsigned char*>(p.mask_a.p) + (int64_t)m.b * p.mask_a.sb + (int64_t)m.cu * p.mask_a.sc + colc;
        }
    }
};

// unscaled 5-point sums of one row for the lane's four columns (rows already reflected by the loader)
__device__ __forceinline__ void lap_row(const D4v& up, const D4v& c, const D4v& dn, double lf, double rt, double* s) {
    s[0] = ((up.v[0] + dn.v[0]) + (lf + c.v[1])) - 4.0 * c.v[0];
    s[1] = ((up.v[1] + dn.v[1]) + (c.v[0] + c.v[2])) - 4.0 * c.v[1];
    s[2] = ((up.v[2] + dn.v[2]) + (c.v[1] + c.v[3])) - 4.0 * c.v[2];
    s[3] = ((up.v[3] + dn.v[3]) + (c.v[2] + rt)) - 4.0 * c.v[3];
}

// a-plane work item -> (batch, channel, first float4 of the block inside the plane's owned rows, float4 count)
struct AItem {
    int b, ch, first4, n4;
};
__device__ __forceinline__ AItem a_decode(const Params& p, const MarchGeom& g, int item) {
    // batch index innermost: the B samples of one block are processed back to back, so observation / mask
    // operands that broadcast over the batch are fetched from HBM once and then hit in L2
    const unsigned t = (unsigned)item / (unsigned)p.B;
    AItem a;
    a.b = (int)((unsigned)item - t * p.B);
    const unsigned blk = t / (unsigned)p.ch_a;
    a.ch = (int)(t - blk * p.ch_a);
    a.first4 = (int)blk * g.a_block4;
    a.n4 = min(g.a_block4, g.a_plane4 - a.first4);
    return a;
}

template <int J>
using IC = std::integral_constant<int, J>;

// compile-time loop: f(IC<0>{}), ..., f(IC<N-1>{})
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (N > 0) {
        static_for<N - 1>(f);
        f(IC<N - 1>{});
    }
}

// Is every row of [first, last] inside the local buffer and inside the global grid (no reflection, no clamp)?
__host__ __device__ __forceinline__ bool rows_inside(const Params& p, int first, int last) {
    return first >= 0 && last <= p.H - 1 && first + p.yg0 >= 0 && last + p.yg0 <= p.Hg - 1;
}

// ---- a-plane streaming items (sum (mask (a - obs))^2 and its gradient), shared by the heat and LLG kernels ---------
// Masks are 0 / 1, so an unobserved pixel is handled by selecting the OBSERVATION operand at 32 bits before it is
// widened: o' = mask ? obs : a makes the difference exactly zero -- one FSEL instead of widening the mask to fp64 and
// multiplying (the round-1 form cost 49 instructions per float4 in the reduce pass; this one ~30).
__device__ __forceinline__ float sel_obs(unsigned m, float o, float v) { return m ? o : v; }

__device__ __forceinline__ void a_item_reduce(const Params& p, const MarchGeom& g, int item, int lane, double& s_a) {
    const AItem a = a_decode(p, g, item);
    const int base = p.ylo * p.W + 4 * a.first4;
    const float* pa = reinterpret_cast<const float*>(p.x0.p) + (int64_t)a.b * p.x0.sb + (int64_t)a.ch * p.x0.sc + base;
    const float* po = reinterpret_cast<const float*>(p.obs_a.p) + (int64_t)a.b * p.obs_a.sb + (int64_t)a.ch * p.obs_a.sc + base;
    const unsigned char* pm = reinterpret_cast<const unsigned char*>(p.mask_a.p) + (int64_t)a.b * p.mask_a.sb + (int64_t)a.ch * p.mask_a.sc + base;
    double s0 = 0.0, s1 = 0.0;
    auto body = [&](int i) {
        const float4 v = ldg4(pa + 4 * i), o = ldg4(po + 4 * i);
        const unsigned k = __ldg(reinterpret_cast<const unsigned*>(pm + 4 * i));
        const double d0 = (double)v.x - (double)(sel_obs(k & 0xffu, o.x, v.x)), d1 = (double)v.y - (double)(sel_obs(k & 0xff00u, o.y, v.y));
        const double d2 = (double)v.z - (double)(sel_obs(k & 0xff0000u, o.z, v.z)), d3 = (double)v.w - (double)(sel_obs(k & 0xff000000u, o.w, v.w));
        s0 = fma(d0, d0, s0);
        s1 = fma(d1, d1, s1);
        s0 = fma(d2, d2, s0);
        s1 = fma(d3, d3, s1);
    };
    int i0 = 0;
    for (; i0 + 256 <= a.n4; i0 += 256) {   // full blocks: eight independent 128-bit loads per array at constant offsets
#pragma unroll
        for (int t = 0; t < 8; ++t) body(i0 + 32 * t + lane);
    }
    if (i0 < a.n4) {                         // ragged tail / small items: the same eight loads, predicated, still issued together
#pragma unroll
        for (int t = 0; t < 8; ++t)
            if (i0 + 32 * t + lane < a.n4) body(i0 + 32 * t + lane);
    }
    s_a += s0 + s1;
}

__device__ __forceinline__ void a_item_vjp(const Params& p, const MarchGeom& g, int item, int lane, double c_a,
                                           float* __restrict__ g_x0, float* __restrict__ g_dxdt) {
    const AItem a = a_decode(p, g, item);
    const int base = p.ylo * p.W + 4 * a.first4;
    const int64_t plane = (int64_t)p.H * p.W;
    float* pg = g_x0 + ((int64_t)a.b * p.C + a.ch) * plane + base;
    float* pgd = g_dxdt ? g_dxdt + ((int64_t)a.b * p.C + a.ch) * plane + base : nullptr;
    if (p.has_a) {   // g = c_a mask (mask (a - obs)); zeros when the mask is empty (sample.py:337-342)
        const float* pa = reinterpret_cast<const float*>(p.x0.p) + (int64_t)a.b * p.x0.sb + (int64_t)a.ch * p.x0.sc + base;
        const float* po = reinterpret_cast<const float*>(p.obs_a.p) + (int64_t)a.b * p.obs_a.sb + (int64_t)a.ch * p.obs_a.sc + base;
        const unsigned char* pm = reinterpret_cast<const unsigned char*>(p.mask_a.p) + (int64_t)a.b * p.mask_a.sb + (int64_t)a.ch * p.mask_a.sc + base;
        auto body = [&](int i) {
            const float4 v = ldg4(pa + 4 * i), o = ldg4(po + 4 * i);
            const unsigned k = __ldg(reinterpret_cast<const unsigned*>(pm + 4 * i));
            float4 w;
            w.x = (float)(c_a * ((double)v.x - (double)(sel_obs(k & 0xffu, o.x, v.x))));
            w.y = (float)(c_a * ((double)v.y - (double)(sel_obs(k & 0xff00u, o.y, v.y))));
            w.z = (float)(c_a * ((double)v.z - (double)(sel_obs(k & 0xff0000u, o.z, v.z))));
            w.w = (float)(c_a * ((double)v.w - (double)(sel_obs(k & 0xff000000u, o.w, v.w))));
            *reinterpret_cast<float4*>(pg + 4 * i) = w;
        };
        int i0 = 0;
        for (; i0 + 256 <= a.n4; i0 += 256) {
#pragma unroll
            for (int t = 0; t < 8; ++t) body(i0 + 32 * t + lane);
        }
        if (i0 < a.n4) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (i0 + 32 * t + lane < a.n4) body(i0 + 32 * t + lane);
        }
    } else {
        for (int i = lane; i < a.n4; i += 32) *reinterpret_cast<float4*>(pg + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (pgd)
        for (int i = lane; i < a.n4; i += 32) *reinterpret_cast<float4*>(pgd + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- the same items streamed through the (otherwise idle) per-lane cp.async ring of the marching kernels --------------
// ncu of the LDG form inside the marching kernels: a third of all stall samples sat on the first use of an a-plane load
// (each warp loads a block, waits ~1-2 us for it, computes, loads the next: nothing in flight while it computes).  Here
// every lane keeps D float4 triples (a, obs, mask) in flight all the time: slot j is refilled with element t + D as soon as
// element t has been read, exactly like the row ring of the u items.  NT = threads per CTA; `sbase` = shared-window address
// of the ring memory (needs NT * D * 36 bytes).
// Slot ownership: the ring memory is shared by all warps of the CTA while OTHER warps run u items, so a lane may only use
// bytes that belong to it in the u-item ring too.  KF = index of the 16-byte field whose region the 4-byte mask slots
// follow: the heat rings (RowRing: u | dudt | obs | masks, D = ring depth) map a -> u slots, obs -> dudt slots, mask -> mask
// slots with KF = 3; a dedicated region (LLG kernels) packs a | obs | mask with KF = 2.
template <int D, int NT, int KF>
struct ARing {
    unsigned a0, o0, k0;
    // `t` = index of the thread among the NT threads that share the region (NT == 32: a warp-private region, t = lane)
    __device__ __forceinline__ ARing(unsigned sbase, int t) : a0(sbase + t * 16), o0(a0 + D * NT * 16), k0(sbase + KF * D * NT * 16 + t * 4) {}
    __device__ __forceinline__ unsigned A(int s) const { return a0 + (unsigned)(s % D) * (NT * 16); }
    __device__ __forceinline__ unsigned O(int s) const { return o0 + (unsigned)(s % D) * (NT * 16); }
    __device__ __forceinline__ unsigned K(int s) const { return k0 + (unsigned)(s % D) * (NT * 4); }
};

template <int D, int NT, int KF, typename F>
__device__ __forceinline__ void a_item_stream(const AItem& a, int lane, unsigned sbase, const float* pa, const float* po, const unsigned char* pm,
                                              F&& consume) {
    const ARing<D, NT, KF> r(sbase, NT == 32 ? lane : (int)threadIdx.x);
    const int T = (a.n4 + 31) >> 5;                                   // iterations of the warp; lane's element of iteration t: lane + 32 t
    auto issue = [&](int slot, int t) {
        const int i = lane + 32 * t;
        if (i < a.n4) {
            cp_async16(r.A(slot), pa + 4 * i);
            cp_async16(r.O(slot), po + 4 * i);
            cp_async4(r.K(slot), pm + 4 * i);
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
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), o = v;
            unsigned k = 0u;
            if (i < a.n4) {
                v = lds128(r.A(j));
                o = lds128(r.O(j));
                k = lds32(r.K(j));
            }
            issue(j, t + D);                                          // element t + D -> the slot just drained
            if (i < a.n4) consume(i, v, o, k);
        });
    }
    cp_async_wait<0>();
}

// (The reduce passes keep the LDG form: streamed through a ring they were slower in every kernel measured -- heat 0.370 vs 0.320 ms,
//  LLG residual 0.377 vs 0.352 ms, LLG soft norm 0.263 vs 0.210 ms; the VJP passes, which also write, prefer the ring.)
template <int D, int NT, int KF>
__device__ __forceinline__ void a_item_vjp_ring(const Params& p, const MarchGeom& g, int item, int lane, unsigned sbase, double c_a,
                                                float* __restrict__ g_x0, float* __restrict__ g_dxdt) {
    if (!p.has_a) {                                                   // empty mask: zeros, nothing to read (sample.py:337-342)
        a_item_vjp(p, g, item, lane, c_a, g_x0, g_dxdt);
        return;
    }
    const AItem a = a_decode(p, g, item);
    const int base = p.ylo * p.W + 4 * a.first4;
    const int64_t plane = (int64_t)p.H * p.W;
    float* pg = g_x0 + ((int64_t)a.b * p.C + a.ch) * plane + base;
    float* pgd = g_dxdt ? g_dxdt + ((int64_t)a.b * p.C + a.ch) * plane + base : nullptr;
    const float* pa = reinterpret_cast<const float*>(p.x0.p) + (int64_t)a.b * p.x0.sb + (int64_t)a.ch * p.x0.sc + base;
    const float* po = reinterpret_cast<const float*>(p.obs_a.p) + (int64_t)a.b * p.obs_a.sb + (int64_t)a.ch * p.obs_a.sc + base;
    const unsigned char* pm = reinterpret_cast<const unsigned char*>(p.mask_a.p) + (int64_t)a.b * p.mask_a.sb + (int64_t)a.ch * p.mask_a.sc + base;
    a_item_stream<D, NT, KF>(a, lane, sbase, pa, po, pm, [&](int i, const float4& v, const float4& o, unsigned k) {
        float4 w;
        w.x = (float)(c_a * ((double)v.x - (double)(sel_obs(k & 0xffu, o.x, v.x))));
        w.y = (float)(c_a * ((double)v.y - (double)(sel_obs(k & 0xff00u, o.y, v.y))));
        w.z = (float)(c_a * ((double)v.z - (double)(sel_obs(k & 0xff0000u, o.z, v.z))));
        w.w = (float)(c_a * ((double)v.w - (double)(sel_obs(k & 0xff000000u, o.w, v.w))));
        *reinterpret_cast<float4*>(pg + 4 * i) = w;
    });
    if (pgd)
        for (int i = lane; i < a.n4; i += 32) *reinterpret_cast<float4*>(pgd + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Each warp owns every nwarps-th u-item (compute/issue bound) and a-item (pure streaming, latency bound) and
// alternates between the two kinds in proportion, odd warps starting with the other kind, so that at any time an
// SM runs a mix of both and the streaming warps' memory stalls overlap the marching warps' arithmetic.
template <typename FU, typename FA>
__device__ __forceinline__ void run_interleaved(int warp0, int nwarps, int n_u, int n_a, int a_first, FU&& do_u, FA&& do_a) {
    const int nu_w = n_u > warp0 ? (n_u - warp0 + nwarps - 1) / nwarps : 0;
    const int na_w = n_a > warp0 ? (n_a - warp0 + nwarps - 1) / nwarps : 0;
    int ju = 0, ja = 0;
    while (ju < nu_w || ja < na_w) {
        const long long lhs = (long long)ju * na_w, rhs = (long long)ja * nu_w;
        const bool take_u = ju < nu_w && (ja >= na_w || (a_first ? lhs < rhs : lhs <= rhs));
        if (take_u) {
            do_u(warp0 + ju * nwarps);
            ++ju;
        } else {
            do_a(warp0 + ja * nwarps);
            ++ja;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass 1 (fast): S_a, S_u, S_pde
// ---------------------------------------------------------------------------------------------------------
// PS (per-sample mode, training loss models/loss.py:143): no global sums -- every row-segment item writes its own
// sum of squared residuals to partials[item] (items of sample b are b, b + B, b + 2 B, ...: per_sample_items_kernel adds
// them in that order), nothing else is touched.
template <bool HAS_D, bool HAS_O, int PA, bool PS = false>
__global__ void __launch_bounds__(kThreads, 3)
heat_march_reduce_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g,
                         double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ sums,
                         int finalize, double* __restrict__ scal, float* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    __shared__ double scratch[3 * (kThreads / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const float* x0 = reinterpret_cast<const float*>(p.x0.p);
    const float* dxp = reinterpret_cast<const float*>(p.dxdt.p);
    double s_a = 0.0, s_u = 0.0, s_p = 0.0;
    const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);

    // ---- a-planes: sum (mask (a - obs))^2; a warp streams one block of a plane with 128-bit loads
    // (LDG form: this pass's ring is 4 deep, shallower than the eight loads per lane the LDG form keeps in flight -- measured
    //  0.370 ms streamed through the ring vs 0.320 ms)
    auto do_a = [&](int item) { a_item_reduce(p, g, item, lane, s_a); };

    // ---- u-planes: iteration `it` handles row j = ys + it with the window ua = u[j-1], ub = u[j], uc = u[j+1].
    //      Ring element s is row ys + s:  uc comes from element it+1, dudt / obs / mask from element it.
    const int LW = 1 << g.lw_log2;
    RowRing<HAS_D, HAS_O, PA, false> ring;
    constexpr int kRing = ring_depth(PA, false);
    ring.init(ring_mem, tid);
    auto do_u = [&](int wi) {
        const MarchLane m = march_decode(p, g, wi, lane);
        ring.bind(p, m, x0, dxp);
        const int n_it = g.R, n_el = g.R + 1;
        ring.begin_item(p, m.ys, n_el - 1);
#pragma unroll
        for (int s = 0; s < kRing; ++s) ring.issue(p, s, s >= 1 && s < n_el, s < n_it, s < n_it);
        const double a_s = __ldg(p.coef + m.b) * p.inv_dx2;
        D4v ua = widen(ring.direct_u(p, m.ys - 1)), ub = widen(ring.direct_u(p, m.ys));
#pragma unroll 3
        for (int it = 0; it < n_it; ++it) {
            cp_async_wait<kRing - 2>();                          // elements <= it + 1 have landed
            const D4v uc = widen(ring.get_u(it + 1));
            const float4 dt = ring.get_d(it), o = ring.get_o(it);
            unsigned k = ring.get_m(it);
            float4 av, oav;
            unsigned ka = 0u;
            if (PA == 1) {
                av = ring.get_a(it);
                oav = ring.get_oa(it);
                ka = ring.get_ma(it);
            }
            const int sn = it + kRing;                           // refill the slot just drained
            ring.issue(p, sn, sn < n_el, sn < n_it, sn < n_it);
            double lf = __shfl_up_sync(0xffffffffu, ub.v[3], 1, LW), rt = __shfl_down_sync(0xffffffffu, ub.v[0], 1, LW);
            if (m.left_edge) lf = ub.v[1];                       // reflect: u[-1] = u[1]
            if (m.right_edge) rt = ub.v[2];
            const bool ok = m.out_ok && m.ys + it < m.ye;
            double s[4];
            lap_row(ua, ub, uc, lf, rt, s);
            const double r0 = (double)dt.x - a_s * s[0], r1 = (double)dt.y - a_s * s[1];
            const double r2 = (double)dt.z - a_s * s[2], r3 = (double)dt.w - a_s * s[3];
            const double okf = ok ? 1.0 : 0.0;
            s_p += okf * ((r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3));
            if (HAS_O) {
                if (!ok) k = 0u;
                const double d0 = u8_to_double(k & 255u) * (ub.v[0] - (double)o.x), d1 = u8_to_double((k >> 8) & 255u) * (ub.v[1] - (double)o.y);
                const double d2 = u8_to_double((k >> 16) & 255u) * (ub.v[2] - (double)o.z), d3 = u8_to_double(k >> 24) * (ub.v[3] - (double)o.w);
                s_u += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            if (PA == 1) {   // paired a-plane row: sum (mask (a - obs))^2
                if (!ok) ka = 0u;
                const double d0 = u8_to_double(ka & 255u) * ((double)av.x - (double)oav.x), d1 = u8_to_double((ka >> 8) & 255u) * ((double)av.y - (double)oav.y);
                const double d2 = u8_to_double((ka >> 16) & 255u) * ((double)av.z - (double)oav.z), d3 = u8_to_double(ka >> 24) * ((double)av.w - (double)oav.w);
                s_a += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            ua = ub;
            ub = uc;
        }
        cp_async_wait<0>();
        if (PS) {   // segmented (LW-lane) butterfly, lane 0 of every segment owns the item's sum
            double v = s_p;
            for (int o = LW >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LW);
            const unsigned si = (unsigned)wi * g.segs_per_warp + (lane >> g.lw_log2);
            if ((lane & (LW - 1)) == 0 && si < (unsigned)g.n_seg_items) partials[si] = v;
            s_p = 0.0;
        }
    };
    run_interleaved(warp0, nwarps, g.n_warp_items, (PA == 0 && p.has_a) ? g.n_a_items : 0, (tid >> 5) & 1, do_u, do_a);
    if (PS) return;

    reduce_epilogue(p, s_a, s_u, s_p, scratch, &is_last, partials, ticket, sums, finalize, scal, trace);
}

// ---------------------------------------------------------------------------------------------------------
// pass 2 (fast): seed gradient
// ---------------------------------------------------------------------------------------------------------
// PS (per-sample mode): `upstream` holds one seed per sample, c_p of an item = 2 upstream[b] (d r^2 / d r), no
// observation terms, `scal` is not read.
template <bool HAS_D, bool HAS_O, int PA, bool PS = false>
__global__ void __launch_bounds__(kThreads, 2)
heat_march_vjp_kernel(const __grid_constant__ Params p, const __grid_constant__ MarchGeom g, const double* __restrict__ scal,
                      const double* __restrict__ upstream, float* __restrict__ g_x0, float* __restrict__ g_dxdt) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const double up = (!PS && upstream) ? __ldg(upstream) : 1.0;
    const double c_a = PS ? 0.0 : __ldg(scal + 4) * up, c_u = PS ? 0.0 : __ldg(scal + 5) * up;
    double c_p = PS ? 0.0 : __ldg(scal + 6) * up;
    const float* x0 = reinterpret_cast<const float*>(p.x0.p);
    const float* dxp = reinterpret_cast<const float*>(p.dxdt.p);
    const int64_t plane = (int64_t)p.H * p.W;
    const int warp0 = blockIdx.x * (kThreads / 32) + (tid >> 5), nwarps = gridDim.x * (kThreads / 32);

    // ---- a-planes: g = c_a mask (mask (a - obs)), zeros when the mask is empty (sample.py:337-342)
    auto do_a = [&](int item) { a_item_vjp_ring<ring_depth(0, true), kThreads, 3>(p, g, item, lane, (unsigned)__cvta_generic_to_shared(ring_mem), c_a, g_x0, g_dxdt); };

    // ---- u-planes.  Iteration `it` computes the residual of row j = ys - 1 + it (window ua = u[j-1], ub = u[j],
    //      uc = u[j+1]) and then emits the gradient of row jo = j - 1 from r2 = r[jo-1], r1 = r[jo], r0 = r[jo+1].
    //      Ring element s is row ys - 2 + s:  uc = element it+2, dudt[j] = element it+1, obs/mask[jo] = element it.
    const int LW = 1 << g.lw_log2;
    RowRing<HAS_D, HAS_O, PA, true> ring;
    constexpr int kRing = ring_depth(PA, true);
    ring.init(ring_mem, tid);
    auto do_u = [&](int wi) {
        const MarchLane m = march_decode(p, g, wi, lane);
        ring.bind(p, m, x0, dxp);
        const int n_it = g.R + 2;
        ring.begin_item(p, m.ys - 2, n_it + 1);
        // fields of element s that are consumed: u for s in [2, n_it+2), dudt for s in [1, n_it+1), obs for s in [2, n_it)
#pragma unroll
        for (int s = 0; s < kRing; ++s) ring.issue(p, s, s >= 2 && s < n_it + 2, s >= 1 && s < n_it + 1, s >= 2 && s < n_it);
        const int ch = p.ch_a + m.cu, colc = ring.colc;
        float* gout = g_x0 + ((int64_t)m.b * p.C + ch) * plane;                                   // (+ colc at the store)
        float* gdout = g_dxdt ? g_dxdt + ((int64_t)m.b * p.C + ch) * plane : nullptr;
        float* gaout = PA ? g_x0 + ((int64_t)m.b * p.C + m.cu) * plane : nullptr;                 // paired a-plane
        float* gadout = (PA && g_dxdt) ? g_dxdt + ((int64_t)m.b * p.C + m.cu) * plane : nullptr;
        if (PS) c_p = 2.0 * __ldg(upstream + m.b);
        const double a_s = __ldg(p.coef + m.b) * p.inv_dx2, kp = -c_p * a_s;
        const double wl1 = m.left_edge ? 2.0 : 1.0, wr2 = m.right_edge ? 2.0 : 1.0;   // transposed-stencil edge weights
        D4v ua = widen(ring.direct_u(p, m.ys - 2)), ub = widen(ring.direct_u(p, m.ys - 1));
        double r2[4] = {0.0, 0.0, 0.0, 0.0}, r1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 3
        for (int it = 0; it < n_it; ++it) {
            const int j = m.ys - 1 + it;
            cp_async_wait<kRing - 3>();                          // elements <= it + 2 have landed
            const D4v uc = widen(ring.get_u(it + 2));
            const float4 dt = ring.get_d(it + 1), o = ring.get_o(it);
            const unsigned k = ring.get_m(it);
            float4 av, oav;
            unsigned ka = 0u;
            if (PA == 1) {
                av = ring.get_a(it);
                oav = ring.get_oa(it);
                ka = ring.get_ma(it);
            }
            const int sn = it + kRing;
            ring.issue(p, sn, sn < n_it + 2, sn < n_it + 1, sn < n_it);
            double lf = __shfl_up_sync(0xffffffffu, ub.v[3], 1, LW), rt = __shfl_down_sync(0xffffffffu, ub.v[0], 1, LW);
            if (m.left_edge) lf = ub.v[1];
            if (m.right_edge) rt = ub.v[2];
            // r of rows outside the grid and of idle lanes is garbage (finite: it is computed from real, clamped
            // data); it is never consumed: the vertical weights below vanish for out-of-grid neighbours and the
            // edge lanes zero their horizontal neighbour.
            double s[4], r0[4];
            lap_row(ua, ub, uc, lf, rt, s);
            r0[0] = (double)dt.x - a_s * s[0];
            r0[1] = (double)dt.y - a_s * s[1];
            r0[2] = (double)dt.z - a_s * s[2];
            r0[3] = (double)dt.w - a_s * s[3];

            double l1 = __shfl_up_sync(0xffffffffu, r1[3], 1, LW), q1 = __shfl_down_sync(0xffffffffu, r1[0], 1, LW);
            if (m.left_edge) l1 = 0.0;
            if (m.right_edge) q1 = 0.0;
            const int jo = j - 1;
            if (it >= 2 && jo < m.ye && m.out_ok) {
                const int gjo = jo + p.yg0;
                // transposed-stencil weights of the rows above / below: 2 from a boundary row, 0 from outside
                const double wu = gjo == 0 ? 0.0 : (gjo == 1 ? 2.0 : 1.0);
                const double wd = gjo == p.Hg - 1 ? 0.0 : (gjo == p.Hg - 2 ? 2.0 : 1.0);
                const double a0 = ((wu * r2[0] + wd * r0[0]) + (l1 + r1[1])) - 4.0 * r1[0];
                const double a1 = ((wu * r2[1] + wd * r0[1]) + (wl1 * r1[0] + r1[2])) - 4.0 * r1[1];
                const double a2 = ((wu * r2[2] + wd * r0[2]) + (r1[1] + wr2 * r1[3])) - 4.0 * r1[2];
                const double a3 = ((wu * r2[3] + wd * r0[3]) + (r1[2] + q1)) - 4.0 * r1[3];
                double v0 = kp * a0, v1 = kp * a1, v2 = kp * a2, v3 = kp * a3;
                if (HAS_O) {   // ua is u[jo]
                    const double m0 = u8_to_double(k & 255u), m1 = u8_to_double((k >> 8) & 255u), m2 = u8_to_double((k >> 16) & 255u), m3 = u8_to_double(k >> 24);
                    v0 += c_u * (m0 * (m0 * (ua.v[0] - (double)o.x)));
                    v1 += c_u * (m1 * (m1 * (ua.v[1] - (double)o.y)));
                    v2 += c_u * (m2 * (m2 * (ua.v[2] - (double)o.z)));
                    v3 += c_u * (m3 * (m3 * (ua.v[3] - (double)o.w)));
                }
                *reinterpret_cast<float4*>((gout + (int64_t)jo * p.W) + colc) = make_float4((float)v0, (float)v1, (float)v2, (float)v3);
                if (gdout)
                    *reinterpret_cast<float4*>((gdout + (int64_t)jo * p.W) + colc) =
                        make_float4((float)(c_p * r1[0]), (float)(c_p * r1[1]), (float)(c_p * r1[2]), (float)(c_p * r1[3]));
                if (PA) {   // paired a-plane, row jo: g = c_a mask (mask (a - obs)); zeros when mask_a is empty
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (PA == 1) {
                        const double m0 = u8_to_double(ka & 255u), m1 = u8_to_double((ka >> 8) & 255u), m2 = u8_to_double((ka >> 16) & 255u), m3 = u8_to_double(ka >> 24);
                        w.x = (float)(c_a * (m0 * (m0 * ((double)av.x - (double)oav.x))));
                        w.y = (float)(c_a * (m1 * (m1 * ((double)av.y - (double)oav.y))));
                        w.z = (float)(c_a * (m2 * (m2 * ((double)av.z - (double)oav.z))));
                        w.w = (float)(c_a * (m3 * (m3 * ((double)av.w - (double)oav.w))));
                    }
                    *reinterpret_cast<float4*>((gaout + (int64_t)jo * p.W) + colc) = w;
                    if (gadout) *reinterpret_cast<float4*>((gadout + (int64_t)jo * p.W) + colc) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r2[i] = r1[i];
                r1[i] = r0[i];
            }
            ua = ub;
            ub = uc;
        }
        cp_async_wait<0>();
    };
    run_interleaved(warp0, nwarps, g.n_warp_items, PA == 0 ? g.n_a_items : 0, (tid >> 5) & 1, do_u, do_a);
}
