/*
 * dpde_b200 -- C ABI of the B200-native physics-guided sampler step.
 *
 * Drop-in boundary for ONE hot path of cmt-dtu-energy/dynamical-pde-diffusion: the guided EDM Heun step of
 * JointSampler.sample (src/diffusion_pde/sampling/sample.py:320-357).  The reference is pure PyTorch and has no
 * FFI of its own; each entry point below states the reference lines whose math it replaces.  A binding is a
 * ctypes / cffi stub; dynamical_pde_diffusion_b200/_ffi.py is the one we ship.
 *
 * Conventions
 *   - The caller owns every buffer; all pointers are DEVICE pointers unless stated otherwise.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); nothing synchronises the host.
 *   - Return value: 0 on success, negative DPDE_ERR_* otherwise; dpde_last_error() gives the message of the
 *     calling thread's last failure.  No exceptions cross the boundary.
 *   - Fields are NCHW with contiguous rows (stride_w == 1, stride_h == W); batch and channel strides are free
 *     (in elements) so channel-slice views such as x_N[:, ch_a:] (sample.py:345-346) pass without a copy, and a
 *     stride of 0 broadcasts (masks (H,W), observations (1,ch,H,W): sample.py:340-342, model_testing.py:174-193).
 *   - Reductions are deterministic: per-CTA partial sums are combined in a fixed order.
 */
#ifndef DPDE_B200_H
#define DPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPDE_ABI_VERSION 2

enum dpde_error { DPDE_OK = 0, DPDE_ERR_INVALID = -1, DPDE_ERR_CUDA = -2, DPDE_ERR_UNSUPPORTED = -3 };
enum dpde_dtype { DPDE_F32 = 0, DPDE_F64 = 1, DPDE_U8 = 2 };

/* PDE residual plugged into the sampler (the reference's loss_fn slot, sample.py:347). */
enum dpde_pde_kind {
    DPDE_PDE_NONE = 0,
    DPDE_PDE_HEAT = 1,         /* heat_loss2,  pde_losses.py:71-96  + laplacian, sample.py:106-134       */
    DPDE_PDE_LLG_NORM = 2,     /* llg_loss2,   pde_losses.py:99-117 (soft |m| = 1)                        */
    DPDE_PDE_LLG_RESIDUAL = 3  /* m x H_eff residual, tests/test_llg_pde_loss.py:70-117 (no demag)        */
};

typedef void* dpde_stream_t; /* cudaStream_t */

/* A (B, ch, H, W) operand; rows contiguous.  ptr == NULL means "absent" (zeros for dxdt, unused for obs/mask). */
typedef struct dpde_view {
    const void* ptr;
    int32_t dtype;    /* dpde_dtype */
    int32_t _pad;
    int64_t stride_b; /* elements between samples;  0 broadcasts over the batch   */
    int64_t stride_c; /* elements between channels; 0 broadcasts over channels    */
} dpde_view;

/* One guidance evaluation: everything sample.py:336-353 reads. */
typedef struct dpde_guidance_desc {
    int32_t B, C, ch_a, H, W;
    int32_t pde_kind;          /* dpde_pde_kind; acts on channels [ch_a, C)                               */
    int32_t has_a, has_u;      /* mask_a.sum() > 0 / mask_u.sum() > 0 (sample.py:339,341), decided once   */
    /* Row-slab domain decomposition (large grids; not in the reference).  slab_H_global == 0: the fields hold the
       whole grid.  Otherwise every field is a local buffer of H rows = slab_halo ghost rows + owned rows +
       slab_halo ghost rows, the first owned row is global row slab_row0 of a grid slab_H_global rows high; sums
       and gradients cover owned rows only (ghost rows are read, never written).  The heat VJP needs
       slab_halo >= 2, LLG_RESIDUAL >= 2, others >= 0. */
    int32_t slab_halo, slab_row0, slab_H_global;
    int32_t _pad;
    dpde_view x0;              /* denoised estimate x_N (B,C,H,W), F32 or F64                             */
    dpde_view dxdt;            /* its time derivative, same dtype; ptr NULL = zeros (X_and_dXdt_dummy)    */
    dpde_view obs_a, mask_a;   /* (.., ch_a, H, W) broadcastable; obs F32/F64; mask F32/F64 (a weight, multiplied
                                  as the reference does) or U8 = a boolean mask whose bytes are 0 or 1       */
    dpde_view obs_u, mask_u;   /* (.., C-ch_a, H, W)                                                      */
    const double* sample_coef; /* HEAT: alpha_b = labels[b,-1] (B,);  LLG_RESIDUAL: h_ext in A/m (B,3)    */
    double dx;                 /* grid spacing (square cells, sample.py:133)                              */
    double w_a, w_u, w_pde;    /* guidance weights of this step (sample.py:348-351)                       */
    /* LLG residual constants (tests/test_llg_pde_loss.py:36-41): r = dmdt - tau (-gamma m x H - alpha m x (m x H)),
       H = h_ext + c_ex lap(m) + c_an (m . e) e */
    double gamma, alpha, c_ex, c_an, tau;
    double easy_axis[3];
} dpde_guidance_desc;

/* scalars[8] written by reduce/finalize: loss_a, loss_u, loss_pde, loss_comb, then the three seed coefficients
   c_a = w_a/loss_a, c_u = w_u/loss_u, c_pde (kind dependent), and one spare. */
#define DPDE_NUM_SCALARS 8

int dpde_abi_version(void);
const char* dpde_last_error(void);

/* Kernel selection: 1 (default) lets eligible problems (fp32 fields, W % 4 == 0, 16-byte aligned operands, fp32
   observations, uint8 masks) take the fast paths -- register row-marching kernels for the heat residual, convert-once
   shared-memory tiles for the LLG m x H_eff residual, 128-bit streaming for the soft-norm loss; 0 forces the generic
   tile kernels, which accept every layout.  Both give the same results (tests run both).  Returns the previous setting.
   TEST / TUNING HOOK: process-wide (an atomic, read once per launch); set it before the launching threads start. */
int dpde_set_fast_path(int enable);

/* Experiment knobs of the row-marching kernels (results never change, only speed): key 0 strip layout (0 (default) =
   per pass: 120 columns + 1 halo lane in the reduce pass, 112 + 2 sector aligned in the VJP; 1 / 2 force one of them), key 2 rows per chunk (0 = automatic: up to 128 in the
   VJP, 64 in the reduce pass), keys 3 / 4 = 1 pair every a-plane with the u-plane of the same index in the reduce /
   VJP pass instead of streaming it as separate work items, key 5 = 1 sends the interior work items of the LLG marching kernels
   through their general loop (A/B measurements of the lean loop), key 6 selects the LLG m x H_eff kernels: 0 (default) =
   row-marching kernels on large grids (W >= 128, >= 2 Mi pixels), convert-once tiles otherwise; 1 = tiles always; 2 = marching
   whenever W >= 128; key 7 = 1 runs the LLG marching kernels without TMA (cp.async.bulk.tensor): reduce pass with the cp.async feed, VJP as the two-CTA
   kernel with register windows instead of the three-CTA kernel (the TMA forms are the default wherever the tensor maps can be encoded).  TEST / TUNING HOOK like dpde_set_fast_path: process-wide atomics; the per-stream thread-safety of the
   compute entry points does not extend to changing these concurrently. */
int dpde_set_tuning(int key, int value);

/* Bytes of scratch the reduce pass needs (per-CTA partial sums + a ticket counter).  The caller zero-fills it
   once after allocation; the library leaves it zeroed-where-needed after every call. */
size_t dpde_guidance_workspace_bytes(void);

/* Pass 1 -- the three global sums of sample.py:340-342 and of loss_fn (pde_losses.py:94,116):
   sums[0] = sum (mask_a (a - obs_a))^2, sums[1] = same for u, sums[2] = sum r^2 (HEAT, LLG_RESIDUAL) or
   sum (1-|m|)^2 (LLG_NORM).  With finalize != 0 the last CTA also runs dpde_guidance_finalize's arithmetic. */
int dpde_guidance_reduce(const dpde_guidance_desc* desc, void* workspace, double* sums, int finalize,
                         double* scalars, float* trace_row, dpde_stream_t stream);

/* sums -> losses, loss_comb (sample.py:353) and seed coefficients; trace_row (4 floats, may be NULL) receives
   [loss_a, loss_u, loss_pde, loss_comb] as sample.py:357 stores them.  Separate entry point so a multi-GPU caller
   can all-reduce `sums` first (batch-coupled semantics / row slabs). */
int dpde_guidance_finalize(const dpde_guidance_desc* desc, const double* sums, double* scalars, float* trace_row,
                           dpde_stream_t stream);

/* ---- Cross-rank exchange of the three partial sums through peer-memory mailboxes (row slabs, coupled batch shards; new:
   the reference is single-process).  Every rank owns a zero-initialised DPDE_MAILBOX_BYTES mailbox in memory from
   dpde_peer_alloc and maps the other ranks' mailboxes (dpde_peer_export / dpde_peer_open); boxes[r] is rank r's mailbox
   as seen from THIS process (boxes[rank] is the local one).  `epoch` is the step number, >= 1, the same on every rank
   and increasing by one per exchange: slot [epoch & 1][source rank] is reused two exchanges later, which the protocol
   itself makes safe (a rank reaches exchange e + 2 only after every rank has finished reading exchange e). */
#define DPDE_MAX_RANKS 8
#define DPDE_MAILBOX_BYTES 512
typedef struct dpde_mailbox {
    int32_t world, rank;
    uint64_t epoch;
    void* boxes[DPDE_MAX_RANKS];
} dpde_mailbox;

/* dpde_guidance_reduce fused with the first half of the exchange: the last CTA of the reduce pass stores this rank's
   three sums into slot [epoch & 1][rank] of EVERY rank's mailbox (plain stores through the mapped peer pointers over
   NVLink) and publishes them with a system-scope release store on the slot's flag.  `sums` receives this rank's
   partial sums as dpde_guidance_reduce would write them; nothing is finalised. */
int dpde_guidance_reduce_post(const dpde_guidance_desc* desc, void* workspace, double* sums, const dpde_mailbox* mbox,
                              dpde_stream_t stream);

/* Second half: blocks the stream until all `world` slots of the local mailbox carry `epoch` (acquire, system scope; after
   timeout_s seconds it gives up and writes 1 to *status), adds them in rank order -- every rank forms the same total --
   into sums[3] and runs dpde_guidance_finalize's arithmetic on it. */
int dpde_mailbox_wait_finalize(const dpde_guidance_desc* desc, const dpde_mailbox* mbox, double timeout_s, int32_t* status,
                               double* sums, double* scalars, float* trace_row, dpde_stream_t stream);

/* Pass 2 -- analytic vector-Jacobian product replacing autograd through sample.py:336-353:
   g_x0 (B,C,H,W contiguous, dtype of x0) = d loss_comb / d x_N;  g_dxdt (same shape, may be NULL) = d / d dxdt.
   `upstream` (device double, may be NULL) multiplies both -- the grad_output of a stand-alone loss. */
int dpde_guidance_vjp(const dpde_guidance_desc* desc, const double* scalars, const double* upstream, void* g_x0,
                      void* g_dxdt, dpde_stream_t stream);

/* laplacian(u, dx), sample.py:106-134, on `planes` (H,W) images (adjoint != 0: its transpose, used as backward). */
int dpde_laplacian(const void* u, void* out, int32_t dtype, int64_t planes, int32_t H, int32_t W,
                   int64_t plane_stride_in, double dx, int32_t adjoint, dpde_stream_t stream);

/* ---- Training-time physics loss (next row f-2): EDMHeatLoss, models/loss.py:143 ------------------------------------
   out[b] = sum_{c,h,w} (dudt - alpha_b laplacian(u, dx))^2  for u, dudt (B, Cu, H, W) with rows contiguous and free
   batch / channel strides (x_0star[:, ch_a:] is a channel-slice view), alpha_b = labels[b, 1].  The caller applies
   1/(H W), the mean / sum over (C, H, W) and pde_loss_coeff / sigma^2 (models/loss.py:143-149) with torch ops.
   `workspace`: dpde_heat_residual_sq_workspace_bytes(B, Cu, H, W) bytes of device scratch.  dudt == NULL means zeros.
   Eligible layouts (as dpde_set_fast_path) run on the row-marching kernels, the rest on a generic grid-stride kernel. */
size_t dpde_heat_residual_sq_workspace_bytes(int32_t B, int32_t Cu, int32_t H, int32_t W);
int dpde_heat_residual_sq(const void* u, const void* dudt, int32_t dtype, int32_t B, int32_t Cu, int32_t H, int32_t W,
                          int64_t stride_b_u, int64_t stride_c_u, int64_t stride_b_dudt, int64_t stride_c_dudt,
                          const double* alpha, double dx, void* workspace, double* out, dpde_stream_t stream);

/* Its VJP: g_u (B,Cu,H,W contiguous, dtype of u) = upstream_b * 2 * (-alpha_b/dx^2) K^T r, g_dudt (may be NULL) =
   upstream_b * 2 * r, with r = dudt - alpha_b laplacian(u) and K^T the transposed reflect-padded stencil. */
int dpde_heat_residual_sq_vjp(const void* u, const void* dudt, int32_t dtype, int32_t B, int32_t Cu, int32_t H, int32_t W,
                              int64_t stride_b_u, int64_t stride_c_u, int64_t stride_b_dudt, int64_t stride_c_dudt,
                              const double* alpha, double dx, const double* upstream, void* g_u, void* g_dudt,
                              dpde_stream_t stream);

/* x = latents * sigma_0 (sample.py:316); also emits the fp32 copy the denoiser reads (sample.py:324). */
int dpde_sampler_init(const double* latents, double sigma0, double* x64, float* x32, int64_t n, dpde_stream_t stream);

/* Euler predictor, sample.py:327-328: x_eu = x_cur + (s_next - s_cur) (x_cur - x0_cur)/s_cur, emitted in fp32 for
   the second denoiser evaluation (sample.py:331). */
int dpde_euler_predict(const double* x_cur, const float* x0_cur, double sigma_cur, double sigma_next, float* x_eu32,
                       int64_t n, dpde_stream_t stream);

/* Backward of the predictor w.r.t. x0_cur: seed = fp32( -((s_next - s_cur) g_eu) / s_cur ). */
int dpde_euler_predict_bwd(const float* g_eu32, double sigma_cur, double sigma_next, float* seed32, int64_t n,
                           dpde_stream_t stream);

/* Heun correction + guidance update, sample.py:330-334,355, fused:
     x_next = x_cur + h (d_cur/2 + d_prime/2) - [ g_eu + (h g_eu)/s_cur + g_cur ],   h = s_next - s_cur
   x0_next == NULL selects the last step (Euler only, sample.py:330).  g_eu / g_cur are the fp32 gradients the
   denoiser backward produced at x_eu and x_cur (either may be NULL = 0).  Writes the fp64 state and its fp32 copy. */
int dpde_heun_guided_update(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                            const float* g_cur, double sigma_cur, double sigma_next, double* x_next64, float* x_next32,
                            int64_t n, dpde_stream_t stream);

/* Same update restricted to elements [first, first + count) of each of `planes` planes of plane_elems elements --
   the owned rows of a row slab, so ghost rows (which the neighbours push, dpde_halo_push) are never written. */
int dpde_heun_guided_update_rows(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                                 const float* g_cur, double sigma_cur, double sigma_next, double* x_next64,
                                 float* x_next32, int64_t planes, int64_t plane_elems, int64_t first, int64_t count,
                                 dpde_stream_t stream);

/* The same update fused with the halo exchange of a row slab (one kernel: compute + transfer over NVLink peer memory).
   The fields are planes x H_local x W local buffers with `halo` ghost rows per side; the owned rows [halo, H_local - halo)
   are updated.  The 2 x halo owned BOUNDARY rows are processed first and every element of them is stored twice: into this
   rank's next-state buffers and straight into the neighbours' ghost rows (up64 / up32: the upper neighbour's next-state
   buffers, H_up rows per plane, written at rows [H_up - halo, H_up); down64 / down32 at rows [0, halo)).  When the last CTA
   that holds boundary rows has fenced its peer stores (system scope), `epoch` is stored with release semantics into
   *flag_up / *flag_down (8-byte words in the neighbours' memory) -- while the remaining CTAs are still updating interior
   rows, so the transfer overlaps the bulk of the update.  NULL pointers on a side = grid boundary.  `ticket`: zeroed 4-byte
   word in local device memory (left zeroed).  Needs H_local >= 4 halo (boundary rows of the two sides disjoint). */
typedef struct dpde_halo_peers {
    double* up64;
    float* up32;
    double* down64;
    float* down32;
    void* flag_up;
    void* flag_down;
    void* ticket;
    uint64_t epoch;
    int32_t H_up, H_down;
} dpde_halo_peers;
int dpde_heun_guided_update_rows_push(const double* x_cur, const float* x0_cur, const float* x0_next, const float* g_eu,
                                      const float* g_cur, double sigma_cur, double sigma_next, double* x_next64,
                                      float* x_next32, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                                      const dpde_halo_peers* peers, dpde_stream_t stream);

/* ---- Peer memory: row-slab halo exchange over NVLink (one process per GPU; nothing like it in the reference) ----
   dpde_peer_alloc returns zero-filled device memory that can be exported to the other ranks of the box
   (cudaIpc); dpde_peer_export / dpde_peer_open move the 64-byte handle / map a neighbour's allocation into this
   process (peer access over NVLink is enabled by the mapping).  The caller keeps allocations alive until every
   rank has closed its mappings. */
#define DPDE_IPC_HANDLE_BYTES 64
int dpde_peer_alloc(size_t bytes, void** ptr);
int dpde_peer_free(void* ptr);
int dpde_peer_export(const void* ptr, unsigned char handle[DPDE_IPC_HANDLE_BYTES]);
int dpde_peer_open(const unsigned char handle[DPDE_IPC_HANDLE_BYTES], void** ptr);
int dpde_peer_close(void* ptr);

/* Halo push: copy this rank's owned boundary rows of `field` (planes x H_local x W, ghost rows included) straight
   into the neighbours' ghost rows through their mapped buffers --
     rows [halo, 2 halo)                   -> dst_up   rows [H_up - halo, H_up)      (dst_up: H_up rows per plane)
     rows [H_local - 2 halo, H_local - halo) -> dst_down rows [0, halo)              (dst_down: H_down rows per plane)
   then, after a system-scope fence, store `value` into *flag_up / *flag_down (8-byte words in the neighbours'
   memory) with release semantics.  NULL destination = no neighbour on that side.  `ticket` is a zeroed 4-byte
   word in LOCAL device memory (left zeroed). */
int dpde_halo_push(const void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                   void* dst_up, int32_t H_up, void* dst_down, int32_t H_down, void* flag_up, void* flag_down,
                   uint64_t value, void* ticket, dpde_stream_t stream);

/* Block the stream until every one of the n (<= 4) LOCAL 8-byte flags is >= value (acquire, system scope), i.e. the
   neighbours' pushes have landed.  After timeout_s seconds it gives up and writes 1 to *status (device int32,
   otherwise untouched) -- a stuck neighbour must not hang the GPU. */
int dpde_flag_wait(const void* const* flags, int32_t n, uint64_t value, double timeout_s, int32_t* status,
                   dpde_stream_t stream);

/* Row-slab halo staging for `planes` local images of H_local rows (ghost rows included), W columns.
   pack:   owned rows [halo, 2 halo) -> send_up, owned rows [H_local - 2 halo, H_local - halo) -> send_down
   unpack: recv_up -> ghost rows [0, halo),      recv_down -> ghost rows [H_local - halo, H_local)
   Each staging buffer holds planes * halo * W elements; a NULL buffer skips that side (grid boundary). */
int dpde_halo_pack(const void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                   void* send_up, void* send_down, dpde_stream_t stream);
int dpde_halo_unpack(void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo,
                     const void* recv_up, const void* recv_down, dpde_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DPDE_B200_H */
