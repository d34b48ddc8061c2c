// Row-slab halo exchange over NVLink peer memory (sm_100a).  New functionality: the reference is single-process.
//
// One process per GPU.  Each rank allocates its state buffers with dpde_peer_alloc, exports them with CUDA IPC and
// maps its two row neighbours' buffers.  After the guided update of a step has written the rank's owned rows,
// dpde_halo_push copies the owned boundary rows STRAIGHT INTO the neighbours' ghost rows with ordinary stores on the
// mapped peer pointers (NVLink 5 through the NVSwitch), fences at system scope and then publishes the step number
// in a flag word that lives in the neighbour's memory (st.release.sys).  The neighbour's next step starts with
// dpde_flag_wait (ld.acquire.sys spin in a one-thread kernel, bounded by a timeout), after which its ghost rows are
// current.  No staging buffer, no pack/unpack pass, no host involvement; 2 rows x W x planes per side per step
// (config 5: 2 x 4096 x 16 planes x 12 B = 1.5 MB per neighbour), so the exchange is latency- not bandwidth-bound
// and overlaps with nothing it needs to wait for: the only consumer is the next step's first kernel.

#include "common.cuh"

namespace dpde {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// V: copy unit (int4 = 16 bytes when rows allow it, else the element type).  row_units = units per row.
template <typename V>
__global__ void __launch_bounds__(kThreads)
halo_push_kernel(const V* __restrict__ field, int64_t planes, int H, int row_units, int halo, V* __restrict__ dst_up,
                 int H_up, V* __restrict__ dst_down, int H_down, unsigned long long* flag_up,
                 unsigned long long* flag_down, unsigned long long value, unsigned int* ticket) {
    const int64_t per_plane = (int64_t)halo * row_units, total = planes * per_plane;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pl = i / per_plane, r = i - pl * per_plane;
        const V* src = field + pl * (int64_t)H * row_units;
        if (dst_up) dst_up[pl * (int64_t)H_up * row_units + (int64_t)(H_up - halo) * row_units + r] = src[(int64_t)halo * row_units + r];
        if (dst_down) dst_down[pl * (int64_t)H_down * row_units + r] = src[(int64_t)(H - 2 * halo) * row_units + r];
    }
    // every thread's peer stores are made visible system-wide before its CTA takes a ticket; the CTA that takes
    // the last ticket publishes the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            if (flag_up) st_release_sys(flag_up, value);
            if (flag_down) st_release_sys(flag_down, value);
        }
    }
}

struct FlagList {
    const unsigned long long* f[4];
    int n;
};

__global__ void flag_wait_kernel(const __grid_constant__ FlagList flags, unsigned long long value,
                                 unsigned long long timeout_ns, int* status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long t0 = global_ns();
    for (int k = 0; k < flags.n; ++k) {
        while (ld_acquire_sys(flags.f[k]) < value) {
            if (global_ns() - t0 > timeout_ns) {
                if (status) *status = 1;
                return;
            }
            __nanosleep(200);
        }
    }
}

}  // namespace
}  // namespace dpde

using namespace dpde;

extern "C" {

int dpde_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return fail(DPDE_ERR_INVALID, "dpde_peer_alloc: ptr is NULL or bytes == 0");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "dpde_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(DPDE_ERR_CUDA, "dpde_peer_alloc: cudaMemset: %s", cudaGetErrorString(e));
    }
    *ptr = p;
    return DPDE_OK;
}

int dpde_peer_free(void* ptr) {
    if (!ptr) return DPDE_OK;
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "dpde_peer_free: %s", cudaGetErrorString(e));
    return DPDE_OK;
}

int dpde_peer_export(const void* ptr, unsigned char handle[DPDE_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == DPDE_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ptr || !handle) return fail(DPDE_ERR_INVALID, "dpde_peer_export: null pointer");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "dpde_peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    memcpy(handle, &h, sizeof(h));
    return DPDE_OK;
}

int dpde_peer_open(const unsigned char handle[DPDE_IPC_HANDLE_BYTES], void** ptr) {
    if (!ptr || !handle) return fail(DPDE_ERR_INVALID, "dpde_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "dpde_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *ptr = p;
    return DPDE_OK;
}

int dpde_peer_close(void* ptr) {
    if (!ptr) return DPDE_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "dpde_peer_close: %s", cudaGetErrorString(e));
    return DPDE_OK;
}

int dpde_halo_push(const void* field, int32_t dtype, int64_t planes, int32_t H_local, int32_t W, int32_t halo, void* dst_up,
                   int32_t H_up, void* dst_down, int32_t H_down, void* flag_up, void* flag_down, uint64_t value,
                   void* ticket, dpde_stream_t stream) {
    const char* who = "dpde_halo_push";
    if (!field || !ticket) return fail(DPDE_ERR_INVALID, "%s: field/ticket is NULL", who);
    if (planes < 0 || W < 1 || halo < 1 || H_local < 3 * halo) return fail(DPDE_ERR_INVALID, "%s: need halo >= 1 and H_local >= 3 halo", who);
    if (dtype != DPDE_F32 && dtype != DPDE_F64) return fail(DPDE_ERR_UNSUPPORTED, "%s: dtype must be f32/f64", who);
    if ((dst_up && H_up < 3 * halo) || (dst_down && H_down < 3 * halo)) return fail(DPDE_ERR_INVALID, "%s: neighbour slab shorter than 3 halo", who);
    if ((dst_up && !flag_up) || (dst_down && !flag_down)) return fail(DPDE_ERR_INVALID, "%s: destination without flag", who);
    if (planes == 0 || (!dst_up && !dst_down)) return DPDE_OK;
    const size_t es = dtype == DPDE_F32 ? 4 : 8, row_bytes = (size_t)W * es;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool v16 = row_bytes % 16 == 0 && al16(field) && (!dst_up || al16(dst_up)) && (!dst_down || al16(dst_down));
    const int row_units = (int)(v16 ? row_bytes / 16 : W);
    const int64_t total = planes * (int64_t)halo * row_units;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > (int64_t)sm_count() * 2) blocks = (int64_t)sm_count() * 2;
    cudaStream_t s = (cudaStream_t)stream;
    auto fu = (unsigned long long*)flag_up;
    auto fd = (unsigned long long*)flag_down;
    if (v16)
        halo_push_kernel<int4><<<(int)blocks, kThreads, 0, s>>>((const int4*)field, planes, H_local, row_units, halo, (int4*)dst_up, H_up,
                                                               (int4*)dst_down, H_down, fu, fd, value, (unsigned int*)ticket);
    else if (dtype == DPDE_F32)
        halo_push_kernel<float><<<(int)blocks, kThreads, 0, s>>>((const float*)field, planes, H_local, row_units, halo, (float*)dst_up, H_up,
                                                                (float*)dst_down, H_down, fu, fd, value, (unsigned int*)ticket);
    else
        halo_push_kernel<double><<<(int)blocks, kThreads, 0, s>>>((const double*)field, planes, H_local, row_units, halo, (double*)dst_up,
                                                                 H_up, (double*)dst_down, H_down, fu, fd, value, (unsigned int*)ticket);
    return check_launch(who);
}

int dpde_flag_wait(const void* const* flags, int32_t n, uint64_t value, double timeout_s, int32_t* status, dpde_stream_t stream) {
    if (n < 0 || n > 4 || (n > 0 && !flags)) return fail(DPDE_ERR_INVALID, "dpde_flag_wait: need 0 <= n <= 4 flags");
    if (!(timeout_s > 0.0)) return fail(DPDE_ERR_INVALID, "dpde_flag_wait: timeout_s must be > 0");
    if (n == 0) return DPDE_OK;
    FlagList fl{};
    fl.n = n;
    for (int k = 0; k < n; ++k) {
        if (!flags[k]) return fail(DPDE_ERR_INVALID, "dpde_flag_wait: flag %d is NULL", k);
        fl.f[k] = (const unsigned long long*)flags[k];
    }
    flag_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fl, value, (unsigned long long)(timeout_s * 1e9), status);
    return check_launch("dpde_flag_wait");
}

}  // extern "C"
