// sharded.cuh -- multi-GPU entry points of the C ABI (included by lyon_cuda.cu): one process per GPU, every rank strokes
// its own contiguous range of paths, the ONLY exchange is an all-gather of the per-rank output totals {V_r, I_r} over
// NCCL (NVLink / NVSwitch), and rank r rebases its indices by sum(V_q, q < r) -- BuffersBuilder::with_vertex_offset
// (geometry_builder.rs:325-329) applied to the rank's slice, so that the concatenation of the per-rank buffers is the
// single-GPU result bit for bit.  Outputs stay resident per GPU.
//
// The all-gather is enqueued on the context's stream right behind the stroke kernels and the rebase kernel reads its base
// from the gathered device buffer: no host round trip between the three.  NCCL is resolved at run time (dlopen of the
// libnccl the application already uses), so the library itself has no link-time dependency on it.
#include <dlfcn.h>

namespace {

struct LycNcclId {
    char internal[128];
};
typedef void* LycNcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(LycNcclId*) = nullptr;
    int (*CommInitRank)(LycNcclComm*, int, LycNcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, LycNcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(LycNcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok() const { return GetUniqueId && CommInitRank && AllGather && CommDestroy; }
};
constexpr int LYC_NCCL_UINT64 = 5;  // ncclUint64 (nccl.h: ncclDataType_t)

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok() ? &api : nullptr;
    tried = true;
    const char* names[] = {getenv("LYON_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
    api.GetUniqueId = (int (*)(LycNcclId*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(LycNcclComm*, int, LycNcclId, int))dlsym(api.handle, "ncclCommInitRank");
    api.AllGather = (int (*)(const void*, void*, size_t, int, LycNcclComm, cudaStream_t))dlsym(api.handle, "ncclAllGather");
    api.CommDestroy = (int (*)(LycNcclComm))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    return api.ok() ? &api : nullptr;
}

// idx[i] += sum of the vertex counts of the ranks before `rank`, read from the gathered totals ({V, I, status} per rank)
__global__ void __launch_bounds__(256) k_rebase_indices_gathered(uint32_t* __restrict__ idx, unsigned long long n,
                                                                 const unsigned long long* __restrict__ gathered, int rank) {
    unsigned long long base64 = 0ull;
    for (int q = 0; q < rank; ++q) base64 += gathered[3 * q];
    const uint32_t base = (uint32_t)base64;
    if (base == 0u) return;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long head = min(n, (unsigned long long)((16u - ((uintptr_t)idx & 15u)) & 15u) / 4ull);
    if (i < head) idx[i] += base;
    uint4* v = reinterpret_cast<uint4*>(idx + head);
    const unsigned long long nv = (n - head) / 4ull;
    for (unsigned long long k = i; k < nv; k += stride) {
        uint4 q = v[k];
        q.x += base;
        q.y += base;
        q.z += base;
        q.w += base;
        v[k] = q;
    }
    const unsigned long long tail0 = head + nv * 4ull;
    if (tail0 + i < n && i < 4ull) idx[tail0 + i] += base;
}

}  // namespace

struct CommState {
    LycNcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    unsigned long long* d_send = nullptr;  // {V_r, I_r, status}
    unsigned long long* d_recv = nullptr;  // nranks x {V, I, status}
    unsigned long long* h_pin = nullptr;   // pinned: 3 (send) + 3 * nranks (recv)
};

static void comm_state_destroy(CommState* s) {
    if (!s) return;
    if (s->comm) {
        if (NcclApi* api = nccl_api()) api->CommDestroy(s->comm);
    }
    if (s->d_send) cudaFree(s->d_send);
    if (s->d_recv) cudaFree(s->d_recv);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    delete s;
}

extern "C" int lyon_cuda_comm_unique_id(uint8_t* id128) {
    if (!id128) return LYON_ERR_INVALID_ARGUMENT;
    NcclApi* api = nccl_api();
    if (!api) return LYON_ERR_CUDA;
    LycNcclId id;
    if (api->GetUniqueId(&id) != 0) return LYON_ERR_CUDA;
    memcpy(id128, id.internal, 128);
    return LYON_OK;
}

extern "C" int lyon_cuda_comm_init(lyon_cuda_ctx* c, const uint8_t* id128, int rank, int nranks) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return LYON_ERR_INVALID_ARGUMENT;
    DeviceGuard device_guard(c->device);
    NcclApi* api = nccl_api();
    if (!api) {
        c->err = "lyon_cuda_comm_init: libnccl.so.2 could not be loaded (set LYON_NCCL_LIB)";
        return LYON_ERR_CUDA;
    }
    comm_state_destroy(c->comm);
    c->comm = nullptr;
    CommState* s = new CommState();
    s->rank = rank;
    s->nranks = nranks;
    LycNcclId id;
    memcpy(id.internal, id128, 128);
    const int r = api->CommInitRank(&s->comm, nranks, id, rank);
    if (r != 0) {
        c->err = std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "failed");
        s->comm = nullptr;
        comm_state_destroy(s);
        return LYON_ERR_CUDA;
    }
    if (cudaMalloc((void**)&s->d_send, 3 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void**)&s->d_recv, 3 * sizeof(unsigned long long) * (size_t)nranks) != cudaSuccess ||
        cudaHostAlloc((void**)&s->h_pin, 3 * sizeof(unsigned long long) * ((size_t)nranks + 1), cudaHostAllocDefault) != cudaSuccess) {
        comm_state_destroy(s);
        return LYON_ERR_OUT_OF_MEMORY;
    }
    c->comm = s;
    return LYON_OK;
}

extern "C" int lyon_cuda_comm_destroy(lyon_cuda_ctx* c) {
    if (!c) return LYON_ERR_INVALID_ARGUMENT;
    DeviceGuard device_guard(c->device);
    cudaStreamSynchronize(c->stream);
    comm_state_destroy(c->comm);
    c->comm = nullptr;
    return LYON_OK;
}

extern "C" int lyon_cuda_stroke_fused_sharded(lyon_cuda_ctx* c, const lyon_path_batch* b, const lyon_stroke_options* o,
                                              lyon_stroke_vertex* d_vertices, uint64_t vertex_capacity, uint32_t* d_indices,
                                              uint64_t index_capacity, uint64_t* d_path_vertex_end, uint64_t* d_path_index_end,
                                              uint64_t* n_vertices, uint64_t* n_indices, uint64_t* vertex_base,
                                              uint64_t* h_totals) {
    if (!c || !n_vertices || !n_indices) return LYON_ERR_INVALID_ARGUMENT;
    DeviceGuard device_guard(c->device);
    CommState* s = c->comm;
    if (!s) {
        c->err = "lyon_cuda_stroke_fused_sharded without lyon_cuda_comm_init";
        return LYON_ERR_INVALID_ARGUMENT;
    }
    NcclApi* api = nccl_api();
    if (!api) return LYON_ERR_CUDA;
    // 1. this rank's slice, ids local to the slice (vertex_base 0).  Whatever happens here, the rank still takes part in
    //    the all-gather below with its status word: a rank that fails must not leave the others waiting in a collective.
    uint64_t nv = 0, ni = 0;
    const int local = lyon_cuda_stroke_fused(c, b, o, d_vertices, vertex_capacity, d_indices, index_capacity, 0, d_path_vertex_end,
                                             d_path_index_end, &nv, &ni);
    *n_vertices = nv;
    *n_indices = ni;
    const std::string local_err = c->err;
    s->h_pin[0] = nv;
    s->h_pin[1] = ni;
    s->h_pin[2] = (unsigned long long)(uint32_t)(-local);  // 0 = LYON_OK
    CU_TRY(c, cudaMemcpyAsync(s->d_send, s->h_pin, 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    // 2. the one exchange of the path: {V_r, I_r, status} from every rank
    const int r = api->AllGather(s->d_send, s->d_recv, 3, LYC_NCCL_UINT64, s->comm, c->stream);
    if (r != 0) {
        c->err = std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(r) : "failed");
        return LYON_ERR_CUDA;
    }
    // 3. rebase this rank's indices; the base is read from the gathered totals on the device
    if (local == LYON_OK && ni) {
        unsigned blocks = grid_for((ni + 3) / 4, 256);
        const unsigned cap = (unsigned)c->num_sms * 32u;
        if (blocks > cap) blocks = cap;
        k_rebase_indices_gathered<<<blocks, 256, 0, c->stream>>>(d_indices, ni, s->d_recv, s->rank);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
    }
    // 4. the totals of every rank for the caller (and the job-wide checks)
    CU_TRY(c, cudaMemcpyAsync(s->h_pin + 3, s->d_recv, 3 * sizeof(unsigned long long) * (size_t)s->nranks, cudaMemcpyDeviceToHost,
                              c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    uint64_t base = 0, all_v = 0;
    int peer_failed = -1;
    for (int q = 0; q < s->nranks; ++q) {
        const unsigned long long* t = s->h_pin + 3 + 3 * q;
        if (q < s->rank) base += t[0];
        all_v += t[0];
        if (t[2] != 0ull && peer_failed < 0) peer_failed = q;
        if (h_totals) {
            h_totals[2 * q] = t[0];
            h_totals[2 * q + 1] = t[1];
        }
    }
    if (vertex_base) *vertex_base = base;
    if (local != LYON_OK) {
        c->err = local_err;
        return local;
    }
    if (peer_failed >= 0 && -(int)(uint32_t)s->h_pin[3 + 3 * peer_failed + 2] == LYON_ERR_BUFFER_TOO_SMALL) {
        // a peer has to grow its buffers and call again: the call is collective, so every rank does
        c->err = "lyon_cuda_stroke_fused_sharded: rank " + std::to_string(peer_failed) + " needs larger output buffers";
        return LYON_ERR_BUFFER_TOO_SMALL;
    }
    if (peer_failed >= 0) {
        c->err = "lyon_cuda_stroke_fused_sharded: rank " + std::to_string(peer_failed) + " failed (status " +
                 std::to_string(-(int)(uint32_t)s->h_pin[3 + 3 * peer_failed + 2]) + ")";
        return LYON_ERR_INTERNAL;
    }
    // TooManyVertices for the job as a whole: the last global vertex id must stay <= u32::MAX - 1 (geometry_builder.rs:578-581)
    if (all_v > 0xFFFFFFFFull) return LYON_ERR_TOO_MANY_VERTICES;
    return LYON_OK;
}
