// host_pipeline.cuh -- host buffers in, host buffers out: lyon_cuda_tessellate_paths_into (included by lyon_cuda.cu).
//
// The call a lyon user makes -- `tessellate_paths(&[Path], &StrokeOptions) -> VertexBuffers` -- for batches whose output
// does not have to fit (or wait) on the device in one piece.  The batch is cut into chunks of whole paths; two
// sub-contexts on two streams alternate, so the PCIe download of chunk i overlaps the upload and the kernels of chunk
// i + 1.  Chunk c is emitted with vertex_base = the vertices of the chunks before it, so the concatenated host
// buffers are exactly what ONE StrokeTessellator::tessellate_path loop over one VertexBuffers produces
// (stroke.rs:164-177, geometry_builder.rs:325-349).
struct HostPipeSlot {
    cudaStream_t stream = nullptr;
    lyon_cuda_ctx* sub = nullptr;
    DevBuf d_in, d_v, d_i;
    uint32_t* h_ranges = nullptr;  // pinned: the chunk's two rebased range arrays
    size_t h_ranges_cap = 0;
};
struct HostPipe {
    HostPipeSlot slot[2];
    double v_per_verb = 0.0, i_per_verb = 0.0;  // output size estimate from the chunks seen so far
};

static void host_pipe_destroy(HostPipe* hp) {
    if (!hp) return;
    for (HostPipeSlot& s : hp->slot) {
        if (s.sub) lyon_cuda_destroy(s.sub);
        if (s.d_in.p) cudaFree(s.d_in.p);
        if (s.d_v.p) cudaFree(s.d_v.p);
        if (s.d_i.p) cudaFree(s.d_i.p);
        if (s.h_ranges) cudaFreeHost(s.h_ranges);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete hp;
}

static int host_pipe_grow(lyon_cuda_ctx* c, HostPipeSlot& s, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return LYON_OK;
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    if (b.p) CU_TRY(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    bytes += bytes / 8 + 4096;
    CU_TRY(c, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return LYON_OK;
}

extern "C" int lyon_cuda_host_alloc(void** p, size_t bytes) {
    if (!p) return LYON_ERR_INVALID_ARGUMENT;
    *p = nullptr;
    return cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? LYON_OK : LYON_ERR_OUT_OF_MEMORY;
}
extern "C" void lyon_cuda_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" int lyon_cuda_tessellate_paths_into(lyon_cuda_ctx* c, const lyon_path_batch* hb, const lyon_stroke_options* o,
                                    lyon_stroke_vertex* h_vertices, uint64_t vertex_capacity, uint32_t* h_indices,
                                    uint64_t index_capacity, uint32_t chunk_paths, uint64_t* n_vertices, uint64_t* n_indices) {
    if (!c || !hb || !o || !n_vertices || !n_indices) return LYON_ERR_INVALID_ARGUMENT;
    *n_vertices = *n_indices = 0;
    if ((vertex_capacity && !h_vertices) || (index_capacity && !h_indices)) return LYON_ERR_INVALID_ARGUMENT;
    if (hb->n_paths == 0 || hb->n_verbs == 0) return LYON_OK;
    if (!hb->verbs || !hb->points || !hb->path_verb_begin || !hb->path_point_begin) return LYON_ERR_INVALID_ARGUMENT;
    if (chunk_paths == 0) chunk_paths = 32768u;
    DeviceGuard guard(c->device);
    if (!c->host_pipe) {
        HostPipe* hp = new HostPipe();
        for (HostPipeSlot& s : hp->slot) {
            if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
                lyon_cuda_create(c->device, (void*)s.stream, &s.sub) != LYON_OK) {
                host_pipe_destroy(hp);
                c->err = "host pipeline: cannot create the two sub-contexts";
                return LYON_ERR_CUDA;
            }
        }
        c->host_pipe = hp;
    }
    HostPipe* hp = c->host_pipe;
    uint64_t tot_v = 0, tot_i = 0;
    bool too_small = false;
    int st = LYON_OK;
    uint32_t chunk = 0;
    for (uint32_t lo = 0; lo < hb->n_paths && st == LYON_OK; lo += chunk_paths, ++chunk) {
        const uint32_t hi = (uint32_t)std::min<uint64_t>((uint64_t)lo + chunk_paths, hb->n_paths);
        HostPipeSlot& s = hp->slot[chunk & 1u];
        const uint32_t np = hi - lo;
        const uint32_t v0 = hb->path_verb_begin[lo], v1 = hb->path_verb_begin[hi];
        const uint32_t p0 = hb->path_point_begin[lo], p1 = hb->path_point_begin[hi];
        if (v1 < v0 || p1 < p0) return LYON_ERR_INVALID_ARGUMENT;
        const size_t b_verbs = (size_t)(v1 - v0), a_verbs = (b_verbs + 255) & ~(size_t)255;
        const size_t b_points = (size_t)(p1 - p0) * 8, a_points = (b_points + 255) & ~(size_t)255;
        const size_t b_ranges = ((size_t)np + 1) * 4, a_ranges = (b_ranges + 255) & ~(size_t)255;
        // the slot's previous chunk (two chunks ago) must have left its buffers: stroke_fused of that chunk synchronised
        // the stream for the sizes, the download that followed has to be waited for before the input area is reused
        if ((st = host_pipe_grow(c, s, s.d_in, a_verbs + a_points + 2 * a_ranges)) != LYON_OK) break;
        if (2 * ((size_t)np + 1) > s.h_ranges_cap) {
            if (cudaStreamSynchronize(s.stream) != cudaSuccess) { st = LYON_ERR_CUDA; break; }
            if (s.h_ranges) cudaFreeHost(s.h_ranges);
            s.h_ranges = nullptr;
            s.h_ranges_cap = 2 * ((size_t)std::max(np, chunk_paths) + 1);
            if (cudaHostAlloc((void**)&s.h_ranges, s.h_ranges_cap * 4, cudaHostAllocDefault) != cudaSuccess) {
                s.h_ranges_cap = 0;
                st = LYON_ERR_OUT_OF_MEMORY;
                break;
            }
        }
        if (cudaStreamSynchronize(s.stream) != cudaSuccess) { st = LYON_ERR_CUDA; break; }
        uint32_t* rv = s.h_ranges;
        uint32_t* rp = s.h_ranges + (np + 1);
        for (uint32_t i = 0; i <= np; ++i) {
            rv[i] = hb->path_verb_begin[lo + i] - v0;
            rp[i] = hb->path_point_begin[lo + i] - p0;
        }
        char* d_in = (char*)s.d_in.p;
        lyon_path_batch db = *hb;
        db.n_paths = np;
        db.n_verbs = v1 - v0;
        db.n_points = p1 - p0;
        db.verbs = (const uint8_t*)d_in;
        db.points = (const float*)(d_in + a_verbs);
        db.path_verb_begin = (const uint32_t*)(d_in + a_verbs + a_points);
        db.path_point_begin = (const uint32_t*)(d_in + a_verbs + a_points + a_ranges);
        if (cudaMemcpyAsync((void*)db.verbs, hb->verbs + v0, b_verbs, cudaMemcpyHostToDevice, s.stream) != cudaSuccess ||
            cudaMemcpyAsync((void*)db.points, hb->points + 2 * (size_t)p0, b_points, cudaMemcpyHostToDevice, s.stream) != cudaSuccess ||
            cudaMemcpyAsync((void*)db.path_verb_begin, rv, b_ranges, cudaMemcpyHostToDevice, s.stream) != cudaSuccess ||
            cudaMemcpyAsync((void*)db.path_point_begin, rp, b_ranges, cudaMemcpyHostToDevice, s.stream) != cudaSuccess) {
            st = LYON_ERR_CUDA;
            break;
        }
        // output buffers of the slot: sized from the chunks seen so far, grown (and the chunk redone) when too small
        uint64_t nv = 0, ni = 0;
        if (hp->v_per_verb > 0.0) {
            if ((st = host_pipe_grow(c, s, s.d_v, (size_t)(hp->v_per_verb * 1.1 * (double)db.n_verbs + 1024) * 32)) != LYON_OK) break;
            if ((st = host_pipe_grow(c, s, s.d_i, (size_t)(hp->i_per_verb * 1.1 * (double)db.n_verbs + 4096) * 4)) != LYON_OK) break;
        }
        for (int attempt = 0; attempt < 2; ++attempt) {
            st = lyon_cuda_stroke_fused(s.sub, &db, o, (lyon_stroke_vertex*)s.d_v.p, s.d_v.cap / 32, (uint32_t*)s.d_i.p,
                                        s.d_i.cap / 4, tot_v, nullptr, nullptr, &nv, &ni);
            if (st != LYON_ERR_BUFFER_TOO_SMALL || attempt == 1) break;
            if ((st = host_pipe_grow(c, s, s.d_v, (size_t)nv * 32)) != LYON_OK) break;
            if ((st = host_pipe_grow(c, s, s.d_i, (size_t)(ni ? ni : 1) * 4)) != LYON_OK) break;
        }
        if (st != LYON_OK) {
            c->err = s.sub->err;
            break;
        }
        if (tot_v + nv > 0xFFFFFFFFull) {
            st = LYON_ERR_TOO_MANY_VERTICES;
            break;
        }
        hp->v_per_verb = std::max(hp->v_per_verb, (double)nv / (double)std::max<uint32_t>(db.n_verbs, 1u));
        hp->i_per_verb = std::max(hp->i_per_verb, (double)ni / (double)std::max<uint32_t>(db.n_verbs, 1u));
        if (tot_v + nv > vertex_capacity || tot_i + ni > index_capacity) {
            too_small = true;  // keep counting: the caller learns the exact sizes
        } else if (!too_small) {
            if ((nv && cudaMemcpyAsync(h_vertices + tot_v, s.d_v.p, (size_t)nv * 32, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess) ||
                (ni && cudaMemcpyAsync(h_indices + tot_i, s.d_i.p, (size_t)ni * 4, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)) {
                st = LYON_ERR_CUDA;
                break;
            }
        }
        tot_v += nv;
        tot_i += ni;
    }
    for (HostPipeSlot& s : hp->slot)
        if (cudaStreamSynchronize(s.stream) != cudaSuccess && st == LYON_OK) st = LYON_ERR_CUDA;
    if (st == LYON_OK)
        for (HostPipeSlot& s : hp->slot) {
            const int e = lyon_cuda_check(s.sub);
            if (e != LYON_OK) {
                st = e;
                c->err = s.sub->err;
                break;
            }
        }
    if (st != LYON_OK) return st;
    *n_vertices = tot_v;
    *n_indices = tot_i;
    return too_small ? LYON_ERR_BUFFER_TOO_SMALL : LYON_OK;
}
