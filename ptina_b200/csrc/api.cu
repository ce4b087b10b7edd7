// api.cu -- the C ABI of include/ptina_b200.h: context, loaders, film, render entry points, parity taps.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "ptb_internal.h"

static thread_local char g_err[512] = "";
void ptb_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

TraceScene ptb_trace_scene(const ptb_ctx* c) {
    TraceScene S;
    S.nodes = c->d_nodes_active ? c->d_nodes_active : c->d_nodes; S.tris = c->d_tris; S.leaf = c->d_leaf; S.slot_of = c->d_slot_of; S.gate = c->d_gate; S.gbox = c->d_gbox; S.tlo = c->d_tlo; S.thi = c->d_thi; S.nlo = c->d_nlo; S.nhi = c->d_nhi; S.list = c->d_list; S.nlist = c->list_n; S.scene_abs = c->scene_abs; S.root_must = c->root_must; S.bmin = c->d_bmin; S.bmax = c->d_bmax; S.child = c->d_child; S.n = c->tree_n;
    S.qnodes = c->d_qnodes;
    S.wnodes = (c->wide4 && c->wnodes_ok) ? c->d_wnodes : nullptr;
    for (int k = 0; k < 3; k++) { S.qbase[k] = c->qbase[k]; S.qext[k] = c->qext[k]; S.qinv[k] = c->qinv[k]; }
    return S;
}
int ptb_effective_policy(const ptb_ctx* c, int requested) {
    if (requested == PTB_TRAVERSE_REFERENCE) return PTB_TRAVERSE_REFERENCE;
    bool ok = c->tree_info.valid && c->tree_info.depth <= PTB_STACK;
    if (!ok) return PTB_TRAVERSE_REFERENCE;
    // AUTO and both ORDERED policies need a proper tree
    return requested == PTB_TRAVERSE_ORDERED_EXACT ? PTB_TRAVERSE_ORDERED_EXACT : PTB_TRAVERSE_ORDERED;
}

namespace {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
template <class T> int dmalloc(T** p, size_t count) {
    PTB_CUDA(cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)));
    return 0;
}
// stage a host array on the device (or pass a device pointer through)
template <class T> struct In {
    const T* dev = nullptr; T* owned = nullptr;
    int set(ptb_ctx* c, const T* src, size_t count, int memspace = PTB_HOST) {
        if (!src) return 0;
        if (memspace == PTB_DEVICE) { dev = src; return 0; }
        PTB_CUDA(cudaMalloc((void**)&owned, sizeof(T) * (count ? count : 1)));
        PTB_CUDA(cudaMemcpyAsync(owned, src, sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
        dev = owned;
        return 0;
    }
    ~In() { if (owned) cudaFree(owned); }
};
template <class T> struct Out {
    T* dev = nullptr; T* host = nullptr; size_t count = 0; bool owned = false;
    int set(T* dst, size_t n, int memspace = PTB_HOST) {
        if (!dst) return 0;
        count = n;
        if (memspace == PTB_DEVICE) { dev = dst; return 0; }
        host = dst; owned = true;
        PTB_CUDA(cudaMalloc((void**)&dev, sizeof(T) * (n ? n : 1)));
        return 0;
    }
    int finish(ptb_ctx* c) {
        if (host) {
            PTB_CUDA(cudaMemcpyAsync(host, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, c->stream));
            PTB_CUDA(cudaStreamSynchronize(c->stream));
        }
        return 0;
    }
    ~Out() { if (owned && dev) cudaFree(dev); }
};
}  // namespace

#define CHECK_CTX(c) do { if (!(c)) { ptb_set_error("null context"); return 1; } } while (0)

// ---- deferred render calls ------------------------------------------------------------------------------------------------
// The reference's drivers call Engine.render() once per sample (exams/benchmark.py:29-33), each call one megakernel.  A one-sample
// wavefront of a 512x512 film is 262 144 paths -- mostly launch latency and kernel tails -- so ptb_render only RECORDS the call
// (the Sobol time advances at once, as path.py:75-77 does) and the recorded samples are submitted as one batch when the batch is
// full or anything could observe the difference: a readback, a film / scene / camera / light / material change, a counter query,
// synchronize.  Same samples, same per-pixel accumulation order: the film is bit-identical to submitting call by call
// (tests/test_gpu_parity.py::test_batching_and_ranges_are_equivalent, test_percall_render_is_coalesced).
static int flush_pending(ptb_ctx* c) {
    if (c->pend_count <= 0) return 0;
    const int engine = c->pend_engine, first = c->pend_first, count = c->pend_count;
    c->pend_count = 0; c->pend_engine = -1;
    DeviceGuard g(c->device);
    return ptb_wf_render(c, engine, first, count, 1, nullptr);
}
#define CHECK_FLUSH(c) do { CHECK_CTX(c); if (flush_pending(c)) return 1; } while (0)

// frees everything a context owns (also the partly built context of a failed ptb_create)
static void release_ctx(ptb_ctx* c) {
    void* ptrs[] = {c->d_verts, c->d_mtlids, c->d_texels, c->d_params, c->d_cache, c->d_sobolV, c->d_sobolP, c->d_mc, c->d_id, c->d_mc_tmp, c->d_id_tmp, c->d_leaf,
                    c->d_child, c->d_bmin, c->d_bmax, c->d_ready, c->d_parentcnt, c->d_range, c->d_height, c->d_nodes, c->d_nodes2, c->d_pl_id[0], c->d_pl_id[1], c->d_pl_depth[0], c->d_pl_depth[1], c->d_pl_nn, c->d_pl_lo[0], c->d_pl_lo[1], c->d_pl_hi[0], c->d_pl_hi[1], c->d_pl_keep, c->d_pl_make, c->d_pl_kpos, c->d_pl_mpos, c->d_pl_S, c->d_qnodes, c->d_wnodes, c->d_tris, c->d_slot_of, c->d_gate, c->d_gbox, c->d_tlo, c->d_thi, c->d_nlo, c->d_nhi, c->d_list, c->d_sort_tmp, c->d_scalars,
                    c->d_film, c->st.ray_o, c->st.ray_d, c->st.hit, c->st.thr, c->st.result, c->xq[0].o, c->xq[0].d, c->xq[1].o, c->xq[1].d, c->sq.o, c->sq.d, c->sq.c, c->tq.e[0], c->tq.e[1], c->tq.e[2], c->tq.e[3], c->tq.e[4], c->tq2.e[0], c->tq2.e[1], c->tq2.e[2], c->tq2.e[3], c->tq2.e[4],
                    c->d_ctrl, c->d_counters, c->d_Xold, c->d_Xnew, c->d_Lold, c->d_resolve, c->d_flags};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->ev_shade) cudaEventDestroy(c->ev_shade);
    if (c->ev_shadow) cudaEventDestroy(c->ev_shadow);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int w = 0; w < PTB_MAX_LANES; w++) {
        for (cudaEvent_t e : {c->lane_shade[w], c->lane_shadow[w], c->lane_join[w], c->ev_acc[w]}) if (e) cudaEventDestroy(e);
        for (cudaStream_t st : {c->lane_main[w], c->lane_side[w]}) if (st) cudaStreamDestroy(st);
    }
    delete c;
}

extern "C" {

const char* ptb_last_error(void) { return g_err; }
int ptb_version(void) { return 200; }

int ptb_create(int device, const ptb_caps* caps, ptb_ctx** out) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        ptb_set_error("no CUDA device: %s (ptina_b200 has no CPU fallback)", cudaGetErrorString(e));
        return 1;
    }
    if (device < 0 || device >= ndev) { ptb_set_error("device %d out of range (%d visible)", device, ndev); return 1; }
    PTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PTB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { ptb_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return 1; }
    ptb_ctx* c = new ptb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    ptb_caps d{};
    if (caps) d = *caps;
    if (d.max_faces <= 0) d.max_faces = 1 << 21;          // things.py:13
    if (d.max_faces > (1 << 24)) { ptb_set_error("max_faces %d above 2^24 (leaf slots are 24-bit fields of the ray queues)", d.max_faces); delete c; return 1; }
    if (d.max_texels <= 0) d.max_texels = 1 << 22;        // things.py:14
    if (d.max_materials <= 0) d.max_materials = 64;
    if (d.max_textures <= 0) d.max_textures = 64;
    if (d.max_lights <= 0) d.max_lights = 64;
    if (d.max_filmsize <= 0) d.max_filmsize = 1 << 21;    // things.py:18
    if (d.max_filmpasses <= 0) d.max_filmpasses = 3;
    if (d.max_paths <= 0) d.max_paths = 1 << 23;
    if (d.max_materials > PTB_MAX_MATERIALS || d.max_textures > PTB_MAX_TEXTURES || d.max_lights > PTB_MAX_LIGHTS || d.max_filmpasses < 3) {
        ptb_set_error("capacities beyond the compiled table sizes (64 materials / textures / lights, >= 3 film passes)");
        release_ctx(c); return 1;
    }
    c->caps = d;
    c->max_paths = d.max_paths;
    size_t nf = (size_t)d.max_faces;
    if (dmalloc(&c->d_verts, nf * 24) || dmalloc(&c->d_mtlids, nf) || dmalloc(&c->d_texels, (size_t)d.max_texels) ||
        dmalloc(&c->d_mc, nf) || dmalloc(&c->d_id, nf) || dmalloc(&c->d_mc_tmp, nf) || dmalloc(&c->d_id_tmp, nf) || dmalloc(&c->d_leaf, nf) ||
        dmalloc(&c->d_child, nf) || dmalloc(&c->d_bmin, nf * 3) || dmalloc(&c->d_bmax, nf * 3) || dmalloc(&c->d_ready, nf) ||
        dmalloc(&c->d_parentcnt, nf * 2) || dmalloc(&c->d_range, nf) || dmalloc(&c->d_height, nf) || dmalloc(&c->d_nodes, nf) || dmalloc(&c->d_qnodes, 2 * nf) ||
        dmalloc(&c->d_tris, nf) || dmalloc(&c->d_slot_of, nf) || dmalloc(&c->d_gate, nf) || dmalloc(&c->d_gbox, 2 * nf) || dmalloc(&c->d_tlo, nf) || dmalloc(&c->d_thi, nf) || dmalloc(&c->d_nlo, nf) || dmalloc(&c->d_nhi, nf) || dmalloc(&c->d_list, PTB_LIST_CAP) || dmalloc(&c->d_nodes2, 8200) || dmalloc(&c->d_pl_id[0], 8200) || dmalloc(&c->d_pl_id[1], 8200) || dmalloc(&c->d_pl_depth[0], 8200) || dmalloc(&c->d_pl_depth[1], 8200) || dmalloc(&c->d_pl_nn, 8200) ||
        dmalloc(&c->d_pl_lo[0], 8200) || dmalloc(&c->d_pl_lo[1], 8200) || dmalloc(&c->d_pl_hi[0], 8200) || dmalloc(&c->d_pl_hi[1], 8200) || dmalloc(&c->d_scalars, 64) || dmalloc(&c->d_film, (size_t)d.max_filmsize * d.max_filmpasses)) {
        release_ctx(c); return 1;
    }
    PTB_CUDA(cudaMemset(c->d_film, 0, sizeof(float4) * (size_t)d.max_filmsize * d.max_filmpasses));
    PTB_CUDA(cudaMemset(c->d_texels, 0, sizeof(float4) * (size_t)d.max_texels));
    if (dmalloc(&c->d_flags, 4)) { release_ctx(c); return 1; }
    PTB_CUDA(cudaMemset(c->d_flags, 0, 16));
    if (ptb_wf_init(c)) { release_ctx(c); return 1; }
    c->coalesce = getenv("PTB_NO_COALESCE") == nullptr;
    c->fast_shade = getenv("PTB_FAST") != nullptr && getenv("PTB_FAST")[0] == '1';
    // defaults of the reference singletons
    memset(&c->h_params, 0, sizeof c->h_params);
    for (int k = 0; k < 4; k++) c->h_params.world_fac[k] = 0.1f;       // light/world.py:14-16 (tex field zero-initialised)
    c->h_params.world_tex = 0;
    LightRec& L = c->h_params.lights[0];                                // light/__init__.py:22-28 default light
    L.type = PTB_LIGHT_POINT; L.size = 0.5f; L.color = mk3(32.f, 32.f, 32.f); L.pos = mk3(1.f, 2.f, 3.f);
    c->h_params.nlights = 1;
    c->params_dirty = true;
    *out = c;
    return 0;
}

int ptb_destroy(ptb_ctx* c) {
    CHECK_CTX(c);
    c->pend_count = 0;               // recorded but never observed: nothing to render
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    release_ctx(c);
    return 0;
}

int ptb_set_stream(ptb_ctx* c, void* s) { CHECK_FLUSH(c); c->stream = (cudaStream_t)s; return 0; }
int ptb_flush(ptb_ctx* c) { CHECK_FLUSH(c); return 0; }
int ptb_set_mode(ptb_ctx* c, int mode) {
    CHECK_FLUSH(c);
    if (mode != PTB_MODE_PARITY && mode != PTB_MODE_FAST) { ptb_set_error("unknown mode %d", mode); return 1; }
    if (c->fast_shade != (mode == PTB_MODE_FAST)) { c->fast_shade = mode == PTB_MODE_FAST; c->params_dirty = true; }   // the per-scene cache is recomputed by the build in use
    return 0;
}
int ptb_set_option(ptb_ctx* c, const char* name, int value) {
    CHECK_FLUSH(c);
    const std::string k = name ? name : "";
    if (k == "coalesce") c->coalesce = value != 0;
    else if (k == "overlap_shadow") c->overlap_shadow = value != 0;
    else if (k == "pt_lanes" || k == "mlt_lanes") {
        if (value < 1 || value > PTB_MAX_LANES) { ptb_set_error("%s outside [1, %d]", k.c_str(), PTB_MAX_LANES); return 1; }
        (k == "pt_lanes" ? c->pt_lanes : c->mlt_lanes) = value;
    }
    else if (k == "wide4") c->wide4 = value != 0;
    else if (k == "pt_split") c->pt_split = value != 0;
    else if (k == "use_ploc") c->use_ploc = value != 0;
    else if (k == "ploc_big") c->ploc_big = value != 0;
    else if (k == "ploc_radius") { if (value < 1 || value > 1024) { ptb_set_error("ploc_radius outside [1, 1024]"); return 1; } c->ploc_radius = value; }
    else { ptb_set_error("unknown option '%s'", k.c_str()); return 1; }
    return 0;
}
int ptb_get_mode(ptb_ctx* c, int* mode) { CHECK_CTX(c); *mode = c->fast_shade ? PTB_MODE_FAST : PTB_MODE_PARITY; return 0; }
int ptb_synchronize(ptb_ctx* c) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    return ptb_stage_collect(c);
}

// ---- Sobol ------------------------------------------------------------------------------------------------------
int ptb_set_sobol_table(ptb_ctx* c, const int32_t* V, int rows, int dim) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (rows != PTB_SOBOL_ROWS || dim <= 0) { ptb_set_error("Sobol table must be [21][dim]"); return 1; }
    if (c->d_sobolV) cudaFree(c->d_sobolV);
    PTB_CUDA(cudaMalloc(&c->d_sobolV, sizeof(int32_t) * rows * dim));
    PTB_CUDA(cudaMemcpy(c->d_sobolV, V, sizeof(int32_t) * rows * dim, cudaMemcpyHostToDevice));
    c->sobol_dim = dim;
    if (c->d_sobolP) { cudaFree(c->d_sobolP); c->d_sobolP = nullptr; c->sobolP_cap = 0; }
    c->sobol_time = 64;     // reset(): time = 0 then skip = 64 updates (sobol.py:75, 92-97)
    return 0;
}
int ptb_sobol_reset(ptb_ctx* c) { CHECK_FLUSH(c); c->sobol_time = 64; return 0; }
int ptb_sobol_get_time(ptb_ctx* c, int* t) { CHECK_CTX(c); *t = c->sobol_time; return 0; }
int ptb_sobol_set_time(ptb_ctx* c, int t) { CHECK_FLUSH(c); c->sobol_time = t; return 0; }
int ptb_sobol_point(ptb_ctx* c, int k, float* P_out) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    Out<float> o;
    if (o.set(P_out, c->sobol_dim)) return 1;
    if (ptb_wf_sobol_points(c, k, 1, 1, o.dev)) return 1;
    return o.finish(c);
}

// ---- loaders ------------------------------------------------------------------------------------------------------
int ptb_load_model(ptb_ctx* c, const float* verts, const int32_t* mtlids, int nfaces, int memspace) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (nfaces < 0 || !(nfaces < c->caps.max_faces)) { ptb_set_error("too many faces"); return 1; }   // model.py:84
    cudaMemcpyKind kind = memspace == PTB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // mtllib.py:30-38 indexes fac[mtlid] for every id but -1: ids outside [-1, max_materials) would read past the tables.  Host arrays
    // are checked here; device arrays by a kernel whose verdict ptb_build_tree reads (it synchronises anyway).
    PTB_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int32_t), c->stream));
    if (memspace != PTB_DEVICE)
        for (int i = 0; i < nfaces; i++)
            if (mtlids[i] < -1 || mtlids[i] >= c->caps.max_materials) { ptb_set_error("face %d: material id %d outside [-1, %d)", i, mtlids[i], c->caps.max_materials); return 1; }
    if (nfaces) {
        PTB_CUDA(cudaMemcpyAsync(c->d_verts, verts, sizeof(float) * 24 * (size_t)nfaces, kind, c->stream));
        PTB_CUDA(cudaMemcpyAsync(c->d_mtlids, mtlids, sizeof(int32_t) * (size_t)nfaces, kind, c->stream));
    }
    if (nfaces && memspace == PTB_DEVICE && ptb_wf_check_mtlids(c, nfaces)) return 1;
    c->nfaces = nfaces;
    c->tree_n = -1;   // stale until build_tree
    return 0;
}
int ptb_load_materials(ptb_ctx* c, const float* fac, const int32_t* tex, int nmat) {
    CHECK_FLUSH(c);
    if (nmat < 0 || nmat > c->caps.max_materials) { ptb_set_error("too many materials (%d > %d)", nmat, c->caps.max_materials); return 1; }
    for (int i = 0; i < nmat * 12; i++)
        if (tex[i] < -1 || tex[i] >= c->caps.max_textures) { ptb_set_error("material %d slot %d: texture id %d outside [-1, %d)", i / 12, i % 12, tex[i], c->caps.max_textures); return 1; }
    memcpy(c->h_params.mat_fac, fac, sizeof(float) * 48 * nmat);
    memcpy(c->h_params.mat_tex, tex, sizeof(int32_t) * 12 * nmat);
    c->params_dirty = true;
    return 0;
}
int ptb_load_images(ptb_ctx* c, const float* texels, int64_t ntexels, const int32_t* nx, const int32_t* ny, const int32_t* base, int nimg, int memspace) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (nimg > c->caps.max_textures) { ptb_set_error("Out of ID!"); return 1; }          // allocator.py:53
    if (ntexels > c->caps.max_texels) { ptb_set_error("Out of memory!"); return 1; }     // allocator.py:25
    for (int i = 0; i < nimg; i++) {
        if (nx[i] <= 0 || ny[i] <= 0 || base[i] < 0 || (int64_t)base[i] + (int64_t)nx[i] * ny[i] > ntexels) { ptb_set_error("image %d outside the texel arena", i); return 1; }
        c->h_params.img_nx[i] = nx[i]; c->h_params.img_ny[i] = ny[i]; c->h_params.img_base[i] = base[i];
    }
    // ids beyond this load are unloaded again (zero texel, ptb_shade.cuh img_fetch): the reference resets its allocators here
    // (image.py:69-75) and sampling a stale id is undefined there
    for (int i = nimg; i < PTB_MAX_TEXTURES; i++) { c->h_params.img_nx[i] = c->h_params.img_ny[i] = 0; c->h_params.img_base[i] = 0; }
    if (ntexels) PTB_CUDA(cudaMemcpyAsync(c->d_texels, texels, sizeof(float4) * (size_t)ntexels, memspace == PTB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    c->params_dirty = true;
    return 0;
}
int ptb_clear_lights(ptb_ctx* c) { CHECK_FLUSH(c); c->h_params.nlights = 0; c->params_dirty = true; return 0; }
int ptb_add_light(ptb_ctx* c, const float pos[3], const float axes[9], const float color[3], float size, int type) {
    CHECK_FLUSH(c);
    int i = c->h_params.nlights;
    if (i >= c->caps.max_lights) { ptb_set_error("too many lights"); return 1; }
    if (type != PTB_LIGHT_POINT && type != PTB_LIGHT_AREA) { ptb_set_error("unknown light type %d", type); return 1; }
    LightRec& L = c->h_params.lights[i];
    L.type = type; L.size = size;
    L.color = mk3(color[0], color[1], color[2]); L.pos = mk3(pos[0], pos[1], pos[2]);
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) L.axes.m[r][k] = axes[3 * r + k];
    c->h_params.nlights = i + 1;
    c->params_dirty = true;
    return 0;
}
int ptb_set_world_light(ptb_ctx* c, const float fac[4], int tex) {
    CHECK_FLUSH(c);
    if (tex < -1 || tex >= c->caps.max_textures) { ptb_set_error("world light: texture id %d outside [-1, %d)", tex, c->caps.max_textures); return 1; }
    for (int k = 0; k < 4; k++) c->h_params.world_fac[k] = fac[k];
    c->h_params.world_tex = tex;
    c->params_dirty = true;
    return 0;
}
int ptb_set_camera(ptb_ctx* c, const float v2w[16], const float w2v[16]) {
    CHECK_FLUSH(c);
    memcpy(c->h_params.v2w, v2w, 64);
    if (w2v) memcpy(c->h_w2v, w2v, 64);
    c->params_dirty = true;
    return 0;
}

// ---- tree ------------------------------------------------------------------------------------------------------------
int ptb_build_tree(ptb_ctx* c, ptb_tree_info* info) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    int rc = ptb_lbvh_build(c);
    if (info) *info = c->tree_info;
    if (rc == 0) {
        int32_t bad = 0;
        PTB_CUDA(cudaMemcpyAsync(&bad, c->d_flags, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        PTB_CUDA(cudaStreamSynchronize(c->stream));
        if (bad) { c->tree_n = -1; ptb_set_error("the loaded model has material ids outside [-1, %d)", c->caps.max_materials); return 1; }
    }
    return rc;
}
int ptb_set_traversal(ptb_ctx* c, int policy) {
    CHECK_FLUSH(c);
    if (policy < 0 || policy > PTB_TRAVERSE_ORDERED_EXACT) { ptb_set_error("unknown traversal policy %d", policy); return 1; }
    c->traversal_request = policy;
    c->tree_info.policy = ptb_effective_policy(c, policy);
    return 0;
}
int ptb_export_tree(ptb_ctx* c, int32_t* mc, int32_t* id, int32_t* child, int32_t* leaf, float* bmin, float* bmax) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    int n = c->tree_n;
    if (n < 0) { ptb_set_error("no tree built"); return 1; }
    cudaStream_t s = c->stream;
    if (mc && n) PTB_CUDA(cudaMemcpyAsync(mc, c->d_mc, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (id && n) PTB_CUDA(cudaMemcpyAsync(id, c->d_id, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (leaf && n) PTB_CUDA(cudaMemcpyAsync(leaf, c->d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (n > 1) {
        if (child) PTB_CUDA(cudaMemcpyAsync(child, c->d_child, 8 * (size_t)(n - 1), cudaMemcpyDeviceToHost, s));
        if (bmin) PTB_CUDA(cudaMemcpyAsync(bmin, c->d_bmin, 12 * (size_t)(n - 1), cudaMemcpyDeviceToHost, s));
        if (bmax) PTB_CUDA(cudaMemcpyAsync(bmax, c->d_bmax, 12 * (size_t)(n - 1), cudaMemcpyDeviceToHost, s));
    }
    PTB_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int ptb_export_traversal(ptb_ctx* c, float* nodes, float* leaf_lo, float* leaf_hi, float* gbox, int32_t* ploc) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    const int n = c->tree_n;
    if (n < 2) { ptb_set_error("no traversal structure (fewer than 2 faces)"); return 1; }
    cudaStream_t s = c->stream;
    const Node64* src = c->d_nodes_active ? c->d_nodes_active : c->d_nodes;
    if (nodes) PTB_CUDA(cudaMemcpyAsync(nodes, src, sizeof(Node64) * (size_t)(n - 1), cudaMemcpyDeviceToHost, s));
    if (leaf_lo) PTB_CUDA(cudaMemcpyAsync(leaf_lo, c->d_tlo, 16 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (leaf_hi) PTB_CUDA(cudaMemcpyAsync(leaf_hi, c->d_thi, 16 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (gbox) PTB_CUDA(cudaMemcpyAsync(gbox, c->d_gbox, 32 * (size_t)n, cudaMemcpyDeviceToHost, s));
    PTB_CUDA(cudaStreamSynchronize(s));
    if (ploc) *ploc = c->d_nodes_active == c->d_nodes2;
    return 0;
}

// ---- film ------------------------------------------------------------------------------------------------------------
int ptb_set_size(ptb_ctx* c, int nx, int ny) {
    CHECK_FLUSH(c);
    if (nx <= 0 || ny <= 0 || (int64_t)nx * ny > c->caps.max_filmsize) { ptb_set_error("film %dx%d exceeds max_filmsize %d", nx, ny, c->caps.max_filmsize); return 1; }
    c->nx = nx; c->ny = ny;
    c->params_dirty = true;
    return 0;
}
int ptb_get_size(ptb_ctx* c, int* nx, int* ny) { CHECK_CTX(c); *nx = c->nx; *ny = c->ny; return 0; }
int ptb_clear(ptb_ctx* c) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    PTB_CUDA(cudaMemsetAsync(c->d_film, 0, sizeof(float4) * (size_t)c->caps.max_filmsize * c->caps.max_filmpasses, c->stream));
    return 0;
}
int ptb_film_ptr(ptb_ctx* c, int pass, void** dev_ptr, int64_t* ntexels) {
    CHECK_FLUSH(c);
    if (pass < 0 || pass >= c->caps.max_filmpasses) { ptb_set_error("film pass %d out of range", pass); return 1; }
    *dev_ptr = c->d_film + (size_t)pass * c->caps.max_filmsize;
    if (ntexels) *ntexels = (int64_t)c->nx * c->ny;
    return 0;
}

// ---- render ----------------------------------------------------------------------------------------------------------
static int render_ready(ptb_ctx* c, int engine) {
    if (engine < PTB_ENGINE_PATH || engine > PTB_ENGINE_MLT) { ptb_set_error("unknown engine %d", engine); return 1; }
    if (c->nx <= 0 || c->ny <= 0) { ptb_set_error("film size not set (ptb_set_size)"); return 1; }
    if (c->tree_n != c->nfaces) { ptb_set_error("BVH is stale: call ptb_build_tree after ptb_load_model"); return 1; }
    if (engine != PTB_ENGINE_MLT && !c->d_sobolV) { ptb_set_error("Sobol direction table not set (ptb_set_sobol_table)"); return 1; }
    return 0;
}
int ptb_render(ptb_ctx* c, int engine, int nsamples) {
    CHECK_CTX(c);
    if (nsamples <= 0) return 0;
    if (render_ready(c, engine)) return 1;          // errors surface at the call that causes them, not at the deferred submission
    if (engine == PTB_ENGINE_MLT) {                  // chains: every call depends on the one before; nothing to merge
        if (flush_pending(c)) return 1;
        if (c->mlt_count <= 0) { ptb_set_error("MLT chains not initialised (ptb_mlt_reset)"); return 1; }
        DeviceGuard g(c->device);
        return ptb_wf_render(c, engine, 0, nsamples, 1, nullptr);
    }
    if (c->pend_count > 0 && c->pend_engine != engine && flush_pending(c)) return 1;
    if (c->pend_count == 0) { c->pend_engine = engine; c->pend_first = c->sobol_time + 1; }   // update() runs before _render (path.py:75-77)
    c->pend_count += nsamples;
    c->sobol_time += nsamples;
    const int cap = ptb_wf_batch_capacity(c);
    if (!c->coalesce || c->profiling) return flush_pending(c);
    // full batches go out at once (the device starts while the caller keeps recording); the remainder waits
    if (c->pend_count >= cap) {
        const int full = c->pend_count / cap * cap, first = c->pend_first;
        c->pend_first += full; c->pend_count -= full;
        if (c->pend_count == 0) c->pend_engine = -1;
        DeviceGuard g(c->device);
        return ptb_wf_render(c, engine, first, full, 1, nullptr);
    }
    return 0;
}
// engine/path.py:96-118 render_tile(i, j, samples): SobolSampler().update(), then samples m = 0..min(samples, 63) (inclusive, sic:
// `if m > samples: continue`) of every pixel of the 64x64 tile (i, j), each with the rotation wanghash3(x, y, m) of the one new
// Sobol point.  Pixels outside the film are skipped (the reference text's `x > nx` / `y > ny` would write one row / column past it).
int ptb_render_tile(ptb_ctx* c, int engine, int i, int j, int samples) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (render_ready(c, engine)) return 1;
    if (engine != PTB_ENGINE_PATH && engine != PTB_ENGINE_BRUTE) { ptb_set_error("render_tile: path or brute engine only"); return 1; }
    c->sobol_time += 1;
    if (i < 0 || j < 0 || i * 64 >= c->nx || j * 64 >= c->ny || samples < 0) return 0;
    const int window[4] = {i * 64, j * 64, 64, 64};
    return ptb_wf_render(c, engine, c->sobol_time, (samples < 63 ? samples : 63) + 1, 1, nullptr, window);
}
// engine/path.py:120-128 render_final(nsamples): every tile, `samples = nsamples; while samples > 0: render_tile(i, j, samples); samples -= 64`
int ptb_render_final(ptb_ctx* c, int engine, int nsamples) {
    CHECK_FLUSH(c);
    for (int i = 0; i < (c->nx + 63) / 64; i++)
        for (int j = 0; j < (c->ny + 63) / 64; j++)
            for (int samples = nsamples; samples > 0; samples -= 64)
                if (ptb_render_tile(c, engine, i, j, samples)) return 1;
    return 0;
}
int ptb_render_range(ptb_ctx* c, int engine, int k_first, int count, int stride) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (count <= 0) return 0;
    return ptb_wf_render(c, engine, k_first, count, stride, nullptr);
}
int ptb_render_sample(ptb_ctx* c, int engine, int k, float* out) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    Out<float> o;
    if (o.set(out, (size_t)c->nx * c->ny * 3)) return 1;
    if (ptb_wf_render(c, engine, k, 1, 1, o.dev)) return 1;
    return o.finish(c);
}
int ptb_mlt_reset(ptb_ctx* c, uint64_t seed, int chain_first, int chain_count) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (chain_count <= 0 || chain_count > c->max_paths) { ptb_set_error("chain count %d outside (0, %lld]", chain_count, (long long)c->max_paths); return 1; }
    if (chain_count != c->mlt_count) {
        cudaFree(c->d_Xold); cudaFree(c->d_Xnew); cudaFree(c->d_Lold);        // cudaFree(nullptr) is a no-op
        c->d_Xold = c->d_Xnew = nullptr; c->d_Lold = nullptr; c->mlt_count = 0;  // a failed allocation below leaves no dangling state
        if (dmalloc(&c->d_Xold, (size_t)chain_count * 32) || dmalloc(&c->d_Xnew, (size_t)chain_count * 32) || dmalloc(&c->d_Lold, (size_t)chain_count)) {
            cudaFree(c->d_Xold); cudaFree(c->d_Xnew); cudaFree(c->d_Lold);
            c->d_Xold = c->d_Xnew = nullptr; c->d_Lold = nullptr;
            return 1;
        }
    }
    c->mlt_seed = seed; c->mlt_first = chain_first; c->mlt_count = chain_count;
    return ptb_wf_mlt_reset(c);
}
int ptb_mlt_state(ptb_ctx* c, int count, float* x_new, float* l_new, float* x_old, float* l_old) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    int n = c->mlt_count;
    if (n <= 0) { ptb_set_error("MLT chains not initialised (ptb_mlt_reset)"); return 1; }
    if (count != n) { ptb_set_error("ptb_mlt_state: buffers sized for %d chains, the context runs %d", count, n); return 1; }
    cudaStream_t s = c->stream;
    std::vector<float> tmp((size_t)n * 4);
    if (x_new) PTB_CUDA(cudaMemcpyAsync(x_new, c->d_Xnew, sizeof(float) * 32 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (x_old) PTB_CUDA(cudaMemcpyAsync(x_old, c->d_Xold, sizeof(float) * 32 * (size_t)n, cudaMemcpyDeviceToHost, s));
    for (int which = 0; which < 2; which++) {
        float* dst = which == 0 ? l_new : l_old;
        if (!dst) continue;
        PTB_CUDA(cudaMemcpyAsync(tmp.data(), which == 0 ? (const void*)c->st.result : (const void*)c->d_Lold, sizeof(float) * 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
        PTB_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < n; i++) { dst[3 * i] = tmp[4 * i]; dst[3 * i + 1] = tmp[4 * i + 1]; dst[3 * i + 2] = tmp[4 * i + 2]; }
    }
    PTB_CUDA(cudaStreamSynchronize(s));
    return 0;
}
int ptb_mlt_set_param(ptb_ctx* c, float lsp, float sigma) { CHECK_FLUSH(c); c->mlt_lsp = lsp; c->mlt_sigma = sigma; return 0; }

static int resolve_common(ptb_ctx* c, int pass, int mode, float* out, int memspace, size_t count) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (pass < 0 || pass >= c->caps.max_filmpasses) { ptb_set_error("film pass %d out of range", pass); return 1; }
    if (memspace == PTB_DEVICE) return ptb_wf_resolve(c, pass, mode, out);
    // host destination: resolve into a context-owned staging buffer (grow-only: no cudaMalloc / cudaFree on the per-frame path)
    if (count > c->resolve_cap) {
        if (c->d_resolve) cudaFree(c->d_resolve);
        c->d_resolve = nullptr; c->resolve_cap = 0;
        PTB_CUDA(cudaMalloc((void**)&c->d_resolve, sizeof(float) * count));
        c->resolve_cap = count;
    }
    if (ptb_wf_resolve(c, pass, mode, c->d_resolve)) return 1;
    PTB_CUDA(cudaMemcpyAsync(out, c->d_resolve, sizeof(float) * count, cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int ptb_get_image(ptb_ctx* c, int pass, float* out, int memspace) { return resolve_common(c, pass, 0, out, memspace, c ? (size_t)c->nx * c->ny * 4 : 0); }
int ptb_fast_export_image(ptb_ctx* c, int pass, float* out, int memspace) { return resolve_common(c, pass, 1, out, memspace, c ? (size_t)c->nx * c->ny * 3 : 0); }
int ptb_get_film(ptb_ctx* c, int pass, float* out, int memspace) { return resolve_common(c, pass, 2, out, memspace, c ? (size_t)c->nx * c->ny * 4 : 0); }

// CUDA-GL interop, declared here because the image has no GL headers for <cuda_gl_interop.h> (GLuint = unsigned int)
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);
int ptb_fast_export_gl(ptb_ctx* c, int pass, unsigned int gl_buffer) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (pass < 0 || pass >= c->caps.max_filmpasses) { ptb_set_error("film pass %d out of range", pass); return 1; }
    cudaGraphicsResource* res = nullptr;
    cudaError_t e = cudaGraphicsGLRegisterBuffer(&res, gl_buffer, cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ptb_set_error("cannot register GL buffer %u with CUDA (%s): is an OpenGL context current on this thread?", gl_buffer, cudaGetErrorName(e));
        return 1;
    }
    int rc = 1;
    void* dev = nullptr; size_t bytes = 0;
    if ((e = cudaGraphicsMapResources(1, &res, c->stream)) != cudaSuccess) ptb_set_error("cudaGraphicsMapResources: %s", cudaGetErrorName(e));
    else {
        if ((e = cudaGraphicsResourceGetMappedPointer(&dev, &bytes, res)) != cudaSuccess) ptb_set_error("cudaGraphicsResourceGetMappedPointer: %s", cudaGetErrorName(e));
        else if (bytes < sizeof(float) * 3 * (size_t)c->nx * c->ny) ptb_set_error("GL buffer holds %zu bytes, the %dx%d image needs %zu", bytes, c->nx, c->ny, sizeof(float) * 3 * (size_t)c->nx * c->ny);
        else rc = ptb_wf_resolve(c, pass, 1, (float*)dev);
        cudaGraphicsUnmapResources(1, &res, c->stream);
    }
    cudaGraphicsUnregisterResource(res);
    return rc;
}

// ---- taps ------------------------------------------------------------------------------------------------------------
int ptb_trace_primary(ptb_ctx* c, int k, float* rays, int32_t* hit, float* depth, int32_t* index, float* uv) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    size_t np = (size_t)c->nx * c->ny;
    Out<float> o_rays, o_depth, o_uv; Out<int32_t> o_hit, o_index;
    if (o_rays.set(rays, np * 6) || o_depth.set(depth, np) || o_uv.set(uv, np * 2) || o_hit.set(hit, np) || o_index.set(index, np)) return 1;
    if (ptb_wf_trace_primary(c, k, o_rays.dev, o_hit.dev, o_depth.dev, o_index.dev, o_uv.dev)) return 1;
    return o_rays.finish(c) || o_depth.finish(c) || o_uv.finish(c) || o_hit.finish(c) || o_index.finish(c);
}
int ptb_intersect(ptb_ctx* c, const float* rays, const int32_t* avoid, int m, int policy, int32_t* hit, float* depth, int32_t* index, float* uv) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (m <= 0) return 0;
    In<float> i_rays; In<int32_t> i_avoid;
    Out<float> o_depth, o_uv; Out<int32_t> o_hit, o_index;
    if (i_rays.set(c, rays, (size_t)m * 6) || i_avoid.set(c, avoid, m)) return 1;
    if (o_hit.set(hit, m) || o_depth.set(depth, m) || o_index.set(index, m) || o_uv.set(uv, (size_t)m * 2)) return 1;
    if (!o_hit.dev || !o_depth.dev || !o_index.dev || !o_uv.dev) { ptb_set_error("ptb_intersect needs all four outputs"); return 1; }
    if (ptb_wf_intersect(c, i_rays.dev, i_avoid.dev, nullptr, m, policy, 0, o_hit.dev, o_depth.dev, o_index.dev, o_uv.dev)) return 1;
    return o_hit.finish(c) || o_depth.finish(c) || o_index.finish(c) || o_uv.finish(c);
}
int ptb_occluded(ptb_ctx* c, const float* rays, const int32_t* avoid, const float* dis, int m, int policy, int32_t* occluded) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (m <= 0) return 0;
    if (!rays || !dis || !occluded) { ptb_set_error("ptb_occluded needs rays, dis and occluded"); return 1; }
    In<float> i_rays, i_dis; In<int32_t> i_avoid; Out<int32_t> o;
    if (i_rays.set(c, rays, (size_t)m * 6) || i_dis.set(c, dis, m) || i_avoid.set(c, avoid, m) || o.set(occluded, m)) return 1;
    if (ptb_wf_intersect(c, i_rays.dev, i_avoid.dev, i_dis.dev, m, policy, 1, o.dev, nullptr, nullptr, nullptr)) return 1;
    return o.finish(c);
}
static int shade_tap(ptb_ctx* c, int what, const float* in0, size_t n0, const float* in1, size_t n1, const int32_t* ini, int m, float* out, size_t nout) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (m <= 0) return 0;
    In<float> a, b; In<int32_t> ii; Out<float> o;
    if (a.set(c, in0, n0) || b.set(c, in1, n1) || ii.set(c, ini, m) || o.set(out, nout)) return 1;
    if (ptb_wf_shade_tap(c, what, a.dev, b.dev, ii.dev, m, o.dev)) return 1;
    return o.finish(c);
}
int ptb_eval_bsdf(ptb_ctx* c, const float* params, const float* geom, int m, float* out) { return shade_tap(c, 0, params, (size_t)m * 14, geom, (size_t)m * 10, nullptr, m, out, (size_t)m * 3); }
int ptb_sample_bsdf(ptb_ctx* c, const float* params, const float* geom, int m, float* out) { return shade_tap(c, 1, params, (size_t)m * 14, geom, (size_t)m * 10, nullptr, m, out, (size_t)m * 7); }
int ptb_eval_bsdf_literal(ptb_ctx* c, const float* params, const float* geom, int m, float* out) { return shade_tap(c, 6, params, (size_t)m * 14, geom, (size_t)m * 10, nullptr, m, out, (size_t)m * 3); }
int ptb_sample_bsdf_literal(ptb_ctx* c, const float* params, const float* geom, int m, float* out) { return shade_tap(c, 7, params, (size_t)m * 14, geom, (size_t)m * 10, nullptr, m, out, (size_t)m * 7); }
int ptb_material_get(ptb_ctx* c, const int32_t* mtlid, const float* uv, int m, float* out) { return shade_tap(c, 2, uv, (size_t)m * 2, nullptr, 0, mtlid, m, out, (size_t)m * 14); }
int ptb_light_hit(ptb_ctx* c, const float* rays, int m, float* out) { return shade_tap(c, 3, rays, (size_t)m * 6, nullptr, 0, nullptr, m, out, (size_t)m * 6); }
int ptb_light_sample(ptb_ctx* c, const float* in, int m, float* out) { return shade_tap(c, 4, in, (size_t)m * 6, nullptr, 0, nullptr, m, out, (size_t)m * 8); }
int ptb_world_at(ptb_ctx* c, const float* dirs, int m, float* out) { return shade_tap(c, 5, dirs, (size_t)m * 3, nullptr, 0, nullptr, m, out, (size_t)m * 3); }

// common.py:346-352 normaldist (with erfinv :337-343): out[i] = sqrt(2) * erfinv(2 in[i] - 1) -- the MLT small-step kernel (mltpath.py:64)
int ptb_normaldist(ptb_ctx* c, const float* in, int m, float* out) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (m <= 0) return 0;
    In<float> a; Out<float> o;
    if (a.set(c, in, m) || o.set(out, m)) return 1;
    if (ptb_wf_normaldist(c, a.dev, m, o.dev)) return 1;
    return o.finish(c);
}

int ptb_set_counting(ptb_ctx* c, int enabled) { CHECK_FLUSH(c); c->counting = enabled & 1; c->profiling = (enabled >> 1) & 1; return 0; }
int ptb_get_counters(ptb_ctx* c, ptb_counters* out) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    DevCounters h;
    PTB_CUDA(cudaMemcpyAsync(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    out->extend_rays = (int64_t)h.extend_rays; out->shadow_rays = (int64_t)h.shadow_rays; out->rays = out->extend_rays + out->shadow_rays;
    out->node_visits = (int64_t)h.nodes; out->box_tests = (int64_t)h.boxes; out->tri_tests = (int64_t)h.tris; out->paths = (int64_t)h.paths;
    out->max_stack = h.max_stack;
    return 0;
}
int ptb_reset_counters(ptb_ctx* c) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    PTB_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), c->stream));
    if (ptb_stage_collect(c)) return 1;
    for (float& f : c->stage_ms) f = 0.0f;
    c->launches = 0;
    return 0;
}
int ptb_get_stage_ms(ptb_ctx* c, float ms[5]) {
    CHECK_FLUSH(c);
    DeviceGuard g(c->device);
    if (ptb_stage_collect(c)) return 1;
    for (int i = 0; i < 5; i++) ms[i] = c->stage_ms[i];
    return 0;
}
int ptb_get_launches(ptb_ctx* c, int64_t* n) { CHECK_CTX(c);   /* host-side counter: recorded render calls stay recorded */ *n = c->launches; return 0; }
int ptb_selftest(ptb_ctx* c, int what, int64_t n, uint64_t seed, int64_t* fails) {
    CHECK_CTX(c);
    DeviceGuard g(c->device);
    long long f = 0;
    int rc = ptb_wf_selftest(c, what, (long long)n, (unsigned long long)seed, &f);
    *fails = f;
    return rc;
}
int ptb_measure_l2(ptb_ctx* c, int mbytes, int iters, float* gbps) {
    CHECK_CTX(c);
    DeviceGuard g(c->device);
    return ptb_wf_measure_l2(c, mbytes, iters, gbps);
}

}  // extern "C"
