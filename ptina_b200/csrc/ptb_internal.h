// ptb_internal.h -- context layout and the launch functions shared between api.cu, lbvh.cu, wavefront.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/ptina_b200.h"
#include "ptb_shade.cuh"
#include "ptb_traverse.cuh"

#define PTB_SOBOL_ROWS 21
#define PTB_LIST_CAP 32

void ptb_set_error(const char* fmt, ...);
#define PTB_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ptb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

// SoA wavefront state, one entry per path in flight (HBM; 5 x 16 B per path)
struct PathState {
    float4* ray_o;    // camera ray origin (written by raygen; read by the taps and the preview pass)
    float4* ray_d;    // camera ray direction as Camera.generate returns it
    float4* hit;      // depth, u, v, face id (int bits; -1 = miss)                        -- written by extend
    float4* thr;      // throughput rgb (path.py:21), w = depth counter (int bits)
    float4* result;   // radiance rgb (path.py:20), w = last_brdf_pdf (path.py:22,61)
};
// Ray queues hold self-contained records in queue order (coalesced for the consumer, stageable with cp.async):
//   extend : o = (origin, path slot)  d = (unit direction, avoid leaf slot)               -- written by raygen / shade
//   shadow : o = (origin, path slot)  d = (direction, distance to the light sample)  c = (contribution if unoccluded, avoid leaf slot)
struct RayQueue { float4* o; float4* d; float4* c; };
// tree queue (k_trace_pre -> k_trace_tree): 80-byte records, see ptb_trace_kernel.cuh
#define PTB_EXP_K 5
struct ExpQ { float4* e[PTB_EXP_K]; };

// device-side control block: queue sizes and work cursors
struct Ctrl {
    int n_in;        // entries in the active queue being consumed
    int cur_extend;
    int n_out;       // entries appended to the next active queue   } 8-byte aligned pair: shade reserves both with ONE 64-bit
    int n_shadow;    // entries in the shadow queue                 } atomic per block (low word n_out, high word n_shadow)
    int cur_shadow;
    int pad[3];
    int n_tree[2];   // entries of the tree queue written by k_trace_pre: [0] extend, [1] shadow / taps
    int cur_tree[2];
};

struct DevCounters {
    unsigned long long extend_rays, shadow_rays, nodes, boxes, tris, paths;
    unsigned int max_stack, pad;
};

// A frame covers the whole film, or -- render_tile (engine/path.py:96-118) -- one 64x64 window of it at (x0, y0); in window mode
// every sample m of a pixel uses the SAME Sobol point with its own dimension rotation wanghash3(x, y, m) (path.py:115).
struct FrameMap { int nx, ny, tiles_y, pps, x0, y0, window, slot_base; FastDiv d_pps, d_ty; FastMod d_dim; };   // pps = path slots per sample (multiple of 32); slot_base: first path slot of the lane
__host__ __device__ inline FrameMap make_frame(int nx, int ny) {
    FrameMap f; f.nx = nx; f.ny = ny; f.x0 = 0; f.y0 = 0; f.window = 0; f.slot_base = 0;
    int tx = (nx + 7) / 8; f.tiles_y = (ny + 3) / 4;
    f.pps = tx * f.tiles_y * 32;
    f.d_pps = make_fastdiv((unsigned)f.pps); f.d_ty = make_fastdiv((unsigned)f.tiles_y); f.d_dim = make_fastmod(0);      // d_dim: set by the caller that knows the Sobol table
    return f;
}
__host__ __device__ inline FrameMap make_window(int nx, int ny, int x0, int y0, int w, int h) {
    FrameMap f; f.nx = nx; f.ny = ny; f.x0 = x0; f.y0 = y0; f.window = 1; f.slot_base = 0;
    int tx = (w + 7) / 8; f.tiles_y = (h + 3) / 4;
    f.pps = tx * f.tiles_y * 32;
    f.d_pps = make_fastdiv((unsigned)f.pps); f.d_ty = make_fastdiv((unsigned)f.tiles_y); f.d_dim = make_fastmod(0);      // d_dim: set by the caller that knows the Sobol table
    return f;
}

// A lane = one wavefront in flight: its queues, control block, streams and events.  Lane 0 is the context's own (whole pools, the
// caller's stream); lane 1 works in the upper half of every pool on streams of its own, so two independent populations of paths
// (the two halves of the MLT chains) can be on the device at once -- each fills the SMs the other's small kernels leave idle.
#define PTB_MAX_LANES 4
struct Ctrl;
struct Lane {
    RayQueue xq[2]; RayQueue sq; ExpQ tq, tq2;
    Ctrl* ctrl = nullptr;
    cudaStream_t s_main = nullptr, s_side = nullptr;
    cudaEvent_t ev_shade = nullptr, ev_shadow = nullptr;
};

enum { ST_RAYGEN = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_ACCUM, ST_COUNT };

struct StageEvent { int stage; cudaEvent_t a, b; };

struct ptb_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    ptb_caps caps{};
    int sm_count = 148;

    // ---- scene (reference layout) ----
    int nfaces = 0;
    float* d_verts = nullptr;       // [nfaces*3][8]   model.py:14
    int32_t* d_mtlids = nullptr;    // [nfaces]        model.py:15
    float4* d_texels = nullptr;     // arena           image.py:17
    SceneParams h_params{};         // host mirror
    SceneParams* d_params = nullptr;
    SceneCache* d_cache = nullptr;  // per-scene constants derived from d_params (k_prepare_cache)
    bool params_dirty = true;
    float h_w2v[16]{};

    // ---- sobol ----
    int32_t* d_sobolV = nullptr;    // [21][dim]
    int sobol_dim = 0;
    int sobol_time = 0;
    float* d_sobolP = nullptr;      // [spp_batch][dim] points of the batch in flight
    int sobolP_cap = 0;

    // ---- tree (reference arrays, lbvh.py:50-59) + packed traversal data ----
    int tree_n = 0;
    int32_t *d_mc = nullptr, *d_id = nullptr, *d_mc_tmp = nullptr, *d_id_tmp = nullptr, *d_leaf = nullptr;
    int2* d_child = nullptr;
    float *d_bmin = nullptr, *d_bmax = nullptr;
    int32_t* d_ready = nullptr;     // sweep stamp per internal node
    int32_t* d_parentcnt = nullptr; // reference counts for validation [2n]
    int2* d_range = nullptr;        // (min slot, max slot) per internal node
    int32_t* d_height = nullptr;
    Node64* d_nodes = nullptr;
    Node64* d_nodes2 = nullptr;     // PLOC traversal tree of small scenes (lbvh.cu k_ploc_small); d_nodes_active = the one traversed
    Node64* d_nodes_active = nullptr;
    int *d_pl_id[2]{}, *d_pl_depth[2]{}, *d_pl_nn = nullptr; float4 *d_pl_lo[2]{}, *d_pl_hi[2]{};   // PLOC cluster buffers
    int *d_pl_keep = nullptr, *d_pl_make = nullptr, *d_pl_kpos = nullptr, *d_pl_mpos = nullptr, *d_pl_S = nullptr;   // multi-block PLOC (big trees)
    int pl_cap = 8200;              // entries the PLOC buffers and d_nodes2 hold
    bool ploc_big = true;           // PTB_NO_PLOC_BIG=1: trees beyond 8193 triangles keep the LBVH topology for traversal
    int ploc_radius = 4;            // PTB_PLOC_RADIUS: search radius of the multi-block PLOC (config 4: 2..4 give the fewest node visits, tools/mega_sweep.py)
    int ploc_rounds = 0;
    bool use_ploc = true;           // PTB_NO_PLOC=1: keep the LBVH topology for traversal
    int trav_depth = -1;
    uint4* d_qnodes = nullptr;      // [2(n-1)] quantised copy of d_nodes (Node32), built for trees that are traversed out of global memory
    uint4* d_wnodes = nullptr;      // [4(n-1)] 4-wide quantised nodes (lbvh.cu k_wide4_nodes), global-memory trees only
    int wnodes_cap = 0; bool wnodes_ok = false;
    bool s16_stack = true;          // PTB_NO_S16=1: resident tree kernels keep the 64-bit stack entries for shallow trees too
    bool wide4 = true;              // "wide4" / PTB_NO_WIDE4=1: walk big trees through the 4-wide nodes
    float qbase[3]{}, qext[3]{1.0f, 1.0f, 1.0f}, qinv[3]{1.0f, 1.0f, 1.0f};
    Tri64* d_tris = nullptr;
    int32_t* d_slot_of = nullptr;   // face id -> leaf slot
    int32_t* d_gate = nullptr;      // leaf slot -> parent internal node (the box that gates its triangle test)
    float4* d_gbox = nullptr;       // [2n] per leaf slot: the reference box (lo, hi) of its gate
    float4 *d_tlo = nullptr, *d_thi = nullptr;   // [n] per leaf slot: inflated triangle bounds (w of tlo: PTB_TF_* flags)
    float4 *d_nlo = nullptr, *d_nhi = nullptr;   // [n-1] per internal node: traversal box (union of the unlisted leaves below)
    int32_t* d_list = nullptr;      // always-test list (leaf slots), PTB_LIST_CAP entries
    int list_n = 0;
    int root_must = 0;
    float scene_abs = 0.0f;         // largest absolute coordinate of the reference root box and the traversal root box
    void* d_sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    int32_t* d_scalars = nullptr;   // small device scratch (bounds as ordered ints, flags)
    ptb_tree_info tree_info{};
    int traversal_request = PTB_TRAVERSE_AUTO;
    float root_lo[3]{}, root_hi[3]{};

    // ---- film (filmtable.py:12-14) ----
    int nx = 0, ny = 0;
    float4* d_film = nullptr;       // [passes][max_filmsize]
    float* d_resolve = nullptr; size_t resolve_cap = 0;   // staging for host-side get_image / fast_export_image

    // ---- wavefront ----
    int64_t max_paths = 0;
    PathState st{};
    RayQueue xq[2]{};               // extend queues (in / out, swapped every bounce)
    RayQueue sq{};                  // shadow queue
    ExpQ tq{};                      // tree queue
    ExpQ tq2{};                     // second tree queue: the shadow stage of bounce b runs beside the extend stage of bounce b+1
    cudaStream_t stream2 = nullptr; // non-blocking side stream of the shadow stage
    cudaEvent_t ev_shade = nullptr, ev_shadow = nullptr;
    // lanes 1..PTB_MAX_LANES-1: main and side stream, shade / shadow events (lane 0 uses stream / stream2 / ev_shade / ev_shadow)
    cudaStream_t lane_main[PTB_MAX_LANES]{}, lane_side[PTB_MAX_LANES]{};
    cudaEvent_t lane_shade[PTB_MAX_LANES]{}, lane_shadow[PTB_MAX_LANES]{}, lane_join[PTB_MAX_LANES]{}, ev_fork = nullptr;
    cudaEvent_t ev_acc[PTB_MAX_LANES]{};   // accumulate of the last chunk of each lane (chunks add to the film in Sobol order)
    int mlt_lanes = 2;              // "mlt_lanes" / PTB_MLT_LANES: concurrent sub-populations of the MLT chains (1 = one wavefront)
    bool pt_split = false;          // "pt_split": also split a render that fits one batch into pt_lanes concurrent chunks (measurement)
    int pt_lanes = 2;               // "pt_lanes" / PTB_PT_LANES: lanes the chunks of a multi-batch render alternate between
    bool overlap_shadow = true;     // PTB_NO_OVERLAP=1 turns the overlap off
    Ctrl* d_ctrl = nullptr;
    DevCounters* d_counters = nullptr;
    int counting = 0;
    int smem_optin = 0;             // opt-in shared memory per block (227 KB on B200)
    bool quant_resident_bvh = false; // PTB_QUANT_RESIDENT_BVH=1: resident trees always as quantised nodes (testing)
    bool no_resident_bvh = false;   // PTB_NO_RESIDENT_BVH=1: never use the shared-memory-resident traversal variant
    int blocks_extend = 0, blocks_shadow = 0, blocks_ref = 0, blocks_generic = 0, blocks_exact = 0;
    std::vector<StageEvent> events;
    std::vector<cudaEvent_t> event_pool;
    int profiling = 0;
    float stage_ms[ST_COUNT]{};
    long long launches = 0;

    // ---- MLT (mltpath.py:9-36) ----
    float *d_Xold = nullptr, *d_Xnew = nullptr;   // [nchains][32]
    float4* d_Lold = nullptr;                      // [nchains]
    int mlt_first = 0, mlt_count = 0;
    uint64_t mlt_seed = 0; uint32_t mlt_iter = 0;
    float mlt_lsp = 0.25f, mlt_sigma = 0.01f;

    // ---- deferred Engine.render() calls (api.cu): consecutive one-sample calls are submitted as one wavefront batch ----
    int pend_engine = -1, pend_first = 0, pend_count = 0;
    bool coalesce = true;           // PTB_NO_COALESCE=1: submit every ptb_render call at once
    bool fast_shade = false;        // PTB_MODE_FAST: the shading stage compiled with FMA contraction and approximate division (shade.cu)
    int32_t* d_flags = nullptr;     // [4] device flags: [0] material id out of range in the loaded model
};

TraceScene ptb_trace_scene(const ptb_ctx* c);
int ptb_effective_policy(const ptb_ctx* c, int requested);
enum { PTB_TREE_RESIDENT = 0, PTB_TREE_RESIDENT_QUANT = 1, PTB_TREE_GLOBAL = 2 };
int ptb_tree_mode(const ptb_ctx* c, int n);               // wavefront.cu: where the tree kernel keeps the nodes

// lbvh.cu
int ptb_lbvh_build(ptb_ctx* c);
// wavefront.cu
int ptb_wf_init(ptb_ctx* c);
int ptb_wf_upload_params(ptb_ctx* c);
int ptb_wf_render(ptb_ctx* c, int engine, int k_first, int count, int stride, float* sample_out_dev, const int* window = nullptr);
int ptb_wf_batch_capacity(const ptb_ctx* c);             // samples of the current film that fit one wavefront batch
int ptb_wf_normaldist(ptb_ctx* c, const float* in_dev, int m, float* out_dev);
int ptb_wf_check_mtlids(ptb_ctx* c, int nfaces);                    // flags material ids outside [-1, max_materials) in d_flags[0]
int ptb_wf_sobol_points(ptb_ctx* c, int k_first, int count, int stride, float* P_dev);
int ptb_wf_trace_primary(ptb_ctx* c, int k, float* rays_dev, int32_t* hit_dev, float* depth_dev, int32_t* index_dev, float* uv_dev);
int ptb_wf_intersect(ptb_ctx* c, const float* rays_dev, const int32_t* avoid_dev, const float* dis_dev, int m, int policy, int anyhit,
                     int32_t* hit_dev, float* depth_dev, int32_t* index_dev, float* uv_dev);
int ptb_wf_resolve(ptb_ctx* c, int pass, int mode, float* out_dev);
// shade.cu, compiled twice (strict / fast arithmetic); the dispatchers below pick by c->fast_shade
void ptb_shade_prepare_cache_strict(ptb_ctx* c);
void ptb_shade_prepare_cache_fast(ptb_ctx* c);
void ptb_shade_launch_strict(ptb_ctx* c, const Lane& L, int engine, const float* rngtab, int dim, int rng_stride, const FrameMap& fm, int cur, cudaStream_t st);
void ptb_shade_launch_fast(ptb_ctx* c, const Lane& L, int engine, const float* rngtab, int dim, int rng_stride, const FrameMap& fm, int cur, cudaStream_t st);
int ptb_wf_shade_tap_strict(ptb_ctx* c, int what, const float* in0_dev, const float* in1_dev, const int32_t* ini_dev, int m, float* out_dev);
int ptb_wf_shade_tap_fast(ptb_ctx* c, int what, const float* in0_dev, const float* in1_dev, const int32_t* ini_dev, int m, float* out_dev);
inline int ptb_wf_shade_tap(ptb_ctx* c, int what, const float* in0_dev, const float* in1_dev, const int32_t* ini_dev, int m, float* out_dev) {
    return c->fast_shade ? ptb_wf_shade_tap_fast(c, what, in0_dev, in1_dev, ini_dev, m, out_dev) : ptb_wf_shade_tap_strict(c, what, in0_dev, in1_dev, ini_dev, m, out_dev);
}
inline void ptb_shade_prepare_cache(ptb_ctx* c) { c->fast_shade ? ptb_shade_prepare_cache_fast(c) : ptb_shade_prepare_cache_strict(c); }
inline void ptb_shade_launch(ptb_ctx* c, const Lane& L, int engine, const float* rngtab, int dim, int rng_stride, const FrameMap& fm, int cur, cudaStream_t st) {
    c->fast_shade ? ptb_shade_launch_fast(c, L, engine, rngtab, dim, rng_stride, fm, cur, st) : ptb_shade_launch_strict(c, L, engine, rngtab, dim, rng_stride, fm, cur, st);
}
Lane ptb_lane(ptb_ctx* c, int which, int nlanes);        // lane `which` of `nlanes` equal shares of the pools
int ptb_wf_measure_l2(ptb_ctx* c, int mbytes, int iters, float* gbps);
int ptb_wf_selftest(ptb_ctx* c, int what, long long n, unsigned long long seed, long long* fails);
int ptb_wf_mlt_reset(ptb_ctx* c);
void ptb_stage_begin(ptb_ctx* c, int stage);
void ptb_stage_end(ptb_ctx* c);
int ptb_stage_collect(ptb_ctx* c);
