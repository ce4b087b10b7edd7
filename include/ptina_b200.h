/* ptina_b200.h -- C ABI of libptina_b200.so (sm_100a CUDA implementation of PTina's per-pixel
 * path-tracing hot path).
 *
 * The reference (archibate/ptina) has no FFI: its boundary is a Python module of singletons
 * (ptina/worker.py:11-87, ptina/things.py:12-28).  Each entry point below names the reference
 * interface it replaces (paths relative to the reference's ptina/ package); the Python facade
 * ptina_b200/ binds them with ctypes and re-exposes the reference's own names.  INTEGRATION.md
 * shows the stub a PTina maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; ptb_last_error() gives the message
 *     (thread-local).  Reference error texts are kept ("too many faces", "AABB step never stop! ...").
 *   - plain pointers and sizes only.  A pointer argument followed by `memspace` may be a host
 *     (PTB_HOST) or device (PTB_DEVICE) pointer; all other pointers are HOST pointers.
 *   - one context = one GPU = the reference's one set of singletons.  Single-threaded,
 *     non-reentrant use per context, exactly like the reference (tools/mtworker.py:39-42).
 *   - all device work is enqueued on the context's stream (ptb_set_stream; default = the legacy
 *     default stream).  Only readbacks (memspace == PTB_HOST outputs), ptb_build_tree and
 *     ptb_synchronize block the host.
 *   - ptb_render only RECORDS its samples; consecutive calls (the reference's one-sample-per-call loop,
 *     exams/benchmark.py:29-33) are submitted as one wavefront batch when the batch is full or when any
 *     other entry point that could observe the difference is called (every entry point below flushes
 *     first; ptb_flush does only that).  The film is bit-identical to call-by-call submission.
 *   - there is NO CPU fallback: without a CUDA device ptb_create fails.
 */
#ifndef PTINA_B200_H
#define PTINA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_ctx ptb_ctx;

enum { PTB_HOST = 0, PTB_DEVICE = 1 };
enum { PTB_ENGINE_PATH = 0, PTB_ENGINE_BRUTE = 1, PTB_ENGINE_PREVIEW = 2, PTB_ENGINE_MLT = 3 };
enum { PTB_LIGHT_POINT = 1, PTB_LIGHT_AREA = 2 };            /* light/__init__.py:11 */
enum { PTB_TRAVERSE_AUTO = 0,      /* ordered + culled when the tree validates, else reference order */
       PTB_TRAVERSE_REFERENCE = 1, /* lbvh.py:313-347 literally: unordered child1-first DFS, no culling */
       PTB_TRAVERSE_ORDERED = 2,   /* near-first, distance-culled; conservative box tests + exact gate test on acceptance */
       PTB_TRAVERSE_ORDERED_EXACT = 3 /* the same order with the exact (division-equivalent) slab test at every node */ };

/* PTB_MODE_PARITY (default): every stage in strict IEEE binary32 -- bit-exact trees / rays / hits, BSDF eval / pdf within 1e-5 of the
 * reference.  PTB_MODE_FAST (opt-in, the role tree/middlebvh.py's "fast, non-parity" path plays in the reference): the shading stage
 * runs its FMA-contracted, approximate-division build; trees, rays and hit ids are unchanged, shaded values agree to ~1e-4 per tap and
 * to the 1e-3 image gate. */
enum { PTB_MODE_PARITY = 0, PTB_MODE_FAST = 1 };

/* things.py:12-19 init_things(...) capacities.  Zero fields take the reference defaults. */
typedef struct ptb_caps {
    int32_t max_faces;      /* 2^21 */
    int32_t max_texels;     /* 2^22 */
    int32_t max_materials;  /* 2^6  */
    int32_t max_textures;   /* 2^6  */
    int32_t max_lights;     /* 2^6  */
    int32_t max_filmsize;   /* 2^21 */
    int32_t max_filmpasses; /* 3    */
    int64_t max_paths;      /* wavefront pool size (paths in flight); 0 = 2^23 */
} ptb_caps;

typedef struct ptb_tree_info {
    int32_t n;            /* faces */
    int32_t aabb_sweeps;  /* level-synchronous sweeps used (lbvh.py:251-261 prints this as "depth") */
    int32_t valid;        /* 1 = proper binary tree (every node one parent, leaf slots in order) */
    int32_t depth;        /* height of the tree in nodes, root = 1 (0 if !valid) */
    int32_t policy;       /* traversal policy in effect (PTB_TRAVERSE_REFERENCE / _ORDERED) */
    float build_ms;       /* device time of the whole build */
    int32_t list_n;       /* triangles on the always-test list of the production traversal (big ones; at most 32, the rest stay in the tree) */
    int32_t trav_depth;   /* height of the tree the production traversal walks (PLOC topology when trav_ploc, else the LBVH's = depth) */
    int32_t trav_ploc;    /* 1 = the traversal tree was rebuilt by PLOC (same answers by the gate lemma, fewer node visits) */
} ptb_tree_info;

typedef struct ptb_counters {
    int64_t rays;         /* extend + shadow rays traced */
    int64_t extend_rays;
    int64_t shadow_rays;
    int64_t node_visits;  /* packed internal nodes fetched (64 B each) */
    int64_t box_tests;    /* child-box slab tests */
    int64_t tri_tests;    /* triangle tests (64 B each) */
    int64_t paths;        /* camera paths started */
    int64_t max_stack;
} ptb_counters;

const char* ptb_last_error(void);
int ptb_version(void);

/* things.py:12-28 init_things + worker.py:11-14 init */
int ptb_create(int device, const ptb_caps* caps, ptb_ctx** out);
int ptb_destroy(ptb_ctx* ctx);
int ptb_set_stream(ptb_ctx* ctx, void* cuda_stream);
int ptb_set_mode(ptb_ctx* ctx, int mode);                           /* PTB_MODE_*; PTB_FAST=1 in the environment makes FAST the default */
int ptb_get_mode(ptb_ctx* ctx, int* mode);
/* scheduling / builder switches that do NOT change results (tests and measurements flip them at run time; the environment variables of
 * INTEGRATION.md section 3 set their defaults): "coalesce", "overlap_shadow", "pt_lanes" and "mlt_lanes" (1..4), "use_ploc", "ploc_big",
 * "ploc_radius", "wide4" (4-wide nodes for trees walked out of global memory).  The tree options take effect at the next ptb_build_tree. */
int ptb_set_option(ptb_ctx* ctx, const char* name, int value);
int ptb_synchronize(ptb_ctx* ctx);                                  /* worker.py:17-18 */
int ptb_flush(ptb_ctx* ctx);                                        /* submit recorded ptb_render calls without waiting */

/* sampling/sobol.py:75-97: V = calc_sobol_vgrid(2^20, 21201) as the i32 field [rows=21][dim]; resets
 * the generator (time = 0 then `skip` = 64 updates). */
int ptb_set_sobol_table(ptb_ctx* ctx, const int32_t* V, int rows, int dim);
int ptb_sobol_reset(ptb_ctx* ctx);                                  /* sobol.py:92-97 */
int ptb_sobol_get_time(ptb_ctx* ctx, int* time);                    /* number of update() calls so far */
int ptb_sobol_set_time(ptb_ctx* ctx, int time);
int ptb_sobol_point(ptb_ctx* ctx, int k, float* P_out);             /* tap: P after k updates, [dim] */

/* model.py:62-86 ModelPool.load -> from_numpy.  verts [nfaces*3][8] = px py pz nx ny nz u v. */
int ptb_load_model(ptb_ctx* ctx, const float* verts, const int32_t* mtlids, int nfaces, int memspace);
/* mtllib.py:58-77 MaterialPool.load after ParameterPair.load: fac [nmat][12][4], tex [nmat][12] */
int ptb_load_materials(ptb_ctx* ctx, const float* fac, const int32_t* tex, int nmat);
/* image.py:69-95 ImagePool.load after the host allocator: arena texels [ntexels][4], per-image nx/ny/base */
int ptb_load_images(ptb_ctx* ctx, const float* texels, int64_t ntexels, const int32_t* nx, const int32_t* ny,
                    const int32_t* base, int nimg, int memspace);
/* light/__init__.py:31-49 LightPool.clear / add (pos = world@(0,0,0,1), axes = world[:3,:3] row-major) */
int ptb_clear_lights(ptb_ctx* ctx);
int ptb_add_light(ptb_ctx* ctx, const float pos[3], const float axes[9], const float color[3], float size, int type);
int ptb_set_world_light(ptb_ctx* ctx, const float fac[4], int tex);  /* light/world.py:18-20 */
/* camera.py:19-22: v2w = float32(inv(pers)), w2v = float32(pers), row-major 4x4 */
int ptb_set_camera(ptb_ctx* ctx, const float v2w[16], const float w2v[16]);

/* tree/lbvh.py:297-305 BVHTree().build(): Morton codes -> radix sort -> hierarchy -> AABBs (+ validation
 * and packing for traversal).  Fails with "AABB step never stop! hierarchy corrupted?" like lbvh.py:258-259. */
int ptb_build_tree(ptb_ctx* ctx, ptb_tree_info* info);
int ptb_set_traversal(ptb_ctx* ctx, int policy);
/* tap: the reference's own arrays (lbvh.py:50-59); any pointer may be NULL.  mc,id,leaf [n]; child [n-1][2];
 * bmin,bmax [n-1][3] */
int ptb_export_tree(ptb_ctx* ctx, int32_t* mc, int32_t* id, int32_t* child, int32_t* leaf, float* bmin, float* bmax);
/* tap: the production traversal structure (no counterpart in the reference; tests validate it).  nodes [n-1][16]: per internal
 * node the boxes of its two children and their ids -- (lo0.xyz, id0) (hi0.xyz, id1) (lo1.xyz, -) (hi1.xyz, -); id: -1 nothing
 * below, bit 30 = exempt from distance culling, low bits < n = leaf slot, otherwise n + internal index; root = node 0; unused
 * entries have both ids -1.  leaf_lo/leaf_hi [n][4]: inflated bounds per leaf slot (lo.w = flags: 1 never hit, 2 ill-conditioned,
 * 4 big, 8 on the always-test list).  gbox [2n][4]: reference box of each slot's gate.  ploc: 1 = PLOC topology, 0 = LBVH's. */
int ptb_export_traversal(ptb_ctx* ctx, float* nodes, float* leaf_lo, float* leaf_hi, float* gbox, int32_t* ploc);

/* filmtable.py:41-45 */
int ptb_set_size(ptb_ctx* ctx, int nx, int ny);
int ptb_get_size(ptb_ctx* ctx, int* nx, int* ny);
int ptb_clear(ptb_ctx* ctx);                                         /* zeroes every pass, filmtable.py:44-45 */
/* device address of film pass `pass` ([nx*ny][4] running sums, index x*ny+y) -- for the multi-GPU reduce */
int ptb_film_ptr(ptb_ctx* ctx, int pass, void** dev_ptr, int64_t* ntexels);

/* engine/path.py:75-77, brute.py:25-27, preview.py:19-21, mltpath.py:85-87: `nsamples` consecutive
 * Engine.render() calls (one Sobol update + one sample per pixel each). */
int ptb_render(ptb_ctx* ctx, int engine, int nsamples);
/* engine/path.py:96-118 render_tile(i, j, samples): one Sobol update, then samples m = 0..min(samples, 63) (inclusive, as the
 * reference text has it) of every pixel of the 64x64 tile (i, j), sample m rotated by wanghash3(x, y, m); pixels outside the film
 * are skipped.  engine/path.py:120-128 render_final(nsamples): every tile, render_tile(i, j, s) for s = nsamples, nsamples-64, ... > 0.
 * (Both are commented out in the reference's engine/path.py and driven by exams/benchtiles.py:25-31.)  PATH or BRUTE engine. */
int ptb_render_tile(ptb_ctx* ctx, int engine, int i, int j, int samples);
int ptb_render_final(ptb_ctx* ctx, int engine, int nsamples);
/* the same for explicit Sobol point indices k_first, k_first+stride, ... (count of them); does not touch
 * the generator's time.  Used to shard a sample range over GPUs. */
int ptb_render_range(ptb_ctx* ctx, int engine, int k_first, int count, int stride);
/* mltpath.py:30-36 reset(); LSP / Sigma fields mltpath.py:23-33 */
int ptb_mlt_reset(ptb_ctx* ctx, uint64_t seed, int chain_first, int chain_count);
int ptb_mlt_set_param(ptb_ctx* ctx, float lsp, float sigma);
/* tap: the chains' last proposal X_new [count][32] and its radiance L_new [count][3] (mltpath.py:56-75), and the current
 * state X_old / L_old; any pointer may be NULL.  `count` = the chain count the buffers are sized for (must match ptb_mlt_reset's) */
int ptb_mlt_state(ptb_ctx* ctx, int count, float* x_new, float* l_new, float* x_old, float* l_old);

/* filmtable.py:47-63 get_image: out [nx][ny][4];  filmtable.py:65-79 fast_export_image: out [ny*nx*3] */
int ptb_get_image(ptb_ctx* ctx, int pass, float* out, int memspace);
int ptb_fast_export_image(ptb_ctx* ctx, int pass, float* out, int memspace);
int ptb_get_film(ptb_ctx* ctx, int pass, float* out, int memspace);  /* raw sums [nx*ny][4] */
/* blender.py:809-820 (TinaDrawData: fast_export_image into a NumPy array that is then uploaded as a GL buffer): the same resolve
 * written straight into an OpenGL buffer object of the caller's current GL context (>= nx*ny*3 floats) through CUDA-GL interop --
 * no host round trip.  Fails with a message when the calling thread has no GL context. */
int ptb_fast_export_gl(ptb_ctx* ctx, int pass, unsigned int gl_buffer);

/* ---- parity taps (tests) ----------------------------------------------------------------------- */
/* primary rays of Sobol point k for every pixel (path.py:85-93 + camera.py:34-39) and their closest hits
 * (lbvh.py:313-347).  rays [nx*ny][6] as generated; hit/index [nx*ny]; depth [nx*ny]; uv [nx*ny][2].
 * Any output may be NULL. */
int ptb_trace_primary(ptb_ctx* ctx, int k, float* rays, int32_t* hit, float* depth, int32_t* index, float* uv);
/* closest hit for arbitrary rays [m][6] (not re-normalised) with per-ray avoid ids (or NULL) */
int ptb_intersect(ptb_ctx* ctx, const float* rays, const int32_t* avoid, int m, int policy,
                  int32_t* hit, float* depth, int32_t* index, float* uv);
/* any-hit (shadow) query: occluded[i] = 1 iff some triangle != avoid is hit with depth <= dis[i] */
int ptb_occluded(ptb_ctx* ctx, const float* rays, const int32_t* avoid, const float* dis, int m, int policy,
                 int32_t* occluded);
/* materials/disney.py:52-106 / 114-233.  params [m][14], geom [m][10] = normal, sign, wi, wo|samp.
 * eval -> out [m][3];  sample -> out [m][7] = outdir, pdf, color */
int ptb_eval_bsdf(ptb_ctx* ctx, const float* params, const float* geom, int m, float* out);
int ptb_sample_bsdf(ptb_ctx* ctx, const float* params, const float* geom, int m, float* out);
/* the same two with every term of disney.py evaluated as written -- the production forms above leave out factors that an exactly-zero
 * material weight (transmission, clearcoat, subsurface) annihilates; the tests compare the two bit for bit */
int ptb_eval_bsdf_literal(ptb_ctx* ctx, const float* params, const float* geom, int m, float* out);
int ptb_sample_bsdf_literal(ptb_ctx* ctx, const float* params, const float* geom, int m, float* out);
int ptb_material_get(ptb_ctx* ctx, const int32_t* mtlid, const float* uv, int m, float* out14);   /* mtllib.py:79-95 */
int ptb_light_hit(ptb_ctx* ctx, const float* rays, int m, float* out6);                            /* light/__init__.py:51-81 */
int ptb_light_sample(ptb_ctx* ctx, const float* hitpos_samp, int m, float* out8);                  /* light/__init__.py:83-121 */
int ptb_world_at(ptb_ctx* ctx, const float* dirs, int m, float* out3);                             /* light/world.py:22-29 */
int ptb_normaldist(ptb_ctx* ctx, const float* samp, int m, float* out);                            /* common.py:337-352 erfinv + normaldist (MLT small steps) */
/* per-pixel radiance of one sample without touching the film: out [nx*ny][3] */
int ptb_render_sample(ptb_ctx* ctx, int engine, int k, float* out);

/* bit 0: count rays / node visits / triangle tests (separate kernel instantiation);
 * bit 1: time every stage with CUDA events on the context's stream */
int ptb_set_counting(ptb_ctx* ctx, int flags);
int ptb_get_counters(ptb_ctx* ctx, ptb_counters* out);
int ptb_reset_counters(ptb_ctx* ctx);
/* device time (ms) spent in each stage since the last reset: raygen, extend, shade, shadow, accumulate */
int ptb_get_stage_ms(ptb_ctx* ctx, float ms[5]);
/* kernels launched by this context since the last ptb_reset_counters */
int ptb_get_launches(ptb_ctx* ctx, int64_t* n);
/* device self-test of the exact-division shortcut used by the traversal: compares div_exact(a, d, RN(1/d)) with the
 * IEEE quotient a / d bit for bit on n random operand pairs (what = 0: slab-test ranges, 1: wide ranges); *fails = mismatches */
int ptb_selftest(ptb_ctx* ctx, int what, int64_t n, uint64_t seed, int64_t* fails);
/* measured L2-resident read bandwidth (GB/s): streams a `mbytes` buffer `iters` times with 16-B loads */
int ptb_measure_l2(ptb_ctx* ctx, int mbytes, int iters, float* gbps);

#ifdef __cplusplus
}
#endif
#endif
