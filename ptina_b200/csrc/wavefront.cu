// wavefront.cu -- the per-pixel path-tracing hot path as a wavefront pipeline.
//
// The reference runs one divergent megakernel per sample (engine/path.py:79-83 `_render` -> do_render path.py:85-93 ->
// path_trace path.py:18-64; engine/brute.py:29-75; engine/mltpath.py:54-83).  Here the same per-path arithmetic is
// split into stages that each keep warps on one code path, with path state in HBM (PathState, SoA) and
// compacted index queues between stages:
//
//   sobol_points -> raygen -> [ extend -> shade -> shadow ] x 5 bounces -> accumulate
//
//   raygen   : Sobol jitter (sobol.py:107-125), Camera.generate (camera.py:34-39)
//   extend   : closest hit (lbvh.py:313-347), persistent warps pulling 32-ray chunks from the queue
//   shade    : light hit + MIS, miss -> environment, shading frame, material fetch, NEE sample -> shadow queue,
//              Disney bounce -> next active queue (queues compacted with warp ballot + block prefix)
//   shadow   : any-hit occlusion, adds the pending NEE contribution
//   accumulate: film[0, x, y] += (rgb, 1) per sample in Sobol order (path.py:93) -- deterministic, no atomics
//
// Random numbers: draw c of pixel (i,j) at Sobol point k is P_k[(wanghash2(i,j) + c) mod 21201]; the draw index is a
// pure function of the bounce number (2 jitter draws, then 3 NEE + 3 bounce draws per bounce), so no per-path
// generator state is carried.
#include <cstdio>
#include <cstdlib>
#include "ptb_wavefront.cuh"
#include "ptb_trace_kernel.cuh"

namespace {

constexpr int BLK = 128;

// ---- sampling/sobol.py:99-105 in closed form: X_k = XOR_{b in gray(k)} V[b+1];  P = X / 2^32 (sobol.py:19-29) ---------
__global__ void k_sobol_points(const int* __restrict__ V, int dim, int k_first, int count, int stride, float* __restrict__ P) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    for (int s = 0; s < count; s++) {
        unsigned k = (unsigned)(k_first + s * stride);
        unsigned g = k ^ (k >> 1), X = 0;
        for (int b = 0; b < 20; b++) if ((g >> b) & 1u) X ^= (unsigned)V[(b + 1) * dim + j];
        // only the top 20 bits are ever set, so the MSB-first sum of construct_float is this exact product
        P[(size_t)s * dim + j] = (float)X * 2.3283064365386963e-10f;
    }
}

__global__ void k_ctrl_begin(Ctrl* c, int n_in) { c->n_in = n_in; c->n_out = 0; c->n_shadow = 0; c->cur_extend = 0; c->cur_shadow = 0; c->n_tree[0] = c->n_tree[1] = 0; c->cur_tree[0] = c->cur_tree[1] = 0; }
__global__ void k_ctrl_tap(Ctrl* c, int m) { c->pad[0] = m; c->pad[1] = 0; c->n_tree[1] = 0; c->cur_tree[1] = 0; }
// overlapped schedule: the shadow stage of the bounce still reads its own fields while the next extend stage is set up
__global__ void k_ctrl_next_extend(Ctrl* c) { c->n_in = c->n_out; c->n_out = 0; c->cur_extend = 0; c->n_tree[0] = 0; c->cur_tree[0] = 0; }
__global__ void k_ctrl_reset_shadow(Ctrl* c) { c->n_shadow = 0; c->cur_shadow = 0; c->n_tree[1] = 0; c->cur_tree[1] = 0; }
__global__ void k_ctrl_next(Ctrl* c) { c->n_in = c->n_out; c->n_out = 0; c->n_shadow = 0; c->cur_extend = 0; c->cur_shadow = 0; c->n_tree[0] = c->n_tree[1] = 0; c->cur_tree[0] = c->cur_tree[1] = 0; }

// ---- path.py:85-93 do_render (first half) / brute.py:66-73 -----------------------------------------------------------
// renorm: path_trace normalises the (already unit) camera direction again before intersecting (path.py:28); PreviewEngine does not
__global__ void __launch_bounds__(BLK) k_raygen(const SceneParams* __restrict__ P, const float* __restrict__ sobolP, int dim, FrameMap fm, int nsamp, int renorm,
                                                PathState st, RayQueue xq, Ctrl* ctrl, DevCounters* ctr) {
    __shared__ int s_warp[BLK / 32]; __shared__ int s_base;
    int total = fm.pps * nsamp;
    int rounded = (total + BLK - 1) / BLK * BLK;
    for (int p0 = blockIdx.x * BLK; p0 < rounded; p0 += gridDim.x * BLK) {
        int p = p0 + threadIdx.x;
        bool live = false;
        V3 ro = v3s(0.0f), rd = v3s(0.0f);
        if (p < total) {
            int s = (int)fastdiv((unsigned)p, fm.d_pps), q = p - s * fm.pps;
            int x, y;
            if (slot_pixel(fm, q, &x, &y)) {
                live = true;
                Rng rng;
                frame_rng(fm, sobolP, dim, s, x, y, &rng);
                float dx = rng.draw(0), dy = rng.draw(1);
                float fx = ((float)x + dx) / (float)fm.nx * 2.0f - 1.0f;
                float fy = ((float)y + dy) / (float)fm.ny * 2.0f - 1.0f;
                camera_generate(P, fx, fy, &ro, &rd);
                const int slot = fm.slot_base + p;
                st.ray_o[slot] = make_float4(ro.x, ro.y, ro.z, 0.0f);
                st.ray_d[slot] = make_float4(rd.x, rd.y, rd.z, 0.0f);
                st.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));           // depth = 0
                st.result[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);                      // last_brdf_pdf = 0
            }
        }
        int pos = block_append<BLK>(live, &ctrl->n_in, s_warp, &s_base);
        if (live) {
            if (renorm) rd = normalized(rd);
            xq.o[pos] = make_float4(ro.x, ro.y, ro.z, __int_as_float(fm.slot_base + p));
            xq.d[pos] = make_float4(rd.x, rd.y, rd.z, __int_as_float(-1));                 // avoid = -1
        }
    }
    if (ctr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr->paths, (unsigned long long)nsamp * fm.nx * fm.ny);
}

// ---- preview.py:23-41: primary hit -> albedo into pass 1, shading normal into pass 2 ---------------------------------------------
__global__ void __launch_bounds__(BLK) k_preview(const SceneParams* __restrict__ P, const float4* __restrict__ texels, const float* __restrict__ verts,
                                                 const int* __restrict__ mtlids, FrameMap fm, int nsamp, PathState st, float4* __restrict__ film1, float4* __restrict__ film2) {
    int q = blockIdx.x * BLK + threadIdx.x;
    if (q >= fm.pps) return;
    int x, y;
    if (!slot_pixel(fm, q, &x, &y)) return;
    size_t pix = (size_t)x * fm.ny + y;
    float4 a = film1[pix], b = film2[pix];
    for (int s = 0; s < nsamp; s++) {
        int p = fm.slot_base + s * fm.pps + q;
        float4 o4 = st.ray_o[p], d4 = st.ray_d[p], h4 = st.hit[p];
        V3 albedo = v3s(0.0f), normal = v3s(0.0f);
        int hit_index = __float_as_int(h4.w);
        if (hit_index >= 0) {
            V3 hitpos; float tu, tv; int mtlid;
            shading_frame(verts, mtlids, hit_index, h4.y, h4.z, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), h4.x, &hitpos, &normal, &tu, &tv, &mtlid);
            albedo = material_get(P, texels, mtlid, tu, tv).basecolor;
        }
        a = make_float4(a.x + albedo.x, a.y + albedo.y, a.z + albedo.z, a.w + 1.0f);
        b = make_float4(b.x + normal.x, b.y + normal.y, b.z + normal.z, b.w + 1.0f);
    }
    film1[pix] = a; film2[pix] = b;
}

// ---- path.py:93: FilmTable()[0, i, j] += V34(clr, 1.0), samples folded in Sobol order ------------------------------------------------
__global__ void __launch_bounds__(BLK) k_accumulate(float4* __restrict__ film, const float4* __restrict__ result, FrameMap fm, int nsamp, float* __restrict__ sample_out) {
    int q = blockIdx.x * BLK + threadIdx.x;
    if (q >= fm.pps) return;
    int x, y;
    if (!slot_pixel(fm, q, &x, &y)) return;
    size_t pix = (size_t)x * fm.ny + y;
    if (sample_out) {   // tap: radiance of the first sample, film untouched
        float4 r = result[fm.slot_base + q];
        sample_out[3 * pix] = r.x; sample_out[3 * pix + 1] = r.y; sample_out[3 * pix + 2] = r.z;
        return;
    }
    float4 f = film[pix];
    for (int s = 0; s < nsamp; s++) {
        float4 r = result[fm.slot_base + (size_t)s * fm.pps + q];
        f = make_float4(f.x + r.x, f.y + r.y, f.z + r.z, f.w + 1.0f);
    }
    film[pix] = f;
}

// ---- filmtable.py:47-79 resolve -----------------------------------------------------------------------------------------------------------
// mode 0: get_image -> out[x][y][4];  mode 1: fast_export_image -> out[(y*nx+x)*3];  mode 2: raw copy
__global__ void __launch_bounds__(BLK) k_resolve(const float4* __restrict__ film, int nx, int ny, int mode, float* __restrict__ out) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= nx * ny) return;
    float4 v = film[i];
    if (mode == 2) { reinterpret_cast<float4*>(out)[i] = v; return; }
    int x = i / ny, y = i - x * ny;
    if (mode == 0) {
        if (v.w != 0.0f) v = make_float4(v.x / v.w, v.y / v.w, v.z / v.w, 1.0f);
        else v = make_float4(0.9f, 0.4f, 0.9f, 0.0f);
        reinterpret_cast<float4*>(out)[i] = v;
    } else {
        if (v.w != 0.0f) { v.x /= v.w; v.y /= v.w; v.z /= v.w; }
        else { v.x = 0.9f; v.y = 0.4f; v.z = 0.9f; }
        size_t base = ((size_t)y * nx + x) * 3;
        out[base] = v.x; out[base + 1] = v.y; out[base + 2] = v.z;
    }
}

// ---- taps ----------------------------------------------------------------------------------------------------------------------------------
// flat ray arrays of ptb_intersect / ptb_occluded -> queue records (directions are used as given, like LinearBVH.intersect)
__global__ void __launch_bounds__(BLK) k_pack_tap(const float* __restrict__ rays, const int* __restrict__ avoid, const float* __restrict__ dis,
                                                  const int* __restrict__ slot_of, int nfaces, int m, RayQueue q) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m) return;
    int av = avoid ? avoid[i] : -1;
    int slot = (av >= 0 && av < nfaces) ? slot_of[av] : -1;
    q.o[i] = make_float4(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2], __int_as_float(i));
    q.d[i] = make_float4(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5], dis ? dis[i] : __int_as_float(slot));
    if (dis) q.c[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(slot));
}
__global__ void __launch_bounds__(BLK) k_gather_primary(FrameMap fm, PathState st, float* rays, int* hit, float* depth, int* index, float* uv, int which) {
    int q = blockIdx.x * BLK + threadIdx.x;
    if (q >= fm.pps) return;
    int x, y;
    if (!slot_pixel(fm, q, &x, &y)) return;
    size_t pix = (size_t)x * fm.ny + y;
    if (which == 0) {
        float4 o = st.ray_o[fm.slot_base + q], d = st.ray_d[fm.slot_base + q];
        rays[6 * pix] = o.x; rays[6 * pix + 1] = o.y; rays[6 * pix + 2] = o.z; rays[6 * pix + 3] = d.x; rays[6 * pix + 4] = d.y; rays[6 * pix + 5] = d.z;
    } else {
        float4 h = st.hit[fm.slot_base + q];
        int id = __float_as_int(h.w);
        if (hit) hit[pix] = id >= 0;
        if (depth) depth[pix] = h.x;
        if (index) index[pix] = id;
        if (uv) { uv[2 * pix] = h.y; uv[2 * pix + 1] = h.z; }
    }
}

// L2-resident read bandwidth probe: every block streams the whole buffer with 16-byte loads
__global__ void __launch_bounds__(256) k_l2_read(const float4* __restrict__ buf, size_t n4, int iters, float* sink) {
    float acc = 0.0f;
    for (int it = 0; it < iters; it++) {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
            float4 v = __ldcg(&buf[i]);
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123456.789f) *sink = acc;
}

// ---- MLT (engine/mltpath.py:9-87) -----------------------------------------------------------------------------------------------------------
// ti.random() is Taichi's per-thread xorshift stream, which cannot be reproduced; a counter-based Philox4x32-10 keyed by
// (seed, chain, iteration) replaces it (statistical parity only).
struct Philox {
    uint32_t key[2], ctr[4], out[4]; int have;
    PTB_D void init(uint64_t seed, uint32_t chain, uint32_t iter) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); ctr[0] = 0; ctr[1] = chain; ctr[2] = iter; ctr[3] = 0x5054494eu; have = 0;
    }
    PTB_D void round_(uint32_t* c, uint32_t* k) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    PTB_D float next() {
        if (have == 0) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
            for (int r = 0; r < 10; r++) { round_(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
            out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
            ctr[0]++; have = 4;
        }
        uint32_t v = out[--have];
        return (float)(v >> 8) * 5.9604644775390625e-8f;   // [0, 1)
    }
};
// common.py:337-352 erfinv (Winitzki, a = 0.147) and normaldist
PTB_D float erfinv_ref(float x) {
    float sgn = x < 0.0f ? -1.0f : 1.0f;
    x = (1.0f - x) * (1.0f + x);
    float lnx = logf(x);
    float tt1 = (float)(2.0 / (3.14159265358979323846 * 0.147)) + 0.5f * lnx;
    float tt2 = (float)(1.0 / 0.147) * lnx;
    return sgn * sqrtf(-tt1 + sqrtf(tt1 * tt1 - tt2));
}
PTB_D float normaldist(float samp) { return 1.41421356237309504880f * erfinv_ref(samp * 2.0f - 1.0f); }
PTB_D float pymodf1(float a) { float r = fmodf(a, 1.0f); if (r < 0.0f) r += 1.0f; return r; }   // Python-style `% 1`

// mltpath.py:30-36 reset
__global__ void __launch_bounds__(BLK) k_mlt_reset(float* Xold, float4* Lold, int nchains, int chain_first, uint64_t seed) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= nchains) return;
    Philox g; g.init(seed, (uint32_t)(chain_first + i), 0xffffffffu);
    Lold[i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);   // L_old = 0, accum = 1
    for (int j = 0; j < 32; j++) Xold[(size_t)i * 32 + j] = g.next();
}
// mltpath.py:56-74: mutate, then camera ray from the first two dims
// chains [first, first + nchains) of this process's population (one lane's share); path slot = chain index, queue position = local index
__global__ void __launch_bounds__(BLK) k_mlt_raygen(const SceneParams* __restrict__ P, const float* __restrict__ Xold, float* __restrict__ Xnew, int first, int nchains, int chain_first,
                                                    uint64_t seed, uint32_t iter, float lsp, float sigma, PathState st, RayQueue xq, Ctrl* ctrl, DevCounters* ctr) {
    const int q = blockIdx.x * BLK + threadIdx.x;
    if (q >= nchains) return;
    const int i = first + q;
    Philox g; g.init(seed, (uint32_t)(chain_first + i), iter);
    const float* xo = Xold + (size_t)i * 32; float* xn = Xnew + (size_t)i * 32;
    if (g.next() < lsp) { for (int j = 0; j < 32; j++) xn[j] = g.next(); }
    else { for (int j = 0; j < 32; j++) xn[j] = pymodf1(xo[j] + sigma * normaldist(g.next())); }
    V3 ro, rd;
    camera_generate(P, xn[0] * 2.0f - 1.0f, xn[1] * 2.0f - 1.0f, &ro, &rd);
    st.thr[i] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(0));
    st.result[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    rd = normalized(rd);                                                                // path.py:28
    xq.o[q] = make_float4(ro.x, ro.y, ro.z, __int_as_float(i));
    xq.d[q] = make_float4(rd.x, rd.y, rd.z, __int_as_float(-1));
    if (q == 0) { ctrl->n_in = nchains; if (ctr) atomicAdd(&ctr->paths, (unsigned long long)nchains); }
}
// mltpath.py:47-52 splat + 75-81 accept/reject
__global__ void __launch_bounds__(BLK) k_mlt_finish(float* __restrict__ Xold, const float* __restrict__ Xnew, float4* __restrict__ Lold, const float4* __restrict__ result,
                                                    int first, int nchains, int chain_first, uint64_t seed, uint32_t iter, int nx, int ny, float4* __restrict__ film) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= nchains) return;
    i += first;
    float4 Ln4 = result[i], Lo4 = Lold[i];
    V3 Ln = mk3(Ln4.x, Ln4.y, Ln4.z), Lo = mk3(Lo4.x, Lo4.y, Lo4.z);
    float AL_new = vavg(Ln) + 1e-10f, AL_old = vavg(Lo) + 1e-10f;
    float accept = fminf(1.0f, AL_new / AL_old);
    V3 splat = Ln * Lo4.w;                                    // L_new * accum
    const float* xn = Xnew + (size_t)i * 32;
    int px = ifloor(xn[0] * (float)nx), py = ifloor(xn[1] * (float)ny);
    if (px >= 0 && px < nx && py >= 0 && py < ny) {
        float* f = reinterpret_cast<float*>(&film[(size_t)px * ny + py]);
        atomicAdd(f, splat.x); atomicAdd(f + 1, splat.y); atomicAdd(f + 2, splat.z); atomicAdd(f + 3, 1.0f);
    }
    Philox g; g.init(seed ^ 0xA5A5A5A5DEADBEEFull, (uint32_t)(chain_first + i), iter);
    if (g.next() < accept) {
        Lold[i] = make_float4(Ln.x, Ln.y, Ln.z, Lo4.w);
        for (int j = 0; j < 32; j++) Xold[(size_t)i * 32 + j] = xn[j];
    }
}

// div_exact(a, d, RN(1/d)) == a / d ?  Operands drawn from a counter-based hash; what = 0: |d| in [1e-6, 1] and |a| up to
// 1e7 (the slab test's ranges), what = 1: both over [1e-18, 1e18].
PTB_D float rand_float(uint32_t h, float lo_exp, float hi_exp) {
    float u = (float)(h >> 9) * (1.0f / 8388608.0f);                  // [0,1)
    float e = lo_exp + (hi_exp - lo_exp) * u;
    float m = 1.0f + (float)(wanghash((int)h) & 0x7FFFFF) * (1.0f / 8388608.0f);
    float v = exp2f(floorf(e)) * m;
    return (h & 1u) ? -v : v;
}
__global__ void __launch_bounds__(256) k_selftest_div(int what, long long n, unsigned long long seed, unsigned long long* fails) {
    unsigned long long bad = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        uint32_t h1 = (uint32_t)wanghash((int)(i * 2654435761u + seed)), h2 = (uint32_t)wanghash((int)(h1 ^ (uint32_t)(i >> 32) ^ 0x9E3779B9u));
        float d = what == 0 ? rand_float(h1, -20.0f, 0.0f) : rand_float(h1, -60.0f, 60.0f);
        float a = what == 0 ? rand_float(h2, -40.0f, 24.0f) : rand_float(h2, -60.0f, 60.0f);
        if ((h2 & 0xFF0u) == 0) a = 0.0f;                             // exact zeros are common (origin on a box plane)
        float q = a / d;
        float r = 1.0f / d;
        float f = div_exact(a, d, r);
        if (__float_as_int(q) != __float_as_int(f) && fabsf(q) > 1e-30f) bad++;
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(fails, bad);
}

__global__ void __launch_bounds__(BLK) k_normaldist_tap(const float* __restrict__ in, int m, float* __restrict__ out) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i < m) out[i] = normaldist(in[i]);
}
__global__ void __launch_bounds__(256) k_check_mtlids(const int* __restrict__ mtlids, int n, int max_materials, int* flag) {
    bool bad = false;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) { const int m = mtlids[i]; bad = bad || m < -1 || m >= max_materials; }
    if (bad) *flag = 1;
}

inline int nblk(long long n, int b = BLK) { return (int)((n + b - 1) / b); }

}  // namespace

// ================================================= host side ==========================================================
// where the tree kernel keeps the nodes of an n-leaf tree: in shared memory as packed 64-byte nodes, in shared memory as quantised
// 32-byte nodes (trees up to twice as big), or in global memory (quantised)
int ptb_tree_mode(const ptb_ctx* c, int n) {
    if (c->no_resident_bvh) return PTB_TREE_GLOBAL;
    if (!c->quant_resident_bvh && TraceSmem<PTB_TRACE_BLK_S>::fixed + TraceSmem<PTB_TRACE_BLK_S>::bvh(n) <= (size_t)c->smem_optin) return PTB_TREE_RESIDENT;
    if (TraceSmem<PTB_TRACE_BLK_S>::fixed + TraceSmem<PTB_TRACE_BLK_S>::qbvh(n) <= (size_t)c->smem_optin) return PTB_TREE_RESIDENT_QUANT;
    return PTB_TREE_GLOBAL;
}
// S16 (ptb_trace_kernel.cuh): can the traversal stack of this tree never hold more than PTB_S16_DEPTH entries?  An entry is pushed at
// a node whose two children are both internal and both hit, so at a node of level L (root = 1) the stack holds at most L entries, and
// such a node has at least two levels below it.  For the PLOC tree the build reports the height in internal-node levels along the
// longest root-to-leaf path (leaves 0, a merge max + 1): at most height - 1 entries.  The LBVH topology's depth is the reference's own
// count (lbvh.cu k_validate), taken as it is.
static bool ptb_s16_ok(const ptb_ctx* c, int n) {
    const int d = c->tree_info.trav_depth;
    const int need = c->tree_info.trav_ploc ? d - 1 : d;
    return c->s16_stack && d >= 1 && need <= PTB_S16_DEPTH && n <= 65536;
}
// one traversal launch: the persistent ordered kernel, or the literal reference-order kernel
// which: 0 = extend, 1 = shadow / taps (selects the tree-queue counters of the control block, reset by k_ctrl_*)
template <class IO>
static void launch_trace_io(ptb_ctx* c, const TraceScene& S, const IO& io, int policy, int which, int* cursor, const int* count_ptr, cudaStream_t st, const ExpQ& tq, Ctrl* ctrl) {
    DevCounters* ctr = c->counting ? c->d_counters : nullptr;
    if (policy == PTB_TRAVERSE_REFERENCE || S.n < 2) {
        if (c->counting) k_trace_simple<IO, 0, true><<<c->blocks_ref, PTB_TRACE_BLK, 0, st>>>(S, io, cursor, count_ptr, ctr);
        else k_trace_simple<IO, 0, false><<<c->blocks_ref, PTB_TRACE_BLK, 0, st>>>(S, io, cursor, count_ptr, ctr);
    } else if (policy == PTB_TRAVERSE_ORDERED_EXACT) {
        if (c->counting) k_trace_simple<IO, 1, true><<<c->blocks_exact, PTB_TRACE_BLK, 0, st>>>(S, io, cursor, count_ptr, ctr);
        else k_trace_simple<IO, 1, false><<<c->blocks_exact, PTB_TRACE_BLK, 0, st>>>(S, io, cursor, count_ptr, ctr);
    } else {
        int* n_tree = &ctrl->n_tree[which]; int* cur_tree = &ctrl->cur_tree[which];
        // phase A: per-ray setup, always-test list, root test; survivors -> tree queue
        if (c->counting) k_trace_pre<IO, true><<<c->blocks_generic, 256, 0, st>>>(S, io, count_ptr, tq, n_tree, ctr);
        else k_trace_pre<IO, false><<<c->blocks_generic, 256, 0, st>>>(S, io, count_ptr, tq, n_tree, ctr);
        // phase B: tree traversal of the survivors
        const int mode = ptb_tree_mode(c, S.n);
        if (mode == PTB_TREE_RESIDENT) {
            // the packed BVH fits in shared memory: one CTA per SM keeps it resident
            // a traversal tree of height <= 16: the stack as 32-bit entries, all of it in shared memory (S16)
            const bool s16 = ptb_s16_ok(c, S.n);
            auto kern = s16 ? (c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK_S, true, false, false, true> : k_trace_tree<IO, false, PTB_TRACE_BLK_S, true, false, false, true>)
                            : (c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK_S, true, false> : k_trace_tree<IO, false, PTB_TRACE_BLK_S, true, false>);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin);
            kern<<<c->sm_count, PTB_TRACE_BLK_S, TraceSmem<PTB_TRACE_BLK_S>::fixed + TraceSmem<PTB_TRACE_BLK_S>::bvh(S.n), st>>>(S, io, tq, cur_tree, n_tree, ctr);
        } else if (mode == PTB_TREE_RESIDENT_QUANT) {
            // twice the size: resident as quantised 32-byte nodes
            const bool s16 = ptb_s16_ok(c, S.n);
            auto kern = s16 ? (c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK_S, true, true, false, true> : k_trace_tree<IO, false, PTB_TRACE_BLK_S, true, true, false, true>)
                            : (c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK_S, true, true> : k_trace_tree<IO, false, PTB_TRACE_BLK_S, true, true>);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin);
            kern<<<c->sm_count, PTB_TRACE_BLK_S, TraceSmem<PTB_TRACE_BLK_S>::fixed + TraceSmem<PTB_TRACE_BLK_S>::qbvh(S.n), st>>>(S, io, tq, cur_tree, n_tree, ctr);
        } else if (S.wnodes) {
            auto kern = c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK, false, true, true> : k_trace_tree<IO, false, PTB_TRACE_BLK, false, true, true>;
            kern<<<c->sm_count * PTB_TRACE_MINB, PTB_TRACE_BLK, TraceSmem<PTB_TRACE_BLK, true>::fixed, st>>>(S, io, tq, cur_tree, n_tree, ctr);
        } else {
            auto kern = c->counting ? k_trace_tree<IO, true, PTB_TRACE_BLK, false, true> : k_trace_tree<IO, false, PTB_TRACE_BLK, false, true>;
            kern<<<c->sm_count * PTB_TRACE_MINB, PTB_TRACE_BLK, TraceSmem<PTB_TRACE_BLK>::fixed, st>>>(S, io, tq, cur_tree, n_tree, ctr);
        }
        c->launches++;
    }
    c->launches++;
}
static void launch_extend(ptb_ctx* c, const Lane& L, const TraceScene& S, int policy, const RayQueue& q) {
    launch_trace_io(c, S, ExtendIO{q.o, q.d, c->st.hit}, policy, 0, &L.ctrl->cur_extend, &L.ctrl->n_in, L.s_main, L.tq, L.ctrl);
}
static void launch_shadow(ptb_ctx* c, const Lane& L, const TraceScene& S, int policy, cudaStream_t st, const ExpQ& tq) {
    launch_trace_io(c, S, ShadowIO{L.sq.o, L.sq.d, L.sq.c, c->st.result}, policy, 1, &L.ctrl->cur_shadow, &L.ctrl->n_shadow, st, tq, L.ctrl);
}
// lane `which` of `nlanes`: an equal share of every pool; lane 0 runs on the context's stream pair, the others on their own
Lane ptb_lane(ptb_ctx* c, int which, int nlanes) {
    Lane L;
    const size_t off = (size_t)(c->max_paths / nlanes) * which;
    for (int k = 0; k < 2; k++) { L.xq[k].o = c->xq[k].o + off; L.xq[k].d = c->xq[k].d + off; L.xq[k].c = nullptr; }
    L.sq.o = c->sq.o + off; L.sq.d = c->sq.d + off; L.sq.c = c->sq.c + off;
    for (int k = 0; k < PTB_EXP_K; k++) { L.tq.e[k] = c->tq.e[k] + off; L.tq2.e[k] = c->tq2.e[k] + off; }
    L.ctrl = c->d_ctrl + which;
    L.s_main = which ? c->lane_main[which] : c->stream; L.s_side = which ? c->lane_side[which] : c->stream2;
    L.ev_shade = which ? c->lane_shade[which] : c->ev_shade; L.ev_shadow = which ? c->lane_shadow[which] : c->ev_shadow;
    return L;
}
// lanes 1.. start after everything already enqueued on the caller's stream, and the caller's stream continues after they are done
static int lanes_fork(ptb_ctx* c, int nlanes) {
    if (nlanes < 2) return 0;
    PTB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    for (int w = 1; w < nlanes; w++) PTB_CUDA(cudaStreamWaitEvent(c->lane_main[w], c->ev_fork, 0));
    return 0;
}
static int lanes_join(ptb_ctx* c, int nlanes) {
    for (int w = 1; w < nlanes; w++) { PTB_CUDA(cudaEventRecord(c->lane_join[w], c->lane_main[w])); PTB_CUDA(cudaStreamWaitEvent(c->stream, c->lane_join[w], 0)); }
    return 0;
}

void ptb_stage_begin(ptb_ctx* c, int stage) {
    if (!c->profiling) return;
    StageEvent ev; ev.stage = stage;
    for (cudaEvent_t* e : {&ev.a, &ev.b}) {
        if (!c->event_pool.empty()) { *e = c->event_pool.back(); c->event_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(ev.a, c->stream);
    c->events.push_back(ev);
}
void ptb_stage_end(ptb_ctx* c) {
    if (!c->profiling) return;
    cudaEventRecord(c->events.back().b, c->stream);
}
int ptb_stage_collect(ptb_ctx* c) {
    if (c->events.empty()) return 0;
    PTB_CUDA(cudaEventSynchronize(c->events.back().b));
    for (auto& ev : c->events) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        c->stage_ms[ev.stage] += ms;
        c->event_pool.push_back(ev.a); c->event_pool.push_back(ev.b);
    }
    c->events.clear();
    return 0;
}

int ptb_wf_init(ptb_ctx* c) {
    int64_t np = c->max_paths;
    float4** arrs[] = {&c->st.ray_o, &c->st.ray_d, &c->st.hit, &c->st.thr, &c->st.result,
                       &c->xq[0].o, &c->xq[0].d, &c->xq[1].o, &c->xq[1].d, &c->sq.o, &c->sq.d, &c->sq.c, &c->tq.e[0], &c->tq.e[1], &c->tq.e[2], &c->tq.e[3], &c->tq.e[4],
                       &c->tq2.e[0], &c->tq2.e[1], &c->tq2.e[2], &c->tq2.e[3], &c->tq2.e[4]};
    for (auto a : arrs) PTB_CUDA(cudaMalloc(a, sizeof(float4) * np));
    PTB_CUDA(cudaMalloc(&c->d_ctrl, PTB_MAX_LANES * sizeof(Ctrl)));          // one control block per lane
    PTB_CUDA(cudaMemset(c->d_ctrl, 0, PTB_MAX_LANES * sizeof(Ctrl)));
    PTB_CUDA(cudaMalloc(&c->d_counters, sizeof(DevCounters)));
    PTB_CUDA(cudaMemset(c->d_counters, 0, sizeof(DevCounters)));
    PTB_CUDA(cudaMalloc(&c->d_params, sizeof(SceneParams)));
    PTB_CUDA(cudaMalloc(&c->d_cache, sizeof(SceneCache)));
    int occ = 0;
    PTB_CUDA(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    c->no_resident_bvh = getenv("PTB_NO_RESIDENT_BVH") != nullptr;
    c->quant_resident_bvh = getenv("PTB_QUANT_RESIDENT_BVH") != nullptr;
    c->overlap_shadow = getenv("PTB_NO_OVERLAP") == nullptr;
    c->use_ploc = getenv("PTB_NO_PLOC") == nullptr;
    c->ploc_big = getenv("PTB_NO_PLOC_BIG") == nullptr;
    c->wide4 = getenv("PTB_NO_WIDE4") == nullptr;
    c->s16_stack = getenv("PTB_NO_S16") == nullptr;
    if (const char* r = getenv("PTB_PLOC_RADIUS")) { const int v = atoi(r); if (v >= 1 && v <= 1024) c->ploc_radius = v; }
    PTB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    PTB_CUDA(cudaEventCreateWithFlags(&c->ev_shade, cudaEventDisableTiming));
    PTB_CUDA(cudaEventCreateWithFlags(&c->ev_shadow, cudaEventDisableTiming));
    for (int w = 1; w < PTB_MAX_LANES; w++) {
        PTB_CUDA(cudaStreamCreateWithFlags(&c->lane_main[w], cudaStreamNonBlocking));
        PTB_CUDA(cudaStreamCreateWithFlags(&c->lane_side[w], cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&c->lane_shade[w], &c->lane_shadow[w], &c->lane_join[w]}) PTB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    for (int w = 0; w < PTB_MAX_LANES; w++) PTB_CUDA(cudaEventCreateWithFlags(&c->ev_acc[w], cudaEventDisableTiming));
    PTB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    if (const char* v = getenv("PTB_MLT_LANES")) { const int k = atoi(v); if (k >= 1 && k <= PTB_MAX_LANES) c->mlt_lanes = k; }
    if (const char* v = getenv("PTB_PT_LANES")) { const int k = atoi(v); if (k >= 1 && k <= PTB_MAX_LANES) c->pt_lanes = k; }
    if (getenv("PTB_MLT_ONE_LANE")) c->mlt_lanes = 1;
    if (getenv("PTB_PT_ONE_LANE")) c->pt_lanes = 1;
    PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_simple<ExtendIO, 1, false>, PTB_TRACE_BLK, 0));
    c->blocks_exact = c->sm_count * (occ > 0 ? occ : 4);
    PTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_simple<ExtendIO, 0, false>, PTB_TRACE_BLK, 0));
    c->blocks_ref = c->sm_count * (occ > 0 ? occ : 4);
    c->blocks_generic = c->sm_count * 8;
    return 0;
}

int ptb_wf_upload_params(ptb_ctx* c) {
    if (!c->params_dirty) return 0;
    c->h_params.nx = c->nx; c->h_params.ny = c->ny;
    // pageable host source: the copy is staged before the call returns, so later host edits cannot race it
    PTB_CUDA(cudaMemcpyAsync(c->d_params, &c->h_params, sizeof(SceneParams), cudaMemcpyHostToDevice, c->stream));
    ptb_shade_prepare_cache(c);
    c->params_dirty = false;
    return 0;
}

int ptb_wf_sobol_points(ptb_ctx* c, int k_first, int count, int stride, float* P_dev) {
    if (!c->d_sobolV) { ptb_set_error("Sobol direction table not set (ptb_set_sobol_table)"); return 1; }
    k_sobol_points<<<nblk(c->sobol_dim, 256), 256, 0, c->stream>>>(c->d_sobolV, c->sobol_dim, k_first, count, stride, P_dev);
    c->launches++;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

static int ensure_sobolP(ptb_ctx* c, int nsamp) {
    if (nsamp > c->sobolP_cap) {
        if (c->d_sobolP) cudaFree(c->d_sobolP);
        PTB_CUDA(cudaMalloc(&c->d_sobolP, sizeof(float) * (size_t)nsamp * c->sobol_dim));
        c->sobolP_cap = nsamp;
    }
    return 0;
}

// bounce loop shared by every engine: extend -> shade -> shadow, up to 5 times (path.py:25)
template <int ENGINE>
static int run_bounces(ptb_ctx* c, const Lane& L, const float* rngtab, int dim, int rng_stride, FrameMap fm) {
    TraceScene S = ptb_trace_scene(c);
    int policy = ptb_effective_policy(c, c->traversal_request);
    cudaStream_t st = L.s_main;
    int cur = 0;
    // The shadow stage of bounce b and the extend stage of bounce b+1 are independent (shadow rays add into `result`, which extend
    // never touches; shade of b+1 waits for both), so they run on two streams: each fills the SMs the other leaves idle while its
    // persistent kernels drain.  The per-path order of the additions into `result` is unchanged.  Off while stages are being timed.
    const bool overlap = ENGINE == PTB_ENGINE_PATH && c->overlap_shadow && !c->profiling;
    for (int depth = 1; depth <= 5; depth++) {
        ptb_stage_begin(c, ST_EXTEND);
        launch_extend(c, L, S, policy, L.xq[cur]);
        ptb_stage_end(c);
        if (overlap && depth > 1) {
            PTB_CUDA(cudaStreamWaitEvent(st, L.ev_shadow, 0));           // the previous shadow stage is done with sq / result
            k_ctrl_reset_shadow<<<1, 1, 0, st>>>(L.ctrl);
            c->launches++;
        }
        ptb_stage_begin(c, ST_SHADE);
        ptb_shade_launch(c, L, ENGINE, rngtab, dim, rng_stride, fm, cur, st);
        ptb_stage_end(c);
        if (ENGINE == PTB_ENGINE_PATH) {
            if (overlap) {
                PTB_CUDA(cudaEventRecord(L.ev_shade, st));
                PTB_CUDA(cudaStreamWaitEvent(L.s_side, L.ev_shade, 0));
                launch_shadow(c, L, S, policy, L.s_side, L.tq2);
                PTB_CUDA(cudaEventRecord(L.ev_shadow, L.s_side));
            } else {
                ptb_stage_begin(c, ST_SHADOW);
                launch_shadow(c, L, S, policy, st, L.tq);
                ptb_stage_end(c);
            }
        }
        if (overlap) k_ctrl_next_extend<<<1, 1, 0, st>>>(L.ctrl);
        else k_ctrl_next<<<1, 1, 0, st>>>(L.ctrl);
        c->launches++;
        cur ^= 1;
    }
    if (overlap) PTB_CUDA(cudaStreamWaitEvent(st, L.ev_shadow, 0));
    PTB_CUDA(cudaGetLastError());
    return 0;
}

// window: NULL = the whole film, `count` Sobol points k_first, k_first+stride, ...; else {x0, y0, w, h} = `count` samples of every pixel of
// that window, all at Sobol point k_first (render_tile, engine/path.py:96-118)
int ptb_wf_render(ptb_ctx* c, int engine, int k_first, int count, int stride, float* sample_out_dev, const int* window) {
    if (c->nx <= 0 || c->ny <= 0) { ptb_set_error("film size not set (ptb_set_size)"); return 1; }
    if (c->tree_n != c->nfaces) { ptb_set_error("BVH is stale: call ptb_build_tree after ptb_load_model"); return 1; }
    if (engine != PTB_ENGINE_MLT && !c->d_sobolV) { ptb_set_error("Sobol direction table not set (ptb_set_sobol_table)"); return 1; }
    if (ptb_wf_upload_params(c)) return 1;
    cudaStream_t st = c->stream;
    DevCounters* ctr = c->counting ? c->d_counters : nullptr;
    FrameMap fm = window ? make_window(c->nx, c->ny, window[0], window[1], window[2], window[3]) : make_frame(c->nx, c->ny);
    fm.d_dim = make_fastmod((unsigned)c->sobol_dim);

    if (engine == PTB_ENGINE_MLT) {
        if (c->mlt_count <= 0) { ptb_set_error("MLT chains not initialised (ptb_mlt_reset)"); return 1; }
        // The chains are independent of each other from one render() to the next, and a wavefront of 2^18 paths leaves the persistent
        // kernels mostly tail: the population runs as two halves on two lanes (own queues, control block and streams), which only
        // meet again when the call returns.  Splats are atomic, chain state is per chain: nothing is shared but the film.
        const int n = c->mlt_count;
        const int nlanes = (c->mlt_lanes > 1 && !c->profiling && n >= 8192 && (long long)n <= c->max_paths) ? c->mlt_lanes : 1;
        if (lanes_fork(c, nlanes)) return 1;
        for (int it = 0; it < count; it++) {
            for (int w = 0; w < nlanes; w++) {
                const Lane L = ptb_lane(c, w, nlanes);
                const int first = (int)((long long)n * w / nlanes), cnt = (int)((long long)n * (w + 1) / nlanes) - first;
                ptb_stage_begin(c, ST_RAYGEN);
                k_ctrl_begin<<<1, 1, 0, L.s_main>>>(L.ctrl, 0);
                k_mlt_raygen<<<nblk(cnt), BLK, 0, L.s_main>>>(c->d_params, c->d_Xold, c->d_Xnew, first, cnt, c->mlt_first, c->mlt_seed, c->mlt_iter, c->mlt_lsp, c->mlt_sigma,
                                                              c->st, L.xq[0], L.ctrl, ctr);
                ptb_stage_end(c);
                c->launches += 2;
                if (run_bounces<PTB_ENGINE_PATH>(c, L, c->d_Xnew, 0, 32, fm)) return 1;
                ptb_stage_begin(c, ST_ACCUM);
                k_mlt_finish<<<nblk(cnt), BLK, 0, L.s_main>>>(c->d_Xold, c->d_Xnew, c->d_Lold, c->st.result, first, cnt, c->mlt_first, c->mlt_seed, c->mlt_iter, c->nx, c->ny, c->d_film);
                ptb_stage_end(c);
                c->launches++;
            }
            c->mlt_iter++;
        }
        if (lanes_join(c, nlanes)) return 1;
        PTB_CUDA(cudaGetLastError());
        return 0;
    }

    int per_batch = (int)(c->max_paths / fm.pps);
    if (window && count > per_batch) { ptb_set_error("a window of %d samples needs %lld path slots, pool has %lld", count, (long long)count * fm.pps, (long long)c->max_paths); return 1; }
    if (per_batch < 1) { ptb_set_error("film %dx%d needs %d path slots per sample, pool has %lld", c->nx, c->ny, fm.pps, (long long)c->max_paths); return 1; }
    if (sample_out_dev) count = 1;
    // A render of several batches (config 4: 4 samples of a 1080p film fill the pool) runs its batches as half-sized chunks that
    // alternate between the two lanes: chunk b+1 traces while chunk b drains, each filling the SMs the other's persistent kernels leave
    // idle.  Chunks add to the film in Sobol order (the accumulate of chunk b waits for that of chunk b-1), so the film is bit-identical.
    int nlanes = (c->pt_lanes > 1 && !c->profiling && !window && !sample_out_dev && engine != PTB_ENGINE_PREVIEW && (count > per_batch || (c->pt_split && count >= c->pt_lanes))) ? c->pt_lanes : 1;
    while (nlanes > 1 && (c->max_paths / nlanes) / fm.pps < 1) nlanes--;
    const bool two = nlanes > 1;
    int chunk = two ? (int)((c->max_paths / nlanes) / fm.pps) : per_batch;
    if (two && count <= per_batch) chunk = (count + nlanes - 1) / nlanes < chunk ? (count + nlanes - 1) / nlanes : chunk;      // "pt_split": one batch as nlanes concurrent chunks
    if (ensure_sobolP(c, two ? nlanes * chunk : (count < chunk ? count : chunk))) return 1;
    if (lanes_fork(c, nlanes)) return 1;
    int b = 0;
    for (int done = 0; done < count; done += chunk, b++) {
        const int ns = count - done < chunk ? count - done : chunk;
        const int w = b % nlanes;
        const Lane L = ptb_lane(c, w, nlanes);
        cudaStream_t ls = L.s_main;
        FrameMap lfm = fm;
        lfm.slot_base = (int)((c->max_paths / nlanes) * w);
        float* P = c->d_sobolP + (size_t)w * chunk * c->sobol_dim;
        ptb_stage_begin(c, ST_RAYGEN);
        k_sobol_points<<<nblk(c->sobol_dim, 256), 256, 0, ls>>>(c->d_sobolV, c->sobol_dim, window ? k_first : k_first + done * stride, window ? 1 : ns, window ? 1 : stride, P);
        k_ctrl_begin<<<1, 1, 0, ls>>>(L.ctrl, 0);
        k_raygen<<<c->blocks_generic, BLK, 0, ls>>>(c->d_params, P, c->sobol_dim, lfm, ns, engine != PTB_ENGINE_PREVIEW, c->st, L.xq[0], L.ctrl, ctr);
        ptb_stage_end(c);
        c->launches += 3;
        if (engine == PTB_ENGINE_PREVIEW) {
            TraceScene S = ptb_trace_scene(c);
            int policy = ptb_effective_policy(c, c->traversal_request);
            ptb_stage_begin(c, ST_EXTEND);
            launch_extend(c, L, S, policy, L.xq[0]);
            ptb_stage_end(c);
            ptb_stage_begin(c, ST_ACCUM);
            size_t pass = (size_t)c->caps.max_filmsize;
            k_preview<<<nblk(fm.pps), BLK, 0, ls>>>(c->d_params, c->d_texels, c->d_verts, c->d_mtlids, lfm, ns, c->st, c->d_film + pass, c->d_film + 2 * pass);
            ptb_stage_end(c);
            c->launches += 1;
            continue;
        }
        int rc = engine == PTB_ENGINE_PATH ? run_bounces<PTB_ENGINE_PATH>(c, L, P, c->sobol_dim, 0, lfm)
                                           : run_bounces<PTB_ENGINE_BRUTE>(c, L, P, c->sobol_dim, 0, lfm);
        if (rc) return 1;
        if (two && b > 0) PTB_CUDA(cudaStreamWaitEvent(ls, c->ev_acc[(w + nlanes - 1) % nlanes], 0));   // Sobol order of the additions into the film
        ptb_stage_begin(c, ST_ACCUM);
        k_accumulate<<<nblk(fm.pps), BLK, 0, ls>>>(c->d_film, c->st.result, lfm, ns, sample_out_dev);
        ptb_stage_end(c);
        if (two) PTB_CUDA(cudaEventRecord(c->ev_acc[w], ls));
        c->launches++;
    }
    if (lanes_join(c, nlanes)) return 1;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_trace_primary(ptb_ctx* c, int k, float* rays_dev, int32_t* hit_dev, float* depth_dev, int32_t* index_dev, float* uv_dev) {
    if (c->nx <= 0 || c->ny <= 0) { ptb_set_error("film size not set (ptb_set_size)"); return 1; }
    if (ptb_wf_upload_params(c)) return 1;
    cudaStream_t st = c->stream;
    FrameMap fm = make_frame(c->nx, c->ny);
    fm.d_dim = make_fastmod((unsigned)c->sobol_dim);
    if (fm.pps > c->max_paths) { ptb_set_error("film larger than the path pool"); return 1; }
    if (ensure_sobolP(c, 1)) return 1;
    if (ptb_wf_sobol_points(c, k, 1, 1, c->d_sobolP)) return 1;
    const Lane L0 = ptb_lane(c, 0, 1);
    k_ctrl_begin<<<1, 1, 0, st>>>(L0.ctrl, 0);
    k_raygen<<<c->blocks_generic, BLK, 0, st>>>(c->d_params, c->d_sobolP, c->sobol_dim, fm, 1, 1, c->st, L0.xq[0], L0.ctrl, nullptr);
    if (rays_dev) k_gather_primary<<<nblk(fm.pps), BLK, 0, st>>>(fm, c->st, rays_dev, nullptr, nullptr, nullptr, nullptr, 0);
    if (hit_dev || depth_dev || index_dev || uv_dev) {
        if (c->tree_n != c->nfaces) { ptb_set_error("BVH is stale: call ptb_build_tree after ptb_load_model"); return 1; }
        TraceScene S = ptb_trace_scene(c);
        int policy = ptb_effective_policy(c, c->traversal_request);
        launch_extend(c, L0, S, policy, L0.xq[0]);
        k_gather_primary<<<nblk(fm.pps), BLK, 0, st>>>(fm, c->st, nullptr, hit_dev, depth_dev, index_dev, uv_dev, 1);
    }
    c->launches += 4;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_intersect(ptb_ctx* c, const float* rays_dev, const int32_t* avoid_dev, const float* dis_dev, int m, int policy, int anyhit,
                     int32_t* hit_dev, float* depth_dev, int32_t* index_dev, float* uv_dev) {
    if (c->tree_n != c->nfaces) { ptb_set_error("BVH is stale: call ptb_build_tree after ptb_load_model"); return 1; }
    if (m > c->max_paths) { ptb_set_error("%d rays exceed the ray queue capacity %lld", m, (long long)c->max_paths); return 1; }
    TraceScene S = ptb_trace_scene(c);
    k_ctrl_tap<<<1, 1, 0, c->stream>>>(c->d_ctrl, m);
    k_pack_tap<<<nblk(m), BLK, 0, c->stream>>>(rays_dev, avoid_dev, anyhit ? dis_dev : nullptr, c->d_slot_of, c->nfaces, m, c->sq);
    c->launches += 2;
    int eff = ptb_effective_policy(c, policy);
    int* cur = &c->d_ctrl->pad[1]; const int* cnt = &c->d_ctrl->pad[0];
    if (anyhit) launch_trace_io(c, S, TapIO<true>{c->sq.o, c->sq.d, c->sq.c, hit_dev, depth_dev, index_dev, uv_dev}, eff, 1, cur, cnt, c->stream, c->tq, c->d_ctrl);
    else launch_trace_io(c, S, TapIO<false>{c->sq.o, c->sq.d, c->sq.c, hit_dev, depth_dev, index_dev, uv_dev}, eff, 1, cur, cnt, c->stream, c->tq, c->d_ctrl);
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_resolve(ptb_ctx* c, int pass, int mode, float* out_dev) {
    const float4* film = c->d_film + (size_t)pass * c->caps.max_filmsize;
    k_resolve<<<nblk((long long)c->nx * c->ny), BLK, 0, c->stream>>>(film, c->nx, c->ny, mode, out_dev);
    c->launches++;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_batch_capacity(const ptb_ctx* c) {
    if (c->nx <= 0 || c->ny <= 0) return 1;
    const long long cap = c->max_paths / make_frame(c->nx, c->ny).pps;
    return cap < 1 ? 1 : (int)cap;
}
int ptb_wf_normaldist(ptb_ctx* c, const float* in_dev, int m, float* out_dev) {
    k_normaldist_tap<<<nblk(m), BLK, 0, c->stream>>>(in_dev, m, out_dev);
    c->launches++;
    PTB_CUDA(cudaGetLastError());
    return 0;
}
int ptb_wf_check_mtlids(ptb_ctx* c, int nfaces) {
    int grid = nblk(nfaces, 256); if (grid > c->sm_count * 8) grid = c->sm_count * 8;
    k_check_mtlids<<<grid, 256, 0, c->stream>>>(c->d_mtlids, nfaces, c->caps.max_materials, c->d_flags);
    c->launches++;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_mlt_reset(ptb_ctx* c) {
    k_mlt_reset<<<nblk(c->mlt_count), BLK, 0, c->stream>>>(c->d_Xold, c->d_Lold, c->mlt_count, c->mlt_first, c->mlt_seed);
    c->launches++;
    c->mlt_iter = 0;
    PTB_CUDA(cudaGetLastError());
    return 0;
}

int ptb_wf_selftest(ptb_ctx* c, int what, long long n, unsigned long long seed, long long* fails) {
    unsigned long long* d = nullptr;
    PTB_CUDA(cudaMalloc(&d, 8));
    PTB_CUDA(cudaMemsetAsync(d, 0, 8, c->stream));
    k_selftest_div<<<c->sm_count * 8, 256, 0, c->stream>>>(what, n, seed, d);
    c->launches++;
    unsigned long long h = 0;
    PTB_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    *fails = (long long)h;
    return 0;
}

int ptb_wf_measure_l2(ptb_ctx* c, int mbytes, int iters, float* gbps) {
    size_t bytes = (size_t)mbytes << 20, n4 = bytes / 16;
    float4* buf = nullptr; float* sink = nullptr;
    PTB_CUDA(cudaMalloc(&buf, bytes));
    PTB_CUDA(cudaMalloc(&sink, 4));
    PTB_CUDA(cudaMemsetAsync(buf, 0, bytes, c->stream));
    int grid = c->sm_count * 8;
    k_l2_read<<<grid, 256, 0, c->stream>>>(buf, n4, 2, sink);   // warm the L2
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
    k_l2_read<<<grid, 256, 0, c->stream>>>(buf, n4, iters, sink);
    cudaEventRecord(b, c->stream);
    PTB_CUDA(cudaEventSynchronize(b));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    *gbps = (float)((double)bytes * iters / (ms * 1e-3) / 1e9);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    c->launches += 2;
    return 0;
}
