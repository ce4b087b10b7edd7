// ptb_wavefront.cuh -- device helpers shared by the wavefront kernels (wavefront.cu) and the shading kernels (shade.cu):
// pixel <-> path-slot mapping, the per-path random source, queue appends, the shading frame of a hit.
#pragma once
#include "ptb_internal.h"

// ---- pixel <-> path-slot mapping: a warp owns an 8x4 pixel tile so primary rays stay coherent (FrameMap: ptb_internal.h) -----
PTB_D bool slot_pixel(const FrameMap& f, int q, int* x, int* y) {
    int tile = q >> 5, lane = q & 31;
    int tx = (int)fastdiv((unsigned)tile, f.d_ty), ty = tile - tx * f.tiles_y;
    *x = f.x0 + tx * 8 + (lane >> 2);
    *y = f.y0 + ty * 4 + (lane & 3);
    return *x < f.nx && *y < f.ny;
}
// random source of sample s of pixel (x, y): Sobol point s of the batch rotated by wanghash2(x, y) (path.py:72-73), or -- window mode --
// the one point of the call rotated by wanghash3(x, y, s)
struct Rng;
PTB_D void frame_rng(const FrameMap& f, const float* __restrict__ tab, int dim, int s, int x, int y, Rng* rng);

// ---- random source: Sobol table of the path's sample (sobol.py:107-125) or the MLT chain vector (sampling/__init__.py:53-64)
struct Rng {
    const float* __restrict__ tab;
    int base, dim;   // dim == 0 -> direct indexing (MLT)
    FastMod fmod;    // i mod dim without a runtime division
    PTB_D float draw(int c) const {
        if (dim == 0) return tab[c];
        int i = (int)((unsigned)base + (unsigned)c);   // i32 wrap of `self.i += 1`
        return __ldg(&tab[fastpymod(i, fmod)]);
    }
};
PTB_D void frame_rng(const FrameMap& f, const float* __restrict__ tab, int dim, int s, int x, int y, Rng* rng) {
    rng->dim = dim; rng->fmod = f.d_dim;
    if (f.window) { rng->tab = tab; rng->base = wanghash(s ^ wanghash2(x, y)); }
    else { rng->tab = tab + (size_t)s * dim; rng->base = wanghash2(x, y); }
}
// k consecutive draws starting at draw c0: one modulo, then increments with wrap-around (identical indices; the i32 wrap of
// `self.i += 1` inside the run falls back to the per-draw modulo)
struct RngRun {
    const Rng& g; int i0, b0; bool fast;
    PTB_D RngRun(const Rng& g_, int c0) : g(g_), i0(0), b0(0), fast(false) {
        if (g.dim != 0) {
            i0 = (int)((unsigned)g.base + (unsigned)c0);
            fast = i0 <= 0x7fffffff - 16;
            b0 = fastpymod(i0, g.fmod);
        } else i0 = c0;
    }
    PTB_D float draw(int j) const {     // j-th draw of the run, j < 16
        if (g.dim == 0) return g.tab[i0 + j];
        int b;
        if (fast) { b = b0 + j; if (b >= g.dim) b -= g.dim; }
        else b = fastpymod((int)((unsigned)i0 + (unsigned)j), g.fmod);
        return __ldg(&g.tab[b]);
    }
};

// ---- queue append: warp ballot + block prefix, one atomic per block, contiguous (coalesced) writes -------------
// returns the queue position reserved for this thread (-1 if !flag).
// TAIL_SYNC = false: the caller alternates between two sets of shared buffers from one call to the next, which makes the trailing
// barrier (it only protects the buffers against the NEXT call's writes) unnecessary: a warp can reach the writes of call k+2 only after
// the barriers of call k+1, which every warp passes after its reads of call k.
template <int NT, bool TAIL_SYNC = true>
PTB_D int block_append(bool flag, int* counter, int* s_warp, int* s_base) {
    unsigned m = __ballot_sync(0xffffffffu, flag);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; i++) { int c = s_warp[i]; s_warp[i] = tot; tot += c; }
        *s_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    int pos = flag ? *s_base + s_warp[w] + rank : -1;
    if (TAIL_SYNC) __syncthreads();
    return pos;
}

// two queues at once (next extend queue + shadow queue): one 64-bit atomic per block on the adjacent counters (n_out, n_shadow)
template <int NT, bool TAIL_SYNC = true>
PTB_D void block_append2(bool fa, bool fb, int* counter_pair, int* s_warp /* [2 * NT/32] */, unsigned long long* s_base, int* pa, int* pb) {
    const unsigned ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_warp[w] = __popc(ma); s_warp[NT / 32 + w] = __popc(mb); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ta = 0, tb = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; i++) { int c = s_warp[i]; s_warp[i] = ta; ta += c; c = s_warp[NT / 32 + i]; s_warp[NT / 32 + i] = tb; tb += c; }
        *s_base = (ta | tb) ? atomicAdd(reinterpret_cast<unsigned long long*>(counter_pair), (unsigned long long)(unsigned)ta | ((unsigned long long)(unsigned)tb << 32)) : 0ull;
    }
    __syncthreads();
    const unsigned long long base = *s_base;
    *pa = fa ? (int)(unsigned)base + s_warp[w] + __popc(ma & ((1u << lane) - 1u)) : -1;
    *pb = fb ? (int)(unsigned)(base >> 32) + s_warp[NT / 32 + w] + __popc(mb & ((1u << lane) - 1u)) : -1;
    if (TAIL_SYNC) __syncthreads();
}

// ---- model.py:88-101 get_geometries + geometries.py:96-108 -------------------------------------------------------------------
PTB_D void shading_frame(const float* __restrict__ verts, const int* __restrict__ mtlids, int f, float u, float v, V3 ro, V3 rd, float depth,
                         V3* hitpos, V3* normal, float* tu, float* tv, int* mtlid) {
    const float4* p = reinterpret_cast<const float4*>(verts + (size_t)f * 24);   // 3 corners x (pos3 nrm3 uv2) = 6 x float4
    float4 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3), c0 = __ldg(p + 4), c1 = __ldg(p + 5);
    float wx = 1.0f - u - v, wy = u, wz = v;
    V3 n0 = mk3(a0.w, a1.x, a1.y), n1 = mk3(b0.w, b1.x, b1.y), n2 = mk3(c0.w, c1.x, c1.y);
    V3 nrm = normalized(wx * n0 + wy * n1 + wz * n2);
    *tu = wx * a1.z + wy * b1.z + wz * c1.z;
    *tv = wx * a1.w + wy * b1.w + wz * c1.w;
    // the next ray's origin: explicit IEEE multiply and add, whatever contraction the translation unit is compiled with
    *hitpos = mk3(__fadd_rn(ro.x, __fmul_rn(depth, rd.x)), __fadd_rn(ro.y, __fmul_rn(depth, rd.y)), __fadd_rn(ro.z, __fmul_rn(depth, rd.z)));
    float sg = -dot(rd, nrm);
    if (sg < 0.0f) nrm = -nrm;
    *normal = nrm;
    *mtlid = __ldg(&mtlids[f]);
}

