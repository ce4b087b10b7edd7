// lbvh.cu -- LBVH build on the device, bit-identical to the reference's tree/lbvh.py:
//   genMortonCodes (lbvh.py:168-183) -> sortMortonCodes (lbvh.py:186-208; the reference round-trips through the host
//   and np.argsort -- here cub::DeviceRadixSort, stable, so ties keep face order) -> genHierarchy (lbvh.py:211-231
//   with determineRange lbvh.py:93-145, findSplit lbvh.py:61-90 and the quirky clz lbvh.py:33-42) ->
//   genAABBs (lbvh.py:251-294, level-synchronous sweeps; min/max are exact so the boxes are identical).
// Then: validation (every node one parent, leaf slots in order, height) and packing into the 64-byte node /
// triangle records the traversal kernels read (ptb_traverse.cuh).
#include <cstring>
#include <cub/cub.cuh>
#include "ptb_internal.h"

namespace {

constexpr int BLK = 256;
#define PTB_SMALL_TREE 8192          /* internal nodes up to which the box sweeps run in one block */

__device__ __forceinline__ void atomicMinF(float* a, float v) {
    if (v >= 0.0f) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxF(float* a, float v) {
    if (v >= 0.0f) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned*)a, __float_as_uint(v));
}

__device__ __forceinline__ V3 face_vertex(const float* __restrict__ verts, int face, int k) {
    const float* p = verts + (size_t)(face * 3 + k) * 8;
    return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
// lbvh.py:161-165 getCenter
__device__ __forceinline__ V3 face_center(const float* __restrict__ verts, int face) {
    return (face_vertex(verts, face, 0) + face_vertex(verts, face, 1) + face_vertex(verts, face, 2)) / 3.0f;
}

// scalars layout (floats): [0..2] centre min, [3..5] centre max ; ints: [8] remaining, [9] invalid flag, [10] root height,
// [11] always-test list length, [12] unused, [13] sweeps, [14] node planes outside the quantisation grid, [15] height of the PLOC traversal tree
__global__ void k_init_scalars(float* s) {
    s[0] = s[1] = s[2] = PTB_INF;
    s[3] = s[4] = s[5] = -PTB_INF;
    ((int*)s)[8] = 0; ((int*)s)[9] = 0; ((int*)s)[10] = 0; ((int*)s)[11] = 0; ((int*)s)[12] = 0; ((int*)s)[13] = 0; ((int*)s)[14] = 0; ((int*)s)[15] = -1;
}

// lbvh.py:172-176: bounds of the triangle CENTRES (atomic min/max; exact, order-free)
__global__ void __launch_bounds__(BLK) k_center_bounds(const float* __restrict__ verts, int n, float* s) {
    V3 lo = v3s(PTB_INF), hi = v3s(-PTB_INF);
    for (int i = blockIdx.x * BLK + threadIdx.x; i < n; i += gridDim.x * BLK) {
        V3 c = face_center(verts, i);
        lo = vmin(lo, c); hi = vmax(hi, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(~0u, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(~0u, lo.y, o)); lo.z = fminf(lo.z, __shfl_xor_sync(~0u, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(~0u, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(~0u, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(~0u, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMinF(&s[0], lo.x); atomicMinF(&s[1], lo.y); atomicMinF(&s[2], lo.z);
        atomicMaxF(&s[3], hi.x); atomicMaxF(&s[4], hi.y); atomicMaxF(&s[5], hi.z);
    }
}

// lbvh.py:12-23 expandBits (wrapping i32 multiply == u32 multiply), lbvh.py:26-30 morton3D
__device__ __forceinline__ int expand_bits(int v_) {
    unsigned v = (unsigned)v_;
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return (int)v;
}
__device__ __forceinline__ int morton3d(V3 v) {
    int wx = expand_bits(clampi(ifloor(v.x * 1024.0f), 0, 1023));
    int wy = expand_bits(clampi(ifloor(v.y * 1024.0f), 0, 1023));
    int wz = expand_bits(clampi(ifloor(v.z * 1024.0f), 0, 1023));
    return wx * 4 + wy * 2 + wz;
}
// lbvh.py:178-183
__global__ void __launch_bounds__(BLK) k_morton(const float* __restrict__ verts, int n, const float* __restrict__ s, int* mc, int* id) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n) return;
    V3 lo = mk3(s[0], s[1], s[2]), hi = mk3(s[3], s[4], s[5]);
    V3 c = face_center(verts, i);
    V3 coord = (c - lo) / (hi - lo);
    mc[i] = morton3d(coord);
    id[i] = i;
}

// lbvh.py:33-42: min(clz32(x)+1, 32) -- XOR values 0 and 1 are indistinguishable, as in the reference
__device__ __forceinline__ int clz_ref(int x) { return min(__clz(x) + 1, 32); }

// lbvh.py:61-90
__device__ int find_split(const int* __restrict__ mc, int l, int r) {
    int lc = mc[l], rc = mc[r];
    if (lc == rc) return (l + r) >> 1;
    int cp = clz_ref(lc ^ rc);
    int m = l, s = r - l;
    while (true) {
        s += 1; s >>= 1;
        int nn = m + s;
        if (nn < r) {
            int sp = clz_ref(lc ^ mc[nn]);
            if (sp > cp) m = nn;
        }
        if (s <= 1) break;
    }
    return m;
}
// lbvh.py:93-145
__device__ void determine_range(const int* __restrict__ mc, int n, int i, int* lo, int* ro) {
    int l = 0, r = n - 1;
    if (i != 0) {
        int ic = mc[i], lc = mc[i - 1], rc = mc[i + 1];
        if (lc == ic && ic == rc) {
            l = i;
            while (i < n - 1) {
                i += 1;
                if (i >= n - 1) break;
                if (mc[i] != mc[i + 1]) break;
            }
            r = i;
        } else {
            int ld = clz_ref(ic ^ lc), rd = clz_ref(ic ^ rc);
            int d = rd > ld ? 1 : -1;
            int delta_min = min(ld, rd);
            int lmax = 2;
            int delta = -1;
            int itmp = i + d * lmax;
            if (0 <= itmp && itmp < n) delta = clz_ref(ic ^ mc[itmp]);
            while (delta > delta_min) {
                lmax <<= 1;
                itmp = i + d * lmax;
                delta = -1;
                if (0 <= itmp && itmp < n) delta = clz_ref(ic ^ mc[itmp]);
            }
            int s = 0;
            for (int t = lmax >> 1; t > 0; t >>= 1) {
                itmp = i + (s + t) * d;
                delta = -1;
                if (0 <= itmp && itmp < n) delta = clz_ref(ic ^ mc[itmp]);
                if (delta > delta_min) s += t;
            }
            l = i; r = i + s * d;
            if (d < 0) { int t = l; l = r; r = t; }
        }
    }
    *lo = l; *ro = r;
}
// lbvh.py:211-231
__global__ void __launch_bounds__(BLK) k_hierarchy(const int* __restrict__ mc, const int* __restrict__ id, int n, int* leaf, int2* child, int* parentcnt) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i < n) leaf[i] = id[i];
    if (i >= n - 1) return;
    int l, r;
    determine_range(mc, n, i, &l, &r);
    int split = find_split(mc, l, r);
    int lhs = split; if (lhs != l) lhs += n;
    int rhs = split + 1; if (rhs != r) rhs += n;
    child[i] = make_int2(lhs, rhs);
    if (lhs >= 0 && lhs < 2 * n) atomicAdd(&parentcnt[lhs], 1);
    if (rhs >= 0 && rhs < 2 * n) atomicAdd(&parentcnt[rhs], 1);
}

// lbvh.py:155-158 getBoundingBox
__device__ __forceinline__ void face_box(const float* __restrict__ verts, int face, V3* lo, V3* hi) {
    V3 a = face_vertex(verts, face, 0), b = face_vertex(verts, face, 1), c = face_vertex(verts, face, 2);
    *lo = vmin(vmin(a, b), c); *hi = vmax(vmax(a, b), c);
}

// lbvh.py:272-294 genAABBSubstep as a Jacobi sweep: a node becomes ready in sweep `stamp` iff both children
// were ready BEFORE this sweep (stamp < current), so there is no intra-sweep race.  Also carries the slot range
// and height used for validation.
// returns true if node i is still not ready after this attempt
__device__ __forceinline__ bool aabb_try(int i, const float* __restrict__ verts, const int* __restrict__ leaf, const int2* __restrict__ child, int n, int stamp,
                                         float* bmin, float* bmax, int* ready, int2* range, int* height, int* scal) {
    if (ready[i] != 0) return false;
    int2 ch = child[i];
    V3 lo[2], hi[2]; int2 rg[2]; int hg[2];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int c = k == 0 ? ch.x : ch.y;
        if (c < 0 || c >= 2 * n - 1) { ok = false; break; }
        if (c < n) {
            face_box(verts, leaf[c], &lo[k], &hi[k]);
            rg[k] = make_int2(c, c); hg[k] = 1;
        } else {
            int j = c - n;
            int st = ready[j];
            if (st == 0 || st >= stamp) { ok = false; break; }
            lo[k] = mk3(bmin[3 * j], bmin[3 * j + 1], bmin[3 * j + 2]);
            hi[k] = mk3(bmax[3 * j], bmax[3 * j + 1], bmax[3 * j + 2]);
            rg[k] = range[j]; hg[k] = height[j];
        }
    }
    if (!ok) return true;
    V3 l = vmin(lo[0], lo[1]), h = vmax(hi[0], hi[1]);
    bmin[3 * i] = l.x; bmin[3 * i + 1] = l.y; bmin[3 * i + 2] = l.z;
    bmax[3 * i] = h.x; bmax[3 * i + 1] = h.y; bmax[3 * i + 2] = h.z;
    range[i] = make_int2(min(rg[0].x, rg[1].x), max(rg[0].y, rg[1].y));
    height[i] = 1 + max(hg[0], hg[1]);
    if (!(rg[0].y < rg[1].x)) scal[9] = 1;     // child0 must cover lower slots than child1
    if (i == 0) scal[10] = height[i];
    ready[i] = stamp;
    return false;
}
__global__ void __launch_bounds__(BLK) k_aabb_sweep(const float* __restrict__ verts, const int* __restrict__ leaf, const int2* __restrict__ child, int n, int stamp,
                                                    float* bmin, float* bmax, int* ready, int2* range, int* height, int* scal) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n - 1) return;
    if (aabb_try(i, verts, leaf, child, n, stamp, bmin, bmax, ready, range, height, scal)) atomicAdd(&scal[8], 1);
}
// small trees: every sweep in ONE block (a __syncthreads between sweeps instead of a launch and a host poll each).
// scal[8] = nodes still not ready at the end, scal[13] = sweeps used.
__global__ void __launch_bounds__(1024) k_aabb_all(const float* __restrict__ verts, const int* __restrict__ leaf, const int2* __restrict__ child, int n,
                                                   float* bmin, float* bmax, int* ready, int2* range, int* height, int* scal) {
    int remaining = 1, stamp = 0;
    while (remaining != 0 && stamp < 64) {
        stamp++;
        int mine = 0;
        for (int i = threadIdx.x; i < n - 1; i += 1024) mine += aabb_try(i, verts, leaf, child, n, stamp, bmin, bmax, ready, range, height, scal) ? 1 : 0;
        __threadfence_block();
        remaining = __syncthreads_count(mine != 0);
    }
    if (threadIdx.x == 0) { scal[8] = remaining; scal[13] = stamp; }
}

__global__ void __launch_bounds__(BLK) k_validate(const int* __restrict__ parentcnt, int n, int* scal) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= 2 * n - 1) return;
    int want = (i == n) ? 0 : 1;
    if (parentcnt[i] != want) scal[9] = 1;
}

// ---- traversal tree (ptb_traverse.cuh "traversal structure") ------------------------------------------------------------------------
// Per leaf slot: the triangle's bounds inflated by eps_T, the distance within which Face.intersect's f32 arithmetic can accept a
// point outside the triangle (derivation in DESIGN.md section 10):
//   eps_T = 256u * cond * (|u| + |v|) + 64u * (|v0|_inf + |u| + |v|),  cond = uu*vv / |D| = 1 / sin^2(angle(u, v)),  u = 2^-24
// valid while cond * (36 + 20 * max(|u|/|v|, |v|/|u|)) <= 1e6.  Flags (w of tlo): 1 = NEVER (D == 0 or not finite: s, t are
// inf / NaN, no ray is ever accepted), 2 = MUST (ill-conditioned: no usable bound -> the leaf STAYS IN THE TREE, bounded by its GATE
// box instead -- the reference box of its parent, which any accepted triangle's ray must pass anyway -- and exempt from distance
// culling), 4 = BIG (box surface above PTB_BIG_FRACTION = 1/4 of the scene's: hurts every ancestor's box -> always-test list while
// there is room; BIG triangles beyond PTB_LIST_CAP simply stay in the tree, which is still conservative, so the list cannot overflow).
#ifndef PTB_BIG_FRACTION
#define PTB_BIG_FRACTION (1.0f / 4.0f)      /* a triangle is BIG when its box surface exceeds this fraction of the scene's (walls: 1/3) */
#endif
__global__ void __launch_bounds__(BLK) k_tri_prep(const float* __restrict__ verts, const int* __restrict__ leaf, int n, const float* __restrict__ scal,
                                                  float4* __restrict__ tlo, float4* __restrict__ thi) {
    int s = blockIdx.x * BLK + threadIdx.x;
    if (s >= n) return;
    int f = leaf[s];
    V3 v0 = face_vertex(verts, f, 0), v1 = face_vertex(verts, f, 1), v2 = face_vertex(verts, f, 2);
    V3 u = v1 - v0, v = v2 - v0;
    float uu = dot(u, u), uv = dot(u, v), vv = dot(v, v);
    float D = uv * uv - uu * vv;
    V3 lo = vmin(vmin(v0, v1), v2), hi = vmax(vmax(v0, v1), v2);
    int flags = 0;
    float eps = 0.0f;
    if (!(fabsf(D) > 0.0f) || !(fabsf(D) <= 3e38f)) flags = PTB_TF_NEVER;
    else {
        float lu = sqrtf(uu), lv = sqrtf(vv);
        float cond = uu * vv / fabsf(D);
        float q = fmaxf(lu, lv) / fminf(lu, lv);
        if (!(cond * (36.0f + 20.0f * q) <= 1e6f)) flags = PTB_TF_MUST;
        else {
            const float U = 5.9604644775390625e-8f;
            float vmaxabs = fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fabsf(v0.z));
            eps = 1.01f * (256.0f * U * cond * (lu + lv) + 64.0f * U * (vmaxabs + lu + lv));
            if (!(eps <= 3e38f)) flags = PTB_TF_MUST;
        }
        // scene extent = extent of the triangle centres (lbvh.py:172-176), the only bounds known before the boxes are built
        V3 e = mk3(scal[3] - scal[0], scal[4] - scal[1], scal[5] - scal[2]), d = hi - lo;
        float scene = e.x * e.y + e.y * e.z + e.z * e.x, own = d.x * d.y + d.y * d.z + d.z * d.x;
        if (own > scene * PTB_BIG_FRACTION) flags |= PTB_TF_BIG;
    }
    tlo[s] = make_float4(__fadd_rd(lo.x, -eps), __fadd_rd(lo.y, -eps), __fadd_rd(lo.z, -eps), __int_as_float(flags));
    thi[s] = make_float4(__fadd_ru(hi.x, eps), __fadd_ru(hi.y, eps), __fadd_ru(hi.z, eps), 0.0f);
}
// always-test list, in slot order (one block): BIG slots while there is room (the rest stay in the tree).  scal[11] = entries.
__global__ void __launch_bounds__(1024) k_build_list(float4* __restrict__ tlo, int n, int cap, int* __restrict__ list, int* scal) {
    __shared__ int s_count, s_warp[32];
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    {
        const int want = PTB_TF_BIG;
        for (int base = 0; base < n; base += 1024) {
            int s = base + threadIdx.x;
            int fl = s < n ? __float_as_int(tlo[s].w) : 0;
            bool take = (fl & want) != 0 && (fl & (PTB_TF_NEVER | PTB_TF_LISTED)) == 0;
            unsigned m = __ballot_sync(0xffffffffu, take);
            int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            if (lane == 0) s_warp[w] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
            for (int k = 0; k < 32; k++) { int c = s_warp[k]; if (k < w) before += c; total += c; }
            int pos = s_count + before + __popc(m & ((1u << lane) - 1u));
            if (take && pos < cap) { list[pos] = s; float4 t = tlo[s]; t.w = __int_as_float(fl | PTB_TF_LISTED); tlo[s] = t; }
            __syncthreads();
            if (threadIdx.x == 0) s_count = min(cap, s_count + total);
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) scal[11] = s_count;
}
// slot_of[face] of a listed triangle carries its list index + 1 above the slot (PTB_SLOT_LIST_SHIFT, ptb_traverse.cuh); after k_pack_tris
__global__ void k_tag_listed(const int* __restrict__ list, const int* __restrict__ scal, const int* __restrict__ leaf, int* slot_of) {
    const int j = threadIdx.x;
    if (j < min(scal[11], PTB_LIST_CAP)) { const int s = list[j]; slot_of[leaf[s]] = s | ((j + 1) << PTB_SLOT_LIST_SHIFT); }
}
// traversal boxes, level by level (`ready` holds the sweep in which the reference box of a node was completed, so the children
// of a node of level `stamp` belong to earlier levels): union of the inflated bounds of the unlisted leaves below the node.
// Empty = (lo, hi) = (+1e30, -1e30).
__device__ __forceinline__ void tbox_node(int i, const int2* __restrict__ child, int n, const float* __restrict__ bmin, const float* __restrict__ bmax,
                                          const float4* __restrict__ tlo, const float4* __restrict__ thi, float4* nlo, float4* nhi) {
    int2 ch = child[i];
    V3 lo = v3s(1e30f), hi = v3s(-1e30f);
    int must = 0;          // an ill-conditioned leaf below: its depth obeys no bound, so nothing above it may be culled by distance
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int c = k == 0 ? ch.x : ch.y;
        float4 a, b;
        if (c < 0 || c >= 2 * n - 1) continue;         // corrupted hierarchy (reported by the build)
        if (c < n) {
            a = tlo[c]; b = thi[c];
            const int fl = __float_as_int(a.w);
            if (fl & (PTB_TF_NEVER | PTB_TF_LISTED)) continue;
            if (fl & PTB_TF_MUST) { must = 1; a = make_float4(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2], 0.0f); b = make_float4(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2], 0.0f); }
        } else { a = nlo[c - n]; b = nhi[c - n]; must |= __float_as_int(a.w); }
        lo = vmin(lo, mk3(a.x, a.y, a.z)); hi = vmax(hi, mk3(b.x, b.y, b.z));
    }
    nlo[i] = make_float4(lo.x, lo.y, lo.z, __int_as_float(must)); nhi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
}
__global__ void __launch_bounds__(BLK) k_tbox_level(const int2* __restrict__ child, const int* __restrict__ ready, int n, int stamp,
                                                    const float* __restrict__ bmin, const float* __restrict__ bmax,
                                                    const float4* __restrict__ tlo, const float4* __restrict__ thi, float4* nlo, float4* nhi) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n - 1 || ready[i] != stamp) return;
    tbox_node(i, child, n, bmin, bmax, tlo, thi, nlo, nhi);
}
__global__ void __launch_bounds__(1024) k_tbox_all(const int2* __restrict__ child, const int* __restrict__ ready, int n, const int* __restrict__ scal,
                                                   const float* __restrict__ bmin, const float* __restrict__ bmax,
                                                   const float4* __restrict__ tlo, const float4* __restrict__ thi, float4* nlo, float4* nhi) {
    const int levels = scal[13];
    for (int lvl = 1; lvl <= levels; lvl++) {
        for (int i = threadIdx.x; i < n - 1; i += 1024) if (ready[i] == lvl) tbox_node(i, child, n, bmin, bmax, tlo, thi, nlo, nhi);
        __threadfence_block();
        __syncthreads();
    }
}
// packed 64-byte traversal node: both children's traversal boxes and ids (-1 = nothing below: listed / never-hit leaf or an
// internal node with an empty box); per leaf slot: its gate (parent) and the gate's REFERENCE box (bmin/bmax of the parent).
__global__ void __launch_bounds__(BLK) k_pack_nodes(const int2* __restrict__ child, const float* __restrict__ bmin, const float* __restrict__ bmax, int n,
                                                    const float4* __restrict__ tlo, const float4* __restrict__ thi, const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                                                    Node64* nodes, int* __restrict__ gate, float4* __restrict__ gbox) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n - 1) return;
    int2 ch = child[i];
    float4 lo[2], hi[2]; int id[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int c = k == 0 ? ch.x : ch.y;
        id[k] = c;
        lo[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); hi[k] = lo[k];
        if (c < 0 || c >= 2 * n - 1) id[k] = -1;
        else if (c < n) {
            lo[k] = tlo[c]; hi[k] = thi[c];
            const int fl = __float_as_int(lo[k].w);
            if (fl & (PTB_TF_NEVER | PTB_TF_LISTED)) id[k] = -1;
            else if (fl & PTB_TF_MUST) {      // ill-conditioned: bounded by its gate box (this node's reference box), never culled by distance
                lo[k] = make_float4(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2], 0.0f); hi[k] = make_float4(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2], 0.0f);
                id[k] |= PTB_NODE_MUST;
            }
            gate[c] = i;
            gbox[2 * c] = make_float4(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2], 0.0f);
            gbox[2 * c + 1] = make_float4(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2], 0.0f);
        } else {
            lo[k] = nlo[c - n]; hi[k] = nhi[c - n];
            if (lo[k].x > hi[k].x) id[k] = -1;
            else if (__float_as_int(lo[k].w)) id[k] |= PTB_NODE_MUST;
        }
    }
    Node64 N;
    N.a = make_float4(lo[0].x, lo[0].y, lo[0].z, __int_as_float(id[0]));
    N.b = make_float4(hi[0].x, hi[0].y, hi[0].z, __int_as_float(id[1]));
    N.c = make_float4(lo[1].x, lo[1].y, lo[1].z, 0.0f);
    N.d = make_float4(hi[1].x, hi[1].y, hi[1].z, 0.0f);
    nodes[i] = N;
}

// ---- a better traversal tree for small scenes: PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) ---------------------
// The traversal tree only has to be conservative (ptb_traverse.cuh): any binary tree over the unlisted triangles whose node boxes
// contain the inflated bounds below them gives the reference's answer once the winner's gate is checked.  The LBVH topology splits at
// the spatial medians of the whole scene; PLOC merges, bottom-up along the Morton order the slots already have, every pair of clusters
// that are each other's nearest neighbour (smallest surface area of the union, search radius PTB_PLOC_R) -- lower SAH cost, shallower.
// One block (trees up to PTB_SMALL_TREE internal nodes); deterministic.  Nodes are numbered downward from m0 - 2 in creation order,
// so the root (created last) is node 0.  scal[15] = height of the tree (-1: not built).
#ifndef PTB_PLOC_R
#define PTB_PLOC_R 8           /* search radius along the Morton order: 16 builds the same trees here, 32 slightly better ones for a slower build */
#endif
struct PlocBufs { int* id[2]; float4* lo[2]; float4* hi[2]; int* depth[2]; int* nn; };
__device__ __forceinline__ float ploc_cost(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}
// exclusive scan over the block of one int per thread (two 16-bit counters packed); *total = sum
__device__ __forceinline__ int ploc_scan(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}
__global__ void __launch_bounds__(1024) k_ploc_small(const float4* __restrict__ tlo, const float4* __restrict__ thi, const float4* __restrict__ gbox, int n,
                                                     PlocBufs B, Node64* __restrict__ nodes, int* __restrict__ scal) {
    __shared__ int s_warp[33];
    const int tid = threadIdx.x;
    // clusters = the unlisted, testable leaf slots, in slot (Morton) order; an ill-conditioned triangle is bounded by its gate box
    int chunk = (n + 1023) / 1024, first = min(n, tid * chunk), last = min(n, first + chunk), cnt = 0, total;
    for (int s = first; s < last; s++) cnt += (__float_as_int(tlo[s].w) & (PTB_TF_NEVER | PTB_TF_LISTED)) == 0;
    int pos = ploc_scan(cnt, s_warp, &total);
    for (int s = first; s < last; s++) {
        const int fl = __float_as_int(tlo[s].w);
        if (fl & (PTB_TF_NEVER | PTB_TF_LISTED)) continue;
        const bool must = (fl & PTB_TF_MUST) != 0;
        B.id[0][pos] = must ? (s | PTB_NODE_MUST) : s;
        B.lo[0][pos] = must ? gbox[2 * s] : tlo[s];
        B.hi[0][pos] = must ? gbox[2 * s + 1] : thi[s];
        B.depth[0][pos] = 0;
        pos++;
    }
    int m = total;
    const int m0 = m;
    if (m0 < 2) { if (tid == 0) scal[15] = -1; return; }
    int created = 0, cur = 0;
    __syncthreads();
    while (m > 1) {
        const int* id = B.id[cur]; const float4* lo = B.lo[cur]; const float4* hi = B.hi[cur]; const int* dep = B.depth[cur];
        for (int i = tid; i < m; i += 1024) {
            const float4 alo = lo[i], ahi = hi[i];
            float best = 0.0f; int bj = -1;
            const int j0 = max(0, i - PTB_PLOC_R), j1 = min(m - 1, i + PTB_PLOC_R);
            for (int j = j0; j <= j1; j++) {
                if (j == i) continue;
                const float c = ploc_cost(alo, ahi, lo[j], hi[j]);
                if (bj < 0 || c < best) { best = c; bj = j; }
            }
            B.nn[i] = bj;
        }
        __syncthreads();
        // mutual nearest neighbours merge (the lower index keeps the slot); counts: clusters kept (low half), nodes created (high half)
        chunk = (m + 1023) / 1024; first = min(m, tid * chunk); last = min(m, first + chunk); cnt = 0;
        for (int i = first; i < last; i++) {
            const int j = B.nn[i];
            const bool mutual = B.nn[j] == i;
            if (mutual && j < i) continue;
            cnt += 1 + ((mutual && i < j) ? 0x10000 : 0);
        }
        const int excl = ploc_scan(cnt, s_warp, &total);
        int kpos = excl & 0xffff, mpos = excl >> 16;
        for (int i = first; i < last; i++) {
            const int j = B.nn[i];
            const bool mutual = B.nn[j] == i;
            if (mutual && j < i) continue;
            int cid = id[i], cdep = dep[i];
            float4 clo = lo[i], chi = hi[i];
            if (mutual) {
                const int node = (m0 - 2) - (created + mpos++);
                const float4 blo = lo[j], bhi = hi[j];
                const int idj = id[j];
                Node64 N;
                N.a = make_float4(clo.x, clo.y, clo.z, __int_as_float(cid));
                N.b = make_float4(chi.x, chi.y, chi.z, __int_as_float(idj));
                N.c = make_float4(blo.x, blo.y, blo.z, 0.0f);
                N.d = make_float4(bhi.x, bhi.y, bhi.z, 0.0f);
                nodes[node] = N;
                clo = make_float4(fminf(clo.x, blo.x), fminf(clo.y, blo.y), fminf(clo.z, blo.z), 0.0f);
                chi = make_float4(fmaxf(chi.x, bhi.x), fmaxf(chi.y, bhi.y), fmaxf(chi.z, bhi.z), 0.0f);
                cid = (n + node) | ((cid | idj) & PTB_NODE_MUST);
                cdep = max(cdep, dep[j]) + 1;
            }
            B.id[cur ^ 1][kpos] = cid; B.lo[cur ^ 1][kpos] = clo; B.hi[cur ^ 1][kpos] = chi; B.depth[cur ^ 1][kpos] = cdep;
            kpos++;
        }
        created += total >> 16;
        m = total & 0xffff;
        cur ^= 1;
        __syncthreads();
    }
    if (tid == 0) scal[15] = B.depth[cur][0];
}

// ---- the same clustering for trees of any size, one kernel per step of a round (round 2) --------------------------------------------------
// Scenes beyond PTB_SMALL_TREE kept the LBVH topology in round 1.  The build of a 1 M-triangle scene is amortised over a render of
// seconds, so the traversal tree is now rebuilt by PLOC at every size: k_plocb_init compacts the unlisted, testable leaf slots into the
// cluster arrays (positions from a device-wide exclusive scan); every round runs k_plocb_nn (nearest neighbour by union surface area
// within +-radius along the current order), k_plocb_flags (who survives / which pairs merge; mutual nearest neighbours merge, the lower
// index keeps the slot), two scans, and k_plocb_merge (writes the merged clusters and their Node64).  Deterministic: ties go to the
// lower index, node numbers are assigned by the scan.  Node (m0 - 2) - k is the k-th node created, so the root is node 0.
// Device scalars (ints): S[0] = clusters in the current round, S[1] = nodes created so far, S[2] = m0, S[3] = height of the last cluster.
__global__ void __launch_bounds__(BLK) k_plocb_flags0(const float4* __restrict__ tlo, int n, int* __restrict__ flag) {
    int s = blockIdx.x * BLK + threadIdx.x;
    if (s < n) flag[s] = (__float_as_int(tlo[s].w) & (PTB_TF_NEVER | PTB_TF_LISTED)) == 0;
}
__global__ void __launch_bounds__(BLK) k_plocb_init(const float4* __restrict__ tlo, const float4* __restrict__ thi, const float4* __restrict__ gbox, int n,
                                                    const int* __restrict__ flag, const int* __restrict__ pos, int* __restrict__ id, float4* __restrict__ lo, float4* __restrict__ hi,
                                                    int* __restrict__ depth, int* __restrict__ S) {
    int s = blockIdx.x * BLK + threadIdx.x;
    if (s >= n) return;
    if (s == n - 1) { const int m0 = pos[s] + flag[s]; S[0] = m0; S[1] = 0; S[2] = m0; S[3] = 0; }
    if (!flag[s]) return;
    const int fl = __float_as_int(tlo[s].w);
    const bool must = (fl & PTB_TF_MUST) != 0;          // ill-conditioned: bounded by its gate box, never culled by distance
    const int k = pos[s];
    id[k] = must ? (s | PTB_NODE_MUST) : s;
    lo[k] = must ? gbox[2 * s] : tlo[s];
    hi[k] = must ? gbox[2 * s + 1] : thi[s];
    depth[k] = 0;
}
__global__ void __launch_bounds__(BLK) k_plocb_nn(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ S, int radius, int* __restrict__ nn) {
    const int m = S[0];
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m) return;
    const float4 alo = lo[i], ahi = hi[i];
    float best = 0.0f; int bj = -1;
    const int j0 = max(0, i - radius), j1 = min(m - 1, i + radius);
    for (int j = j0; j <= j1; j++) {
        if (j == i) continue;
        const float c = ploc_cost(alo, ahi, __ldg(&lo[j]), __ldg(&hi[j]));
        if (bj < 0 || c < best) { best = c; bj = j; }
    }
    nn[i] = bj;
}
__global__ void __launch_bounds__(BLK) k_plocb_flags(const int* __restrict__ nn, const int* __restrict__ S, int* __restrict__ keep, int* __restrict__ make) {
    const int m = S[0];
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    keep[i] = !(mutual && j < i);
    make[i] = mutual && i < j;
}
__global__ void __launch_bounds__(BLK) k_plocb_merge(const int* __restrict__ nn, const int* __restrict__ keep, const int* __restrict__ kpos, const int* __restrict__ mpos,
                                                     const int* __restrict__ id, const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ depth,
                                                     int* __restrict__ id2, float4* __restrict__ lo2, float4* __restrict__ hi2, int* __restrict__ depth2,
                                                     int n, const int* __restrict__ S, Node64* __restrict__ nodes) {
    const int m = S[0], created = S[1], m0 = S[2];
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m || !keep[i]) return;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;           // keep[i] && mutual  =>  i < j
    int cid = id[i], cdep = depth[i];
    float4 clo = lo[i], chi = hi[i];
    if (mutual) {
        const int node = (m0 - 2) - (created + mpos[i]);
        const float4 blo = lo[j], bhi = hi[j];
        const int idj = id[j];
        Node64 N;
        N.a = make_float4(clo.x, clo.y, clo.z, __int_as_float(cid));
        N.b = make_float4(chi.x, chi.y, chi.z, __int_as_float(idj));
        N.c = make_float4(blo.x, blo.y, blo.z, 0.0f);
        N.d = make_float4(bhi.x, bhi.y, bhi.z, 0.0f);
        nodes[node] = N;
        clo = make_float4(fminf(clo.x, blo.x), fminf(clo.y, blo.y), fminf(clo.z, blo.z), 0.0f);
        chi = make_float4(fmaxf(chi.x, bhi.x), fmaxf(chi.y, bhi.y), fmaxf(chi.z, bhi.z), 0.0f);
        cid = (n + node) | ((cid | idj) & PTB_NODE_MUST);
        cdep = max(cdep, depth[j]) + 1;
    }
    const int k = kpos[i];
    id2[k] = cid; lo2[k] = clo; hi2[k] = chi; depth2[k] = cdep;
}
// after the merge of a round: new cluster count, nodes created so far, height of the (eventually only) cluster 0
__global__ void k_plocb_next(const int* __restrict__ keep, const int* __restrict__ make, const int* __restrict__ kpos, const int* __restrict__ mpos,
                             const int* __restrict__ depth2, int* S, int* __restrict__ scal) {
    const int m = S[0];
    const int m2 = kpos[m - 1] + keep[m - 1];
    S[1] += mpos[m - 1] + make[m - 1];
    S[0] = m2;
    S[3] = depth2[0];
    if (m2 == 1) scal[15] = depth2[0];
}

// Node32 (ptb_traverse.cuh): the boxes of a packed node on the 15-bit grid, rounded outward.  The candidate from the f32 estimate is
// corrected against the exact value of the plane it decodes to (base + (1 + q/32768) * ext: a 24-bit plus a 16-bit number of
// nearby exponents, exact in double), so lo' <= lo and hi' >= hi hold as real numbers whatever the rounding of the estimate.
// scal[14] counts planes outside the grid (cannot happen: the grid covers both root boxes; checked by the host all the same).
struct QuantGrid { float base[3], ext[3]; };
__device__ __forceinline__ unsigned quant_plane(float x, float base, float ext, bool up, int* bad) {
    const double b = (double)base, e = (double)ext, xd = (double)x;
    double t = ((xd - b) / e - 1.0) * 32768.0;
    int q = up ? (int)ceil(t) : (int)floor(t);
    q = q < 0 ? 0 : (q > 32767 ? 32767 : q);
    for (int it = 0; it < 4; it++) {
        const double pl = b + (1.0 + (double)q * (1.0 / 32768.0)) * e;
        if (up ? pl >= xd : pl <= xd) break;
        if (up ? q == 32767 : q == 0) break;
        q += up ? 1 : -1;
    }
    const double pl = b + (1.0 + (double)q * (1.0 / 32768.0)) * e;
    if (!(up ? pl >= xd : pl <= xd)) *bad = 1;
    return 0x8000u | (unsigned)q;
}
__global__ void __launch_bounds__(BLK) k_quant_nodes(const Node64* __restrict__ nodes, int n, QuantGrid g, uint4* __restrict__ qnodes, int* __restrict__ scal) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n - 1) return;
    const Node64 N = nodes[i];
    const int id0 = __float_as_int(N.a.w), id1 = __float_as_int(N.b.w);
    const float lo[2][3] = {{N.a.x, N.a.y, N.a.z}, {N.c.x, N.c.y, N.c.z}}, hi[2][3] = {{N.b.x, N.b.y, N.b.z}, {N.d.x, N.d.y, N.d.z}};
    unsigned f[12];
    int bad = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const bool used = (k == 0 ? id0 : id1) >= 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int dummy = 0;
            f[6 * k + a] = quant_plane(used ? lo[k][a] : g.base[a] + g.ext[a], g.base[a], g.ext[a], false, used ? &bad : &dummy);
            f[6 * k + 3 + a] = quant_plane(used ? hi[k][a] : g.base[a] + g.ext[a], g.base[a], g.ext[a], true, used ? &bad : &dummy);
        }
    }
    if (bad) atomicAdd(&scal[14], 1);
    // one word per axis and box: lo in the low half, hi in the high half (the traversal picks the entry / exit plane of an axis out of
    // ONE word with a per-ray PRMT selector, ptb_traverse.cuh slab_quant)
    qnodes[2 * i] = make_uint4(f[0] | (f[3] << 16), f[1] | (f[4] << 16), f[2] | (f[5] << 16), f[6] | (f[9] << 16));
    qnodes[2 * i + 1] = make_uint4(f[7] | (f[10] << 16), f[8] | (f[11] << 16), (unsigned)id0, (unsigned)id1);
}

// 4-wide quantised node of binary node i (64 B): its two children, each internal one replaced by ITS two children -- up to four
// (box, id) entries on the same 15-bit grid as Node32 (three words per box, one per axis: lo in the low half, hi in the high half), then
// the four ids (-1: none).  Every binary node gets one (index = binary index); a traversal from the root only ever reaches those at
// even depth, so the hot footprint equals the binary quantised array's.
__global__ void __launch_bounds__(BLK) k_wide4_nodes(const Node64* __restrict__ nodes, int n, QuantGrid g, uint4* __restrict__ wnodes, int* __restrict__ scal) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= n - 1) return;
    const Node64 N = nodes[i];
    float lo[4][3], hi[4][3]; int id[4]; int m = 0;
    const int cid[2] = {__float_as_int(N.a.w), __float_as_int(N.b.w)};
    const float4 clo[2] = {N.a, N.c}, chi[2] = {N.b, N.d};
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (cid[k] < 0) continue;
        const int c = cid[k] & PTB_NODE_ID;
        if (c < n) {                                       // leaf: itself
            lo[m][0] = clo[k].x; lo[m][1] = clo[k].y; lo[m][2] = clo[k].z; hi[m][0] = chi[k].x; hi[m][1] = chi[k].y; hi[m][2] = chi[k].z; id[m++] = cid[k];
        } else {                                           // internal: its children (their own ids carry their own flags)
            const Node64 M = nodes[c - n];
            const int gid[2] = {__float_as_int(M.a.w), __float_as_int(M.b.w)};
            const float4 glo[2] = {M.a, M.c}, ghi[2] = {M.b, M.d};
            int took = 0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (gid[j] < 0) continue;
                lo[m][0] = glo[j].x; lo[m][1] = glo[j].y; lo[m][2] = glo[j].z; hi[m][0] = ghi[j].x; hi[m][1] = ghi[j].y; hi[m][2] = ghi[j].z; id[m++] = gid[j]; took++;
            }
            (void)took;
        }
    }
    unsigned w[16];
    int bad = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool used = k < m;
        unsigned f[6];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int dummy = 0;
            f[a] = quant_plane(used ? lo[k][a] : g.base[a] + g.ext[a], g.base[a], g.ext[a], false, used ? &bad : &dummy);
            f[3 + a] = quant_plane(used ? hi[k][a] : g.base[a] + g.ext[a], g.base[a], g.ext[a], true, used ? &bad : &dummy);
        }
        w[3 * k] = f[0] | (f[3] << 16); w[3 * k + 1] = f[1] | (f[4] << 16); w[3 * k + 2] = f[2] | (f[5] << 16);
        w[12 + k] = used ? (unsigned)id[k] : 0xFFFFFFFFu;
    }
    if (bad) atomicAdd(&scal[14], 1);
#pragma unroll
    for (int q = 0; q < 4; q++) wnodes[4 * (size_t)i + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// packed 64-byte triangle by leaf slot; u, v, n, uu, uv, vv, D are the f32 expressions of geometries.py:121-141
__global__ void __launch_bounds__(BLK) k_pack_tris(const float* __restrict__ verts, const int* __restrict__ leaf, int n, Tri64* tris, int* slot_of) {
    int s = blockIdx.x * BLK + threadIdx.x;
    if (s >= n) return;
    int f = leaf[s];
    V3 v0 = face_vertex(verts, f, 0), v1 = face_vertex(verts, f, 1), v2 = face_vertex(verts, f, 2);
    V3 u = v1 - v0, v = v2 - v0;
    V3 nrm = cross(u, v);
    float uu = dot(u, u), uv = dot(u, v), vv = dot(v, v);
    float D = uv * uv - uu * vv;
    Tri64 T;
    T.a = make_float4(v0.x, v0.y, v0.z, 1.0f / D);
    T.b = make_float4(u.x, u.y, u.z, uu);
    T.c = make_float4(v.x, v.y, v.z, uv);
    T.d = make_float4(nrm.x, nrm.y, nrm.z, vv);
    tris[s] = T;
    slot_of[f] = s;      // leaf[] is a permutation of the faces, so this is its inverse
}

inline int nblk(int n) { return (n + BLK - 1) / BLK; }

}  // namespace

// PLOC over a tree of any size (kernels k_plocb_*).  The round loop reads the cluster count back once per round: the build of a big
// scene is paid once per render, not per frame.  Leaves scal[15] = height of the tree, or -1 if it did not converge (LBVH topology stays).
static int ptb_ploc_big(ptb_ctx* c, int n) {
    cudaStream_t st = c->stream;
    if (c->pl_cap < n) {      // grow-only cluster buffers and the second node array
        void** bufs[] = {(void**)&c->d_pl_id[0], (void**)&c->d_pl_id[1], (void**)&c->d_pl_depth[0], (void**)&c->d_pl_depth[1], (void**)&c->d_pl_nn, (void**)&c->d_pl_keep, (void**)&c->d_pl_make,
                         (void**)&c->d_pl_kpos, (void**)&c->d_pl_mpos, (void**)&c->d_pl_lo[0], (void**)&c->d_pl_lo[1], (void**)&c->d_pl_hi[0], (void**)&c->d_pl_hi[1], (void**)&c->d_nodes2};
        const size_t elem[] = {4, 4, 4, 4, 4, 4, 4, 4, 4, 16, 16, 16, 16, sizeof(Node64)};
        PTB_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 14; k++) { cudaFree(*bufs[k]); *bufs[k] = nullptr; PTB_CUDA(cudaMalloc(bufs[k], elem[k] * (size_t)(n + 8))); }
        c->pl_cap = n;
    }
    if (!c->d_pl_S) PTB_CUDA(cudaMalloc((void**)&c->d_pl_S, 16));
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, c->d_pl_keep, c->d_pl_kpos, n, st);
    if (need > c->sort_tmp_bytes) {
        PTB_CUDA(cudaStreamSynchronize(st));
        if (c->d_sort_tmp) cudaFree(c->d_sort_tmp);
        PTB_CUDA(cudaMalloc(&c->d_sort_tmp, need));
        c->sort_tmp_bytes = need;
    }
    size_t tmp = c->sort_tmp_bytes;
    PTB_CUDA(cudaMemsetAsync(c->d_nodes2, 0xFF, sizeof(Node64) * (size_t)(n - 1), st));      // ids -1: slots the PLOC tree leaves unused
    k_plocb_flags0<<<nblk(n), BLK, 0, st>>>(c->d_tlo, n, c->d_pl_keep);
    PTB_CUDA(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp, tmp, c->d_pl_keep, c->d_pl_kpos, n, st));
    k_plocb_init<<<nblk(n), BLK, 0, st>>>(c->d_tlo, c->d_thi, c->d_gbox, n, c->d_pl_keep, c->d_pl_kpos, c->d_pl_id[0], c->d_pl_lo[0], c->d_pl_hi[0], c->d_pl_depth[0], c->d_pl_S);
    c->launches += 3;
    int hS[4] = {0, 0, 0, 0};
    PTB_CUDA(cudaMemcpyAsync(hS, c->d_pl_S, sizeof hS, cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
    int m = hS[0], cur = 0, rounds = 0;
    if (m < 2) return 0;                                   // scal[15] stays -1: nothing to cluster
    for (; m > 1 && rounds < 256; rounds++) {
        k_plocb_nn<<<nblk(m), BLK, 0, st>>>(c->d_pl_lo[cur], c->d_pl_hi[cur], c->d_pl_S, c->ploc_radius, c->d_pl_nn);
        k_plocb_flags<<<nblk(m), BLK, 0, st>>>(c->d_pl_nn, c->d_pl_S, c->d_pl_keep, c->d_pl_make);
        tmp = c->sort_tmp_bytes;
        PTB_CUDA(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp, tmp, c->d_pl_keep, c->d_pl_kpos, m, st));
        tmp = c->sort_tmp_bytes;
        PTB_CUDA(cub::DeviceScan::ExclusiveSum(c->d_sort_tmp, tmp, c->d_pl_make, c->d_pl_mpos, m, st));
        k_plocb_merge<<<nblk(m), BLK, 0, st>>>(c->d_pl_nn, c->d_pl_keep, c->d_pl_kpos, c->d_pl_mpos, c->d_pl_id[cur], c->d_pl_lo[cur], c->d_pl_hi[cur], c->d_pl_depth[cur],
                                               c->d_pl_id[cur ^ 1], c->d_pl_lo[cur ^ 1], c->d_pl_hi[cur ^ 1], c->d_pl_depth[cur ^ 1], n, c->d_pl_S, c->d_nodes2);
        k_plocb_next<<<1, 1, 0, st>>>(c->d_pl_keep, c->d_pl_make, c->d_pl_kpos, c->d_pl_mpos, c->d_pl_depth[cur ^ 1], c->d_pl_S, c->d_scalars);
        c->launches += 6;
        PTB_CUDA(cudaMemcpyAsync(hS, c->d_pl_S, sizeof hS, cudaMemcpyDeviceToHost, st));
        PTB_CUDA(cudaStreamSynchronize(st));
        if (hS[0] >= m) break;                             // no pair merged: cannot happen (the global minimum is mutual), guards the loop
        m = hS[0];
        cur ^= 1;
    }
    c->ploc_rounds = rounds;
    return 0;
}

int ptb_lbvh_build(ptb_ctx* c) {
    const int n = c->nfaces;
    cudaStream_t st = c->stream;
    c->tree_n = n;
    c->tree_info = ptb_tree_info{};
    c->tree_info.n = n;
    struct Events {          // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
    PTB_CUDA(cudaEventCreate(&ev.a)); PTB_CUDA(cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    PTB_CUDA(cudaEventRecord(e0, st));
    float* scal = (float*)c->d_scalars;
    int nint = n > 1 ? n - 1 : 1;
    k_init_scalars<<<1, 1, 0, st>>>(scal);
    PTB_CUDA(cudaMemsetAsync(c->d_child, 0, sizeof(int2) * nint, st));
    PTB_CUDA(cudaMemsetAsync(c->d_bmin, 0, sizeof(float) * 3 * nint, st));
    PTB_CUDA(cudaMemsetAsync(c->d_bmax, 0, sizeof(float) * 3 * nint, st));
    PTB_CUDA(cudaMemsetAsync(c->d_ready, 0, sizeof(int) * nint, st));
    PTB_CUDA(cudaMemsetAsync(c->d_parentcnt, 0, sizeof(int) * 2 * (size_t)(n > 0 ? n : 1), st));
    int sweeps = 1, valid = 0, depth = 0;
    if (n > 0) {
        int grid = nblk(n) < c->sm_count * 8 ? nblk(n) : c->sm_count * 8;
        k_center_bounds<<<grid, BLK, 0, st>>>(c->d_verts, n, scal);
        k_morton<<<nblk(n), BLK, 0, st>>>(c->d_verts, n, scal, c->d_mc_tmp, c->d_id_tmp);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, c->d_mc_tmp, c->d_mc, c->d_id_tmp, c->d_id, n, 0, 30, st);
        if (need > c->sort_tmp_bytes) {
            if (c->d_sort_tmp) cudaFree(c->d_sort_tmp);
            PTB_CUDA(cudaMalloc(&c->d_sort_tmp, need));
            c->sort_tmp_bytes = need;
        }
        PTB_CUDA(cub::DeviceRadixSort::SortPairs(c->d_sort_tmp, need, c->d_mc_tmp, c->d_mc, c->d_id_tmp, c->d_id, n, 0, 30, st));
        k_hierarchy<<<nblk(n), BLK, 0, st>>>(c->d_mc, c->d_id, n, c->d_leaf, c->d_child, c->d_parentcnt);
        c->launches += 5;
    }
    if (n > 1) {
        // genAABBs lbvh.py:251-261: sweep until every node is ready, give up after 64
        int first_check = 1;
        while ((1 << first_check) < n) first_check++;   // a tree over n leaves is at least log2(n) high
        int remaining = -1;
        int h_scal[16];
        const bool small = n - 1 <= PTB_SMALL_TREE;
        if (small) {
            k_aabb_all<<<1, 1024, 0, st>>>(c->d_verts, c->d_leaf, c->d_child, n, c->d_bmin, c->d_bmax, c->d_ready, c->d_range, c->d_height, c->d_scalars);
            c->launches++;
            remaining = 0;      // read back with the final synchronisation below
        } else
        for (sweeps = 1; sweeps <= 64; sweeps++) {
            PTB_CUDA(cudaMemsetAsync(&c->d_scalars[8], 0, sizeof(int), st));
            k_aabb_sweep<<<nblk(n - 1), BLK, 0, st>>>(c->d_verts, c->d_leaf, c->d_child, n, sweeps, c->d_bmin, c->d_bmax, c->d_ready, c->d_range, c->d_height, c->d_scalars);
            c->launches++;
            if (sweeps >= first_check - 1 || sweeps == 64) {
                PTB_CUDA(cudaMemcpyAsync(h_scal, c->d_scalars, sizeof h_scal, cudaMemcpyDeviceToHost, st));
                PTB_CUDA(cudaStreamSynchronize(st));
                remaining = h_scal[8];
                if (remaining == 0) break;
            }
        }
        if (remaining != 0) {
            c->tree_info.aabb_sweeps = sweeps;
            ptb_set_error("AABB step never stop! hierarchy corrupted?");
            return 2;
        }
        k_validate<<<nblk(2 * n - 1), BLK, 0, st>>>(c->d_parentcnt, n, c->d_scalars);
        // traversal structure: inflated leaf bounds, always-test list, pruned traversal boxes, packed nodes
        k_tri_prep<<<nblk(n), BLK, 0, st>>>(c->d_verts, c->d_leaf, n, scal, c->d_tlo, c->d_thi);
        k_build_list<<<1, 1024, 0, st>>>(c->d_tlo, n, PTB_LIST_CAP, c->d_list, c->d_scalars);
        if (small) k_tbox_all<<<1, 1024, 0, st>>>(c->d_child, c->d_ready, n, c->d_scalars, c->d_bmin, c->d_bmax, c->d_tlo, c->d_thi, c->d_nlo, c->d_nhi);
        else for (int lvl = 1; lvl <= sweeps; lvl++) k_tbox_level<<<nblk(n - 1), BLK, 0, st>>>(c->d_child, c->d_ready, n, lvl, c->d_bmin, c->d_bmax, c->d_tlo, c->d_thi, c->d_nlo, c->d_nhi);
        k_pack_nodes<<<nblk(n - 1), BLK, 0, st>>>(c->d_child, c->d_bmin, c->d_bmax, n, c->d_tlo, c->d_thi, c->d_nlo, c->d_nhi, c->d_nodes, c->d_gate, c->d_gbox);
        c->launches += small ? 5 : 4 + sweeps;
        if (!small && c->use_ploc && c->ploc_big) {
            if (ptb_ploc_big(c, n)) return 1;
        }
        if (small && c->use_ploc) {
            PlocBufs B;
            for (int k = 0; k < 2; k++) { B.id[k] = c->d_pl_id[k]; B.lo[k] = c->d_pl_lo[k]; B.hi[k] = c->d_pl_hi[k]; B.depth[k] = c->d_pl_depth[k]; }
            B.nn = c->d_pl_nn;
            PTB_CUDA(cudaMemsetAsync(c->d_nodes2, 0xFF, sizeof(Node64) * (size_t)(n - 1), st));     // ids -1: slots PLOC leaves unused
            k_ploc_small<<<1, 1024, 0, st>>>(c->d_tlo, c->d_thi, c->d_gbox, n, B, c->d_nodes2, c->d_scalars);
            c->launches++;
        }
    }
    if (n > 0) { k_pack_tris<<<nblk(n), BLK, 0, st>>>(c->d_verts, c->d_leaf, n, c->d_tris, c->d_slot_of); c->launches++; }
    if (n > 1) { k_tag_listed<<<1, PTB_LIST_CAP, 0, st>>>(c->d_list, c->d_scalars, c->d_leaf, c->d_slot_of); c->launches++; }
    PTB_CUDA(cudaEventRecord(e1, st));
    {
        int h_scal[16];
        PTB_CUDA(cudaMemcpyAsync(h_scal, c->d_scalars, sizeof h_scal, cudaMemcpyDeviceToHost, st));
        float h_root[6] = {0, 0, 0, 0, 0, 0};
        float h_troot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (n > 1) {
            PTB_CUDA(cudaMemcpyAsync(h_root, c->d_bmin, 12, cudaMemcpyDeviceToHost, st));
            PTB_CUDA(cudaMemcpyAsync(h_root + 3, c->d_bmax, 12, cudaMemcpyDeviceToHost, st));
            PTB_CUDA(cudaMemcpyAsync(h_troot, c->d_nlo, 16, cudaMemcpyDeviceToHost, st));
            PTB_CUDA(cudaMemcpyAsync(h_troot + 4, c->d_nhi, 16, cudaMemcpyDeviceToHost, st));
        }
        PTB_CUDA(cudaStreamSynchronize(st));
        c->scene_abs = 0.0f;
        for (int k = 0; k < 6; k++) c->scene_abs = fmaxf(c->scene_abs, fabsf(h_root[k]));
        c->root_must = 0;
        if (h_troot[0] <= h_troot[4]) { int m; memcpy(&m, &h_troot[3], 4); c->root_must = m != 0; }
        if (h_troot[0] <= h_troot[4]) for (int k = 0; k < 3; k++) c->scene_abs = fmaxf(c->scene_abs, fmaxf(fabsf(h_troot[k]), fabsf(h_troot[4 + k])));
        if (n > 1 && n - 1 <= PTB_SMALL_TREE) {
            sweeps = h_scal[13];
            if (h_scal[8] != 0) {
                c->tree_info.aabb_sweeps = sweeps;
                ptb_set_error("AABB step never stop! hierarchy corrupted?");
                return 2;
            }
        }
        valid = (n > 1) && h_scal[9] == 0;
        depth = valid ? h_scal[10] : 0;
        c->list_n = n > 1 ? h_scal[11] : 0;
        for (int k = 0; k < 3; k++) { c->root_lo[k] = h_root[k]; c->root_hi[k] = h_root[3 + k]; }
        // the PLOC tree replaces the LBVH topology for traversal when it was built and fits the traversal stack
        c->trav_depth = (n > 1 && c->use_ploc && (n - 1 <= PTB_SMALL_TREE || c->ploc_big)) ? h_scal[15] : -1;
        c->d_nodes_active = (c->trav_depth >= 1 && c->trav_depth <= PTB_STACK) ? c->d_nodes2 : c->d_nodes;
        // quantised nodes for a tree too big to be resident as 64-byte nodes (wavefront.cu launch_trace_io decides the same way)
        if (valid && ptb_tree_mode(c, n) != PTB_TREE_RESIDENT) {
            QuantGrid g;
            bool finite = true;
            for (int k = 0; k < 3; k++) {
                float lo = h_root[k], hi = h_root[3 + k];
                if (h_troot[0] <= h_troot[4]) { lo = fminf(lo, h_troot[k]); hi = fmaxf(hi, h_troot[4 + k]); }
                finite = finite && std::isfinite(lo) && std::isfinite(hi) && lo <= hi;
                // plane = base + v * ext, v in [1, 2 - 2^-15]: ext a power of two, base rounded down, widened until it covers [lo, hi]
                double ext = ldexp(1.0, (int)ceil(log2(fmax((double)hi - (double)lo, 1e-30) * 1.01)));
                float base = 0.0f;
                for (int it = 0; it < 8; it++, ext *= 2.0) {
                    const double b = (double)lo - ext;
                    base = (float)b;
                    if ((double)base > b) base = nextafterf(base, -INFINITY);
                    if ((double)base + ext <= (double)lo && (double)base + (2.0 - 1.0 / 32768.0) * ext >= (double)hi) break;
                }
                g.base[k] = base; g.ext[k] = (float)ext;
                c->qbase[k] = base; c->qext[k] = (float)ext; c->qinv[k] = (float)(1.0 / ext);
                finite = finite && std::isfinite(c->qext[k]) && std::isfinite(c->qinv[k]) && c->qinv[k] > 0.0f;
            }
            int bad = 1;
            c->wnodes_ok = false;
            if (finite) {
                k_quant_nodes<<<nblk(n - 1), BLK, 0, st>>>(c->d_nodes_active, n, g, c->d_qnodes, c->d_scalars);
                c->launches++;
                // 4-wide nodes for a tree walked out of global memory; three stack pushes per step: height <= 40 keeps the 64-entry stack safe
                const int tdepth = c->d_nodes_active == c->d_nodes2 ? c->trav_depth : depth;
                if (c->wide4 && ptb_tree_mode(c, n) == PTB_TREE_GLOBAL && tdepth >= 1 && tdepth <= 40) {
                    if (c->wnodes_cap < n) {
                        PTB_CUDA(cudaStreamSynchronize(st));
                        cudaFree(c->d_wnodes); c->d_wnodes = nullptr;
                        PTB_CUDA(cudaMalloc((void**)&c->d_wnodes, sizeof(uint4) * 4 * (size_t)n));
                        c->wnodes_cap = n;
                    }
                    k_wide4_nodes<<<nblk(n - 1), BLK, 0, st>>>(c->d_nodes_active, n, g, c->d_wnodes, c->d_scalars);
                    c->launches++;
                    c->wnodes_ok = true;
                }
                PTB_CUDA(cudaMemcpyAsync(&bad, &c->d_scalars[14], sizeof(int), cudaMemcpyDeviceToHost, st));
                PTB_CUDA(cudaStreamSynchronize(st));
            }
            if (bad) valid = 0;        // no usable grid (non-finite bounds): the literal traversal takes over
        }
    }
    PTB_CUDA(cudaGetLastError());
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    c->tree_info.aabb_sweeps = sweeps;
    c->tree_info.valid = valid;
    c->tree_info.depth = depth;
    c->tree_info.build_ms = ms;
    c->tree_info.list_n = c->list_n;
    c->tree_info.trav_ploc = c->d_nodes_active == c->d_nodes2;
    c->tree_info.trav_depth = c->tree_info.trav_ploc ? c->trav_depth : depth;
    c->tree_info.policy = ptb_effective_policy(c, c->traversal_request);
    return 0;
}
