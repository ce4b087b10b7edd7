// ptb_traverse.cuh -- ray/box, ray/triangle predicates and BVH traversal.
//
// Reference: Box.intersect geometries.py:23-46, Face.intersect geometries.py:117-148,
// LinearBVH.intersect tree/lbvh.py:313-347.
//
// Results contract (bit-exact hit ids): the reference tests a triangle iff the exact slab test of every
// ancestor's box passes, and keeps the smallest f32 depth with strict `<`, visiting child1 before child0.
// Both traversals below evaluate the SAME f32 predicates (same operations, same order, IEEE division);
// they differ only in visiting order:
//   * trace_reference : the reference's own order (unordered child1-first DFS, no culling).
//   * trace_ordered   : near child first, sub-trees whose entry distance is beyond the current best
//                       (plus a relative guard band) are skipped, and ties in depth are resolved to the
//                       leaf the reference would have visited first (= the larger sorted-leaf slot, valid
//                       for proper trees only; ptb_build_tree checks that and falls back otherwise).
#pragma once
#include "ptb_math.cuh"

// 64-byte packed internal node: both children's boxes and ids in four 16-byte loads.
//   a = (lo0.xyz, c0)  b = (hi0.xyz, c1)  c = (lo1.xyz, -)  d = (hi1.xyz, -)
// child id < n : leaf slot (box = that triangle's bounds, used for ordering/culling only -- the reference
// does not box-test leaves);  id >= n : internal node id-n.
struct __align__(16) Node64 { float4 a, b, c, d; };
// 64-byte packed triangle, indexed by sorted leaf slot:
//   a = (v0.xyz, face id)  b = (u.xyz, uu)  c = (v.xyz, uv)  d = (n.xyz, vv)     u=v1-v0, v=v2-v0, n=u x v
struct __align__(16) Tri64 { float4 a, b, c, d; };

struct TraceScene {
    const Node64* __restrict__ nodes;    // [n-1]
    const Tri64* __restrict__ tris;      // [n]  by leaf slot
    // reference arrays (lbvh.py:50-59) for the literal traversal
    const float* __restrict__ bmin;      // [n-1][3]
    const float* __restrict__ bmax;      // [n-1][3]
    const int2* __restrict__ child;      // [n-1]
    float root_lo[3], root_hi[3];
    int n;
};

struct HitRec { int hit; float depth; int index; float u, v; int slot; };

struct TraceCounters { unsigned long long nodes, boxes, tris; unsigned int max_stack; };

#define PTB_STACK 64
#define PTB_CULL_GUARD 2.44140625e-4f   /* 2^-12 relative guard band on distance culling */

// geometries.py:23-46, literal.  Returns hit; *near_out = entry distance (>= 0).
PTB_D bool slab_ref(float lox, float loy, float loz, float hix, float hiy, float hiz, V3 o, V3 d, float* near_out) {
    float tnear = 0.0f, tfar = PTB_INF;
    bool hit = true;
    if (fabsf(d.x) < PTB_EPS) { if (o.x < lox || o.x > hix) hit = false; }
    else {
        float i1 = (lox - o.x) / d.x, i2 = (hix - o.x) / d.x;
        if (i1 > i2) { float t = i1; i1 = i2; i2 = t; }
        tfar = fminf(tfar, i2); tnear = fmaxf(tnear, i1);
        if (tnear > tfar) hit = false;
    }
    if (fabsf(d.y) < PTB_EPS) { if (o.y < loy || o.y > hiy) hit = false; }
    else {
        float i1 = (loy - o.y) / d.y, i2 = (hiy - o.y) / d.y;
        if (i1 > i2) { float t = i1; i1 = i2; i2 = t; }
        tfar = fminf(tfar, i2); tnear = fmaxf(tnear, i1);
        if (tnear > tfar) hit = false;
    }
    if (fabsf(d.z) < PTB_EPS) { if (o.z < loz || o.z > hiz) hit = false; }
    else {
        float i1 = (loz - o.z) / d.z, i2 = (hiz - o.z) / d.z;
        if (i1 > i2) { float t = i1; i1 = i2; i2 = t; }
        tfar = fminf(tfar, i2); tnear = fmaxf(tnear, i1);
        if (tnear > tfar) hit = false;
    }
    *near_out = tnear;
    return hit;
}

// geometries.py:117-148 on the packed record (u, v, n, uu, uv, vv are the same f32 expressions the
// reference evaluates per test; D is recomputed).  Returns hit; depth/s/t as the reference.
PTB_D bool tri_ref(const Tri64& T, V3 ro, V3 rd, float* depth, float* s_out, float* t_out) {
    V3 v0 = mk3(T.a.x, T.a.y, T.a.z), u = mk3(T.b.x, T.b.y, T.b.z), v = mk3(T.c.x, T.c.y, T.c.z), nrm = mk3(T.d.x, T.d.y, T.d.z);
    float b = dot(nrm, rd);
    if (!(fabsf(b) >= PTB_EPS)) return false;
    V3 w0 = ro - v0;
    float a = -dot(nrm, w0);
    float r = a / b;
    if (!(r > 0.0f)) return false;
    V3 ip = ro + r * rd;
    float uu = T.b.w, uv = T.c.w, vv = T.d.w;
    V3 w = ip - v0;
    float wu = dot(w, u), wv = dot(w, v);
    float D = uv * uv - uu * vv;
    float s = (uv * wv - vv * wu) / D;
    float t = (uv * wu - uu * wv) / D;
    *s_out = s; *t_out = t; *depth = r;
    return (0.0f <= s && s <= 1.0f) && (0.0f <= t && s + t <= 1.0f);
}

// ---- tree/lbvh.py:313-347, literal order ------------------------------------------------------------------
template <bool COUNT>
PTB_D HitRec trace_reference(const TraceScene& S, V3 ro, V3 rd, int avoid, TraceCounters* C) {
    int stack[PTB_STACK];
    int sp = 0;
    const int n = S.n;
    stack[sp++] = n;
    HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
    int ntimes = 0;
    while (ntimes < n && sp != 0) {
        int curr = stack[--sp];
        if (curr < n) {
            Tri64 T = S.tris[curr];
            int index = __float_as_int(T.a.w);
            if (index != avoid) {
                if (COUNT) C->tris++;
                float dep, s, t;
                if (tri_ref(T, ro, rd, &dep, &s, &t) && dep < ret.depth) {
                    ret.depth = dep; ret.index = index; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = curr;
                }
            }
            continue;
        }
        int i = curr - n;
        float nr;
        if (COUNT) C->boxes++;
        if (!slab_ref(S.bmin[3 * i], S.bmin[3 * i + 1], S.bmin[3 * i + 2], S.bmax[3 * i], S.bmax[3 * i + 1], S.bmax[3 * i + 2], ro, rd, &nr)) continue;
        ntimes++;
        if (COUNT) C->nodes++;
        int2 ch = S.child[i];
        if (sp + 2 <= PTB_STACK) { stack[sp++] = ch.x; stack[sp++] = ch.y; }
        if (COUNT) C->max_stack = max(C->max_stack, (unsigned)sp);
    }
    return ret;
}

// ---- ordered traversal over packed nodes -----------------------------------------------------------------------
// ANYHIT: stop at the first accepted triangle with depth <= tmax (shadow rays: the reference calls the
// ray occluded iff its CLOSEST hit has depth <= dis, which holds iff ANY reachable triangle has).
template <bool ANYHIT, bool COUNT>
PTB_D HitRec trace_ordered(const TraceScene& S, V3 ro, V3 rd, int avoid, float tmax, TraceCounters* C) {
    HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
    const int n = S.n;
    if (n < 2) {
        if (n == 1) return trace_reference<COUNT>(S, ro, rd, avoid, C);
        return ret;
    }
    float limit = ANYHIT ? tmax : PTB_INF;             // distances beyond `limit` cannot change the answer
    float cull = limit + limit * PTB_CULL_GUARD;
    {   // the root's own box (the reference pops and tests it first)
        float nr;
        if (COUNT) C->boxes++;
        if (!slab_ref(S.root_lo[0], S.root_lo[1], S.root_lo[2], S.root_hi[0], S.root_hi[1], S.root_hi[2], ro, rd, &nr)) return ret;
    }
    int stack_id[PTB_STACK];
    float stack_near[PTB_STACK];
    int sp = 0;
    int cur = 0;   // internal node index
    while (true) {
        const Node64 N = S.nodes[cur];
        if (COUNT) C->nodes++;
        int c0 = __float_as_int(N.a.w), c1 = __float_as_int(N.b.w);
        float n0, n1;
        bool h0 = slab_ref(N.a.x, N.a.y, N.a.z, N.b.x, N.b.y, N.b.z, ro, rd, &n0);
        bool h1 = slab_ref(N.c.x, N.c.y, N.c.z, N.d.x, N.d.y, N.d.z, ro, rd, &n1);
        if (COUNT) C->boxes += 2;
        bool leaf0 = c0 < n, leaf1 = c1 < n;
        // leaves: always tested when reached (no box predicate in the reference); internal: exact predicate,
        // then distance culling with a guard band
        bool go0 = leaf0 ? true : (h0 && !(n0 > cull));
        bool go1 = leaf1 ? true : (h1 && !(n1 > cull));
#pragma unroll
        for (int k = 0; k < 2; k++) {
            bool isleaf = k == 0 ? leaf0 : leaf1;
            if (!isleaf) continue;
            int slot = k == 0 ? c0 : c1;
            const Tri64 T = S.tris[slot];
            int index = __float_as_int(T.a.w);
            if (index == avoid) continue;
            if (COUNT) C->tris++;
            float dep, s, t;
            if (tri_ref(T, ro, rd, &dep, &s, &t)) {
                if (ANYHIT) {
                    if (dep <= tmax) { ret.hit = 1; ret.depth = dep; ret.index = index; ret.u = s; ret.v = t; ret.slot = slot; return ret; }
                } else if (dep < ret.depth || (dep == ret.depth && ret.hit && slot > ret.slot)) {
                    ret.depth = dep; ret.index = index; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = slot;
                    cull = dep + dep * PTB_CULL_GUARD;
                }
            }
        }
        bool d0 = go0 && !leaf0 && !(n0 > cull), d1 = go1 && !leaf1 && !(n1 > cull);
        if (d0 && d1) {
            // nearer first; on equal entry distance take child1 first like the reference
            bool first1 = !(n0 < n1);
            int nearc = first1 ? c1 : c0, farc = first1 ? c0 : c1;
            float farn = first1 ? n0 : n1;
            if (sp < PTB_STACK) { stack_id[sp] = farc - n; stack_near[sp] = farn; sp++; }
            if (COUNT) C->max_stack = max(C->max_stack, (unsigned)sp);
            cur = nearc - n;
            continue;
        }
        if (d0) { cur = c0 - n; continue; }
        if (d1) { cur = c1 - n; continue; }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            if (!(stack_near[sp] > cull)) { cur = stack_id[sp]; found = true; break; }
        }
        if (!found) break;
    }
    return ret;
}
