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
// child id < n : leaf slot (box = that triangle's bounds; unused by the predicates -- the reference does not box-test
// leaves);  id >= n : internal node id-n.
struct __align__(16) Node64 { float4 a, b, c, d; };
// traversal nodes (lbvh.cu k_pack_nodes): child id -1 = nothing below; PTB_NODE_MUST set = an ill-conditioned triangle below,
// whose computed depth obeys no bound: the child is never culled by distance
#define PTB_NODE_MUST 0x40000000
#define PTB_NODE_ID 0x3FFFFFFF
// flags of a leaf slot (w of its inflated lower bound, lbvh.cu k_tri_prep): never hit (degenerate), ill-conditioned (no leaf bound:
// bounded by its gate box, exempt from distance culling), big (candidate for the always-test list), on the always-test list
#define PTB_TF_NEVER 1
#define PTB_TF_MUST 2
#define PTB_TF_BIG 4
#define PTB_TF_LISTED 8
// slot_of[face] and the avoid field of the ray queues: the leaf slot in the low 24 bits (faces < 2^24, ptb_create); for a triangle on
// the always-test list, its list index + 1 in bits 24..29 (lbvh.cu k_tag_listed), so that k_trace_pre gets the entry to skip from a
// shift instead of comparing every entry's slot.  A tagged value never equals a tree leaf's slot (listed triangles are not in the
// traversal tree), so the tree kernels compare it as it is; code that indexes by slot masks it.
#define PTB_SLOT_MASK 0x00FFFFFF
#define PTB_SLOT_LIST_SHIFT 24
// 64-byte packed triangle, indexed by sorted leaf slot:
//   a = (v0.xyz, 1/D)  b = (u.xyz, uu)  c = (v.xyz, uv)  d = (n.xyz, vv)     u=v1-v0, v=v2-v0, n=u x v, D = uv*uv - uu*vv
// Every field is the f32 expression geometries.py:121-141 evaluates per test (they depend on the triangle only), so
// precomputing them changes no bit; 1/D is the correctly rounded reciprocal used by div_exact below.
struct __align__(16) Tri64 { float4 a, b, c, d; };

struct TraceScene {
    const Node64* __restrict__ nodes;    // [n-1]
    const Tri64* __restrict__ tris;      // [n]  by leaf slot
    const int* __restrict__ leaf;        // [n]  slot -> face id      (lbvh.py:55)
    const int* __restrict__ slot_of;     // [n]  face id -> slot      (inverse of leaf, proper trees only)
    const int* __restrict__ gate;        // [n]  slot -> internal node whose child the leaf is (its box gates the triangle test)
    const float4* __restrict__ gbox;     // [2n] slot -> reference box (lo, hi) of its gate
    const float4* __restrict__ tlo;      // [n]  slot -> inflated bounds of its triangle (lo.w: PTB_TF_* flags), hi
    const float4* __restrict__ thi;
    const float4* __restrict__ nlo;      // [n-1] traversal boxes of the internal nodes (lo, hi); [0] = the root's
    const float4* __restrict__ nhi;
    float scene_abs;                     // largest absolute coordinate of the (inflated) scene bounds
    int root_must;                       // the traversal root has an ill-conditioned triangle below it (no distance cull)
    const int* __restrict__ list;        // always-test list (leaf slots): big or ill-conditioned triangles kept out of the traversal tree
    int nlist;
    // reference arrays (lbvh.py:50-59) for the literal traversal
    const float* __restrict__ bmin;      // [n-1][3]
    const float* __restrict__ bmax;      // [n-1][3]
    const int2* __restrict__ child;      // [n-1]
    int n;
    // quantised copy of `nodes` for trees that do not fit in shared memory (32 B per node, see Node32)
    const uint4* __restrict__ qnodes;    // [2(n-1)]
    const uint4* __restrict__ wnodes;    // [4(n-1)] 4-wide quantised nodes (64 B: 12 box words + 4 child ids), or NULL
    float qbase[3], qext[3], qinv[3];    // plane = qbase + v * qext, v in [1, 2) on a 15-bit grid; qext a power of two, qinv = 1 / qext
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

// Correctly rounded a / d from the correctly rounded reciprocal r = RN(1/d): two Markstein correction steps.
// q0 = RN(a*r) is within ~2 ulp; after one step the estimate is faithful, after the second it is RN(a/d) (Markstein's
// theorem: faithful q, exact residual via FMA, correctly rounded reciprocal).  5 FMA-pipe instructions instead of the
// ~12 + MUFU of a full IEEE division.  Exactness vs `a / d` is checked on the device over 2^30 operand pairs
// (ptb_selftest, tests/test_gpu_parity.py) -- operands in the denormal range are outside the guarantee.
PTB_D float div_exact(float a, float d, float r) {
    float q = a * r;
    float e = __fmaf_rn(-q, d, a);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-q, d, a);
    return __fmaf_rn(e, r, q);
}

// per-ray constants of the fast slab test
struct RayPre { V3 o, d, r; bool par; };
PTB_D RayPre ray_pre(V3 o, V3 d) {
    RayPre P; P.o = o; P.d = d;
    P.r = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    P.par = fabsf(d.x) < PTB_EPS || fabsf(d.y) < PTB_EPS || fabsf(d.z) < PTB_EPS;
    return P;
}
// Same boolean and same tnear as slab_ref (identical quotients; min/max instead of the swap; the per-axis early
// `near > far` checks are implied by the final one because near only grows and far only shrinks).  Rays with an
// axis-parallel component (|d| < 1e-6) take the literal routine.
PTB_D bool slab_fast(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayPre& P, float* near_out) {
    if (P.par) return slab_ref(lox, loy, loz, hix, hiy, hiz, P.o, P.d, near_out);
    float x1 = div_exact(lox - P.o.x, P.d.x, P.r.x), x2 = div_exact(hix - P.o.x, P.d.x, P.r.x);
    float y1 = div_exact(loy - P.o.y, P.d.y, P.r.y), y2 = div_exact(hiy - P.o.y, P.d.y, P.r.y);
    float z1 = div_exact(loz - P.o.z, P.d.z, P.r.z), z2 = div_exact(hiz - P.o.z, P.d.z, P.r.z);
    float tnear = fmaxf(fmaxf(0.0f, fminf(x1, x2)), fmaxf(fminf(y1, y2), fminf(z1, z2)));
    float tfar = fminf(fminf(PTB_INF, fmaxf(x1, x2)), fminf(fmaxf(y1, y2), fmaxf(z1, z2)));
    *near_out = tnear;
    return !(tnear > tfar);
}

// ---- conservative slab test (production traversal) ------------------------------------------------------------------------------
// The reference's slab test is monotone in the box: for B' inside B (lo' >= lo, hi' <= hi componentwise) every f32 step of
// Box.intersect -- RN(p - o), RN(x / d), the swap, min/max -- is a monotone function of the plane coordinate p, so the t-interval
// of B' is contained in the t-interval of B axis by axis (and `origin inside the slab` of B' implies it for B on a parallel
// axis).  Hence  hit(B') => hit(B)  EXACTLY, in f32.  Internal boxes are exact min/max unions of their children, so along a
// root-to-leaf chain "every ancestor's box passes" is equivalent to "the box of the leaf's parent passes": the reference tests a
// triangle iff the exact slab test of its parent's box (its GATE) passes.  The production traversal therefore only needs box
// tests that never reject a box the exact test accepts; it finds candidate triangles with the cheap test below (1 FMA per
// plane) and runs the exact division test once, on the gate of a triangle that is about to be accepted (gate_passes).
//
// Error bound.  Reference: t_ref = RN(RN(p - o) / d) = T(1+e1)(1+e2), T = (p-o)/d, |e| <= u = 2^-24.
// Here: r = RN(1/d), c = RN(o*r), t = RN(p*r - c) (one FMA) = (T(1+e3) - (o/d)(1+e3)e4)(1+e5).
// => |t - t_ref| <= 4.1u|t| + 1.1u|c|.  With A = max_axis |c| the bounds  L(t) = t - 8u|t| - 4uA  <=  t_ref  <=  U(t) = t + 8u|t| + 4uA
// are monotone in t, so they commute with the per-axis min/max and with the clamps (0 and 1e6 are exact):
//   tnear_ref >= lb := max(0, near)(1 - 8u) - a2,   tfar_ref <= ub := far + 8u|far| + a2',   a2 + a2' folded into lb (a2 = 8uA + 1e-30;
// the 1e-30 covers results in the denormal range).  The exact test passes only if tnear_ref <= tfar_ref, hence only if lb <= ub.
// Comparisons are written so that a NaN (overflowing coordinates) counts as "hit".  Rays with an axis-parallel component,
// non-finite or huge components are not handled here: the kernel sends them through trace_ordered (exact tests).
#define PTB_CONS_KAPPA 4.76837158203125e-7f      /* 2^-21 = 8u */
struct RayCons { V3 o, d, r, nc; float a2; };
PTB_D bool ray_is_special(V3 o, V3 d) {
    const float big = 1e30f;
    bool ok = fabsf(d.x) >= PTB_EPS && fabsf(d.y) >= PTB_EPS && fabsf(d.z) >= PTB_EPS && fabsf(d.x) <= big && fabsf(d.y) <= big && fabsf(d.z) <= big &&
              fabsf(o.x) <= big && fabsf(o.y) <= big && fabsf(o.z) <= big;
    return !ok;            // NaNs fail every comparison above
}
PTB_D RayCons ray_cons(V3 o, V3 d) {
    RayCons R; R.o = o; R.d = d;
    R.r = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    R.nc = mk3(-(o.x * R.r.x), -(o.y * R.r.y), -(o.z * R.r.z));
    R.a2 = __fmaf_rn(fmaxf(fmaxf(fabsf(R.nc.x), fabsf(R.nc.y)), fabsf(R.nc.z)), PTB_CONS_KAPPA, 1e-30f);
    return R;
}
// *lb: lower bound of the reference's tnear for this box (valid whether or not the box is hit); returns false only if the
// reference's test certainly fails.  *sure = the reference's test CERTAINLY passes: an upper bound of its tnear is <= a lower bound of its tfar
// (tnear_ref <= tn(1+8u) + a2,  tfar_ref >= tf - 8u|tf|, a2 carrying both absolute terms).  A triangle whose gate is `sure` needs
// no exact gate test; only rays grazing the gate (difference within ~1e-6 relative) are ambiguous.
PTB_D bool slab_cons2(float lox, float loy, float loz, float hix, float hiy, float hiz, const RayCons& R, float* lb, bool* sure) {
    const float x1 = __fmaf_rn(lox, R.r.x, R.nc.x), x2 = __fmaf_rn(hix, R.r.x, R.nc.x);
    const float y1 = __fmaf_rn(loy, R.r.y, R.nc.y), y2 = __fmaf_rn(hiy, R.r.y, R.nc.y);
    const float z1 = __fmaf_rn(loz, R.r.z, R.nc.z), z2 = __fmaf_rn(hiz, R.r.z, R.nc.z);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fmaxf(fminf(z1, z2), 0.0f));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fminf(fmaxf(z1, z2), PTB_INF));
    const float l = __fmaf_rn(tn, 1.0f - PTB_CONS_KAPPA, -R.a2);
    const float ub = __fmaf_rn(fabsf(tf), PTB_CONS_KAPPA, tf);
    const float hi_n = __fmaf_rn(tn, 1.0f + PTB_CONS_KAPPA, R.a2);
    const float lo_f = __fmaf_rn(fabsf(tf), -PTB_CONS_KAPPA, tf);
    *lb = l;
    *sure = hi_n <= lo_f;              // false on NaN
    return !(l > ub);
}

// ---- traversal structure (production kernel) -------------------------------------------------------------------------------------
// With the gate lemma above, the reference's answer is: the smallest f32 depth (ties: largest leaf slot) over the triangles T that
// (i) pass Face.intersect in f32 and (ii) whose gate passes Box.intersect in f32.  Any structure that enumerates a superset of the
// triangles satisfying (i) gives the same answer once (ii) is checked per candidate.  The traversal tree keeps the LBVH topology
// but bounds every node by the union of the INFLATED bounds of the triangles below it -- inflated by eps_T, the distance within
// which the f32 arithmetic of Face.intersect can accept a point outside the triangle (lbvh.cu k_tri_prep) -- and leaves
// out triangles that would spoil it: big ones (their bounds make every ancestor's box huge) and ill-conditioned ones (no usable
// eps_T) go to a short always-test list.
// Ray side of the bound: if T is accepted at depth r, the real point X = ro + r*rd lies within eps_T + delta of T's bounds on every
// axis, delta = 64u(|ro|_1 + r|rd|_1).  Since X is then within eps_T + delta of the scene, r|rd_a| <= S + |o_a| per axis (S = largest
// absolute coordinate of the inflated scene), which bounds delta by the per-ray constant 65u(2|ro|_1 + 3S) -- no r left in it.
// A spatial slack delta on axis a moves that axis' slab interval by delta*|1/d_a| at both ends: folded, with the right sign, into
// the two FMA constants  nc1 = -o/d - delta/d  (lo planes)  and  nc2 = -o/d + delta/d  (hi planes).  The rounding of the 1-FMA
// arithmetic itself is covered by the relative margin 8u and the absolute margin a2, as for the gate test.
struct RayTrav { V3 nc1, nc2; };
PTB_D float trav_delta(V3 o, float scene_abs) {
    const float U = 5.9604644775390625e-8f;
    return 65.0f * U * (2.0f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z)) + 3.0f * scene_abs);
}
PTB_D RayTrav ray_trav(const RayCons& R, float delta) {
    RayTrav Q;
    Q.nc1 = mk3(__fmaf_rn(-delta, R.r.x, R.nc.x), __fmaf_rn(-delta, R.r.y, R.nc.y), __fmaf_rn(-delta, R.r.z, R.nc.z));
    Q.nc2 = mk3(__fmaf_rn(delta, R.r.x, R.nc.x), __fmaf_rn(delta, R.r.y, R.nc.y), __fmaf_rn(delta, R.r.z, R.nc.z));
    return Q;
}
// slab test of a traversal-tree box: false only if no triangle below can be accepted; *lb <= depth of anything accepted below.
PTB_D bool slab_trav(float4 lo, float4 hi, const RayCons& R, const RayTrav& Q, float* lb) {
    const float x1 = __fmaf_rn(lo.x, R.r.x, Q.nc1.x), x2 = __fmaf_rn(hi.x, R.r.x, Q.nc2.x);
    const float y1 = __fmaf_rn(lo.y, R.r.y, Q.nc1.y), y2 = __fmaf_rn(hi.y, R.r.y, Q.nc2.y);
    const float z1 = __fmaf_rn(lo.z, R.r.z, Q.nc1.z), z2 = __fmaf_rn(hi.z, R.r.z, Q.nc2.z);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fmaxf(fminf(z1, z2), 0.0f));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fminf(fmaxf(z1, z2), PTB_INF));
    const float l = __fmaf_rn(tn, 1.0f - PTB_CONS_KAPPA, -R.a2);
    const float ub = __fmaf_rn(fabsf(tf), PTB_CONS_KAPPA, tf);
    *lb = l;
    return !(l > ub);
}

// geometries.py:117-148 on the packed record, literal (true divisions).  Returns hit; depth/s/t as the reference.
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

// The same test with (i) the depth compared against `tlimit` BEFORE the barycentrics are computed (a candidate with
// depth > tlimit can never be accepted, so skipping the rest changes nothing) and (ii) s, t through div_exact.
PTB_D bool tri_fast(const Tri64& T, V3 ro, V3 rd, float tlimit, float* depth, float* s_out, float* t_out) {
    V3 v0 = mk3(T.a.x, T.a.y, T.a.z), nrm = mk3(T.d.x, T.d.y, T.d.z);
    float b = dot(nrm, rd);
    if (!(fabsf(b) >= PTB_EPS)) return false;
    float a = -dot(nrm, ro - v0);
    // RN(a / b) > 0 needs a != 0 and equal signs: decided without the division (which takes its slow path for a == 0 -- common,
    // the origin lies on the plane of every coplanar neighbour of the surface it left)
    if (a == 0.0f || ((a < 0.0f) != (b < 0.0f))) return false;
    float r = a / b;
    if (!(r > 0.0f) || r > tlimit) return false;
    V3 u = mk3(T.b.x, T.b.y, T.b.z), v = mk3(T.c.x, T.c.y, T.c.z);
    float uu = T.b.w, uv = T.c.w, vv = T.d.w, rD = T.a.w;
    V3 w = (ro + r * rd) - v0;
    float wu = dot(w, u), wv = dot(w, v);
    float D = uv * uv - uu * vv;
    float s = div_exact(uv * wv - vv * wu, D, rD);
    float t = div_exact(uv * wu - uu * wv, D, rD);
    *s_out = s; *t_out = t; *depth = r;
    return (0.0f <= s && s <= 1.0f) && (0.0f <= t && s + t <= 1.0f);
}

// ---- quantised traversal nodes -------------------------------------------------------------------------------------------------------
// A tree too big for shared memory is walked out of L2, and every step waits for the slowest of a warp's 32 scattered fetches, so
// what matters is that the whole node array stays cache resident and one fetch brings a node.  Node32 stores the two child boxes on
// a per-axis 15-bit grid over the scene, rounded OUTWARD (the traversal tree only has to be conservative: a larger box can add
// candidates, never lose one):  plane = qbase + v * qext,  v = 1 + q / 32768.  The 16-bit field is 0x8000 | q, so that one PRMT
// with 0x3F000000 turns it into the f32 v, and the plane is never formed: with A = qext * r (exact, qext is a power of two) and
// B = fma(qbase, r, nc) the slab distance is fma(v, A, B) -- the product is exact, so against fma(plane, r, nc) the only new
// error is the rounding of B, at most u|B| per plane, which ray_quant folds into B itself.
//   words 0-5: one per box and axis, (lo | hi << 16): box 0 x, y, z, box 1 x, y, z; 6, 7: child ids
// Which of an axis' two planes the ray enters through is a property of the ray (the sign of 1/d), so nothing has to be compared per
// box: a per-ray PRMT selector takes the entry plane's field out of the axis' word, its complement the exit plane's, and the slab
// distances are max3 / min3 of three FMAs each.  Same values as min / max of the two distances per axis: for r > 0, lo <= hi, A > 0 and
// B(lo) <= B(hi) make the lo plane's distance the smaller one by the monotonicity of the FMA, and the other way round for r < 0.
struct RayQuant { V3 A, Bn, Bf; unsigned sx, sy, sz; };        // Bn / Bf: constants of the entry / exit planes; s*: PRMT selector of the entry plane
#define PTB_QSEL_LO 0x7104u
#define PTB_QSEL_HI 0x7324u
#define PTB_QSEL_FLIP 0x0220u
PTB_D RayQuant ray_quant(const TraceScene& S, const RayCons& R, const RayTrav& Q) {
    RayQuant G;
    G.A = mk3(S.qext[0] * R.r.x, S.qext[1] * R.r.y, S.qext[2] * R.r.z);
    // B is rounded once (|error| <= u|B|): moved by 4u|B| (at least 3u|B| after the rounding of this fma) to the side that pushes the
    // plane outward -- down for the lo planes, up for the hi planes when r > 0, the other way when r < 0.  Per axis, like delta: an
    // absolute margin shared by the axes would be scaled by the largest |1/d| and make nearly axis-parallel rays enter every box.
    const float K4 = 2.384185791015625e-7f;        // 4u
    const V3 k = mk3(copysignf(K4, R.r.x), copysignf(K4, R.r.y), copysignf(K4, R.r.z));
    V3 b1 = mk3(__fmaf_rn(S.qbase[0], R.r.x, Q.nc1.x), __fmaf_rn(S.qbase[1], R.r.y, Q.nc1.y), __fmaf_rn(S.qbase[2], R.r.z, Q.nc1.z));
    V3 b2 = mk3(__fmaf_rn(S.qbase[0], R.r.x, Q.nc2.x), __fmaf_rn(S.qbase[1], R.r.y, Q.nc2.y), __fmaf_rn(S.qbase[2], R.r.z, Q.nc2.z));
    const V3 B1 = mk3(__fmaf_rn(-k.x, fabsf(b1.x), b1.x), __fmaf_rn(-k.y, fabsf(b1.y), b1.y), __fmaf_rn(-k.z, fabsf(b1.z), b1.z));      // lo planes
    const V3 B2 = mk3(__fmaf_rn(k.x, fabsf(b2.x), b2.x), __fmaf_rn(k.y, fabsf(b2.y), b2.y), __fmaf_rn(k.z, fabsf(b2.z), b2.z));         // hi planes
    const bool px = !(R.r.x < 0.0f), py = !(R.r.y < 0.0f), pz = !(R.r.z < 0.0f);
    G.Bn = mk3(px ? B1.x : B2.x, py ? B1.y : B2.y, pz ? B1.z : B2.z);
    G.Bf = mk3(px ? B2.x : B1.x, py ? B2.y : B1.y, pz ? B2.z : B1.z);
    G.sx = px ? PTB_QSEL_LO : PTB_QSEL_HI; G.sy = py ? PTB_QSEL_LO : PTB_QSEL_HI; G.sz = pz ? PTB_QSEL_LO : PTB_QSEL_HI;
    return G;
}
PTB_D float quant_lo(unsigned w) { return __uint_as_float(__byte_perm(w, 0x3F000000u, PTB_QSEL_LO)); }
PTB_D float quant_hi(unsigned w) { return __uint_as_float(__byte_perm(w, 0x3F000000u, PTB_QSEL_HI)); }
PTB_D float quant_sel(unsigned w, unsigned sel) { return __uint_as_float(__byte_perm(w, 0x3F000000u, sel)); }
// box = the three axis words of a Node32 / wide-node entry
PTB_D bool slab_quant(unsigned wx, unsigned wy, unsigned wz, const RayQuant& G, float a2, float* lb) {
    const float xn = __fmaf_rn(quant_sel(wx, G.sx), G.A.x, G.Bn.x), xf = __fmaf_rn(quant_sel(wx, G.sx ^ PTB_QSEL_FLIP), G.A.x, G.Bf.x);
    const float yn = __fmaf_rn(quant_sel(wy, G.sy), G.A.y, G.Bn.y), yf = __fmaf_rn(quant_sel(wy, G.sy ^ PTB_QSEL_FLIP), G.A.y, G.Bf.y);
    const float zn = __fmaf_rn(quant_sel(wz, G.sz), G.A.z, G.Bn.z), zf = __fmaf_rn(quant_sel(wz, G.sz ^ PTB_QSEL_FLIP), G.A.z, G.Bf.z);
    const float tn = fmaxf(fmaxf(xn, yn), fmaxf(zn, 0.0f));
    const float tf = fminf(fminf(xf, yf), fminf(zf, PTB_INF));
    const float l = __fmaf_rn(tn, 1.0f - PTB_CONS_KAPPA, -a2);
    const float ub = __fmaf_rn(fabsf(tf), PTB_CONS_KAPPA, tf);
    *lb = l;
    return !(l > ub);
}

// The exact gate: Box.intersect (division form) on the box of internal node g, as the reference evaluates it when it pops g.
PTB_D bool gate_passes(const TraceScene& S, int g, V3 ro, V3 rd) {
    float nr;
    return slab_ref(S.bmin[3 * g], S.bmin[3 * g + 1], S.bmin[3 * g + 2], S.bmax[3 * g], S.bmax[3 * g + 1], S.bmax[3 * g + 2], ro, rd, &nr);
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
            int index = S.leaf[curr];
            if (index != avoid) {
                const Tri64 T = S.tris[curr];
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

// ---- ordered traversal with the EXACT predicates, over the reference's own arrays (child / bmin / bmax) ----------------------
// Near child first; ties in depth go to the larger leaf slot.  No distance culling: the f32 depth of a triangle is not bounded by
// the f32 entry distance of the reference boxes around it (grazing rays / sliver triangles), so only the reference's own hit / miss
// gating is used -- every triangle the reference tests is tested.  The checker for the production kernel
// (PTB_TRAVERSE_ORDERED_EXACT) and the tracer of the rays it sets aside (axis-parallel, non-finite).
// ANYHIT: stop at the first accepted triangle with depth <= tmax (shadow rays: the reference calls the ray occluded
// iff its CLOSEST hit has depth <= dis, which holds iff ANY reachable triangle has one).
template <bool ANYHIT, bool COUNT>
PTB_D HitRec trace_ordered(const TraceScene& S, V3 ro, V3 rd, int avoid, float tmax, TraceCounters* C) {
    HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
    const int n = S.n;
    if (n < 2) {
        if (n == 1) return trace_reference<COUNT>(S, ro, rd, avoid, C);
        return ret;
    }
    const RayPre P = ray_pre(ro, rd);
    float nr;
    if (COUNT) C->boxes++;
    // the root's own box (the reference pops and tests it first)
    if (!slab_fast(S.bmin[0], S.bmin[1], S.bmin[2], S.bmax[0], S.bmax[1], S.bmax[2], P, &nr)) return ret;
    const int avoid_slot = avoid >= 0 ? (S.slot_of[avoid] & PTB_SLOT_MASK) : -1;
    float best = ANYHIT ? fminf(tmax, PTB_INF) : PTB_INF;
    const float cull = PTB_INF * 4.0f;        // nothing is culled by distance
    int stack_id[PTB_STACK];
    float stack_near[PTB_STACK];
    int sp = 0;
    int cur = 0;          // internal node whose box passed, -1 = none
    float cur_near = 0.0f;
    while (true) {
        int pend0 = -1, pend1 = -1;
        while (cur >= 0) {
            const int2 ch = S.child[cur];
            if (COUNT) C->nodes++;
            const int c0 = ch.x, c1 = ch.y;
            const bool leaf0 = c0 < n, leaf1 = c1 < n;
            float n0 = 0.0f, n1 = 0.0f;
            bool d0 = false, d1 = false;
            if (!leaf0) { const int j = c0 - n; d0 = slab_fast(S.bmin[3 * j], S.bmin[3 * j + 1], S.bmin[3 * j + 2], S.bmax[3 * j], S.bmax[3 * j + 1], S.bmax[3 * j + 2], P, &n0) && !(n0 > cull); if (COUNT) C->boxes++; }
            if (!leaf1) { const int j = c1 - n; d1 = slab_fast(S.bmin[3 * j], S.bmin[3 * j + 1], S.bmin[3 * j + 2], S.bmax[3 * j], S.bmax[3 * j + 1], S.bmax[3 * j + 2], P, &n1) && !(n1 > cull); if (COUNT) C->boxes++; }
            if (leaf0 && c0 != avoid_slot) pend0 = c0;
            if (leaf1 && c1 != avoid_slot) { if (pend0 < 0) pend0 = c1; else pend1 = c1; }
            if (d0 && d1) {
                const bool first1 = !(n0 < n1);
                if (sp < PTB_STACK) { stack_id[sp] = (first1 ? c0 : c1) - n; stack_near[sp] = first1 ? n0 : n1; sp++; }
                if (COUNT) C->max_stack = max(C->max_stack, (unsigned)sp);
                cur = (first1 ? c1 : c0) - n; cur_near = first1 ? n1 : n0;
            } else if (d0) { cur = c0 - n; cur_near = n0; }
            else if (d1) { cur = c1 - n; cur_near = n1; }
            else cur = -1;
            if (pend0 >= 0) break;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int slot = k == 0 ? pend0 : pend1;
            if (slot < 0) continue;
            const Tri64 T = S.tris[slot];
            if (COUNT) C->tris++;
            float dep, s, t;
            if (tri_fast(T, ro, rd, best, &dep, &s, &t)) {
                if (ANYHIT) {
                    if (dep < PTB_INF) { ret.hit = 1; ret.depth = dep; ret.u = s; ret.v = t; ret.slot = slot; ret.index = S.leaf[slot]; return ret; }
                } else if (dep < ret.depth || (ret.hit && slot > ret.slot)) {     // here dep <= best == ret.depth
                    ret.depth = dep; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = slot;
                    best = dep;
                }
            }
        }
        if (cur < 0) {
            if (sp == 0) break;
            --sp;
            cur = stack_id[sp]; cur_near = stack_near[sp];
        }
    }
    if (!ANYHIT && ret.hit) ret.index = S.leaf[ret.slot];
    return ret;
}
