// ptb_trace_kernel.cuh -- the traversal kernels.
//
// Production policy (PTB_TRAVERSE_ORDERED), two kernels per ray queue:
//  k_trace_pre   : one ray per thread, uniform control flow.  Decodes the queue record, computes the per-ray constants of the
//                  conservative slab test, tests the always-test list (every thread of a warp tests the same triangle: broadcast
//                  loads; identical boxes -- the two triangles of a wall quad -- are tested once), and tests the root box of the traversal tree.  A ray with nothing left to do (no box of the tree in
//                  reach -- most rays of a room-like scene) is finished here; the others are appended, as self-contained 80-byte
//                  records, to the tree queue.  Axis-parallel / non-finite rays, which the conservative test cannot handle, are
//                  traced here with the exact routine (trace_ordered).
//  k_trace_tree  : persistent warps, one ray per lane, traversal state in registers (stack: first entries in shared memory, the rest in
//                  local memory; shallow resident trees keep all of it in shared memory as 32-bit entries, S16).  A warp stages the records of the tree queue 16 at a time into a double-buffered shared-memory
//                  tile with cp.async, so the DRAM latency of the next tile is covered by the traversal of the current one; a lane
//                  that finishes its ray takes the next staged record.  Every iteration the warp votes for ONE kind of step -- a
//                  node step (fetch a 64-B node, two conservative slab tests, push / descend) or a leaf step (gate test + triangle
//                  test from the lane's ring of pending leaves in shared memory) -- whichever more lanes can take.  For scenes
//                  whose packed BVH fits, each CTA keeps it resident in shared memory (SMEM variant).
//  Box tests are conservative (1 FMA per plane); the exact reference test is evaluated on the gate box of a triangle that is about
//  to be accepted and that the conservative bounds cannot decide (ptb_traverse.cuh).  Results are bit-identical to trace_reference.
//
//  k_trace_simple: one ray per thread.  POLICY 0 = the reference's literal traversal (lbvh.py:313-347): used when the tree is not a
//                  proper tree, when PTB_TRAVERSE_REFERENCE is requested, and by the tests as the on-device checker.
//                  POLICY 1 = trace_ordered (exact slab test at every node): the PTB_TRAVERSE_ORDERED_EXACT checker.
#pragma once
#include "ptb_internal.h"

#define PTB_TRACE_BLK 128
#define PTB_TILE 16                 /* ray records per staged tile */
#ifndef PTB_POOL
#define PTB_POOL 64                 /* queue positions a warp reserves per atomic */
#endif
#ifndef PTB_FETCH_MIN
#define PTB_FETCH_MIN 8             /* idle lanes that trigger a refill from the staged tile: resident trees */
#endif
#ifndef PTB_FETCH_MIN_G
#define PTB_FETCH_MIN_G 4           /* the same for trees walked out of global memory (a ray lives longer there, so an idle lane costs more: +1.3 % on config 4) */
#endif

struct RayIn { int item; V3 ro, rd; int avoid_slot; float tmax; V3 c; };

// Queue records (float4 arrays indexed by queue position), written by raygen / shade / k_pack_tap:
//   q0 = (origin, path slot)
//   q1 = (unit direction, avoid leaf slot)   extend   |  (direction, distance to the light sample)                  shadow
//   q2 =  --                                          |  (contribution if unoccluded, avoid leaf slot)               shadow
// Tree-queue records (ExpQ, written by k_trace_pre):
//   e0 = (origin, path slot)  e1 = (direction, avoid leaf slot)  e2 = (1/d, delta)  e3 = (-o/d, -)
//   e4 = provisional closest hit over the always-test list (depth, u, v, leaf slot or -1)   extend
//      = (contribution if unoccluded, distance to the light sample)                            shadow

// ---- extend queue: closest hit for path p (path.py:28-29) ---------------------------------------------------------------------------
struct ExtendIO {
    static constexpr bool kAnyHit = false;
    static constexpr int K = 2;
    const float4* __restrict__ q0; const float4* __restrict__ q1; float4* hit;
    PTB_D const float4* rec(int k) const { return k == 0 ? q0 : q1; }
    PTB_D void decode(const float4* r, RayIn* in) const {
        in->item = __float_as_int(r[0].w); in->ro = mk3(r[0].x, r[0].y, r[0].z);
        in->rd = mk3(r[1].x, r[1].y, r[1].z); in->avoid_slot = __float_as_int(r[1].w); in->tmax = PTB_INF; in->c = v3s(0.0f);
    }
    PTB_D void store(int p, const HitRec& h, V3) const { hit[p] = make_float4(h.depth, h.u, h.v, __int_as_float(h.hit ? h.index : -1)); }
};
// ---- shadow queue: Ray(hitpos, li.dir) against avoid = hit triangle; unoccluded -> add the pending contribution (path.py:49-55).
// The direction is used as sampled (not re-normalised, like the reference).  Exactly one shadow ray per path per launch touches
// result[p], so the reduction `red.add` computes the same single rounded sum as a load-add-store, without making the warp wait
// for the load.
struct ShadowIO {
    static constexpr bool kAnyHit = true;
    static constexpr int K = 3;
    const float4* __restrict__ q0; const float4* __restrict__ q1; const float4* __restrict__ q2; float4* result;
    PTB_D const float4* rec(int k) const { return k == 0 ? q0 : (k == 1 ? q1 : q2); }
    PTB_D void decode(const float4* r, RayIn* in) const {
        in->item = __float_as_int(r[0].w); in->ro = mk3(r[0].x, r[0].y, r[0].z);
        in->rd = mk3(r[1].x, r[1].y, r[1].z); in->tmax = r[1].w;
        in->c = mk3(r[2].x, r[2].y, r[2].z); in->avoid_slot = __float_as_int(r[2].w);
    }
    PTB_D void store(int p, const HitRec& h, V3 c) const {
        if (h.hit) return;
        float* r = reinterpret_cast<float*>(&result[p]);
        atomicAdd(r, c.x); atomicAdd(r + 1, c.y); atomicAdd(r + 2, c.z);
    }
};
// ---- parity taps: the same record formats (written by k_pack_tap), results into flat arrays -------------------------------------------
template <bool ANYHIT>
struct TapIO {
    static constexpr bool kAnyHit = ANYHIT;
    static constexpr int K = ANYHIT ? 3 : 2;
    const float4* __restrict__ q0; const float4* __restrict__ q1; const float4* __restrict__ q2;
    int* hit; float* depth; int* index; float* uv;
    PTB_D const float4* rec(int k) const { return k == 0 ? q0 : (k == 1 ? q1 : q2); }
    PTB_D void decode(const float4* r, RayIn* in) const {
        in->item = __float_as_int(r[0].w); in->ro = mk3(r[0].x, r[0].y, r[0].z);
        in->rd = mk3(r[1].x, r[1].y, r[1].z); in->c = v3s(0.0f);
        if (ANYHIT) { in->tmax = r[1].w; in->avoid_slot = __float_as_int(r[K - 1].w); }
        else { in->tmax = PTB_INF; in->avoid_slot = __float_as_int(r[1].w); }
    }
    PTB_D void store(int i, const HitRec& h, V3) const {
        hit[i] = h.hit;
        if (!ANYHIT) { depth[i] = h.depth; index[i] = h.index; uv[2 * i] = h.u; uv[2 * i + 1] = h.v; }
    }
};

// record `i` of the queue straight from global memory (one-ray-per-thread kernels)
template <class IO>
PTB_D void fetch_direct(const IO& io, int i, RayIn* in) {
    float4 r[IO::K];
#pragma unroll
    for (int k = 0; k < IO::K; k++) r[k] = io.rec(k)[i];
    io.decode(r, in);
}

template <bool COUNT>
PTB_D void flush_counters(const TraceCounters& C, unsigned long long nrays, bool shadow, DevCounters* ctr) {
    if (!COUNT) return;
    unsigned long long a = C.nodes, b = C.boxes, t = C.tris, r = nrays; unsigned int m = C.max_stack;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o);
        r += __shfl_xor_sync(0xffffffffu, r, o); m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&ctr->nodes, a); atomicAdd(&ctr->boxes, b); atomicAdd(&ctr->tris, t);
        atomicAdd(shadow ? &ctr->shadow_rays : &ctr->extend_rays, r); atomicMax(&ctr->max_stack, m);
    }
}

// one ray per thread, chunks of 32 per warp.  POLICY 0: the reference's literal order; 1: trace_ordered (exact slab tests)
template <class IO, int POLICY, bool COUNT>
__global__ void __launch_bounds__(PTB_TRACE_BLK) k_trace_simple(TraceScene S, IO io, int* cursor, const int* count_ptr, DevCounters* ctr) {
    const int lane = threadIdx.x & 31;
    const int count = *count_ptr;
    TraceCounters C; C.nodes = C.boxes = C.tris = 0; C.max_stack = 0;
    unsigned long long nrays = 0;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        int idx = base + lane;
        if (idx < count) {
            RayIn in;
            fetch_direct(io, idx, &in);
            const int avoid = in.avoid_slot >= 0 ? S.leaf[in.avoid_slot & PTB_SLOT_MASK] : -1;
            HitRec h;
            if (POLICY == 0) {
                h = trace_reference<COUNT>(S, in.ro, in.rd, avoid, &C);
                if (IO::kAnyHit) h.hit = !(h.hit == 0 || h.depth > in.tmax);       // path.py:50  occ.hit == 0 or occ.depth > li.dis
            } else {
                h = trace_ordered<IO::kAnyHit, COUNT>(S, in.ro, in.rd, avoid, in.tmax, &C);
            }
            io.store(in.item, h, in.c);
            if (COUNT) nrays++;
        }
    }
    flush_counters<COUNT>(C, nrays, IO::kAnyHit, ctr);
}

// ---- the always-test list (shared-memory copy: slots, packed triangles, gate boxes) ------------------------------------------------
// The reference tests a listed triangle only if its gate passes Box.intersect (conservative test, exact when the ray grazes the gate).
PTB_D bool list_gate_ok(const TraceScene& S, const float4 (*s_gate)[2], int j, int slot, const RayCons& R, V3 ro, V3 rd) {
    const float4 glo = s_gate[j][0], ghi = s_gate[j][1];
    float gl; bool gsure;
    return slab_cons2(glo.x, glo.y, glo.z, ghi.x, ghi.y, ghi.z, R, &gl, &gsure) && (gsure || gate_passes(S, S.gate[slot], ro, rd));
}
// The list is scanned in two phases.  Phase 1 (list_candidates, uniform control flow: every lane looks at the same box, broadcast
// shared-memory reads, no branches): the conservative slab test of the entries' INFLATED bounds -- the leaf bound of ptb_traverse.cuh,
// the same test a tree leaf gets from its parent's node step -- marks the entries whose triangle the ray can possibly be accepted by
// (an ill-conditioned entry has no bound and is always marked).  Inside a room that is the two triangles of the wall the ray leaves
// through, not all ten.  The two triangles of an axis-aligned wall quad have the SAME bounds, so the block's prologue (list_setup)
// reduces the list to its distinct boxes, each with the bit mask of the entries it stands for: five slab tests for the ten walls of a
// Cornell room, and a passing box ORs its mask in (no per-entry work; the entry to avoid is cleared afterwards from the tag in the
// avoid field, PTB_SLOT_LIST_SHIFT).  Phase 2 (list_scan): every lane walks its OWN marked entries (per-lane shared-memory index) with
// the exact triangle test.  Before, all lanes ran the exact test on all entries, and its early exits (plane behind the ray, beyond the
// best hit) left 20 of 32 lanes active in the long part.
#ifndef PTB_LIST_FILTER
#define PTB_LIST_FILTER 1           /* 0: every entry is a candidate (A/B builds) */
#endif
struct ListBoxes { float4 box[PTB_LIST_CAP][2]; unsigned mask[PTB_LIST_CAP]; int nuniq; unsigned must, all; };
// one warp of the block: distinct boxes of s_box[0..nlist) in order of first appearance, and for each the entries that share it
PTB_D void list_setup(const float4 (*s_box)[2], int nlist, ListBoxes* U) {
    const int j = threadIdx.x;                           // called by threads 0..31
    const bool in = j < nlist;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (in) { a = s_box[j][0]; b = s_box[j][1]; }
    int g = j;
    for (int k = 0; k < j && in; k++) {
        const float4 c = s_box[k][0], d = s_box[k][1];
        if (a.x == c.x && a.y == c.y && a.z == c.z && b.x == d.x && b.y == d.y && b.z == d.z) { g = k; break; }
    }
    const unsigned leaders = __ballot_sync(0xffffffffu, in && g == j);
    const unsigned must = __ballot_sync(0xffffffffu, in && (__float_as_int(a.w) & PTB_TF_MUST) != 0);
    if (in && g == j) {
        const int u = __popc(leaders & ((1u << j) - 1u));
        U->box[u][0] = a; U->box[u][1] = b; U->mask[u] = 0u;
    }
    __syncwarp();
    if (in) atomicOr(&U->mask[__popc(leaders & ((1u << g) - 1u))], 1u << j);
    if (j == 0) { U->nuniq = __popc(leaders); U->must = must; U->all = nlist >= 32 ? 0xffffffffu : (1u << nlist) - 1u; }
}
PTB_D unsigned list_candidates(const ListBoxes& U, int avoid_slot, const RayCons& R, const RayTrav& Q, float cull) {
    unsigned cand = PTB_LIST_FILTER ? U.must : U.all;
    if (PTB_LIST_FILTER) {
        const int nu = U.nuniq;
        for (int k = 0; k < nu; k++) {
            float lb;
            const bool ok = slab_trav(U.box[k][0], U.box[k][1], R, Q, &lb) && !(lb > cull);
            if (ok) cand |= U.mask[k];
        }
    }
    const int tag = avoid_slot >> PTB_SLOT_LIST_SHIFT;  // 0 (not listed, or no triangle to avoid: -1 >> 24 = -1 is excluded below) or list index + 1
    if (avoid_slot >= 0 && tag != 0) cand &= ~(1u << (tag - 1));
    return cand;
}
// Closest accepted triangle among the marked entries (ties: larger slot) into ret / best; ANYHIT: stop at the first one.  GATED = false
// leaves the gates out: the winner of the ungated pass is the winner of the gated one whenever its own gate passes (it is the
// minimum over a superset), which the caller checks once per ray instead of once per candidate -- the gates of listed triangles are
// boxes near the root, so the gated pass is a rare second pass.  *jw = list index of the winner.
template <bool ANYHIT, bool GATED, bool COUNT>
PTB_D bool list_scan(const TraceScene& S, const int* s_slot, const float4 (*s_tri)[4], const float4 (*s_gate)[2], unsigned cand, const RayIn& in, const RayCons& R,
                     HitRec& ret, float& best, int* jw, TraceCounters& C) {
    while (cand != 0u) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1u;
        const int slot = s_slot[j];
        Tri64 T; T.a = s_tri[j][0]; T.b = s_tri[j][1]; T.c = s_tri[j][2]; T.d = s_tri[j][3];
        if (COUNT) C.tris++;
        float dep, s, t;
        if (tri_fast(T, in.ro, in.rd, best, &dep, &s, &t)) {
            const bool better = ANYHIT ? dep < PTB_INF : (dep < ret.depth || (ret.hit && slot > ret.slot));
            if (better && (!GATED || list_gate_ok(S, s_gate, j, slot, R, in.ro, in.rd))) {
                ret.depth = dep; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = slot; best = dep; *jw = j;
                if (ANYHIT) return true;
            }
        }
    }
    return false;
}

// ---- production phase A: per-ray setup, always-test list, root test; survivors go to the tree queue ---------------------------------
template <class IO, bool COUNT>
__global__ void __launch_bounds__(256) k_trace_pre(TraceScene S, IO io, const int* count_ptr, ExpQ xq, int* tree_count, DevCounters* ctr) {
    constexpr bool ANYHIT = IO::kAnyHit;
    const int count = *count_ptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TraceCounters C; C.nodes = C.boxes = C.tris = 0; C.max_stack = 0;
    unsigned long long nrays = 0;
    __shared__ int s_warp[2][8], s_base[2];
    int par = 0;
    // the always-test list in shared memory: slots, packed triangles, gate boxes (broadcast reads, no dependent global loads)
    __shared__ int s_slot[PTB_LIST_CAP];
    __shared__ float4 s_tri[PTB_LIST_CAP][4];
    __shared__ float4 s_gate[PTB_LIST_CAP][2];
    __shared__ float4 s_box[PTB_LIST_CAP][2];          // inflated bounds (lo.w: PTB_TF_* flags)
    __shared__ ListBoxes s_uniq;                       // their distinct boxes (list_setup)
    const int nlist = min(S.nlist, PTB_LIST_CAP);
    for (int j = threadIdx.x; j < nlist * 4; j += 256) {
        const int slot = S.list[j >> 2];
        s_tri[j >> 2][j & 3] = reinterpret_cast<const float4*>(S.tris + slot)[j & 3];
        if ((j & 3) < 2) s_gate[j >> 2][j & 3] = S.gbox[2 * slot + (j & 3)];
        if ((j & 3) == 2) s_box[j >> 2][0] = S.tlo[slot];
        if ((j & 3) == 3) s_box[j >> 2][1] = S.thi[slot];
        if ((j & 3) == 0) s_slot[j >> 2] = slot;
    }
    __syncthreads();
    if (threadIdx.x < 32) list_setup(s_box, nlist, &s_uniq);
    __syncthreads();
    const int rounded = (count + 255) & ~255;          // every thread of a block runs the same number of iterations
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < rounded; idx += gridDim.x * 256) {
        bool live = false;
        RayIn in; in.item = -1; in.ro = v3s(0.0f); in.rd = v3s(0.0f); in.avoid_slot = -1; in.tmax = 0.0f; in.c = v3s(0.0f);
        RayCons R; R.o = v3s(0.0f); R.d = v3s(0.0f); R.r = v3s(0.0f); R.nc = v3s(0.0f); R.a2 = 0.0f;
        float delta = 0.0f;
        HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
        if (idx < count) {
            {   // start the next iteration's record on its way (the loop is otherwise serialised on this DRAM round trip)
                const int nxt = idx + gridDim.x * 256;
                if (nxt < count) {
#pragma unroll
                    for (int k = 0; k < IO::K; k++) asm volatile("prefetch.global.L1 [%0];" ::"l"(io.rec(k) + nxt));
                }
            }
            fetch_direct(io, idx, &in);
            if (COUNT) nrays++;
            if (ray_is_special(in.ro, in.rd)) {
                // axis-parallel / non-finite: exact tests over the reference's own arrays (rare)
                const int avoid = in.avoid_slot >= 0 ? S.leaf[in.avoid_slot & PTB_SLOT_MASK] : -1;
                ret = trace_ordered<ANYHIT, COUNT>(S, in.ro, in.rd, avoid, in.tmax, &C);
                io.store(in.item, ret, in.c);
            } else {
                R = ray_cons(in.ro, in.rd);
                float best = ANYHIT ? fminf(in.tmax, PTB_INF) : PTB_INF;
                const float best0 = best;
                int jw = -1;
                delta = trav_delta(in.ro, S.scene_abs);
                const RayTrav Q = ray_trav(R, delta);
                const unsigned cand = list_candidates(s_uniq, in.avoid_slot, R, Q, best + best * PTB_CULL_GUARD);
                if (COUNT) C.boxes += nlist;
                bool occluded = list_scan<ANYHIT, false, COUNT>(S, s_slot, s_tri, s_gate, cand, in, R, ret, best, &jw, C);
                if (ret.hit && !list_gate_ok(S, s_gate, jw, ret.slot, R, in.ro, in.rd)) {
                    // the ungated winner is not a triangle the reference would have tested: start over, gate checked per candidate
                    ret.hit = 0; ret.depth = PTB_INF; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1; best = best0;
                    TraceCounters unused; unused.nodes = unused.boxes = unused.tris = 0; unused.max_stack = 0;
                    occluded = list_scan<ANYHIT, true, false>(S, s_slot, s_tri, s_gate, cand, in, R, ret, best, &jw, unused);
                }
                // anything left in the tree?  its root box = union of the inflated bounds of every triangle not in the list
                if (!occluded && S.n >= 2) {
                    float lb;
                    if (COUNT) C.boxes++;
                    live = S.nlo[0].x <= S.nhi[0].x && slab_trav(S.nlo[0], S.nhi[0], R, Q, &lb) && (S.root_must || !(lb > best + best * PTB_CULL_GUARD));
                }
                if (!live) {
                    if (ret.hit) ret.index = S.leaf[ret.slot];
                    io.store(in.item, ret, in.c);
                }
            }
        }
        // append the survivors to the tree queue: one atomic per block (every launch hammers the same counter), contiguous records
        // (two sets of buffers used alternately: no third barrier, see block_append)
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_warp[par][warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const int c = s_warp[par][w]; s_warp[par][w] = tot; tot += c; }
            s_base[par] = tot ? atomicAdd(tree_count, tot) : 0;
        }
        __syncthreads();
        const int wbase = s_base[par] + s_warp[par][warp];
        par ^= 1;
        if (live) {
            const int pos = wbase + __popc(m & ((1u << lane) - 1u));
            xq.e[0][pos] = make_float4(in.ro.x, in.ro.y, in.ro.z, __int_as_float(in.item));
            xq.e[1][pos] = make_float4(in.rd.x, in.rd.y, in.rd.z, __int_as_float(in.avoid_slot));
            xq.e[2][pos] = make_float4(R.r.x, R.r.y, R.r.z, delta);
            xq.e[3][pos] = make_float4(R.nc.x, R.nc.y, R.nc.z, 0.0f);
            xq.e[4][pos] = ANYHIT ? make_float4(in.c.x, in.c.y, in.c.z, in.tmax) : make_float4(ret.depth, ret.u, ret.v, __int_as_float(ret.hit ? ret.slot : -1));
        }
    }
    flush_counters<COUNT>(C, nrays, ANYHIT, ctr);
}

// ---- cp.async helpers (16-byte global -> shared copies, per-thread groups) --------------------------------------------------------------
PTB_D void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
PTB_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> PTB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// 32-byte quantised node (Node32, ptb_traverse.cuh): one 256-bit load
PTB_D void ld_qnode256(const uint4* p, uint4* a, uint4* b) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z), "=r"(b->w) : "l"(p));
}

// 64-byte triangle record of a big tree (global-memory variant): two 256-bit loads that do not allocate in L1 -- a ray tests ~2 leaf
// triangles out of a million, so the line is not coming back, and the L1 is better spent on the top levels of the node array
#ifndef PTB_STREAM_TRIS
#define PTB_STREAM_TRIS 1
#endif
PTB_D Tri64 ld_tri_stream(const Tri64* p) {
    Tri64 T;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(T.a.x), "=f"(T.a.y), "=f"(T.a.z), "=f"(T.a.w), "=f"(T.b.x), "=f"(T.b.y), "=f"(T.b.z), "=f"(T.b.w) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(T.c.x), "=f"(T.c.y), "=f"(T.c.z), "=f"(T.c.w), "=f"(T.d.x), "=f"(T.d.y), "=f"(T.d.z), "=f"(T.d.w) : "l"(reinterpret_cast<const char*>(p) + 32));
    return T;
}

#define PTB_PQ 4                    /* pending-leaf ring entries per lane */
#ifndef PTB_VOTE_N
#define PTB_VOTE_N 1                /* node step iff PTB_VOTE_N * (lanes ready for one) > PTB_VOTE_L * (lanes with a pending leaf) */
#define PTB_VOTE_L 2
#endif
#ifndef PTB_NODE_REPS
#define PTB_NODE_REPS 1             /* node steps per vote */
#endif
#ifndef PTB_TRACE_BLK_S
#define PTB_TRACE_BLK_S 768         /* threads of the shared-memory-resident variant (one CTA per SM) */
#endif
#ifndef PTB_TRACE_MINB
#define PTB_TRACE_MINB 7            /* resident CTAs per SM the global-memory variant is compiled for (72 registers) */
#endif

// The traversal stack: its first PTB_SSTACK entries live in shared memory (column = thread: no bank conflicts), deeper ones in local
// memory.  As a local array alone (round 1) the stack was 37-39 % of the kernel's L1 sector traffic -- lanes sit at different depths,
// so a pop pulls a 32-byte sector for 8 bytes (ncu: 2.5-2.9 bytes used per sector) -- with an L1 hit rate of 31 % (config 4) / 63 %
// (config 2), i.e. one L2 round trip per two or three pops, and it evicted the nodes and triangles the L1 is there for.
#ifndef PTB_SSTACK_S
#define PTB_SSTACK_S 8              /* shared-memory stack entries per thread, resident variants (one 768-thread CTA per SM) */
#endif
#ifndef PTB_SSTACK_G
#define PTB_SSTACK_G 8              /* the same for the global-memory variant (7 CTAs of 128 threads per SM) */
#endif
// dynamic shared memory layout of k_trace_tree (bytes), shared by the kernel and the host launch code
template <int BLK, bool W4 = false>
struct TraceSmem {
    static constexpr int sstack = BLK == PTB_TRACE_BLK ? PTB_SSTACK_G : PTB_SSTACK_S;
    static constexpr int pq = W4 ? 2 * PTB_PQ : PTB_PQ;     // pending-leaf ring entries per lane (a 4-wide node step can add four)
    static constexpr size_t tile = (size_t)(BLK / 32) * 2 * PTB_EXP_K * PTB_TILE * sizeof(float4);
    static constexpr size_t queue = (size_t)pq * BLK * (sizeof(int) + sizeof(float));
    static constexpr size_t stack = (size_t)sstack * BLK * sizeof(unsigned long long);
    static constexpr size_t fixed = tile + queue + stack;
    __host__ __device__ static size_t bvh(int n) { return (size_t)(n > 1 ? n - 1 : 0) * sizeof(Node64); }
    __host__ __device__ static size_t qbvh(int n) { return (size_t)(n > 1 ? n - 1 : 0) * 32; }          // quantised (Node32)
};

// ---- production phase B: traversal of the tree for the rays k_trace_pre queued ----------------------------------------------------
// SMEM = true: the nodes of the traversal tree are copied into shared memory by each CTA (one CTA of PTB_TRACE_BLK_S threads per
// SM), as four 16-byte "quarter" arrays of the packed 64-byte nodes (QN = false) or, for trees up to twice that size, two 16-byte
// halves of the quantised 32-byte nodes (QN = true), so that a divergent node fetch costs shared-memory wavefronts instead of one
// L1 wavefront per lane and 16-byte load.  Triangles, gate boxes and the local-memory stack stay behind L1, which keeps ~100 KB
// next to the carve-out.  SMEM = false (bigger trees): quantised nodes out of global memory, one 256-bit load per step, several
// CTAs per SM to cover the L2 latency.
// W4 (global-memory variant only): 4-wide quantised nodes (lbvh.cu k_wide4_nodes: a binary node with its internal children replaced by
// their children, 64 bytes) -- half as many dependent node fetches per ray, which is what the big-scene kernel waits for.
// S16 (resident variants, traversal trees of height <= PTB_S16_DEPTH with fewer than 65536 nodes): the whole stack lives in the
// shared-memory region as 32-bit entries -- node id in the low half, the entry distance cut to its upper 16 bits (bfloat16, rounded
// toward zero after a clamp at 0: still a lower bound of any depth below, which is all a popped entry's distance is used for) -- so
// a push is one STS, a pop one LDS, and the local-memory tail of the stack with its two branches is gone.
#define PTB_S16_DEPTH 16
template <class IO, bool COUNT, int BLK, bool SMEM, bool QN, bool W4 = false, bool S16 = false>
__global__ void __launch_bounds__(BLK, SMEM ? 1 : PTB_TRACE_MINB) k_trace_tree(TraceScene S, IO io, ExpQ xq, int* cursor, const int* count_ptr, DevCounters* ctr) {
    static_assert(!W4 || (!SMEM && QN), "4-wide nodes: global-memory quantised variant only");
    static_assert(!S16 || (SMEM && !W4 && PTB_S16_DEPTH * sizeof(unsigned) <= TraceSmem<BLK, W4>::sstack * sizeof(unsigned long long)), "S16: resident variants, same shared-memory region");
    using Smem = TraceSmem<BLK, W4>;
    constexpr int PQ = Smem::pq;
    constexpr bool ANYHIT = IO::kAnyHit;
    constexpr int K = PTB_EXP_K;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int count = *count_ptr;
    const int n = S.n;
    TraceCounters C; C.nodes = C.boxes = C.tris = 0; C.max_stack = 0;

    extern __shared__ __align__(16) unsigned char s_raw[];
    // staged ray records: [warp][buffer][k][record]
    float4 (*s_tile)[2][K][PTB_TILE] = reinterpret_cast<float4 (*)[2][K][PTB_TILE]>(s_raw);
    // pending leaves: ring of PTB_PQ (slot, lower bound of the depth) per lane, column = thread (no bank conflicts)
    int (*s_qslot)[BLK] = reinterpret_cast<int (*)[BLK]>(s_raw + Smem::tile);
    float (*s_qnear)[BLK] = reinterpret_cast<float (*)[BLK]>(s_raw + Smem::tile + (size_t)PQ * BLK * sizeof(int));
    // traversal stack, entries [0, SSTACK): column = thread
    constexpr int SSTACK = Smem::sstack;
    unsigned long long (*s_stack)[BLK] = reinterpret_cast<unsigned long long (*)[BLK]>(s_raw + Smem::tile + Smem::queue);
    unsigned (*s_stack32)[BLK] = reinterpret_cast<unsigned (*)[BLK]>(s_raw + Smem::tile + Smem::queue);        // S16: the same region
    // resident nodes: quarter q of node i at s_node[q * (n-1) + i]
    float4* s_node = reinterpret_cast<float4*>(s_raw + Smem::fixed);
    if (SMEM) {
        if (blockIdx.x * PTB_TILE >= count) return;        // more CTAs than tiles of work: skip the copy
        if constexpr (QN) {                                 // quantised nodes: two 16-byte halves per node
            const float4* gn = reinterpret_cast<const float4*>(S.qnodes);
            for (int i = threadIdx.x; i < 2 * (n - 1); i += BLK) s_node[(i & 1) * (n - 1) + (i >> 1)] = gn[i];
        } else {
            const float4* gn = reinterpret_cast<const float4*>(S.nodes);
            for (int i = threadIdx.x; i < 4 * (n - 1); i += BLK) s_node[(i & 3) * (n - 1) + (i >> 2)] = gn[i];
        }
        __syncthreads();
    }

    // ---- warp-uniform staging state ----
    int tile_n0 = 0, tile_n1 = 0;              // number of valid records per buffer
    int cur_buf = 0, tile_pos = 0;             // consumption cursor in the current buffer
    bool tile_ready = false, exhausted = false;
    int pool_next = 0, pool_end = 0;           // queue positions reserved by this warp (PTB_POOL per atomic on the shared cursor)
    auto fill = [&](int b) {
        if (pool_next >= pool_end) {
            int p = 0;
            if (lane == 0) p = atomicAdd(cursor, PTB_POOL);
            p = __shfl_sync(FULL, p, 0);
            pool_next = p; pool_end = p + PTB_POOL;
        }
        const int base = pool_next;
        pool_next += PTB_TILE;
        const int nv = max(0, min(PTB_TILE, count - base));
#pragma unroll
        for (int e = lane; e < K * PTB_TILE; e += 32) {
            const int k = e / PTB_TILE, i = e % PTB_TILE;
            if (i < nv) cp_async16(&s_tile[warp][b][k][i], xq.e[k] + base + i);
        }
        cp_async_commit();
        if (b == 0) tile_n0 = nv; else tile_n1 = nv;
    };
    fill(0); fill(1);

    // ---- per-lane ray state ----
    int item = -1, avoid_slot = -1;            // item >= 0: this lane owns a ray (its result is stored when the lane next goes idle)
    RayCons R; R.o = v3s(0.0f); R.d = v3s(0.0f); R.r = v3s(0.0f); R.nc = v3s(0.0f); R.a2 = 0.0f;
    RayTrav Q; Q.nc1 = v3s(0.0f); Q.nc2 = v3s(0.0f);   // slab constants of the traversal-tree boxes (ray_trav)
    RayQuant G; G.A = v3s(0.0f); G.Bn = v3s(0.0f); G.Bf = v3s(0.0f); G.sx = G.sy = G.sz = PTB_QSEL_LO;      // the same for quantised nodes (ray_quant)
    V3 contrib = v3s(0.0f);
    float best = 0.0f, cull = 0.0f;
    HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
    // stack entries: low word = internal node, high word = lower bound of the depth of anything below it.  A pop is issued as soon
    // as a lane runs out of children (end of its node step) into `pe`, and consumed at the start of its next node step, so the
    // local-memory load has a whole iteration to land.
    unsigned long long stack[(!S16 && PTB_STACK - SSTACK > 0) ? PTB_STACK - SSTACK : 1];      // entries beyond the shared-memory part
    unsigned long long pe = 0;                 // popped entry, valid iff `popped`
    bool popped = false;
    int sp = 0;
    int cur = -1;                              // node to visit next; -1: none (take `pe` if popped)
    float cur_near = 0.0f;
    int q_head = 0, q_count = 0;

    while (true) {
        // a lane can take a node step if it has a node (current or stacked) and room for two more pending leaves; a leaf step if
        // a leaf is pending.  A lane with neither is idle: finished (result not stored yet) or never started.
        const bool node_ok = (cur != -1 || popped) && q_count <= PQ - (W4 ? 4 : 2);
        const bool leaf_ok = q_count > 0;
        const unsigned mn = __ballot_sync(FULL, node_ok), ml = __ballot_sync(FULL, leaf_ok);
        const unsigned idle = ~(mn | ml);
        if (!exhausted && (__popc(idle) >= (SMEM ? PTB_FETCH_MIN : PTB_FETCH_MIN_G) || idle == FULL)) {
            // ---- idle lanes store their result and take the next staged records ----------------------------------------------------------
            if (!tile_ready) { cp_async_wait<1>(); __syncwarp(); tile_ready = true; }     // the older of the two groups in flight has landed
            const int nv = cur_buf == 0 ? tile_n0 : tile_n1;
            if (nv == 0) exhausted = true;                                                 // positions are handed out in order: nothing is left
            const bool me = (idle >> lane) & 1u;
            if (me && item >= 0) { io.store(item, ret, contrib); item = -1; }
            const int k = tile_pos + __popc(idle & lt_mask);
            if (me && k < nv) {
                const float4 e0 = s_tile[warp][cur_buf][0][k], e1 = s_tile[warp][cur_buf][1][k], e2 = s_tile[warp][cur_buf][2][k],
                             e3 = s_tile[warp][cur_buf][3][k], e4 = s_tile[warp][cur_buf][4][k];
                item = __float_as_int(e0.w); avoid_slot = __float_as_int(e1.w);
                R.o = mk3(e0.x, e0.y, e0.z); R.d = mk3(e1.x, e1.y, e1.z); R.r = mk3(e2.x, e2.y, e2.z); R.nc = mk3(e3.x, e3.y, e3.z);
                R.a2 = __fmaf_rn(fmaxf(fmaxf(fabsf(R.nc.x), fabsf(R.nc.y)), fabsf(R.nc.z)), PTB_CONS_KAPPA, 1e-30f);
                Q = ray_trav(R, e2.w);
                if constexpr (QN) G = ray_quant(S, R, Q);           // quantised nodes: R.r and Q are not used past this point
                ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
                if (ANYHIT) { contrib = mk3(e4.x, e4.y, e4.z); best = fminf(e4.w, PTB_INF); }
                else {                                                      // provisional closest hit over the always-test list
                    ret.slot = __float_as_int(e4.w); ret.hit = ret.slot >= 0;
                    ret.depth = e4.x; ret.u = e4.y; ret.v = e4.z;
                    if (ret.hit) ret.index = S.leaf[ret.slot];
                    best = ret.depth;
                }
                cull = best + best * PTB_CULL_GUARD;
                sp = 0; popped = false; cur_near = 0.0f; q_head = 0; q_count = 0;
                cur = 0;                // the root (its children's boxes are tested in its node step)
            }
            tile_pos = min(tile_pos + __popc(idle), nv);
            if (tile_pos >= nv && !exhausted) {        // tile consumed: switch to the prefetched one and refill this buffer
                __syncwarp();
                fill(cur_buf);
                cur_buf ^= 1; tile_pos = 0; tile_ready = false;
            }
            continue;
        }
        if ((mn | ml) == 0u) break;                    // exhausted and nothing in flight
        if (__popc(mn) * PTB_VOTE_N > __popc(ml) * PTB_VOTE_L) {
            // ---- node step: one 64-byte node, both children's slab tests ----------------------------------------------------------------------
#pragma unroll 1
            for (int rep = 0; rep < PTB_NODE_REPS; rep++)
            if (rep == 0 ? node_ok : ((cur != -1 || popped) && q_count <= PQ - (W4 ? 4 : 2))) {
                if (cur == -1) {
                    if constexpr (S16) { cur = (int)((unsigned)pe & 0xffffu); cur_near = __uint_as_float((unsigned)pe & 0xffff0000u); }
                    else { cur = (int)(unsigned)pe; cur_near = __int_as_float((int)(pe >> 32)); }
                    popped = false;
                }
                if (cur_near > cull) cur = -1;          // nothing below can beat the best hit found since it was pushed / chosen
                else if constexpr (W4) {
                    uint4 qa, qb, qc, qd;
                    ld_qnode256(S.wnodes + 4 * (size_t)cur, &qa, &qb); ld_qnode256(S.wnodes + 4 * (size_t)cur + 2, &qc, &qd);
                    const unsigned bw[12] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w, qc.x, qc.y, qc.z, qc.w};
                    const int wid[4] = {(int)qd.x, (int)qd.y, (int)qd.z, (int)qd.w};
                    if (COUNT) { C.nodes++; C.boxes += 4; }
                    float nn[4]; int cc[4];
                    int m = 0;                              // internal children hit
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float nk;
                        bool hk = slab_quant(bw[3 * k], bw[3 * k + 1], bw[3 * k + 2], G, R.a2, &nk);
                        hk = hk && wid[k] >= 0;
                        const int ck = wid[k] & PTB_NODE_ID;
                        if (wid[k] & PTB_NODE_MUST) nk = 0.0f;        // an ill-conditioned triangle below: no distance bound holds
                        hk = hk && !(nk > cull);
                        const bool leafk = ck < n;
                        if (hk && leafk && ck != avoid_slot) { const int q = (q_head + q_count) & (PQ - 1); s_qslot[q][threadIdx.x] = ck; s_qnear[q][threadIdx.x] = nk; q_count++; }
                        const bool dk = hk && !leafk;
                        nn[k] = dk ? nk : 3.0e38f; cc[k] = ck - n;
                        m += dk ? 1 : 0;
                    }
                    // nearest first: sort the four (near, node) pairs (misses carry +huge), 5 compare-exchanges
#define PTB_CSWAP(a, b) { const bool sw = nn[b] < nn[a]; const float tn = sw ? nn[a] : nn[b]; const int tc = sw ? cc[a] : cc[b]; nn[a] = sw ? nn[b] : nn[a]; cc[a] = sw ? cc[b] : cc[a]; nn[b] = tn; cc[b] = tc; }
                    PTB_CSWAP(0, 1) PTB_CSWAP(2, 3) PTB_CSWAP(0, 2) PTB_CSWAP(1, 3) PTB_CSWAP(1, 2)
#undef PTB_CSWAP
                    // farther ones onto the stack, farthest first; the nearest is next (a proper tree of height <= 40 cannot overflow: checked by the host)
#pragma unroll
                    for (int k = 3; k >= 1; k--) if (k < m) {
                        const unsigned long long entry = ((unsigned long long)(unsigned)__float_as_int(nn[k]) << 32) | (unsigned)cc[k];
                        if (sp < SSTACK) s_stack[sp][threadIdx.x] = entry; else stack[sp - SSTACK] = entry;
                        sp++;
                    }
                    if (COUNT) C.max_stack = max(C.max_stack, (unsigned)sp);
                    if (m > 0) { cur = cc[0]; cur_near = nn[0]; } else cur = -1;
                } else {
                    int w0, w1;                          // child ids; -1: nothing below
                    float n0, n1;
                    bool h0, h1;
                    if constexpr (!QN) {
                        Node64 N;
                        N.a = s_node[cur]; N.b = s_node[(n - 1) + cur]; N.c = s_node[2 * (n - 1) + cur]; N.d = s_node[3 * (n - 1) + cur];
                        w0 = __float_as_int(N.a.w); w1 = __float_as_int(N.b.w);
                        h0 = slab_trav(N.a, N.b, R, Q, &n0); h1 = slab_trav(N.c, N.d, R, Q, &n1);
                    } else {
                        uint4 qa, qb;
                        if constexpr (SMEM) { qa = reinterpret_cast<const uint4*>(s_node)[cur]; qb = reinterpret_cast<const uint4*>(s_node)[(n - 1) + cur]; }
                        else ld_qnode256(S.qnodes + 2 * cur, &qa, &qb);
                        w0 = (int)qb.z; w1 = (int)qb.w;
                        h0 = slab_quant(qa.x, qa.y, qa.z, G, R.a2, &n0);
                        h1 = slab_quant(qa.w, qb.x, qb.y, G, R.a2, &n1);
                    }
                    if (COUNT) { C.nodes++; C.boxes += 2; }
                    h0 = h0 && w0 >= 0; h1 = h1 && w1 >= 0;
                    const int c0 = w0 & PTB_NODE_ID, c1 = w1 & PTB_NODE_ID;
                    const bool must0 = (w0 & PTB_NODE_MUST) != 0, must1 = (w1 & PTB_NODE_MUST) != 0;
                    if (must0) n0 = 0.0f;               // an ill-conditioned triangle below: no distance bound holds
                    if (must1) n1 = 0.0f;
                    h0 = h0 && !(n0 > cull); h1 = h1 && !(n1 > cull);
                    const bool leaf0 = c0 < n, leaf1 = c1 < n;
                    const bool p0 = h0 && leaf0 && c0 != avoid_slot, p1 = h1 && leaf1 && c1 != avoid_slot;
                    if (p0) { const int k = (q_head + q_count) & (PQ - 1); s_qslot[k][threadIdx.x] = c0; s_qnear[k][threadIdx.x] = n0; q_count++; }
                    if (p1) { const int k = (q_head + q_count) & (PQ - 1); s_qslot[k][threadIdx.x] = c1; s_qnear[k][threadIdx.x] = n1; q_count++; }
                    const bool d0 = h0 && !leaf0, d1 = h1 && !leaf1;
                    const bool first1 = !(n0 < n1);     // nearer first
                    if (d0 && d1) {
                        // a proper tree of height <= PTB_STACK cannot overflow (checked at build time)
                        if constexpr (S16) {
                            s_stack32[sp][threadIdx.x] = (__float_as_uint(fmaxf(first1 ? n0 : n1, 0.0f)) & 0xffff0000u) | (unsigned)((first1 ? c0 : c1) - n);
                        } else {
                            const unsigned long long entry = ((unsigned long long)(unsigned)__float_as_int(first1 ? n0 : n1) << 32) | (unsigned)((first1 ? c0 : c1) - n);
                            if (sp < SSTACK) s_stack[sp][threadIdx.x] = entry; else stack[sp - SSTACK] = entry;
                        }
                        sp++;
                        if (COUNT) C.max_stack = max(C.max_stack, (unsigned)sp);
                        cur = (first1 ? c1 : c0) - n; cur_near = first1 ? n1 : n0;
                    } else if (d0) { cur = c0 - n; cur_near = n0; }
                    else if (d1) { cur = c1 - n; cur_near = n1; }
                    else cur = -1;
                }
                if (cur == -1 && sp > 0) {                                  // out of children: issue the pop now
                    --sp;
                    if constexpr (S16) pe = s_stack32[sp][threadIdx.x];
                    else pe = sp < SSTACK ? s_stack[sp][threadIdx.x] : stack[sp - SSTACK];
                    popped = true;
                }
            }
        } else {
            // ---- leaf step: the oldest pending leaf of every lane that has one ---------------------------------------------------------------
            if (leaf_ok) {
                const int slot = s_qslot[q_head][threadIdx.x]; const float lnear = s_qnear[q_head][threadIdx.x];
                q_head = (q_head + 1) & (PQ - 1); q_count--;
                if (!(lnear > cull)) {
                    const Tri64 T = (!SMEM && PTB_STREAM_TRIS) ? ld_tri_stream(S.tris + slot) : S.tris[slot];
                    if (COUNT) C.tris++;
                    float dep, s, t;
                    if (tri_fast(T, R.o, R.d, best, &dep, &s, &t)) {
                        // here dep <= best.  The reference tests this triangle only if its gate box passes Box.intersect: conservative
                        // test (certain pass / certain miss), exact test when the ray grazes the gate
                        const bool better = ANYHIT ? dep < PTB_INF : (dep < ret.depth || (ret.hit && slot > ret.slot));
                        if (better) {
                            const float4 glo = S.gbox[2 * slot], ghi = S.gbox[2 * slot + 1];
                            float gl; bool gsure;
                            RayCons Rg = R;
                            if constexpr (QN) Rg.r = mk3(G.A.x * S.qinv[0], G.A.y * S.qinv[1], G.A.z * S.qinv[2]);     // exact: A = qext * r, qext = 2^k
                            if (slab_cons2(glo.x, glo.y, glo.z, ghi.x, ghi.y, ghi.z, Rg, &gl, &gsure) && (gsure || gate_passes(S, S.gate[slot], R.o, R.d))) {
                                ret.depth = dep; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = slot;
                                ret.index = S.leaf[slot];                         // face id: only read when the ray is stored
                                if (ANYHIT) { cur = -1; sp = 0; popped = false; q_count = 0; }    // occluded: done (stored when the lane is refilled)
                                else { best = dep; cull = dep + dep * PTB_CULL_GUARD; }
                            }
                        }
                    }
                }
            }
        }
    }
    if (item >= 0) io.store(item, ret, contrib);
    cp_async_wait<0>();
    flush_counters<COUNT>(C, 0ull, ANYHIT, ctr);
}
