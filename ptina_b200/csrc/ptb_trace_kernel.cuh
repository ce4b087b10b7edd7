// ptb_trace_kernel.cuh -- the traversal kernels.
//
//  k_trace      : production kernel.  Persistent warps; every lane owns one ray and keeps its traversal state in
//                 registers (+ a local-memory stack).  A lane that finishes its ray takes the next one from a warp-local
//                 pool of queue indices (refilled 256 at a time with one atomic), so lanes do not idle while the
//                 slowest ray of a 32-ray batch finishes.  Every iteration the warp votes for ONE kind of step -- a node
//                 step (fetch a 64-B node, two exact slab tests, push/descend) or a leaf step (one triangle test from the
//                 lane's small queue of pending leaves) -- whichever more lanes can take, so that the lanes of a warp run
//                 the same code.  Deferring a triangle test never changes the result: acceptance is `depth < best`, ties
//                 go to the larger leaf slot.  Same predicates as trace_ordered in ptb_traverse.cuh; results are
//                 bit-identical to trace_reference.
//  k_trace_ref  : the reference's literal traversal (lbvh.py:313-347), one ray per thread; used when the tree is not
//                 a proper tree, when PTB_TRAVERSE_REFERENCE is requested, and by the tests as the on-device checker.
//
// Ray sources/sinks (IO): the wavefront's extend and shadow queues, and flat arrays for the parity taps.
#pragma once
#include "ptb_internal.h"

#define PTB_TRACE_BLK 128
#define PTB_TRACE_CHUNK 256
#ifndef PTB_NODE_STEPS
#define PTB_NODE_STEPS 3
#endif

// ---- extend queue: closest hit for path p (path.py:28-29) ----------------------------------------------------------------
struct ExtendIO {
    PathState st; const int* __restrict__ queue;
    static constexpr bool kAnyHit = false;
    PTB_D void load(int idx, int* item, V3* ro, V3* rd, int* avoid, float* tmax) const {
        int p = queue[idx];
        float4 o4 = st.ray_o[p], d4 = st.ray_d[p];
        *item = p; *ro = mk3(o4.x, o4.y, o4.z);
        *rd = normalized(mk3(d4.x, d4.y, d4.z));                  // path.py:28  r.d = r.d.normalized()
        st.ray_d[p] = make_float4(rd->x, rd->y, rd->z, d4.w);
        *avoid = __float_as_int(d4.w); *tmax = PTB_INF;
    }
    PTB_D void store(int p, const HitRec& h) const { st.hit[p] = make_float4(h.depth, h.u, h.v, __int_as_float(h.hit ? h.index : -1)); }
};
// ---- shadow queue: Ray(hitpos, li.dir) against avoid = hit triangle; unoccluded -> add the pending contribution (path.py:49-55)
struct ShadowIO {
    PathState st; const int* __restrict__ queue;
    static constexpr bool kAnyHit = true;
    PTB_D void load(int idx, int* item, V3* ro, V3* rd, int* avoid, float* tmax) const {
        int p = queue[idx];
        float4 o4 = st.ray_o[p], d4 = st.sh_d[p];
        *item = p; *ro = mk3(o4.x, o4.y, o4.z); *rd = mk3(d4.x, d4.y, d4.z);   // not re-normalised, as in the reference
        *avoid = __float_as_int(st.ray_d[p].w); *tmax = d4.w;
    }
    PTB_D void store(int p, const HitRec& h) const {
        if (h.hit) return;
        float4 r = st.result[p], c4 = st.sh_c[p];
        st.result[p] = make_float4(r.x + c4.x, r.y + c4.y, r.z + c4.z, r.w);
    }
};
// ---- parity taps: flat ray arrays ----------------------------------------------------------------------------------------------
template <bool ANYHIT>
struct TapIO {
    const float* __restrict__ rays; const int* __restrict__ avoid; const float* __restrict__ dis;
    int* hit; float* depth; int* index; float* uv;
    static constexpr bool kAnyHit = ANYHIT;
    PTB_D void load(int idx, int* item, V3* ro, V3* rd, int* av, float* tmax) const {
        *item = idx;
        *ro = mk3(rays[6 * idx], rays[6 * idx + 1], rays[6 * idx + 2]); *rd = mk3(rays[6 * idx + 3], rays[6 * idx + 4], rays[6 * idx + 5]);
        *av = avoid ? avoid[idx] : -1; *tmax = ANYHIT ? dis[idx] : PTB_INF;
    }
    PTB_D void store(int i, const HitRec& h) const {
        hit[i] = h.hit;
        if (!ANYHIT) { depth[i] = h.depth; index[i] = h.index; uv[2 * i] = h.u; uv[2 * i + 1] = h.v; }
    }
};

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

// reference policy: one ray per thread, static chunks of 32 per warp
template <class IO, bool COUNT>
__global__ void __launch_bounds__(PTB_TRACE_BLK) k_trace_ref(TraceScene S, IO io, int* cursor, const int* count_ptr, DevCounters* ctr) {
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
            int item, avoid; V3 ro, rd; float tmax;
            io.load(idx, &item, &ro, &rd, &avoid, &tmax);
            HitRec h = trace_reference<COUNT>(S, ro, rd, avoid, &C);
            if (IO::kAnyHit) h.hit = !(h.hit == 0 || h.depth > tmax);       // path.py:50  occ.hit == 0 or occ.depth > li.dis
            io.store(item, h);
            if (COUNT) nrays++;
        }
    }
    flush_counters<COUNT>(C, nrays, IO::kAnyHit, ctr);
}

// Pending-leaf queue of a lane: a ring of PTB_PQ (slot, entry distance) pairs in shared memory (column = thread, so the
// dynamic index costs no divergence and no bank conflict).
#define PTB_PQ 4
struct LeafQueue {
    int* slots; float* nears; int head, count;
    PTB_D void clear() { head = 0; count = 0; }
    PTB_D void push(int s, float nr) {
        const int k = ((head + count) & (PTB_PQ - 1)) * PTB_TRACE_BLK;
        slots[k] = s; nears[k] = nr;
        count++;
    }
    PTB_D void pop(int* s, float* nr) {
        const int k = head * PTB_TRACE_BLK;
        *s = slots[k]; *nr = nears[k];
        head = (head + 1) & (PTB_PQ - 1);
        count--;
    }
};

template <class IO, bool COUNT>
__global__ void __launch_bounds__(PTB_TRACE_BLK) k_trace(TraceScene S, IO io, int* cursor, const int* count_ptr, DevCounters* ctr) {
    constexpr bool ANYHIT = IO::kAnyHit;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int count = *count_ptr;
    const int n = S.n;
    TraceCounters C; C.nodes = C.boxes = C.tris = 0; C.max_stack = 0;
    unsigned long long nrays = 0;

    int pool_next = 0, pool_end = 0;      // warp-uniform pool of queue indices
    bool exhausted = false;
    // ---- per-lane ray state ----
    bool have = false;
    int item = -1, avoid_slot = -1;
    RayPre P; P.o = v3s(0.0f); P.d = v3s(0.0f); P.r = v3s(0.0f); P.par = false;
    float best = 0.0f, cull = 0.0f;
    HitRec ret; ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
    int stack_id[PTB_STACK]; float stack_near[PTB_STACK];
    int sp = 0, cur = -1;
    float cur_near = 0.0f;
    __shared__ int q_slots[PTB_PQ * PTB_TRACE_BLK];
    __shared__ float q_nears[PTB_PQ * PTB_TRACE_BLK];
    LeafQueue Q; Q.slots = q_slots + threadIdx.x; Q.nears = q_nears + threadIdx.x; Q.clear();

    while (true) {
        // ---- fetch: idle lanes take the next queue entries (batched: at least 4 idle lanes, or nothing else to do) -------------
        const unsigned idle = __ballot_sync(FULL, !have);
        if (idle != 0u && !exhausted && (__popc(idle) >= 4 || idle == FULL)) {
            if (pool_next >= pool_end) {
                int b = 0;
                if (lane == 0) b = atomicAdd(cursor, PTB_TRACE_CHUNK);
                b = __shfl_sync(FULL, b, 0);
                pool_next = b; pool_end = min(b + PTB_TRACE_CHUNK, count);
                if (b >= count) { exhausted = true; pool_end = pool_next; }
            }
            const int idx = pool_next + __popc(idle & lt_mask);
            if (!have && idx < pool_end) {
                V3 ro, rd; int avoid; float tmax;
                io.load(idx, &item, &ro, &rd, &avoid, &tmax);
                P = ray_pre(ro, rd);
                avoid_slot = avoid >= 0 ? S.slot_of[avoid] : -1;
                best = ANYHIT ? fminf(tmax, PTB_INF) : PTB_INF;
                cull = best + best * PTB_CULL_GUARD;
                ret.hit = 0; ret.depth = PTB_INF; ret.index = -1; ret.u = 0.0f; ret.v = 0.0f; ret.slot = -1;
                sp = 0; cur_near = 0.0f; Q.clear();
                float nr;
                if (COUNT) { C.boxes++; nrays++; }
                // the root's own box (the reference pops and tests it first)
                cur = slab_fast(S.root_lo[0], S.root_lo[1], S.root_lo[2], S.root_hi[0], S.root_hi[1], S.root_hi[2], P, &nr) ? 0 : -1;
                have = true;
            }
            pool_next = min(pool_next + __popc(idle), pool_end);
        }
        // ---- bookkeeping: culled current node -> pop; nothing left -> the ray is finished ---------------------------------------------
        if (have) {
            if (cur >= 0 && cur_near > cull) cur = -1;                   // the best hit improved since this node was chosen
            if (cur < 0) {
                while (sp > 0) {
                    --sp;
                    if (!(stack_near[sp] > cull)) { cur = stack_id[sp]; cur_near = stack_near[sp]; break; }
                }
                if (cur < 0 && Q.count == 0) {
                    if (!ANYHIT && ret.hit) ret.index = S.leaf[ret.slot];
                    io.store(item, ret);
                    have = false;
                }
            }
        }
        const bool node_ok = have && cur >= 0 && Q.count <= PTB_PQ - 2;
        const bool leaf_ok = have && Q.count > 0;
        const unsigned mn = __ballot_sync(FULL, node_ok), ml = __ballot_sync(FULL, leaf_ok);
        if ((mn | ml) == 0u) {
            if (exhausted && __ballot_sync(FULL, have) == 0u) break;
            continue;
        }
        if (mn != 0u && __popc(mn) > __popc(ml)) {
            // ---- node step: one 64-byte node, both children's exact slab tests (all participating lanes run the same code) -----------------
            if (node_ok) {
                const Node64 N = S.nodes[cur];
                if (COUNT) { C.nodes++; C.boxes += 2; }
                const int c0 = __float_as_int(N.a.w), c1 = __float_as_int(N.b.w);
                float n0, n1;
                const bool h0 = slab_fast(N.a.x, N.a.y, N.a.z, N.b.x, N.b.y, N.b.z, P, &n0);
                const bool h1 = slab_fast(N.c.x, N.c.y, N.c.z, N.d.x, N.d.y, N.d.z, P, &n1);
                // leaves are always tested (no box predicate in the reference) unless their bounds are entered beyond the cull distance
                const bool leaf0 = c0 < n, leaf1 = c1 < n;
                if (leaf0 && c0 != avoid_slot && !(h0 && n0 > cull)) Q.push(c0, h0 ? n0 : 0.0f);
                if (leaf1 && c1 != avoid_slot && !(h1 && n1 > cull)) Q.push(c1, h1 ? n1 : 0.0f);
                const bool d0 = !leaf0 && h0 && !(n0 > cull), d1 = !leaf1 && h1 && !(n1 > cull);
                if (d0 && d1) {
                    const bool first1 = !(n0 < n1);      // nearer first; equal entry distance: child1 first like the reference
                    if (sp < PTB_STACK) { stack_id[sp] = (first1 ? c0 : c1) - n; stack_near[sp] = first1 ? n0 : n1; sp++; }
                    if (COUNT) C.max_stack = max(C.max_stack, (unsigned)sp);
                    cur = (first1 ? c1 : c0) - n; cur_near = first1 ? n1 : n0;
                } else if (d0) { cur = c0 - n; cur_near = n0; }
                else if (d1) { cur = c1 - n; cur_near = n1; }
                else cur = -1;
            }
        } else {
            // ---- leaf step: the oldest pending triangle of every lane that has one -----------------------------------------------------------
            if (leaf_ok) {
                int slot; float lnear;
                Q.pop(&slot, &lnear);
                if (!(lnear > cull)) {
                    const Tri64 T = S.tris[slot];
                    if (COUNT) C.tris++;
                    float dep, s, t;
                    if (tri_fast(T, P.o, P.d, best, &dep, &s, &t)) {
                        if (ANYHIT) {
                            if (dep < PTB_INF) {                              // occluded: done with this ray
                                ret.hit = 1; ret.depth = dep; ret.u = s; ret.v = t; ret.slot = slot; ret.index = S.leaf[slot];
                                io.store(item, ret);
                                have = false;
                            }
                        } else if (dep < ret.depth || (ret.hit && slot > ret.slot)) {     // here dep <= best == ret.depth
                            ret.depth = dep; ret.u = s; ret.v = t; ret.hit = 1; ret.slot = slot;
                            best = dep; cull = dep + dep * PTB_CULL_GUARD;
                        }
                    }
                }
            }
        }
    }
    flush_counters<COUNT>(C, nrays, ANYHIT, ctr);
}
