// shade.cu -- the shading stage of the wavefront (and its parity taps): everything between two traversal stages of a path.
//
// Own translation unit because it is compiled TWICE into the library (ptina_b200/build.py):
//   strict (default, PTB_MODE_PARITY): -fmad=false -prec-div=true -prec-sqrt=true like the traversal, the LBVH build, ray generation
//          and the film (lbvh.cu, wavefront.cu), which owe the reference BITS.  Shading owes it 1e-5 per BSDF eval / pdf, and holds it.
//   fast   (opt-in, PTB_MODE_FAST, -DPTB_SHADE_FAST): FMA contraction, approximate division and square root.  Measured on config 2: the
//          shade stage 2.58 -> 1.95 ms (frame 7.85 -> 7.22 ms, almost all of it from division / square root: FMA alone gives 2.46 ms);
//          the converged-image gate (rel-RMSE <= 1e-3 at 1024 spp) still holds, the 1e-5 tap gate does NOT (sampled directions move by
//          up to 1e-4 where sqrt(1 - h^2) cancels), which is why it is a mode and not the default.
// What must not move stays explicit either way: the hit position that becomes the next ray's origin (shading_frame) is written with
// __fmul_rn / __fadd_rn.
#include "ptb_wavefront.cuh"

namespace {

constexpr int BLK = 128;

// ---- shade: the body of the while loop of path_trace (path.py:25-62) / BruteEngine.trace (brute.py:35-60) after the hit ---------
// per-scene constants (materials without textures, light frames), recomputed whenever the parameters are uploaded
__global__ void k_prepare_cache(const SceneParams* __restrict__ P, const float4* __restrict__ texels, SceneCache* SC) {
    int i = threadIdx.x;
    if (i <= PTB_MAX_MATERIALS) {
        int mtlid = i == PTB_MAX_MATERIALS ? -1 : i;
        bool plain = true;
        if (mtlid >= 0) for (int s = 0; s < PTB_NSLOTS; s++) plain = plain && P->mat_tex[mtlid][s] == -1;
        SC->plain[i] = plain;
        if (plain) SC->mat[i] = material_get(P, texels, mtlid, 0.0f, 0.0f);
    }
    if (i <= PTB_MAX_LIGHTS) SC->light[i] = light_cache(P->lights[i]);
}

// 1024 threads per SM at 64 registers; the block size sets how many warps run the same instruction stream (the kernel is
// ~130 KB of SASS: with many small blocks at different places of it the instruction caches thrash, ncu `stall_no_inst`)
#ifndef PTB_SHADE_BLK
#define PTB_SHADE_BLK 512
#endif
constexpr int SBLK = PTB_SHADE_BLK;
#ifndef PTB_SHADE_MINB
#define PTB_SHADE_MINB (1024 / PTB_SHADE_BLK)      /* resident blocks per SM the kernel is compiled for (register cap = 65536 / threads) */
#endif
#ifndef PTB_SHADE_PF2
#define PTB_SHADE_PF2 1          /* 1: prefetch the next iteration's path state into L1 (matball shade 4.65 -> 4.46 ms; config 2 unchanged), 0: off */
#endif
template <int ENGINE>
__global__ void __launch_bounds__(SBLK, PTB_SHADE_MINB) k_shade(const SceneParams* __restrict__ P, const SceneCache* __restrict__ SC, const float4* __restrict__ texels, const float* __restrict__ verts,
                                               const int* __restrict__ mtlids, const int* __restrict__ slot_of, const float* __restrict__ rngtab, int dim, int rng_stride,
                                               FrameMap fm, PathState st, RayQueue q_in, RayQueue q_out, RayQueue q_shadow, Ctrl* ctrl) {
    // two sets of append buffers, used alternately (block_append: no trailing barrier)
    __shared__ int s_warp[2][SBLK / 32]; __shared__ int s_base[2];
    __shared__ int s_warp2[2][2 * (SBLK / 32)]; __shared__ unsigned long long s_base2[2];
    int par = 0;
    const int count = ctrl->n_in;
    const int rounded = (count + SBLK - 1) / SBLK * SBLK;
    for (int i0 = blockIdx.x * SBLK; i0 < rounded; i0 += gridDim.x * SBLK) {
        int idx = i0 + threadIdx.x;
        bool alive = false, want_shadow = false;
        int p = -1, avoid_slot = -1;
        V3 next_o = v3s(0.0f), next_d = v3s(0.0f), sh_dir = v3s(0.0f), sh_contrib = v3s(0.0f);
        float sh_dis = 0.0f;
        if (idx < count) {
            {   // start the next iteration's records on their way
                const int nxt = idx + gridDim.x * SBLK;
                if (nxt < count) { asm volatile("prefetch.global.L1 [%0];" ::"l"(q_in.o + nxt)); asm volatile("prefetch.global.L1 [%0];" ::"l"(q_in.d + nxt)); }
            }
            float4 o4 = q_in.o[idx], d4 = q_in.d[idx];
            p = __float_as_int(o4.w);
            float4 h4 = st.hit[p], t4 = st.thr[p], r4 = st.result[p];
            V3 ro = mk3(o4.x, o4.y, o4.z), rd = mk3(d4.x, d4.y, d4.z);
            V3 thr = mk3(t4.x, t4.y, t4.z), result = mk3(r4.x, r4.y, r4.z);
            float last_pdf = r4.w;
            int depth = __float_as_int(t4.w) + 1;                       // path.py:26 depth += 1
            int hit_index = __float_as_int(h4.w);
            bool hit = hit_index >= 0;
            Rng rng;
            if (dim == 0) { rng.tab = rngtab + (size_t)p * rng_stride; rng.base = 0; rng.dim = 0; rng.fmod = fm.d_dim; }
            else {
                const int lp = p - fm.slot_base;                      // slot within the lane's share of the pool
                int s = (int)fastdiv((unsigned)lp, fm.d_pps), q = lp - s * fm.pps, x, y;
                slot_pixel(fm, q, &x, &y);
                frame_rng(fm, rngtab, dim, s, x, y, &rng);
            }
            // path.py:31-35 / brute.py:41-43
            LitHit lit = light_hit_cached(P, SC, ro, rd);
            if (lit.hit != 0 && (!hit || lit.dis < h4.x)) {
                if (ENGINE == PTB_ENGINE_PATH) {
                    float mis = power_heuristic(last_pdf, lit.pdf);
                    result = result + thr * (mis * lit.color);
                } else {
                    result = result + thr * lit.color;
                }
            }
            if (!hit) {
                result = result + thr * world_at(P, texels, rd);        // path.py:37-39
            } else {
                avoid_slot = __ldg(&slot_of[hit_index]);
                V3 hitpos, normal; float tu, tv; int mtlid;
                shading_frame(verts, mtlids, hit_index, h4.y, h4.z, ro, rd, h4.x, &hitpos, &normal, &tu, &tv, &mtlid);
                const int mslot = mtlid < 0 ? PTB_MAX_MATERIALS : mtlid;
                Disney mat;
                if (SC->plain[mslot]) mat = SC->mat[mslot];              // untextured: precomputed with the same arithmetic
                else mat = material_get(P, texels, mtlid, tu, tv);
                float sign = -dot(rd, normal);                            // path.py:44-46 (recomputed on the flipped normal)
                if (sign < 0.0f) normal = -normal;
                V3 wi = -rd;
                int c0;
                const RngRun run(rng, ENGINE == PTB_ENGINE_PATH ? 2 + 6 * (depth - 1) : 2 + 3 * (depth - 1));
                if (ENGINE == PTB_ENGINE_PATH) {
                    c0 = 0;
                    // path.py:48-56 next-event estimation
                    V3 ls = mk3(run.draw(c0), run.draw(c0 + 1), run.draw(c0 + 2));
                    LitSample li = light_sample_cached(P, SC, hitpos, ls);
                    if (any_gt(li.color, 0.0f)) {
                        V3 brdf_clr = disney_brdf(mat, normal, sign, wi, li.dir);
                        float brdf_pdf = vavg(brdf_clr);                  // sic: path.py:53
                        float mis = power_heuristic(li.pdf, brdf_pdf);
                        V3 direct = mis * li.color * brdf_clr * dot_or_zero(normal, li.dir);
                        V3 contrib = thr * direct;
                        // a contribution that is exactly zero (light sample below the horizon, black throughput) adds nothing whether or
                        // not the shadow ray is blocked, so the ray is not traced; NaNs compare unequal and still go through
                        if (!(contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f)) {
                            sh_dir = li.dir; sh_dis = li.dis; sh_contrib = contrib;
                            want_shadow = true;
                        }
                    }
                    c0 += 3;
                } else {
                    c0 = 0;
                }
                // path.py:58-62 / brute.py:56-60
                BSDFSample bs = disney_bounce(mat, normal, sign, wi, mk3(run.draw(c0), run.draw(c0 + 1), run.draw(c0 + 2)));
                thr = thr * bs.color;
                last_pdf = bs.pdf;                                         // path.py:61
                st.thr[p] = make_float4(thr.x, thr.y, thr.z, __int_as_float(depth));
                // loop condition path.py:25 / brute.py:35
                if (ENGINE == PTB_ENGINE_PATH) alive = depth < 5 && any_gt(thr, 0.0f) && any_ne0(bs.outdir);
                else alive = depth < 5 && any_gt(thr, PTB_EPS);
                next_o = hitpos;
                if (alive) next_d = normalized(bs.outdir);                 // path.py:28  r.d = r.d.normalized() at the top of the next iteration
            }
            st.result[p] = make_float4(result.x, result.y, result.z, last_pdf);
#if PTB_SHADE_PF2
            {   // the next iteration's path state: its slot comes out of the queue record prefetched above (an L1 hit by now), and
                // the three gathers it addresses would otherwise start only after that record has been read at the top of the loop
                const int nxt = idx + gridDim.x * SBLK;
                if (nxt < count) {
                    const int pn = __float_as_int(reinterpret_cast<const float*>(q_in.o + nxt)[3]);
#if PTB_SHADE_PF2 == 1
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(st.hit + pn)); asm volatile("prefetch.global.L1 [%0];" ::"l"(st.thr + pn)); asm volatile("prefetch.global.L1 [%0];" ::"l"(st.result + pn));
#else
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(st.hit + pn)); asm volatile("prefetch.global.L2 [%0];" ::"l"(st.thr + pn)); asm volatile("prefetch.global.L2 [%0];" ::"l"(st.result + pn));
#endif
                }
            }
#endif
        }
        int pos, ps = -1;
        if (ENGINE == PTB_ENGINE_PATH) block_append2<SBLK, false>(alive, want_shadow, &ctrl->n_out, s_warp2[par], &s_base2[par], &pos, &ps);
        else pos = block_append<SBLK, false>(alive, &ctrl->n_out, s_warp[par], &s_base[par]);
        par ^= 1;
        if (alive) {
            q_out.o[pos] = make_float4(next_o.x, next_o.y, next_o.z, __int_as_float(p));
            q_out.d[pos] = make_float4(next_d.x, next_d.y, next_d.z, __int_as_float(avoid_slot));   // avoid = hit.index (as its leaf slot)
        }
        if (ENGINE == PTB_ENGINE_PATH) {
            if (want_shadow) {
                q_shadow.o[ps] = make_float4(next_o.x, next_o.y, next_o.z, __int_as_float(p));
                q_shadow.d[ps] = make_float4(sh_dir.x, sh_dir.y, sh_dir.z, sh_dis);
                q_shadow.c[ps] = make_float4(sh_contrib.x, sh_contrib.y, sh_contrib.z, __int_as_float(avoid_slot));
            }
        }
    }
}

PTB_D Disney disney_from(const float* p) {
    Disney m;
    m.basecolor = mk3(p[0], p[1], p[2]); m.metallic = p[3]; m.roughness = p[4]; m.specular = p[5]; m.specularTint = p[6];
    m.subsurface = p[7]; m.sheen = p[8]; m.sheenTint = p[9]; m.clearcoat = p[10]; m.clearcoatGloss = p[11]; m.transmission = p[12]; m.ior = p[13];
    disney_init(m);
    return m;
}
// what: 0 eval_bsdf, 1 sample_bsdf, 2 material_get, 3 light_hit, 4 light_sample, 5 world_at, 6 / 7 eval / sample with every term as written (checkers)
__global__ void __launch_bounds__(BLK) k_shade_tap(const SceneParams* __restrict__ P, const float4* __restrict__ texels, int what, const float* __restrict__ in0,
                                                   const float* __restrict__ in1, const int* __restrict__ ini, int m, float* __restrict__ out) {
    int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= m) return;
    if (what == 0 || what == 1 || what == 6 || what == 7) {
        Disney d = disney_from(in0 + 14 * i);
        const float* g = in1 + 10 * i;
        V3 nrm = mk3(g[0], g[1], g[2]), wi = mk3(g[4], g[5], g[6]), x = mk3(g[7], g[8], g[9]);
        if (what == 0 || what == 6) {
            V3 r = what == 0 ? disney_brdf(d, nrm, g[3], wi, x) : disney_brdf_literal(d, nrm, g[3], wi, x);
            out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
        } else {
            BSDFSample s = what == 1 ? disney_bounce(d, nrm, g[3], wi, x) : disney_bounce_literal(d, nrm, g[3], wi, x);
            float* o = out + 7 * i;
            o[0] = s.outdir.x; o[1] = s.outdir.y; o[2] = s.outdir.z; o[3] = s.pdf; o[4] = s.color.x; o[5] = s.color.y; o[6] = s.color.z;
        }
    } else if (what == 2) {
        Disney d = material_get(P, texels, ini[i], in0[2 * i], in0[2 * i + 1]);
        float* o = out + 14 * i;
        o[0] = d.basecolor.x; o[1] = d.basecolor.y; o[2] = d.basecolor.z; o[3] = d.metallic; o[4] = d.roughness; o[5] = d.specular; o[6] = d.specularTint;
        o[7] = d.subsurface; o[8] = d.sheen; o[9] = d.sheenTint; o[10] = d.clearcoat; o[11] = d.clearcoatGloss; o[12] = d.transmission; o[13] = d.ior;
    } else if (what == 3) {
        const float* r = in0 + 6 * i;
        LitHit l = light_hit(P, mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]));
        float* o = out + 6 * i;
        o[0] = (float)l.hit; o[1] = l.dis; o[2] = l.pdf; o[3] = l.color.x; o[4] = l.color.y; o[5] = l.color.z;
    } else if (what == 4) {
        const float* r = in0 + 6 * i;
        LitSample l = light_sample(P, mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]));
        float* o = out + 8 * i;
        o[0] = l.dis; o[1] = l.dir.x; o[2] = l.dir.y; o[3] = l.dir.z; o[4] = l.pdf; o[5] = l.color.x; o[6] = l.color.y; o[7] = l.color.z;
    } else {
        V3 c = world_at(P, texels, mk3(in0[3 * i], in0[3 * i + 1], in0[3 * i + 2]));
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}

inline int nblk(long long n, int b = BLK) { return (int)((n + b - 1) / b); }

}  // namespace

#ifdef PTB_SHADE_FAST
#define SHADE_FN(name) name##_fast
#else
#define SHADE_FN(name) name##_strict
#endif

void SHADE_FN(ptb_shade_prepare_cache)(ptb_ctx* c) {
    k_prepare_cache<<<1, 128, 0, c->stream>>>(c->d_params, c->d_texels, c->d_cache);
    c->launches++;
}

void SHADE_FN(ptb_shade_launch)(ptb_ctx* c, const Lane& L, int engine, const float* rngtab, int dim, int rng_stride, const FrameMap& fm, int cur, cudaStream_t st) {
    const int grid = c->sm_count * PTB_SHADE_MINB;
    if (engine == PTB_ENGINE_PATH)
        k_shade<PTB_ENGINE_PATH><<<grid, SBLK, 0, st>>>(c->d_params, c->d_cache, c->d_texels, c->d_verts, c->d_mtlids, c->d_slot_of, rngtab, dim, rng_stride, fm, c->st,
                                                     L.xq[cur], L.xq[cur ^ 1], L.sq, L.ctrl);
    else
        k_shade<PTB_ENGINE_BRUTE><<<grid, SBLK, 0, st>>>(c->d_params, c->d_cache, c->d_texels, c->d_verts, c->d_mtlids, c->d_slot_of, rngtab, dim, rng_stride, fm, c->st,
                                                      L.xq[cur], L.xq[cur ^ 1], L.sq, L.ctrl);
    c->launches++;
}

int SHADE_FN(ptb_wf_shade_tap)(ptb_ctx* c, int what, const float* in0_dev, const float* in1_dev, const int32_t* ini_dev, int m, float* out_dev) {
    if (ptb_wf_upload_params(c)) return 1;
    k_shade_tap<<<nblk(m), BLK, 0, c->stream>>>(c->d_params, c->d_texels, what, in0_dev, in1_dev, ini_dev, m, out_dev);
    c->launches++;
    PTB_CUDA(cudaGetLastError());
    return 0;
}
