// ptb_shade.cuh -- device-side shading: texture fetch, material table, Disney BSDF, lights, environment,
// camera.  Each function names the reference code whose arithmetic it reproduces (paths relative to
// the reference's ptina/ package).  Float results are held to <=1e-5 relative against the oracle.
#pragma once
#include "ptb_math.cuh"

#define PTB_MAX_MATERIALS 64
#define PTB_MAX_TEXTURES 64
#define PTB_MAX_LIGHTS 64
#define PTB_NSLOTS 12

struct LightRec {          // light/__init__.py:13-19
    int type; float size;
    V3 color, pos;
    M33 axes;
};

// Everything the shading kernels read that is not per-triangle; lives in one device buffer, read through
// the read-only path (uniform addresses -> broadcast).
struct SceneParams {
    float v2w[16];                                              // camera.py:10-12
    float mat_fac[PTB_MAX_MATERIALS][PTB_NSLOTS][4];            // mtllib.py:11-13 (x12, mtllib.py:42-56)
    int mat_tex[PTB_MAX_MATERIALS][PTB_NSLOTS];
    int img_nx[PTB_MAX_TEXTURES], img_ny[PTB_MAX_TEXTURES], img_base[PTB_MAX_TEXTURES];   // image.py:14-16
    LightRec lights[PTB_MAX_LIGHTS + 1];                        // +1: _sample's inclusive clamp reads [count]
    int nlights;
    float world_fac[4]; int world_tex;                          // light/world.py:10-16
    int nx, ny;                                                 // filmtable.py:11
};

struct Disney {            // materials/disney.py:14-50
    V3 basecolor; float metallic, roughness, specular, specularTint, subsurface, sheen, sheenTint, clearcoat,
        clearcoatGloss, transmission, ior;
    V3 speccolor, sheencolor; float alpha, clearcoatAlpha;
    int elim;              // PTB_ELIM_* : terms of brdf / bounce that are exactly zero for this material and may be left out (disney_init)
};
struct BSDFSample { V3 outdir; float pdf; V3 color; };   // materials/__init__.py:8-18

// ---- image.py:21-24, 137-148 + common.py:183-192 (wrap-mode bilinear, x-major texels) ------------
PTB_D V4 img_fetch(const SceneParams* P, const float4* __restrict__ texels, int id, int x, int y) {
    int nx = P->img_nx[id], ny = P->img_ny[id];
    if (nx <= 0 || ny <= 0) return mk4(0.0f, 0.0f, 0.0f, 0.0f);   // unloaded image id: `x % 0` in the reference (undefined); fenced
    x = pymod(x, nx); y = pymod(y, ny);
    float4 t = __ldg(&texels[(size_t)P->img_base[id] + (size_t)x * ny + y]);
    return mk4(t.x, t.y, t.z, t.w);
}
PTB_D V4 img_sample(const SceneParams* P, const float4* __restrict__ texels, int id, float u, float v) {
    float px = u * (float)(P->img_nx[id] - 1), py = v * (float)(P->img_ny[id] - 1);
    int ix = ifloor(px), iy = ifloor(py);
    float fx = px - (float)ix, fy = py - (float)iy;
    float gx = 1.0f - fx, gy = 1.0f - fy;
    return img_fetch(P, texels, id, ix + 1, iy + 1) * fx * fy + img_fetch(P, texels, id, ix + 1, iy) * fx * gy +
           img_fetch(P, texels, id, ix, iy) * gx * gy + img_fetch(P, texels, id, ix, iy + 1) * gx * fy;
}

// ---- mtllib.py:30-38 ParameterPair.get ----------------------------------------------------------------
PTB_D V4 param_get(const SceneParams* P, const float4* __restrict__ texels, int slot, int mtlid, float u, float v, float dflt) {
    V4 fac = mk4(dflt, dflt, dflt, dflt);
    if (mtlid != -1) {
        const float* f = P->mat_fac[mtlid][slot];
        fac = mk4(f[0], f[1], f[2], f[3]);
        int texid = P->mat_tex[mtlid][slot];
        if (texid != -1) fac = fac * img_sample(P, texels, texid, u, v);
    }
    return fac;
}

// ---- terms that vanish exactly ----------------------------------------------------------------------------------
// Disney.brdf / Disney.bounce (disney.py:52-233) always evaluate every lobe and multiply by the material's weights.  When a weight is
// EXACTLY zero -- transmission, clearcoat, subsurface of every material of configs 1, 2 and 4 -- the product is +-0 provided the other
// factor is finite, and adding +-0 leaves the sum as it is (bit for bit, but for the sign of a zero result, which nothing downstream can
// see: the value is only added, multiplied or compared).  So the factor need not be computed: the transmission Fresnel term (two
// square roots, three divisions), the clearcoat lobe (a logarithm, two square roots, three divisions) and the subsurface term (a
// division) are 250 of the ~1340 instructions of a path vertex.  Finiteness is not assumed but guarded: a material qualifies only if
// every parameter is finite and of ordinary size (`sane`: |x| <= 1e4, ior >= 1e-3, so that both refraction indices are positive), and
// each call site adds the cheap run-time condition under which the skipped factor is provably finite (stated there); when a guard fails
// the literal expression runs, so non-finite values (NaN from 0 * inf) come out exactly where the reference produces them.
// disney_brdf_t<false> / disney_bounce_t<false> are the literal evaluations; tests/test_gpu_parity.py compares the two bit for bit.
#define PTB_ELIM_TRANS 1          /* transmission == 0 */
#define PTB_ELIM_COAT 2           /* clearcoat == 0 and clearcoatAlpha in [1e-3, 0.5] */
#define PTB_ELIM_SS 4             /* subsurface == 0 */
#ifndef PTB_DEAD_TERMS
#define PTB_DEAD_TERMS 1          /* 0: always the literal expressions (A/B builds) */
#endif
PTB_D int disney_elim(const Disney& m) {
    const float B = 1e4f;
    bool sane = fabsf(m.basecolor.x) <= B && fabsf(m.basecolor.y) <= B && fabsf(m.basecolor.z) <= B && fabsf(m.metallic) <= B && fabsf(m.roughness) <= B &&
                fabsf(m.specular) <= B && fabsf(m.specularTint) <= B && fabsf(m.subsurface) <= B && fabsf(m.sheen) <= B && fabsf(m.sheenTint) <= B &&
                fabsf(m.clearcoat) <= B && fabsf(m.clearcoatGloss) <= B && fabsf(m.transmission) <= B && m.ior >= 1e-3f && m.ior <= B;      // NaNs fail
    sane = sane && fabsf(m.sheencolor.x) <= B && fabsf(m.sheencolor.y) <= B && fabsf(m.sheencolor.z) <= B;
    int e = 0;
    if (sane && m.transmission == 0.0f) e |= PTB_ELIM_TRANS;
    if (sane && m.clearcoat == 0.0f && m.clearcoatAlpha >= 1e-3f && m.clearcoatAlpha <= 0.5f) e |= PTB_ELIM_COAT;
    if (sane && m.subsurface == 0.0f) e |= PTB_ELIM_SS;
    return e;
}
// ---- materials/disney.py:14-50 --------------------------------------------------------------------------
PTB_D void disney_init(Disney& m) {
    V3 tint = v3s(1.0f);
    float lum = dot(m.basecolor, mk3(0.3f, 0.6f, 0.1f));
    if (lum > PTB_EPS) tint = m.basecolor / lum;
    m.speccolor = lerp3(m.metallic, (m.specular * 0.08f) * lerp3(m.specularTint, v3s(1.0f), tint), m.basecolor);
    m.sheencolor = lerp3(m.sheenTint, v3s(1.0f), tint);
    m.alpha = fmaxf(0.001f, m.roughness * m.roughness);
    m.clearcoatAlpha = lerpf(m.clearcoatGloss, 0.1f, 0.001f);
    m.elim = disney_elim(m);
}
// ---- mtllib.py:79-95 MaterialPool.get ---------------------------------------------------------------------
PTB_D Disney material_get(const SceneParams* P, const float4* __restrict__ texels, int mtlid, float u, float v) {
    Disney m;
    V4 b = param_get(P, texels, 0, mtlid, u, v, 0.8f);
    m.basecolor = mk3(b.x, b.y, b.z);
    m.metallic = param_get(P, texels, 1, mtlid, u, v, 0.0f).x;
    m.roughness = param_get(P, texels, 2, mtlid, u, v, 0.4f).x;
    m.specular = param_get(P, texels, 3, mtlid, u, v, 0.5f).x;
    m.specularTint = param_get(P, texels, 4, mtlid, u, v, 0.4f).x;
    m.subsurface = param_get(P, texels, 5, mtlid, u, v, 0.0f).x;
    m.sheen = param_get(P, texels, 6, mtlid, u, v, 0.0f).x;
    m.sheenTint = param_get(P, texels, 7, mtlid, u, v, 0.4f).x;
    m.clearcoat = param_get(P, texels, 8, mtlid, u, v, 0.0f).x;
    m.clearcoatGloss = param_get(P, texels, 9, mtlid, u, v, 0.5f).x;
    m.transmission = param_get(P, texels, 10, mtlid, u, v, 0.0f).x;
    m.ior = param_get(P, texels, 11, mtlid, u, v, 1.45f).x;
    disney_init(m);
    return m;
}

// ---- materials/microfacet.py:8-77 ---------------------------------------------------------------------------
PTB_D float schlickFresnel(float cost) {
    float c = clampf(1.0f - cost, 0.0f, 1.0f);
    float c2 = c * c;
    return c2 * c2 * c;
}
PTB_D float dielectricFresnel(float etai, float etao, float cosi) {
    float sini = sqrtf(fmaxf(0.0f, 1.0f - cosi * cosi));
    float sint = etao / etai * sini;
    float ret = 1.0f;
    if (sint < 1.0f) {
        float cost = sqrtf(fmaxf(0.0f, 1.0f - sint * sint));
        float a1 = etai * cosi, a2 = etao * cost;
        float b1 = etao * cosi, b2 = etai * cost;
        float para = (a1 - a2) / (a1 + a2);
        float perp = (b1 - b2) / (b1 + b2);
        ret = 0.5f * (para * para + perp * perp);
    }
    return ret;
}
PTB_D float GTR1(float cosh, float alpha) {
    float a2 = alpha * alpha;
    float t = 1.0f + (a2 - 1.0f) * (cosh * cosh);
    return (a2 - 1.0f) / (PTB_PI * logf(a2) * t);
}
PTB_D float GTR2(float cosh, float alpha) {
    float a2 = alpha * alpha;
    float t = 1.0f + (a2 - 1.0f) * (cosh * cosh);
    return a2 / (PTB_PI * (t * t));
}
PTB_D float smithGGX(float cosi, float alpha) {
    float a = alpha * alpha, b = cosi * cosi;
    return 1.0f / (cosi + sqrtf(a + b - a * b));
}
PTB_D V3 sample_GTR1(float u, float v, float alpha) {   // microfacet.py:68-71, division outside the sqrt as written there
    u = sqrtf(powf(alpha, 2.0f - 2.0f * u) - 1.0f) / (alpha * alpha - 1.0f);
    return spherical(u, v);
}
PTB_D V3 sample_GTR2(float u, float v, float alpha) {   // microfacet.py:74-77
    u = sqrtf((1.0f - u) / (1.0f - u * (1.0f - alpha * alpha)));
    return spherical(u, v);
}

// ---- materials/disney.py:52-106 Disney.brdf (value excludes the cosine), every term as written --------------------------------------
PTB_D V3 disney_brdf_literal(const Disney& m, V3 normal, float sign, V3 indir, V3 outdir) {
    float etai = 1.0f, etao = m.ior;
    if (sign < 0.0f) { etai = m.ior; etao = 1.0f; }
    V3 halfdir = normalized(indir + outdir);
    float cosi = dot(indir, normal), coso = dot(outdir, normal);
    float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
    V3 result = v3s(0.0f);
    if (coso < 0.0f) {
        if (cosi >= 0.0f) {
            float Ds = GTR2(cosh, m.alpha);
            float fdf = dielectricFresnel(etao, etai, cosoh);
            V3 transmit = (1.0f / PTB_PI) * m.basecolor * (1.0f - fdf) * Ds;
            result = transmit * (1.0f - m.metallic) * m.transmission;
        }
    } else {
        float Fi = schlickFresnel(cosi), Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2.0f * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float Fss90 = (cosoh * cosoh) * m.roughness;
        float Fss = lerpf(Fi, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
        float ss = 1.25f * (Fss * (1.0f / (cosi + coso) - 0.5f) + 0.5f);
        float Foh = schlickFresnel(cosoh);
        V3 Fsheen = (Foh * m.sheen) * m.sheencolor;
        float fdf = dielectricFresnel(etao, etai, cosoh);
        float Ds = GTR2(cosh, m.alpha);
        V3 Fs = lerp3(Foh, m.speccolor, v3s(1.0f));
        float Gs = smithGGX(cosi, m.alpha) * smithGGX(coso, m.alpha);
        float Dr = GTR1(cosh, m.clearcoatAlpha);
        float Gr = smithGGX(cosi, 0.25f) * smithGGX(coso, 0.25f);
        float Fr = lerpf(Foh, 0.04f, 1.0f);
        V3 diffuse = ((1.0f / PTB_PI) * lerpf(m.subsurface, Fd, ss)) * m.basecolor + Fsheen;
        V3 specular = (Gs * Fs) * Ds + v3s(0.25f * m.clearcoat * Gr * Fr * Dr);
        V3 transmit = ((1.0f / PTB_PI) * fdf * Ds) * m.basecolor;
        result = diffuse * (1.0f - m.metallic) * (1.0f - m.transmission);
        result = result + transmit * (1.0f - m.metallic) * m.transmission;
        result = result + specular * (1.0f - m.transmission);
    }
    return result;
}

// materials/__init__.py:21-48: lobe pick that re-uses one uniform by rescaling it
struct Choice {
    float pdf, w;
    PTB_D int pick(float r) {
        if (w < r) { w /= r; pdf *= r; return 1; }
        w = (w - r) / (1.0f - r); pdf *= 1.0f - r; return 0;
    }
    // the same; a rate of exactly zero (no clearcoat, no transmission) is decided without the division: for w >= 0 or NaN the test
    // w < 0 fails, (w - 0) / (1 - 0) = w and pdf * 1 = pdf
    PTB_D int pick_z(float r) {
        if (r == 0.0f && !(w < 0.0f)) return 0;
        return pick(r);
    }
};

// ---- materials/disney.py:114-233 Disney.bounce, as written ------------------------------------------------------------------
PTB_D BSDFSample disney_bounce_literal(const Disney& m, V3 normal, float sign, V3 indir, V3 samp) {
    BSDFSample res; res.outdir = v3s(0.0f); res.pdf = 0.0f; res.color = v3s(0.0f);
    float etai = 1.0f, etao = m.ior;
    if (sign < 0.0f) { etai = m.ior; etao = 1.0f; }
    float eta = etai / etao;
    float Fi = schlickFresnel(dot(indir, normal));
    V3 Fs = lerp3(Fi, m.speccolor, v3s(1.0f));
    Choice choice; choice.pdf = 1.0f; choice.w = samp.z;
    float specrate = lerpf(m.transmission, lerpf(m.metallic, vavg(Fs), 1.0f), 1.0f);
    float coatrate = 0.04f * m.clearcoat;
    specrate = lerpf(specrate, 0.1f, 1.0f);
    if (coatrate != 0.0f) coatrate = lerpf(coatrate, 0.1f, 1.0f);

    if (choice.pick(coatrate)) {
        float alpha = m.clearcoatAlpha;
        V3 halfdir = matvec(tanspace(normal), sample_GTR1(samp.x, samp.y, alpha));
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0.0f) {
            float Dr = GTR1(cosh, alpha);
            float Fr = lerpf(schlickFresnel(cosoh), 0.04f, 1.0f);
            res.outdir = outdir;
            float partial = m.clearcoat * Fr * coso / cosoh;
            res.pdf = Dr * partial;
            res.color = v3s(partial / choice.pdf);
        }
    } else if (choice.pick(specrate)) {
        float alpha = m.alpha;
        V3 halfdir = matvec(tanspace(normal), sample_GTR2(samp.x, samp.y, alpha));
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot_or_zero(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0.0f && coso > 0.0f && cosh > 0.0f) {
            float Ds = GTR2(cosh, alpha);
            if (choice.pick(m.transmission)) {
                float fdf = dielectricFresnel(etao, etai, cosoh);
                float reflrate = lerpf(fdf, 0.2f, 1.0f);
                if (choice.pick(reflrate)) {
                    res.outdir = outdir;
                    res.pdf = Ds * fdf;
                    res.color = m.basecolor * fdf * m.transmission / choice.pdf;
                } else {
                    V3 T;
                    if (refract(-indir, halfdir, eta, &T)) {
                        res.outdir = T;
                        res.pdf = Ds * (1.0f - fdf);
                        res.color = m.basecolor * (1.0f - fdf) * m.transmission / choice.pdf;
                    }
                }
            } else {
                V3 Fs2 = lerp3(schlickFresnel(cosoh), m.speccolor, v3s(1.0f));
                res.outdir = outdir;
                float partial = 0.5f / (cosoh * smithGGX(coso, alpha));
                res.pdf = Ds * vavg(Fs2) * partial;
                res.color = Fs2 * partial * (1.0f - m.transmission) / choice.pdf;
            }
        }
    } else {
        V3 outdir = matvec(tanspace(normal), spherical(sqrtf(samp.x), samp.y));
        V3 halfdir = normalized(indir + outdir);
        float cosi = dot(indir, normal), coso = dot(outdir, normal);
        float cosoh = dot_or_zero(halfdir, outdir);
        float Fi2 = schlickFresnel(cosi), Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2.0f * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi2, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float Fss90 = (cosoh * cosoh) * m.roughness;
        float Fss = lerpf(Fi2, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
        float ss = 1.25f * (Fss * (1.0f / (cosi + coso) - 0.5f) + 0.5f);
        V3 Fsheen = (schlickFresnel(cosoh) * m.sheen) * m.sheencolor;
        V3 diffuse = ((1.0f / PTB_PI) * lerpf(m.subsurface, Fd, ss)) * m.basecolor + Fsheen;
        res.outdir = outdir;
        res.pdf = 1.0f / PTB_PI;
        res.color = diffuse * PTB_PI * (1.0f - m.metallic) * (1.0f - m.transmission) / choice.pdf;
    }
    return res;
}


// ---- the production forms: the same values (see "terms that vanish exactly" above) -----------------------------------------------
// magnitudes under which the skipped factors below are provably finite; directions are unit vectors in the renderer (the taps may pass anything)
PTB_D bool elim_dirs_ok(float cosi, float coso, float cosoh) { return fabsf(cosi) <= 4.0f && fabsf(coso) <= 4.0f && cosoh <= 4.0f; }
PTB_D V3 disney_brdf(const Disney& m, V3 normal, float sign, V3 indir, V3 outdir) {
#if !PTB_DEAD_TERMS
    return disney_brdf_literal(m, normal, sign, indir, outdir);
#else
    const int el = m.elim;
    float etai = 1.0f, etao = m.ior;
    if (sign < 0.0f) { etai = m.ior; etao = 1.0f; }
    V3 halfdir = normalized(indir + outdir);
    float cosi = dot(indir, normal), coso = dot(outdir, normal);
    float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
    const bool dirs_ok = elim_dirs_ok(cosi, coso, cosoh);
    V3 result = v3s(0.0f);
    if (coso < 0.0f) {
        if (cosi >= 0.0f) {
            // transmission == 0: the value is (finite) * 0.  Finite: cosh <= 0.99999 keeps GTR2's t = 1 + (a2-1) cosh^2 >= 2e-5 (a2 <= 1e16),
            // dielectricFresnel is in [0, 1] for positive finite indices, basecolor and metallic are bounded by `sane`.
            if (!((el & PTB_ELIM_TRANS) && dirs_ok && cosh <= 0.99999f)) {
                float Ds = GTR2(cosh, m.alpha);
                float fdf = dielectricFresnel(etao, etai, cosoh);
                V3 transmit = (1.0f / PTB_PI) * m.basecolor * (1.0f - fdf) * Ds;
                result = transmit * (1.0f - m.metallic) * m.transmission;
            }
        }
    } else {
        float Fi = schlickFresnel(cosi), Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2.0f * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float Foh = schlickFresnel(cosoh);
        V3 Fsheen = (Foh * m.sheen) * m.sheencolor;
        float Ds = GTR2(cosh, m.alpha);
        V3 Fs = lerp3(Foh, m.speccolor, v3s(1.0f));
        float Gs = smithGGX(cosi, m.alpha) * smithGGX(coso, m.alpha);
        // subsurface == 0: lerp(0, Fd, ss) = Fd * 1 + ss * 0 = Fd for finite ss; |cosi + coso| > 1e-20 bounds its 1 / (cosi + coso)
        float dl;
        if ((el & PTB_ELIM_SS) && dirs_ok && fabsf(cosi + coso) > 1e-20f) dl = Fd;
        else {
            float Fss90 = (cosoh * cosoh) * m.roughness;
            float Fss = lerpf(Fi, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
            float ss = 1.25f * (Fss * (1.0f / (cosi + coso) - 0.5f) + 0.5f);
            dl = lerpf(m.subsurface, Fd, ss);
        }
        V3 diffuse = ((1.0f / PTB_PI) * dl) * m.basecolor + Fsheen;
        // clearcoat == 0: 0.25 * 0 * Gr * Fr * Dr.  Gr <= 16 for cosines in [0, 4] (smithGGX(c, 0.25) <= 4), Fr in [0.04, 1], and
        // GTR1's t >= 2e-5 for cosh <= 0.99999 with a2 = clearcoatAlpha^2 in [1e-6, 0.25] (log a2 in [-13.9, -1.38])
        V3 specular = (Gs * Fs) * Ds;
        if (!((el & PTB_ELIM_COAT) && dirs_ok && cosi >= 0.0f && cosh <= 0.99999f)) {
            float Dr = GTR1(cosh, m.clearcoatAlpha);
            float Gr = smithGGX(cosi, 0.25f) * smithGGX(coso, 0.25f);
            float Fr = lerpf(Foh, 0.04f, 1.0f);
            specular = specular + v3s(0.25f * m.clearcoat * Gr * Fr * Dr);
        }
        result = diffuse * (1.0f - m.metallic) * (1.0f - m.transmission);
        // transmission == 0: transmit * (1 - metallic) * 0 with transmit = (1/pi) fdf Ds basecolor finite for |Ds| <= 1e30
        if (!((el & PTB_ELIM_TRANS) && dirs_ok && fabsf(Ds) <= 1e30f)) {
            float fdf = dielectricFresnel(etao, etai, cosoh);
            V3 transmit = ((1.0f / PTB_PI) * fdf * Ds) * m.basecolor;
            result = result + transmit * (1.0f - m.metallic) * m.transmission;
        }
        result = result + specular * (1.0f - m.transmission);
    }
    return result;
#endif
}

// Disney.bounce.  Besides the vanishing terms: (i) the three lobes all end in `tanspace(normal) @ spherical(h, samp.y)` -- a square
// root, a division, a sincos and two cross products -- with only h depending on the lobe; a warp whose lanes picked different lobes
// used to run that tail once per lobe, now h is chosen inside the branch and the tail runs once for all lanes (same operations on
// the same operands: same bits); (ii) eta = etai / etao is only read by refract(), so the division moves into that branch.
PTB_D BSDFSample disney_bounce(const Disney& m, V3 normal, float sign, V3 indir, V3 samp) {
#if !PTB_DEAD_TERMS
    return disney_bounce_literal(m, normal, sign, indir, samp);
#else
    const int el = m.elim;
    BSDFSample res; res.outdir = v3s(0.0f); res.pdf = 0.0f; res.color = v3s(0.0f);
    float etai = 1.0f, etao = m.ior;
    if (sign < 0.0f) { etai = m.ior; etao = 1.0f; }
    float Fi = schlickFresnel(dot(indir, normal));
    V3 Fs = lerp3(Fi, m.speccolor, v3s(1.0f));
    Choice choice; choice.pdf = 1.0f; choice.w = samp.z;
    float specrate = lerpf(m.transmission, lerpf(m.metallic, vavg(Fs), 1.0f), 1.0f);
    float coatrate = 0.04f * m.clearcoat;
    specrate = lerpf(specrate, 0.1f, 1.0f);
    if (coatrate != 0.0f) coatrate = lerpf(coatrate, 0.1f, 1.0f);

    int lobe; float h;
    if (choice.pick_z(coatrate)) {
        lobe = 0;
        const float alpha = m.clearcoatAlpha;
        h = sqrtf(powf(alpha, 2.0f - 2.0f * samp.x) - 1.0f) / (alpha * alpha - 1.0f);       // sample_GTR1
    } else if (choice.pick_z(specrate)) {
        lobe = 1;
        h = sqrtf((1.0f - samp.x) / (1.0f - samp.x * (1.0f - m.alpha * m.alpha)));           // sample_GTR2
    } else {
        lobe = 2;
        h = sqrtf(samp.x);
    }
    const V3 dirv = matvec(tanspace(normal), spherical(h, samp.y));
    if (lobe == 0) {
        float alpha = m.clearcoatAlpha;
        V3 halfdir = dirv;
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0.0f) {
            float Dr = GTR1(cosh, alpha);
            float Fr = lerpf(schlickFresnel(cosoh), 0.04f, 1.0f);
            res.outdir = outdir;
            float partial = m.clearcoat * Fr * coso / cosoh;
            res.pdf = Dr * partial;
            res.color = v3s(partial / choice.pdf);
        }
    } else if (lobe == 1) {
        float alpha = m.alpha;
        V3 halfdir = dirv;
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot_or_zero(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal), cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0.0f && coso > 0.0f && cosh > 0.0f) {
            float Ds = GTR2(cosh, alpha);
            if (choice.pick_z(m.transmission)) {
                float fdf = dielectricFresnel(etao, etai, cosoh);
                float reflrate = lerpf(fdf, 0.2f, 1.0f);
                if (choice.pick(reflrate)) {
                    res.outdir = outdir;
                    res.pdf = Ds * fdf;
                    res.color = m.basecolor * fdf * m.transmission / choice.pdf;
                } else {
                    V3 T;
                    if (refract(-indir, halfdir, etai / etao, &T)) {
                        res.outdir = T;
                        res.pdf = Ds * (1.0f - fdf);
                        res.color = m.basecolor * (1.0f - fdf) * m.transmission / choice.pdf;
                    }
                }
            } else {
                V3 Fs2 = lerp3(schlickFresnel(cosoh), m.speccolor, v3s(1.0f));
                res.outdir = outdir;
                float partial = 0.5f / (cosoh * smithGGX(coso, alpha));
                res.pdf = Ds * vavg(Fs2) * partial;
                res.color = Fs2 * partial * (1.0f - m.transmission) / choice.pdf;
            }
        }
    } else {
        V3 outdir = dirv;
        V3 halfdir = normalized(indir + outdir);
        float cosi = dot(indir, normal), coso = dot(outdir, normal);
        float cosoh = dot_or_zero(halfdir, outdir);
        float Fi2 = schlickFresnel(cosi), Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2.0f * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi2, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float dl;
        if ((el & PTB_ELIM_SS) && elim_dirs_ok(cosi, coso, cosoh) && fabsf(cosi + coso) > 1e-20f) dl = Fd;       // as in disney_brdf
        else {
            float Fss90 = (cosoh * cosoh) * m.roughness;
            float Fss = lerpf(Fi2, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
            float ss = 1.25f * (Fss * (1.0f / (cosi + coso) - 0.5f) + 0.5f);
            dl = lerpf(m.subsurface, Fd, ss);
        }
        V3 Fsheen = (schlickFresnel(cosoh) * m.sheen) * m.sheencolor;
        V3 diffuse = ((1.0f / PTB_PI) * dl) * m.basecolor + Fsheen;
        res.outdir = outdir;
        res.pdf = 1.0f / PTB_PI;
        res.color = diffuse * PTB_PI * (1.0f - m.metallic) * (1.0f - m.transmission) / choice.pdf;
    }
    return res;
#endif
}

// engine/path.py:10-14
PTB_D float power_heuristic(float a, float b) {
    a = clampf(a, PTB_EPS, PTB_INF); a = a * a;
    b = clampf(b, PTB_EPS, PTB_INF); b = b * b;
    return a / (a + b);
}

// ---- geometries.py:158-179 Sphere.intersect, geometries.py:57-73 Area.intersect ------------------------------------
PTB_D float sphere_intersect(V3 pos, float rad2, V3 ro, V3 rd) {
    V3 op = pos - ro;
    float b = dot(op, rd);
    float det = b * b + rad2 - norm_sqr(op);
    float ret = 0.0f;
    if (!(det < 0.0f)) {
        det = sqrtf(det);
        float t = b - det;
        if (t > PTB_EPS) ret = t;
        else { t = b + det; if (t > PTB_EPS) ret = t; }
    }
    return ret;
}
PTB_D int area_intersect(V3 pos, V3 dirx, V3 diry, V3 ro, V3 rd, float* tout) {
    float t = PTB_INF; int hit = 0;
    V3 nrm = normalized(cross(dirx, diry));
    float NoD = dot(nrm, rd);
    if (NoD > PTB_EPS) {
        t = dot(nrm, pos - ro) / NoD;
        V3 hd = ro + t * rd - pos;
        float u = dot(hd, dirx) / norm_sqr(dirx);
        float v = dot(hd, diry) / norm_sqr(diry);
        if (-1.0f < u && u < 1.0f && -1.0f < v && v < 1.0f) hit = 1;
    }
    *tout = t;
    return hit;
}
struct LitHit { int hit; float dis, pdf; V3 color; };
// ---- light/__init__.py:51-81 LightPool.hit: FIRST light in index order with 0 < t < dis wins (break) ----------------
PTB_D LitHit light_hit(const SceneParams* P, V3 ro, V3 rd) {
    LitHit ret; ret.hit = 0; ret.dis = PTB_INF; ret.pdf = 0.0f; ret.color = v3s(0.0f);
    int n = P->nlights;
    for (int i = 0; i < n; i++) {
        const LightRec& L = P->lights[i];
        float t = 0.0f, area = 0.0f;
        if (L.type == 1) {
            t = sphere_intersect(L.pos, L.size * L.size, ro, rd);
            area = PTB_PI * (L.size * L.size);
        } else if (L.type == 2) {
            V3 dirx = matvec(L.axes, mk3(L.size, 0.0f, 0.0f));
            V3 diry = matvec(L.axes, mk3(0.0f, L.size, 0.0f));
            float tt;
            if (area_intersect(L.pos, dirx, diry, ro, rd, &tt)) { t = tt; area = 4.0f * (L.size * L.size); }
        }
        if (0.0f < t && t < ret.dis) {
            ret.dis = t; ret.pdf = (t * t) / area; ret.color = L.color; ret.hit = 1;
            break;
        }
    }
    return ret;
}
struct LitSample { float dis; V3 dir; float pdf; V3 color; };
// ---- light/__init__.py:83-121 LightPool.sample / _sample -----------------------------------------------------------------
PTB_D LitSample light_sample(const SceneParams* P, V3 hitpos, V3 samp) {
    LitSample ret; ret.dis = PTB_INF; ret.dir = v3s(0.0f); ret.pdf = 0.0f; ret.color = v3s(0.0f);
    int n = P->nlights;
    if (n != 0) {
        int i = clampi(ifloor(samp.z * (float)n), 0, n);   // inclusive upper clamp as in the reference
        const LightRec& L = P->lights[i];
        V3 color = L.color, litpos = v3s(PTB_INF), nrm = v3s(0.0f);
        float area = 0.0f;
        if (L.type == 1) {
            litpos = L.pos + L.size * spherical(samp.x, samp.y);
            area = PTB_PI * (L.size * L.size);
        } else if (L.type == 2) {
            V3 disp = matvec(L.axes, mk3(samp.x * 2.0f - 1.0f, samp.y * 2.0f - 1.0f, 0.0f));
            nrm = matvec(L.axes, mk3(0.0f, 0.0f, 1.0f));
            litpos = L.pos + L.size * disp;
            area = 4.0f * (L.size * L.size);
        }
        V3 toli = litpos - hitpos;
        float dis = norm(toli);
        V3 dir = toli / dis;
        float pdf = (dis * dis) / area;
        color = color / pdf;
        if (any_ne0(nrm)) color = color * dot_or_zero(nrm, dir);
        ret.dis = dis; ret.dir = dir; ret.pdf = pdf; ret.color = color;
    }
    return ret;
}
// ---- per-scene constants hoisted out of the per-vertex code (k_prepare_cache): the f32 expressions below are the ones
// material_get / disney_init / light_hit / light_sample evaluate per call, computed once with the same operations -> same bits.
struct LightCache { V3 dirx, diry, nrm_hit, nrm_smp; float nsqx, nsqy, area, rad2; };
struct SceneCache {
    Disney mat[PTB_MAX_MATERIALS + 1];          // [mtlid], last = the default material (mtlid -1); valid where plain != 0
    int plain[PTB_MAX_MATERIALS + 1];           // no textured slot: the material does not depend on (u, v)
    LightCache light[PTB_MAX_LIGHTS + 1];
};
PTB_D LightCache light_cache(const LightRec& L) {
    LightCache c;
    c.dirx = matvec(L.axes, mk3(L.size, 0.0f, 0.0f));
    c.diry = matvec(L.axes, mk3(0.0f, L.size, 0.0f));
    c.nrm_hit = normalized(cross(c.dirx, c.diry));
    c.nrm_smp = matvec(L.axes, mk3(0.0f, 0.0f, 1.0f));
    c.nsqx = norm_sqr(c.dirx); c.nsqy = norm_sqr(c.diry);
    c.rad2 = L.size * L.size;
    c.area = L.type == 1 ? PTB_PI * (L.size * L.size) : 4.0f * (L.size * L.size);
    return c;
}
// light_hit with the hoisted constants (same arithmetic as light_hit / area_intersect above)
PTB_D LitHit light_hit_cached(const SceneParams* P, const SceneCache* SC, V3 ro, V3 rd) {
    LitHit ret; ret.hit = 0; ret.dis = PTB_INF; ret.pdf = 0.0f; ret.color = v3s(0.0f);
    int n = P->nlights;
    for (int i = 0; i < n; i++) {
        const LightRec& L = P->lights[i];
        const LightCache& C = SC->light[i];
        float t = 0.0f, area = 0.0f;
        if (L.type == 1) {
            t = sphere_intersect(L.pos, C.rad2, ro, rd);
            area = C.area;
        } else if (L.type == 2) {
            float NoD = dot(C.nrm_hit, rd);
            if (NoD > PTB_EPS) {
                float tt = dot(C.nrm_hit, L.pos - ro) / NoD;
                V3 hd = ro + tt * rd - L.pos;
                float u = dot(hd, C.dirx) / C.nsqx;
                float v = dot(hd, C.diry) / C.nsqy;
                if (-1.0f < u && u < 1.0f && -1.0f < v && v < 1.0f) { t = tt; area = C.area; }
            }
        }
        if (0.0f < t && t < ret.dis) {
            ret.dis = t; ret.pdf = (t * t) / area; ret.color = L.color; ret.hit = 1;
            break;
        }
    }
    return ret;
}
PTB_D LitSample light_sample_cached(const SceneParams* P, const SceneCache* SC, V3 hitpos, V3 samp) {
    LitSample ret; ret.dis = PTB_INF; ret.dir = v3s(0.0f); ret.pdf = 0.0f; ret.color = v3s(0.0f);
    int n = P->nlights;
    if (n != 0) {
        int i = clampi(ifloor(samp.z * (float)n), 0, n);   // inclusive upper clamp as in the reference
        const LightRec& L = P->lights[i];
        const LightCache& C = SC->light[i];
        V3 color = L.color, litpos = v3s(PTB_INF), nrm = v3s(0.0f);
        float area = 0.0f;
        if (L.type == 1) {
            litpos = L.pos + L.size * spherical(samp.x, samp.y);
            area = C.area;
        } else if (L.type == 2) {
            V3 disp = matvec(L.axes, mk3(samp.x * 2.0f - 1.0f, samp.y * 2.0f - 1.0f, 0.0f));
            nrm = C.nrm_smp;
            litpos = L.pos + L.size * disp;
            area = C.area;
        }
        V3 toli = litpos - hitpos;
        float dis = norm(toli);
        V3 dir = toli / dis;
        float pdf = (dis * dis) / area;
        color = color / pdf;
        if (any_ne0(nrm)) color = color * dot_or_zero(nrm, dir);
        ret.dis = dis; ret.dir = dir; ret.pdf = pdf; ret.color = color;
    }
    return ret;
}

// ---- light/world.py:22-29 WorldLight.at + common.py:234-239 dir2tex ----------------------------------------------------------
PTB_D V3 world_at(const SceneParams* P, const float4* __restrict__ texels, V3 dir) {
    V4 fac = mk4(P->world_fac[0], P->world_fac[1], P->world_fac[2], P->world_fac[3]);
    int texid = P->world_tex;
    if (texid != -1) {
        V3 d = normalized(mk3(dir.x, dir.z, -dir.y));
        float s = atan2f(d.z, d.x) / PTB_PI * 0.5f + 0.5f;
        float t = atan2f(d.y, sqrtf(d.x * d.x + d.z * d.z)) / PTB_PI + 0.5f;
        fac = fac * img_sample(P, texels, texid, s, t);
    }
    return mk3(fac.x, fac.y, fac.z);
}

// ---- camera.py:34-39 Camera.generate ----------------------------------------------------------------------------------------
PTB_D void camera_generate(const SceneParams* P, float x, float y, V3* ro, V3* rd) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float* m = &P->v2w[4 * i];
        a[i] = m[0] * x + m[1] * y + m[2] * -1.0f + m[3] * 1.0f;
        b[i] = m[0] * x + m[1] * y + m[2] * 1.0f + m[3] * 1.0f;
    }
    V3 o = mk3(a[0] / a[3], a[1] / a[3], a[2] / a[3]);
    V3 o1 = mk3(b[0] / b[3], b[1] / b[3], b[2] / b[3]);
    *ro = o;
    *rd = normalized(o1 - o);
}
