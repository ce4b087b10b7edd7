// ptina_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of archibate/ptina's per-pixel path-tracing hot path, written to be
// read next to the reference: every function cites the reference file:line it follows
// (paths relative to /root/reference/ptina/).  It is the parity checker for the CUDA
// library (tests/, __graft_entry__.smoke()) and the timed CPU baseline (bench.py
// cpu_baseline / --impl reference).  Nothing under ptina_b200/ may import, link or
// call it.
//
// Semantics: IEEE-754 binary32, operations in source order, no FMA contraction, no
// reassociation (build with -O2 -ffp-contract=off -fno-fast-math), i32 wraps, `%` and
// `//` Python-style, int() truncates, Matrix@ accumulates k ascending, .dot sums left
// to right, .normalized() = (1/|v|)*v  (Taichi 0.7 defaults: float=f32, int=i32).
//
// PARITY PINNING: Taichi / pysobol cannot be installed here, so the reference cannot
// be executed natively.  The oracle is pinned instead against golden vectors produced
// by running the reference's OWN Python sources (imported from /root/reference) under
// the Taichi-semantics interpreter shim in oracle/tishim/ (see tests/golden/README.md,
// tests/golden/make_golden.py).  Where a claim rests only on this restatement it says
// "parity unpinned" (MLT's ti.random stream).
//
// Algorithm differences from the reference: none on purpose.  The traversal here is the
// reference's unordered right-child-first DFS with a per-thread stack.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// common.py:32-33
constexpr float EPS = 1e-6f;
constexpr float INF = 1e6f;
constexpr float PI_F = 3.14159265358979323846f;   // ti.pi -> f32
constexpr float TAU_F = 6.28318530717958647692f;  // ti.tau -> f32
constexpr int SOBOL_DIM = 21201;                  // sampling/sobol.py:75
constexpr int SOBOL_ROWS = 21;                    // L + 1, L = 20

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 v3(float a) { return {a, a, a}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator+(V3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline V3 operator-(float s, V3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // (a*b).sum(), left to right
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float norm_sqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline float norm(V3 a) { return sqrtf(norm_sqr(a)); }
inline V3 normalized(V3 a) { float inv = 1.0f / norm(a); return inv * a; }  // Taichi Matrix.normalized
inline V3 vmin(V3 a, V3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline float vavg(V3 a) { return (a.x + a.y + a.z) / 3.0f; }               // common.py:73-77
inline bool vany_gt(V3 a, float s) { return a.x > s || a.y > s || a.z > s; }
inline bool vany_ne0(V3 a) { return a.x != 0 || a.y != 0 || a.z != 0; }

inline V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }  // common.py:164-165
inline int clampi(int x, int lo, int hi) { return std::min(hi, std::max(lo, x)); }
inline int ifloor(float x) {                                                           // common.py:169-170
    float f = floorf(x);
    if (!(f == f)) return INT32_MIN;  // x86 cvttss2si(NaN) = INT_MIN; clamps to 0 wherever the reference uses it
    if (f >= 2147483648.0f || f < -2147483648.0f) return INT32_MIN;
    return (int)f;
}
inline float dot_or_zero(V3 a, V3 b) { return fmaxf(0.0f, dot(a, b)); }                // common.py:178-180
inline float lerpf(float f, float a, float b) { return a * (1 - f) + b * f; }          // common.py:269-271
inline V3 lerp3(float f, V3 a, V3 b) { return a * (1 - f) + b * f; }
inline V3 lerp3(V3 f, V3 a, V3 b) { return a * (1.0f - f) + b * f; }
inline int pymod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

struct M33 { float m[3][3]; };
inline V3 matvec(const M33& a, V3 v) {  // Matrix @ Vector, k ascending
    V3 r;
    r.x = a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z;
    r.y = a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z;
    r.z = a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z;
    return r;
}

// common.py:213-217
inline M33 tanspace(V3 nrm) {
    V3 up = {233.f, 666.f, 512.f};
    V3 bitan = normalized(cross(nrm, up));
    V3 tan = cross(bitan, nrm);
    M33 r;
    r.m[0][0] = tan.x; r.m[0][1] = bitan.x; r.m[0][2] = nrm.x;
    r.m[1][0] = tan.y; r.m[1][1] = bitan.y; r.m[1][2] = nrm.y;
    r.m[2][0] = tan.z; r.m[2][1] = bitan.z; r.m[2][2] = nrm.z;
    return r;
}
// common.py:221-225
inline V3 spherical(float h, float p) {
    float ux = cosf(p * TAU_F), uy = sinf(p * TAU_F);
    float s = sqrtf(fmaxf(0.0f, 1 - h * h));
    return {s * ux, s * uy, h};
}
// common.py:247-249
inline V3 reflect(V3 I, V3 N) { return I - (2 * dot(N, I)) * N; }
// common.py:252-260
inline int refract(V3 I, V3 N, float eta, V3* T) {
    *T = I * 0.0f;
    float NoI = dot(N, I);
    float discr = 1 - (eta * eta) * (1 - NoI * NoI);
    if (discr > 0) {
        *T = normalized(eta * I - N * (eta * NoI + sqrtf(discr)));
        return 1;
    }
    return 0;
}

// ---- sampling/__init__.py:8-23 ------------------------------------------------------
inline int32_t wanghash(int32_t x) {
    uint32_t v = (uint32_t)x;
    v = (v ^ 61u) ^ (v >> 16);
    v *= 9u;
    v ^= v << 4;   // left shift, as written in the reference
    v *= 0x27d4eb2du;
    v ^= v >> 15;
    return (int32_t)v;
}
inline int32_t wanghash2(int32_t x, int32_t y) { return wanghash(y ^ wanghash(x)); }

struct Ray { V3 o, d; };
struct Hit { int hit; float depth; int index; float u, v; };

struct Light { int type; V3 color, pos; M33 axes; float size; };

struct Disney {
    V3 basecolor; float metallic, roughness, specular, specularTint, subsurface, sheen, sheenTint,
        clearcoat, clearcoatGloss, transmission, ior;
    V3 tintcolor, speccolor, sheencolor; float alpha, clearcoatAlpha;
};

// ---- materials/microfacet.py:8-77 ----------------------------------------------------
inline float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }  // x**5
inline float schlickFresnel(float cost) { return pow5(clampf(1 - cost, 0, 1)); }
inline float dielectricFresnel(float etai, float etao, float cosi) {
    float sini = sqrtf(fmaxf(0.0f, 1 - cosi * cosi));
    float sint = etao / etai * sini;
    float ret = 1.0f;
    if (sint < 1) {
        float cost = sqrtf(fmaxf(0.0f, 1 - sint * sint));
        float a1 = etai * cosi, a2 = etao * cost;
        float b1 = etao * cosi, b2 = etai * cost;
        float para = (a1 - a2) / (a1 + a2);
        float perp = (b1 - b2) / (b1 + b2);
        ret = 0.5f * (para * para + perp * perp);
    }
    return ret;
}
inline float GTR1(float cosh, float alpha) {
    float alpha2 = alpha * alpha;
    float t = 1 + (alpha2 - 1) * (cosh * cosh);
    return (alpha2 - 1) / (PI_F * logf(alpha2) * t);
}
inline float GTR2(float cosh, float alpha) {
    float alpha2 = alpha * alpha;
    float t = 1 + (alpha2 - 1) * (cosh * cosh);
    return alpha2 / (PI_F * (t * t));
}
inline float smithGGX(float cosi, float alpha) {
    float a = alpha * alpha, b = cosi * cosi;
    return 1 / (cosi + sqrtf(a + b - a * b));
}
inline V3 sample_GTR1(float u, float v, float alpha) {  // microfacet.py:68-71 (sic: /(alpha^2-1) outside sqrt)
    u = sqrtf(powf(alpha, 2 - 2 * u) - 1) / (alpha * alpha - 1);
    return spherical(u, v);
}
inline V3 sample_GTR2(float u, float v, float alpha) {  // microfacet.py:74-77
    u = sqrtf((1 - u) / (1 - u * (1 - alpha * alpha)));
    return spherical(u, v);
}

// ---- materials/disney.py:14-50 -------------------------------------------------------
inline void disney_init(Disney& m) {
    m.tintcolor = v3(1.0f);
    float luminance = dot(m.basecolor, V3{0.3f, 0.6f, 0.1f});
    if (luminance > EPS) m.tintcolor = m.basecolor / luminance;
    m.speccolor = lerp3(m.metallic, (m.specular * 0.08f) * lerp3(m.specularTint, v3(1.0f), m.tintcolor), m.basecolor);
    m.sheencolor = lerp3(m.sheenTint, v3(1.0f), m.tintcolor);
    m.alpha = fmaxf(0.001f, m.roughness * m.roughness);
    m.clearcoatAlpha = lerpf(m.clearcoatGloss, 0.1f, 0.001f);
}

// ---- materials/disney.py:52-106 ------------------------------------------------------
inline V3 disney_brdf(const Disney& m, V3 normal, float sign, V3 indir, V3 outdir) {
    float etai = 1.0f, etao = m.ior;
    if (sign < 0) { etai = m.ior; etao = 1.0f; }
    V3 halfdir = normalized(indir + outdir);
    float cosi = dot(indir, normal);
    float coso = dot(outdir, normal);
    float cosh = dot_or_zero(halfdir, normal);
    float cosoh = dot_or_zero(halfdir, outdir);
    V3 result = v3(0.0f);
    if (coso < 0) {
        if (cosi >= 0) {
            float Ds = GTR2(cosh, m.alpha);
            float fdf = dielectricFresnel(etao, etai, cosoh);
            V3 transmit = (1 / PI_F) * m.basecolor * (1 - fdf) * Ds;
            result = transmit * (1 - m.metallic) * m.transmission;
        }
    } else {
        float Fi = schlickFresnel(cosi);
        float Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2 * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float Fss90 = (cosoh * cosoh) * m.roughness;
        float Fss = lerpf(Fi, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
        float ss = 1.25f * (Fss * (1 / (cosi + coso) - 0.5f) + 0.5f);
        float Foh = schlickFresnel(cosoh);
        V3 Fsheen = (Foh * m.sheen) * m.sheencolor;
        float fdf = dielectricFresnel(etao, etai, cosoh);
        float Ds = GTR2(cosh, m.alpha);
        V3 Fs = lerp3(Foh, m.speccolor, v3(1.0f));
        float Gs = smithGGX(cosi, m.alpha) * smithGGX(coso, m.alpha);
        float Dr = GTR1(cosh, m.clearcoatAlpha);
        float Gr = smithGGX(cosi, 0.25f) * smithGGX(coso, 0.25f);
        float Fr = lerpf(Foh, 0.04f, 1.0f);
        V3 diffuse = ((1 / PI_F) * lerpf(m.subsurface, Fd, ss)) * m.basecolor + Fsheen;
        V3 specular = (Gs * Fs) * Ds + v3(0.25f * m.clearcoat * Gr * Fr * Dr);
        V3 transmit = ((1 / PI_F) * fdf * Ds) * m.basecolor;
        result = diffuse * (1 - m.metallic) * (1 - m.transmission);
        result = result + transmit * (1 - m.metallic) * m.transmission;
        result = result + specular * (1 - m.transmission);
    }
    return result;
}

struct BSDFSample { V3 outdir; float pdf; V3 color; };

// materials/__init__.py:21-48
struct Choice {
    float pdf, w;
    int operator()(float r) {
        if (w < r) { w /= r; pdf *= r; return 1; }
        w = (w - r) / (1 - r); pdf *= 1 - r; return 0;
    }
};

// ---- materials/disney.py:114-233 -----------------------------------------------------
inline BSDFSample disney_bounce(const Disney& m, V3 normal, float sign, V3 indir, V3 samp) {
    BSDFSample result{v3(0.0f), 0.0f, v3(0.0f)};
    float etai = 1.0f, etao = m.ior;
    if (sign < 0) { etai = m.ior; etao = 1.0f; }
    float eta = etai / etao;

    float cosi = dot(indir, normal);
    float Fi = schlickFresnel(cosi);
    V3 Fs = lerp3(Fi, m.speccolor, v3(1.0f));

    Choice choice{1.0f, samp.z};
    float specrate = lerpf(m.transmission, lerpf(m.metallic, vavg(Fs), 1.0f), 1.0f);
    float coatrate = 0.04f * m.clearcoat;
    specrate = lerpf(specrate, 0.1f, 1.0f);
    if (coatrate != 0) coatrate = lerpf(coatrate, 0.1f, 1.0f);

    if (choice(coatrate)) {
        float alpha = m.clearcoatAlpha;
        V3 halfdir = matvec(tanspace(normal), sample_GTR1(samp.x, samp.y, alpha));
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal);
        float cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0) {
            float Dr = GTR1(cosh, alpha);
            float Foh = schlickFresnel(cosoh);
            float Fr = lerpf(Foh, 0.04f, 1.0f);
            result.outdir = outdir;
            float partial = m.clearcoat * Fr * coso / cosoh;
            result.pdf = Dr * partial;
            result.color = v3(partial / choice.pdf);
        } else {
            result.pdf = 0.0f;
            result.color = v3(0.0f);
        }
    } else if (choice(specrate)) {
        float alpha = m.alpha;
        V3 halfdir = matvec(tanspace(normal), sample_GTR2(samp.x, samp.y, alpha));
        V3 outdir = reflect(-indir, halfdir);
        float coso = dot_or_zero(outdir, normal);
        float cosh = dot_or_zero(halfdir, normal);
        float cosoh = dot_or_zero(halfdir, outdir);
        if (cosoh > 0 && coso > 0 && cosh > 0) {
            float Ds = GTR2(cosh, alpha);
            if (choice(m.transmission)) {
                float fdf = dielectricFresnel(etao, etai, cosoh);
                float reflrate = lerpf(fdf, 0.2f, 1.0f);
                if (choice(reflrate)) {
                    result.outdir = outdir;
                    result.pdf = Ds * fdf;
                    result.color = m.basecolor * fdf * m.transmission / choice.pdf;
                } else {
                    V3 T;
                    int has_r = refract(-indir, halfdir, eta, &T);
                    if (has_r) {
                        result.outdir = T;
                        result.pdf = Ds * (1 - fdf);
                        result.color = m.basecolor * (1 - fdf) * m.transmission / choice.pdf;
                    }
                }
            } else {
                float Foh = schlickFresnel(cosoh);
                V3 Fs2 = lerp3(Foh, m.speccolor, v3(1.0f));
                result.outdir = outdir;
                float partial = 0.5f / (cosoh * smithGGX(coso, alpha));
                result.pdf = Ds * vavg(Fs2) * partial;
                result.color = Fs2 * partial * (1 - m.transmission) / choice.pdf;
            }
        } else {
            result.pdf = 0.0f;
            result.color = v3(0.0f);
        }
    } else {
        V3 outdir = matvec(tanspace(normal), spherical(sqrtf(samp.x), samp.y));
        V3 halfdir = normalized(indir + outdir);
        float cosi2 = dot(indir, normal);
        float coso = dot(outdir, normal);
        float cosoh = dot_or_zero(halfdir, outdir);
        float Fi2 = schlickFresnel(cosi2);
        float Fo = schlickFresnel(coso);
        float Fd90 = 0.5f + 2 * (cosoh * cosoh) * m.roughness;
        float Fd = lerpf(Fi2, 1.0f, Fd90) * lerpf(Fo, 1.0f, Fd90);
        float Fss90 = (cosoh * cosoh) * m.roughness;
        float Fss = lerpf(Fi2, 1.0f, Fss90) * lerpf(Fo, 1.0f, Fss90);
        float ss = 1.25f * (Fss * (1 / (cosi2 + coso) - 0.5f) + 0.5f);
        float Foh = schlickFresnel(cosoh);
        V3 Fsheen = (Foh * m.sheen) * m.sheencolor;
        V3 diffuse = ((1 / PI_F) * lerpf(m.subsurface, Fd, ss)) * m.basecolor + Fsheen;
        result.outdir = outdir;
        result.pdf = 1 / PI_F;
        result.color = diffuse * PI_F * (1 - m.metallic) * (1 - m.transmission) / choice.pdf;
    }
    return result;
}

// engine/path.py:10-14
inline float power_heuristic(float a, float b) {
    a = clampf(a, EPS, INF); a = a * a;
    b = clampf(b, EPS, INF); b = b * b;
    return a / (a + b);
}

struct Counters { long long rays, node_visits, box_tests, tri_tests, max_stack; };

struct Oracle {
    // model.py:13-16
    int nfaces = 0;
    std::vector<float> verts;     // [nfaces*3*8]
    std::vector<int32_t> mtlids;  // [nfaces]
    // mtllib.py:42-56 : 12 ParameterPair tables x 64 materials (zero-initialised fields)
    float mat_fac[64][12][4];
    int32_t mat_tex[64][12];
    // image.py:12-19
    std::vector<V4> texels;
    int img_nx[64], img_ny[64], img_base[64];
    // light/__init__.py:13-19
    Light lights[65];
    int nlights = 0;
    // light/world.py:10-16
    V4 world_fac{0.1f, 0.1f, 0.1f, 0.1f};
    int world_tex = 0;   // zero-initialised field (world.py:12)
    // camera.py:10-12
    float V2W[4][4];
    // filmtable.py:12-14
    int nx = 0, ny = 0;
    std::vector<V4> film[3];
    // tree/lbvh.py:50-59
    int n = 0;
    std::vector<int32_t> mc, id, child, leaf, bready;
    std::vector<V3> bmin, bmax;
    int aabb_sweeps = 0;
    // sampling/sobol.py:77-80
    std::vector<int32_t> sobolV;  // [21][D]
    int sobol_dim = SOBOL_DIM;
    int max_stack_seen = 0;

    Oracle() {
        memset(mat_fac, 0, sizeof mat_fac);
        memset(mat_tex, 0, sizeof mat_tex);
        memset(img_nx, 0, sizeof img_nx); memset(img_ny, 0, sizeof img_ny); memset(img_base, 0, sizeof img_base);
        memset(V2W, 0, sizeof V2W);
        memset(lights, 0, sizeof lights);
    }

    // ---- model.py:22-37 -------------------------------------------------------------
    inline V3 vpos(int face, int k) const { const float* p = &verts[(size_t)(face * 3 + k) * 8]; return {p[0], p[1], p[2]}; }
    inline V3 vnrm(int face, int k) const { const float* p = &verts[(size_t)(face * 3 + k) * 8]; return {p[3], p[4], p[5]}; }
    inline void vuv(int face, int k, float* u, float* v) const { const float* p = &verts[(size_t)(face * 3 + k) * 8]; *u = p[6]; *v = p[7]; }

    // ---- sampling/sobol.py:99-105 (closed form of reset()+k updates; see header) --------
    void sobol_point(int k, float* P) const {
        // X_k[j] = XOR over set bits b of gray(k) of V[b+1][j];  gray(k)=k^(k>>1)  (Antonov-Saleev)
        uint32_t g = (uint32_t)k ^ ((uint32_t)k >> 1);
        for (int j = 0; j < sobol_dim; j++) {
            uint32_t X = 0;
            for (int b = 0; b < 20; b++) if (g >> b & 1) X ^= (uint32_t)sobolV[(size_t)(b + 1) * sobol_dim + j];
            // construct_float sobol.py:19-29 : sum of bit_k * 2^-(k+1), MSB first (exact in f32: 20 bits)
            float ret = 0.0f, term = 0.5f; uint32_t value = X;
            while (value) { if (value & 0x80000000u) ret += term; value <<= 1; term *= 0.5f; }
            P[j] = ret;
        }
    }

    // ---- geometries.py:23-46 ----------------------------------------------------------
    static inline int box_intersect(V3 lo, V3 hi, const Ray& r) {
        float near = 0.0f, far = INF;
        int hit = 1;
        const float* l = &lo.x; const float* h = &hi.x; const float* o = &r.o.x; const float* d = &r.d.x;
        for (int i = 0; i < 3; i++) {
            if (fabsf(d[i]) < EPS) {
                if (o[i] < l[i] || o[i] > h[i]) hit = 0;
            } else {
                float i1 = (l[i] - o[i]) / d[i];
                float i2 = (h[i] - o[i]) / d[i];
                if (i1 > i2) std::swap(i1, i2);
                far = fminf(far, i2);
                near = fmaxf(near, i1);
                if (near > far) hit = 0;
            }
        }
        return hit;
    }

    // ---- geometries.py:117-148 --------------------------------------------------------
    inline Hit face_intersect(int index, const Ray& ray) const {
        V3 v0 = vpos(index, 0), v1 = vpos(index, 1), v2 = vpos(index, 2);
        V3 ro = ray.o, rd = ray.d;
        V3 u = v1 - v0, v = v2 - v0;
        V3 nrm = cross(u, v);
        Hit h{0, INF * 2, index, 0.f, 0.f};
        float b = dot(nrm, rd);
        if (fabsf(b) >= EPS) {
            V3 w0 = ro - v0;
            float a = -dot(nrm, w0);
            float r = a / b;
            if (r > 0) {
                V3 ip = ro + r * rd;
                float uu = dot(u, u), uv = dot(u, v), vv = dot(v, v);
                V3 w = ip - v0;
                float wu = dot(w, u), wv = dot(w, v);
                float D = uv * uv - uu * vv;
                float s = (uv * wv - vv * wu) / D;
                float t = (uv * wu - uu * wv) / D;
                h.u = s; h.v = t;
                if (0 <= s && s <= 1) {
                    if (0 <= t && s + t <= 1) { h.depth = r; h.hit = 1; }
                }
            }
        }
        return h;
    }

    // ---- tree/lbvh.py:313-347 (stack = stack.py:10-60, depth 32) -------------------------
    Hit bvh_intersect(const Ray& ray, int avoid, Counters* c) const {
        int stack[64]; int sp = 0, maxsp = 0;
        stack[sp++] = n;
        Hit ret{0, INF, -1, 0.f, 0.f};
        int ntimes = 0;
        if (c) c->rays++;
        while (ntimes < n && sp != 0) {
            int curr = stack[--sp];
            if (curr < n) {
                int index = leaf[curr];
                if (index != avoid) {
                    if (c) c->tri_tests++;
                    Hit h = face_intersect(index, ray);
                    if (h.hit != 0 && h.depth < ret.depth) { ret.depth = h.depth; ret.index = index; ret.u = h.u; ret.v = h.v; ret.hit = 1; }
                }
                continue;
            }
            int i = curr - n;
            if (c) c->box_tests++;
            if (box_intersect(bmin[i], bmax[i], ray) == 0) continue;
            ntimes++;
            if (c) c->node_visits++;
            if (sp + 2 > 64) { fprintf(stderr, "[oracle] traversal stack overflow (>64)\n"); abort(); }
            stack[sp++] = child[2 * i + 0];
            stack[sp++] = child[2 * i + 1];
            if (sp > maxsp) maxsp = sp;
        }
        if (c && maxsp > c->max_stack) c->max_stack = maxsp;
        return ret;
    }

    // brute-force closest hit over all triangles in the reference's visiting order for a VALID
    // tree (leaf slots descending); independent cross-check of bvh_intersect for rays whose
    // ancestors' boxes all pass.
    Hit brute_intersect(const Ray& ray, int avoid) const {
        Hit ret{0, INF, -1, 0.f, 0.f};
        for (int s = n - 1; s >= 0; s--) {
            int index = leaf[s];
            if (index == avoid) continue;
            Hit h = face_intersect(index, ray);
            if (h.hit != 0 && h.depth < ret.depth) { ret.depth = h.depth; ret.index = index; ret.u = h.u; ret.v = h.v; ret.hit = 1; }
        }
        return ret;
    }

    // ---- tree/lbvh.py:12-42 ---------------------------------------------------------------
    static inline int32_t expandBits(int32_t v_) {
        uint32_t v = (uint32_t)v_;  // i32 wrapping multiply == u32 multiply
        v = (v * 0x00010001u) & 0xFF0000FFu;
        v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u;
        v = (v * 0x00000005u) & 0x49249249u;
        return (int32_t)v;
    }
    static inline int32_t morton3D(V3 v) {
        int32_t wx = expandBits(clampi(ifloor(v.x * 1024), 0, 1023));
        int32_t wy = expandBits(clampi(ifloor(v.y * 1024), 0, 1023));
        int32_t wz = expandBits(clampi(ifloor(v.z * 1024), 0, 1023));
        return wx * 4 + wy * 2 + wz * 1;
    }
    static inline int clz_ref(int32_t x) {  // lbvh.py:33-42: min(clz32(x)+1, 32) for x >= 0
        int r = 0;
        while (true) {
            int32_t f = x >> (31 - r);
            if (f == 1 || r == 31) { r += 1; break; }
            r += 1;
        }
        return r;
    }

    // lbvh.py:61-90
    int findSplit(int l, int r) const {
        int m = 0;
        int32_t lc = mc[l], rc = mc[r];
        if (lc == rc) {
            m = (l + r) >> 1;
        } else {
            int cp = clz_ref(lc ^ rc);
            m = l;
            int s = r - l;
            while (true) {
                s += 1; s >>= 1;
                int nn = m + s;
                if (nn < r) {
                    int32_t nc = mc[nn];
                    int sp = clz_ref(lc ^ nc);
                    if (sp > cp) m = nn;
                }
                if (s <= 1) break;
            }
        }
        return m;
    }
    // lbvh.py:93-145
    void determineRange(int nn, int i, int* lo, int* ro) const {
        int l = 0, r = nn - 1;
        if (i != 0) {
            int32_t ic = mc[i], lc = mc[i - 1], rc = mc[i + 1];
            if (lc == ic && ic == rc) {
                l = i;
                while (i < nn - 1) {
                    i += 1;
                    if (i >= nn - 1) break;
                    if (mc[i] != mc[i + 1]) break;
                }
                r = i;
            } else {
                int ld = clz_ref(ic ^ lc), rd = clz_ref(ic ^ rc);
                int d = -1;
                if (rd > ld) d = 1;
                int delta_min = std::min(ld, rd);
                int lmax = 2;
                int delta = -1;
                int itmp = i + d * lmax;
                if (0 <= itmp && itmp < nn) delta = clz_ref(ic ^ mc[itmp]);
                while (delta > delta_min) {
                    lmax <<= 1;
                    itmp = i + d * lmax;
                    delta = -1;
                    if (0 <= itmp && itmp < nn) delta = clz_ref(ic ^ mc[itmp]);
                }
                int s = 0;
                int t = lmax >> 1;
                while (t > 0) {
                    itmp = i + (s + t) * d;
                    delta = -1;
                    if (0 <= itmp && itmp < nn) delta = clz_ref(ic ^ mc[itmp]);
                    if (delta > delta_min) s += t;
                    t >>= 1;
                }
                l = i; r = i + s * d;
                if (d < 0) std::swap(l, r);
            }
        }
        *lo = l; *ro = r;
    }

    // lbvh.py:168-305 : build().  Returns number of AABB sweeps, or -1 for 'AABB step never stop!'
    int build_tree() {
        n = nfaces;
        mc.assign(n, 0); id.assign(n, 0); leaf.assign(n, 0);
        child.assign((size_t)2 * std::max(n - 1, 1), 0);
        bmin.assign(std::max(n - 1, 1), v3(0)); bmax.assign(std::max(n - 1, 1), v3(0)); bready.assign(std::max(n, 1), 0);
        // genMortonCodes lbvh.py:168-183
        V3 cmin = v3(INF), cmax = v3(-INF);
        for (int i = 0; i < n; i++) {
            V3 center = (vpos(i, 0) + vpos(i, 1) + vpos(i, 2)) / 3.0f;  // getCenter lbvh.py:161-165
            cmax = vmax(cmax, center); cmin = vmin(cmin, center);
        }
        for (int i = 0; i < n; i++) {
            V3 center = (vpos(i, 0) + vpos(i, 1) + vpos(i, 2)) / 3.0f;
            V3 coord = (center - cmin) / (cmax - cmin);
            mc[i] = morton3D(coord);
            id[i] = i;
        }
        // sortMortonCodes lbvh.py:186-208 : np.argsort; tie order canonicalised to STABLE
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return mc[a] < mc[b]; });
        {
            std::vector<int32_t> mc2(n), id2(n);
            for (int i = 0; i < n; i++) { mc2[i] = mc[order[i]]; id2[i] = id[order[i]]; }
            mc.swap(mc2); id.swap(id2);
        }
        // genHierarchy lbvh.py:211-231
        for (int i = 0; i < n; i++) leaf[i] = id[i];
        for (int i = 0; i < n - 1; i++) {
            int l, r; determineRange(n, i, &l, &r);
            int split = findSplit(l, r);
            int lhs = split; if (lhs != l) lhs += n;
            int rhs = split + 1; if (rhs != r) rhs += n;
            child[2 * i + 0] = lhs; child[2 * i + 1] = rhs;
        }
        // genAABBs lbvh.py:234-294
        int count = 1;
        while (!aabb_substep()) {
            count += 1;
            if (count > 64) { aabb_sweeps = count; return -1; }
        }
        aabb_sweeps = count;
        return count;
    }
    bool node_bbox(int i, V3* lo, V3* hi, const std::vector<int32_t>& ready) const {  // getNodeBoundingBox lbvh.py:234-248
        if (i < n) {
            int f = leaf[i];
            V3 a = vpos(f, 0), b = vpos(f, 1), c = vpos(f, 2);
            *lo = vmin(vmin(a, b), c); *hi = vmax(vmax(a, b), c);
            return true;
        }
        i -= n;
        *lo = bmin[i]; *hi = bmax[i];
        return ready[i] != 0;
    }
    bool aabb_substep() {  // genAABBSubstep lbvh.py:272-294 (sequential sweep; boxes are order-independent)
        for (int i = 0; i < n - 1; i++) {
            if (bready[i]) continue;
            V3 lo1, hi1, lo2, hi2;
            bool r1 = node_bbox(child[2 * i + 0], &lo1, &hi1, bready);
            bool r2 = node_bbox(child[2 * i + 1], &lo2, &hi2, bready);
            if (r1 && r2) { bmin[i] = vmin(lo1, lo2); bmax[i] = vmax(hi1, hi2); bready[i] = 1; }
        }
        for (int i = 0; i < n - 1; i++) if (bready[i] == 0) return false;
        return true;
    }

    // ---- image.py:21-24,137-148, common.py:183-192 -----------------------------------------
    inline V4 img_fetch(int id, int x, int y) const {
        if (img_nx[id] <= 0 || img_ny[id] <= 0) return V4{0, 0, 0, 0};  // unloaded image: `x % 0` in the reference (UB); fenced
        x = pymod(x, img_nx[id]); y = pymod(y, img_ny[id]);
        return texels[(size_t)img_base[id] + (size_t)x * img_ny[id] + y];
    }
    inline V4 img_sample(int id, float u, float v) const {
        float px = u * (float)(img_nx[id] - 1), py = v * (float)(img_ny[id] - 1);
        int Ix = ifloor(px), Iy = ifloor(py);
        float x0 = px - (float)Ix, x1 = py - (float)Iy;
        float y0 = 1 - x0, y1 = 1 - x1;
        return img_fetch(id, Ix + 1, Iy + 1) * x0 * x1 + img_fetch(id, Ix + 1, Iy) * x0 * y1 +
               img_fetch(id, Ix, Iy) * y0 * y1 + img_fetch(id, Ix, Iy + 1) * y0 * x1;
    }
    // ---- mtllib.py:30-38 --------------------------------------------------------------------
    inline V4 param_get(int slot, int mtlid, float u, float v, float dflt) const {
        V4 fac{dflt, dflt, dflt, dflt};
        if (mtlid != -1) {
            const float* f = mat_fac[mtlid][slot];
            fac = V4{f[0], f[1], f[2], f[3]};
            int texid = mat_tex[mtlid][slot];
            if (texid != -1) fac = fac * img_sample(texid, u, v);
        }
        return fac;
    }
    // ---- mtllib.py:79-95 --------------------------------------------------------------------
    inline Disney material_get(int mtlid, float u, float v) const {
        Disney m;
        V4 b = param_get(0, mtlid, u, v, 0.8f);
        m.basecolor = {b.x, b.y, b.z};
        m.metallic = param_get(1, mtlid, u, v, 0.0f).x;
        m.roughness = param_get(2, mtlid, u, v, 0.4f).x;
        m.specular = param_get(3, mtlid, u, v, 0.5f).x;
        m.specularTint = param_get(4, mtlid, u, v, 0.4f).x;
        m.subsurface = param_get(5, mtlid, u, v, 0.0f).x;
        m.sheen = param_get(6, mtlid, u, v, 0.0f).x;
        m.sheenTint = param_get(7, mtlid, u, v, 0.4f).x;
        m.clearcoat = param_get(8, mtlid, u, v, 0.0f).x;
        m.clearcoatGloss = param_get(9, mtlid, u, v, 0.5f).x;
        m.transmission = param_get(10, mtlid, u, v, 0.0f).x;
        m.ior = param_get(11, mtlid, u, v, 1.45f).x;
        disney_init(m);
        return m;
    }
    // ---- model.py:88-101, geometries.py:96-108 ---------------------------------------------
    inline void get_geometries(const Hit& hit, const Ray& r, V3* hitpos, V3* normal, float* sign, Disney* mat) const {
        int f = hit.index;
        float u = hit.u, v = hit.v;
        float wx = 1 - u - v, wy = u, wz = v;
        V3 nrm = normalized(wx * vnrm(f, 0) + wy * vnrm(f, 1) + wz * vnrm(f, 2));
        float t0u, t0v, t1u, t1v, t2u, t2v;
        vuv(f, 0, &t0u, &t0v); vuv(f, 1, &t1u, &t1v); vuv(f, 2, &t2u, &t2v);
        float tu = wx * t0u + wy * t1u + wz * t2u;
        float tv = wx * t0v + wy * t1v + wz * t2v;
        *hitpos = r.o + hit.depth * r.d;
        float sg = -dot(r.d, nrm);
        if (sg < 0) nrm = -nrm;
        *normal = nrm; *sign = sg;
        *mat = material_get(mtlids[f], tu, tv);
    }

    // ---- geometries.py:158-179 ----------------------------------------------------------------
    static inline float sphere_intersect(V3 pos, float rad2, const Ray& ray) {
        float ret = 0.0f;
        V3 op = pos - ray.o;
        float b = dot(op, ray.d);
        float det = b * b + rad2 - norm_sqr(op);
        if (det < 0) ret = 0.0f;
        else {
            det = sqrtf(det);
            float t = b - det;
            if (t > EPS) ret = t;
            else { t = b + det; ret = t > EPS ? t : 0.0f; }
        }
        return ret;
    }
    // ---- geometries.py:57-73 --------------------------------------------------------------------
    static inline int area_intersect(V3 pos, V3 dirx, V3 diry, const Ray& ray, float* tout) {
        float t = INF; int hit = 0;
        V3 nrm = normalized(cross(dirx, diry));
        float NoD = dot(nrm, ray.d);
        if (NoD > EPS) {
            t = dot(nrm, pos - ray.o) / NoD;
            V3 hitdisp = ray.o + t * ray.d - pos;
            float u = dot(hitdisp, dirx) / norm_sqr(dirx);
            float v = dot(hitdisp, diry) / norm_sqr(diry);
            if (-1 < u && u < 1 && -1 < v && v < 1) hit = 1;
        }
        *tout = t;
        return hit;
    }
    struct LitHit { int hit; float dis, pdf; V3 color; };
    // ---- light/__init__.py:51-81 ------------------------------------------------------------------
    inline LitHit light_hit(const Ray& ray) const {
        LitHit ret{0, INF, 0.0f, v3(0.0f)};
        for (int i = 0; i < nlights; i++) {
            const Light& L = lights[i];
            float t = 0.0f, area = 0.0f;
            if (L.type == 1) {
                t = sphere_intersect(L.pos, L.size * L.size, ray);
                area = PI_F * (L.size * L.size);
            } else if (L.type == 2) {
                V3 dirx = matvec(L.axes, V3{L.size, 0.0f, 0.0f});
                V3 diry = matvec(L.axes, V3{0.0f, L.size, 0.0f});
                float tt;
                if (area_intersect(L.pos, dirx, diry, ray, &tt)) { t = tt; area = 4 * (L.size * L.size); }
            }
            if (0 < t && t < ret.dis) {
                ret.dis = t; ret.pdf = (ret.dis * ret.dis) / area; ret.color = L.color; ret.hit = 1;
                break;
            }
        }
        return ret;
    }
    struct LitSample { float dis; V3 dir; float pdf; V3 color; };
    // ---- light/__init__.py:83-121 -----------------------------------------------------------------
    inline LitSample light_sample(V3 hitpos, V3 samp) const {
        LitSample ret{INF, v3(0.0f), 0.0f, v3(0.0f)};
        if (nlights != 0) {
            int i = clampi(ifloor(samp.z * (float)nlights), 0, nlights);  // inclusive upper clamp, sic
            const Light& L = lights[i];
            V3 color = L.color;
            V3 litpos = v3(INF), nrm = v3(0.0f);
            float area = 0.0f;
            if (L.type == 1) {
                V3 disp = spherical(samp.x, samp.y);
                litpos = L.pos + L.size * disp;
                area = PI_F * (L.size * L.size);
            } else if (L.type == 2) {
                V3 disp = matvec(L.axes, V3{samp.x * 2 - 1, samp.y * 2 - 1, 0.0f});
                nrm = matvec(L.axes, V3{0.0f, 0.0f, 1.0f});
                litpos = L.pos + L.size * disp;
                area = 4 * (L.size * L.size);
            }
            V3 toli = litpos - hitpos;
            float dis = norm(toli);
            V3 dir = toli / dis;
            float pdf = (dis * dis) / area;
            color = color / pdf;
            if (vany_ne0(nrm)) color = color * dot_or_zero(nrm, dir);
            ret = LitSample{dis, dir, pdf, color};
        }
        return ret;
    }
    // ---- light/world.py:22-29, common.py:234-239 ------------------------------------------------
    inline V3 world_at(V3 dir) const {
        V4 fac = world_fac;
        if (world_tex != -1) {
            V3 d2{dir.x, dir.z, -dir.y};
            d2 = normalized(d2);
            float s = atan2f(d2.z, d2.x) / PI_F * 0.5f + 0.5f;
            float t = atan2f(d2.y, sqrtf(d2.x * d2.x + d2.z * d2.z)) / PI_F + 0.5f;
            fac = fac * img_sample(world_tex, s, t);
        }
        return {fac.x, fac.y, fac.z};
    }

    // ---- camera.py:34-39 ---------------------------------------------------------------------------
    inline Ray camera_generate(float x, float y) const {
        float a[4], b[4];
        for (int i = 0; i < 4; i++) {
            a[i] = V2W[i][0] * x + V2W[i][1] * y + V2W[i][2] * -1.0f + V2W[i][3] * 1.0f;
            b[i] = V2W[i][0] * x + V2W[i][1] * y + V2W[i][2] * 1.0f + V2W[i][3] * 1.0f;
        }
        V3 ro{a[0] / a[3], a[1] / a[3], a[2] / a[3]};
        V3 ro1{b[0] / b[3], b[1] / b[3], b[2] / b[3]};
        return Ray{ro, normalized(ro1 - ro)};
    }

    struct Rng {  // sampling/sobol.py:107-125 (Proxy) or sampling/__init__.py:53-64 (RNGProxy, MLT)
        const float* P; int32_t i; int dim; const float* X; int j;
        inline float random() {
            if (X) return X[j++];
            float r = P[pymod(i, dim)];
            i = (int32_t)((uint32_t)i + 1u);
            return r;
        }
    };

    // ---- engine/path.py:18-64 ---------------------------------------------------------------------
    V3 path_trace(Ray r, Rng& rng, Counters* c) const {
        int avoid = -1, depth = 0;
        V3 result = v3(0.0f), throughput = v3(1.0f);
        float last_brdf_pdf = 0.0f;
        while (depth < 5 && vany_gt(throughput, 0.0f) && vany_ne0(r.d)) {
            depth += 1;
            r.d = normalized(r.d);
            Hit hit = bvh_intersect(r, avoid, c);
            LitHit lit = light_hit(r);
            if (lit.hit != 0 && (hit.hit == 0 || lit.dis < hit.depth)) {
                float mis = power_heuristic(last_brdf_pdf, lit.pdf);
                V3 direct_li = mis * lit.color;
                result = result + throughput * direct_li;
            }
            if (hit.hit == 0) {
                result = result + throughput * world_at(r.d);
                break;
            }
            avoid = hit.index;
            V3 hitpos, normal; float sign; Disney material;
            get_geometries(hit, r, &hitpos, &normal, &sign, &material);
            sign = -dot(r.d, normal);
            if (sign < 0) normal = -normal;

            float s0 = rng.random(), s1 = rng.random(), s2 = rng.random();
            LitSample li = light_sample(hitpos, V3{s0, s1, s2});
            if (vany_gt(li.color, 0.0f)) {
                Hit occ = bvh_intersect(Ray{hitpos, li.dir}, avoid, c);
                if (occ.hit == 0 || occ.depth > li.dis) {
                    V3 brdf_clr = disney_brdf(material, normal, sign, -r.d, li.dir);
                    float brdf_pdf = vavg(brdf_clr);
                    float mis = power_heuristic(li.pdf, brdf_pdf);
                    V3 direct_li = mis * li.color * brdf_clr * dot_or_zero(normal, li.dir);
                    result = result + throughput * direct_li;
                }
            }
            float b0 = rng.random(), b1 = rng.random(), b2 = rng.random();
            BSDFSample brdf = disney_bounce(material, normal, sign, -r.d, V3{b0, b1, b2});
            throughput = throughput * brdf.color;
            r.o = hitpos;
            r.d = brdf.outdir;
            last_brdf_pdf = brdf.pdf;
        }
        return result;
    }
    // ---- engine/brute.py:29-62 -------------------------------------------------------------------
    V3 brute_trace(Ray r, Rng& rng, Counters* c) const {
        int avoid = -1, depth = 0;
        V3 result = v3(0.0f), throughput = v3(1.0f);
        while (depth < 5 && vany_gt(throughput, EPS)) {
            depth += 1;
            r.d = normalized(r.d);
            Hit hit = bvh_intersect(r, avoid, c);
            LitHit lit = light_hit(r);
            if (lit.hit != 0 && (hit.hit == 0 || lit.dis < hit.depth)) result = result + throughput * lit.color;
            if (hit.hit == 0) { result = result + throughput * world_at(r.d); break; }
            avoid = hit.index;
            V3 hitpos, normal; float sign; Disney material;
            get_geometries(hit, r, &hitpos, &normal, &sign, &material);
            sign = -dot(r.d, normal);
            if (sign < 0) normal = -normal;
            float b0 = rng.random(), b1 = rng.random(), b2 = rng.random();
            BSDFSample brdf = disney_bounce(material, normal, sign, -r.d, V3{b0, b1, b2});
            throughput = throughput * brdf.color;
            r.o = hitpos;
            r.d = brdf.outdir;
        }
        return result;
    }
    // engine/path.py:85-93 / brute.py:64-75 : one sample of pixel (i,j) at Sobol point table P
    inline Ray primary_ray(int i, int j, Rng& rng) const {
        float dx = rng.random(), dy = rng.random();
        float x = ((float)i + dx) / (float)nx * 2 - 1;
        float y = ((float)j + dy) / (float)ny * 2 - 1;
        return camera_generate(x, y);
    }
};

}  // namespace

// =============================== extern "C" surface (ctypes) =================================
extern "C" {

void* ora_create() { return new Oracle(); }
void ora_destroy(void* h) { delete (Oracle*)h; }

int ora_set_sobol(void* h, const int32_t* V, int rows, int dim) {
    Oracle* o = (Oracle*)h;
    if (rows != SOBOL_ROWS) return -1;
    o->sobolV.assign(V, V + (size_t)rows * dim);
    o->sobol_dim = dim;
    return 0;
}
int ora_sobol_point(void* h, int k, float* P) { ((Oracle*)h)->sobol_point(k, P); return 0; }

int ora_load_model(void* h, const float* verts, const int32_t* mtlids, int nfaces) {
    Oracle* o = (Oracle*)h;
    o->nfaces = nfaces;
    o->verts.assign(verts, verts + (size_t)nfaces * 24);
    o->mtlids.assign(mtlids, mtlids + nfaces);
    return 0;
}
int ora_load_materials(void* h, const float* fac, const int32_t* tex, int nmat) {
    Oracle* o = (Oracle*)h;
    if (nmat > 64) return -1;
    memcpy(o->mat_fac, fac, sizeof(float) * nmat * 48);
    memcpy(o->mat_tex, tex, sizeof(int32_t) * nmat * 12);
    return 0;
}
int ora_load_images(void* h, const float* texels, long long ntexels, const int32_t* nx, const int32_t* ny, const int32_t* base, int nimg) {
    Oracle* o = (Oracle*)h;
    if (nimg > 64) return -1;
    o->texels.resize(ntexels);
    if (ntexels) memcpy(o->texels.data(), texels, sizeof(float) * 4 * ntexels);
    for (int i = 0; i < nimg; i++) { o->img_nx[i] = nx[i]; o->img_ny[i] = ny[i]; o->img_base[i] = base[i]; }
    return 0;
}
int ora_clear_lights(void* h) { ((Oracle*)h)->nlights = 0; return 0; }
int ora_add_light(void* h, const float* pos, const float* axes, const float* color, float size, int type) {
    Oracle* o = (Oracle*)h;
    if (o->nlights >= 64) return -1;
    Light& L = o->lights[o->nlights];
    L.type = type; L.size = size;
    L.pos = {pos[0], pos[1], pos[2]}; L.color = {color[0], color[1], color[2]};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L.axes.m[i][j] = axes[i * 3 + j];
    return o->nlights++;
}
int ora_set_world_light(void* h, const float* fac, int tex) {
    Oracle* o = (Oracle*)h;
    o->world_fac = {fac[0], fac[1], fac[2], fac[3]}; o->world_tex = tex;
    return 0;
}
int ora_set_camera(void* h, const float* V2W) { memcpy(((Oracle*)h)->V2W, V2W, 64); return 0; }
int ora_set_size(void* h, int nx, int ny) {
    Oracle* o = (Oracle*)h;
    o->nx = nx; o->ny = ny;
    for (auto& f : o->film) f.assign((size_t)nx * ny, V4{0, 0, 0, 0});
    return 0;
}
int ora_clear(void* h) {  // filmtable.py:44-45 zeroes every pass
    Oracle* o = (Oracle*)h;
    for (auto& f : o->film) std::fill(f.begin(), f.end(), V4{0, 0, 0, 0});
    return 0;
}
int ora_build_tree(void* h) { return ((Oracle*)h)->build_tree(); }
int ora_export_tree(void* h, int32_t* mc, int32_t* id, int32_t* child, int32_t* leaf, float* bmin, float* bmax) {
    Oracle* o = (Oracle*)h;
    int n = o->n;
    memcpy(mc, o->mc.data(), 4 * n); memcpy(id, o->id.data(), 4 * n); memcpy(leaf, o->leaf.data(), 4 * n);
    if (n > 1) {
        memcpy(child, o->child.data(), 8 * (n - 1));
        memcpy(bmin, o->bmin.data(), 12 * (n - 1)); memcpy(bmax, o->bmax.data(), 12 * (n - 1));
    }
    return n;
}

// Tree validator: every leaf slot referenced exactly once, every internal node != root
// referenced exactly once, DFS from the root visits 2n-1 nodes, child0 covers [l..split] and child1
// [split+1..r] so leaf slots come out in ascending order, max depth reported.
// returns max depth (root = 1) or a negative error code.
int ora_validate_tree(void* h) {
    Oracle* o = (Oracle*)h;
    int n = o->n;
    if (n < 2) return n == 1 ? 1 : -1;
    std::vector<int> refs(2 * n, 0);
    for (int i = 0; i < n - 1; i++) for (int k = 0; k < 2; k++) {
        int c = o->child[2 * i + k];
        if (c < 0 || c >= 2 * n - 1) return -2;
        refs[c]++;
    }
    for (int s = 0; s < n; s++) if (refs[s] != 1) return -3;
    for (int i = 1; i < n - 1; i++) if (refs[n + i] != 1) return -4;
    if (refs[n] != 0) return -5;
    std::vector<std::pair<int, int>> st;
    st.push_back({n, 1});
    long long visited = 0; int maxd = 0; int expect = n - 1;
    while (!st.empty()) {
        auto [c, d] = st.back(); st.pop_back();
        visited++;
        if (visited > 2LL * n) return -6;
        maxd = std::max(maxd, d);
        if (c < n) { if (c != expect) return -7; expect--; continue; }
        st.push_back({o->child[2 * (c - n)], d + 1});
        st.push_back({o->child[2 * (c - n) + 1], d + 1});
    }
    if (visited != 2LL * n - 1) return -8;
    return maxd;
}

// rays: [m][6] (o, d as given -- NOT re-normalised), avoid: [m] or NULL.  policy 0 = reference DFS,
// 1 = brute force over all triangles.
int ora_intersect(void* h, const float* rays, const int32_t* avoid, int m, int policy,
                  int32_t* hit, float* depth, int32_t* index, float* uv, long long* counters) {
    Oracle* o = (Oracle*)h;
    Counters tot{0, 0, 0, 0, 0};
#pragma omp parallel
    {
        Counters c{0, 0, 0, 0, 0};
#pragma omp for schedule(dynamic, 256)
        for (int q = 0; q < m; q++) {
            Ray r{{rays[6 * q], rays[6 * q + 1], rays[6 * q + 2]}, {rays[6 * q + 3], rays[6 * q + 4], rays[6 * q + 5]}};
            int av = avoid ? avoid[q] : -1;
            Hit ht = policy == 0 ? o->bvh_intersect(r, av, &c) : o->brute_intersect(r, av);
            hit[q] = ht.hit; depth[q] = ht.depth; index[q] = ht.index; uv[2 * q] = ht.u; uv[2 * q + 1] = ht.v;
        }
#pragma omp critical
        { tot.rays += c.rays; tot.node_visits += c.node_visits; tot.box_tests += c.box_tests; tot.tri_tests += c.tri_tests; tot.max_stack = std::max(tot.max_stack, c.max_stack); }
    }
    if (counters) { counters[0] = tot.rays; counters[1] = tot.node_visits; counters[2] = tot.box_tests; counters[3] = tot.tri_tests; counters[4] = tot.max_stack; }
    return 0;
}

// primary rays of Sobol point k for every pixel: rays[nx*ny][6] (index x*ny+y), and their closest hits
int ora_primary(void* h, int k, float* rays, int32_t* hit, float* depth, int32_t* index, float* uv) {
    Oracle* o = (Oracle*)h;
    std::vector<float> P(o->sobol_dim);
    o->sobol_point(k, P.data());
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < o->nx; i++)
        for (int j = 0; j < o->ny; j++) {
            Oracle::Rng rng{P.data(), wanghash2(i, j), o->sobol_dim, nullptr, 0};
            Ray r = o->primary_ray(i, j, rng);
            size_t q = (size_t)i * o->ny + j;
            if (rays) { rays[6 * q] = r.o.x; rays[6 * q + 1] = r.o.y; rays[6 * q + 2] = r.o.z; rays[6 * q + 3] = r.d.x; rays[6 * q + 4] = r.d.y; rays[6 * q + 5] = r.d.z; }
            if (hit) {
                Ray rn = r; rn.d = normalized(rn.d);  // path.py:28 normalises before intersecting
                Hit ht = o->bvh_intersect(rn, -1, nullptr);
                hit[q] = ht.hit; depth[q] = ht.depth; index[q] = ht.index; uv[2 * q] = ht.u; uv[2 * q + 1] = ht.v;
            }
        }
    return 0;
}

// engine: 0 = PathEngine, 1 = BruteEngine.  Renders Sobol points k_first .. k_first+nsamples-1 and
// accumulates film pass 0 (+= (rgb,1) per sample, in k order).  counters (optional): rays, node_visits,
// box_tests, tri_tests, max_stack.  nthreads <= 0 -> OpenMP default.
int ora_render(void* h, int engine, int k_first, int nsamples, int nthreads, long long* counters) {
    Oracle* o = (Oracle*)h;
    if ((int)o->sobolV.size() == 0 || o->nx * o->ny == 0) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    Counters tot{0, 0, 0, 0, 0};
    std::vector<float> P(o->sobol_dim);
    for (int s = 0; s < nsamples; s++) {
        o->sobol_point(k_first + s, P.data());
#pragma omp parallel
        {
            Counters c{0, 0, 0, 0, 0};
#pragma omp for schedule(dynamic, 8)
            for (int i = 0; i < o->nx; i++)
                for (int j = 0; j < o->ny; j++) {
                    Oracle::Rng rng{P.data(), wanghash2(i, j), o->sobol_dim, nullptr, 0};
                    Ray r = o->primary_ray(i, j, rng);
                    V3 clr = engine == 0 ? o->path_trace(r, rng, &c) : o->brute_trace(r, rng, &c);
                    V4& f = o->film[0][(size_t)i * o->ny + j];
                    f = f + V4{clr.x, clr.y, clr.z, 1.0f};
                }
#pragma omp critical
            { tot.rays += c.rays; tot.node_visits += c.node_visits; tot.box_tests += c.box_tests; tot.tri_tests += c.tri_tests; tot.max_stack = std::max(tot.max_stack, c.max_stack); }
        }
    }
    if (counters) { counters[0] = tot.rays; counters[1] = tot.node_visits; counters[2] = tot.box_tests; counters[3] = tot.tri_tests; counters[4] = tot.max_stack; }
    return 0;
}

// engine/preview.py:19-41 PreviewEngine.render() x nsamples: Sobol points k_first .., primary hit of the camera ray AS GENERATED
// (preview.py:33 does not re-normalise it, unlike path.py:28), film 1 += (albedo, 1), film 2 += (normal, 1); misses add (0, 1)
int ora_render_preview(void* h, int k_first, int nsamples) {
    Oracle* o = (Oracle*)h;
    if ((int)o->sobolV.size() == 0 || o->nx * o->ny == 0) return -1;
    std::vector<float> P(o->sobol_dim);
    for (int s = 0; s < nsamples; s++) {
        o->sobol_point(k_first + s, P.data());
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < o->nx; i++)
            for (int j = 0; j < o->ny; j++) {
                Oracle::Rng rng{P.data(), wanghash2(i, j), o->sobol_dim, nullptr, 0};
                Ray r = o->primary_ray(i, j, rng);
                V3 albedo = v3(0.0f), normal = v3(0.0f);
                Hit hit = o->bvh_intersect(r, -1, nullptr);
                if (hit.hit == 1) {
                    V3 hitpos; float sign; Disney material;
                    o->get_geometries(hit, r, &hitpos, &normal, &sign, &material);
                    albedo = material.basecolor;
                }
                size_t q = (size_t)i * o->ny + j;
                o->film[1][q] = o->film[1][q] + V4{albedo.x, albedo.y, albedo.z, 1.0f};
                o->film[2][q] = o->film[2][q] + V4{normal.x, normal.y, normal.z, 1.0f};
            }
    }
    return 0;
}

// engine/path.py:96-118 render_tile(i, j, samples) at Sobol point k (the reference text is commented out; exams/benchtiles.py:25
// calls it): samples m = 0..min(samples, 63) inclusive (`if m > samples: continue`) of every pixel of the 64x64 tile, sample m drawn
// through get_proxy(wanghash3(x, y, m)).  Pixels outside the film are skipped: the text's `x > nx` / `y > ny` lets x == nx write
// past the image and y == ny alias pixel (x+1, 0); tests/golden/make_golden_r2.py records what the text itself does there.
int ora_render_tile(void* h, int engine, int k, int ti, int tj, int samples) {
    Oracle* o = (Oracle*)h;
    if ((int)o->sobolV.size() == 0 || o->nx * o->ny == 0) return -1;
    std::vector<float> P(o->sobol_dim);
    o->sobol_point(k, P.data());
    const int ms = std::min(samples, 63);
#pragma omp parallel for schedule(dynamic, 1)
    for (int kx = 0; kx < 64; kx++)
        for (int l = 0; l < 64; l++) {
            const int x = ti * 64 + kx, y = tj * 64 + l;
            if (x >= o->nx || y >= o->ny) continue;
            for (int m = 0; m <= ms; m++) {
                Oracle::Rng rng{P.data(), wanghash(m ^ wanghash2(x, y)), o->sobol_dim, nullptr, 0};
                Ray r = o->primary_ray(x, y, rng);
                V3 clr = engine == 0 ? o->path_trace(r, rng, nullptr) : o->brute_trace(r, rng, nullptr);
                V4& f = o->film[0][(size_t)x * o->ny + y];
                f = f + V4{clr.x, clr.y, clr.z, 1.0f};
            }
        }
    return 0;
}

// common.py:337-352, literal: erfinv (Winitzki, a = 0.147; the constants 2/(pi*0.147) and 1/0.147 are Python floats folded to f32
// when they meet an f32 operand) and normaldist(samp) = sqrt(2) * erfinv(samp * 2 - 1)
static inline float erfinv_ref(float x) {
    float sgn = x < 0 ? -1.0f : 1.0f;
    x = (1.0f - x) * (1.0f + x);
    float lnx = logf(x);
    float tt1 = (float)(2.0 / (3.14159265358979323846 * 0.147)) + 0.5f * lnx;
    float tt2 = (float)(1.0 / 0.147) * lnx;
    return sgn * sqrtf(-tt1 + sqrtf(tt1 * tt1 - tt2));
}
int ora_normaldist(const float* in, int m, float* out) {
    for (int q = 0; q < m; q++) out[q] = sqrtf(2.0f) * erfinv_ref(in[q] * 2.0f - 1.0f);
    return 0;
}
int ora_erfinv(const float* in, int m, float* out) {
    for (int q = 0; q < m; q++) out[q] = erfinv_ref(in[q]);
    return 0;
}

// per-pixel radiance of ONE sample (no film), for golden comparisons: out[nx*ny][3]
int ora_render_sample(void* h, int engine, int k, float* out) {
    Oracle* o = (Oracle*)h;
    std::vector<float> P(o->sobol_dim);
    o->sobol_point(k, P.data());
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < o->nx; i++)
        for (int j = 0; j < o->ny; j++) {
            Oracle::Rng rng{P.data(), wanghash2(i, j), o->sobol_dim, nullptr, 0};
            Ray r = o->primary_ray(i, j, rng);
            V3 clr = engine == 0 ? o->path_trace(r, rng, nullptr) : o->brute_trace(r, rng, nullptr);
            size_t q = (size_t)i * o->ny + j;
            out[3 * q] = clr.x; out[3 * q + 1] = clr.y; out[3 * q + 2] = clr.z;
        }
    return 0;
}

// path_trace driven by explicit per-path random vectors X[m][32] (RNGProxy, sampling/__init__.py:53-64),
// as MLTPathEngine does (mltpath.py:71-74): ray = Camera.generate(X0*2-1, X1*2-1); out[m][3]
int ora_trace_from_samples(void* h, const float* X, int m, int ndims, float* out) {
    Oracle* o = (Oracle*)h;
#pragma omp parallel for schedule(dynamic, 64)
    for (int q = 0; q < m; q++) {
        Oracle::Rng rng{nullptr, 0, 0, X + (size_t)q * ndims, 0};
        float a = rng.random(), b = rng.random();
        Ray r = o->camera_generate(a * 2 - 1, b * 2 - 1);
        V3 clr = o->path_trace(r, rng, nullptr);
        out[3 * q] = clr.x; out[3 * q + 1] = clr.y; out[3 * q + 2] = clr.z;
    }
    return 0;
}

int ora_get_film(void* h, int pass, float* out) {
    Oracle* o = (Oracle*)h;
    memcpy(out, o->film[pass].data(), sizeof(V4) * o->film[pass].size());
    return 0;
}
// filmtable.py:47-63
int ora_get_image(void* h, int pass, float* out) {
    Oracle* o = (Oracle*)h;
    for (size_t q = 0; q < o->film[pass].size(); q++) {
        V4 val = o->film[pass][q];
        if (val.w != 0) { val.x /= val.w; val.y /= val.w; val.z /= val.w; val.w = 1.0f; }
        else val = V4{0.9f, 0.4f, 0.9f, 0.0f};
        out[4 * q] = val.x; out[4 * q + 1] = val.y; out[4 * q + 2] = val.z; out[4 * q + 3] = val.w;
    }
    return 0;
}
// filmtable.py:65-79
int ora_fast_export_image(void* h, int pass, float* out) {
    Oracle* o = (Oracle*)h;
    for (int x = 0; x < o->nx; x++)
        for (int y = 0; y < o->ny; y++) {
            V4 val = o->film[pass][(size_t)x * o->ny + y];
            if (val.w != 0) { val.x /= val.w; val.y /= val.w; val.z /= val.w; }
            else { val.x = 0.9f; val.y = 0.4f; val.z = 0.9f; }
            size_t base = ((size_t)y * o->nx + x) * 3;
            out[base] = val.x; out[base + 1] = val.y; out[base + 2] = val.z;
        }
    return 0;
}

// BSDF taps.  params[m][14] = basecolor rgb, metallic, roughness, specular, specularTint, subsurface,
// sheen, sheenTint, clearcoat, clearcoatGloss, transmission, ior.  geom[m][10] = normal, sign, wi, wo|samp.
static Disney mk_disney(const float* p) {
    Disney m;
    m.basecolor = {p[0], p[1], p[2]}; m.metallic = p[3]; m.roughness = p[4]; m.specular = p[5]; m.specularTint = p[6];
    m.subsurface = p[7]; m.sheen = p[8]; m.sheenTint = p[9]; m.clearcoat = p[10]; m.clearcoatGloss = p[11];
    m.transmission = p[12]; m.ior = p[13];
    disney_init(m);
    return m;
}
int ora_eval_bsdf(const float* params, const float* geom, int m, float* out) {
    for (int q = 0; q < m; q++) {
        Disney d = mk_disney(params + 14 * q);
        const float* g = geom + 10 * q;
        V3 r = disney_brdf(d, {g[0], g[1], g[2]}, g[3], {g[4], g[5], g[6]}, {g[7], g[8], g[9]});
        out[3 * q] = r.x; out[3 * q + 1] = r.y; out[3 * q + 2] = r.z;
    }
    return 0;
}
// out[m][7] = outdir xyz, pdf, color rgb
int ora_sample_bsdf(const float* params, const float* geom, int m, float* out) {
    for (int q = 0; q < m; q++) {
        Disney d = mk_disney(params + 14 * q);
        const float* g = geom + 10 * q;
        BSDFSample s = disney_bounce(d, {g[0], g[1], g[2]}, g[3], {g[4], g[5], g[6]}, {g[7], g[8], g[9]});
        float* r = out + 7 * q;
        r[0] = s.outdir.x; r[1] = s.outdir.y; r[2] = s.outdir.z; r[3] = s.pdf; r[4] = s.color.x; r[5] = s.color.y; r[6] = s.color.z;
    }
    return 0;
}
// material fetch tap (mtllib.py:79-95 + image.py bilinear): out[m][14] raw parameters
int ora_material_get(void* h, const int32_t* mtlid, const float* uv, int m, float* out) {
    Oracle* o = (Oracle*)h;
    for (int q = 0; q < m; q++) {
        Disney d = o->material_get(mtlid[q], uv[2 * q], uv[2 * q + 1]);
        float* r = out + 14 * q;
        r[0] = d.basecolor.x; r[1] = d.basecolor.y; r[2] = d.basecolor.z; r[3] = d.metallic; r[4] = d.roughness; r[5] = d.specular;
        r[6] = d.specularTint; r[7] = d.subsurface; r[8] = d.sheen; r[9] = d.sheenTint; r[10] = d.clearcoat; r[11] = d.clearcoatGloss;
        r[12] = d.transmission; r[13] = d.ior;
    }
    return 0;
}
// light taps: rays[m][6] -> out[m][6] = hit, dis, pdf, color rgb   (light/__init__.py:51-81)
int ora_light_hit(void* h, const float* rays, int m, float* out) {
    Oracle* o = (Oracle*)h;
    for (int q = 0; q < m; q++) {
        Ray r{{rays[6 * q], rays[6 * q + 1], rays[6 * q + 2]}, {rays[6 * q + 3], rays[6 * q + 4], rays[6 * q + 5]}};
        Oracle::LitHit l = o->light_hit(r);
        float* w = out + 6 * q;
        w[0] = (float)l.hit; w[1] = l.dis; w[2] = l.pdf; w[3] = l.color.x; w[4] = l.color.y; w[5] = l.color.z;
    }
    return 0;
}
// in[m][6] = hitpos, samp -> out[m][8] = dis, dir xyz, pdf, color rgb   (light/__init__.py:83-121)
int ora_light_sample(void* h, const float* in, int m, float* out) {
    Oracle* o = (Oracle*)h;
    for (int q = 0; q < m; q++) {
        const float* g = in + 6 * q;
        Oracle::LitSample l = o->light_sample({g[0], g[1], g[2]}, {g[3], g[4], g[5]});
        float* w = out + 8 * q;
        w[0] = l.dis; w[1] = l.dir.x; w[2] = l.dir.y; w[3] = l.dir.z; w[4] = l.pdf; w[5] = l.color.x; w[6] = l.color.y; w[7] = l.color.z;
    }
    return 0;
}
// world light tap: dirs[m][3] -> out[m][3]   (light/world.py:22-29)
int ora_world_at(void* h, const float* dirs, int m, float* out) {
    Oracle* o = (Oracle*)h;
    for (int q = 0; q < m; q++) {
        V3 c = o->world_at({dirs[3 * q], dirs[3 * q + 1], dirs[3 * q + 2]});
        out[3 * q] = c.x; out[3 * q + 1] = c.y; out[3 * q + 2] = c.z;
    }
    return 0;
}
int ora_wanghash2(int x, int y) { return wanghash2(x, y); }
int ora_morton3d(float x, float y, float z) { return Oracle::morton3D({x, y, z}); }
int ora_clz(int x) { return Oracle::clz_ref(x); }
int ora_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
