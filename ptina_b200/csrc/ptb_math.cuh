// ptb_math.cuh -- f32 vector helpers with the reference's evaluation order.
//
// The whole library is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so the
// expressions below are IEEE binary32 operations in source order (no FMA contraction).  That is what
// makes Morton codes, primary rays and the box / triangle predicates bit-identical to a strict-IEEE CPU
// evaluation of the reference (SURVEY.md 7 "Bit-exactness").  Conventions follow Taichi 0.7:
// dot = left-to-right sum, normalized() = (1/|v|) * v, Matrix @ accumulates k ascending
// (reference: ptina/common.py:32-33, 73-77, 164-180, 213-271).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PTB_EPS 1e-6f
#define PTB_INF 1e6f
#define PTB_PI 3.14159265358979323846f
#define PTB_TAU 6.28318530717958647692f
#define PTB_HD __host__ __device__ __forceinline__
#define PTB_D __device__ __forceinline__

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

PTB_HD V3 mk3(float a, float b, float c) { V3 r; r.x = a; r.y = b; r.z = c; return r; }
PTB_HD V3 v3s(float a) { return mk3(a, a, a); }
PTB_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PTB_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PTB_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PTB_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PTB_HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PTB_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PTB_HD V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PTB_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PTB_HD V3 operator-(float s, V3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
PTB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PTB_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PTB_HD float norm_sqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PTB_HD float norm(V3 a) { return sqrtf(norm_sqr(a)); }
PTB_HD V3 normalized(V3 a) { float inv = 1.0f / norm(a); return inv * a; }
PTB_HD V3 vmin(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PTB_HD V3 vmax(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PTB_HD float vavg(V3 a) { return (a.x + a.y + a.z) / 3.0f; }
PTB_HD bool any_gt(V3 a, float s) { return a.x > s || a.y > s || a.z > s; }
PTB_HD bool any_ne0(V3 a) { return a.x != 0.0f || a.y != 0.0f || a.z != 0.0f; }

PTB_HD V4 mk4(float a, float b, float c, float d) { V4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
PTB_HD V4 operator*(V4 a, V4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PTB_HD V4 operator*(V4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
PTB_HD V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

PTB_HD float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
PTB_HD int clampi(int x, int lo, int hi) { return min(hi, max(lo, x)); }
PTB_HD float dot_or_zero(V3 a, V3 b) { return fmaxf(0.0f, dot(a, b)); }
PTB_HD float lerpf(float f, float a, float b) { return a * (1.0f - f) + b * f; }
PTB_HD V3 lerp3(float f, V3 a, V3 b) { return a * (1.0f - f) + b * f; }
PTB_HD int pymod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// ---- exact integer division / modulo by a divisor fixed per launch (a runtime `/` or `%` is ~25-35 instructions; these are 2-10) ----
// FastDiv: floor(n / d) for 0 <= n < 2^31.  d a power of two: a shift.  Otherwise, with L = ceil(log2 d) and M = floor(2^(31+L) / d) + 1
// (fits 32 bits), M*d = 2^(31+L) + e with 0 < e <= d, so n*M / 2^(31+L) = n/d + n*e / (d * 2^(31+L)) and the excess is below 2^-L < 1/d:
// it never reaches the next integer.  tests/test_fastdiv_cpu.py checks both against Python's // and % on boundary and random inputs.
struct FastDiv { unsigned d, M; int sh; };
inline __host__ __device__ FastDiv make_fastdiv(unsigned d) {
    FastDiv f; f.d = d; f.M = 0; f.sh = 0;
    int L = 0; while ((1ull << L) < d) L++;
    if ((d & (d - 1)) == 0) f.sh = L;
    else { f.M = (unsigned)(((1ull << (31 + L)) / d) + 1ull); f.sh = L - 1; }
    return f;
}
PTB_D unsigned fastdiv(unsigned n, const FastDiv& f) { return f.M ? (__umulhi(n, f.M) >> f.sh) : (n >> f.sh); }
// FastMod: Python-style i mod m for ANY int32 i (the Sobol dimension rotation wraps through the whole i32 range, sobol.py:107-125).
// q = floor(n / m) for the unsigned n = i mod 2^32 via M = floor(2^64 / m) + 1 (excess below 2^-32 < 1/m); a negative i is n - 2^32, so
// its residue is (n mod m) - (2^32 mod m), brought back into [0, m).
struct FastMod { unsigned m, K; unsigned long long M; };
inline __host__ __device__ FastMod make_fastmod(unsigned m) {
    FastMod f; f.m = m; f.K = 0; f.M = 0;
    if (m > 1) { f.M = 0xFFFFFFFFFFFFFFFFull / m + 1ull; f.K = (unsigned)((1ull << 32) % m); }      // m == 1: M stays 0 and every residue is 0
    return f;
}
PTB_D int fastpymod(int i, const FastMod& f) {
    if (f.M == 0ull) return 0;
    const unsigned n = (unsigned)i;
    const unsigned q = (unsigned)__umul64hi((unsigned long long)n, f.M);
    int r = (int)(n - q * f.m);
    if (i < 0) { r -= (int)f.K; if (r < 0) r += (int)f.m; }
    return r;
}

// int(ti.floor(x)): NaN / out-of-range -> INT_MIN (x86 cvttss2si), which every call site then clamps
PTB_D int ifloor(float x) {
    float f = floorf(x);
    if (!(f == f) || f >= 2147483648.0f || f < -2147483648.0f) return INT32_MIN;
    return (int)f;
}

struct M33 { float m[3][3]; };
PTB_HD V3 matvec(const M33& a, V3 v) {
    return mk3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
               a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
               a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}

// common.py:213-217
PTB_D M33 tanspace(V3 nrm) {
    V3 bitan = normalized(cross(nrm, mk3(233.f, 666.f, 512.f)));
    V3 tan = cross(bitan, nrm);
    M33 r;
    r.m[0][0] = tan.x; r.m[0][1] = bitan.x; r.m[0][2] = nrm.x;
    r.m[1][0] = tan.y; r.m[1][1] = bitan.y; r.m[1][2] = nrm.y;
    r.m[2][0] = tan.z; r.m[2][1] = bitan.z; r.m[2][2] = nrm.z;
    return r;
}
// common.py:221-225
PTB_D V3 spherical(float h, float p) {
    float s, c;
    sincosf(p * PTB_TAU, &s, &c);
    float r = sqrtf(fmaxf(0.0f, 1.0f - h * h));
    return mk3(r * c, r * s, h);
}
// common.py:247-249
PTB_D V3 reflect(V3 I, V3 N) { return I - (2.0f * dot(N, I)) * N; }
// common.py:252-260
PTB_D int refract(V3 I, V3 N, float eta, V3* T) {
    *T = I * 0.0f;
    float NoI = dot(N, I);
    float discr = 1.0f - (eta * eta) * (1.0f - NoI * NoI);
    if (discr > 0.0f) {
        *T = normalized(eta * I - N * (eta * NoI + sqrtf(discr)));
        return 1;
    }
    return 0;
}

// sampling/__init__.py:8-23 (u32 arithmetic; `value << 4` is a LEFT shift in the reference)
PTB_HD int32_t wanghash(int32_t x) {
    uint32_t v = (uint32_t)x;
    v = (v ^ 61u) ^ (v >> 16);
    v *= 9u;
    v ^= v << 4;
    v *= 0x27d4eb2du;
    v ^= v >> 15;
    return (int32_t)v;
}
PTB_HD int32_t wanghash2(int32_t x, int32_t y) { return wanghash(y ^ wanghash(x)); }
