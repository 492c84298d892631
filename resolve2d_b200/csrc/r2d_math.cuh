// r2d_math.cuh — f32 vector math, constants and the trig used by resolve2d's `rotate2`, usable from
// device code (nvcc, --fmad=false) and from host test harnesses (g++ -ffp-contract=off).
//
// Follows  src/core/nmath.zig:5-155, src/core/simulation_constants.zig:3-22  and, for sinf/cosf, the musl-derived
// routines Zig 0.14.1's compiler-rt links into the reference (lib/compiler_rt/{sin,cos,trig,rem_pio2f}.zig;
// constants recovered from the shipped wasm).  CUDA's sinf/cosf/__sinf and glibc's are
// different algorithms and must not be used here: bit-exact rectangle AABBs and SAT axes depend on this one.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define R2D_HD __host__ __device__ __forceinline__
#else
#define R2D_HD inline
#endif

namespace r2d {

// simulation_constants.zig:3-22
constexpr float MIN_MANIFOLD_IMPULSE = 1e-4f;
constexpr float BAUMGARTE = 0.02f;
constexpr float BAUMGARTE_SLOP = 0.005f;
constexpr float SAT_OVERLAP_THRESHOLD = 1e-4f;
constexpr float COLLISION_MARGIN = 0.01f;
constexpr float NMATH_WARN_DIVIDING_BELOW = 1e-3f;
constexpr float AABB_EPS_OVERLAP = 0.01f;
constexpr float CONSTRAINT_GRADIENT_DIVISION_LIMIT = 1e-4f;
constexpr float ALLOWED_CONSTRAINT_VALUE = 1e-6f;

struct v2 {
    float x, y;
};

R2D_HD v2 mk2(float x, float y) {
    v2 r;
    r.x = x;
    r.y = y;
    return r;
}
R2D_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
R2D_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
R2D_HD float inf32() { return u2f(0x7f800000u); }

// IEEE single-rounded primitives.  On the device the explicit round-to-nearest intrinsics can never be contracted
// into an FMA or replaced by an approximate sequence, whatever the compile flags.
R2D_HD float fsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
R2D_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
R2D_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
R2D_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
R2D_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
R2D_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
R2D_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
R2D_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}

// Zig @max/@min lower to fmaxf/fminf (NaN-ignoring); written out so host and device agree on signed zeros too
// (SURVEY A.8b).
R2D_HD float fmax_z(float x, float y) { return (x != x) ? y : ((y != y) ? x : (x < y ? y : x)); }
R2D_HD float fmin_z(float x, float y) { return (x != x) ? y : ((y != y) ? x : (x < y ? x : y)); }
// std.math.clamp(v, lo, hi) = @max(lo, @min(v, hi))
R2D_HD float clamp_z(float v, float lo, float hi) { return fmax_z(lo, fmin_z(v, hi)); }
R2D_HD float fabs_z(float x) { return u2f(f2u(x) & 0x7fffffffu); }

// nmath.zig:5-7 — strict open interval (Q23)
R2D_HD bool approx_eql(float a, float b, float eps) { return a > fsub(b, eps) && a < fadd(b, eps); }

// nmath.zig:54-155
R2D_HD v2 add2(v2 a, v2 b) { return mk2(fadd(a.x, b.x), fadd(a.y, b.y)); }
R2D_HD v2 sub2(v2 a, v2 b) { return mk2(fsub(a.x, b.x), fsub(a.y, b.y)); }
R2D_HD v2 scale2(v2 a, float s) { return mk2(fmul(a.x, s), fmul(a.y, s)); }
R2D_HD float dot2(v2 a, v2 b) { return fadd(fmul(a.x, b.x), fmul(a.y, b.y)); }
R2D_HD float cross2(v2 a, v2 b) { return fsub(fmul(a.x, b.y), fmul(a.y, b.x)); }
R2D_HD float length2sq(v2 a) { return dot2(a, a); }
R2D_HD float length2(v2 a) { return fsqrt(length2sq(a)); }
R2D_HD v2 normalize2(v2 a) {  // :95-103 (Q18)
    const float len = length2(a);
    if (len < NMATH_WARN_DIVIDING_BELOW) return mk2(0.0f, 0.0f);
    return scale2(a, fdiv(1.0f, len));
}
R2D_HD v2 negate2(v2 a) { return mk2(-a.x, -a.y); }                  // :105-107
R2D_HD v2 negate_mul(v2 a) { return mk2(fmul(a.x, -1.0f), fmul(a.y, -1.0f)); }  // Vector2.negate :48-51
R2D_HD v2 addmult2(v2 a, v2 b, float s) { return add2(a, scale2(b, s)); }
R2D_HD v2 submult2(v2 a, v2 b, float s) { return sub2(a, scale2(b, s)); }
R2D_HD bool approx_eql2(v2 a, v2 b, float eps) { return approx_eql(a.x, b.x, eps) && approx_eql(a.y, b.y, eps); }
R2D_HD v2 rot90cw(v2 a) { return mk2(a.y, -a.x); }    // :142-144
R2D_HD v2 rot90ccw(v2 a) { return mk2(-a.y, a.x); }   // :146-148
// rotate2 with the cos/sin of the angle already evaluated (nmath.zig:129-136)
R2D_HD v2 rotate_cs(v2 a, float c, float s) {
    return mk2(fsub(fmul(a.x, c), fmul(a.y, s)), fadd(fmul(a.x, s), fmul(a.y, c)));
}

// ---- trig (musl sinf/cosf as ported by Zig compiler-rt; f64 inside, no FMA) ---------------------------------
R2D_HD float k_cosdf(double x) {  // trig.zig __cosdf
    const double C0 = -0x1.ffffffd0c5e81p-2, C1 = 0x1.55553e1053a42p-5, C2 = -0x1.6c087e80f1e27p-10,
                 C3 = 0x1.99342e0ee5069p-16;
    const double z = dmul(x, x);
    const double w = dmul(z, z);
    const double r = dadd(C2, dmul(z, C3));
    return (float)dadd(dadd(dadd(1.0, dmul(z, C0)), dmul(w, C1)), dmul(dmul(w, z), r));
}
R2D_HD float k_sindf(double x) {  // trig.zig __sindf
    const double S1 = -0x1.5555554cbac77p-3, S2 = 0x1.11110896efbb2p-7, S3 = -0x1.a00f9e2cae774p-13,
                 S4 = 0x1.6cd878c3b46a7p-19;
    const double z = dmul(x, x);
    const double w = dmul(z, z);
    const double r = dadd(S3, dmul(z, S4));
    const double s = dmul(z, x);
    return (float)dadd(dadd(x, dmul(s, dadd(S1, dmul(z, S2)))), dmul(dmul(s, w), r));
}
// rem_pio2f.zig, medium-size branch (|x| < 2^28 * pi/2).  Larger arguments need rem_pio2_large, which no scene can
// reach (|angle| > 4.2e8 rad); they yield NaN here — a documented restriction.
R2D_HD int rem_pio2f(float x, double* y) {
    const double toint = 0x1.8p52, pio4 = 0x1.921fb6p-1, invpio2 = 0x1.45f306dc9c883p-1,
                 pio2_1 = 0x1.921fb50000000p+0, pio2_1t = 0x1.110b4611a6263p-26;
    const double xd = (double)x;
    double fn = dsub(dadd(dmul(xd, invpio2), toint), toint);
    int n = (int)fn;
    *y = dsub(dsub(xd, dmul(fn, pio2_1)), dmul(fn, pio2_1t));
    if (*y < -pio4) {
        n -= 1;
        fn = dsub(fn, 1.0);
        *y = dsub(dsub(xd, dmul(fn, pio2_1)), dmul(fn, pio2_1t));
    } else if (*y > pio4) {
        n += 1;
        fn = dadd(fn, 1.0);
        *y = dsub(dsub(xd, dmul(fn, pio2_1)), dmul(fn, pio2_1t));
    }
    return n;
}
R2D_HD float sin_ref(float x) {  // sin.zig sinf
    const double s1pio2 = 0x1.921fb54442d18p+0, s2pio2 = 0x1.921fb54442d18p+1, s3pio2 = 0x1.2d97c7f3321d2p+2,
                 s4pio2 = 0x1.921fb54442d18p+2;
    uint32_t ix = f2u(x);
    const bool sign = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    const double xd = (double)x;
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) return x;
        return k_sindf(xd);
    }
    if (ix <= 0x407b53d1u) {
        if (ix <= 0x4016cbe3u) {
            if (sign) return -k_cosdf(dadd(xd, s1pio2));
            return k_cosdf(dsub(xd, s1pio2));
        }
        return k_sindf(sign ? -dadd(xd, s2pio2) : -dsub(xd, s2pio2));
    }
    if (ix <= 0x40e231d5u) {
        if (ix <= 0x40afeddfu) {
            if (sign) return k_cosdf(dadd(xd, s3pio2));
            return -k_cosdf(dsub(xd, s3pio2));
        }
        return k_sindf(sign ? dadd(xd, s4pio2) : dsub(xd, s4pio2));
    }
    if (ix >= 0x4dc90fdbu) return fsub(x, x) + u2f(0x7fc00000u);  // inf/NaN, and the unsupported huge range
    double y;
    const int n = rem_pio2f(x, &y);
    switch (n & 3) {
        case 0: return k_sindf(y);
        case 1: return k_cosdf(y);
        case 2: return k_sindf(-y);
        default: return -k_cosdf(y);
    }
}
R2D_HD float cos_ref(float x) {  // cos.zig cosf
    const double c1pio2 = 0x1.921fb54442d18p+0, c2pio2 = 0x1.921fb54442d18p+1, c3pio2 = 0x1.2d97c7f3321d2p+2,
                 c4pio2 = 0x1.921fb54442d18p+2;
    uint32_t ix = f2u(x);
    const bool sign = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    const double xd = (double)x;
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) return 1.0f;
        return k_cosdf(xd);
    }
    if (ix <= 0x407b53d1u) {
        if (ix > 0x4016cbe3u) return -k_cosdf(sign ? dadd(xd, c2pio2) : dsub(xd, c2pio2));
        if (sign) return k_sindf(dadd(xd, c1pio2));
        return k_sindf(dsub(c1pio2, xd));
    }
    if (ix <= 0x40e231d5u) {
        if (ix > 0x40afeddfu) return k_cosdf(sign ? dadd(xd, c4pio2) : dsub(xd, c4pio2));
        if (sign) return k_sindf(dsub(-xd, c3pio2));
        return k_sindf(dsub(xd, c3pio2));
    }
    if (ix >= 0x4dc90fdbu) return fsub(x, x) + u2f(0x7fc00000u);
    double y;
    const int n = rem_pio2f(x, &y);
    switch (n & 3) {
        case 0: return k_cosdf(y);
        case 1: return k_sindf(-y);
        case 2: return -k_cosdf(y);
        default: return k_sindf(y);
    }
}

// sin_ref(x) and cos_ref(x) together.  In every argument range the two functions evaluate __sindf and __cosdf on the SAME
// reduced argument up to its sign (x -+ k pi/2; IEEE subtraction is antisymmetric, __cosdf is even and __sindf odd bit for
// bit: only z = x*x and sign-symmetric sums enter), so ONE polynomial of each kind serves both — half the arithmetic, and a
// warp whose lanes fall in different ranges runs two polynomial bodies instead of the sixteen inlined copies of the
// separate functions.  Checked against sin_ref / cos_ref on all 2^32 float bit patterns (tests/test_trig_exhaustive.py).
R2D_HD void sincos_ref(float x, float* s_out, float* c_out) {
    const double pio2 = 0x1.921fb54442d18p+0, pi = 0x1.921fb54442d18p+1, pio2_3 = 0x1.2d97c7f3321d2p+2, pi2 = 0x1.921fb54442d18p+2;
    uint32_t ix = f2u(x);
    const bool sign = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    const double xd = (double)x;
    if (ix >= 0x4dc90fdbu) {   // inf / NaN, and the unsupported huge range
        *s_out = *c_out = fsub(x, x) + u2f(0x7fc00000u);
        return;
    }
    double a;       // the common argument
    int form;       // 0: s = S, c = C   1: s = C, c = S   2: s = -C, c = S   3: s = S, c = -C   4: s = C, c = -S   5: s = -S, c = -C   6: s = -C, c = -S... (see below)
    if (ix <= 0x3f490fdau) {
        if (ix < 0x39800000u) {
            *s_out = x;
            *c_out = 1.0f;
            return;
        }
        a = xd; form = 0;                                    // sin = S(x),            cos = C(x)
    } else if (ix <= 0x407b53d1u) {
        if (ix <= 0x4016cbe3u) {
            if (sign) { a = dadd(xd, pio2); form = 2; }      // sin = -C(x + pi/2),     cos = S(x + pi/2)
            else      { a = dsub(pio2, xd); form = 1; }      // sin = C(x - pi/2) = C(a), cos = S(pi/2 - x) = S(a)
        } else {
            a = sign ? -dadd(xd, pi) : -dsub(xd, pi); form = 3;   // sin = S(-(x -+ pi)) = S(a), cos = -C(x -+ pi) = -C(a)
        }
    } else if (ix <= 0x40e231d5u) {
        if (ix <= 0x40afeddfu) {
            if (sign) { a = dsub(-xd, pio2_3); form = 1; }   // sin = C(x + 3pi/2) = C(a), cos = S(-x - 3pi/2) = S(a)
            else      { a = dsub(xd, pio2_3); form = 2; }    // sin = -C(x - 3pi/2),    cos = S(x - 3pi/2)
        } else {
            a = sign ? dadd(xd, pi2) : dsub(xd, pi2); form = 0;
        }
    } else {
        double y;
        const int n = rem_pio2f(x, &y);
        a = y;
        switch (n & 3) {
            case 0: form = 0; break;                         // sin = S(y),  cos = C(y)
            case 1: form = 4; break;                         // sin = C(y),  cos = S(-y) = -S(y)
            case 2: form = 5; break;                         // sin = S(-y) = -S(y), cos = -C(y)
            default: form = 2; break;                        // sin = -C(y), cos = S(y)
        }
    }
    const float S = k_sindf(a), C = k_cosdf(a);
    switch (form) {
        case 0: *s_out = S; *c_out = C; break;
        case 1: *s_out = C; *c_out = S; break;
        case 2: *s_out = -C; *c_out = S; break;
        case 3: *s_out = S; *c_out = -C; break;
        case 4: *s_out = C; *c_out = -S; break;
        default: *s_out = -S; *c_out = -C; break;
    }
}

// SpatialHash.hash (SpatialHash.zig:78-81): u64(xi*92837111 ^ yi*689287499) % table_size  (`*` binds tighter than `^`)
R2D_HD uint64_t cell_hash(int64_t xi, int64_t yi, uint64_t table_size) {
    const uint64_t h = ((uint64_t)xi * 92837111ull) ^ ((uint64_t)yi * 689287499ull);  // two's-complement wrap == i64 product
    return h % table_size;
}
// The same value without the 64-bit division (a ~150-instruction subroutine on the GPU, four per body: 40 % of the grid
// kernels' instructions): `magic` = floor((2^64 - 1) / table_size), one per world, from the host.  q = mulhi(h, magic) is
// floor(h / table_size) or up to 2 less, so h - q * table_size needs at most two corrections.  Exact for every h.
R2D_HD uint64_t hash_magic(uint64_t table_size) { return table_size ? ~0ull / table_size : 0ull; }
R2D_HD uint64_t cell_hash_magic(int64_t xi, int64_t yi, uint64_t table_size, uint64_t magic) {
    const uint64_t h = ((uint64_t)xi * 92837111ull) ^ ((uint64_t)yi * 689287499ull);
#if defined(__CUDA_ARCH__)
    const uint64_t q = __umul64hi(h, magic);
#else
    const uint64_t q = (uint64_t)(((unsigned __int128)h * magic) >> 64);
#endif
    uint64_t r = h - q * table_size;
    if (r >= table_size) r -= table_size;
    if (r >= table_size) r -= table_size;
    return r;
}
// @intFromFloat(@floor(v / cell_size)) (SpatialHash.zig:93-96)
R2D_HD int64_t cell_coord(float v, float cell_size) { return (int64_t)floorf(fdiv(v, cell_size)); }

}  // namespace r2d
