// ag_device.cuh -- device-side building blocks of the scene_0 hot path (sm_100a).
//
// Everything that decides a flag is float64 with ONE rounding per operation (__dadd_rn /
// __dmul_rn / __ddiv_rn are never contracted into FMAs), in the reference's operation order.
// Reference citations are file:line relative to the reference root.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abstract_gym_b200.h"

namespace agd {

// ---------------------------------------------------------------- kernel-side grid descriptor
struct GridDev {
    const uint32_t *bits;
    const uint32_t *bits_t;   // transposed copy (lines = columns) or nullptr
    const unsigned char *hier;   // two-level form (ag_grid.hier: 8x8 tiles + summary bitmap) or nullptr
    const double *min_x;
    const double *min_y;
    double side;       // E/(S-1)
    double half;       // E/2
    double inv_side;   // 1/side (traversal only, never a decision quantity)
    double margin;     // conservative widening of the traversal, metres
    int32_t S;
    int32_t wpr;
    int32_t n_grids;
    int32_t stage;     // 1: block-uniform grid, stage bits+tables into shared memory
    int64_t stride_words;
    int64_t envs_per_grid;
    int32_t T, cwpr;          // hier: tiles per side, summary words per row
    int32_t hier_tiles_bytes; // hier: bytes of the tile array (the two summary bitmaps follow it)
    int32_t hier_coarse_words;// hier: words of one summary bitmap (row-major, then column-major)
    int32_t hier_bytes;       // hier: bytes per grid
};

// what one thread sees: either shared-memory copies or global pointers
struct GridView {
    const uint32_t *bits;
    const uint32_t *bits_t;   // nullptr when the grid has no transposed copy
    const double *min_x;
    const double *min_y;
    const unsigned long long *tiles;   // two-level form: 8x8 tiles, T*T of them; nullptr = none
    const uint32_t *coarse;            // ... and the T x T summary bitmap (lines = tile rows)
    const uint32_t *coarse_t;          // ... transposed (lines = tile columns)
};
__device__ __forceinline__ void view_hier(GridView &V, const GridDev &G, const unsigned char *h) {
    V.tiles = reinterpret_cast<const unsigned long long *>(h);
    V.coarse = h ? reinterpret_cast<const uint32_t *>(h + G.hier_tiles_bytes) : nullptr;
    V.coarse_t = h ? V.coarse + G.hier_coarse_words : nullptr;
}

__device__ __forceinline__ int64_t grid_of_env(const GridDev &G, int64_t gid) {
    return G.n_grids == 1 ? 0 : (int64_t)(((uint64_t)gid / (uint64_t)G.envs_per_grid) % (uint64_t)G.n_grids);
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. SC'11)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms in [0,1) with 53 random bits each: (a>>5, b>>6) -> (a*2^26+b)/2^53, the bit
// recipe of np.random.rand() (scenario/scene_0.py:84-85,180-181 draw one double at a time).
// Every operation below is exact in float64.
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint64_t gid, uint32_t draw, uint32_t stream,
                                                double &u0, double &u1) {
    uint32_t w[4];
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), draw, stream, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    u0 = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) * (1.0 / 9007199254740992.0);
    u1 = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------- FK  robot/two_joint_robot.py:31-47
struct Arm { double ex, ey, gx, gy; };   // elbow, end effector

__device__ __forceinline__ Arm forward_kinematics(double j1, double j2, double l1, double l2) {
    double s1, c1, s2, c2;
    sincos(j1, &s1, &c1);
    sincos(j2, &s2, &c2);
    Arm a;
    a.ex = __dmul_rn(c1, l1);                          // :45
    a.ey = __dmul_rn(s1, l1);                          // :46
    a.gx = __dadd_rn(a.ex, __dmul_rn(c2, l2));         // :36  (cos(j1)*l1) + (cos(j2)*l2)
    a.gy = __dadd_rn(a.ey, __dmul_rn(s2, l2));         // :37
    return a;
}

// ---------------------------------------------------------------- line  utils/geometry.py:14-32
struct LineD { double a, b, c, p0x, p0y, p1x, p1y; };

__device__ __forceinline__ LineD make_line(double p0x, double p0y, double p1x, double p1y) {
    LineD L;
    L.p0x = p0x; L.p0y = p0y; L.p1x = p1x; L.p1y = p1y;
    if (p0x == p1x) {                                  // :19-23
        L.a = 1.0; L.b = 0.0; L.c = -p0x;
    } else if (p0y == p1y) {                           // :24-28
        L.a = 0.0; L.b = 1.0; L.c = -p0y;
    } else {
        const double dx = __dsub_rn(p1x, p0x), dy = __dsub_rn(p1y, p0y);
        L.a = __ddiv_rn(1.0, dx);                      // :29
        L.b = __ddiv_rn(-1.0, dy);                     // :30
        L.c = __dsub_rn(__ddiv_rn(p0y, dy), __ddiv_rn(p0x, dx));   // :31
    }
    return L;
}

// ---------------------------------------------------------------- predicate  utils/collision_checker.py:23-85
// Returns the reference's CollisionChecker(line, square).collision_check().  `axis` counts the
// a==0 / b==0 branch of check_sections (:59-68), which raises AttributeError in the reference and
// is defined here as the interval-overlap test that code evidently intends.
__device__ __forceinline__ bool segment_square_exact(const LineD &L, double min_x, double min_y, double max_x,
                                                     double max_y, double eps, int &axis) {
    const double ax0 = __dmul_rn(L.a, min_x), ax1 = __dmul_rn(L.a, max_x);
    const double by0 = __dmul_rn(L.b, min_y), by1 = __dmul_rn(L.b, max_y);
    const double v1 = __dadd_rn(__dadd_rn(ax0, by0), L.c);     // :27
    const double v2 = __dadd_rn(__dadd_rn(ax0, by1), L.c);     // :28
    const double v3 = __dadd_rn(__dadd_rn(ax1, by0), L.c);     // :29
    const double v4 = __dadd_rn(__dadd_rn(ax1, by1), L.c);     // :30
    const bool pos = (v1 > 0.0) | (v2 > 0.0) | (v3 > 0.0) | (v4 > 0.0);   // :41
    const bool neg = (v1 < 0.0) | (v2 < 0.0) | (v3 < 0.0) | (v4 < 0.0);   // :42
    if (!(pos && neg)) return false;                             // :43-46
    if (L.a == 0.0) {                                            // :59-63
        ++axis;
        return !(fmax(L.p0x, L.p1x) < min_x || fmin(L.p0x, L.p1x) > max_x);
    }
    if (L.b == 0.0) {                                            // :64-68
        ++axis;
        return !(fmax(L.p0y, L.p1y) < min_y || fmin(L.p0y, L.p1y) > max_y);
    }
    const double nc = -L.c;
    const double x1 = __ddiv_rn(__dsub_rn(nc, by0), L.a);        // :69  (-c - b*min_y)/a
    const double x2 = __ddiv_rn(__dsub_rn(nc, by1), L.a);        // :71
    // :74,:76-78: sort {x1, x2, min_x, max_x}; the middle two are max(mins), min(maxes)
    const double second = fmax(fmin(x1, x2), min_x);
    const double third = fmin(fmax(x1, x2), max_x);
    const double den = __dsub_rn(L.p1x, L.p0x);
    const double lam1 = __ddiv_rn(__dsub_rn(second, L.p0x), den);   // :79,:90-91
    const double lam2 = __ddiv_rn(__dsub_rn(third, L.p0x), den);    // :80
    return (1.0 > lam1 && lam1 > eps) || (1.0 > lam2 && lam2 > eps);   // :82
}

// ---------------------------------------------------------------- debug build: index assertions
// `python -m abstract_gym_b200.build --debug` compiles libabstract_gym_b200_debug.so with -DAG_DEBUG_BOUNDS: every
// computed index into the bit grid, the cell-corner arrays, the record planes, the action rows and the event sink is
// checked on the device and a violation traps the kernel (the launch then fails with a CUDA error, which the host entry
// reports).  The release build compiles the checks away.
#ifdef AG_DEBUG_BOUNDS
#define AG_CHECK_INDEX(i, n)                                                                                          \
    do {                                                                                                              \
        if (!((long long)(i) >= 0 && (long long)(i) < (long long)(n))) {                                              \
            printf("AG_DEBUG_BOUNDS %s:%d: index %lld outside [0, %lld)\n", __FILE__, __LINE__, (long long)(i),        \
                   (long long)(n));                                                                                   \
            __trap();                                                                                                 \
        }                                                                                                             \
    } while (0)
#else
#define AG_CHECK_INDEX(i, n) ((void)0)
#endif

// ---------------------------------------------------------------- cell index helpers
// Row r spans y in [E/2 - r*s, E/2 - (r-1)*s]; column c spans x in [c*s - E/2, (c+1)*s - E/2]
// (environment/occupancy_grid.py:59-67: S cells of side E/(S-1), y flipped, not centred).
__device__ __forceinline__ int row_of(const GridDev &G, double y) {
    const double f = floor((G.half - y) * G.inv_side) + 1.0;
    return (int)fmin(fmax(f, -1.0), (double)G.S);
}
__device__ __forceinline__ int col_of(const GridDev &G, double x) {
    const double f = floor((x + G.half) * G.inv_side);
    return (int)fmin(fmax(f, -1.0), (double)G.S);
}

// ---------------------------------------------------------------- one link against every occupied cell
// scenario/scene_0.py:67-75 for one link: the reference's own loop.  Used by link_exact for nearly axis-aligned links.
template <bool WANT_FIRST>
__device__ __noinline__ bool link_brute(const GridDev &G, const GridView &V, double p0x, double p0y, double p1x, double p1y,
                                        double eps, int &fh, int &axis) {
    const LineD L = make_line(p0x, p0y, p1x, p1y);
    bool hit = false;
    for (int r = 0; r < G.S; ++r)
        for (int w = 0; w < G.wpr; ++w) {
            uint32_t word = V.bits[r * G.wpr + w];
            while (word) {
                const int c = (w << 5) + __ffs(word) - 1;
                word &= word - 1;
                AG_CHECK_INDEX(c, G.S);      // padding bits of the last word of a row must be clear
                const double mnx = V.min_x[c], mny = V.min_y[r];
                if (segment_square_exact(L, mnx, mny, __dadd_rn(mnx, G.side), __dadd_rn(mny, G.side), eps, axis)) {
                    if (!WANT_FIRST) return true;
                    hit = true;
                    fh = min(fh, r * G.S + c);
                }
            }
        }
    return hit;
}

// ---------------------------------------------------------------- EXACT engine: one link
// Conservative traversal: every cell whose closed square comes within G.margin of the segment
// is visited (a superset of the cells the reference can flag, because a hit needs a point of
// the square's boundary on the segment); the reference predicate runs on occupied ones only.
// Returns hit; with WANT_FIRST it visits everything and keeps min(row*S+col) in fh.
template <bool WANT_FIRST>
__device__ __forceinline__ bool link_exact(const GridDev &G, const GridView &V, double p0x, double p0y, double p1x,
                                           double p1y, double eps, int &fh, int &axis) {
    // A nearly axis-aligned link (|dx| or |dy| < 1e-7 m): the reference's lam = (x - p0.x) / (p1.x - p0.x)
    // (collision_checker.py:79-80,90-91) then carries rounding noise of ~ulp/|dx|, which can move its accepted range far
    // beyond the segment's end, i.e. outside any bounded neighbourhood of the link.  Probability ~1e-7 per link and
    // step: take the reference's own loop over every occupied cell.
    if (fmin(fabs(p1x - p0x), fabs(p1y - p0y)) < 1.0e-7) return link_brute<WANT_FIRST>(G, V, p0x, p0y, p1x, p1y, eps, fh, axis);
    const double m = G.margin;
    int r_lo = row_of(G, fmax(p0y, p1y) + m), r_hi = row_of(G, fmin(p0y, p1y) - m);
    if (r_lo > G.S - 1 || r_hi < 0) return false;
    r_lo = max(r_lo, 0); r_hi = min(r_hi, G.S - 1);
    const double dx = p1x - p0x, dy = p1y - p0y;
    const bool clip = fabs(dy) > 1e-12;
    const double inv_dy = clip ? 1.0 / dy : 0.0;
    bool hit = false, have_line = false;
    LineD L;
    for (int r = r_lo; r <= r_hi; ++r) {
        double xa = p0x, xb = p1x;
        if (clip) {
            const double yb = G.half - (double)r * G.side - m, yt = G.half - (double)(r - 1) * G.side + m;
            const double t0 = (yb - p0y) * inv_dy, t1 = (yt - p0y) * inv_dy;
            const double ta = fmin(fmax(fmin(t0, t1), 0.0), 1.0), tb = fmin(fmax(fmax(t0, t1), 0.0), 1.0);
            xa = p0x + ta * dx; xb = p0x + tb * dx;
        }
        int c_lo = col_of(G, fmin(xa, xb) - m), c_hi = col_of(G, fmax(xa, xb) + m);
        if (c_lo > G.S - 1 || c_hi < 0) continue;
        c_lo = max(c_lo, 0); c_hi = min(c_hi, G.S - 1);
        for (int w = c_lo >> 5; w <= (c_hi >> 5); ++w) {
            uint32_t mask = 0xFFFFFFFFu;
            if (w == (c_lo >> 5)) mask &= 0xFFFFFFFFu << (c_lo & 31);
            if (w == (c_hi >> 5)) mask &= 0xFFFFFFFFu >> (31 - (c_hi & 31));
            AG_CHECK_INDEX(r * G.wpr + w, G.S * G.wpr);
            uint32_t word = V.bits[r * G.wpr + w] & mask;
            while (word) {
                const int c = (w << 5) + __ffs(word) - 1;
                word &= word - 1;
                AG_CHECK_INDEX(c, G.S);
                if (!have_line) { L = make_line(p0x, p0y, p1x, p1y); have_line = true; }
                const double mnx = V.min_x[c], mny = V.min_y[r];
                if (segment_square_exact(L, mnx, mny, __dadd_rn(mnx, G.side), __dadd_rn(mny, G.side), eps, axis)) {
                    if (!WANT_FIRST) return true;
                    hit = true;
                    fh = min(fh, r * G.S + c);
                }
            }
        }
    }
    return hit;
}

// ---------------------------------------------------------------- BRUTE engine: every occupied cell
// scenario/scene_0.py:67-75 as written: loop over all obstacles, both links each.
template <bool WANT_FIRST>
__device__ __forceinline__ bool arm_brute(const GridDev &G, const GridView &V, const Arm &A, double eps, int &fh,
                                          int &axis) {
    const LineD L1 = make_line(0.0, 0.0, A.ex, A.ey);          // scene_0.py:65
    const LineD L2 = make_line(A.ex, A.ey, A.gx, A.gy);        // scene_0.py:66
    bool hit = false;
    for (int r = 0; r < G.S; ++r)
        for (int w = 0; w < G.wpr; ++w) {
            uint32_t word = V.bits[r * G.wpr + w];
            while (word) {
                const int c = (w << 5) + __ffs(word) - 1;
                word &= word - 1;
                AG_CHECK_INDEX(c, G.S);      // padding bits of the last word of a row must be clear
                const double mnx = V.min_x[c], mny = V.min_y[r];
                const double mxx = __dadd_rn(mnx, G.side), mxy = __dadd_rn(mny, G.side);
                if (segment_square_exact(L1, mnx, mny, mxx, mxy, eps, axis) ||
                    segment_square_exact(L2, mnx, mny, mxx, mxy, eps, axis)) {
                    if (!WANT_FIRST) return true;
                    hit = true;
                    fh = min(fh, r * G.S + c);
                }
            }
        }
    return hit;
}

template <int ENGINE, bool WANT_FIRST>
__device__ __forceinline__ bool arm_collides(const GridDev &G, const GridView &V, const Arm &A, double eps, int &fh,
                                             int &axis) {
    if (ENGINE == AG_ENGINE_BRUTE) return arm_brute<WANT_FIRST>(G, V, A, eps, fh, axis);
    // reconverge between the two lane-dependent traversals (see arm_fast in ag_fast.cuh)
    const unsigned lanes = __activemask();
    bool hit = link_exact<WANT_FIRST>(G, V, 0.0, 0.0, A.ex, A.ey, eps, fh, axis);
    __syncwarp(lanes);
    if (!hit || WANT_FIRST) hit |= link_exact<WANT_FIRST>(G, V, A.ex, A.ey, A.gx, A.gy, eps, fh, axis);
    __syncwarp(lanes);
    return hit;
}

// ---------------------------------------------------------------- reach test  scenario/scene_0.py:115-133
__device__ __forceinline__ bool target_reached_joint(const ag_params &P, double j1, double j2) {   // :123-127
    return fabs(__dsub_rn(j1, P.target_j1)) < P.reach_eps && fabs(__dsub_rn(j2, P.target_j2)) < P.reach_eps;
}
__device__ __forceinline__ bool target_reached_cart(const ag_params &P, const Arm &A) {            // :129-130
    return fabs(__dsub_rn(P.target_x, A.gx)) < P.reach_eps && fabs(__dsub_rn(P.target_y, A.gy)) < P.reach_eps;
}
__device__ __forceinline__ bool target_reached(const ag_params &P, double j1, double j2, const Arm &A) {
    return P.choose_j_tar ? target_reached_joint(P, j1, j2) : target_reached_cart(P, A);
}
// tgt: per-env override of Scene.target_c (double[2]) or nullptr
__device__ __forceinline__ bool target_reached_at(const ag_params &P, double j1, double j2, const Arm &A, const double *tgt) {
    if (P.choose_j_tar) return target_reached_joint(P, j1, j2);
    if (tgt == nullptr) return target_reached_cart(P, A);
    const double2 t = *reinterpret_cast<const double2 *>(tgt);
    return fabs(__dsub_rn(t.x, A.gx)) < P.reach_eps && fabs(__dsub_rn(t.y, A.gy)) < P.reach_eps;   // :129-130
}

// ---------------------------------------------------------------- shared-memory staging of the grid
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage this block's grid (bits + corner tables) into dynamic shared memory.  Layout:
//   [mbarrier 16 B][bits: stride_words*4 B][bits_t: the same, if present][hier: hier_bytes, if present][min_x: Spad*8][min_y: Spad*8]
__device__ __forceinline__ GridView stage_grid(const GridDev &G, int64_t block_gid0, unsigned char *smem) {
    GridView V;
    const int64_t g = grid_of_env(G, block_gid0);
    const uint32_t *gbits = G.bits + g * G.stride_words;
    const uint32_t *gbits_t = G.bits_t ? G.bits_t + g * G.stride_words : nullptr;
    const unsigned char *ghier = G.hier ? G.hier + g * (int64_t)G.hier_bytes : nullptr;
    if (!G.stage) {
        V.bits = gbits; V.bits_t = gbits_t; V.min_x = G.min_x; V.min_y = G.min_y;
        view_hier(V, G, ghier);
        return V;
    }
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *sbits = reinterpret_cast<uint32_t *>(smem + 16);
    const uint32_t bit_bytes = (uint32_t)G.stride_words * 4u;
    const uint32_t two_bytes = G.bits_t ? 2u * bit_bytes : bit_bytes;
    const uint32_t hier_bytes = G.hier ? (uint32_t)G.hier_bytes : 0u;
    const uint32_t all_bytes = two_bytes + hier_bytes;
    unsigned char *shier = smem + 16 + two_bytes;
    const int spad = (G.S + 1) & ~1;
    double *sx = reinterpret_cast<double *>(smem + 16 + all_bytes);
    double *sy = sx + spad;
    if (bit_bytes >= 2048u) {
        // one mbarrier, one expected byte count, one or two bulk copies (row-major and transposed bits)
        if (threadIdx.x == 0) {
            const uint32_t b = smem_u32(mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(all_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sbits)), "l"(gbits), "r"(bit_bytes), "r"(b) : "memory");
            if (gbits_t)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sbits) + bit_bytes), "l"(gbits_t), "r"(bit_bytes), "r"(b) : "memory");
            if (ghier)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(shier)), "l"(ghier), "r"(hier_bytes), "r"(b) : "memory");
        }
        __syncthreads();   // the barrier is initialised before anyone polls it
        const uint32_t b = smem_u32(mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(b) : "memory");
        }
    } else {
        for (uint32_t i = threadIdx.x; i < (uint32_t)G.stride_words; i += blockDim.x) {
            sbits[i] = gbits[i];
            if (gbits_t) sbits[G.stride_words + i] = gbits_t[i];
        }
        for (uint32_t i = threadIdx.x; i < hier_bytes / 4u; i += blockDim.x)
            reinterpret_cast<uint32_t *>(shier)[i] = reinterpret_cast<const uint32_t *>(ghier)[i];
    }
    for (int i = threadIdx.x; i < G.S; i += blockDim.x) { sx[i] = G.min_x[i]; sy[i] = G.min_y[i]; }
    __syncthreads();
    V.bits = sbits; V.bits_t = G.bits_t ? sbits + G.stride_words : nullptr; V.min_x = sx; V.min_y = sy;
    view_hier(V, G, G.hier ? shier : nullptr);
    return V;
}

// ---------------------------------------------------------------- block-level statistics reduction
// per-thread counters -> warp shuffles -> shared atomics -> ONE global atomic per block per slot
__device__ __forceinline__ void block_accumulate_stats(const long long (&loc)[AG_ST_COUNT], unsigned long long *gstats,
                                                       unsigned long long *s_acc /* [AG_ST_COUNT] shared, zeroed */) {
#pragma unroll
    for (int i = 0; i < AG_ST_COUNT; ++i) {
        long long v = loc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&s_acc[i], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x < AG_ST_COUNT && gstats != nullptr) {
        const unsigned long long v = s_acc[threadIdx.x];
        if (v != 0) atomicAdd(&gstats[threadIdx.x], v);
    }
}

}  // namespace agd
