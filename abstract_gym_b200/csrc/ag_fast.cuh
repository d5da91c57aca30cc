// ag_fast.cuh -- FAST engine: float32 interval filter in front of the float64 reference predicate.
//
// Idea (exact-predicate style): every decision of a step (does link L hit cell C, is the target
// reached) is first evaluated in float32 together with a bound on how far the float32 quantities
// can be from the real-arithmetic values of the float64 inputs.  A decision that clears its bound
// IS the reference's decision (the reference's own float64 rounding noise is orders of magnitude
// below the bounds, except for nearly axis-aligned links, which the filter refuses to decide);
// anything else is "undecided" and re-evaluated by the EXACT engine.  FAST == EXACT by
// construction; tests/test_gpu_parity.py checks it on >10^7 poses per grid class.
//
// Real-arithmetic restatement of utils/collision_checker.py:34-85 used by the filter (derivation
// in DESIGN.md "FAST engine"):  with cr_ij = (X_i-p0x)*dy - (Y_j-p0y)*dx,
//     hit  <=>  P2 and OVL and not B
//     P2  : some cr > 0 and some cr < 0          (line strictly separates the corners, :41-43)
//     OVL : the open x- and y-ranges of the segment overlap the square's   (t_in < 1, t_out > eps)
//     B   : both end points lie in the closed square                       (t_in <= eps, t_out >= 1)
// No division appears, so no error amplification for shallow links.
//
// Error budget (metres), scene_0-class geometry (|coordinates| < 2, links < 1):
//   AG_DELTA_P  float32 FK coordinate vs float64 FK coordinate      <= 3.0e-7 (measured 1.8e-7)
//   AG_DELTA_C  float32 corner (and +side) vs float64 corner        <= 1.5e-7
//   AG_M        any difference of a link coordinate and a corner    <= 8.0e-7  (delta_p+delta_c+rounding)
#pragma once
#include "ag_device.cuh"

// cold paths: out of line (ABI call) or inlined; see DESIGN.md "register allocation of K4"
#ifdef AG_COLD_INLINE
#define AG_COLD static __device__ __forceinline__
#else
#define AG_COLD static __device__ __noinline__
#endif

namespace agd {

constexpr float AG_DELTA_P = 3.0e-7f;
constexpr float AG_DELTA_C = 1.5e-7f;
constexpr float AG_M = 8.0e-7f;          // margin on (link coordinate - corner) comparisons
constexpr float AG_K1 = 1.2e-6f;         // cross-product error coefficient (see narrow_f32)
constexpr float AG_MIN_DXY = 1.0e-5f;    // below this |dx| or |dy| the filter refuses to decide
constexpr int AG_LIST_MAX = 8;           // obstacle-list broad phase for grids with <= 8 occupied cells
// Broad-phase margin of broad_list(): AG_M plus the rounding of its centre / half-extent arithmetic
// (square centre 1.2e-7, link centre and half extent 1.2e-7 each, two subtractions 2.4e-7; |coordinates| < 2).
constexpr float AG_M_BROAD = 2.0e-6f;

struct ArmF { float ex, ey, gx, gy; };

// per-thread float32 constants, hoisted out of the step loop
struct FastConst {
    float l1, l2, side, half, inv_side, tx, ty, reach_eps;
};
__host__ __device__ __forceinline__ FastConst make_fast_const(const ag_params &P, const GridDev &G) {
    FastConst C;
    C.l1 = (float)P.link_1; C.l2 = (float)P.link_2;
    C.side = (float)G.side; C.half = (float)G.half; C.inv_side = (float)G.inv_side;
    C.tx = (float)P.target_x; C.ty = (float)P.target_y; C.reach_eps = (float)P.reach_eps;
    return C;
}

// sin(pi*f)/f and cos(pi*f) for f in [-0.5, 0.5] (half turns) are degree-4 polynomials in u = f*f, near-minimax
// fits (approximation error 1.4e-8 / 4.7e-8).  Evaluated in float32 the sine and cosine are within 2.1e-7 of the
// float64 values, the link end points within 1.4e-7 m (4e6 random angles, emulated op for op in numpy: DESIGN.md
// "FAST engine"); both angles go through the same Horner chain at once (fast_forward_kinematics below).

// Packed float32x2 arithmetic (Blackwell FFMA2 / FMUL2: two float32 operations per issue slot); the
// two halves carry the two joint angles through the same polynomial.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)),
        "l"(*reinterpret_cast<unsigned long long *>(&b)), "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)),
        "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// half-turn remainder f (float32) and signed length of one float64 angle: j = pi*(k + f), |f| <= 0.5;
// sin(j) = (-1)^k sin(pi f), cos(j) = (-1)^k cos(pi f): the sign is folded into the link length with one
// XOR, so there is no select on the ALU pipe.  The reduction stays in float64 (3 DP ops).
__device__ __forceinline__ void reduce_half_turns(double j, float len, float &f, float &sl) {
    const double t = j * 0.31830988618379067154;               // half turns
    const double tk = t + 6755399441055744.0;                  // 1.5*2^52: rounds to nearest integer
    const uint32_t k = (uint32_t)__double2loint(tk);
    f = (float)(t - (tk - 6755399441055744.0));
    sl = __uint_as_float(__float_as_uint(len) ^ (k << 31));
}

// ok=false when an angle is outside the range where the reduction above is trustworthy (also NaN)
__device__ __forceinline__ ArmF fast_forward_kinematics(double j1, double j2, const FastConst &C, bool &ok) {
    float2 f, sl;
    reduce_half_turns(j1, C.l1, f.x, sl.x);
    reduce_half_turns(j2, C.l2, f.y, sl.y);
    ok = (fabs(j1) < 1048576.0) && (fabs(j2) < 1048576.0);     // two DSETP on the (idle) FP64 pipe
    // both angles at once, FFMA2
    const float2 u = mul2(f, f);
    float2 ps = splat2(0.07765940576791763f);
    ps = fma2(ps, u, splat2(-0.5982921719551086f));
    ps = fma2(ps, u, splat2(2.5500776767730713f));
    ps = fma2(ps, u, splat2(-5.167710304260254f));
    ps = fma2(ps, u, splat2(3.1415927410125732f));
    float2 pc = splat2(0.2196967899799347f);
    pc = fma2(pc, u, splat2(-1.3318802118301392f));
    pc = fma2(pc, u, splat2(4.058412075042725f));
    pc = fma2(pc, u, splat2(-4.934792995452881f));
    pc = fma2(pc, u, splat2(0.9999999403953552f));
    const float2 s = mul2(mul2(ps, f), sl), c = mul2(pc, sl);  // (sin j1 * l1, sin j2 * l2), (cos j1 * l1, cos j2 * l2)
    ArmF a;
    a.ex = c.x; a.ey = s.x;
    a.gx = c.y + a.ex; a.gy = s.y + a.ey;
    return a;
}

// ---------------------------------------------------------------- narrow phase, one (link, cell)
// returns 0 = certainly no hit, 1 = certainly hit, 2 = undecided.
struct LinkF {
    float p0x, p0y, dx, dy, xlo, xhi, ylo, yhi, ecr;
    bool degenerate;
};

__device__ __forceinline__ LinkF make_link_f(float p0x, float p0y, float p1x, float p1y, float side) {
    LinkF L;
    L.p0x = p0x; L.p0y = p0y; L.dx = p1x - p0x; L.dy = p1y - p0y;
    L.xlo = fminf(p0x, p1x); L.xhi = fmaxf(p0x, p1x); L.ylo = fminf(p0y, p1y); L.yhi = fmaxf(p0y, p1y);
    // |error of cr| <= (delta_d + 2^-23)*(|ux|+|uy|) + delta_u*(|dx|+|dy|) with delta_d = 6.4e-7, delta_u = 5.5e-7;
    // for a candidate cell (ranges overlap within AG_M) |ux| <= |dx| + side + AG_M, same for y.
    L.ecr = AG_K1 * (2.0f * (fabsf(L.dx) + fabsf(L.dy)) + 2.02f * side + 4.0e-6f);
    L.degenerate = fminf(fabsf(L.dx), fabsf(L.dy)) < AG_MIN_DXY;
    return L;
}

__device__ __forceinline__ int narrow_f32(const LinkF &L, float min_x, float min_y, float max_x, float max_y) {
    // corner sign test: max/min over the four cross products
    const float ux0 = min_x - L.p0x, ux1 = max_x - L.p0x, uy0 = min_y - L.p0y, uy1 = max_y - L.p0y;
    const float a0 = ux0 * L.dy, a1 = ux1 * L.dy, b0 = uy0 * L.dx, b1 = uy1 * L.dx;
    const float cmax = fmaxf(a0, a1) - fminf(b0, b1), cmin = fminf(a0, a1) - fmaxf(b0, b1);
    if (!((cmax > -L.ecr) && (cmin < L.ecr))) return 0;   // the line certainly misses the open square
    // OVL from four differences: ranges certainly disjoint => no hit, whatever the link's direction (for an exactly
    // axis-aligned link the defined behaviour is the closed interval-overlap test, which also needs the overlap)
    const float a_min = fminf(fminf(L.xhi - min_x, max_x - L.xlo), fminf(L.yhi - min_y, max_y - L.ylo));
    if (!(a_min > -AG_M)) return 0;
    // A nearly axis-aligned link: the reference switches formulas at exactly dx == 0 / dy == 0 and
    // its lambda arithmetic gets noisy below |dx| ~ 1e-9; leave every such case to the EXACT engine.
    if (L.degenerate) return 2;
    // B from four more differences: segment wholly inside => no hit
    const float b_min = fminf(fminf(L.xlo - min_x, max_x - L.xhi), fminf(L.ylo - min_y, max_y - L.yhi));
    if (b_min > AG_M) return 0;
    const bool p2_certain = (cmax > L.ecr) && (cmin < -L.ecr);
    return (p2_certain && (a_min > AG_M) && (b_min < -AG_M)) ? 1 : 2;
}

// ---------------------------------------------------------------- obstacle-list broad phase (<= AG_LIST_MAX cells)
// Built per block in shared memory from the staged bit grid, row-major order.
struct FastList {
    int m;                          // number of occupied cells, or -1 if the list form does not apply
    float hm;                       // side/2 + AG_M_BROAD: centre-distance threshold of the broad phase
    float2 ctr[AG_LIST_MAX];        // square centres as float32
    float4 sq[AG_LIST_MAX];         // (min_x, min_y, max_x, max_y) as float32
    int cell[AG_LIST_MAX];          // (row << 16) | col
    const double *vmin_x, *vmin_y;  // the block's float64 corner tables (second-level filter)
};

// warp 0 builds the list (S <= 32: one row word per lane)
__device__ __forceinline__ void build_fast_list(const GridDev &G, const GridView &V, FastList *fl) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int m = -1;
        if (G.S <= 32) {
            const uint32_t word = lane < G.S ? V.bits[lane] : 0u;
            int cnt = __popc(word), pre = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, pre, o);
                if (lane >= o) pre += v;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, pre, 31);
            if (total <= AG_LIST_MAX) {
                m = total;
                int slot = pre - cnt;
                uint32_t w = word;
                const float side = (float)G.side;
                while (w) {
                    const int c = __ffs(w) - 1;
                    w &= w - 1;
                    const float mnx = (float)V.min_x[c], mny = (float)V.min_y[lane];
                    fl->ctr[slot] = make_float2(fmaf(0.5f, side, mnx), fmaf(0.5f, side, mny));
                    fl->cell[slot] = (lane << 16) | c;
                    fl->sq[slot++] = make_float4(mnx, mny, mnx + side, mny + side);
                }
            }
        }
        // m < 0 (list form does not apply): an infinite threshold sends every lane through the slow branch
        if (lane == 0) { fl->vmin_x = V.min_x; fl->vmin_y = V.min_y; fl->m = m; fl->hm = m < 0 ? __int_as_float(0x7f800000) : fmaf(0.5f, (float)G.side, AG_M_BROAD); }
    }
    __syncthreads();
}

// Branch-free broad phase of the hot loop.  For link l (centre c_l, half extents h_l) and square k (centre o_k,
// half side h): the closed boxes come within AG_M of each other  =>  max(|c_lx-o_kx| - h_lx, |c_ly-o_ky| - h_ly) < h + AG_M.
// Returns the minimum of that measure over both links and all squares (compare with fl.hm).  8 FADD (FMA pipe)
// + 2 FMNMX + 1 FMNMX3 (ALU pipe) per square.
struct ArmBoxes { float c1x, c1y, c2x, c2y, h2x, h2y; };      // link centres; half extents: link 1 = |c1|, link 2 = h2

__device__ __forceinline__ ArmBoxes make_arm_boxes(const ArmF &a) {
    ArmBoxes b;
    b.c1x = 0.5f * a.ex; b.c1y = 0.5f * a.ey;                                // link 1: (0,0) -> elbow
    b.c2x = fmaf(0.5f, a.gx, b.c1x); b.c2y = fmaf(0.5f, a.gy, b.c1y);        // link 2: elbow -> end effector
    b.h2x = fabsf(b.c2x - a.ex); b.h2y = fabsf(b.c2y - a.ey);
    return b;
}

// separation measures of both links against the square centred at o (compare with FastList::hm)
__device__ __forceinline__ void box_measures(const ArmBoxes &b, float2 o, float &m1, float &m2) {
    m1 = fmaxf(fabsf(b.c1x - o.x) - fabsf(b.c1x), fabsf(b.c1y - o.y) - fabsf(b.c1y));
    m2 = fmaxf(fabsf(b.c2x - o.x) - b.h2x, fabsf(b.c2y - o.y) - b.h2y);
}

__device__ __forceinline__ float broad_list(const FastList &fl, const ArmF &a) {
    const ArmBoxes b = make_arm_boxes(a);
    const int m = fl.m;
    float acc = 1.0e30f, m1, m2;
    if (m == 3) {                            // scene_0's manual map (occupancy_grid.py:45-47): straight-line code
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            box_measures(b, fl.ctr[k], m1, m2);
            acc = fminf(acc, fminf(m1, m2));
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < m; ++k) {        // warp-uniform trip count; kept rolled so the hot loop stays compact
            box_measures(b, fl.ctr[k], m1, m2);
            acc = fminf(acc, fminf(m1, m2));
        }
    }
    return acc;
}

// 0 / 1 certain, 2 undecided.  Warp-uniform loop over the (few) obstacles: the broad-phase measures pick
// the (link, square) pairs, then the narrow phase runs on the surviving pairs.
// Broad phase over the obstacle list: bit 2k = link 1 may touch square k, bit 2k+1 = link 2 may touch square k.
// check_link1 = false: the caller already knows that link 1 is clear of every square (the hazard bit of ag_rollout_lut.cu)
__device__ __forceinline__ uint32_t list_candidates(const FastList *fl, const ArmF &a, bool check_link1 = true) {
    const int m = fl->m;
    const float hm = fl->hm;
    const ArmBoxes bx = make_arm_boxes(a);
    uint32_t cand = 0;
#pragma unroll 1
    for (int k = 0; k < m; ++k) {
        float m1, m2;
        box_measures(bx, fl->ctr[k], m1, m2);
        cand |= (((check_link1 && m1 < hm) ? 1u : 0u) | (m2 < hm ? 2u : 0u)) << (2 * k);
    }
    return cand;
}

// Narrow phase on the surviving (link, cell) pairs: 0 / 1 certain, 2 undecided.  The link is rebuilt per pair from `a`
// (a dozen instructions) instead of keeping two LinkF live: registers matter more here.
__device__ __forceinline__ int narrow_candidates(const FastList *fl, const ArmF &a, const FastConst &C, uint32_t cand) {
    int result = 0;
#pragma unroll 1
    while (cand) {
        const int b = __ffs(cand) - 1;
        cand &= cand - 1;
        const float4 q = fl->sq[b >> 1];
        const bool second = (b & 1) != 0;
        const LinkF L = make_link_f(second ? a.ex : 0.0f, second ? a.ey : 0.0f, second ? a.gx : a.ex,
                                    second ? a.gy : a.ey, C.side);
        const int v = narrow_f32(L, q.x, q.y, q.z, q.w);
        if (v == 1) return 1;
        result |= v;
    }
    return result;
}

// 0 / 1 certain, 2 undecided.  Warp-uniform loop over the (few) obstacles: the broad-phase measures pick
// the (link, square) pairs, then the narrow phase runs on the surviving pairs.
__device__ __forceinline__ int arm_fast_list(const FastList *fl, const ArmF &a, const FastConst &C, bool check_link1 = true) {
    return narrow_candidates(fl, a, C, list_candidates(fl, a, check_link1));
}

// ---------------------------------------------------------------- traversal broad phase (any grid), one link
// conservative float32 traversal of the bit grid (same scheme as link_exact, wider margins).
// floor() is replaced by a round-to-nearest magic-number add biased so that the low index can
// only come out lower and the high index only higher (no F2I/FRND on the XU pipe).
__device__ __forceinline__ int round_magic(float v) { return __float_as_int(v + 12582912.0f) - 0x4B400000; }

// The traversal walks the bit grid line by line: a "line" is a row of the row-major bits, or a column of the
// transposed copy; positions within a line are columns resp. rows.  When the grid carries the transposed copy the
// lines run along the link's MAJOR axis (rows for a shallow link, columns for a steep one), so that the link crosses
// as few lines as possible: at most len/side/sqrt(2) instead of up to len/side, on average 0.37 instead of 0.64 of it;
// a line's position interval then spans |dp/dl| >= 1 cells.  (Until late in round 2 the choice was the other way
// round -- most lines, one-cell intervals: 20 % slower on the 256 x 256 maps.)
struct LinkScan {
    float p0x, p0y, p1x, p1y, side;
    LinkF L;
    bool have_link;
    bool swapped;      // lines are columns, positions are rows
    int result;
};

// narrow phase over the set bits of one (already masked) word of line `line`; true = certain hit
__device__ __forceinline__ bool scan_word(const GridView &V, uint32_t word, int w, int line, LinkScan &K) {
    while (word) {
        const int pos = (w << 5) + __ffs(word) - 1;
        word &= word - 1;
        if (!K.have_link) { K.L = make_link_f(K.p0x, K.p0y, K.p1x, K.p1y, K.side); K.have_link = true; }
        const int r = K.swapped ? pos : line, c = K.swapped ? line : pos;
        const float mnx = (float)V.min_x[c], mny = (float)V.min_y[r];
        const int v = narrow_f32(K.L, mnx, mny, mnx + K.side, mny + K.side);
        if (v == 1) return true;
        K.result |= v;          // 0 or 2
    }
    return false;
}

// occupied cells of one line at positions [p_lo, p_hi] -> narrow phase; true = certain hit (result accumulates 0 / 2)
__device__ __forceinline__ bool scan_line(const GridView &V, const uint32_t *__restrict__ linep, int line, int p_lo, int p_hi,
                                          LinkScan &K) {
    const int w0 = p_lo >> 5, w1 = p_hi >> 5;
    const uint32_t mlo = 0xFFFFFFFFu << (p_lo & 31), mhi = 0xFFFFFFFFu >> (31 - (p_hi & 31));
    if (w0 == w1) {                                     // the usual case: the interval sits inside one word
        const uint32_t word = linep[w0] & mlo & mhi;
        return word != 0 && scan_word(V, word, w0, line, K);
    }
    for (int w = w0; w <= w1; ++w) {
        uint32_t word = linep[w];
        if (w == w0) word &= mlo;
        if (w == w1) word &= mhi;
        if (word != 0 && scan_word(V, word, w, line, K)) return true;
    }
    return false;
}

__device__ __forceinline__ int link_fast(const GridDev &G, const GridView &V, const FastConst &C, float p0x, float p0y,
                                         float p1x, float p1y) {
    const float mcell = fmaxf(2.0e-6f * C.inv_side, 1.0e-3f) + 0.001f;   // margin in cells
    const int S1 = G.S - 1;
    const float dx = p1x - p0x, dy = p1y - p0y;
    // cell coordinates: column of x = round(x*inv_side + xoff); row of y = round(-y*inv_side + roff)
    // (environment/occupancy_grid.py:59-67: row 0 is the top row, S cells of side E/(S-1), not centred)
    const float xoff = C.half * C.inv_side - 0.5f, roff = C.half * C.inv_side + 0.5f;
    const float ua = fmaf(p0x, C.inv_side, xoff), ub = fmaf(p1x, C.inv_side, xoff);      // columns of the end points
    const float va = fmaf(-p0y, C.inv_side, roff), vb = fmaf(-p1y, C.inv_side, roff);    // rows of the end points
    LinkScan K;
    K.p0x = p0x; K.p0y = p0y; K.p1x = p1x; K.p1y = p1y; K.side = C.side; K.have_link = false; K.result = 0;
    // lines along the link's major axis: rows for a shallow link, columns (transposed copy) for a steep one
    K.swapped = (V.bits_t != nullptr) && (fabsf(dy) > fabsf(dx));
    // line coordinate l (rows, or columns when swapped) and position coordinate p of the two end points
    const float la = K.swapped ? ua : va, lb = K.swapped ? ub : vb;
    const float pa = K.swapped ? va : ua, pb = K.swapped ? vb : ub;
    int l_lo = round_magic(fminf(la, lb) - mcell), l_hi = round_magic(fmaxf(la, lb) + mcell);
    if (l_lo > S1 || l_hi < 0) return 0;
    l_lo = max(l_lo, 0); l_hi = min(l_hi, S1);
    const float pseg_lo = fminf(pa, pb), pseg_hi = fmaxf(pa, pb);
    const uint32_t *linep = (K.swapped ? V.bits_t : V.bits) + l_lo * G.wpr;
    // d(line coordinate) along the link vs d(position coordinate): |dl| <= |dp| with the transposed copy; a link
    // that is more than 64 positions per line (nearly parallel to the lines) is not tracked
    const float dl = lb - la, dp = pb - pa;
    const bool tracked = (l_hi - l_lo >= 2) && (fabsf(dl) * 64.0f >= fabsf(dp));
    if (!tracked) {
        // few lines, or (no transposed copy) a shallow link walked by rows: every line gets the whole position range
        const int p_lo = max(round_magic(pseg_lo - mcell), 0), p_hi = min(round_magic(pseg_hi + mcell), S1);
        if (p_lo > p_hi) return 0;
        for (int l = l_lo; l <= l_hi; ++l, linep += G.wpr) {
            AG_CHECK_INDEX(l, G.S); AG_CHECK_INDEX(p_lo, G.S); AG_CHECK_INDEX(p_hi, G.S);
            if (scan_line(V, linep, l, p_lo, p_hi, K)) return 1;
        }
        return K.result;
    }
    // One line per iteration, the position interval follows the link incrementally: along the link p is linear in
    // l with slope s = dp/dl, cell (line) boundaries sit at half-integers of the line coordinate, so the interval
    // of line l is [p(l - 0.5), p(l + 0.5)]: one FMA per line from the fixed start (no accumulation).  It is
    // widened by mm cells -- mcell of slack in l costs |s|*mcell in p, plus the float32 error of s and the start
    // (<= 5e-5*inv_side) -- and clamped to the link's own position range.
    const float s = dp * __frcp_rn(dl);
    const float mm = mcell + 5.0e-5f * C.inv_side + fabsf(s) * mcell;
    const float clamp_lo = pseg_lo - mm, clamp_hi = pseg_hi + mm;
    const float pstart = fmaf(((float)l_lo - 0.5f) - la, s, pa);             // p at the near boundary of line l_lo
    float pprev = pstart, k = 1.0f;
    for (int l = l_lo; l <= l_hi; ++l, linep += G.wpr, k += 1.0f) {
        const float pcur = fmaf(k, s, pstart);                               // p at the far boundary of line l
        const float lo = fmaxf(fminf(pprev, pcur) - mm, clamp_lo), hi = fminf(fmaxf(pprev, pcur) + mm, clamp_hi);
        pprev = pcur;
        const int p_lo = max(round_magic(lo), 0), p_hi = min(round_magic(hi), S1);
        if (p_lo > p_hi) continue;
        AG_CHECK_INDEX(l, G.S); AG_CHECK_INDEX(p_lo, G.S); AG_CHECK_INDEX(p_hi, G.S);
        AG_CHECK_INDEX(linep - (K.swapped ? V.bits_t : V.bits) + (p_hi >> 5), G.S * G.wpr);
        if (scan_line(V, linep, l, p_lo, p_hi, K)) return 1;
    }
    return K.result;
}

// ---------------------------------------------------------------- two-level traversal (grids with ag_grid.hier)
// The same conservative line walk as link_fast, but over the T x T summary bitmap (one bit per 8x8 tile): 8x fewer
// lines and words, and only the cells of OCCUPIED tiles the link passes are examined -- each tile is one 64-bit word
// whose set bits go to the narrow phase (narrow_f32's first test discards the cells the line misses in ~15
// instructions).  Tile coordinates: fine column c = round(u) with u = x*inv_side + xoff covers u in [c-1/2, c+1/2), so
// tile column C = c >> 3 covers (u + 1/2)/8 - 1/2 in [C-1/2, C+1/2); rows alike.
//
// Two phases per pose, because the lanes of a warp meet their occupied tiles on different lines: with the narrow phase
// nested in the line loop every line trip ran ~80 instructions for the few lanes that had a tile there while the others
// waited (profiles/r2 "h_c4": 2.8 threads per instruction in that part, 5.1 overall).  Phase A walks the summary lines
// of BOTH links and only pushes the occupied tiles (tile index | link << 15) onto a per-lane queue in shared memory;
// phase B drains the queue, so all lanes run the narrow phase together.
// A cell the float32 narrow phase cannot settle is settled on the spot by the reference predicate in float64
// (segment_square_exact on the float64 arm, get_arm()) -- before, one undecided cell sent the lane through the whole
// float64 traversal of the bit grid (~4000 instructions on a 1024 x 1024 map, 17 % of all instructions of config 4).
// Nearly axis-aligned links still return 2 (the caller's float64 path knows how the reference treats them).
constexpr int AG_TILEQ = 24;
constexpr int AG_TILEQ_COLS = 256;      // threads per block of every kernel that reaches the traversal (AG_BLOCK)
constexpr int AG_HIER_MAX_T = 128;      // a queue entry is (tile row << 7 | tile column) | link << 15: S <= 1024

// phase A for one link: push(tile row << 7 | tile column) for every occupied tile within the link's margin
template <typename Push>
__device__ __forceinline__ void link_tiles(const GridDev &G, const GridView &V, const FastConst &C, float p0x, float p0y, float p1x,
                                           float p1y, Push push) {
    // margin in tiles: the fine walk's margin (cells) / 8, plus slack for the float32 tile coordinates
    const float mcell = (fmaxf(2.0e-6f * C.inv_side, 1.0e-3f) + 0.001f) * 0.125f + 0.001f;
    const int T1 = G.T - 1;
    const float inv8 = C.inv_side * 0.125f;
    const float hi8 = C.half * inv8;
    const float xoff = hi8 - 0.5f, roff = hi8 + (0.125f - 0.5f);          // ((xoff_fine + 1/2)/8 - 1/2), ((roff_fine + 1/2)/8 - 1/2)
    const float ca = fmaf(p0x, inv8, xoff), cb = fmaf(p1x, inv8, xoff);    // tile columns of the end points
    const float ra = fmaf(-p0y, inv8, roff), rb = fmaf(-p1y, inv8, roff);  // tile rows of the end points
    // The lines run along the link's major axis, so that it crosses as few of them as possible: tile rows for a
    // shallow link, tile columns (the transposed summary) for a steep one -- at most len/sqrt(2) lines, 0.37 len on
    // average instead of 0.64 len; a line's position interval then spans |dp/dl| >= 1 tiles (one or two words).
    const bool swapped = fabsf(p1y - p0y) > fabsf(p1x - p0x);
    const float la = swapped ? ca : ra, lb = swapped ? cb : rb;
    const float pa = swapped ? ra : ca, pb = swapped ? rb : cb;
    int l_lo = round_magic(fminf(la, lb) - mcell), l_hi = round_magic(fmaxf(la, lb) + mcell);
    if (l_lo > T1 || l_hi < 0) return;
    l_lo = max(l_lo, 0); l_hi = min(l_hi, T1);
    const float pseg_lo = fminf(pa, pb), pseg_hi = fmaxf(pa, pb);
    const uint32_t *linep = (swapped ? V.coarse_t : V.coarse) + l_lo * G.cwpr;
    const float dl = lb - la, dp = pb - pa;
    const bool tracked = (l_hi - l_lo >= 2) && (fabsf(dl) * 64.0f >= fabsf(dp));
    // the position interval of tile row l follows the link: [p(l - 1/2), p(l + 1/2)] widened as in link_fast;
    // untracked (few lines, or a very shallow link): every line gets the link's whole position range
    const float s = tracked ? dp * __frcp_rn(dl) : 0.0f;
    const float mm = mcell + 5.0e-5f * inv8 + fabsf(s) * mcell;
    const float clamp_lo = pseg_lo - mm, clamp_hi = pseg_hi + mm;
    const float pstart = fmaf(((float)l_lo - 0.5f) - la, s, pa);
    float pprev = pstart, k = 1.0f;
    for (int l = l_lo; l <= l_hi; ++l, linep += G.cwpr, k += 1.0f) {
        float lo = clamp_lo, hi = clamp_hi;
        if (tracked) {
            const float pcur = fmaf(k, s, pstart);
            lo = fmaxf(fminf(pprev, pcur) - mm, clamp_lo); hi = fminf(fmaxf(pprev, pcur) + mm, clamp_hi);
            pprev = pcur;
        }
        const int p_lo = max(round_magic(lo), 0), p_hi = min(round_magic(hi), T1);
        if (p_lo > p_hi) continue;
        AG_CHECK_INDEX(l, G.T); AG_CHECK_INDEX(p_lo, G.T); AG_CHECK_INDEX(p_hi, G.T);
        const int w0 = p_lo >> 5, w1 = p_hi >> 5;
        const uint32_t mlo = 0xFFFFFFFFu << (p_lo & 31), mhi = 0xFFFFFFFFu >> (31 - (p_hi & 31));
        for (int w = w0; w <= w1; ++w) {
            uint32_t word = linep[w];
            if (w == w0) word &= mlo;
            if (w == w1) word &= mhi;
            while (word) {
                const int pos = (w << 5) + __ffs(word) - 1;
                word &= word - 1;
                push(swapped ? ((pos << 7) | l) : ((l << 7) | pos));                 // (tile row << 7) | tile column
            }
        }
    }
}

// both links of one pose: 0 / 1 certain, 2 undecided.  get_arm(): the pose's float64 arm (reference arithmetic), evaluated
// only when a cell needs the float64 predicate; eps: CollisionChecker's section epsilon.
template <typename GetArm>
__device__ __forceinline__ int arm_fast_hier(const GridDev &G, const GridView &V, const FastConst &C, const ArmF &a, GetArm get_arm,
                                             double eps) {
    __shared__ unsigned short s_tileq[AG_TILEQ][AG_TILEQ_COLS];
    AG_CHECK_INDEX(G.T, AG_HIER_MAX_T + 1);
    AG_CHECK_INDEX(threadIdx.x, AG_TILEQ_COLS);
    const int tx = threadIdx.x;
    int cnt = 0, result = 0;
    bool hit = false;
    // phase B: drain this lane's queue
    auto drain = [&]() {
        for (int i = 0; i < cnt && !hit; ++i) {
            const unsigned e = s_tileq[i][tx];
            const bool second = (e >> 15) != 0;
            const int tr = (int)((e >> 7) & 0x7Fu), tc = (int)(e & 0x7Fu);
            AG_CHECK_INDEX(tr, G.T); AG_CHECK_INDEX(tc, G.T);
            unsigned long long tile = V.tiles[tr * G.T + tc];
            const LinkF L = make_link_f(second ? a.ex : 0.0f, second ? a.ey : 0.0f, second ? a.gx : a.ex, second ? a.gy : a.ey, C.side);
            while (tile) {
                const int b = __ffsll((long long)tile) - 1;
                tile &= tile - 1;
                const int r = (tr << 3) + (b >> 3), c = (tc << 3) + (b & 7);
                AG_CHECK_INDEX(r, G.S); AG_CHECK_INDEX(c, G.S);
                const double mnxd = V.min_x[c], mnyd = V.min_y[r];
                const float mnx = (float)mnxd, mny = (float)mnyd;
                int v = narrow_f32(L, mnx, mny, mnx + C.side, mny + C.side);
                if (v == 2 && !L.degenerate) {                              // settle this cell with the reference predicate
                    const Arm A = get_arm();
                    const LineD Ld = second ? make_line(A.ex, A.ey, A.gx, A.gy) : make_line(0.0, 0.0, A.ex, A.ey);
                    int axis = 0;
                    v = segment_square_exact(Ld, mnxd, mnyd, __dadd_rn(mnxd, G.side), __dadd_rn(mnyd, G.side), eps, axis) ? 1 : 0;
                }
                if (v == 1) { hit = true; break; }
                result |= v;
            }
        }
        cnt = 0;
    };
    auto push_link = [&](unsigned link_bit) {
        return [&, link_bit](int rc) {
            if (cnt == AG_TILEQ) drain();                                   // rare: more than AG_TILEQ occupied tiles on the way
            s_tileq[cnt++][tx] = (unsigned short)((unsigned)rc | link_bit);
        };
    };
    // phase A, both links (no reconvergence point is needed between them: nothing lane-dependent is nested inside)
    link_tiles(G, V, C, 0.0f, 0.0f, a.ex, a.ey, push_link(0u));
    if (!hit) link_tiles(G, V, C, a.ex, a.ey, a.gx, a.gy, push_link(0x8000u));
    if (!hit) drain();
    return hit ? 1 : result;
}

// broad-phase selection: a compile-time choice for the rollout kernel (keeps its register
// allocation small), a run-time one (BP_ANY) for K1..K3
enum { BP_ANY = 0, BP_LIST = 1, BP_TRAVERSAL = 2 };

// 0 / 1 certain, 2 undecided
template <int BP, typename GetArm>
__device__ __forceinline__ int arm_fast(const GridDev &G, const GridView &V, const FastList *fl, const FastConst &C,
                                        const ArmF &a, GetArm get_arm, double eps) {
    if (BP == BP_LIST) return (fl != nullptr && fl->m >= 0) ? arm_fast_list(fl, a, C) : 2;
    if (BP == BP_ANY && fl != nullptr && fl->m >= 0) return arm_fast_list(fl, a, C);
    if (V.tiles != nullptr) {                                               // uniform over the launch
        const unsigned lanes_h = __activemask();
        const int vh = arm_fast_hier(G, V, C, a, get_arm, eps);
        __syncwarp(lanes_h);
        return vh;
    }
    // The two traversals have lane-dependent trip counts.  Without an explicit reconvergence point between them
    // the lanes that finish link 1 first run ahead ALONE into link 2 (independent thread scheduling): ncu showed
    // link 2's row loop at 1.0 active thread per instruction (profiles/r1_c4_*).  Hence: no early return between
    // the links, and a __syncwarp over the lanes that entered together.
    const unsigned lanes = __activemask();
    const int v1 = link_fast(G, V, C, 0.0f, 0.0f, a.ex, a.ey);
    __syncwarp(lanes);
    const int v2 = (v1 == 1) ? 0 : link_fast(G, V, C, a.ex, a.ey, a.gx, a.gy);
    __syncwarp(lanes);
    return (v1 == 1 || v2 == 1) ? 1 : (v1 | v2);
}

// reach test filter, scenario/scene_0.py:129-130 ; 0/1 certain, 2 undecided
__device__ __forceinline__ int reach_fast_at(const FastConst &C, const ArmF &a, float tx, float ty) {
    const float m = AG_DELTA_P + 2.0e-7f;
    const float worst = fmaxf(fabsf(tx - a.gx), fabsf(ty - a.gy));        // reached <=> worst < eps
    if (worst > C.reach_eps + m) return 0;
    if (worst < C.reach_eps - m) return 1;
    return 2;
}
__device__ __forceinline__ int reach_fast(const FastConst &C, const ArmF &a) { return reach_fast_at(C, a, C.tx, C.ty); }

// ---------------------------------------------------------------- cold path: the float64 reference arithmetic
// Kept out of line so that the hot loop's register allocation is not dictated by it.
// Second-level filter, float64: the same division-free decision as narrow_f32 on the float64 arm, with
// margins that only have to cover (a) our FK vs the reference's (CUDA sincos <= 2 ulp, glibc <= 1 ulp:
// <= 4e-16 m) and (b) the reference's own rounding noise, which for |dx|,|dy| >= AG_MIN_DXY stays below
// 1e-12 m (x1 = (-c - b*min_y)/a has absolute error ~ eps*|min_y*dx/dy|).  It settles the float32 filter's
// undecided band (e.g. scene_0's link 1, whose 0.4 m reach is tangent to two squares) without the
// reference's four divisions per pair; what it cannot settle goes to the EXACT engine.
constexpr double AG_M64 = 1.0e-9;

__device__ __forceinline__ int narrow_f64(double p0x, double p0y, double p1x, double p1y, double min_x, double min_y,
                                          double max_x, double max_y, double side) {
    const double dx = p1x - p0x, dy = p1y - p0y;
    const double ecr = AG_M64 * (2.0 * (fabs(dx) + fabs(dy)) + 2.0 * side);
    const double ux0 = min_x - p0x, ux1 = max_x - p0x, uy0 = min_y - p0y, uy1 = max_y - p0y;
    const double a0 = ux0 * dy, a1 = ux1 * dy, b0 = uy0 * dx, b1 = uy1 * dx;
    const double cmax = fmax(a0, a1) - fmin(b0, b1), cmin = fmin(a0, a1) - fmax(b0, b1);
    if (!((cmax > -ecr) && (cmin < ecr))) return 0;                       // the line certainly misses the open square
    if (fmin(fabs(dx), fabs(dy)) < (double)AG_MIN_DXY) return 2;
    const double xlo = fmin(p0x, p1x), xhi = fmax(p0x, p1x), ylo = fmin(p0y, p1y), yhi = fmax(p0y, p1y);
    const double a_min = fmin(fmin(xhi - min_x, max_x - xlo), fmin(yhi - min_y, max_y - ylo));
    const double b_min = fmin(fmin(xlo - min_x, max_x - xhi), fmin(ylo - min_y, max_y - yhi));
    if (!(a_min > -AG_M64) || b_min > AG_M64) return 0;                   // ranges certainly disjoint, or segment wholly inside
    const bool p2_certain = (cmax > ecr) && (cmin < -ecr);
    return (p2_certain && (a_min > AG_M64) && (b_min < -AG_M64)) ? 1 : 2;
}

// 0 / 1 certain, 2 undecided: both links against every square of the obstacle list, float64
__device__ __forceinline__ int arm_f64_list(const GridDev &G, const FastList *fl, const Arm &A) {
    int result = 0;
    const int m = fl->m;
#pragma unroll 1
    for (int k = 0; k < 2 * m; ++k) {
        const float4 qf = fl->sq[k >> 1];        // float32 copy of the corner: only used to skip far squares
        const bool second = (k & 1) != 0;
        const double p0x = second ? A.ex : 0.0, p0y = second ? A.ey : 0.0, p1x = second ? A.gx : A.ex, p1y = second ? A.gy : A.ey;
        const double far = 1.0e-5;
        if (fmax(p0x, p1x) < (double)qf.x - far || fmin(p0x, p1x) > (double)qf.z + far ||
            fmax(p0y, p1y) < (double)qf.y - far || fmin(p0y, p1y) > (double)qf.w + far) continue;
        const int cell = fl->cell[k >> 1];       // (row << 16) | col of the square: exact float64 corners
        const double mnx = fl->vmin_x[cell & 0xFFFF], mny = fl->vmin_y[cell >> 16];
        const int v = narrow_f64(p0x, p0y, p1x, p1y, mnx, mny, __dadd_rn(mnx, G.side), __dadd_rn(mny, G.side), G.side);
        if (v == 1) return 1;
        result |= v;
    }
    return result;
}

// ---------------------------------------------------------------- cold path: float64
// Kept out of line so that the hot loop's register allocation is not dictated by it.
// c / r: the float32 filter's verdicts (0, 1, or 2 = undecided).
// (tx, ty): the cartesian target of this env (Scene.target_c, or its per-env override)
AG_COLD int cold_exact_decide_at(const ag_params &P, const GridDev &G, const GridView &V, const FastList *fl, double q1,
                                 double q2, int c, int r, double tx, double ty) {
    const Arm A = forward_kinematics(q1, q2, P.link_1, P.link_2);
    int fh = 0, axis = 0;
    if (c == 2 && fl != nullptr && fl->m >= 0) c = arm_f64_list(G, fl, A);
    const bool hit = (c == 2) ? arm_collides<AG_ENGINE_EXACT, false>(G, V, A, P.section_eps, fh, axis) : (c == 1);
    const bool reached = (r == 2) ? (fabs(__dsub_rn(tx, A.gx)) < P.reach_eps && fabs(__dsub_rn(ty, A.gy)) < P.reach_eps)   // scene_0.py:129-130
                                  : (r == 1);
    return (hit ? 1 : 0) | (reached ? 2 : 0) | (axis << 2);
}
__device__ __forceinline__ int cold_exact_decide(const ag_params &P, const GridDev &G, const GridView &V, const FastList *fl,
                                                 double q1, double q2, int c, int r) {
    return cold_exact_decide_at(P, G, V, fl, q1, q2, c, r, P.target_x, P.target_y);
}

// One step's two decisions for the FAST engine: bit0 collision, bit1 target reached, bits 2.. axis-aligned count.
// want_reach=false (reset candidates, K2, K3): only the collision bit is meaningful.
// tgt: this env's cartesian target (double[2], per-env override of Scene.target_c) or nullptr = P.target_x/y
template <int BP>
__device__ __forceinline__ int fast_decide(const ag_params &P, const GridDev &G, const GridView &V, const FastList *fl,
                                           const FastConst &C, double q1, double q2, bool want_reach,
                                           const double *tgt = nullptr) {
    bool ok;
    const ArmF a = fast_forward_kinematics(q1, q2, C, ok);
    const int c = ok ? arm_fast<BP>(G, V, fl, C, a, [&]() { return forward_kinematics(q1, q2, P.link_1, P.link_2); }, P.section_eps) : 2;
    int r = 0;
    double txd = P.target_x, tyd = P.target_y;
    if (want_reach) {
        if (P.choose_j_tar) r = target_reached_joint(P, q1, q2) ? 1 : 0;
        else {
            if (tgt != nullptr) { const double2 t2 = *reinterpret_cast<const double2 *>(tgt); txd = t2.x; tyd = t2.y; }
            r = ok ? reach_fast_at(C, a, (float)txd, (float)tyd) : 2;
        }
    }
    if (c == 2 || r == 2) return cold_exact_decide_at(P, G, V, fl, q1, q2, c, r, txd, tyd);
    return c | (r << 1);
}

// FAST collision_check when the float64 arm is already known (K1 needs it for its outputs)
__device__ __forceinline__ bool fast_arm_collides(const ag_params &P, const GridDev &G, const GridView &V,
                                                  const FastList *fl, const FastConst &C, const Arm &A, int &axis) {
    ArmF a;
    a.ex = (float)A.ex; a.ey = (float)A.ey; a.gx = (float)A.gx; a.gy = (float)A.gy;   // error 6e-8 < AG_DELTA_P
    const bool ok = fmax(fmax(fabs(A.ex), fabs(A.ey)), fmax(fabs(A.gx), fabs(A.gy))) < 1.0e3;
    const int v = ok ? arm_fast<BP_ANY>(G, V, fl, C, a, [&]() { return A; }, P.section_eps) : 2;
    if (v != 2) return v == 1;
    int fh = 0;
    return arm_collides<AG_ENGINE_EXACT, false>(G, V, A, P.section_eps, fh, axis);
}

}  // namespace agd
