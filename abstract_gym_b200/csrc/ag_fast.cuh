// ag_fast.cuh -- FAST engine: float32 interval filter in front of the float64 reference predicate.
//
// Idea (exact-predicate style): every decision of the step (corner sign test, entry/exit lambda
// test, reach test) is first evaluated in float32 together with a rigorous bound on how far the
// float32 value can be from the real-arithmetic value of the float64 inputs.  If every decision
// clears its bound the float32 answer IS the reference's answer (the reference's own float64
// rounding is ~1e-16, eight orders of magnitude below the filter bounds); otherwise the lane is
// "undecided" and re-evaluated by the EXACT engine.  The result is therefore identical to EXACT
// by construction; tests/test_gpu_parity.py checks it on 10^8+ poses.
//
// Error budget (metres unless noted), scene_0-class geometry (|coordinates| <= ~2):
//   FK:   quarter-turn reduction in float64, minimax sin/cos polynomials in float32:
//         |sin,cos error| <= 1.4e-7  ->  elbow/EE coordinates within AG_DELTA_P = 3e-7.
//   corners: (float)min_x[c] and +side: <= 1.0e-7 (AG_DELTA_C).
#pragma once
#include "ag_device.cuh"

namespace agd {

constexpr float AG_DELTA_P = 3.0e-7f;    // |float32 FK coordinate - float64 FK coordinate|
constexpr float AG_DELTA_C = 1.5e-7f;    // |float32 corner - float64 corner|
constexpr float AG_MIN_DXY = 1.0e-5f;    // below this |dx| or |dy| the filter refuses to decide

struct ArmF { float ex, ey, gx, gy; };

// sin(pi/2*f), cos(pi/2*f) for f in [-0.5, 0.5] (quarter turns); max abs error 9e-8 (host-checked,
// see DESIGN.md "FAST engine error budget").
__device__ __forceinline__ void sincos_quarter(float f, float &s, float &c) {
    const float u = f * f;
    float ps = -0.0046021631049598674f;
    ps = fmaf(ps, u, 0.07968022285816816f);
    ps = fmaf(ps, u, -0.6459634781534226f);
    ps = fmaf(ps, u, 1.5707963219600267f);
    s = ps * f;
    float pc = 0.0009036298853670689f;
    pc = fmaf(pc, u, -0.020860070409767503f);
    pc = fmaf(pc, u, 0.2536692037972198f);
    pc = fmaf(pc, u, -1.2337005406402657f);
    c = fmaf(pc, u, 0.9999999999525445f);
}

// float32 sin/cos of a float64 angle: j*(2/pi) and the rounding to the nearest quarter turn stay
// in float64 (3 DP ops), the rest is float32.  ok=false for |j| beyond the range where the
// float64 reduction is trustworthy.
__device__ __forceinline__ void sincos_f32_of_f64(double j, float &s, float &c, bool &ok) {
    const double t = j * 0.63661977236758134308;               // quarter turns
    const double tk = t + 6755399441055744.0;                  // 1.5*2^52: rounds to nearest integer
    const int k = __double2loint(tk);
    const float f = (float)(t - (tk - 6755399441055744.0));
    ok = ok && (fabs(j) < 1.0e6);
    float sq, cq;
    sincos_quarter(f, sq, cq);
    const float a = (k & 1) ? cq : sq, b = (k & 1) ? sq : cq;  // q=1,3 swap
    s = (k & 2) ? -a : a;                                      // q: 0:(s,c) 1:(c,-s) 2:(-s,-c) 3:(-c,s)
    c = ((k + 1) & 2) ? -b : b;
}

__device__ __forceinline__ ArmF fast_forward_kinematics(double j1, double j2, float l1, float l2, bool &ok) {
    float s1, c1, s2, c2;
    sincos_f32_of_f64(j1, s1, c1, ok);
    sincos_f32_of_f64(j2, s2, c2, ok);
    ArmF a;
    a.ex = c1 * l1; a.ey = s1 * l1;
    a.gx = fmaf(c2, l2, a.ex); a.gy = fmaf(s2, l2, a.ey);
    return a;
}

// ---------------------------------------------------------------- narrow phase, one (link, cell)
// returns 0 = certainly no hit, 1 = certainly hit, 2 = undecided.
// Real-arithmetic restatement of utils/collision_checker.py:34-85: with
//   cr_ij = (X_i - p0x)*dy - (Y_j - p0y)*dx           (sign(v_ij) = sign(cr_ij)*sign(dx*dy))
//   t_in = max(min(tx0,tx1), min(ty0,ty1)), t_out = min(max(tx0,tx1), max(ty0,ty1)),
//   tx_i = (X_i - p0x)/dx, ty_j = (Y_j - p0y)/dy
// the reference returns  (some cr > 0 and some cr < 0) and (eps < t_in < 1 or eps < t_out < 1).
struct LinkF { float p0x, p0y, dx, dy, rdx, rdy, et; bool degenerate; };

__device__ __forceinline__ LinkF make_link_f(float p0x, float p0y, float p1x, float p1y) {
    LinkF L;
    L.p0x = p0x; L.p0y = p0y; L.dx = p1x - p0x; L.dy = p1y - p0y;
    L.degenerate = fminf(fabsf(L.dx), fabsf(L.dy)) < AG_MIN_DXY;
    L.rdx = __frcp_rn(L.dx); L.rdy = __frcp_rn(L.dy);
    // |t error| for |t| <= 2 (clamped): numerator error + |t| * denominator error, both <= K2
    const float K2 = 2.0f * AG_DELTA_P + AG_DELTA_C + 2.4e-7f;
    L.et = 3.0f * K2 * fmaxf(fabsf(L.rdx), fabsf(L.rdy)) + 1.0e-6f;
    return L;
}

__device__ __forceinline__ int narrow_f32(const LinkF &L, float min_x, float min_y, float side) {
    if (L.degenerate) return 2;
    const float ux0 = min_x - L.p0x, ux1 = (min_x + side) - L.p0x;
    const float uy0 = min_y - L.p0y, uy1 = (min_y + side) - L.p0y;
    // corner sign test
    const float a0 = ux0 * L.dy, a1 = ux1 * L.dy, b0 = uy0 * L.dx, b1 = uy1 * L.dx;
    const float c00 = a0 - b0, c01 = a0 - b1, c10 = a1 - b0, c11 = a1 - b1;
    const float K1 = 2.0f * AG_DELTA_P + AG_DELTA_C + 3.6e-7f;
    const float ecr = K1 * (fmaxf(fabsf(ux0), fabsf(ux1)) + fmaxf(fabsf(uy0), fabsf(uy1)) + fabsf(L.dx) + fabsf(L.dy));
    const float cmax = fmaxf(fmaxf(c00, c01), fmaxf(c10, c11)), cmin = fminf(fminf(c00, c01), fminf(c10, c11));
    const bool pos = cmax > ecr, neg = cmin < -ecr;
    if (!(pos && neg)) {
        // all four certainly on one side (or exactly... never: a zero is inside the band) -> miss
        const bool all_certain = fminf(fminf(fabsf(c00), fabsf(c01)), fminf(fabsf(c10), fabsf(c11))) > ecr;
        return all_certain ? 0 : 2;
    }
    // entry / exit parameters, clamped to [-1, 2] (monotone, 1-Lipschitz; thresholds 0 and 1 inside)
    const float tx0 = fminf(fmaxf(ux0 * L.rdx, -1.0f), 2.0f), tx1 = fminf(fmaxf(ux1 * L.rdx, -1.0f), 2.0f);
    const float ty0 = fminf(fmaxf(uy0 * L.rdy, -1.0f), 2.0f), ty1 = fminf(fmaxf(uy1 * L.rdy, -1.0f), 2.0f);
    const float t_in = fmaxf(fminf(tx0, tx1), fminf(ty0, ty1));
    const float t_out = fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1));
    const float e = 1.01f * L.et;   // strictly above the bound, so t > e implies real t > section_eps (1e-10)
    const bool in_sure = (t_in > e) && (t_in < 1.0f - e), out_sure = (t_out > e) && (t_out < 1.0f - e);
    if (in_sure || out_sure) return 1;
    const bool in_out = (t_in < -e) || (t_in > 1.0f + e), out_out = (t_out < -e) || (t_out > 1.0f + e);
    return (in_out && out_out) ? 0 : 2;
}

// ---------------------------------------------------------------- broad phase + narrow phase, one link
// conservative float32 traversal of the bit grid (same scheme as link_exact, wider margins)
__device__ __forceinline__ int link_fast(const GridDev &G, const GridView &V, float p0x, float p0y, float p1x,
                                         float p1y) {
    const float side = (float)G.side, half = (float)G.half, inv_side = (float)G.inv_side;
    const float m = fmaxf(2.0e-6f, 1.0e-3f * side);
    const float Sf = (float)G.S;
    // rows: r = floor((half - y)/side) + 1
    float rf_lo = floorf((half - (fmaxf(p0y, p1y) + m)) * inv_side) + 1.0f;
    float rf_hi = floorf((half - (fminf(p0y, p1y) - m)) * inv_side) + 1.0f;
    if (rf_lo > Sf - 1.0f || rf_hi < 0.0f) return 0;
    const int r_lo = (int)fmaxf(rf_lo, 0.0f), r_hi = (int)fminf(rf_hi, Sf - 1.0f);
    const float dx = p1x - p0x, dy = p1y - p0y;
    const bool clip = (r_hi - r_lo >= 2) && (fabsf(dy) * 64.0f >= fabsf(dx));
    const float inv_dy = clip ? __frcp_rn(dy) : 0.0f;
    const float mx = clip ? m + 5.0e-5f : m;
    int result = 0;
    bool have_link = false;
    LinkF L;
    for (int r = r_lo; r <= r_hi; ++r) {
        float xa = p0x, xb = p1x;
        if (clip) {
            const float yb = half - (float)r * side - m, yt = half - (float)(r - 1) * side + m;
            const float t0 = (yb - p0y) * inv_dy, t1 = (yt - p0y) * inv_dy;
            const float ta = __saturatef(fminf(t0, t1)), tb = __saturatef(fmaxf(t0, t1));
            xa = fmaf(ta, dx, p0x); xb = fmaf(tb, dx, p0x);
        }
        const float cf_lo = floorf((fminf(xa, xb) - mx + half) * inv_side);
        const float cf_hi = floorf((fmaxf(xa, xb) + mx + half) * inv_side);
        if (cf_lo > Sf - 1.0f || cf_hi < 0.0f) continue;
        const int c_lo = (int)fmaxf(cf_lo, 0.0f), c_hi = (int)fminf(cf_hi, Sf - 1.0f);
        for (int w = c_lo >> 5; w <= (c_hi >> 5); ++w) {
            uint32_t mask = 0xFFFFFFFFu;
            if (w == (c_lo >> 5)) mask &= 0xFFFFFFFFu << (c_lo & 31);
            if (w == (c_hi >> 5)) mask &= 0xFFFFFFFFu >> (31 - (c_hi & 31));
            uint32_t word = V.bits[r * G.wpr + w] & mask;
            while (word) {
                const int c = (w << 5) + __ffs(word) - 1;
                word &= word - 1;
                if (!have_link) { L = make_link_f(p0x, p0y, p1x, p1y); have_link = true; }
                const int v = narrow_f32(L, (float)V.min_x[c], (float)V.min_y[r], side);
                if (v == 1) return 1;
                result |= v;          // 0 or 2
            }
        }
    }
    return result;
}

// 0 / 1 certain, 2 undecided
__device__ __forceinline__ int arm_fast(const GridDev &G, const GridView &V, const ArmF &a) {
    const int v1 = link_fast(G, V, 0.0f, 0.0f, a.ex, a.ey);
    if (v1 == 1) return 1;
    const int v2 = link_fast(G, V, a.ex, a.ey, a.gx, a.gy);
    if (v2 == 1) return 1;
    return v1 | v2;
}

// reach test filter, scenario/scene_0.py:129-130 ; 0/1 certain, 2 undecided
__device__ __forceinline__ int reach_fast(const ag_params &P, const ArmF &a) {
    const float eps = (float)P.reach_eps, m = AG_DELTA_P + 2.0e-7f;
    const float ax = fabsf((float)P.target_x - a.gx), ay = fabsf((float)P.target_y - a.gy);
    if (ax > eps + m || ay > eps + m) return 0;
    if (ax < eps - m && ay < eps - m) return 1;
    return 2;
}

// FAST collision_check of a pose (K2/K3): float32 first, EXACT for undecided lanes.
__device__ __forceinline__ bool fast_pose_collides(const ag_params &P, const GridDev &G, const GridView &V, double j1,
                                                   double j2, int &axis) {
    bool ok = true;
    const ArmF a = fast_forward_kinematics(j1, j2, (float)P.link_1, (float)P.link_2, ok);
    const int v = ok ? arm_fast(G, V, a) : 2;
    if (v != 2) return v == 1;
    const Arm A = forward_kinematics(j1, j2, P.link_1, P.link_2);
    int fh = 0;
    return arm_collides<AG_ENGINE_EXACT, false>(G, V, A, P.section_eps, fh, axis);
}

// FAST collision_check when the float64 arm is already known (K1 needs it for its outputs)
__device__ __forceinline__ bool fast_arm_collides(const ag_params &P, const GridDev &G, const GridView &V, const Arm &A,
                                                  int &axis) {
    ArmF a;
    a.ex = (float)A.ex; a.ey = (float)A.ey; a.gx = (float)A.gx; a.gy = (float)A.gy;   // error 6e-8 < AG_DELTA_P
    const int v = arm_fast(G, V, a);
    if (v != 2) return v == 1;
    int fh = 0;
    return arm_collides<AG_ENGINE_EXACT, false>(G, V, A, P.section_eps, fh, axis);
}

// One step's two decisions (collision flag, target reached) for the rollout kernel.
template <int ENGINE>
__device__ __forceinline__ void step_decide(const ag_params &P, const GridDev &G, const GridView &V, double q1,
                                            double q2, bool &hit, bool &reached, int &axis) {
    if constexpr (ENGINE == AG_ENGINE_FAST) {
        bool ok = true;
        const ArmF a = fast_forward_kinematics(q1, q2, (float)P.link_1, (float)P.link_2, ok);
        const int c = ok ? arm_fast(G, V, a) : 2;
        int r;
        if (P.choose_j_tar) r = target_reached_joint(P, q1, q2) ? 1 : 0;
        else r = ok ? reach_fast(P, a) : 2;
        if (c == 2 || r == 2) {                     // undecided lane: the float64 reference arithmetic
            const Arm A = forward_kinematics(q1, q2, P.link_1, P.link_2);
            int fh = 0;
            hit = (c == 2) ? arm_collides<AG_ENGINE_EXACT, false>(G, V, A, P.section_eps, fh, axis) : (c == 1);
            reached = (r == 2) ? target_reached_cart(P, A) : (r == 1);
        } else {
            hit = c == 1; reached = r == 1;
        }
    } else {
        const Arm A = forward_kinematics(q1, q2, P.link_1, P.link_2);
        int fh = 0;
        hit = arm_collides<ENGINE, false>(G, V, A, P.section_eps, fh, axis);
        reached = P.choose_j_tar ? target_reached_joint(P, q1, q2) : target_reached_cart(P, A);
    }
}

}  // namespace agd
