// ag_dense.cu -- K4 for dense / high-resolution maps (BASELINE configs 4 and 5): the rollout loop
// experiment/experiment_0.py:20-34 with a WARP-COOPERATIVE, LOAD-BALANCED collision check.
//
// On these maps a pose check is a walk over hundreds of cells (a 0.4 m link crosses 256 cells of a 1024 x 1024
// grid), an episode ends every second or third step and Scene.reset() needs ~3 candidates, so the pose check IS the
// hot path.  The lane-per-pose traversal of k_rollout_async (ag_kernels.cu) is bound by divergence: every lane walks
// a different number of lines and meets occupied cells at different times; ncu counted 9.7 of 32 threads per
// instruction and 118 k warp-instructions per warp-step on config 4 (profiles/r1_c4_traversal.json).
//
// Here the lanes of a warp still own one environment each (same STEP / RESET state machine, one pose check per lane
// and round), but the traversal work of a round is pooled:
//   * every lane publishes its link's walk (minor-axis line range + the line -> cell-interval map, 12 words) in
//     shared memory; a warp scan turns the line counts into task offsets;
//   * the (pose, line) TASKS are dealt round-robin to the 32 lanes: task j belongs to the owner whose offset range
//     holds j (offsets are monotone, so a lane finds its next owner by stepping forward); one fused multiply-add
//     gives the cell interval on that line (at most two words wide when a link is walked along its minor axis),
//     one or two loads and a mask give the occupied cells in it;
//   * words with occupied cells go to a small per-warp queue (ballot compaction); whenever 32 are waiting the warp
//     runs the float32 narrow phase on them, one word per lane, and ORs "certain hit" / "undecided" into the owner's
//     result word;
//   * link 2 is pooled the same way for the lanes link 1 has not already hit; undecided poses (and angles outside
//     the filter's range) go through the float64 filter and the reference arithmetic on the owner lane.
// Every lane is busy whatever the lines-per-pose distribution: ~300 warp-instructions per pose check on the 256 x 256
// maps instead of ~580 for "one pose at a time, lane = line" (both measured with ncu, round 2).  Decisions are those
// of the FAST engine (same margins, same narrow_f32), so flags, records and statistics equal the EXACT engine's and
// the reference's (tests/test_gpu_parity.py runs this kernel with AG_DENSE_POOLED=1).
#include <cstdlib>

#include "ag_rollout.cuh"

using namespace agd;

void ag_note_launch();

namespace {

constexpr int DB = 256;                     // block size (a block shares one staged grid)

constexpr int DW = DB / 32;
constexpr int QCAP = 64;                    // queue of occupied words per warp

// One link's walk as its owner publishes it: line l (0-based from l_lo) covers cells [lo, hi] with
//   lo = max(min(p(k), p(k+1)) - mm, clamp_lo),  hi = min(max(p(k), p(k+1)) + mm, clamp_hi),  p(k) = pstart + k * s.
// (A short walk -- fewer than three lines -- uses the link's whole cell range on every line: mm = +inf.)
struct __align__(16) WalkS {
    float pstart, s, mm, clamp_lo;
    float clamp_hi, p0x, p0y, p1x;
    float p1y; int l_lo; int swapped; int pad;
};

struct WarpPool {
    WalkS walk[32];
    int start[33];                          // exclusive scan of the line counts; start[32] = total
    int result[32];                         // per owner: 4 = certain hit, 2 = undecided
    uint4 queue[QCAP];                      // (owner, line, first cell of the word, occupied bits)
    unsigned long long planes[32];          // per-env grids in global memory: the owner's plane offset in words
};

// Same geometry as link_fast (ag_fast.cuh): cell coordinates, margins, minor-axis choice, incremental interval.
// Returns the number of lines (0: the link misses the grid).
__device__ __forceinline__ int publish_walk(const GridDev &G, const FastConst &C, float p0x, float p0y, float p1x, float p1y,
                                            WalkS &out) {
    const float mcell = fmaxf(2.0e-6f * C.inv_side, 1.0e-3f) + 0.001f;          // margin in cells
    const int S1 = G.S - 1;
    const float dx = p1x - p0x, dy = p1y - p0y;
    const float xoff = C.half * C.inv_side - 0.5f, roff = C.half * C.inv_side + 0.5f;
    const float ua = fmaf(p0x, C.inv_side, xoff), ub = fmaf(p1x, C.inv_side, xoff);      // columns of the end points
    const float va = fmaf(-p0y, C.inv_side, roff), vb = fmaf(-p1y, C.inv_side, roff);    // rows of the end points
    const bool swapped = fabsf(dx) > fabsf(dy);                                  // lines are columns (transposed bits)
    const float la = swapped ? ua : va, lb = swapped ? ub : vb;
    const float pa = swapped ? va : ua, pb = swapped ? vb : ub;
    int l_lo = round_magic(fminf(la, lb) - mcell), l_hi = round_magic(fmaxf(la, lb) + mcell);
    if (l_lo > S1 || l_hi < 0) return 0;
    l_lo = max(l_lo, 0); l_hi = min(l_hi, S1);
    const float pseg_lo = fminf(pa, pb), pseg_hi = fmaxf(pa, pb);
    const float dl = lb - la, dp = pb - pa;
    const bool tracked = (l_hi - l_lo >= 2) && (fabsf(dl) * 64.0f >= fabsf(dp));
    WalkS W;
    if (tracked) {
        const float s = dp * __frcp_rn(dl);
        W.s = s;
        W.mm = mcell + 5.0e-5f * C.inv_side + fabsf(s) * mcell;
        W.clamp_lo = pseg_lo - W.mm; W.clamp_hi = pseg_hi + W.mm;
        W.pstart = fmaf(((float)l_lo - 0.5f) - la, s, pa);                       // p at the near boundary of line l_lo
    } else {
        W.s = 0.0f; W.pstart = pa; W.mm = __int_as_float(0x7f800000);
        W.clamp_lo = pseg_lo - mcell; W.clamp_hi = pseg_hi + mcell;
    }
    W.p0x = p0x; W.p0y = p0y; W.p1x = p1x; W.p1y = p1y;
    W.l_lo = l_lo; W.swapped = swapped ? 1 : 0; W.pad = 0;
    out = W;
    return l_hi - l_lo + 1;
}

// the float32 narrow phase of one queued word (owner's link against the occupied cells of one word of one line)
__device__ __forceinline__ void narrow_entry(const GridDev &G, const GridView &V, const FastConst &C, WarpPool &wp, uint4 q) {
    const int owner = (int)q.x, line = (int)q.y, wbase = (int)q.z;
    uint32_t word = q.w;
    const WalkS &W = wp.walk[owner];
    const LinkF L = make_link_f(W.p0x, W.p0y, W.p1x, W.p1y, C.side);
    const bool swapped = W.swapped != 0;
    int v = 0;
    while (word) {
        const int pos = wbase + __ffs(word) - 1;
        word &= word - 1;
        const int r = swapped ? pos : line, c = swapped ? line : pos;
        const float mnx = (float)V.min_x[c], mny = (float)V.min_y[r];
        const int vv = narrow_f32(L, mnx, mny, mnx + C.side, mny + C.side);
        v |= vv == 1 ? 4 : vv;
    }
    if (v) atomicOr(&wp.result[owner], v);
}

// All published walks of the warp against the grid.  n: this lane's line count (0: nothing to check).  Returns this
// lane's verdict: 0 / 1 certain, 2 undecided.
__device__ __forceinline__ int pooled_walks(const GridDev &G, const GridView &Vown, const FastConst &C, WarpPool &wp, int n, int lane) {
    // task offsets: exclusive scan of n
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0) return 0;
    wp.start[lane] = incl - n;
    if (lane == 31) wp.start[32] = total;
    wp.result[lane] = 0;
    __syncwarp();
    const int S1 = G.S - 1;
    const bool per_env_planes = !G.stage && G.n_grids > 1;
    int qn = 0;                                                                  // queued words (warp-uniform)
    int owner = 0;
    for (int j0 = 0; j0 < total; j0 += 32) {                                     // warp-uniform trip count
        const int j = j0 + lane;
        uint32_t m0 = 0, m1 = 0;
        int w0 = 0, w1 = 0, line = 0;
        bool wide = false;
        if (j < total) {
            while (wp.start[owner + 1] <= j) ++owner;                            // offsets are monotone: step forward
            const WalkS &W = wp.walk[owner];
            const int k = j - wp.start[owner];
            line = W.l_lo + k;
            const float pprev = fmaf((float)k, W.s, W.pstart), pcur = fmaf((float)k + 1.0f, W.s, W.pstart);
            const float lo = fmaxf(fminf(pprev, pcur) - W.mm, W.clamp_lo), hi = fminf(fmaxf(pprev, pcur) + W.mm, W.clamp_hi);
            const int p_lo = max(round_magic(lo), 0), p_hi = min(round_magic(hi), S1);
            if (p_lo <= p_hi) {
                w0 = p_lo >> 5; w1 = p_hi >> 5;
                const uint32_t *base = W.swapped ? Vown.bits_t : Vown.bits;
                if (per_env_planes) base = (W.swapped ? G.bits_t : G.bits) + wp.planes[owner];
                const uint32_t *lp = base + line * G.wpr;
                const uint32_t mlo = 0xFFFFFFFFu << (p_lo & 31), mhi = 0xFFFFFFFFu >> (31 - (p_hi & 31));
                if (w1 == w0) {
                    m0 = lp[w0] & mlo & mhi;
                } else {                                                         // the interval straddles a word boundary
                    m0 = lp[w0] & mlo;
                    m1 = lp[w1] & mhi;
                    for (int w = w0 + 1; w < w1; ++w) wide |= lp[w] != 0;        // (minor-axis walk: never more than two words)
                }
            }
            if (wide) atomicOr(&wp.result[owner], 2);                            // leave it to the float64 engine
        }
        // occupied words -> queue (ballot compaction), first words then second words
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t word = half ? m1 : m0;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, word != 0);
            if (mask) {
                if (word != 0) wp.queue[qn + __popc(mask & ((1u << lane) - 1u))] = make_uint4((uint32_t)owner, (uint32_t)line, (uint32_t)((half ? w1 : w0) << 5), word);
                qn += __popc(mask);
                __syncwarp();
                if (qn >= 32) {                                                  // a full warp of narrow-phase work
                    qn -= 32;
                    narrow_entry(G, Vown, C, wp, wp.queue[qn + lane]);
                    __syncwarp();
                }
            }
        }
    }
    if (qn > 0) {
        if (lane < qn) narrow_entry(G, Vown, C, wp, wp.queue[lane]);
    }
    __syncwarp();
    const int r = wp.result[lane];
    return (r & 4) ? 1 : (r & 2);
}

template <bool HAS_ACT, bool HAS_RESET_U, bool RECORD>
__global__ void __launch_bounds__(DB, 3)
k_rollout_coop(const __grid_constant__ ag_params P, const __grid_constant__ GridDev G, const __grid_constant__ FastConst C,
               const __grid_constant__ RolloutDev A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ FastList s_fl;
    __shared__ unsigned long long s_acc[AG_ST_COUNT + AG_DIAG_COUNT];
    __shared__ WarpPool s_pool[DW];
    if (threadIdx.x < AG_ST_COUNT + AG_DIAG_COUNT) s_acc[threadIdx.x] = 0;
    const BlockCtx B = block_prologue<AG_ENGINE_FAST>(G, A.env_id0, A.n, smem, &s_fl);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = e < A.n;
    const int64_t ec = active ? e : A.n - 1;
    long long loc[AG_ST_COUNT];
#pragma unroll
    for (int i = 0; i < AG_ST_COUNT; ++i) loc[i] = 0;
    // per-lane grid view (block-uniform when the grid is staged); the owner's view is broadcast with its pose
    const GridView Vown = G.stage ? B.V : thread_view(G, smem, A.env_id0 + ec);
    if (!G.stage && G.n_grids > 1) s_pool[threadIdx.x >> 5].planes[lane] = (unsigned long long)(grid_of_env(G, A.env_id0 + ec) * G.stride_words);
    double q1 = A.j1[ec], q2 = A.j2[ec];
    float rw = A.reward[ec];
    uint32_t fl = A.flags[ec], el = A.ep_len[ec], rc = A.reset_ctr[ec];
    const uint32_t sc0 = A.step_ctr[ec];
    if (active) A.step_ctr[e] = sc0 + (uint32_t)A.K;
    const uint64_t gid = (uint64_t)(A.env_id0 + ec);
    const double *ru = HAS_RESET_U ? A.reset_u + ec * A.R * 2 : nullptr;
    const double *tgt = A.targets ? A.targets + 2 * ec : nullptr;
    int t = active ? 0 : A.K, tries = 0;
    bool resetting = false;
    for (;;) {
        // ---------------- phase A, per lane: the next pose of this lane's state machine
        const bool busy = t < A.K || resetting;
        if (!__any_sync(0xFFFFFFFFu, busy)) break;
        bool stuck = false;
        if (busy) {
            if (!resetting) {                                                    // ---- STEP: experiment_0.py:21-22
                double d1, d2;
                if (HAS_ACT) {
                    const float2 a = __ldcs(reinterpret_cast<const float2 *>(A.actions) + (int64_t)t * A.row_stride + e);
                    d1 = (double)a.x; d2 = (double)a.y;
                } else {
                    double u0, u1;
                    philox_uniform2(A.seed, gid, sc0 + (uint32_t)t, 0u, u0, u1);
                    d1 = __dmul_rn(__dsub_rn(u0, 0.5), P.action_scale);          // scene_0.py:84
                    d2 = __dmul_rn(__dsub_rn(u1, 0.5), P.action_scale);          // :85
                }
                q1 = __dadd_rn(q1, d1); q2 = __dadd_rn(q2, d2);                  // two_joint_robot.py:71-72
            } else if (tries >= P.max_reset_tries || (HAS_RESET_U && rc >= (uint32_t)A.R)) {
                stuck = true;                                                    // give up: keep the last candidate
            } else {                                                             // ---- RESET: scene_0.py:179-181
                double u0, u1;
                if (HAS_RESET_U) {
                    const double2 u = reinterpret_cast<const double2 *>(ru)[rc];
                    u0 = u.x; u1 = u.y;
                } else {
                    philox_uniform2(A.seed, gid, rc, 1u, u0, u1);
                }
                ++rc; ++tries;
                q1 = __dmul_rn(__dmul_rn(u0, 3.141592653589793), 2.0);           // scene_0.py:180  rand()*pi*2.0
                q2 = __dmul_rn(__dmul_rn(u1, 3.141592653589793), 2.0);           // :181
            }
        }
        bool ok = false;
        ArmF a;
        a.ex = a.ey = a.gx = a.gy = 0.0f;
        if (busy && !stuck) a = fast_forward_kinematics(q1, q2, C, ok);
        // ---------------- phase B, whole warp: the traversal work of all 32 poses, pooled
        int c = (busy && !stuck && !ok) ? 2 : 0;                                 // angles outside the filter's range: float64
        {
            WarpPool &wp = s_pool[threadIdx.x >> 5];
            const bool check = busy && !stuck && ok;
            int n = 0;
            if (check) n = publish_walk(G, C, 0.0f, 0.0f, a.ex, a.ey, wp.walk[lane]);              // link 1: origin -> elbow
            const int v1 = pooled_walks(G, Vown, C, wp, n, lane);
            __syncwarp();
            n = 0;
            if (check && v1 != 1) n = publish_walk(G, C, a.ex, a.ey, a.gx, a.gy, wp.walk[lane]);   // link 2: elbow -> end effector
            const int v2 = pooled_walks(G, Vown, C, wp, n, lane);
            __syncwarp();
            if (check) c = (v1 == 1 || v2 == 1) ? 1 : (v1 | v2);
        }
        // ---------------- phase C, per lane: the decision and its consequences
        if (busy) {
            int d = 0;
            if (!stuck) {
                int r = 0;
                if (!resetting) {
                    if (P.choose_j_tar) r = target_reached_joint(P, q1, q2) ? 1 : 0;
                    else if (!ok) r = 2;
                    else {
                        float tx = C.tx, ty = C.ty;
                        if (tgt != nullptr) { tx = (float)tgt[0]; ty = (float)tgt[1]; }
                        r = reach_fast_at(C, a, tx, ty);
                    }
                }
                if (c == 2 || r == 2) {
                    double txd = P.target_x, tyd = P.target_y;
                    if (tgt != nullptr) { txd = tgt[0]; tyd = tgt[1]; }
                    d = cold_exact_decide_at(P, G, Vown, nullptr, q1, q2, c, r, txd, tyd);
                } else {
                    d = c | (r << 1);
                }
            }
            loc[AG_ST_AXIS_ALIGNED] += d >> 2;
            if (!resetting) {
                if (d & 1) { rw = (float)P.reward_collision; fl |= AG_FLAG_COLLISION; }   // scene_0.py:95-97
                if (d & 2) { rw = (float)P.reward_reach; fl |= AG_FLAG_DONE; }            // :98-100
                store_record<RECORD>(A, (int64_t)t * A.row_stride + e, q1, q2, rw, fl);    // experiment_0.py:23-25
                emit_event(A, e, t, rw, fl);
                ++el; ++t;
                if (fl) {                                                        // experiment_0.py:30-34
                    ++loc[AG_ST_EPISODES];
                    loc[AG_ST_COLLISIONS] += (fl & AG_FLAG_COLLISION) ? 1 : 0;
                    loc[AG_ST_SUCCESSES] += (fl & AG_FLAG_DONE) ? 1 : 0;
                    loc[AG_ST_EP_LEN_SUM] += el;
                    loc[AG_ST_RETURN_MILLI] += __float2int_rn(rw * 1e-3f);
                    rw = 0.0f; fl = 0; el = 0;                                   // scene_0.py:111-113
                    resetting = (d & 1) != 0;                                    // random_valid_pose() only while colliding
                    tries = 0;
                }
            } else if (stuck) {
                ++loc[AG_ST_STUCK_RESETS];
                resetting = false;
            } else if (!(d & 1)) {
                resetting = false;                                               // candidate accepted
            }
        }
        __syncwarp();
    }
    if (active) {
        loc[AG_ST_ENV_STEPS] = A.K;
        A.j1[e] = q1; A.j2[e] = q2; A.reward[e] = rw; A.flags[e] = (uint8_t)fl;
        A.ep_len[e] = el; A.reset_ctr[e] = rc;
    }
    block_accumulate_stats(loc, A.stats, s_acc);
}

template <typename Kern>
ag_status set_smem_d(Kern k, size_t smem) {
    if (smem == 0) return AG_OK;
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, k);
    if (e != cudaSuccess) return (ag_status)e;
    if (smem + at.sharedSizeBytes > 48 * 1024 && (int)smem > at.maxDynamicSharedSizeBytes) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (ag_status)e;
    }
    return AG_OK;
}

template <bool HA, bool HR, bool REC>
ag_status launch_c(const ag_params &P, const GridDev &G, const RolloutDev &A, size_t smem, cudaStream_t s) {
    auto k = k_rollout_coop<HA, HR, REC>;
    ag_status st = set_smem_d(k, smem);
    if (st) return st;
    k<<<(unsigned)((A.n + DB - 1) / DB), DB, smem, s>>>(P, G, make_fast_const(P, G), A);
    ag_note_launch();
    return (ag_status)cudaGetLastError();
}

}  // namespace

// Opt-in (AG_DENSE_POOLED=1): measured on a B200 (round 2, valid maps) it equals the lane-asynchronous kernel on the
// 256 x 256 per-batch maps (62.6 vs 58.2 ms per 2^20 x 64 launch) and loses on the 1024 x 1024 map (365 vs 256 ms):
// ~105 warp-instructions per 32 line tasks and the per-round scans outweigh the balance it buys.  DESIGN.md section 4.
// It needs the transposed bit planes (every link is walked along its long axis, rows or columns).
bool rollout_coop_applies(const ag_params &P, const GridDev &G, const RolloutDev &A) {
    static const bool pooled = std::getenv("AG_DENSE_POOLED") != nullptr;
    (void)P; (void)A;
    return pooled && G.bits_t != nullptr;
}

ag_status launch_rollout_coop(const ag_params &P, const GridDev &G, const RolloutDev &A, size_t smem, cudaStream_t s) {
    const bool ha = A.actions != nullptr, hr = A.reset_u != nullptr, rec = A.rec_j1 != nullptr;
#define AG_RC(HA, HR, REC) return launch_c<HA, HR, REC>(P, G, A, smem, s)
    if (ha) { if (hr) { if (rec) AG_RC(true, true, true); else AG_RC(true, true, false); }
              else    { if (rec) AG_RC(true, false, true); else AG_RC(true, false, false); } }
    else    { if (hr) { if (rec) AG_RC(false, true, true); else AG_RC(false, true, false); }
              else    { if (rec) AG_RC(false, false, true); else AG_RC(false, false, false); } }
#undef AG_RC
}
