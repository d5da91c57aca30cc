/*
 * ag_oracle.c -- CPU restatement of abstract_gym's scene_0 step/reset hot path (float64).
 *
 * TEST INFRASTRUCTURE ONLY (see ag_oracle.h).  Plain C, one IEEE rounding per operation, the
 * reference's operation order.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * sin/cos are glibc libm, which numpy's float64 np.sin/np.cos equal bit-for-bit on this image
 * (re-checked by tests/test_oracle_golden.py::test_libm_matches_numpy).
 */
#include "ag_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void ago_default_params(ago_params *p) {
    p->link_1 = 0.4;              /* robot/two_joint_robot.py:12 */
    p->link_2 = 0.3;              /* robot/two_joint_robot.py:13 */
    p->target_x = -0.2;           /* scenario/scene_0.py:17 */
    p->target_y = -0.3;
    p->target_j1 = 1.1;           /* scenario/scene_0.py:30 */
    p->target_j2 = -0.2;
    p->reach_eps = 2e-3;          /* scenario/scene_0.py:122 */
    p->section_eps = 1e-10;       /* utils/collision_checker.py:81 */
    p->reward_collision = -1e3;   /* scenario/scene_0.py:96 */
    p->reward_reach = 1e4;        /* scenario/scene_0.py:99 */
    p->action_scale = 0.1;        /* scenario/scene_0.py:78 */
    p->choose_j_tar = 0;          /* scenario/scene_0.py:31 */
    p->max_reset_tries = 64;
}

int ago_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ grid -> squares (G1) */

/* environment/occupancy_grid.py:59-67: coord * E / (S-1) - E/2.0 ; y *= -1 ; max = min + side */
static void cell_to_square(int32_t col, int32_t row, int32_t S, double E, ago_square *s) {
    double side = E / (double)(S - 1);                       /* occupancy_grid.py:28 */
    double half = E / 2.0;
    double x = ((double)col * E) / (double)(S - 1) - half;   /* occupancy_grid.py:59-60 */
    double y = ((double)row * E) / (double)(S - 1) - half;
    y = y * -1.0;                                            /* occupancy_grid.py:64 (keeps -0.0) */
    s->min_x = x;
    s->min_y = y;
    s->max_x = x + side;                                     /* occupancy_grid.py:67 */
    s->max_y = y + side;
}

void ago_cells_to_squares(const int32_t *cols, const int32_t *rows, int64_t m, int32_t S,
                          double env_size, ago_square *squares) {
    for (int64_t i = 0; i < m; ++i) cell_to_square(cols[i], rows[i], S, env_size, &squares[i]);
}

int64_t ago_grid_squares(const uint8_t *occ, int32_t S, double env_size,
                         ago_square *squares, int32_t *cell_index, int64_t cap) {
    int64_t m = 0;
    /* np.where(occ != 0) enumerates row-major: occupancy_grid.py:37,90 */
    for (int32_t r = 0; r < S; ++r)
        for (int32_t c = 0; c < S; ++c)
            if (occ[(int64_t)r * S + c]) {
                if (m < cap) {
                    if (squares) cell_to_square(c, r, S, env_size, &squares[m]);
                    if (cell_index) cell_index[m] = r * S + c;
                }
                ++m;
            }
    return m;
}

/* ------------------------------------------------------------------ line + predicate (L1,P1-P3) */

void ago_line_function(double p0x, double p0y, double p1x, double p1y, double *a, double *b, double *c) {
    if (p0x == p1x) {            /* utils/geometry.py:19-23 vertical */
        *b = 0.0; *a = 1.0; *c = -p0x;
        return;
    }
    if (p0y == p1y) {            /* utils/geometry.py:24-28 horizontal */
        *a = 0.0; *b = 1.0; *c = -p0y;
        return;
    }
    *a = 1.0 / (p1x - p0x);                              /* utils/geometry.py:29 */
    *b = -1.0 / (p1y - p0y);                             /* utils/geometry.py:30 */
    *c = p0y / (p1y - p0y) - p0x / (p1x - p0x);          /* utils/geometry.py:31 */
}

void ago_corner_values(double a, double b, double c, const ago_square *s, double v[4]) {
    v[0] = a * s->min_x + b * s->min_y + c;              /* utils/collision_checker.py:27 */
    v[1] = a * s->min_x + b * s->max_y + c;              /* :28 */
    v[2] = a * s->max_x + b * s->min_y + c;              /* :29 */
    v[3] = a * s->max_x + b * s->max_y + c;              /* :30 */
}

static double dmin(double x, double y) { return x < y ? x : y; }
static double dmax(double x, double y) { return x > y ? x : y; }

static int cmp_double(const void *pa, const void *pb) {
    double x = *(const double *)pa, y = *(const double *)pb;
    return (x > y) - (x < y);
}

/* utils/collision_checker.py:48-85 */
static int check_sections(double a, double b, double c, double p0x, double p0y, double p1x, double p1y,
                          const ago_square *s, double eps, int64_t *axis_aligned) {
    if (a == 0.0) {
        /* :59-63 reads Line.max_x/min_x which do not exist (AttributeError in the reference).
         * Defined here as the evident intent: x-interval overlap.  Counted. */
        if (axis_aligned) ++*axis_aligned;
        if (dmax(p0x, p1x) < s->min_x || dmin(p0x, p1x) > s->max_x) return 0;
        return 1;
    }
    if (b == 0.0) {              /* :64-68, same remark, y-interval overlap */
        if (axis_aligned) ++*axis_aligned;
        if (dmax(p0y, p1y) < s->min_y || dmin(p0y, p1y) > s->max_y) return 0;
        return 1;
    }
    double xs[4];
    xs[0] = (-c - b * s->min_y) / a;                     /* :69 sx_y_min */
    xs[1] = (-c - b * s->max_y) / a;                     /* :71 sx_y_max */
    xs[2] = s->min_x;                                    /* :74 (sy_x_min :73 is never used) */
    xs[3] = s->max_x;                                    /* :76 */
    qsort(xs, 4, sizeof(double), cmp_double);            /* :77-78 sort by x */
    double lam1 = (xs[1] - p0x) / (p1x - p0x);           /* :79, :90-91 */
    double lam2 = (xs[2] - p0x) / (p1x - p0x);           /* :80 */
    if ((1.0 > lam1 && lam1 > eps) || (1.0 > lam2 && lam2 > eps)) return 1;   /* :82 */
    return 0;
}

int ago_segment_square(double p0x, double p0y, double p1x, double p1y, const ago_square *s,
                       double section_eps, int64_t *axis_aligned) {
    double a, b, c, v[4];
    ago_line_function(p0x, p0y, p1x, p1y, &a, &b, &c);   /* utils/collision_checker.py:21 */
    ago_corner_values(a, b, c, s, v);
    int pos = 0, neg = 0;                                /* utils/collision_checker.py:41-43 */
    for (int i = 0; i < 4; ++i) { pos += v[i] > 0.0; neg += v[i] < 0.0; }
    if (pos > 0 && neg > 0) return check_sections(a, b, c, p0x, p0y, p1x, p1y, s, section_eps, axis_aligned);
    return 0;
}

double ago_segment_square_margin(double p0x, double p0y, double p1x, double p1y, const ago_square *s,
                                 double section_eps, double link_len) {
    double a, b, c, v[4];
    ago_line_function(p0x, p0y, p1x, p1y, &a, &b, &c);
    ago_corner_values(a, b, c, s, v);
    double h = hypot(a, b), m = INFINITY;
    int pos = 0, neg = 0;
    for (int i = 0; i < 4; ++i) {
        m = dmin(m, fabs(v[i]) / h);
        pos += v[i] > 0.0; neg += v[i] < 0.0;
    }
    if (!(pos > 0 && neg > 0) || a == 0.0 || b == 0.0) return m;
    double xs[4] = { (-c - b * s->min_y) / a, (-c - b * s->max_y) / a, s->min_x, s->max_x };
    qsort(xs, 4, sizeof(double), cmp_double);
    for (int i = 1; i <= 2; ++i) {
        double lam = (xs[i] - p0x) / (p1x - p0x);
        m = dmin(m, fabs(lam - 1.0) * link_len);
        m = dmin(m, fabs(lam - section_eps) * link_len);
    }
    /* a tie between candidate xs changes which two are "the middle two" */
    m = dmin(m, dmin(dmin(fabs(xs[1] - xs[0]), fabs(xs[2] - xs[1])), fabs(xs[3] - xs[2])));
    return m;
}

/* ------------------------------------------------------------------ FK (F1,F2) */

void ago_forward_kinematics(double j1, double j2, double l1, double l2,
                            double *elbow_x, double *elbow_y, double *ee_x, double *ee_y) {
    *elbow_x = cos(j1) * l1;                             /* robot/two_joint_robot.py:45 */
    *elbow_y = sin(j1) * l1;                             /* :46 */
    *ee_x = cos(j1) * l1 + cos(j2) * l2;                 /* :36 */
    *ee_y = sin(j1) * l1 + sin(j2) * l2;                 /* :37 */
}

/* ------------------------------------------------------------------ IK / joint-space motion */

/* robot/two_joint_robot.py:74-113.  sol = (j1_1, j2_1, j1_2, j2_2); returns 1 if the target is in the reachable
 * annulus |l1-l2| < r <= l1+l2 (and r != 0), else 0 with sol untouched.  corrected != 0 replaces
 * alpha = arccos(x/r) (which drops the sign of y) by atan2(y, x).
 * Pinned to the reference within 4 ulp, not bit-exact: numpy's arccos is not glibc's acos (tests/test_oracle_golden.py). */
/* the reference squares with pow(x, 2) (Python float / numpy scalar __pow__ -> libm pow); gcc would fold a literal
 * exponent into x*x, which is not always the same double, so the exponent is kept opaque */
static double pw2(double x) {
    volatile double two = 2.0;
    return pow(x, two);
}

int ago_inverse_kinematics(double tx, double ty, double l1, double l2, int corrected, double *sol) {
    const double R = l1 + l2;                                           /* :80 total_length() */
    const double r = fabs(l1 - l2);                                     /* :81 */
    const double radius = sqrt(pw2(tx) + pw2(ty));                      /* :82 */
    if (!(r < radius && radius <= R)) return 0;                         /* :83-86 */
    if (radius == 0) return 0;                                          /* :96-98 */
    const double cos_theta = (pw2(radius) + pw2(l1) - pw2(l2)) / (2.0 * l1 * radius);   /* :99 */
    const double theta = acos(cos_theta);                               /* :100 */
    const double alpha = corrected ? atan2(ty, tx) : acos(tx / radius); /* :101-102 */
    const double j1_1 = alpha - theta, j1_2 = alpha + theta;            /* :103-104 */
    const double cos_beta = (pw2(l1) + pw2(l2) - pw2(radius)) / (2.0 * l1 * l2);        /* :105 */
    sol[0] = j1_1;
    sol[1] = M_PI - acos(cos_beta) + j1_1;                              /* :106 */
    sol[2] = j1_2;
    sol[3] = j1_2 - (M_PI - acos(cos_beta));                            /* :107 */
    return 1;
}

/* robot/two_joint_robot.py:49-62: `steps` equal increments alpha*(target - init), accumulated one by one */
void ago_move_to_joint_pose(double *j1, double *j2, double t1, double t2, int32_t steps) {
    const double alpha = 1.0 / steps;                                   /* :57 */
    const double i1 = *j1, i2 = *j2;
    for (int32_t i = 0; i < steps; ++i) {
        *j1 += alpha * (t1 - i1);                                       /* :61 */
        *j2 += alpha * (t2 - i2);                                       /* :62 */
    }
}

/* ------------------------------------------------------------------ scene (C1,R1,ST,RS) */

int ago_collision_check(const ago_params *p, double j1, double j2, const ago_square *sq,
                        const int32_t *cell_index, int64_t m, int32_t *first_hit, double *margin,
                        int64_t *axis_aligned) {
    double ex, ey, gx, gy;
    ago_forward_kinematics(j1, j2, p->link_1, p->link_2, &ex, &ey, &gx, &gy);
    int want_all = (first_hit != NULL) || (margin != NULL);
    int hit = 0;
    int32_t fh = -1;
    double mg = INFINITY;
    for (int64_t i = 0; i < m; ++i) {                    /* scenario/scene_0.py:67 */
        /* l1 = (0,0)->elbow :65 ; l2 = elbow->EE :66 */
        int c1 = ago_segment_square(0.0, 0.0, ex, ey, &sq[i], p->section_eps, axis_aligned);
        int c2 = 0;
        if (!c1 || want_all)
            c2 = ago_segment_square(ex, ey, gx, gy, &sq[i], p->section_eps, axis_aligned);
        if (margin) {
            mg = dmin(mg, ago_segment_square_margin(0.0, 0.0, ex, ey, &sq[i], p->section_eps, p->link_1));
            mg = dmin(mg, ago_segment_square_margin(ex, ey, gx, gy, &sq[i], p->section_eps, p->link_2));
        }
        if (c1 || c2) {
            hit = 1;
            if (!want_all) return 1;                     /* early return :69-70,:73-74 */
            int32_t ci = cell_index ? cell_index[i] : (int32_t)i;
            if (fh < 0 || ci < fh) fh = ci;
        }
    }
    if (first_hit) *first_hit = fh;
    if (margin) *margin = mg;
    return hit;
}

int ago_target_reached(const ago_params *p, double j1, double j2) {
    if (p->choose_j_tar)                                 /* scenario/scene_0.py:123-127 */
        return fabs(j1 - p->target_j1) < p->reach_eps && fabs(j2 - p->target_j2) < p->reach_eps;
    double ex, ey, gx, gy;
    ago_forward_kinematics(j1, j2, p->link_1, p->link_2, &ex, &ey, &gx, &gy);
    return fabs(p->target_x - gx) < p->reach_eps && fabs(p->target_y - gy) < p->reach_eps; /* :129-130 */
}

void ago_step_batch(const ago_params *p, const ago_square *sq, const int32_t *cell_index, int64_t m,
                    double *j1, double *j2, const double *actions, double *reward, uint8_t *flags,
                    double *ee_xy, double *dist_xy, int32_t *first_hit, double *margin,
                    int64_t *axis_aligned, int64_t n) {
    int64_t aa = 0;
#pragma omp parallel for reduction(+ : aa) schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        j1[e] += actions[2 * e];                         /* robot/two_joint_robot.py:71 */
        j2[e] += actions[2 * e + 1];                     /* :72 */
        int32_t fh = -1;
        double mg = 0.0;
        int hit = ago_collision_check(p, j1[e], j2[e], sq, cell_index, m, first_hit ? &fh : NULL,
                                      margin ? &mg : NULL, &aa);
        if (hit) {                                       /* scenario/scene_0.py:95-97 */
            reward[e] = p->reward_collision;
            flags[e] |= AGO_FLAG_COLLISION;
        }
        if (ago_target_reached(p, j1[e], j2[e])) {       /* :98-100 */
            reward[e] = p->reward_reach;
            flags[e] |= AGO_FLAG_DONE;
        }
        if (first_hit) first_hit[e] = fh;
        if (margin) margin[e] = mg;
        if (ee_xy || dist_xy) {
            double ex, ey, gx, gy;
            ago_forward_kinematics(j1[e], j2[e], p->link_1, p->link_2, &ex, &ey, &gx, &gy);
            if (ee_xy) { ee_xy[2 * e] = gx; ee_xy[2 * e + 1] = gy; }
            if (dist_xy) {
                dist_xy[2 * e] = fabs(p->target_x - gx);
                dist_xy[2 * e + 1] = fabs(p->target_y - gy);
            }
        }
    }
    if (axis_aligned) *axis_aligned += aa;
}

void ago_collision_batch(const ago_params *p, const ago_square *sq, const int32_t *cell_index, int64_t m,
                         const double *j1, const double *j2, uint8_t *hit, int32_t *first_hit,
                         double *margin, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        int32_t fh = -1;
        double mg = 0.0;
        hit[e] = (uint8_t)ago_collision_check(p, j1[e], j2[e], sq, cell_index, m,
                                              first_hit ? &fh : NULL, margin ? &mg : NULL, NULL);
        if (first_hit) first_hit[e] = fh;
        if (margin) margin[e] = mg;
    }
}

/* ------------------------------------------------------------------ experiment_0 loop, one env */

int64_t ago_experiment_loop(const ago_params *p, const ago_square *sq, int64_t m,
                            double *j1, double *j2, const double *draws, int64_t n_draws,
                            int64_t steps, double *rec, int64_t *n_resets, int64_t *reset_steps,
                            int64_t reset_cap) {
    int64_t k = 0, nres = 0;
    double reward = 0.0;
    int done = 0, coll = 0;
    /* s.random_valid_pose()  experiment_0.py:16 -> scene_0.py:179-181 */
    while (ago_collision_check(p, *j1, *j2, sq, NULL, m, NULL, NULL, NULL)) {
        if (k + 2 > n_draws) return -1;
        *j1 = draws[k++] * M_PI * 2.0;                   /* scenario/scene_0.py:180 */
        *j2 = draws[k++] * M_PI * 2.0;                   /* :181 */
    }
    for (int64_t i = 0; i < steps; ++i) {                /* experiment_0.py:20 */
        if (k + 2 > n_draws) return -1;
        double d1 = (draws[k++] - 0.5) * p->action_scale;    /* scenario/scene_0.py:84 */
        double d2 = (draws[k++] - 0.5) * p->action_scale;    /* :85 */
        *j1 += d1;                                       /* step: scene_0.py:94 */
        *j2 += d2;
        if (ago_collision_check(p, *j1, *j2, sq, NULL, m, NULL, NULL, NULL)) {
            reward = p->reward_collision; coll = 1;
        }
        if (ago_target_reached(p, *j1, *j2)) {
            reward = p->reward_reach; done = 1;
        }
        double *r = rec + 7 * i;                         /* experiment_0.py:23-25 */
        r[0] = *j1; r[1] = *j2; r[2] = d1; r[3] = d2; r[4] = reward; r[5] = done; r[6] = coll;
        if (done || coll) {                              /* experiment_0.py:30-34 */
            if (nres < reset_cap && reset_steps) reset_steps[nres] = i;
            ++nres;
            while (ago_collision_check(p, *j1, *j2, sq, NULL, m, NULL, NULL, NULL)) {
                if (k + 2 > n_draws) return -1;
                *j1 = draws[k++] * M_PI * 2.0;
                *j2 = draws[k++] * M_PI * 2.0;
            }
            coll = 0; done = 0; reward = 0.0;            /* scenario/scene_0.py:111-113 */
        }
    }
    if (n_resets) *n_resets = nres;
    return k;
}

/* ------------------------------------------------------------------ Philox4x32-10 */

void ago_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void ago_philox_uniform2(uint64_t seed, uint64_t env_id, uint32_t draw, uint32_t stream, double u[2]) {
    uint32_t ctr[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), draw, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t w[4];
    ago_philox4x32_10(ctr, key, w);
    /* numpy legacy genrand_res53: (a>>5, b>>6) -> (a*2^26 + b) / 2^53 */
    u[0] = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
    u[1] = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) / 9007199254740992.0;
}

/* ------------------------------------------------------------------ batched K-step rollout */

int ago_rollout(const ago_params *p, const ago_rollout_args *a) {
    if (a->n < 0 || a->K < 0 || a->n_grids < 1 || a->envs_per_grid < 1) return -1;
    int64_t tot[AGO_ST_COUNT];
    memset(tot, 0, sizeof tot);
    int nthreads = a->threads > 0 ? a->threads : ago_num_threads();
    (void)nthreads;
#pragma omp parallel num_threads(nthreads)
    {
        int64_t st[AGO_ST_COUNT];
        memset(st, 0, sizeof st);
#pragma omp for schedule(static)
        for (int64_t e = 0; e < a->n; ++e) {
            uint64_t gid = (uint64_t)(a->env_id0 + e);
            int64_t g = (int64_t)((gid / (uint64_t)a->envs_per_grid) % (uint64_t)a->n_grids);
            const ago_square *sq = a->sq + a->sq_offsets[g];
            int64_t m = a->sq_offsets[g + 1] - a->sq_offsets[g];
            double j1 = a->j1[e], j2 = a->j2[e];
            double reward = a->reward[e];
            uint8_t flags = a->flags[e];
            uint32_t sc = a->step_ctr[e], rc = a->reset_ctr[e], el = a->ep_len[e];
            for (int32_t t = 0; t < a->K; ++t) {
                double d1, d2;
                if (a->actions_f32) {
                    const float *ap = a->actions_f32 + ((int64_t)t * a->n + e) * 2;
                    d1 = (double)ap[0]; d2 = (double)ap[1];
                } else {
                    double u[2];
                    ago_philox_uniform2(a->seed, gid, sc, 0u, u);
                    d1 = (u[0] - 0.5) * p->action_scale;     /* scenario/scene_0.py:84-85 */
                    d2 = (u[1] - 0.5) * p->action_scale;
                }
                ++sc;
                j1 += d1; j2 += d2;                          /* two_joint_robot.py:71-72 */
                if (ago_collision_check(p, j1, j2, sq, NULL, m, NULL, NULL, &st[AGO_ST_AXIS_ALIGNED])) {
                    reward = p->reward_collision; flags |= AGO_FLAG_COLLISION;
                }
                if (ago_target_reached(p, j1, j2)) {
                    reward = p->reward_reach; flags |= AGO_FLAG_DONE;
                }
                if (a->rec_j1) {
                    int64_t o = (int64_t)t * a->n + e;
                    a->rec_j1[o] = (float)j1; a->rec_j2[o] = (float)j2;
                    a->rec_reward[o] = (float)reward; a->rec_flags[o] = flags;
                }
                ++el; ++st[AGO_ST_ENV_STEPS];
                if (flags) {                                 /* experiment_0.py:30-34 */
                    ++st[AGO_ST_EPISODES];
                    st[AGO_ST_COLLISIONS] += (flags & AGO_FLAG_COLLISION) != 0;
                    st[AGO_ST_SUCCESSES] += (flags & AGO_FLAG_DONE) != 0;
                    st[AGO_ST_EP_LEN_SUM] += el;
                    st[AGO_ST_RETURN_MILLI] += llround(reward / 1000.0);
                    el = 0;
                    int tries = 0;                           /* scenario/scene_0.py:179-181, bounded */
                    while (ago_collision_check(p, j1, j2, sq, NULL, m, NULL, NULL, &st[AGO_ST_AXIS_ALIGNED])) {
                        double u[2];
                        if (tries >= p->max_reset_tries || (a->reset_u && (int64_t)rc >= a->R)) {
                            ++st[AGO_ST_STUCK_RESETS];
                            break;
                        }
                        if (a->reset_u) {
                            const double *up = a->reset_u + ((int64_t)e * a->R + rc) * 2;
                            u[0] = up[0]; u[1] = up[1];
                        } else {
                            ago_philox_uniform2(a->seed, gid, rc, 1u, u);
                        }
                        ++rc; ++tries;
                        j1 = u[0] * M_PI * 2.0;              /* (u*pi)*2.0 */
                        j2 = u[1] * M_PI * 2.0;
                    }
                    reward = 0.0; flags = 0;                 /* scenario/scene_0.py:111-113 */
                }
            }
            a->j1[e] = j1; a->j2[e] = j2;
            a->reward[e] = (float)reward; a->flags[e] = flags;
            a->step_ctr[e] = sc; a->reset_ctr[e] = rc; a->ep_len[e] = el;
        }
#pragma omp critical
        for (int i = 0; i < AGO_ST_COUNT; ++i) tot[i] += st[i];
    }
    if (a->stats)
        for (int i = 0; i < AGO_ST_COUNT; ++i) a->stats[i] += tot[i];
    return 0;
}
