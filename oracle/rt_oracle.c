/*
 * rt_oracle.c — CPU restatement of the ppo-radiotherapy environment-step path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ppo-radiotherapy_b200/ may link,
 * import or call this file; it exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs have a checker that can
 * travel to the GPU box (the Python reference under /root/reference cannot).
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors for
 * this path, so this restatement is pinned against outputs of the
 * reference itself, executed unmodified in the build container under NumPy 2.3.5
 * / SciPy 1.18.1 by oracle/gen_golden.py and committed under tests/golden/
 * (tests/test_oracle_golden.py re-checks them on every run).
 *
 * Every function cites the reference file:line it follows.  All float32
 * arithmetic is written one IEEE operation per statement and the file MUST be
 * compiled with -ffp-contract=off (no FMA): NumPy rounds after every scalar op.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* NumPy float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
 * FLOAT_pairwise_sum; reached from np.sum at environment.py:166,167,178,234,235).
 * The reduction result is initialised to 0 and the contiguous array is fed to
 * one inner-loop call, so np.sum(a) == 0 + pairwise(a, n).                    */
static float np_pairwise_f32(const float *a, long n)
{
    if (n < 8) {
        float res = 0.0f;
        for (long i = 0; i < n; i++) res = res + a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res = res + a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_f32(a, n2) + np_pairwise_f32(a + n2, n - n2);
    }
}

ORC_API float orc_np_sum_f32(const float *a, long n) { return 0.0f + np_pairwise_f32(a, n); }

/* ------------------------------------------------------------------------ */
/* draw_line.py:4-102  beam_voxels.
 * Emits the in-bounds splat writes in the reference's order (x outer, dy, dz
 * inner) as (linear index, weight) pairs; at most 4*(max(G)+1) of them.
 * Returns the pair count, or -1 for the ValueError at draw_line.py:23-24.     */
#define ORC_EPS 1e-6 /* python float; becomes float32 at each use (NEP 50) */

ORC_API int orc_beam_trace(const double pos_d[3], const double dir_d[3], const int grid[3],
                           int *idx_out, float *w_out, int *n_slabs)
{
    const float eps = (float)ORC_EPS;
    float p[3], v[3];
    if (n_slabs) *n_slabs = 0;
    for (int i = 0; i < 3; i++) { p[i] = (float)pos_d[i]; v[i] = (float)dir_d[i]; } /* :19-20 */

    /* :22  np.linalg.norm(float32[3]) -> sqrt(sdot(v,v)); OpenBLAS sdot rounds each
     * product to float32 and accumulates in double, result cast to float32.     */
    float q0 = v[0] * v[0], q1 = v[1] * v[1], q2 = v[2] * v[2];
    double acc = (double)q0;
    acc = acc + (double)q1;
    acc = acc + (double)q2;
    float norm = sqrtf((float)acc);
    if (norm < eps) return -1;                                   /* :23-24 */
    for (int i = 0; i < 3; i++) v[i] = v[i] / norm;              /* :25 */

    float t_entry = -INFINITY, t_exit = INFINITY;
    for (int i = 0; i < 3; i++) {                                /* :31-43 */
        float gm1 = (float)(grid[i] - 1);
        float te, tx;
        if (fabsf(v[i]) > eps) {
            float t1 = (-p[i]) / v[i];
            float t2 = (gm1 - p[i]) / v[i];
            te = t1 < t2 ? t1 : t2;
            tx = t1 < t2 ? t2 : t1;
        } else {
            if (p[i] < 0.0f || p[i] > gm1) return 0;
            te = -INFINITY;
            tx = INFINITY;
        }
        if (te > t_entry) t_entry = te;                          /* :44 */
        if (tx < t_exit) t_exit = tx;                            /* :45 */
    }
    if (t_entry > t_exit) return 0;                              /* :46-47 */

    int dom = 0;                                                 /* :49-50 first max wins */
    float best = fabsf(v[0]);
    for (int i = 1; i < 3; i++) if (fabsf(v[i]) > best) { best = fabsf(v[i]); dom = i; }
    int o0 = dom == 0 ? 1 : 0;                                   /* :51 */
    int o1 = dom == 2 ? 1 : 2;
    int step = v[dom] > 0.0f ? 1 : -1;                           /* :53 */

    float a0 = t_entry * v[dom];                                 /* :55-60 */
    float s0 = p[dom] + a0;
    int x0 = (int)floorf(s0);
    float a1 = t_exit * v[dom];
    float s1 = p[dom] + a1;
    int x1 = (int)floorf(s1);

    float my = t_entry * v[o0];                                  /* :62-63 */
    float y = p[o0] + my;
    float mz = t_entry * v[o1];
    float z = p[o1] + mz;

    float den = v[dom] + eps;                                    /* :65-66 */
    float gy = v[o0] / den;
    float gz = v[o1] / den;
    float sgy = gy * (float)step;
    float sgz = gz * (float)step;

    int n = 0, slabs = 0;
    for (int x = x0; (x - x1) * step <= 0; x += step) {          /* :68-100 */
        int idx[3];
        float yf = floorf(y), zf = floorf(z);
        float fy = y - yf, fz = z - zf;                          /* :76-82 */
        idx[dom] = x;
        idx[o0] = (int)yf;
        idx[o1] = (int)zf;
        for (int dy = 0; dy < 2; dy++)
            for (int dz = 0; dz < 2; dz++) {                     /* :84-96 */
                float w = dy == 0 ? (1.0f - fy) : fy;
                float wz = dz == 0 ? (1.0f - fz) : fz;
                w = w * wz;
                int ix = idx[0], iy = idx[1] + dy, iz = idx[2] + dz; /* :88-90 axis quirk */
                if (ix >= 0 && ix < grid[0] && iy >= 0 && iy < grid[1] && iz >= 0 && iz < grid[2]) {
                    idx_out[n] = (ix * grid[1] + iy) * grid[2] + iz;
                    w_out[n] = w;
                    n++;
                }
            }
        y = y + sgy;                                             /* :98-99 */
        z = z + sgz;
        slabs++;
    }
    if (n_slabs) *n_slabs = slabs;
    return n;
}

/* draw_line.py:17,96,102: dense float32 volume, output[...] += weight in order. */
ORC_API int orc_beam_voxels(const double pos[3], const double dir[3], const int grid[3], float *out)
{
    int idx[4 * 260];
    float w[4 * 260];
    long nv = (long)grid[0] * grid[1] * grid[2];
    memset(out, 0, sizeof(float) * nv);
    int n = orc_beam_trace(pos, dir, grid, idx, w, 0);
    if (n < 0) return -1;
    for (int k = 0; k < n; k++) out[idx[k]] = out[idx[k]] + w[k];
    return n;
}

/* Batched form of the same function for the parity tests: per ray the MERGED trace
 * (each voxel once, weights summed in write order, first-touch order), row stride cap.
 * count[k] = -1 marks the ValueError rays.                                         */
ORC_API void orc_beam_batch(const double *pos, const double *dir, int m, const int grid[3], int cap,
                            int *idx_out, float *w_out, int *count_out)
{
    int idx[4 * 260];
    float w[4 * 260];
    for (int k = 0; k < m; k++) {
        int n = orc_beam_trace(pos + 3 * k, dir + 3 * k, grid, idx, w, 0);
        int *io = idx_out + (long)k * cap;
        float *wo = w_out + (long)k * cap;
        if (n < 0) { count_out[k] = -1; continue; }
        int u = 0;
        for (int a = 0; a < n; a++) {
            int found = -1;
            /* duplicates only come from the previous slab: look back over the last 8 entries */
            for (int b = u - 1; b >= 0 && b >= u - 8; b--) if (io[b] == idx[a]) { found = b; break; }
            if (found >= 0) wo[found] = wo[found] + w[a];
            else if (u < cap) { io[u] = idx[a]; wo[u] = w[a]; u++; }
        }
        count_out[k] = u;
    }
}

/* ------------------------------------------------------------------------ */
/* transforms.py:62-69  apply_translation (float64). */
ORC_API void orc_apply_translation(const double pos[3], const double t[3], const double bounds[3],
                                   double out_pos[3], double overshoot[3])
{
    for (int i = 0; i < 3; i++) {
        double tp = pos[i] + t[i];
        double bp = tp < 0.0 ? 0.0 : (tp > bounds[i] ? bounds[i] : tp);
        out_pos[i] = bp;
        overshoot[i] = fabs(tp - bp);
    }
}

static double norm3(const double a[3])
{
    double s = a[0] * a[0];
    s = s + a[1] * a[1];
    s = s + a[2] * a[2];
    return sqrt(s);
}

/* transforms.py:7-59  apply_rotation (float64).  Lines 25-26 call
 * scipy.spatial.transform.Rotation (third party, not under /root/reference;
 * pinned scipy==1.14.1 in environment.yaml:170, 1.18.1 in the container): the
 * published algorithm is restated here from scipy/spatial/transform/_rotation_xp.py
 * from_rotvec (159-179), as_matrix (302-333) and apply (631-645).               */
ORC_API void orc_apply_rotation(const double dir_in[3], const double rv[3], double min_angle,
                                double out_dir[3], double *overshoot)
{
    double d[3], q[4], r[3];
    double n = norm3(dir_in);                                    /* :23 */
    for (int i = 0; i < 3; i++) d[i] = dir_in[i] / n;

    double angle = norm3(rv);                                    /* from_rotvec */
    double scale;
    if (angle <= 1e-3) {
        double a2 = angle * angle;
        scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
    } else {
        scale = sin(angle / 2.0) / angle;
    }
    q[0] = scale * rv[0]; q[1] = scale * rv[1]; q[2] = scale * rv[2];
    q[3] = cos(angle / 2.0);

    double x = q[0], y = q[1], z = q[2], w = q[3];               /* as_matrix */
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    double m[3][3];
    m[0][0] = x2 - y2 - z2 + w2;  m[0][1] = 2.0 * (xy - zw);     m[0][2] = 2.0 * (xz + yw);
    m[1][0] = 2.0 * (xy + zw);    m[1][1] = -x2 + y2 - z2 + w2;  m[1][2] = 2.0 * (yz - xw);
    m[2][0] = 2.0 * (xz - yw);    m[2][1] = 2.0 * (yz + xw);     m[2][2] = -x2 - y2 + z2 + w2;
    for (int i = 0; i < 3; i++) {                                /* apply */
        double s = m[i][0] * d[0];
        s = s + m[i][1] * d[1];
        s = s + m[i][2] * d[2];
        r[i] = s;
    }
    n = norm3(r);                                                /* :27 */
    for (int i = 0; i < 3; i++) r[i] = r[i] / n;

    double zc = r[0] < -1.0 ? -1.0 : (r[0] > 1.0 ? 1.0 : r[0]);  /* :29 */
    double sg = zc > 0.0 ? 1.0 : (zc < 0.0 ? -1.0 : 0.0);        /* :30 */
    double ang = acos(zc);                                       /* :31 */
    if (sg < 0.0) ang = M_PI - ang;                              /* :32-33 */

    double nd[3];
    if (fabs(ang) < min_angle) {                                 /* :35-51 */
        double tz = sg * cos(min_angle);
        double mag = sqrt(1.0 - tz * tz);
        double px = r[1], py = r[2];
        double pn = sqrt(px * px + py * py);
        if (pn < 1e-8) { px = 1.0; py = 0.0; }
        else {
            px = px / pn;
            py = py / pn;
        }
        nd[0] = tz; nd[1] = px * mag; nd[2] = py * mag;
    } else {
        nd[0] = r[0]; nd[1] = r[1]; nd[2] = r[2];                /* :53 */
    }
    n = norm3(nd);                                               /* :55 */
    for (int i = 0; i < 3; i++) out_dir[i] = nd[i] / n;
    double os = min_angle - ang;                                 /* :57 */
    *overshoot = os > 0.0 ? os : 0.0;
}

/* One pose update for m independent (pos, dir, action) triples:
 * environment.py:112-143 (map_translation / map_rotation) + transforms.py:7-69.   */
ORC_API void orc_pose_batch(const double *pos, const double *dir, const float *actions, int m,
                            const int grid[3], double *pos_out, double *dir_out,
                            double *os_t, double *os_r)
{
    for (int k = 0; k < m; k++) {
        double tr[3], bounds[3], rv[3];
        for (int i = 0; i < 3; i++) {
            float a = actions[6 * k + i];
            a = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
            double sc = (double)a * (double)grid[i];
            tr[i] = sc * 0.2;
            bounds[i] = (double)grid[i];
            float r = actions[6 * k + 3 + i];
            r = r < -1.0f ? -1.0f : (r > 1.0f ? 1.0f : r);
            float sr = r * (float)M_PI;
            rv[i] = (double)(sr * 0.5f);
        }
        orc_apply_translation(pos + 3 * k, tr, bounds, pos_out + 3 * k, os_t + 3 * k);
        orc_apply_rotation(dir + 3 * k, rv, M_PI / 4.0, dir_out + 3 * k, os_r + k);
    }
}

/* ------------------------------------------------------------------------ */
/* environment.py:15-273  RadiotherapyEnv on dense float32 volumes, as the
 * reference holds them.                                                      */
typedef struct {
    int grid[3];
    long nv;
    const float *lungs;      /* shared, :29,39 */
    float *tumours;          /* :87-97 */
    float *dose;             /* :104-105 */
    float *tmp;              /* scratch for dense products */
    double pos[3], dir[3];   /* :101-102 */
    int t;                   /* :81 */
    int done;
} orc_env;

ORC_API orc_env *orc_env_create(const float *lungs, const int grid[3])
{
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    memcpy(e->grid, grid, sizeof(int) * 3);
    e->nv = (long)grid[0] * grid[1] * grid[2];
    e->lungs = lungs;
    e->tumours = (float *)calloc(e->nv, sizeof(float));
    e->dose = (float *)calloc(e->nv, sizeof(float));
    e->tmp = (float *)calloc(e->nv, sizeof(float));
    return e;
}

ORC_API void orc_env_destroy(orc_env *e)
{
    if (!e) return;
    free(e->tumours); free(e->dose); free(e->tmp); free(e);
}

/* environment.py:77-105  reset; the tumour volume is given as its voxel list
 * (the caller picks the file — :90 uses the global NumPy RNG).               */
ORC_API void orc_env_reset(orc_env *e, const int *vox, int n_vox)
{
    memset(e->tumours, 0, sizeof(float) * e->nv);
    for (int k = 0; k < n_vox; k++) e->tumours[vox[k]] = 1.0f;
    for (int i = 0; i < 3; i++) e->pos[i] = (double)e->grid[i] / 2.0;   /* :101 */
    e->dir[0] = 0.0; e->dir[1] = 1.0; e->dir[2] = 0.0;                  /* :102 */
    memset(e->dose, 0, sizeof(float) * e->nv);                          /* :105 */
    e->t = 0;
    e->done = 0;
}

/* environment.py:145-148,259-268  get_vector_observation (float64[9]). */
ORC_API void orc_env_vector_obs(const orc_env *e, double obs[9])
{
    double c[3] = {0, 0, 0};
    long cnt = 0;
    for (int i = 0; i < e->grid[0]; i++)
        for (int j = 0; j < e->grid[1]; j++)
            for (int k = 0; k < e->grid[2]; k++)
                if (e->tumours[((long)i * e->grid[1] + j) * e->grid[2] + k] == 1.0f) {
                    c[0] += i; c[1] += j; c[2] += k; cnt++;    /* exact integer sums */
                }
    for (int i = 0; i < 3; i++) {
        double g = (double)e->grid[i];
        obs[i] = e->pos[i] / g * 2.0 - 1.0;
        obs[3 + i] = e->dir[i];
        obs[6 + i] = (c[i] / (double)cnt) / g * 2.0 - 1.0;
    }
}

/* environment.py:245-257  get_volumes: clip(stack[lungs,tumours,dose,view],0,1). */
ORC_API void orc_env_volumes(const orc_env *e, float *out /* 4*nv */)
{
    long nv = e->nv;
    float *b1 = (float *)malloc(sizeof(float) * nv);
    float *b2 = (float *)malloc(sizeof(float) * nv);
    double horiz[3] = {1.0, 0.0, 0.0};
    orc_beam_voxels(e->pos, e->dir, e->grid, b1);                /* :246 */
    orc_beam_voxels(e->pos, horiz, e->grid, b2);                 /* :247-249 */
    for (long i = 0; i < nv; i++) {
        float view = b1[i] + b2[i];                              /* :250 */
        float vals[4] = {e->lungs[i], e->tumours[i], e->dose[i], view};
        for (int c = 0; c < 4; c++) {
            float x = vals[c];
            x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);         /* :257 */
            out[c * nv + i] = x;
        }
    }
    free(b1); free(b2);
}

/* Output record of one step (all float64 unless noted).
 *  out[0..8]   observation (vector mode)                      :259-268
 *  out[9]      reward                                         :218
 *  out[10..12] tumour, lung, distance reward components       :214-216
 *  out[13]     doses.tumour  out[14] doses.lung               :234-235
 *  out[15..17] translation overshoot  out[18] rotation overshoot :237-240
 *  out[19]     lung voxels above threshold (integer valued)   :177
 * returns done (0/1), or -1 on the beam_voxels ValueError.                  */
#define ORC_STEP_OUT 20

ORC_API int orc_env_step(orc_env *e, const float action[6], double out[ORC_STEP_OUT])
{
    const long nv = e->nv;
    e->t += 1;                                                   /* :194 */

    /* :112-126 map_translation: clip(f32) * int64 shape -> float64, * 0.2 */
    double tr[3], bounds[3];
    float rvf[3];
    double rv[3];
    for (int i = 0; i < 3; i++) {
        float a = action[i];
        a = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
        double s = (double)a * (double)e->grid[i];
        tr[i] = s * 0.2;
        bounds[i] = (double)e->grid[i];
    }
    /* :128-143 map_rotation: float32 array * pi * 0.5 stays float32 */
    for (int i = 0; i < 3; i++) {
        float a = action[3 + i];
        a = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
        float s = a * (float)M_PI;
        rvf[i] = s * 0.5f;
        rv[i] = (double)rvf[i];
    }
    double npos[3], ndir[3], os_t[3], os_r;
    orc_apply_translation(e->pos, tr, bounds, npos, os_t);       /* :202-204 */
    orc_apply_rotation(e->dir, rv, M_PI / 4.0, ndir, &os_r);     /* :205-207 */
    memcpy(e->pos, npos, sizeof(npos));
    memcpy(e->dir, ndir, sizeof(ndir));

    /* :107-110 add_beam: dose = clip(dose + beam*0.1, 0, 1), dense */
    float *beam = e->tmp;
    if (orc_beam_voxels(e->pos, e->dir, e->grid, beam) < 0) return -1;
    for (long i = 0; i < nv; i++) {
        float b = beam[i] * 0.1f;
        float d = e->dose[i] + b;
        d = d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
        e->dose[i] = d;
    }

    /* :164-171 tumour_dose_reward (float32 throughout) */
    for (long i = 0; i < nv; i++) e->tmp[i] = e->dose[i] * e->tumours[i];
    float total_tumour_dose = orc_np_sum_f32(e->tmp, nv);
    float total_tumour = orc_np_sum_f32(e->tumours, nv);
    float ratio = total_tumour_dose / total_tumour;
    float tumour_reward = ratio * 10.0f;

    /* :173-182 lungs_dose_reward: int64 count / float32 sum -> float64 */
    long count = 0;
    for (long i = 0; i < nv; i++) {
        float mask = e->lungs[i] * (1.0f - e->tumours[i]);
        float ld = e->dose[i] * mask;
        e->tmp[i] = mask;
        if (ld > 0.2f) count++;
    }
    float total_lung = orc_np_sum_f32(e->tmp, nv);
    double lung_reward = (double)count / (double)total_lung * -1.0;

    /* :150-162 distance_to_tumour_reward (float64) */
    double bestn = INFINITY, bv[3] = {0, 0, 0};
    for (int i = 0; i < e->grid[0]; i++)
        for (int j = 0; j < e->grid[1]; j++)
            for (int k = 0; k < e->grid[2]; k++)
                if (e->tumours[((long)i * e->grid[1] + j) * e->grid[2] + k] == 1.0f) {
                    double dv[3] = {(double)i - e->pos[0], (double)j - e->pos[1], (double)k - e->pos[2]};
                    double nn = norm3(dv);
                    if (nn < bestn) { bestn = nn; bv[0] = dv[0]; bv[1] = dv[1]; bv[2] = dv[2]; }
                }
    double gs[3] = {(double)e->grid[0], (double)e->grid[1], (double)e->grid[2]};
    double dist_reward = norm3(bv) / norm3(gs) * -1.0;

    double reward = (double)tumour_reward + lung_reward + dist_reward;   /* :218 */

    /* :184-191,220 termination: float32 ratio >= 0.9 (python float -> float32) */
    int done = (ratio >= 0.9f) || (e->t >= 100);
    e->done = done;

    /* :234-235 info doses */
    for (long i = 0; i < nv; i++) e->tmp[i] = e->dose[i] * e->tumours[i];
    double dose_t = (double)orc_np_sum_f32(e->tmp, nv);
    for (long i = 0; i < nv; i++) e->tmp[i] = e->dose[i] * e->lungs[i];
    double dose_l = (double)orc_np_sum_f32(e->tmp, nv);

    orc_env_vector_obs(e, out);
    out[9] = reward;
    out[10] = (double)tumour_reward;
    out[11] = lung_reward;
    out[12] = dist_reward;
    out[13] = dose_t;
    out[14] = dose_l;
    out[15] = os_t[0]; out[16] = os_t[1]; out[17] = os_t[2];
    out[18] = os_r;
    out[19] = (double)count;
    return done;
}

ORC_API const float *orc_env_dose(const orc_env *e) { return e->dose; }
ORC_API void orc_env_pose(const orc_env *e, double pose[6])
{
    memcpy(pose, e->pos, sizeof(double) * 3);
    memcpy(pose + 3, e->dir, sizeof(double) * 3);
}
ORC_API void orc_env_set_pose(orc_env *e, const double pose[6])
{
    memcpy(e->pos, pose, sizeof(double) * 3);
    memcpy(e->dir, pose + 3, sizeof(double) * 3);
}

/* ------------------------------------------------------------------------ */
/* A batch of independent episodes with gymnasium-1.0.0 SyncVectorEnv NEXT_STEP
 * autoreset semantics (train.py:93,151; third party, not under /root/reference):
 * the call after a terminal step ignores the action, resets the env and reports
 * reward 0 / terminated False.  Tumour for episode k of env i is
 * tumour_ids[k*n + i] (explicit schedule; the reference draws it from the global
 * NumPy RNG, environment.py:90).  out is [T][n][ORC_STEP_OUT], done is [T][n].
 * Only envs [i0, i1) are advanced: the Python wrapper fans disjoint ranges out to
 * threads (ctypes drops the GIL), each env being an independent reference instance. */
ORC_API long orc_rollout(const float *lungs, const int grid[3],
                         const int *vox_offsets, const int *vox,
                         const int *tumour_ids, int n_episodes_max,
                         const float *actions /* [T][n][6] */, int T, int n,
                         double *out, signed char *done_out, int i0, int i1)
{
    long steps = 0;
    for (int i = i0; i < i1 && i < n; i++) {
        orc_env *e = orc_env_create(lungs, grid);
        int ep = 0;
        int tid = tumour_ids[(long)ep * n + i];
        orc_env_reset(e, vox + vox_offsets[tid], vox_offsets[tid + 1] - vox_offsets[tid]);
        int needs_reset = 0;
        for (int t = 0; t < T; t++) {
            double *o = out + ((long)t * n + i) * ORC_STEP_OUT;
            if (needs_reset) {
                ep = ep + 1 < n_episodes_max ? ep + 1 : ep;
                tid = tumour_ids[(long)ep * n + i];
                orc_env_reset(e, vox + vox_offsets[tid], vox_offsets[tid + 1] - vox_offsets[tid]);
                memset(o, 0, sizeof(double) * ORC_STEP_OUT);
                orc_env_vector_obs(e, o);
                done_out[(long)t * n + i] = 0;
                needs_reset = 0;
            } else {
                int d = orc_env_step(e, actions + ((long)t * n + i) * 6, o);
                done_out[(long)t * n + i] = (signed char)d;
                needs_reset = d > 0;
            }
            steps++;
        }
        orc_env_destroy(e);
    }
    return steps;
}

/* ------------------------------------------------------------------------ */
/* train.py:164-181  GAE, float32, evaluation order of the torch expression:
 * delta = (r + (gamma*v_next)*nnt) - v ; A = delta + ((gamma*lambda)*nnt)*A_next */
ORC_API void orc_gae(const float *rewards, const float *values, const float *dones,
                     const float *next_value, const float *next_done, int T, int n,
                     double gamma, double gae_lambda, float *adv, float *ret)
{
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    for (int i = 0; i < n; i++) {
        float last = 0.0f;
        for (int t = T - 1; t >= 0; t--) {
            float nnt, nv;
            if (t == T - 1) { nnt = 1.0f - next_done[i]; nv = next_value[i]; }
            else { nnt = 1.0f - dones[(long)(t + 1) * n + i]; nv = values[(long)(t + 1) * n + i]; }
            float a = g * nv;
            a = a * nnt;
            float d = rewards[(long)t * n + i] + a;
            d = d - values[(long)t * n + i];
            float b = gl * nnt;
            b = b * last;
            last = d + b;
            adv[(long)t * n + i] = last;
            ret[(long)t * n + i] = last + values[(long)t * n + i];
        }
    }
}
