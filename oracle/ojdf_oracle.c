/*
 * oracle/ojdf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's per-frame gather / scatter arithmetic
 * (suryanshkumar/online-joint-depthfusion-and-semantic @ a4f9e19):
 *   modules/extractor.py:82-120   compute_coordinates        -> ojdf_oracle_unproject
 *   modules/extractor.py:309-345  extract_values             -> ojdf_oracle_extract (ray samples)
 *   modules/extractor.py:533-593  interpolation_weights      -> corner indices / weights
 *   modules/extractor.py:596-681  get_index_mask / trilinear -> gather + 8-term f64 sum
 *   modules/integrator.py:29-88   Integrator.forward (TSDF)  -> ojdf_oracle_integrate
 *   modules/integrator.py:90-124  Integrator.forward (sem.)  -> ojdf_oracle_integrate
 *   modules/pipeline.py:137-171   _prepare_volume_update     -> ojdf_oracle_integrate_frame
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library -- always as the checker or the reported CPU baseline,
 * never on the product path (the product path is the CUDA library and fails loudly
 * without it).
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the pin is the reference itself: the tests/golden npz fixtures were
 * produced by importing /root/reference's own Extractor / Integrator on CPU
 * (tests/golden/make_golden.py, 1 torch thread) and this file reproduces every
 * array in them bit for bit (tests/test_oracle_golden.py).
 *
 * The arithmetic is order-sensitive on purpose; build with -ffp-contract=off.
 * The reference rounds every f64 product and sum separately (ATen elementwise
 * kernels), sums the 8 corner terms in ATen's AVX2 row-reduction order, and
 * accumulates index_add_ sequentially in entry order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef _Float16 ojdf_half;

static inline float h2f(uint16_t h) { ojdf_half x; memcpy(&x, &h, 2); return (float)x; }
static inline uint16_t f2h(float f) { ojdf_half x = (ojdf_half)f; uint16_t h; memcpy(&h, &x, 2); return h; }

int ojdf_oracle_version(void) { return 1; }

/* ---- minimal pthread parallel-for (no OpenMP runtime in this image) -------------- */
static int g_threads = 1;
void ojdf_oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int ojdf_oracle_get_threads(void) { return g_threads; }
int ojdf_oracle_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n < 1 ? 1 : (int)n; }

typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; int64_t lo, hi; void *ctx; } range_job;
static void *range_tramp(void *p) { range_job *j = (range_job *)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void parallel_for(int64_t n, range_fn fn, void *ctx)
{
    int T = g_threads;
    if (T > n) T = (int)(n > 0 ? n : 1);
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256]; range_job jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].lo = n * t / T; jobs[t].hi = n * (t + 1) / T;
        if (pthread_create(&th[t], NULL, range_tramp, &jobs[t]) != 0) { fn(jobs[t].lo, jobs[t].hi, ctx); th[t] = 0; }
    }
    for (int t = 0; t < T; ++t) if (th[t]) pthread_join(th[t], NULL);
}

/* ---- A.1 unprojection, f32 (modules/extractor.py:82-120) -------------------------
 * pixel (col*z, row*z, z) -> Kinv (3x3 row major) -> E (3x4 row major, cam->world).
 * fma_chain=1: t=fl(a0*b0); t=fma(a1,b1,t); ...   (what MKL sgemm does at N>=65536)
 * fma_chain=0: every product and sum rounded separately, left to right.
 * The reference leaves this order to the BLAS library; both forms are within an
 * ulp or two of each other, and everything downstream is checked on identical
 * world points. */
void ojdf_oracle_unproject(const float *depth, int h, int w, const float *Kinv, const float *E,
                           int fma_chain, float *world)
{
    for (int r = 0; r < h; ++r) {
        for (int c = 0; c < w; ++c) {
            const int n = r * w + c;
            const float z = depth[n];
            const float p[3] = { (float)c * z, (float)r * z, z };
            float q[4];
            for (int i = 0; i < 3; ++i) {
                float t = Kinv[3 * i] * p[0];
                if (fma_chain) { t = fmaf(Kinv[3 * i + 1], p[1], t); t = fmaf(Kinv[3 * i + 2], p[2], t); }
                else { t = t + Kinv[3 * i + 1] * p[1]; t = t + Kinv[3 * i + 2] * p[2]; }
                q[i] = t;
            }
            q[3] = 1.0f;
            for (int i = 0; i < 3; ++i) {
                float t = E[4 * i] * q[0];
                if (fma_chain) {
                    t = fmaf(E[4 * i + 1], q[1], t); t = fmaf(E[4 * i + 2], q[2], t); t = fmaf(E[4 * i + 3], q[3], t);
                } else {
                    t = t + E[4 * i + 1] * q[1]; t = t + E[4 * i + 2] * q[2]; t = t + E[4 * i + 3] * q[3];
                }
                world[3 * n + i] = t;
            }
        }
    }
}

/* sign() as torch.sign: -1, 0, +1 */
static inline double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }

/* One ray sample: corner indices, corner weights (modules/extractor.py:533-593). */
static inline void corners(const double p[3], int64_t idx[8][3], double wgt[8])
{
    double fl[3], nb[3], a[3], ai[3];
    for (int d = 0; d < 3; ++d) {
        fl[d] = floor(p[d]);
        const double ctr = fl[d] + 0.5;
        nb[d] = sgn(ctr - p[d]);
        a[d] = fabs(p[d] - ctr);
        ai[d] = 1.0 - a[d];
    }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const int c = 4 * i + 2 * j + k;
                const double w1 = i ? a[0] : ai[0], w2 = j ? a[1] : ai[1], w3 = k ? a[2] : ai[2];
                idx[c][0] = (int64_t)(i ? fl[0] + nb[0] : fl[0]);
                idx[c][1] = (int64_t)(j ? fl[1] + nb[1] : fl[1]);
                idx[c][2] = (int64_t)(k ? fl[2] + nb[2] : fl[2]);
                wgt[c] = (w1 * w2) * w3;
            }
}

/* Ray samples of one pixel (modules/extractor.py:309-345), P odd, centre at P/2. */
static inline void ray_points(const float *world3, const float *eye, const double *origin, double res,
                              int P, double *pts /* P*3 */)
{
    double cv[3], dl[3];
    for (int d = 0; d < 3; ++d) {
        cv[d] = ((double)world3[d] - origin[d]) / res;
        const double ev = ((double)eye[d] - origin[d]) / res;
        dl[d] = cv[d] - ev;
    }
    const double nrm = sqrt((dl[0] * dl[0] + dl[1] * dl[1]) + dl[2] * dl[2]);
    const double den = nrm > 1e-12 ? nrm : 1e-12;          /* F.normalize eps clamp */
    const int half = P / 2;
    for (int d = 0; d < 3; ++d) {
        const double dir = dl[d] / den;
        pts[3 * half + d] = cv[d];
        for (int i = 1; i <= half; ++i) {
            const double step = (double)i * dir;
            pts[3 * (half + i) + d] = cv[d] + step;
            pts[3 * (half - i) + d] = cv[d] - step;
        }
    }
}

/* ---- A.2/A.3 ray samples + trilinear gather (modules/extractor.py:309-345,533-681) --
 * world (N,3) f32; eye (3) f32; origin (3) f64; volumes (X,Y,Z) fp16 bit patterns.
 * out_vals/out_wts (N,P) f32 always; out_points (N,P,3) f64, out_idx (N,P,8,3) i64,
 * out_w (N,P,8) f64 optional (NULL to skip). */
typedef struct {
    const float *world; const float *eye; const double *origin; double res;
    const uint16_t *tsdf; const uint16_t *wvol; int X, Y, Z, P;
    float *out_vals; float *out_wts; double *out_points; int64_t *out_idx; double *out_w;
} extract_ctx;

static void extract_range(int64_t lo, int64_t hi, void *vctx)
{
    const extract_ctx *a = (const extract_ctx *)vctx;
    const float *world = a->world, *eye = a->eye; const double *origin = a->origin; const double res = a->res;
    const uint16_t *tsdf = a->tsdf, *wvol = a->wvol; const int X = a->X, Y = a->Y, Z = a->Z, P = a->P;
    float *out_vals = a->out_vals, *out_wts = a->out_wts; double *out_points = a->out_points;
    int64_t *out_idx = a->out_idx; double *out_w = a->out_w;
    for (int64_t n = lo; n < hi; ++n) {
        double pts[3 * 33];
        ray_points(world + 3 * n, eye, origin, res, P, pts);
        for (int k = 0; k < P; ++k) {
            int64_t idx[8][3]; double wgt[8], tv[8], tw[8];
            corners(pts + 3 * k, idx, wgt);
            for (int c = 0; c < 8; ++c) {
                const int ok = idx[c][0] >= 0 && idx[c][0] < X && idx[c][1] >= 0 && idx[c][1] < Y &&
                               idx[c][2] >= 0 && idx[c][2] < Z;
                float v = -0.1f, wv = 0.0f;                 /* modules/extractor.py:663-664 */
                if (ok) {
                    const int64_t lin = (idx[c][0] * Y + idx[c][1]) * (int64_t)Z + idx[c][2];
                    v = h2f(tsdf[lin]); wv = h2f(wvol[lin]);
                }
                tv[c] = (double)v * wgt[c];
                tw[c] = (double)wv * wgt[c];
            }
            /* ATen AVX2 row-sum order for 8 contiguous f64 */
            out_vals[n * P + k] = (float)((((tv[0] + tv[4]) + (tv[1] + tv[5])) + (tv[2] + tv[6])) + (tv[3] + tv[7]));
            out_wts[n * P + k]  = (float)((((tw[0] + tw[4]) + (tw[1] + tw[5])) + (tw[2] + tw[6])) + (tw[3] + tw[7]));
            if (out_points) memcpy(out_points + (n * P + k) * 3, pts + 3 * k, 3 * sizeof(double));
            if (out_idx) memcpy(out_idx + (n * P + k) * 24, idx, 24 * sizeof(int64_t));
            if (out_w) memcpy(out_w + (n * P + k) * 8, wgt, 8 * sizeof(double));
        }
    }
}

void ojdf_oracle_extract(const float *world, int64_t N, const float *eye, const double *origin, double res,
                         const uint16_t *tsdf, const uint16_t *wvol, int X, int Y, int Z, int P,
                         float *out_vals, float *out_wts, double *out_points, int64_t *out_idx, double *out_w)
{
    extract_ctx a = { world, eye, origin, res, tsdf, wvol, X, Y, Z, P, out_vals, out_wts, out_points, out_idx, out_w };
    parallel_for(N, extract_range, &a);
}

/* ---- A.4/A.5 integration in the reference's `updates` form (modules/integrator.py:15-126)
 * values (M1) f32 (already clamped by the caller, modules/pipeline.py:157-159),
 * idx (M1,8,3) i64, w (M1,8) f64; entry e = m*8+c.  Volumes are updated in place.
 * ids (M1) u8 / scores (M1) f32 per sample; do_sem = DATA.semantics && test.
 * Uses the same 2 x G^3 fp32 scratch as the reference (modules/integrator.py:59-67). */
int ojdf_oracle_integrate(const float *values, const int64_t *idx, const double *w, int64_t M1,
                          uint16_t *tsdf, uint16_t *wvol, int X, int Y, int Z,
                          const uint8_t *ids, const float *scores, uint8_t *ids_vol, uint16_t *scores_vol,
                          int do_sem)
{
    const int64_t G = (int64_t)X * Y * Z, M = M1 * 8;
    float *cw = (float *)calloc((size_t)G, sizeof(float));
    float *cu = (float *)calloc((size_t)G, sizeof(float));
    int64_t *lin = (int64_t *)malloc((size_t)(M > 0 ? M : 1) * sizeof(int64_t));
    if (!cw || !cu || !lin) { free(cw); free(cu); free(lin); return -1; }
    for (int64_t e = 0; e < M; ++e) {
        const int64_t *ix = idx + 3 * e;
        const int ok = ix[0] >= 0 && ix[0] < X && ix[1] >= 0 && ix[1] < Y && ix[2] >= 0 && ix[2] < Z;
        lin[e] = ok ? (ix[0] * Y + ix[1]) * (int64_t)Z + ix[2] : -1;
        if (!ok) continue;
        const float wf = (float)w[e];
        const float up = wf * values[e / 8];               /* separately rounded */
        cw[lin[e]] += wf;                                   /* index_add_, entry order */
        cu[lin[e]] += up;
    }
    /* all reads of the old volumes happen before any write (gather-then-scatter);
       every entry of a voxel computes the same value, so update each voxel once and
       mark it done with the scratch (NaN marks "already written"). */
    uint16_t *sc_new = NULL; uint8_t *id_new = NULL; uint8_t *id_wr = NULL;
    if (do_sem) {
        sc_new = (uint16_t *)malloc((size_t)(M > 0 ? M : 1) * 2);
        id_new = (uint8_t *)malloc((size_t)(M > 0 ? M : 1));
        id_wr = (uint8_t *)malloc((size_t)(M > 0 ? M : 1));
        for (int64_t e = 0; e < M; ++e) {
            if (lin[e] < 0) continue;
            const float so = h2f(scores_vol[lin[e]]);
            const uint8_t io = ids_vol[lin[e]];
            const float s = scores[e / 8];
            const uint8_t id = ids[e / 8];
            sc_new[e] = f2h(s > so ? s : so);
            id_new[e] = s > so ? id : io;
            id_wr[e] = io != id;
        }
    }
    for (int64_t e = 0; e < M; ++e) {
        if (lin[e] < 0) continue;
        const int64_t l = lin[e];
        if (cw[l] != cw[l]) continue;                       /* already stored */
        const float wo = h2f(wvol[l]), vo = h2f(tsdf[l]);
        const float W = cw[l], U = cu[l];
        wvol[l] = f2h(wo + W);
        tsdf[l] = f2h((wo * vo + U) / (wo + W));            /* 0/0 -> NaN is stored, as the reference */
        cw[l] = NAN;
    }
    if (do_sem) {
        for (int64_t e = 0; e < M; ++e) {                   /* index_put_, last entry wins */
            if (lin[e] < 0) continue;
            if (id_wr[e]) ids_vol[lin[e]] = id_new[e];
            scores_vol[lin[e]] = sc_new[e];
        }
    }
    free(cw); free(cu); free(lin); free(sc_new); free(id_new); free(id_wr);
    return 0;
}

typedef struct {
    const float *world; const float *est; const float *eye; const double *origin; double res;
    int P, tail; float clampv; const uint8_t *pix_ids; const float *pix_scores; const int64_t *slot;
    float *values; int64_t *idx; double *w; uint8_t *ids; float *scores;
} frame_ctx;

static void frame_range(int64_t lo, int64_t hi, void *vctx)
{
    const frame_ctx *a = (const frame_ctx *)vctx;
    for (int64_t n = lo; n < hi; ++n) {
        if (a->slot[n] < 0) continue;
        double pts[3 * 33];
        ray_points(a->world + 3 * n, a->eye, a->origin, a->res, a->P, pts);
        for (int k = 0; k < a->tail; ++k) {
            const int64_t m = a->slot[n] * a->tail + k;
            int64_t ci[8][3]; double cwt[8];
            corners(pts + 3 * k, ci, cwt);
            memcpy(a->idx + m * 24, ci, sizeof(ci));
            memcpy(a->w + m * 8, cwt, sizeof(cwt));
            float x = a->est[n * a->P + k];
            x = x < -a->clampv ? -a->clampv : (x > a->clampv ? a->clampv : x);   /* torch.clamp; NaN propagates */
            a->values[m] = x;
            a->ids[m] = a->pix_ids ? a->pix_ids[n] : 0;
            a->scores[m] = a->pix_scores ? a->pix_scores[n] : 0.0f;
        }
    }
}

/* ---- whole-frame form: what Pipeline.fuse does between the network and the volumes
 * (modules/pipeline.py:137-171 + modules/integrator.py:15-126) without materialising
 * the updates dict.  est (N,P) f32 network output; filt_depth (N) f32 (0 = masked);
 * tail = n_tail_points; clampv = DATA.init_value.  pix_ids/pix_scores per pixel. */
int ojdf_oracle_integrate_frame(const float *world, const float *filt_depth, const float *est, int64_t N,
                                const float *eye, const double *origin, double res, int P, int tail,
                                float clampv, uint16_t *tsdf, uint16_t *wvol, int X, int Y, int Z,
                                const uint8_t *pix_ids, const float *pix_scores,
                                uint8_t *ids_vol, uint16_t *scores_vol, int do_sem)
{
    int64_t nv = 0;
    for (int64_t n = 0; n < N; ++n) nv += filt_depth[n] != 0.0f;
    const int64_t M1 = nv * tail;
    float *values = (float *)malloc((size_t)(M1 > 0 ? M1 : 1) * sizeof(float));
    int64_t *idx = (int64_t *)malloc((size_t)(M1 > 0 ? M1 : 1) * 24 * sizeof(int64_t));
    double *w = (double *)malloc((size_t)(M1 > 0 ? M1 : 1) * 8 * sizeof(double));
    uint8_t *ids = (uint8_t *)malloc((size_t)(M1 > 0 ? M1 : 1));
    float *scores = (float *)malloc((size_t)(M1 > 0 ? M1 : 1) * sizeof(float));
    if (!values || !idx || !w || !ids || !scores) { free(values); free(idx); free(w); free(ids); free(scores); return -1; }
    int64_t *slot = (int64_t *)malloc((size_t)(N > 0 ? N : 1) * sizeof(int64_t));
    int64_t v = 0;
    for (int64_t n = 0; n < N; ++n) slot[n] = filt_depth[n] != 0.0f ? v++ : -1;   /* nonzero(): ascending n */
    frame_ctx fc = { world, est, eye, origin, res, P, tail, clampv, pix_ids, pix_scores, slot, values, idx, w, ids, scores };
    parallel_for(N, frame_range, &fc);
    const int rc = ojdf_oracle_integrate(values, idx, w, M1, tsdf, wvol, X, Y, Z, ids, scores, ids_vol, scores_vol, do_sem);
    free(values); free(idx); free(w); free(ids); free(scores); free(slot);
    return rc;
}
