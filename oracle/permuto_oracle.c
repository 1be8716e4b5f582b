/*
 * oracle/permuto_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the permutohedral-lattice
 * bilateral filter that the reference runs for its DenseCRF loss.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (libtcamcrf.so) never does.
 *
 * It follows the x86-64 (SSE) code path of the reference, because that is the
 * path that is compiled on every machine the reference runs on
 * (`__SSE__` => SSE_PERMUTOHEDRAL, dlib/crf/crfwrapper/bilateralfilter/
 * permutohedral.hpp:48-51):
 *
 *   po_build          <- Permutohedral::init           permutohedral.cpp:115-297
 *   po_filter         <- Permutohedral::compute        permutohedral.cpp:507-572
 *   keytab_*          <- HashTable::{hash,find,grow}   permutohedral.cpp:13-100
 *   po_bilateral*     <- initializePermutohedral/bilateralfilter/_batch
 *                                                      bilateralfilter.cpp:4-55
 *   po_color*         <- colour variants               colorbilateralfilter.cpp:4-54
 *
 * Numerics that matter for parity and are reproduced on purpose:
 *   - scale factors are evaluated in double and then narrowed to float
 *     (permutohedral.cpp:156-159);
 *   - every per-pixel operation is a separately rounded fp32 operation (SSE2
 *     has no FMA): compile this file with -ffp-contract=off;
 *   - the nearest-simplex rounding is round-half-to-even (cvtps2dq under the
 *     default MXCSR, permutohedral.cpp:162-165,193) -> lrintf();
 *   - rank comparisons are fp32 `<` (permutohedral.cpp:203-209);
 *   - vertex ids are handed out in first-seen order, pixel-major then
 *     remainder-major (permutohedral.cpp:245-251) -- ids therefore do not
 *     depend on the hash function, and `offset` can be compared one to one
 *     with the reference's `offset_`;
 *   - splat accumulates in pixel order, blur is a Jacobi ping-pong over the
 *     axes 0..d with a zero sentinel for missing neighbours, slice multiplies
 *     (bary*alpha) first and then the value (permutohedral.cpp:526-567).
 *
 * A reference quirk reproduced on purpose: it processes pixels four at a time
 * and also inserts the zero-feature padding pixels of the last partial block
 * (permutohedral.cpp:173,238-251).  Those vertices never receive a splat, but
 * they exist: they pick up values from their neighbours during the blur and
 * hand them on, so they DO change the output (rel ~1e-2 near black pixels at
 * the image origin) whenever N % 4 != 0.  po_build inserts them like the
 * reference does, and M counts them.
 *
 * Parity pinned: tests/test_oracle.py checks this file bit for bit against
 * the reference's own C++ compiled unmodified into oracle/_ref/ (when
 * /root/reference is present) and against the committed fixtures in
 * tests/golden/ that were generated from that build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PO_MAX_D 15

typedef struct po_lattice {
    int d;          /* feature dimension */
    int n;          /* number of input points (pixels) */
    int m;          /* number of lattice vertices */
    int32_t *offset;  /* [n*(d+1)]  vertex id per (pixel, remainder) */
    float *bary;      /* [n*(d+1)]  barycentric weight per (pixel, remainder) */
    int32_t *nbr;     /* [(d+1)*m*2] (n1,n2) per (axis, vertex), -1 = absent */
    int16_t *keys;    /* [m*d] lattice coordinates of each vertex */
} po_lattice;

/* ------------------------------------------------------------------ */
/* key table: open addressing + linear probing, ids in first-seen order */
/* (permutohedral.cpp:13-100).  Capacity is a power of two and the hash */
/* is a different mixer than the reference's; neither is observable.    */
/* ------------------------------------------------------------------ */
typedef struct {
    int d;
    size_t cap, used;
    int32_t *slot;   /* -1 = empty, else vertex id */
    int16_t *keys;   /* dense [id*d] */
    size_t keys_cap;
} keytab;

static size_t keytab_hash(const int16_t *k, int d)
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < d; i++) {
        h ^= (uint16_t)k[i];
        h *= 1099511628211ull;
    }
    h ^= h >> 29;
    return (size_t)h;
}

static int keytab_init(keytab *t, int d, size_t expect)
{
    size_t cap = 64;
    while (cap < 2 * expect) cap <<= 1;
    t->d = d;
    t->cap = cap;
    t->used = 0;
    t->slot = (int32_t *)malloc(cap * sizeof(int32_t));
    t->keys_cap = expect + 16;
    t->keys = (int16_t *)malloc(t->keys_cap * (size_t)d * sizeof(int16_t));
    if (!t->slot || !t->keys) return -1;
    memset(t->slot, 0xff, cap * sizeof(int32_t));
    return 0;
}

static int keytab_rehash(keytab *t)
{
    size_t ncap = t->cap * 2;
    int32_t *ns = (int32_t *)malloc(ncap * sizeof(int32_t));
    if (!ns) return -1;
    memset(ns, 0xff, ncap * sizeof(int32_t));
    for (size_t id = 0; id < t->used; id++) {
        size_t h = keytab_hash(t->keys + id * t->d, t->d) & (ncap - 1);
        while (ns[h] >= 0) h = (h + 1) & (ncap - 1);
        ns[h] = (int32_t)id;
    }
    free(t->slot);
    t->slot = ns;
    t->cap = ncap;
    return 0;
}

/* returns the id of `k`, inserting it when create != 0; -1 when absent */
static int32_t keytab_find(keytab *t, const int16_t *k, int create)
{
    const int d = t->d;
    if (create && 2 * t->used >= t->cap)
        if (keytab_rehash(t)) return -2;
    size_t h = keytab_hash(k, d) & (t->cap - 1);
    for (;;) {
        int32_t e = t->slot[h];
        if (e < 0) {
            if (!create) return -1;
            if (t->used == t->keys_cap) {
                size_t nk = t->keys_cap * 2;
                int16_t *p = (int16_t *)realloc(t->keys, nk * (size_t)d * sizeof(int16_t));
                if (!p) return -2;
                t->keys = p;
                t->keys_cap = nk;
            }
            memcpy(t->keys + t->used * d, k, (size_t)d * sizeof(int16_t));
            t->slot[h] = (int32_t)t->used;
            return (int32_t)t->used++;
        }
        if (memcmp(t->keys + (size_t)e * d, k, (size_t)d * sizeof(int16_t)) == 0)
            return e;
        h = (h + 1) & (t->cap - 1);
    }
}

static void keytab_release(keytab *t)
{
    free(t->slot);
    free(t->keys);
    t->slot = NULL;
    t->keys = NULL;
}

/* ------------------------------------------------------------------ */
/* per-point embedding (permutohedral.cpp:168-252, one SSE lane)        */
/* ------------------------------------------------------------------ */

/* Diagonal of the elevation matrix E, evaluated like the reference does:
 * a float inv_std_dev, then a double expression narrowed to float
 * (permutohedral.cpp:156-159). */
void po_scale_factors(int d, float *sf)
{
    float inv_std_dev = sqrt(2.0 / 3.0) * (d + 1);
    for (int i = 0; i < d; i++)
        sf[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);
}

/* Embeds one feature vector: writes the d+1 multiples-of-(d+1) `rem0`, the
 * d+1 ranks and the d+1 barycentric weights (indexed by remainder). */
static void po_embed(const float *f, int d, const float *sf,
                     float *rem0, int *rank, float *bary_out)
{
    float elevated[PO_MAX_D + 1];
    float frank[PO_MAX_D + 1];
    float b[PO_MAX_D + 2];
    const float inv_dp1 = 1.0f / (d + 1);
    const float dp1 = (float)(d + 1);

    /* y = E p  (permutohedral.cpp:177-184) */
    float sm = 0.0f;
    for (int j = d; j > 0; j--) {
        float cf = f[j - 1] * sf[j - 1];
        float t = (float)j * cf;
        elevated[j] = sm - t;
        sm = sm + cf;
    }
    elevated[0] = sm;

    /* nearest 0-coloured vertex, round half to even (:187-197) */
    float sum = 0.0f;
    for (int i = 0; i <= d; i++) {
        float v = inv_dp1 * elevated[i];
        v = (float)lrintf(v);
        rem0[i] = v * dp1;
        sum = sum + v;
    }

    /* rank of each coordinate's residual (:200-210) */
    for (int i = 0; i <= d; i++) frank[i] = 0.0f;
    for (int i = 0; i < d; i++) {
        float di = elevated[i] - rem0[i];
        for (int j = i + 1; j <= d; j++) {
            float dj = elevated[j] - rem0[j];
            float c = (di < dj) ? 1.0f : 0.0f;
            frank[i] = frank[i] + c;
            frank[j] = frank[j] + (1.0f - c);
        }
    }

    /* bring the point back onto the plane sum == 0 (:213-219) */
    for (int i = 0; i <= d; i++) {
        frank[i] = frank[i] + sum;
        float add = (frank[i] < 0.0f) ? dp1 : 0.0f;
        float sub = (frank[i] >= dp1) ? dp1 : 0.0f;
        frank[i] = frank[i] + (add - sub);
        rem0[i] = rem0[i] + (add - sub);
    }

    /* barycentric coordinates (:222-240) */
    for (int i = 0; i <= d + 1; i++) b[i] = 0.0f;
    for (int i = 0; i <= d; i++) {
        float v = (elevated[i] - rem0[i]) * inv_dp1;
        int p = (int)((float)d - frank[i]);
        b[p] = b[p] + v;
        b[p + 1] = b[p + 1] - v;
    }
    b[0] = b[0] + (1.0f + b[d + 1]);

    for (int i = 0; i <= d; i++) {
        rank[i] = (int)frank[i];
        bary_out[i] = b[i];
    }
}

void po_free(po_lattice *L)
{
    if (!L) return;
    free(L->offset);
    free(L->bary);
    free(L->nbr);
    free(L->keys);
    free(L);
}

/* Builds the lattice of `n` d-dimensional features (row-major [n][d]). */
po_lattice *po_build(const float *feature, int d, int n)
{
    if (d < 1 || d > PO_MAX_D || n < 0) return NULL;
    po_lattice *L = (po_lattice *)calloc(1, sizeof(po_lattice));
    if (!L) return NULL;
    const int dp1 = d + 1;
    L->d = d;
    L->n = n;
    L->offset = (int32_t *)calloc((size_t)(n + 4) * dp1, sizeof(int32_t));
    L->bary = (float *)calloc((size_t)(n + 4) * dp1, sizeof(float));

    keytab tab;
    if (!L->offset || !L->bary || keytab_init(&tab, d, (size_t)n + 16)) {
        po_free(L);
        return NULL;
    }

    float sf[PO_MAX_D];
    po_scale_factors(d, sf);

    /* canonical simplex (permutohedral.cpp:148-153) */
    int canonical[(PO_MAX_D + 1) * (PO_MAX_D + 1)];
    for (int i = 0; i <= d; i++) {
        for (int j = 0; j <= d - i; j++) canonical[i * dp1 + j] = i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * dp1 + j] = i - dp1;
    }

    const int n_padded = (n + 3) & ~3; /* the reference's 4-wide blocks */
    float zero_f[PO_MAX_D] = {0};
    for (int p = 0; p < n_padded; p++) {
        const float *f = p < n ? feature + (size_t)p * d : zero_f;
        float rem0[PO_MAX_D + 1], bary[PO_MAX_D + 1];
        int rank[PO_MAX_D + 1];
        int16_t key[PO_MAX_D];
        po_embed(f, d, sf, rem0, rank, bary);
        for (int r = 0; r <= d; r++) {
            for (int i = 0; i < d; i++)
                key[i] = (int16_t)(rem0[i] + (float)canonical[r * dp1 + rank[i]]);
            int32_t id = keytab_find(&tab, key, 1);
            if (id < 0) {
                keytab_release(&tab);
                po_free(L);
                return NULL;
            }
            L->offset[(size_t)p * dp1 + r] = id;
            L->bary[(size_t)p * dp1 + r] = bary[r];
        }
    }

    /* neighbour table (permutohedral.cpp:272-295) */
    const int m = (int)tab.used;
    L->m = m;
    L->nbr = (int32_t *)malloc((size_t)dp1 * (m > 0 ? m : 1) * 2 * sizeof(int32_t));
    if (!L->nbr) {
        keytab_release(&tab);
        po_free(L);
        return NULL;
    }
    for (int j = 0; j <= d; j++) {
        for (int v = 0; v < m; v++) {
            const int16_t *key = tab.keys + (size_t)v * d;
            int16_t lo[PO_MAX_D + 1], hi[PO_MAX_D + 1];
            for (int k = 0; k < d; k++) {
                lo[k] = (int16_t)(key[k] - 1);
                hi[k] = (int16_t)(key[k] + 1);
            }
            if (j < d) { /* axis d only exists implicitly (coords sum to 0) */
                lo[j] = (int16_t)(key[j] + d);
                hi[j] = (int16_t)(key[j] - d);
            }
            L->nbr[((size_t)j * m + v) * 2 + 0] = keytab_find(&tab, lo, 0);
            L->nbr[((size_t)j * m + v) * 2 + 1] = keytab_find(&tab, hi, 0);
        }
    }
    /* hand the dense key array over to the lattice */
    L->keys = tab.keys;
    tab.keys = NULL;
    keytab_release(&tab);
    return L;
}

/* out = alpha * Slice(Blur(Splat(in))) for one scalar plane of n values
 * (permutohedral.cpp:507-572 with value_size == 1). */
int po_filter(const po_lattice *L, const float *in, float *out)
{
    const int d = L->d, dp1 = d + 1, n = L->n, m = L->m;
    float *val = (float *)calloc((size_t)m + 2, sizeof(float));
    float *nxt = (float *)calloc((size_t)m + 2, sizeof(float));
    if (!val || !nxt) {
        free(val);
        free(nxt);
        return -1;
    }
    /* splat: index 0 is the zero sentinel, vertex v lives at v+1 */
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * dp1 + j] + 1;
            float w = L->bary[(size_t)i * dp1 + j];
            float t = w * in[i];
            val[o] = val[o] + t;
        }
    /* blur along each of the d+1 axes, ping-pong */
    for (int j = 0; j <= d; j++) {
        for (int v = 0; v < m; v++) {
            int n1 = L->nbr[((size_t)j * m + v) * 2 + 0] + 1;
            int n2 = L->nbr[((size_t)j * m + v) * 2 + 1] + 1;
            float s = val[n1] + val[n2];
            float h = 0.5f * s;
            nxt[v + 1] = val[v + 1] + h;
        }
        float *t = val;
        val = nxt;
        nxt = t;
    }
    /* slice */
    const float alpha = 1.0f / (1 + powf(2, -d));
    for (int i = 0; i < n; i++) {
        float acc = 0.0f;
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * dp1 + j] + 1;
            float w = L->bary[(size_t)i * dp1 + j] * alpha;
            float t = w * val[o];
            acc = acc + t;
        }
        out[i] = acc;
    }
    free(val);
    free(nxt);
    return 0;
}

/* accessors for ctypes */
int po_num_vertices(const po_lattice *L) { return L->m; }
const int32_t *po_offsets(const po_lattice *L) { return L->offset; }
const float *po_barycentric(const po_lattice *L) { return L->bary; }
const int32_t *po_neighbours(const po_lattice *L) { return L->nbr; }
const int16_t *po_keys(const po_lattice *L) { return L->keys; }

/* ------------------------------------------------------------------ */
/* feature construction + per-image / batch drivers                     */
/* ------------------------------------------------------------------ */

/* 5-D features (col/sxy, row/sxy, R/srgb, G/srgb, B/srgb) of a planar CHW
 * image (bilateralfilter.cpp:4-19). */
po_lattice *po_lattice_bilateral(const float *image, int H, int W,
                                 float sigmargb, float sigmaxy)
{
    const int P = H * W;
    float *feat = (float *)malloc((size_t)P * 5 * sizeof(float));
    if (!feat) return NULL;
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) {
            int idx = r * W + c;
            feat[idx * 5 + 0] = (float)c / sigmaxy;
            feat[idx * 5 + 1] = (float)r / sigmaxy;
            feat[idx * 5 + 2] = image[0 * P + idx] / sigmargb;
            feat[idx * 5 + 3] = image[1 * P + idx] / sigmargb;
            feat[idx * 5 + 4] = image[2 * P + idx] / sigmargb;
        }
    po_lattice *L = po_build(feat, 5, P);
    free(feat);
    return L;
}

/* DIM-D colour-only features (colorbilateralfilter.cpp:4-18). */
po_lattice *po_lattice_color(const float *image, int H, int W,
                             float sigmargb, int DIM)
{
    const int P = H * W;
    float *feat = (float *)malloc((size_t)P * DIM * sizeof(float));
    if (!feat) return NULL;
    for (int idx = 0; idx < P; idx++)
        for (int z = 0; z < DIM; z++)
            feat[idx * DIM + z] = image[(size_t)z * P + idx] / sigmargb;
    po_lattice *L = po_build(feat, DIM, P);
    free(feat);
    return L;
}

static int po_filter_planes(const po_lattice *L, const float *in, float *out,
                            int K, int P)
{
    /* one scalar pass per class (bilateralfilter.cpp:31-37) */
    for (int k = 0; k < K; k++)
        if (po_filter(L, in + (size_t)k * P, out + (size_t)k * P)) return -1;
    return 0;
}

int po_bilateralfilter(const float *image, const float *in, float *out,
                       int K, int H, int W, float sigmargb, float sigmaxy)
{
    po_lattice *L = po_lattice_bilateral(image, H, W, sigmargb, sigmaxy);
    if (!L) return -1;
    int rc = po_filter_planes(L, in, out, K, H * W);
    po_free(L);
    return rc;
}

int po_bilateralfilter_batch(const float *images, const float *ins, float *outs,
                             int N, int K, int H, int W,
                             float sigmargb, float sigmaxy)
{
    /* bilateralfilter.cpp:42-55, without the OpenMP pragma */
    const size_t P = (size_t)H * W;
    for (int n = 0; n < N; n++)
        if (po_bilateralfilter(images + n * 3 * P, ins + n * K * P,
                               outs + n * K * P, K, H, W, sigmargb, sigmaxy))
            return -1;
    return 0;
}

int po_colorbilateralfilter(const float *image, const float *in, float *out,
                            int K, int H, int W, float sigmargb, int DIM)
{
    po_lattice *L = po_lattice_color(image, H, W, sigmargb, DIM);
    if (!L) return -1;
    int rc = po_filter_planes(L, in, out, K, H * W);
    po_free(L);
    return rc;
}

int po_colorbilateralfilter_batch(const float *images, const float *ins,
                                  float *outs, int N, int K, int H, int W,
                                  float sigmargb, int DIM)
{
    /* the image stride is 3*H*W whatever DIM is (colorbilateralfilter.cpp:50) */
    const size_t P = (size_t)H * W;
    for (int n = 0; n < N; n++)
        if (po_colorbilateralfilter(images + n * 3 * P, ins + n * K * P,
                                    outs + n * K * P, K, H, W, sigmargb, DIM))
            return -1;
    return 0;
}
