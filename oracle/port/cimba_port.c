/*
 * cimba_port.c - TEST INFRASTRUCTURE ONLY (the parity oracle, "port" kind).
 *
 * CPU restatement of the hot path of ambonvik/cimba in plain C11: sfc64 +
 * ziggurat samplers, the hashheap future-event list, the event dispatcher, the
 * resource-guard wait list, cmb_objectqueue and cmb_resourcepool, the running
 * summaries, and the three queueing workloads of SURVEY.md section 8d.
 *
 * The reference runs each simulated process on its own stack
 * (src/cmi_coroutine.c); here a process is a resume-point index plus locals -
 * the same re-expression the CUDA engine uses - while the data structures keep
 * the reference's shapes (1-based binary heap with slot 0 as the pop scratch,
 * the same sift loops, the same comparators), so pop order follows from the
 * same algorithm, not merely from the same ordering relation.
 *
 * Compiled with -ffp-contract=off: the reference build has no FMA contraction
 * (no -march=native, meson.build:22-27).
 *
 * Parity: pinned - see cimba_port.h and tests/test_oracle_*.py.
 */
#include "cimba_port.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#include "zig_tables.h"

/* ===================================================================== RNG */

/* src/cmb_random.c:70-80 (MurmurHash3 finaliser over seed + nonce) */
uint64_t port_fmix64(uint64_t seed, uint64_t nonce)
{
    uint64_t h = seed + nonce;
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

/* src/cmb_random.c:54-62 */
uint64_t port_sfc64(port_rng *r)
{
    const uint64_t out = r->a + r->b + r->d++;
    r->a = r->b ^ (r->b >> 11);
    r->b = r->c + (r->c << 3);
    r->c = ((r->c << 24) | (r->c >> 40)) + out;
    return out;
}

/* src/cmb_random.c:91-124: splitmix64 x4 -> a,b,c,d, then 20 discarded draws */
void port_rng_init(port_rng *r, uint64_t seed)
{
    uint64_t sm = seed;
    uint64_t *dst[4] = { &r->a, &r->b, &r->c, &r->d };
    for (int i = 0; i < 4; i++) {
        uint64_t z = (sm += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        *dst[i] = z ^ (z >> 31);
    }
    for (int i = 0; i < 20; i++) {
        (void)port_sfc64(r);
    }
}

/* include/cmb_random.h:149-152 */
double port_random(port_rng *r)
{
    return ldexp((double)(port_sfc64(r) >> 11), -53);
}

/* include/cmb_random.h:165-173 */
double port_uniform(port_rng *r, double lo, double hi)
{
    return lo + (hi - lo) * port_random(r);
}

/* include/cmb_random.h:749-754 */
unsigned port_bernoulli(port_rng *r, double p)
{
    return (port_random(r) <= p) ? 1u : 0u;
}

/* include/cmb_random.h:840-846 */
long port_dice(port_rng *r, long a, long b)
{
    const double x = (double)(b - a + 1) * port_random(r);
    return (long)floor((double)a + x);
}

/* src/cmb_random.c:198-206: map integer coordinates inside overhang j to doubles */
static double exp_overhang_x(unsigned j, uint64_t u)
{
    return ldexp(zt_exp_x[j], 64) + (zt_exp_x[j - 1] - zt_exp_x[j]) * (double)u;
}

static double exp_overhang_y(unsigned j, uint64_t u)
{
    return ldexp(zt_exp_y[j - 1], 64) + (zt_exp_y[j] - zt_exp_y[j - 1]) * (double)u;
}

/* src/cmb_random.c:216-285: the 1.56 % of draws that miss the ziggurat body */
static double exp_slow(port_rng *r, uint64_t ux)
{
    double shift = 0.0;
    for (;;) {
        uint64_t uy = port_sfc64(r);
        unsigned j = (unsigned)(uy & 0xff);
        if (port_sfc64(r) >= zt_exp_prob[j]) {
            j = zt_exp_alias[j];
        }
        if (j > 0u) {
            for (;;) {
                if (uy > (UINT64_MAX - ux)) {           /* reflect into the triangle */
                    uy = UINT64_MAX - uy;
                    ux = UINT64_MAX - ux;
                }
                const uint64_t gap = (UINT64_MAX - ux) - uy;
                const double x = exp_overhang_x(j, ux);
                if (gap >= zt_exp_concavity[j]) {
                    return x + shift;
                }
                if (exp_overhang_y(j, uy) <= exp(-x)) {
                    return x + shift;
                }
                uy = port_sfc64(r);
                ux = port_sfc64(r);
            }
        }
        shift += ZT_EXP_TAIL;                           /* memoryless tail */
        ux = port_sfc64(r);
        const unsigned i = (unsigned)(ux & 0xff);
        if (i <= ZT_EXP_MAX) {
            return zt_exp_x[i] * (double)ux + shift;
        }
    }
}

/* include/cmb_random.h:319-329 */
double port_std_exponential(port_rng *r)
{
    const uint64_t u = port_sfc64(r);
    const unsigned i = (unsigned)(u & 0xff);
    return (i <= ZT_EXP_MAX) ? zt_exp_x[i] * (double)u : exp_slow(r, u);
}

/* include/cmb_random.h:344-352 */
double port_exponential(port_rng *r, double mean)
{
    return mean * port_std_exponential(r);
}

/* include/cmb_random.h:366-378 */
double port_erlang(port_rng *r, unsigned k, double m)
{
    double x = 0.0;
    for (unsigned i = 0u; i < k; i++) {
        x += port_exponential(r, m);
    }
    return x;
}

/* src/cmb_random.c:321-345 helpers */
static double nor_overhang_x(unsigned j, int64_t ix)
{
    return ldexp(zt_nor_x[j], 63) + (zt_nor_x[j - 1] - zt_nor_x[j]) * (double)ix;
}

static double nor_overhang_y(unsigned j, uint64_t uy)
{
    return ldexp(zt_nor_y[j - 1], 63) + (zt_nor_y[j] - zt_nor_y[j - 1]) * (double)uy;
}

static int64_t draw63(port_rng *r)
{
    return (int64_t)(port_sfc64(r) & (uint64_t)INT64_MAX);
}

static double nor_pdf_scaled(double x)
{
    return exp(-0.5 * x * x);
}

/* src/cmb_random.c:352-451 */
static double nor_slow(port_rng *r, int64_t ix)
{
    const double sign = (ix < 0) ? -1.0 : 1.0;
    ix &= INT64_MAX;

    int64_t iy = draw63(r);
    unsigned j = (unsigned)(iy & 0xff);
    if (ix >= zt_nor_prob[j]) {
        j = zt_nor_alias[j];
    }

    if (j > ZT_NOR_INFLECTION) {                        /* convex overhang */
        for (;;) {
            const double x = nor_overhang_x(j, ix);
            const int64_t gap = (INT64_MAX - ix) - iy;
            if (gap >= 0) {
                return sign * x;
            }
            if (gap + zt_nor_convexity[j] >= 0) {
                if (nor_overhang_y(j, (uint64_t)iy) < nor_pdf_scaled(x)) {
                    return sign * x;
                }
            }
            ix = draw63(r);
            iy = draw63(r);
        }
    }
    else if (j == 0u) {                                 /* tail (Marsaglia) */
        double x, z;
        do {
            x = ZT_NOR_INV_TAIL * port_exponential(r, 1.0);
            z = port_exponential(r, 1.0);
        } while (2 * z <= x * x);
        return sign * (x + ZT_NOR_TAIL);
    }
    else if (j < ZT_NOR_INFLECTION) {                   /* concave overhang */
        for (;;) {
            if (iy > INT64_MAX - ix) {
                iy = INT64_MAX - iy;
                ix = INT64_MAX - ix;
            }
            const double x = nor_overhang_x(j, ix);
            const int64_t gap = (INT64_MAX - ix) - iy;
            if (gap >= zt_nor_concavity[j]) {
                return sign * x;
            }
            if (nor_overhang_y(j, (uint64_t)iy) <= nor_pdf_scaled(x)) {
                return sign * x;
            }
            ix = draw63(r);
            iy = draw63(r);
        }
    }
    else {                                              /* the inflection layer */
        for (;;) {
            const double x = nor_overhang_x(j, ix);
            const int64_t gap = (INT64_MAX - ix) - iy;
            if (gap >= zt_nor_concavity[j]) {
                return sign * x;
            }
            if (gap + zt_nor_convexity[j] > 0) {
                if (nor_overhang_y(j, (uint64_t)iy) < nor_pdf_scaled(x)) {
                    return sign * x;
                }
            }
            ix = draw63(r);
            iy = draw63(r);
        }
    }
}

/* include/cmb_random.h:206-215: the sign rides in the signed integer */
double port_std_normal(port_rng *r)
{
    const int64_t ix = (int64_t)port_sfc64(r);
    const unsigned i = (unsigned)(ix & 0xff);
    return (i <= ZT_NOR_MAX) ? zt_nor_x[i] * (double)ix : nor_slow(r, ix);
}

/* include/cmb_random.h:230-235 */
double port_normal(port_rng *r, double mu, double sigma)
{
    return mu + sigma * port_std_normal(r);
}

int port_rng_draws(uint64_t seed, int kind, double p0, double p1, uint64_t n, double *out)
{
    port_rng r;
    port_rng_init(&r, seed);
    for (uint64_t i = 0u; i < n; i++) {
        switch (kind) {
        case 0: { uint64_t u = port_sfc64(&r); memcpy(&out[i], &u, 8); break; }
        case 1: out[i] = port_exponential(&r, p0); break;
        case 2: out[i] = port_std_normal(&r); break;
        case 3: out[i] = port_random(&r); break;
        case 4: out[i] = port_normal(&r, p0, p1); break;
        case 5: out[i] = port_erlang(&r, (unsigned)p0, p1); break;
        case 6: out[i] = port_uniform(&r, p0, p1); break;
        case 7: out[i] = (double)port_dice(&r, (long)p0, (long)p1); break;
        case 8: out[i] = (double)port_bernoulli(&r, p0); break;
        default: return -1;
        }
    }
    return 0;
}

/* =================================================== the rest of cmb_random
 *
 * include/cmb_random.h:189-940 and src/cmb_random.c:299-313, 465-766, restated.  Kind
 * numbers as in ref_rng_draws_ex (oracle/ref_build/ref_driver.c).  Everything here is a
 * thin layer over sfc64 / the two ziggurats plus sqrt (IEEE exact) and, for some, libm's
 * log / exp / pow - on the CPU that is the same glibc the reference links, so this
 * restatement is bit-identical to it; the CUDA side documents which kinds can differ in
 * the last place because of that.
 */

/* src/cmb_random.c:500-520 */
double port_triangular(port_rng *r, double min, double mode, double max)
{
    const double u = port_random(r);
    if (u < (mode - min) / (max - min)) {
        return min + sqrt(u * (max - min) * (mode - min));
    }
    return max - sqrt((1.0 - u) * (max - min) * (max - mode));
}

double port_lognormal(port_rng *r, double m, double s)         /* include/cmb_random.h:249-257 */
{
    return exp(port_normal(r, m, s));
}

double port_logistic(port_rng *r, double m, double s)          /* :267-273 */
{
    const double x = port_random(r);
    return m + s * log(x / (1.0 - x));
}

double port_cauchy(port_rng *r, double mode, double scale)     /* :290-299 */
{
    const double x = port_std_normal(r);
    double y;
    while ((y = port_std_normal(r)) == 0.0) {}
    return mode + scale * x / y;
}

double port_hypoexponential(port_rng *r, unsigned n, const double *ma)     /* :394-408 */
{
    double x = 0.0;
    for (unsigned i = 0u; i < n; i++) {
        x += port_exponential(r, ma[i]);
    }
    return x;
}

unsigned port_loaded_dice(port_rng *r, unsigned n, const double *pa)       /* src/cmb_random.c:644-662 */
{
    const double x = port_random(r);
    double q = 0.0;
    unsigned ui;
    for (ui = 0u; ui < n; ui++) {
        q += pa[ui];
        if (x < q) {
            break;
        }
    }
    return ui;
}

double port_hyperexponential(port_rng *r, unsigned n, const double *ma, const double *pa)  /* :299-313 */
{
    const unsigned ui = port_loaded_dice(r, n, pa);
    return port_exponential(r, ma[ui]);
}

/* Marsaglia & Tsang, src/cmb_random.c:465-497 */
double port_std_gamma(port_rng *r, double shape)
{
    const double d = shape - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    double x, v;
    for (;;) {
        do {
            x = port_std_normal(r);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        const double w = v * v * v;
        const double u = port_random(r);
        if ((u < 1.0 - 0.331 * (x * x) * (x * x))
            || (log(u) < (0.5 * x * x) + (d * (1.0 - w + log(w))))) {
            return d * w;
        }
    }
}

/* include/cmb_random.h:451-463.  The two factors of the shape < 1 branch are unsequenced in
 * the reference; gcc evaluates the std_gamma call first (checked against oracle/_ref) */
double port_gamma(port_rng *r, double shape, double scale)
{
    if (shape >= 1.0) {
        return scale * port_std_gamma(r, shape);
    }
    const double g = port_std_gamma(r, shape + 1.0);
    const double u = port_random(r);
    return scale * (g * pow(u, 1.0 / shape));
}

double port_std_beta(port_rng *r, double a, double b)          /* :476-487 */
{
    const double x = port_std_gamma(r, a);
    const double y = port_std_gamma(r, b);
    return x / (x + y);
}

double port_beta(port_rng *r, double a, double b, double min, double max)  /* :500-512 */
{
    return min + (max - min) * port_std_beta(r, a, b);
}

double port_PERT_mod(port_rng *r, double min, double mode, double max, double lambda)  /* src/cmb_random.c:523-538 */
{
    const double rng = max - min;
    const double a = 1.0 + lambda * (mode - min) / rng;
    const double b = 1.0 + lambda * (max - mode) / rng;
    return min + rng * port_std_beta(r, a, b);
}

double port_weibull(port_rng *r, double shape, double scale)   /* include/cmb_random.h:571-582 */
{
    const double u = port_exponential(r, 1.0);
    return scale * pow(u, 1.0 / shape);
}

double port_pareto(port_rng *r, double shape, double mode)     /* :595-605 */
{
    return mode / pow(port_random(r), 1.0 / shape);
}

double port_chisquared(port_rng *r, double k)                  /* :618-626 */
{
    return port_gamma(r, k / 2.0, 2.0);
}

double port_F_dist(port_rng *r, double a, double b)            /* :639-653 */
{
    const double x = port_chisquared(r, a) / a;
    double y;
    while ((y = port_chisquared(r, b) / b) == 0.0) {}
    return x / y;
}

double port_std_t_dist(port_rng *r, double v)                  /* :668-679 */
{
    const double x = port_std_normal(r);
    double y;
    while ((y = port_chisquared(r, v)) == 0.0) {}
    return x / sqrt(y / v);
}

double port_rayleigh(port_rng *r, double s)                    /* :714-725 */
{
    const double x = port_normal(r, 0.0, s);
    const double y = port_normal(r, 0.0, s);
    return sqrt(x * x + y * y);
}

/* src/cmb_random.c:541-552: 64 coin flips per sfc64 word, most significant bit first */
typedef struct { uint64_t bits; unsigned pos; } port_flipper;

static int port_flip(port_rng *r, port_flipper *f)
{
    if (f->pos == 0u) {
        f->bits = port_sfc64(r);
        f->pos = 64u;
    }
    return (int)((f->bits >> --f->pos) & 1u);
}

unsigned port_geometric(port_rng *r, double p)                 /* :558-573 */
{
    const double denom = -log(1.0 - p);
    return (unsigned)ceil(port_std_exponential(r) / denom);
}

unsigned port_binomial(port_rng *r, unsigned n, double p)      /* :576-588 */
{
    unsigned s = 0u;
    for (unsigned i = 0u; i < n; i++) {
        s += port_bernoulli(r, p);
    }
    return s;
}

unsigned port_negative_binomial(port_rng *r, unsigned m, double p)         /* :594-606 */
{
    unsigned f = 0u;
    for (unsigned i = 0u; i < m; i++) {
        f += port_geometric(r, p) - 1u;
    }
    return f;
}

unsigned port_poisson(port_rng *r, double rate)                /* :612-632 */
{
    const double m = 1.0 / rate;
    double t = 0.0;
    unsigned ctr = 0u;
    for (;;) {
        t += port_exponential(r, m);
        if (t <= 1.0) {
            ctr++;
        }
        else {
            break;
        }
    }
    return ctr;
}

/* Vose alias tables, src/cmb_random.c:672-752; sampling include/cmb_random.h:922-933 */
static uint64_t alias_secure(double p)
{
    if (p <= 0.0) {
        return 0u;
    }
    if (p >= 1.0) {
        return UINT64_MAX;
    }
    return (uint64_t)(p * (double)UINT64_MAX);
}

void port_alias_create(unsigned n, const double *pa, uint64_t *uprob, unsigned *alias)
{
    double *work = calloc(n, sizeof(double));
    unsigned *small = calloc(n, sizeof(unsigned));
    unsigned *large = calloc(n, sizeof(unsigned));
    double psum = 0.0;
    for (unsigned i = 0u; i < n; i++) {
        psum += pa[i];
        uprob[i] = 0u;
        alias[i] = 0u;
    }
    unsigned ns = 0u, nl = 0u;
    for (unsigned i = 0u; i < n; i++) {
        work[i] = pa[i] * n / psum;
        if (work[i] < 1.0) {
            small[ns++] = i;
        }
        else {
            large[nl++] = i;
        }
    }
    while (ns > 0u && nl > 0u) {
        const unsigned l = small[--ns];
        const unsigned g = large[--nl];
        uprob[l] = alias_secure(work[l]);
        alias[l] = g;
        work[g] = (work[g] + work[l]) - 1.0;
        if (work[g] < 1.0) {
            small[ns++] = g;
        }
        else {
            large[nl++] = g;
        }
    }
    while (nl > 0u) {
        uprob[large[--nl]] = UINT64_MAX;
    }
    while (ns > 0u) {
        uprob[small[--ns]] = UINT64_MAX;
    }
    free(large);
    free(small);
    free(work);
}

unsigned port_alias_sample(port_rng *r, unsigned n, const uint64_t *uprob, const unsigned *alias)
{
    const unsigned idx = (unsigned)floor(n * port_random(r));
    const bool c = port_sfc64(r) >= uprob[idx];
    return c ? alias[idx] : idx;
}

int port_rng_draws_ex(uint64_t seed, int kind, const double *par, uint32_t npar, uint64_t n, double *out)
{
    (void)npar;
    port_rng r;
    port_rng_init(&r, seed);
    port_flipper flips = { 0u, 0u };
    uint64_t uprob[64];
    unsigned alias[64];
    if (kind == 30) {
        if ((unsigned)par[0] > 64u) {
            return -1;
        }
        port_alias_create((unsigned)par[0], &par[1], uprob, alias);
    }
    for (uint64_t i = 0u; i < n; i++) {
        switch (kind) {
        case 9:  out[i] = port_triangular(&r, par[0], par[1], par[2]); break;
        case 10: out[i] = port_lognormal(&r, par[0], par[1]); break;
        case 11: out[i] = port_logistic(&r, par[0], par[1]); break;
        case 12: out[i] = port_cauchy(&r, par[0], par[1]); break;
        case 13: out[i] = port_hypoexponential(&r, (unsigned)par[0], &par[1]); break;
        case 14: out[i] = port_hyperexponential(&r, (unsigned)par[0], &par[1], &par[1 + (unsigned)par[0]]); break;
        case 15: out[i] = port_gamma(&r, par[0], par[1]); break;
        case 16: out[i] = port_beta(&r, par[0], par[1], par[2], par[3]); break;
        case 17: out[i] = port_PERT_mod(&r, par[0], par[1], par[2], 4.0); break;
        case 18: out[i] = port_weibull(&r, par[0], par[1]); break;
        case 19: out[i] = port_pareto(&r, par[0], par[1]); break;
        case 20: out[i] = port_chisquared(&r, par[0]); break;
        case 21: out[i] = port_F_dist(&r, par[0], par[1]); break;
        case 22: out[i] = par[0] + par[1] * port_std_t_dist(&r, par[2]); break;
        case 23: out[i] = port_rayleigh(&r, par[0]); break;
        case 24: out[i] = (double)port_flip(&r, &flips); break;
        case 25: out[i] = (double)port_geometric(&r, par[0]); break;
        case 26: out[i] = (double)port_binomial(&r, (unsigned)par[0], par[1]); break;
        case 27: out[i] = (double)port_negative_binomial(&r, (unsigned)par[0], par[1]); break;
        case 28: out[i] = (double)port_poisson(&r, par[0]); break;
        case 29: out[i] = (double)port_loaded_dice(&r, (unsigned)par[0], &par[1]); break;
        case 30: out[i] = (double)port_alias_sample(&r, (unsigned)par[0], uprob, alias); break;
        case 31: out[i] = port_std_gamma(&r, par[0]); break;
        case 32: out[i] = port_PERT_mod(&r, par[0], par[1], par[2], par[3]); break;
        case 33: out[i] = (double)port_negative_binomial(&r, (unsigned)par[0], par[1]); break;
        default: return -1;
        }
    }
    return 0;
}

/* =============================================================== summaries */

/* src/cmb_datasummary.c:37-50 */
void port_summary_init(port_summary *s)
{
    s->count = 0u;
    s->min = DBL_MAX;
    s->max = -DBL_MAX;
    s->m1 = s->m2 = s->m3 = s->m4 = 0.0;
}

/* src/cmb_datasummary.c:144-166 (Meng's update order) */
uint64_t port_summary_add(port_summary *s, double y)
{
    s->max = (y > s->max) ? y : s->max;
    s->min = (y < s->min) ? y : s->min;

    const double d = y - s->m1;
    const double d_2 = d * d;
    const double d_3 = d * d_2;
    const double n = (double)(++s->count);
    const double d_n = d / n;
    const double d_n_2 = d_n * d_n;
    const double d_n_3 = d_n_2 * d_n;

    s->m1 += d_n;
    s->m2 += d * (d - d_n);
    s->m3 += d * (d_2 - d_n_2) - 3.0 * d_n * s->m2;
    s->m4 += d * (d_3 - d_n_3) - 6.0 * d_n_2 * s->m2 - 4.0 * d_n * s->m3;
    return s->count;
}

/* src/cmb_datasummary.c:93-131 (Pebay pairwise merge; tgt may alias a source) */
uint64_t port_summary_merge(port_summary *tgt, const port_summary *a, const port_summary *b)
{
    port_summary c;
    port_summary_init(&c);
    c.count = a->count + b->count;
    c.min = (a->min < b->min) ? a->min : b->min;
    c.max = (a->max > b->max) ? a->max : b->max;

    const double n1 = (double)a->count;
    const double n2 = (double)b->count;
    const double n = (double)c.count;
    const double d21 = b->m1 - a->m1;
    const double d21_n = d21 / n;
    const double d21_n_2 = d21_n * d21_n;
    const double d21_n_3 = d21_n * d21_n_2;

    c.m1 = a->m1 + n2 * d21_n;
    c.m2 = a->m2 + b->m2 + n1 * n2 * d21 * d21_n;
    c.m3 = a->m3 + b->m3
         + n1 * n2 * (n1 - n2) * d21 * d21_n_2
         + 3.0 * (n1 * b->m2 - n2 * a->m2) * d21_n;
    c.m4 = a->m4 + b->m4
         + n1 * n2 * (n1 * n1 - n1 * n2 + n2 * n2) * d21 * d21_n_3
         + 6.0 * (n1 * n1 * b->m2 + n2 * n2 * a->m2) * d21_n_2
         + 4.0 * (n1 * b->m3 - n2 * a->m3) * d21_n;
    *tgt = c;
    return tgt->count;
}

void port_wsummary_init(port_wsummary *s)
{
    port_summary_init(&s->ds);
    s->wsum = 0.0;
}

/* src/cmb_wtdsummary.c:82-137 */
uint64_t port_wsummary_add(port_wsummary *s, double x, double w)
{
    port_summary *d = &s->ds;
    if (w == 0.0) {
        return d->count;
    }
    if (d->count == 0u) {
        d->count = 1u;
        d->max = x;
        d->min = x;
        d->m1 = x;
        d->m2 = d->m3 = d->m4 = 0.0;
        s->wsum = w;
        return d->count;
    }

    d->max = (x > d->max) ? x : d->max;
    d->min = (x < d->min) ? x : d->min;
    d->count++;

    const double w1 = s->wsum;
    const double w2 = w;
    const double ws = w1 + w2;
    const double d21 = x - d->m1;
    const double d21_w = d21 / ws;
    const double d21_w_2 = d21_w * d21_w;
    const double d21_w_3 = d21_w * d21_w_2;

    const double m1 = d->m1 + w2 * d21_w;
    const double m2 = d->m2 + w1 * w2 * d21 * d21_w;
    const double m3 = d->m3
                    + w1 * w2 * (w1 - w2) * d21 * d21_w_2
                    - 3.0 * w2 * d->m2 * d21_w;
    const double m4 = d->m4
                    + w1 * w2 * (w1 * w1 - w1 * w2 + w2 * w2) * d21 * d21_w_3
                    + 6.0 * w2 * w2 * d->m2 * d21_w_2
                    - 4.0 * w2 * d->m3 * d21_w;
    d->m1 = m1;
    d->m2 = m2;
    d->m3 = m3;
    d->m4 = m4;
    s->wsum = ws;
    return d->count;
}

/* src/cmb_wtdsummary.c:152-194 */
uint64_t port_wsummary_merge(port_wsummary *tgt, const port_wsummary *a, const port_wsummary *b)
{
    port_wsummary t;
    port_wsummary_init(&t);
    const port_summary *p = &a->ds, *q = &b->ds;
    t.ds.count = p->count + q->count;
    t.ds.min = (p->min < q->min) ? p->min : q->min;
    t.ds.max = (p->max > q->max) ? p->max : q->max;

    const double w1 = a->wsum;
    const double w2 = b->wsum;
    const double ws = w1 + w2;
    const double d21 = q->m1 - p->m1;
    const double d21_w = d21 / ws;
    const double d21_w_2 = d21_w * d21_w;
    const double d21_w_3 = d21_w * d21_w_2;

    t.wsum = ws;
    t.ds.m1 = p->m1 + w2 * d21_w;
    t.ds.m2 = p->m2 + q->m2 + w1 * w2 * d21 * d21_w;
    t.ds.m3 = p->m3 + q->m3
            + w1 * w2 * (w1 - w2) * d21 * d21_w_2
            + 3.0 * (w1 * q->m2 - w2 * p->m2) * d21_w;
    t.ds.m4 = p->m4 + q->m4
            + w1 * w2 * (w1 * w1 - w1 * w2 + w2 * w2) * d21 * d21_w_3
            + 6.0 * (w1 * w1 * q->m2 + w2 * w2 * p->m2) * d21_w_2
            + 4.0 * (w1 * q->m3 - w2 * p->m3) * d21_w;
    *tgt = t;
    return t.ds.count;
}

static void flat_summary(const port_summary *s, double *out)
{
    out[0] = (double)s->count;
    out[1] = s->min;
    out[2] = s->max;
    out[3] = s->m1;
    out[4] = s->m2;
    out[5] = s->m3;
    out[6] = s->m4;
}

int port_datasummary_of(const double *x, uint64_t n, double *out)
{
    port_summary s;
    port_summary_init(&s);
    for (uint64_t i = 0u; i < n; i++) {
        port_summary_add(&s, x[i]);
    }
    flat_summary(&s, out);
    return 0;
}

int port_datasummary_split_merge(const double *x, uint64_t na, uint64_t n, double *out)
{
    port_summary a, b, m;
    port_summary_init(&a);
    port_summary_init(&b);
    for (uint64_t i = 0u; i < na; i++) {
        port_summary_add(&a, x[i]);
    }
    for (uint64_t i = na; i < n; i++) {
        port_summary_add(&b, x[i]);
    }
    port_summary_merge(&m, &a, &b);
    flat_summary(&m, out);
    return 0;
}

int port_wtdsummary_of(const double *x, const double *w, uint64_t n, double *out)
{
    port_wsummary s;
    port_wsummary_init(&s);
    for (uint64_t i = 0u; i < n; i++) {
        port_wsummary_add(&s, x[i], w[i]);
    }
    flat_summary(&s.ds, out);
    out[7] = s.wsum;
    return 0;
}

int port_wtdsummary_split_merge(const double *x, const double *w, uint64_t na, uint64_t n, double *out)
{
    port_wsummary a, b, m;
    port_wsummary_init(&a);
    port_wsummary_init(&b);
    for (uint64_t i = 0u; i < na; i++) {
        port_wsummary_add(&a, x[i], w[i]);
    }
    for (uint64_t i = na; i < n; i++) {
        port_wsummary_add(&b, x[i], w[i]);
    }
    port_wsummary_merge(&m, &a, &b);
    flat_summary(&m.ds, out);
    out[7] = m.wsum;
    return 0;
}

/* ================================================================ hashheap */

/* src/cmi_hashheap.h:53-59 without the hash_index back-pointer: the key->slot
 * map only accelerates lookups (src/cmi_hashheap.c:587-622); a linear search
 * finds the same slot, so heap contents and order are unaffected. */
typedef struct {
    uint64_t key;
    double   d;         /* rank_d64 */
    int64_t  i;         /* rank_i64 */
    int64_t  item[4];
} heap_tag;

typedef bool (*heap_before)(const heap_tag *a, const heap_tag *b);

typedef struct {
    heap_tag *slot;         /* 1-based; slot[0] = last popped (src/cmi_hashheap.c:496-498) */
    uint64_t cap, count, issued;
    heap_before before;
} heap;

/* src/cmi_hashheap.c:55-80: time asc, priority desc, key asc */
static bool fel_before(const heap_tag *a, const heap_tag *b)
{
    if (a->d < b->d) return true;
    if (a->d > b->d) return false;
    if (a->i > b->i) return true;
    if (a->i < b->i) return false;
    return a->key < b->key;
}

/* src/cmb_resourceguard.c:71-90, copied in meaning INCLUDING its fall-through
 * when a has the lower priority */
static bool guard_before(const heap_tag *a, const heap_tag *b)
{
    if (a->i > b->i) return true;
    if (a->d < b->d) return true;
    if (a->key < b->key) return true;
    return false;
}

static void heap_init(heap *h, unsigned exp2, heap_before before)
{
    h->cap = 1u << exp2;
    h->slot = calloc(h->cap + 1u, sizeof(heap_tag));
    h->count = 0u;
    h->issued = 0u;
    h->before = before;
}

static void heap_free(heap *h)
{
    free(h->slot);
    h->slot = NULL;
}

/* src/cmi_hashheap.c:277-316 */
static void sift_up(heap *h, uint64_t k)
{
    const heap_tag moving = h->slot[k];
    uint64_t parent;
    while ((parent = (k >> 1)) > 0u) {
        if (!h->before(&moving, &h->slot[parent])) {
            break;
        }
        h->slot[k] = h->slot[parent];
        k = parent;
    }
    h->slot[k] = moving;
}

/* src/cmi_hashheap.c:321-370 */
static void sift_down(heap *h, uint64_t k)
{
    const heap_tag moving = h->slot[k];
    const uint64_t last_parent = h->count >> 1;
    while (k <= last_parent) {
        uint64_t child = k << 1;
        if (child + 1u <= h->count && h->before(&h->slot[child + 1u], &h->slot[child])) {
            child++;
        }
        if (h->before(&moving, &h->slot[child])) {
            break;
        }
        h->slot[k] = h->slot[child];
        k = child;
    }
    h->slot[k] = moving;
}

/* src/cmi_hashheap.c:428-478; doubles on demand like :381-421 */
static uint64_t heap_push(heap *h, uint64_t key, double d, int64_t i,
                          int64_t p0, int64_t p1, int64_t p2)
{
    if (h->count == h->cap) {
        h->cap <<= 1;
        h->slot = realloc(h->slot, (h->cap + 1u) * sizeof(heap_tag));
    }
    const uint64_t at = ++h->count;
    h->issued += 1u;
    if (key == 0u) {
        key = h->issued;
    }
    heap_tag *t = &h->slot[at];
    t->key = key;
    t->d = d;
    t->i = i;
    t->item[0] = p0;
    t->item[1] = p1;
    t->item[2] = p2;
    t->item[3] = 0;
    sift_up(h, at);
    return key;
}

/* src/cmi_hashheap.c:486-524: result left in slot[0] */
static bool heap_pop(heap *h)
{
    if (h->count == 0u) {
        return false;
    }
    h->slot[0] = h->slot[1];
    if (h->count > 1u) {
        h->slot[1] = h->slot[h->count];
        h->count--;
        if (h->count > 1u) {
            sift_down(h, 1u);
        }
    }
    else {
        h->count = 0u;
    }
    return true;
}

/* src/cmi_hashheap.c:529-579 (lookup by linear search, see heap_tag note) */
static bool heap_remove(heap *h, uint64_t key)
{
    uint64_t at = 0u;
    for (uint64_t k = 1u; k <= h->count; k++) {
        if (h->slot[k].key == key) {
            at = k;
            break;
        }
    }
    if (at == 0u) {
        return false;
    }
    if (at == h->count) {
        h->count--;
        return true;
    }
    const bool down = h->before(&h->slot[at], &h->slot[h->count]);
    h->slot[at] = h->slot[h->count];
    h->count--;
    if (down) {
        sift_down(h, at);
    }
    else {
        sift_up(h, at);
    }
    return true;
}

int port_heap_script(uint64_t n, const int *ops, const double *vals_d,
                     const int64_t *vals_i, uint64_t *out_key)
{
    heap h;
    heap_init(&h, 3u, fel_before);
    for (uint64_t s = 0u; s < n; s++) {
        switch (ops[s]) {
        case 0: out_key[s] = heap_push(&h, 0u, vals_d[s], vals_i[s], 0, 0, 0); break;
        case 1: out_key[s] = heap_pop(&h) ? h.slot[0].key : 0u; break;
        case 2: out_key[s] = heap_remove(&h, (uint64_t)vals_i[s]) ? 1u : 0u; break;
        default: heap_free(&h); return -1;
        }
    }
    heap_free(&h);
    return 0;
}

/* ======================================================== simulation kernel */

enum { SIG_SUCCESS = 0, SIG_PREEMPTED = -1, SIG_INTERRUPTED = -2 };    /* include/cmb_process.h:59-99 */
enum { ACT_START = 1, ACT_WAKE_TIME, ACT_WAKE_RESOURCE };
enum { ST_CREATED = 0, ST_RUNNING, ST_FINISHED };

struct sim;
struct proc;
typedef void (*proc_body)(struct sim *s, struct proc *p, int64_t sig);

typedef struct proc {
    proc_body body;
    int       pc;           /* resume point */
    int       status;
    int64_t   prio;
    int       id;
    /* model locals */
    uint64_t  n_done;
    double    stamp;
    struct proc *next_free;
} proc;

typedef struct {
    heap waiting;           /* cmb_resourceguard "is a" hashheap (src/cmb_resourceguard.c:93-106) */
} guard;

typedef struct sim {
    port_rng rng;
    double   now;           /* src/cmb_event.c:39 */
    heap     fel;           /* src/cmb_event.c:44 */
    uint64_t guard_seq;     /* src/cmb_resourceguard.c:64 (thread-local; per trial here, monotone either way) */
    proc    *current;
    /* workload */
    int      model;
    uint64_t num_objects;
    double   arr_mean, srv_mean;
    /* objectqueue (src/cmb_objectqueue.c:45-52): FIFO of arrival stamps */
    double  *ring;
    uint64_t ring_cap, ring_head, ring_len;
    guard    q_front;
    /* the queue's history (model 9): cmb_timeseries_add fused with cmb_timeseries_summarize */
    bool     recording;
    uint64_t rec_n;
    double   rec_x, rec_t;
    port_wsummary hist;
    /* resourcepool (src/cmb_resourcepool.c) */
    uint64_t pool_cap, pool_in_use;
    guard    pool_guard;
    proc   **cust;
    unsigned cust_count, cust_cap;
    proc    *cust_free;
    /* results */
    port_result *res;
    uint64_t trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
} sim;

/* src/cmb_event.c:123-140 */
static uint64_t schedule(sim *s, int action, proc *subject, int64_t arg, double t, int64_t prio)
{
    return heap_push(&s->fel, 0u, t, prio, action, (int64_t)(intptr_t)subject, arg);
}

/* src/cmb_process.c:127-135 */
static void process_start(sim *s, proc *p)
{
    schedule(s, ACT_START, p, 0, s->now, p->prio);
}

/* src/cmb_process.c:262-285 + 316-333: schedule own wake-up; the caller then yields */
static void process_hold(sim *s, proc *p, double dur)
{
    schedule(s, ACT_WAKE_TIME, p, SIG_SUCCESS, s->now + dur, p->prio);
}

/* src/cmb_resourceguard.c:125-146: join the wait list; the caller then yields */
static void guard_wait(sim *s, guard *g, proc *p)
{
    const uint64_t key = ++s->guard_seq;
    heap_push(&g->waiting, key, s->now, p->prio, (int64_t)(intptr_t)p, 0, 0);
}

/* src/cmb_resourceguard.c:202-242: wake at most the head, if its demand holds */
static bool guard_signal(sim *s, guard *g, bool demand_holds)
{
    if (g->waiting.count == 0u || !demand_holds) {
        return false;
    }
    proc *p = (proc *)(intptr_t)g->waiting.slot[1].item[0];
    heap_pop(&g->waiting);
    schedule(s, ACT_WAKE_RESOURCE, p, SIG_SUCCESS, s->now, p->prio);
    return true;
}

/* src/cmb_process.c:671-684 via the trampoline (cmi_coroutine_context.asm:140-148):
 * nothing held, nothing awaited, nobody waiting in these workloads; the final
 * cmb_event_pattern_cancel(ANY, p, ANY) (src/cmb_process.c:618-619) is kept. */
static void process_exit(sim *s, proc *p)
{
    for (uint64_t k = 1u; k <= s->fel.count; ) {
        if ((proc *)(intptr_t)s->fel.slot[k].item[1] == p) {
            heap_remove(&s->fel, s->fel.slot[k].key);
            k = 1u;
        }
        else {
            k++;
        }
    }
    p->status = ST_FINISHED;
}

/* ---- cmb_objectqueue as a ring of stamps ---- */

static void ring_push(sim *s, double v)
{
    if (s->ring_len == s->ring_cap) {
        double *bigger = malloc(2u * s->ring_cap * sizeof(double));
        for (uint64_t k = 0u; k < s->ring_len; k++) {
            bigger[k] = s->ring[(s->ring_head + k) % s->ring_cap];
        }
        free(s->ring);
        s->ring = bigger;
        s->ring_head = 0u;
        s->ring_cap *= 2u;
    }
    s->ring[(s->ring_head + s->ring_len) % s->ring_cap] = v;
    s->ring_len++;
}

/* record_sample (src/cmb_objectqueue.c:151-159) -> cmb_timeseries_add (src/cmb_timeseries.c:106-141):
 * a new sample fixes the duration of the previous one, which is all cmb_timeseries_summarize
 * (:167-188) ever feeds to cmb_wtdsummary_add - so the history need not be stored */
static void record_sample(sim *s)
{
    if (!s->recording) {
        return;
    }
    if (s->rec_n > 0u) {
        (void)port_wsummary_add(&s->hist, s->rec_x, s->now - s->rec_t);
    }
    s->rec_x = (double)s->ring_len;
    s->rec_t = s->now;
    s->rec_n++;
}

/* src/cmb_objectqueue.c:262-314, unlimited capacity: append, signal the front guard */
static void objectqueue_put(sim *s, double stamp)
{
    ring_push(s, stamp);
    record_sample(s);
    if (s->ring_len > s->res->max_queue) {
        s->res->max_queue = s->ring_len;
    }
    guard_signal(s, &s->q_front, s->ring_len > 0u);     /* demand = has_content (:119-131) */
}

/* src/cmb_objectqueue.c:203-239: true if an object was taken (rear guard has no waiters) */
static bool objectqueue_try_get(sim *s, double *stamp)
{
    if (s->ring_len == 0u) {
        return false;
    }
    *stamp = s->ring[s->ring_head];
    s->ring_head = (s->ring_head + 1u) % s->ring_cap;
    s->ring_len--;
    record_sample(s);
    return true;
}

/* ---- the two queue processes: benchmark/MM1_multi.c:52-89 as resume points ---- */

static double draw_interarrival(sim *s)
{
    if (s->model == 1) {
        return port_erlang(&s->rng, 2u, 0.5 * s->arr_mean);
    }
    return port_exponential(&s->rng, s->arr_mean);
}

static double draw_service(sim *s)
{
    if (s->model == 1) {
        double v;
        do {
            v = port_normal(&s->rng, s->srv_mean, 0.25 * s->srv_mean);
        } while (v < 0.0);
        return v;
    }
    return port_exponential(&s->rng, s->srv_mean);
}

static void source_body(sim *s, proc *p, int64_t sig)
{
    (void)sig;
    switch (p->pc) {
    case 0:
        p->n_done = 0u;
        for (;;) {
            if (p->n_done >= s->num_objects) {
                process_exit(s, p);
                return;
            }
            process_hold(s, p, draw_interarrival(s));
            p->pc = 1;
            return;
    case 1:
            objectqueue_put(s, s->now);
            p->n_done++;
        }
    }
}

static void server_body(sim *s, proc *p, int64_t sig)
{
    (void)sig;
    switch (p->pc) {
    case 0:
        for (;;) {
            while (!objectqueue_try_get(s, &p->stamp)) {
                guard_wait(s, &s->q_front, p);
                p->pc = 1;
                return;
    case 1:     ;
            }
            process_hold(s, p, draw_service(s));
            p->pc = 2;
            return;
    case 2:
            s->res->sum_wait += s->now - p->stamp;
            s->res->objects += 1u;
        }
    }
}

/* ---- M/M/c: oracle/ref_build/ref_driver.c c_source_body / c_customer_body ---- */

/* src/cmb_resourcepool.c:362-533 with req = 1, no preemption, nobody interrupts */
static void customer_body(sim *s, proc *p, int64_t sig)
{
    (void)sig;
    switch (p->pc) {
    case 0:
        for (;;) {
            if (s->pool_cap - s->pool_in_use >= 1u) {
                s->pool_in_use += 1u;
                /* "in case someone else can use the leftovers" (:408-409) */
                guard_signal(s, &s->pool_guard, s->pool_cap - s->pool_in_use > 0u);
                break;
            }
            guard_wait(s, &s->pool_guard, p);
            p->pc = 1;
            return;
    case 1:     ;
        }
        process_hold(s, p, port_exponential(&s->rng, s->srv_mean));
        p->pc = 2;
        return;
    case 2:
        /* src/cmb_resourcepool.c:561-605 */
        s->pool_in_use -= 1u;
        guard_signal(s, &s->pool_guard, s->pool_cap - s->pool_in_use > 0u);
        s->res->sum_wait += s->now - p->stamp;
        s->res->objects += 1u;
        p->next_free = s->cust_free;
        s->cust_free = p;
        process_exit(s, p);
        return;
    }
}

static void generator_body(sim *s, proc *p, int64_t sig)
{
    (void)sig;
    switch (p->pc) {
    case 0:
        p->n_done = 0u;
        for (;;) {
            if (p->n_done >= s->num_objects) {
                process_exit(s, p);
                return;
            }
            process_hold(s, p, port_exponential(&s->rng, s->arr_mean));
            p->pc = 1;
            return;
    case 1: {
            proc *cu = s->cust_free;
            if (cu != NULL) {
                s->cust_free = cu->next_free;
            }
            else {
                cu = calloc(1, sizeof(*cu));
                cu->body = customer_body;
                cu->id = (int)(s->cust_count + 1u);
                if (s->cust_count == s->cust_cap) {
                    s->cust_cap *= 2u;
                    s->cust = realloc(s->cust, s->cust_cap * sizeof(proc *));
                }
                s->cust[s->cust_count++] = cu;
            }
            cu->stamp = s->now;
            process_start(s, cu);
            p->n_done++;
            }
        }
    }
}

/* ---- dispatcher: src/cmb_event.c:229-252, 259-267 ---- */

static void dispatch_all(sim *s)
{
    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > s->res->max_fel) {
            s->res->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < s->trace_cap) {
            s->trace_key[n] = ev.key;
            s->trace_time[n] = s->now;
        }
        n++;
        proc *p = (proc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:                                 /* src/cmb_process.c:115-122, src/cmi_coroutine.c:172-196 */
            p->status = ST_RUNNING;
            p->pc = 0;
            s->current = p;
            p->body(s, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:                             /* src/cmb_process.c:292-308 */
            s->current = p;
            p->body(s, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:                         /* src/cmb_resourceguard.c:168-180 */
            if (p->status == ST_RUNNING) {
                s->current = p;
                p->body(s, p, ev.item[2]);
            }
            break;
        }
        s->current = NULL;
    }
    s->res->events = n;
    s->res->t_end = s->now;
}

static void run_one(int model, int servers, uint64_t seed, uint64_t num_objects,
                    double arr_mean, double srv_mean, uint64_t trace_cap,
                    uint64_t *trace_key, double *trace_time, port_result *out)
{
    sim s;
    memset(&s, 0, sizeof(s));
    memset(out, 0, sizeof(*out));
    s.res = out;
    s.model = model;
    s.num_objects = num_objects;
    s.arr_mean = arr_mean;
    s.srv_mean = srv_mean;
    s.trace_cap = trace_cap;
    s.trace_key = trace_key;
    s.trace_time = trace_time;

    port_rng_init(&s.rng, seed);                        /* benchmark/MM1_multi.c:96 (seeded) */
    s.now = 0.0;
    heap_init(&s.fel, 3u, fel_before);                  /* src/cmb_event.c:47,74-81 */

    proc source, server;
    memset(&source, 0, sizeof(source));
    memset(&server, 0, sizeof(server));
    if (model == 2) {
        s.pool_cap = (uint64_t)servers;
        heap_init(&s.pool_guard.waiting, 3u, guard_before);
        s.cust_cap = 64u;
        s.cust = malloc(s.cust_cap * sizeof(proc *));
        source.body = generator_body;
        process_start(&s, &source);
    }
    else {
        s.ring_cap = 64u;
        s.ring = malloc(s.ring_cap * sizeof(double));
        heap_init(&s.q_front.waiting, 3u, guard_before);
        if (model == 9) {                               /* cmb_objectqueue_recording_start */
            s.recording = true;
            port_wsummary_init(&s.hist);
            record_sample(&s);
        }
        source.body = source_body;
        server.body = server_body;
        server.id = 1;
        process_start(&s, &source);                     /* benchmark/MM1_multi.c:107-111 */
        process_start(&s, &server);
    }

    dispatch_all(&s);

    if (model == 2) {
        out->max_queue = s.cust_count;
        for (unsigned k = 0u; k < s.cust_count; k++) {
            free(s.cust[k]);
        }
        free(s.cust);
        heap_free(&s.pool_guard.waiting);
    }
    else {
        if (model == 9) {                               /* cmb_objectqueue_recording_stop + summarize */
            record_sample(&s);
            const double v[7] = { s.hist.ds.min, s.hist.ds.max, s.hist.ds.m1, s.hist.ds.m2,
                                  s.hist.ds.m3, s.hist.ds.m4, s.hist.wsum };
            out->counter[0] = s.hist.ds.count;
            memcpy(&out->counter[1], v, sizeof(v));
        }
        free(s.ring);
        heap_free(&s.q_front.waiting);
    }
    heap_free(&s.fel);
}

/* ===================================================== general kernel (model 3)
 *
 * The interrupt / cancel / stop half of the reference's process layer, restated:
 * awaitable lists (src/cmb_process.c:220-260), cmb_process_hold with its
 * interrupted branch (:262-285), cmb_process_interrupt and its wake-up event
 * (:628-666), cmi_process_cancel_awaiteds (:581-620), cmb_process_stop
 * (:698-723), cmb_event_cancel / cmb_event_pattern_cancel (src/cmb_event.c:285-302,
 * 385-425), cmb_resourceguard_wait with its self-cancel (src/cmb_resourceguard.c:
 * 125-163), cmb_resourceguard_remove keyed by process address (:270-290, never
 * hits: SURVEY.md quirk 2), the bounded cmb_objectqueue with both guards
 * (src/cmb_objectqueue.c:203-314) and event/guard ordering under unequal
 * priorities (quirk 1).  Workload: oracle/ref_build/ref_driver.c model 3.
 */
enum { ACT_WAKE_INTERRUPT = 4, ACT_USER_END = 5 };
enum { AW_TIME = 0, AW_RESOURCE = 1 };
#define G_WORKERS 6
#define G_NPUT 3

struct gsim;
typedef struct gproc {
    int       pc, status, id, kind;     /* kind: 0 putter, 1 getter, 2 nuisance */
    int64_t   prio;
    /* awaits list, most recent first (cmi_slist push-front) */
    struct { int type; uint64_t handle; void *ptr; } awaits[8];
    int       n_awaits;
    uint64_t  hold_handle, guard_key;
    double    stamp;
} gproc;

typedef struct gsim {
    port_rng rng;
    double   now;
    heap     fel;
    uint64_t guard_seq;
    uint64_t current_key;               /* cmb_event_current() */
    heap     front, rear;               /* the queue's two guards */
    double  *ring;
    uint64_t cap, head, len;
    double   put_mean, get_mean;
    bool     use_pq;                    /* model 13: the objects sit in a cmb_priorityqueue (priority desc, FIFO) */
    heap     pq;
    bool     recording;                 /* models 11, 13: the queue's history, fused as in record_sample() above */
    uint64_t rec_n;
    double   rec_x, rec_t;
    port_wsummary hist;
    gproc    worker[G_WORKERS], nuisance;
    port_result *res;
    uint64_t trace_cap, *trace_key;
    double  *trace_time;
} gsim;

static void aw_push(gproc *p, int type, uint64_t handle, void *ptr)
{
    for (int k = p->n_awaits; k > 0; k--) {
        p->awaits[k] = p->awaits[k - 1];
    }
    p->awaits[0].type = type;
    p->awaits[0].handle = handle;
    p->awaits[0].ptr = ptr;
    p->n_awaits++;
}

/* src/cmb_process.c:239-260: first entry of that type (and value, unless wildcard) */
static bool aw_remove(gproc *p, int type, bool any, uint64_t handle, void *ptr)
{
    for (int k = 0; k < p->n_awaits; k++) {
        if (p->awaits[k].type == type
            && (any || (type == AW_TIME ? p->awaits[k].handle == handle : p->awaits[k].ptr == ptr))) {
            for (int m = k; m + 1 < p->n_awaits; m++) {
                p->awaits[m] = p->awaits[m + 1];
            }
            p->n_awaits--;
            return true;
        }
    }
    return false;
}

static uint64_t g_schedule(gsim *s, int action, void *subject, int64_t arg, double t, int64_t prio)
{
    return heap_push(&s->fel, 0u, t, prio, action, (int64_t)(intptr_t)subject, arg);
}

/* src/cmb_event.c:285-302 (no event waiters in this workload) */
static bool g_event_cancel(gsim *s, uint64_t handle)
{
    return heap_remove(&s->fel, handle);
}

/* src/cmb_event.c:385-425 with (ANY action, subject, ANY object): two passes */
static void g_cancel_events_of(gsim *s, void *subject)
{
    uint64_t hit[64];
    unsigned n = 0u;
    for (uint64_t k = 1u; k <= s->fel.count && n < 64u; k++) {
        if ((void *)(intptr_t)s->fel.slot[k].item[1] == subject) {
            hit[n++] = s->fel.slot[k].key;
        }
    }
    for (unsigned k = 0u; k < n; k++) {
        g_event_cancel(s, hit[k]);
    }
}

/* src/cmb_resourceguard.c:270-290: looks the waiter up by PROCESS ADDRESS although
 * entries are keyed by sequence number - reproduced literally */
static bool g_guard_remove(heap *g, gproc *p)
{
    return heap_remove(g, (uint64_t)(uintptr_t)p);
}

/* src/cmb_process.c:581-620 */
static void g_cancel_awaiteds(gsim *s, gproc *p)
{
    while (p->n_awaits > 0) {
        const int type = p->awaits[0].type;
        const uint64_t handle = p->awaits[0].handle;
        void *ptr = p->awaits[0].ptr;
        for (int m = 0; m + 1 < p->n_awaits; m++) {
            p->awaits[m] = p->awaits[m + 1];
        }
        p->n_awaits--;
        if (type == AW_TIME) {
            (void)g_event_cancel(s, handle);
        }
        else {
            (void)g_guard_remove((heap *)ptr, p);
        }
    }
    g_cancel_events_of(s, p);
}

/* cmb_process_hold up to its yield (src/cmb_process.c:262-273, 316-333) */
static void g_hold_begin(gsim *s, gproc *p, double dur)
{
    p->hold_handle = g_schedule(s, ACT_WAKE_TIME, p, SIG_SUCCESS, s->now + dur, p->prio);
    aw_push(p, AW_TIME, p->hold_handle, NULL);
}

/* ... and after it (:274-284, 338-349) */
static int64_t g_hold_end(gsim *s, gproc *p, int64_t sig)
{
    if (sig != SIG_SUCCESS) {
        (void)aw_remove(p, AW_TIME, false, p->hold_handle, NULL);
        (void)g_event_cancel(s, p->hold_handle);
        (void)aw_remove(p, AW_TIME, false, p->hold_handle, NULL);
    }
    return sig;
}

/* src/cmb_resourceguard.c:125-152 */
static void g_wait_begin(gsim *s, heap *g, gproc *p)
{
    p->guard_key = ++s->guard_seq;
    heap_push(g, p->guard_key, s->now, p->prio, (int64_t)(intptr_t)p, 0, 0);
    aw_push(p, AW_RESOURCE, 0u, g);
}

/* :153-162 */
static int64_t g_wait_end(gsim *s, heap *g, gproc *p, int64_t sig)
{
    (void)s;
    if (sig != SIG_SUCCESS) {
        (void)heap_remove(g, p->guard_key);
    }
    (void)aw_remove(p, AW_RESOURCE, false, 0u, g);
    return sig;
}

/* src/cmb_resourceguard.c:202-226 */
static void g_signal(gsim *s, heap *g, bool demand_holds)
{
    if (g->count > 0u && demand_holds) {
        gproc *p = (gproc *)(intptr_t)g->slot[1].item[0];
        heap_pop(g);
        g_schedule(s, ACT_WAKE_RESOURCE, p, SIG_SUCCESS, s->now, p->prio);
    }
}

static bool prioq_before(const heap_tag *a, const heap_tag *b);    /* src/cmb_priorityqueue.c:43-54, defined with model 6 */

/* record_sample of the bounded queue, src/cmb_objectqueue.c:151-159 */
static void g_record(gsim *s)
{
    if (!s->recording) {
        return;
    }
    if (s->rec_n > 0u) {
        (void)port_wsummary_add(&s->hist, s->rec_x, s->now - s->rec_t);
    }
    s->rec_x = (double)s->len;
    s->rec_t = s->now;
    s->rec_n++;
}

static void g_note(gsim *s, int64_t sig, unsigned which)
{
    if (sig != SIG_SUCCESS) {
        s->res->counter[which] += 1u;
        s->res->counter[5] += (uint64_t)sig;
    }
}

/* the three process bodies of ref_driver.c model 3 as resume points */
static void g_body(gsim *s, gproc *p, int64_t sig)
{
    switch (p->pc) {
    case 0:
        for (;;) {
            g_hold_begin(s, p, port_exponential(&s->rng, p->kind == 0 ? s->put_mean
                                                        : p->kind == 1 ? s->get_mean : 1.0));
            p->pc = 1;
            return;
    case 1:
            sig = g_hold_end(s, p, sig);
            if (p->kind == 2) {
                const long victim = port_dice(&s->rng, 0, G_WORKERS - 1);
                const int64_t isig = port_dice(&s->rng, 1, 10);
                const int64_t ipri = port_dice(&s->rng, -5, 5);
                s->res->counter[7] += 1u;
                /* cmb_process_interrupt, src/cmb_process.c:653-666 */
                g_schedule(s, ACT_WAKE_INTERRUPT, &s->worker[victim], isig, s->now, ipri);
                continue;
            }
            g_note(s, sig, 2u);
            if (p->kind == 0) {
                p->stamp = s->now;
                for (;;) {                              /* cmb_objectqueue_put */
                    if (s->len < s->cap) {
                        if (s->use_pq) {                /* cmb_priorityqueue_put with the putter's priority, :237-262 */
                            int64_t bits;
                            memcpy(&bits, &p->stamp, 8);
                            heap_push(&s->pq, 0u, 0.0, p->prio, bits, 0, 0);
                        }
                        else {
                            s->ring[(s->head + s->len) % s->cap] = p->stamp;
                        }
                        s->len++;
                        g_record(s);
                        g_signal(s, &s->front, s->len > 0u);
                        s->res->counter[0] += 1u;
                        break;
                    }
                    g_wait_begin(s, &s->rear, p);
                    p->pc = 2;
                    return;
    case 2:
                    sig = g_wait_end(s, &s->rear, p, sig);
                    if (sig != SIG_SUCCESS) {
                        g_note(s, sig, 3u);
                        break;
                    }
                }
            }
            else {
                for (;;) {                              /* cmb_objectqueue_get */
                    if (s->len > 0u) {
                        double stamp;
                        if (s->use_pq) {                /* cmb_priorityqueue_get, :189-212 */
                            heap_pop(&s->pq);
                            memcpy(&stamp, &s->pq.slot[0].item[0], 8);
                        }
                        else {
                            stamp = s->ring[s->head];
                            s->head = (s->head + 1u) % s->cap;
                        }
                        s->len--;
                        g_record(s);
                        g_signal(s, &s->rear, s->len < s->cap);
                        s->res->counter[1] += 1u;
                        s->res->sum_wait += s->now - stamp;
                        break;
                    }
                    g_wait_begin(s, &s->front, p);
                    p->pc = 3;
                    return;
    case 3:
                    sig = g_wait_end(s, &s->front, p, sig);
                    if (sig != SIG_SUCCESS) {
                        g_note(s, sig, 4u);
                        break;
                    }
                }
            }
        }
    }
}

/* cmb_process_stop, src/cmb_process.c:698-723 */
static void g_stop(gsim *s, gproc *p)
{
    if (p->status != ST_RUNNING) {
        return;
    }
    p->status = ST_FINISHED;
    g_cancel_awaiteds(s, p);
}

static void run_guarded(int capacity, uint64_t seed, uint64_t duration,
                        double put_mean, double get_mean, uint64_t trace_cap,
                        uint64_t *trace_key, double *trace_time, port_result *out, bool record, bool use_pq)
{
    gsim *s = calloc(1, sizeof(*s));
    memset(out, 0, sizeof(*out));
    s->res = out;
    s->put_mean = put_mean;
    s->get_mean = get_mean;
    s->trace_cap = trace_cap;
    s->trace_key = trace_key;
    s->trace_time = trace_time;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&s->front, 3u, guard_before);
    heap_init(&s->rear, 3u, guard_before);
    s->cap = (uint64_t)capacity;
    s->use_pq = use_pq;
    heap_init(&s->pq, 3u, prioq_before);
    if (record) {                                       /* cmb_objectqueue_recording_start: the empty queue at t = 0 */
        s->recording = true;
        port_wsummary_init(&s->hist);
        g_record(s);
    }
    s->ring = calloc(s->cap, sizeof(double));

    for (int i = 0; i < G_WORKERS; i++) {
        gproc *p = &s->worker[i];
        p->id = i;
        p->kind = (i < G_NPUT) ? 0 : 1;
        p->prio = port_dice(&s->rng, -5, 5);
        g_schedule(s, ACT_START, p, 0, s->now, p->prio);
    }
    s->nuisance.id = G_WORKERS;
    s->nuisance.kind = 2;
    g_schedule(s, ACT_START, &s->nuisance, 0, s->now, 0);
    g_schedule(s, ACT_USER_END, s, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        s->current_key = ev.key;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        gproc *p = (gproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->status = ST_RUNNING;
            p->pc = 0;
            g_body(s, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:                             /* src/cmb_process.c:292-308 */
            (void)aw_remove(p, AW_TIME, false, ev.key, NULL);
            g_body(s, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
            if (p->status == ST_RUNNING) {
                g_body(s, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_INTERRUPT:                        /* src/cmb_process.c:628-643 */
            g_cancel_awaiteds(s, p);
            g_body(s, p, ev.item[2]);
            break;
        case ACT_USER_END:                              /* g_end_event in ref_driver.c */
            for (int i = 0; i < G_WORKERS; i++) {
                g_stop(s, &s->worker[i]);
            }
            g_stop(s, &s->nuisance);
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->counter[6] = s->len;
    out->objects = out->counter[1];
    if (record) {                                       /* recording_stop + cmb_timeseries_summarize */
        g_record(s);
        memcpy(&out->counter[6], &s->hist.ds.m1, 8);
        out->max_queue = s->hist.ds.count;
    }
    heap_free(&s->fel);
    heap_free(&s->front);
    heap_free(&s->rear);
    free(s->ring);
    heap_free(&s->pq);
    free(s);
}

/* ======================================== model 4: resource pool with pre-emption
 *
 * cmi_pool_acquire_inner (src/cmb_resourcepool.c:362-533) in full, release (:561-605),
 * update_record (:324-355), reset_holder (:217-236), resourcepool_drop_holder
 * (:98-121), reprioritize_holder (:127-137) -> cmi_hashheap_reprioritize
 * (src/cmi_hashheap.c:679-711), holder_queue_check (:75-92),
 * cmb_process_priority_set for a running process (src/cmb_process.c:150-198),
 * cmi_process_drop_resources (:507-527).  Workload: ref_driver.c model 4.
 * Holder keys are process index + 1 (the reference keys by process address; the
 * driver allocates its processes in one array, so both orders agree).
 */
#define P_MICE 3
#define P_RODENTS 5

/* src/cmb_resourcepool.c:75-92: lowest priority first, then LARGER key first */
static bool holder_before(const heap_tag *a, const heap_tag *b)
{
    if (a->i < b->i) return true;
    if (a->i == b->i && a->key > b->key) return true;
    return false;
}

typedef struct pproc {
    gproc    g;                 /* pc, status, prio, awaits, hold_handle, guard_key */
    bool     holds_pool;        /* a cmi_process_holdable tag for the pool is on the resources list */
    uint64_t held, req, rem, initially_held;
} pproc;

typedef struct psim {
    gsim     s;                 /* rng, now, fel, guard_seq; s.front is the pool's guard */
    heap     holders;
    uint64_t cap, in_use;
    pproc    proc[P_RODENTS + 1];
} psim;

static uint64_t holder_slot(psim *w, uint64_t key)
{
    for (uint64_t k = 1u; k <= w->holders.count; k++) {
        if (w->holders.slot[k].key == key) {
            return k;
        }
    }
    return 0u;
}

static uint64_t p_held_by(psim *w, pproc *p)             /* :302-318 */
{
    const uint64_t k = holder_slot(w, (uint64_t)(p - w->proc) + 1u);
    return k ? (uint64_t)w->holders.slot[k].item[1] : 0u;
}

static void p_update_record(psim *w, pproc *p, uint64_t amount)     /* :324-355 */
{
    const uint64_t key = (uint64_t)(p - w->proc) + 1u;
    const uint64_t k = holder_slot(w, key);
    if (k != 0u) {
        w->holders.slot[k].item[1] += (int64_t)amount;
    }
    else {
        p->holds_pool = true;
        heap_push(&w->holders, key, 0.0, p->g.prio, (int64_t)(intptr_t)p, (int64_t)amount, 0);
    }
}

static void p_signal_guard(psim *w)
{
    g_signal(&w->s, &w->s.front, w->cap - w->in_use > 0u);          /* is_available, :198-211 */
}

/* cmi_hashheap_reprioritize, src/cmi_hashheap.c:679-711 */
static void heap_reprioritize(heap *h, uint64_t key, double d, int64_t i)
{
    uint64_t at = 0u;
    for (uint64_t k = 1u; k <= h->count; k++) {
        if (h->slot[k].key == key) {
            at = k;
            break;
        }
    }
    if (at == 0u) {
        return;
    }
    const heap_tag old = h->slot[at];
    h->slot[at].d = d;
    h->slot[at].i = i;
    if (h->before(&old, &h->slot[at])) {
        sift_down(h, at);
    }
    else {
        sift_up(h, at);
    }
}

/* cmb_process_priority_set for the running process (its awaits list is empty) */
static void p_priority_set(psim *w, pproc *p, int64_t pri)
{
    p->g.prio = pri;
    if (p->holds_pool) {
        heap_reprioritize(&w->holders, (uint64_t)(p - w->proc) + 1u, 0.0, pri);
    }
}

static void p_release(psim *w, pproc *p, uint64_t amount)           /* :561-605 */
{
    const uint64_t key = (uint64_t)(p - w->proc) + 1u;
    const uint64_t k = holder_slot(w, key);
    if ((uint64_t)w->holders.slot[k].item[1] == amount) {
        heap_remove(&w->holders, key);
        p->holds_pool = false;
    }
    else {
        w->holders.slot[k].item[1] -= (int64_t)amount;
    }
    w->in_use -= amount;
    p_signal_guard(w);
}

static void p_check(psim *w, pproc *p)
{
    if (p_held_by(w, p) != p->held) {
        w->s.res->counter[7] += 1u;
    }
}

static void p_take_signal(psim *w, pproc *p, int64_t sig)
{
    if (sig == SIG_PREEMPTED) {
        w->s.res->counter[2] += 1u;
        p->held = 0u;
    }
    else if (sig != SIG_SUCCESS) {
        w->s.res->counter[3] += 1u;
    }
    w->s.res->counter[4] += (uint64_t)sig;
}

static void p_rodent(psim *w, pproc *p, int64_t sig)
{
    gsim *s = &w->s;
    const bool rat = (p - w->proc) >= P_MICE;
    switch (p->g.pc) {
    case 0:
        for (;;) {
            p_check(w, p);
            p->req = (uint64_t)port_dice(&s->rng, 1, 5);
            if (!rat) {
                p_priority_set(w, p, port_dice(&s->rng, -5, 5));
            }
            /* ---- cmi_pool_acquire_inner */
            p->initially_held = p_held_by(w, p);
            p->rem = p->req;
            for (;;) {
                const uint64_t avail = w->cap - w->in_use;
                if (avail >= p->rem) {
                    w->in_use += p->rem;
                    p_update_record(w, p, p->rem);
                    p_signal_guard(w);
                    sig = SIG_SUCCESS;
                    goto acquired;
                }
                else if (avail > 0u) {
                    w->in_use += avail;
                    p->rem -= avail;
                    p_update_record(w, p, avail);
                }
                if (rat) {
                    while (w->holders.count > 0u && w->holders.slot[1].i < p->g.prio) {
                        heap_pop(&w->holders);
                        pproc *victim = (pproc *)(intptr_t)w->holders.slot[0].item[0];
                        const uint64_t loot = (uint64_t)w->holders.slot[0].item[1];
                        victim->holds_pool = false;                 /* cmi_process_remove_holdable */
                        g_schedule(s, ACT_WAKE_INTERRUPT, victim, SIG_PREEMPTED, s->now, victim->g.prio);
                        if (loot < p->rem) {
                            p_update_record(w, p, loot);
                            p->rem -= loot;
                        }
                        else {
                            p_update_record(w, p, p->rem);
                            w->in_use -= loot - p->rem;
                            p_signal_guard(w);
                            sig = SIG_SUCCESS;
                            goto acquired;
                        }
                    }
                }
                g_wait_begin(s, &s->front, &p->g);
                p->g.pc = 1;
                return;
    case 1:
                sig = g_wait_end(s, &s->front, &p->g, sig);
                if (sig == SIG_PREEMPTED) {
                    goto acquired;
                }
                else if (sig != SIG_SUCCESS) {
                    if (p->initially_held > 0u) {                   /* reset_holder, :217-236 */
                        const uint64_t k = holder_slot(w, (uint64_t)(p - w->proc) + 1u);
                        const uint64_t surplus = (uint64_t)w->holders.slot[k].item[1] - p->initially_held;
                        w->holders.slot[k].item[1] = (int64_t)p->initially_held;
                        w->in_use -= surplus;
                        p_signal_guard(w);
                    }
                    else {
                        w->in_use -= p_held_by(w, p);
                        if (heap_remove(&w->holders, (uint64_t)(p - w->proc) + 1u)) {
                            p->holds_pool = false;
                        }
                    }
                    goto acquired;
                }
            }
acquired:
            if (sig == SIG_SUCCESS) {
                p->held += p->req;
                s->res->counter[rat ? 1 : 0] += 1u;
                p_check(w, p);
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
                p->g.pc = 2;
                return;
    case 2:
                sig = g_hold_end(s, &p->g, sig);
                if (sig == SIG_SUCCESS) {
                    uint64_t rel = (uint64_t)port_dice(&s->rng, 1, 5);
                    if (rel > p->held || port_dice(&s->rng, 0, 1) == 1) {
                        rel = p->held;
                    }
                    p_release(w, p, rel);
                    p->held -= rel;
                    s->res->counter[5] += rel;
                    s->res->sum_wait += s->now * (double)rel;
                }
                else {
                    p_take_signal(w, p, sig);
                }
            }
            else {
                p_take_signal(w, p, sig);
            }
            p_check(w, p);
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 3;
            return;
    case 3:
            sig = g_hold_end(s, &p->g, sig);
            if (sig != SIG_SUCCESS) {
                p_take_signal(w, p, sig);
            }
        }
    }
}

static void p_cat(psim *w, pproc *p, int64_t sig)
{
    gsim *s = &w->s;
    switch (p->g.pc) {
    case 0:
        for (;;) {
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 1;
            return;
    case 1:
            (void)g_hold_end(s, &p->g, sig);
            {
                const long victim = port_dice(&s->rng, 0, P_RODENTS - 1);
                const int64_t loud = port_dice(&s->rng, 10, 100);
                const int64_t isig = (port_dice(&s->rng, 0, 1) == 1) ? SIG_INTERRUPTED : loud;
                g_schedule(s, ACT_WAKE_INTERRUPT, &w->proc[victim], isig, s->now, 0);
            }
        }
    }
}

static void p_resume(psim *w, pproc *p, int64_t sig)
{
    if (p - w->proc == P_RODENTS) {
        p_cat(w, p, sig);
    }
    else {
        p_rodent(w, p, sig);
    }
}

/* cmb_process_stop: cancel awaiteds, THEN drop resources (src/cmb_process.c:714-719) */
static void p_stop(psim *w, pproc *p)
{
    if (p->g.status != ST_RUNNING) {
        return;
    }
    p->g.status = ST_FINISHED;
    g_cancel_awaiteds(&w->s, &p->g);
    if (p->holds_pool) {                                /* resourcepool_drop_holder, :98-121 */
        p->holds_pool = false;
        const uint64_t key = (uint64_t)(p - w->proc) + 1u;
        const uint64_t k = holder_slot(w, key);
        if (k != 0u) {
            w->in_use -= (uint64_t)w->holders.slot[k].item[1];
            heap_remove(&w->holders, key);
            p_signal_guard(w);
        }
    }
}

static void run_preempt(int capacity, uint64_t seed, uint64_t duration, uint64_t trace_cap,
                        uint64_t *trace_key, double *trace_time, port_result *out)
{
    psim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&s->front, 3u, guard_before);
    heap_init(&w->holders, 3u, holder_before);
    w->cap = (uint64_t)capacity;

    for (int i = 0; i < P_RODENTS; i++) {
        w->proc[i].g.prio = port_dice(&s->rng, -5, 5);
        g_schedule(s, ACT_START, &w->proc[i], 0, s->now, w->proc[i].g.prio);
    }
    g_schedule(s, ACT_START, &w->proc[P_RODENTS], 0, s->now, 0);
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        pproc *p = (pproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->g.status = ST_RUNNING;
            p->g.pc = 0;
            p_resume(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(&p->g, AW_TIME, false, ev.key, NULL);
            p_resume(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
            if (p->g.status == ST_RUNNING) {
                p_resume(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_INTERRUPT:
            g_cancel_awaiteds(s, &p->g);
            p_resume(w, p, ev.item[2]);
            break;
        case ACT_USER_END:
            for (int i = 0; i <= P_RODENTS; i++) {
                p_stop(w, &w->proc[i]);
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->counter[6] = w->in_use;
    out->objects = out->counter[0] + out->counter[1];
    heap_free(&s->fel);
    heap_free(&s->front);
    heap_free(&w->holders);
    free(w);
}

/* ======================================== model 5: buffer + binary resource
 *
 * cmb_buffer_get / cmb_buffer_put with partial fulfilment (src/cmb_buffer.c:194-264,
 * 279-346) and cmb_resource acquire / release / preempt with its wake-up event and
 * drop handler (src/cmb_resource.c:45-56, 182-320).  Workload: ref_driver.c model 5.
 */
enum { ACT_WAKE_PREEMPT = 6 };
#define B_PROCS 6

typedef struct bproc {
    gproc    g;
    bool     holds_tool;
    uint64_t want, rem, moved;
    double   since;
} bproc;

typedef struct bsim {
    gsim     s;                 /* s.front / s.rear are the buffer's guards */
    heap     tool_guard;
    bproc   *tool_holder;
    uint64_t cap, level;
    double   put_mean, get_mean;
    int      fillers, drainers;         /* model 5: 2 + 2 (+ 2 tool workers); model 12 (test/test_buffer.c): 3 + 3 */
    long     amount_max;                /* 8 / 15 */
    bool     recording;                 /* model 12: the level history, fused as in record_sample() above */
    uint64_t rec_n;
    double   rec_x, rec_t;
    port_wsummary hist;
    bproc    proc[B_PROCS + 1];
} bsim;

/* record_sample, src/cmb_buffer.c:129-136 */
static void b_record(bsim *w)
{
    if (!w->recording) {
        return;
    }
    if (w->rec_n > 0u) {
        (void)port_wsummary_add(&w->hist, w->rec_x, w->s.now - w->rec_t);
    }
    w->rec_x = (double)w->level;
    w->rec_t = w->s.now;
    w->rec_n++;
}

static void b_note(bsim *w, int64_t sig)
{
    if (sig != SIG_SUCCESS) {
        w->s.res->counter[6] += (uint64_t)sig;
    }
}

static void b_body(bsim *w, bproc *p, int64_t sig)
{
    gsim *s = &w->s;
    const int id = (int)(p - w->proc);
    switch (p->g.pc) {
    case 0:
        if (id == B_PROCS) {                            /* nuisance */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
                p->g.pc = 10;
                return;
    case 10:
                (void)g_hold_end(s, &p->g, sig);
                {
                    const long victim = port_dice(&s->rng, 0, B_PROCS - 1);
                    const int64_t isig = port_dice(&s->rng, 1, 10);
                    const int64_t ipri = port_dice(&s->rng, -5, 5);
                    g_schedule(s, ACT_WAKE_INTERRUPT, &w->proc[victim], isig, s->now, ipri);
                }
            }
        }
        if (id < w->fillers) {                          /* filler */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, w->put_mean));
                p->g.pc = 20;
                return;
    case 20:
                b_note(w, g_hold_end(s, &p->g, sig));
                p->want = (uint64_t)port_dice(&s->rng, 1, w->amount_max);
                p->rem = p->want;                       /* cmb_buffer_put: *amntp and rem_claim move together */
                for (;;) {
                    if (w->cap - w->level >= p->rem) {
                        w->level += p->rem;
                        b_record(w);
                        p->rem = 0u;
                        g_signal(s, &s->front, w->level > 0u);
                        if (w->level < w->cap) {
                            g_signal(s, &s->rear, w->level < w->cap);
                        }
                        sig = SIG_SUCCESS;
                        break;
                    }
                    else if (w->level < w->cap) {
                        const uint64_t grab = w->cap - w->level;
                        w->level = w->cap;
                        b_record(w);
                        p->rem -= grab;
                        g_signal(s, &s->front, w->level > 0u);
                    }
                    g_signal(s, &s->front, w->level > 0u);
                    g_wait_begin(s, &s->rear, &p->g);
                    p->g.pc = 21;
                    return;
    case 21:
                    sig = g_wait_end(s, &s->rear, &p->g, sig);
                    if (sig != SIG_SUCCESS) {
                        break;
                    }
                }
                s->res->counter[0] += p->want - p->rem;
                if (sig != SIG_SUCCESS) {
                    s->res->counter[2] += 1u;
                    b_note(w, sig);
                }
            }
        }
        if (id < w->fillers + w->drainers) {            /* drainer */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, w->get_mean));
                p->g.pc = 30;
                return;
    case 30:
                b_note(w, g_hold_end(s, &p->g, sig));
                p->rem = (uint64_t)port_dice(&s->rng, 1, w->amount_max);
                p->moved = 0u;
                for (;;) {                              /* cmb_buffer_get */
                    if (w->level >= p->rem) {
                        w->level -= p->rem;
                        b_record(w);
                        p->moved += p->rem;
                        g_signal(s, &s->rear, w->level < w->cap);
                        if (w->level > 0u) {
                            g_signal(s, &s->front, w->level > 0u);
                        }
                        sig = SIG_SUCCESS;
                        break;
                    }
                    else if (w->level > 0u) {
                        const uint64_t grab = w->level;
                        w->level = 0u;
                        b_record(w);
                        p->moved += grab;
                        p->rem -= grab;
                        g_signal(s, &s->rear, w->level < w->cap);
                    }
                    g_signal(s, &s->rear, w->level < w->cap);
                    g_wait_begin(s, &s->front, &p->g);
                    p->g.pc = 31;
                    return;
    case 31:
                    sig = g_wait_end(s, &s->front, &p->g, sig);
                    if (sig != SIG_SUCCESS) {
                        break;
                    }
                }
                s->res->counter[1] += p->moved;
                if (sig != SIG_SUCCESS) {
                    s->res->counter[3] += 1u;
                    b_note(w, sig);
                }
            }
        }
        for (;;) {                                      /* workers 4 (polite) and 5 (pushy) */
            if (id == 5 && w->tool_holder != NULL && p->g.prio >= w->tool_holder->g.prio) {
                /* cmb_resource_preempt, kick-out branch (src/cmb_resource.c:282-299) */
                bproc *victim = w->tool_holder;
                victim->holds_tool = false;
                g_cancel_awaiteds(s, &p->g);            /* sic: the CALLER's awaiteds */
                w->tool_holder = NULL;
                g_schedule(s, ACT_WAKE_PREEMPT, victim, SIG_PREEMPTED, s->now, victim->g.prio);
                w->tool_holder = p;
                p->holds_tool = true;
                sig = SIG_SUCCESS;
            }
            else if (w->tool_holder == NULL) {          /* free: grab (acquire :196-203, preempt :277-281) */
                w->tool_holder = p;
                p->holds_tool = true;
                sig = SIG_SUCCESS;
            }
            else {                                      /* wait politely (:206-222) */
                g_wait_begin(s, &w->tool_guard, &p->g);
                p->g.pc = 40;
                return;
    case 40:
                sig = g_wait_end(s, &w->tool_guard, &p->g, sig);
                if (sig == SIG_SUCCESS) {
                    w->tool_holder = p;
                    p->holds_tool = true;
                }
            }
            if (sig == SIG_SUCCESS) {
                s->res->counter[4] += 1u;
                p->since = s->now;
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
                p->g.pc = 41;
                return;
    case 41:
                sig = g_hold_end(s, &p->g, sig);
                if (sig == SIG_PREEMPTED) {
                    s->res->counter[5] += 1u;
                    b_note(w, sig);
                }
                else {
                    b_note(w, sig);
                    p->holds_tool = false;              /* cmb_resource_release, :234-250 */
                    w->tool_holder = NULL;
                    g_signal(s, &w->tool_guard, w->tool_holder == NULL);
                    s->res->sum_wait += s->now - p->since;
                }
            }
            else {
                b_note(w, sig);
            }
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 42;
            return;
    case 42:
            b_note(w, g_hold_end(s, &p->g, sig));
        }
    }
}

static void b_stop(bsim *w, bproc *p)
{
    if (p->g.status != ST_RUNNING) {
        return;
    }
    p->g.status = ST_FINISHED;
    g_cancel_awaiteds(&w->s, &p->g);
    if (p->holds_tool) {                                /* resource_drop_holder, :45-56 */
        p->holds_tool = false;
        w->tool_holder = NULL;
        g_signal(&w->s, &w->tool_guard, true);
    }
}

static void run_buffer(int capacity, uint64_t seed, uint64_t duration, double put_mean, double get_mean,
                       uint64_t trace_cap, uint64_t *trace_key, double *trace_time, port_result *out, bool plain)
{
    bsim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    w->put_mean = put_mean;
    w->get_mean = get_mean;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&s->front, 3u, guard_before);
    heap_init(&s->rear, 3u, guard_before);
    heap_init(&w->tool_guard, 3u, guard_before);
    w->cap = (uint64_t)capacity;
    w->fillers = plain ? 3 : 2;
    w->drainers = plain ? 3 : 2;
    w->amount_max = plain ? 15 : 8;
    if (plain) {                                        /* cmb_buffer_recording_start: level 0 at t = 0 */
        w->recording = true;
        port_wsummary_init(&w->hist);
        b_record(w);
    }

    for (int i = 0; i < B_PROCS; i++) {
        w->proc[i].g.prio = port_dice(&s->rng, -5, 5);
        g_schedule(s, ACT_START, &w->proc[i], 0, s->now, w->proc[i].g.prio);
    }
    g_schedule(s, ACT_START, &w->proc[B_PROCS], 0, s->now, 0);
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        bproc *p = (bproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->g.status = ST_RUNNING;
            p->g.pc = 0;
            b_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(&p->g, AW_TIME, false, ev.key, NULL);
            b_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
        case ACT_WAKE_PREEMPT:                          /* src/cmb_resource.c:256-268: no cancel_awaiteds */
            if (p->g.status == ST_RUNNING) {
                b_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_INTERRUPT:
            g_cancel_awaiteds(s, &p->g);
            b_body(w, p, ev.item[2]);
            break;
        case ACT_USER_END:
            for (int i = 0; i <= B_PROCS; i++) {
                b_stop(w, &w->proc[i]);
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->counter[7] = w->level;
    out->objects = out->counter[1];
    if (plain) {                                        /* recording_stop + cmb_timeseries_summarize */
        b_record(w);
        memcpy(&out->counter[4], &w->hist.ds.m1, 8);
        out->max_queue = w->hist.ds.count;
    }
    heap_free(&s->fel);
    heap_free(&s->front);
    heap_free(&s->rear);
    heap_free(&w->tool_guard);
    free(w);
}

/* ======================================== model 6: priority queue + condition
 *
 * cmb_priorityqueue_put/get (src/cmb_priorityqueue.c:189-284), position (:286-320),
 * cancel / reprioritize by handle (include/cmb_priorityqueue.h:152-185), the queue's
 * order (:43-54); cmb_condition_wait / signal with its two passes over the guard in
 * array order and its own wake-up event (src/cmb_condition.c:63-167).
 * Workload: ref_driver.c model 6.
 */
enum { ACT_WAKE_CONDITION = 7 };
#define C_PROCS 7

/* src/cmb_priorityqueue.c:43-54 */
static bool prioq_before(const heap_tag *a, const heap_tag *b)
{
    if (a->i != b->i) {
        return a->i > b->i;
    }
    return a->key < b->key;
}

typedef struct cproc {
    gproc    g;
    uint64_t handle;
    int64_t  weight, pri;
} cproc;

typedef struct csim {
    gsim     s;                 /* s.front / s.rear guard the priority queue */
    heap     pq, cv;            /* cv = the condition's guard */
    uint64_t cap;
    uint64_t last_handle[2];
    long     level, threshold[2];
    double   put_mean, get_mean;
    cproc    proc[C_PROCS + 1];
} csim;

static void c_note(csim *w, int64_t sig)
{
    if (sig != SIG_SUCCESS) {
        w->s.res->counter[6] += (uint64_t)sig;
    }
}

/* cmb_condition_signal, src/cmb_condition.c:120-167 */
static uint64_t c_condition_signal(csim *w)
{
    gsim *s = &w->s;
    uint64_t hit[16];
    uint64_t cnt = 0u;
    for (uint64_t k = 1u; k <= w->cv.count; k++) {
        cproc *p = (cproc *)(intptr_t)w->cv.slot[k].item[0];
        const long thr = w->threshold[(p - w->proc) - 5];
        if (w->level >= thr) {
            hit[cnt++] = w->cv.slot[k].key;
            g_schedule(s, ACT_WAKE_CONDITION, p, SIG_SUCCESS, s->now, p->g.prio);
        }
    }
    for (uint64_t k = 0u; k < cnt; k++) {
        heap_remove(&w->cv, hit[k]);
    }
    return cnt;
}

/* cmb_priorityqueue_position, src/cmb_priorityqueue.c:286-320 */
static uint64_t c_position(csim *w, uint64_t handle)
{
    uint64_t at = 0u;
    for (uint64_t k = 1u; k <= w->pq.count; k++) {
        if (w->pq.slot[k].key == handle) {
            at = k;
            break;
        }
    }
    if (at == 0u) {
        return 0u;
    }
    uint64_t ahead = 0u;
    for (uint64_t k = 1u; k <= w->pq.count; k++) {
        if (k != at && prioq_before(&w->pq.slot[k], &w->pq.slot[at])) {
            ahead++;
        }
    }
    return ahead + 1u;
}

static void c_body(csim *w, cproc *p, int64_t sig)
{
    gsim *s = &w->s;
    const int id = (int)(p - w->proc);
    switch (p->g.pc) {
    case 0:
        if (id == C_PROCS) {                            /* nuisance */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
                p->g.pc = 10;
                return;
    case 10:
                (void)g_hold_end(s, &p->g, sig);
                {
                    const long victim = port_dice(&s->rng, 0, C_PROCS - 1);
                    const int64_t isig = port_dice(&s->rng, 1, 10);
                    const int64_t ipri = port_dice(&s->rng, -5, 5);
                    g_schedule(s, ACT_WAKE_INTERRUPT, &w->proc[victim], isig, s->now, ipri);
                }
            }
        }
        if (id < 2) {                                   /* producer */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, w->put_mean));
                p->g.pc = 20;
                return;
    case 20:
                c_note(w, g_hold_end(s, &p->g, sig));
                p->weight = port_dice(&s->rng, 1, 9);
                p->pri = port_dice(&s->rng, -3, 3);
                for (;;) {                              /* cmb_priorityqueue_put */
                    if (w->pq.count < w->cap) {
                        p->handle = heap_push(&w->pq, 0u, 0.0, p->pri, p->weight, 0, 0);
                        g_signal(s, &s->front, w->pq.count > 0u);
                        sig = SIG_SUCCESS;
                        break;
                    }
                    g_wait_begin(s, &s->rear, &p->g);
                    p->g.pc = 21;
                    return;
    case 21:
                    sig = g_wait_end(s, &s->rear, &p->g, sig);
                    if (sig != SIG_SUCCESS) {
                        break;
                    }
                }
                if (sig == SIG_SUCCESS) {
                    s->res->counter[0] += 1u;
                    w->last_handle[id] = p->handle;
                }
                else {
                    s->res->counter[2] += 1u;
                    c_note(w, sig);
                }
            }
        }
        if (id == 2) {                                  /* consumer */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, w->get_mean));
                p->g.pc = 30;
                return;
    case 30:
                c_note(w, g_hold_end(s, &p->g, sig));
                for (;;) {                              /* cmb_priorityqueue_get */
                    if (w->pq.count > 0u) {
                        heap_pop(&w->pq);
                        p->weight = w->pq.slot[0].item[0];
                        g_signal(s, &s->rear, w->pq.count < w->cap);
                        sig = SIG_SUCCESS;
                        break;
                    }
                    g_wait_begin(s, &s->front, &p->g);
                    p->g.pc = 31;
                    return;
    case 31:
                    sig = g_wait_end(s, &s->front, &p->g, sig);
                    if (sig != SIG_SUCCESS) {
                        break;
                    }
                }
                if (sig == SIG_SUCCESS) {
                    s->res->counter[1] += (uint64_t)p->weight;
                    s->res->sum_wait += s->now * (double)(uint64_t)p->weight;
                }
                else {
                    s->res->counter[2] += 1u;
                    c_note(w, sig);
                }
            }
        }
        if (id == 3) {                                  /* shuffler */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.5));
                p->g.pc = 40;
                return;
    case 40:
                c_note(w, g_hold_end(s, &p->g, sig));
                {
                    const uint64_t handle = w->last_handle[port_dice(&s->rng, 0, 1)];
                    if (handle == 0u) {
                        continue;
                    }
                    const uint64_t pos = c_position(w, handle);
                    s->res->counter[3] += pos;
                    if (pos > 0u) {
                        if (port_dice(&s->rng, 0, 1) == 1) {
                            heap_reprioritize(&w->pq, handle, 0.0, port_dice(&s->rng, -3, 3));
                        }
                        else {
                            (void)heap_remove(&w->pq, handle);
                            s->res->counter[3] += 1000u;
                        }
                    }
                }
            }
        }
        if (id == 4) {                                  /* tide */
            for (;;) {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
                p->g.pc = 50;
                return;
    case 50:
                c_note(w, g_hold_end(s, &p->g, sig));
                w->level = port_dice(&s->rng, 0, 5);
                s->res->counter[4] += c_condition_signal(w);
            }
        }
        for (;;) {                                      /* waiters 5, 6 */
            {
                bool through = true;
                while (w->level < w->threshold[id - 5]) {
                    g_wait_begin(s, &w->cv, &p->g);     /* cmb_condition_wait = cmb_resourceguard_wait */
                    p->g.pc = 60;
                    return;
    case 60:
                    through = true;
                    sig = g_wait_end(s, &w->cv, &p->g, sig);
                    if (sig != SIG_SUCCESS) {
                        c_note(w, sig);
                        through = false;
                        break;
                    }
                }
                if (through) {
                    s->res->counter[5] += 1u;
                }
            }
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 61;
            return;
    case 61:
            c_note(w, g_hold_end(s, &p->g, sig));
        }
    }
}

static void run_prioq(int capacity, uint64_t seed, uint64_t duration, double put_mean, double get_mean,
                      uint64_t trace_cap, uint64_t *trace_key, double *trace_time, port_result *out)
{
    csim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    w->put_mean = put_mean;
    w->get_mean = get_mean;
    w->threshold[0] = 2;
    w->threshold[1] = 4;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&s->front, 3u, guard_before);
    heap_init(&s->rear, 3u, guard_before);
    heap_init(&w->cv, 3u, guard_before);
    heap_init(&w->pq, 3u, prioq_before);
    w->cap = (uint64_t)capacity;

    for (int i = 0; i < C_PROCS; i++) {
        w->proc[i].g.prio = port_dice(&s->rng, -5, 5);
        g_schedule(s, ACT_START, &w->proc[i], 0, s->now, w->proc[i].g.prio);
    }
    g_schedule(s, ACT_START, &w->proc[C_PROCS], 0, s->now, 0);
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        cproc *p = (cproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->g.status = ST_RUNNING;
            p->g.pc = 0;
            c_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(&p->g, AW_TIME, false, ev.key, NULL);
            c_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
            if (p->g.status == ST_RUNNING) {
                c_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_CONDITION:                        /* src/cmb_condition.c:85-103 */
            (void)aw_remove(&p->g, AW_RESOURCE, true, 0u, NULL);
            if (p->g.status == ST_RUNNING) {
                c_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_INTERRUPT:
            g_cancel_awaiteds(s, &p->g);
            c_body(w, p, ev.item[2]);
            break;
        case ACT_USER_END:
            for (int i = 0; i <= C_PROCS; i++) {
                g_stop(s, &w->proc[i].g);
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->counter[7] = w->pq.count;
    out->objects = out->counter[0];
    heap_free(&s->fel);
    heap_free(&s->front);
    heap_free(&s->rear);
    heap_free(&w->cv);
    heap_free(&w->pq);
    free(w);
}

/* ======================================== model 7: the hold model (large event list)
 *
 * ref_driver.c model 7: `servers` workers in cmb_process_hold loops + a ticker + an end
 * event.  Every worker owns exactly one pending event, so the end event's
 * cmb_process_stop loop (cancel each hold, src/cmb_process.c:698-723) empties the list.
 */
static void run_hold(int workers, uint64_t seed, uint64_t duration, double mean,
                     uint64_t trace_cap, uint64_t *trace_key, double *trace_time, port_result *out)
{
    port_rng rng;
    heap fel;
    double now = 0.0;
    memset(out, 0, sizeof(*out));
    port_rng_init(&rng, seed);
    heap_init(&fel, 3u, fel_before);
    const int64_t ticker = workers;
    for (int64_t i = 0; i <= ticker; i++) {
        heap_push(&fel, 0u, now, 0, ACT_START, i, 0);               /* cmb_process_start */
    }
    heap_push(&fel, 0u, (double)duration, 0, ACT_USER_END, -1, 0);

    uint64_t n = 0u;
    for (;;) {
        if (fel.count > out->max_fel) {
            out->max_fel = fel.count;
        }
        if (!heap_pop(&fel)) {
            break;
        }
        const heap_tag ev = fel.slot[0];
        now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = now;
        }
        n++;
        const int64_t who = ev.item[1];
        if ((int)ev.item[0] == ACT_USER_END) {
            fel.count = 0u;                             /* every pending event is a stopped process's hold */
            continue;
        }
        if ((int)ev.item[0] == ACT_WAKE_TIME) {
            if (who == ticker) {
                out->counter[1] += 1u;
            }
            else {
                out->counter[0] += 1u;
                out->sum_wait += now;
            }
        }
        const double dur = (who == ticker) ? 1.0 : port_exponential(&rng, mean);
        heap_push(&fel, 0u, now + dur, 0, ACT_WAKE_TIME, who, SIG_SUCCESS);
    }
    out->events = n;
    out->t_end = now;
    out->objects = out->counter[0];
    heap_free(&fel);
}

/* ======================================== model 8: timers, waits, observers
 *
 * cmb_process_timer_add / timer_cancel / timers_clear / timer_set (src/cmb_process.c:316-381,
 * include/cmb_process.h:280-289), cmb_process_yield + cmb_process_resume (:751-760),
 * cmb_process_wait_process / wait_event with their wake-up events (:386-483, src/cmb_event.c:176-221),
 * cmb_process_exit (:671-684) and restart of a FINISHED process, cmb_event_reschedule /
 * reprioritize / cancel with waiter notification (src/cmb_event.c:285-344), and signal
 * forwarding to a guard's observers (src/cmb_resourceguard.c:227-239).
 * Workload: ref_driver.c model 8.
 */
enum { AW_PROCESS = 2, AW_EVENT = 3 };
enum { ACT_WAKE_PROCESS = 8, ACT_WAKE_EVENT = 9, ACT_RESUME = 10, ACT_BELL = 11 };
enum { SIG_STOPPED = -3, SIG_CANCELLED = -4, SIG_TIMEOUT = -5 };
enum { T_SIG_ALARM = 77, T_SIG_DOZE = 55, T_SIG_NUDGE = 9 };
#define T_PROCS 8

typedef struct tproc {
    gproc    g;
    bool     holds_desk;
    uint64_t timer, bell;
    double   since;
    long     jobs, j;
    int      waiters[T_PROCS];          /* processes waiting for this one, most recent first */
    int      n_waiters;
} tproc;

typedef struct tsim {
    gsim     s;
    heap     desk_guard, cv;            /* cv = the condition's guard, an observer of desk_guard */
    tproc   *desk_holder;
    uint64_t bell;
    bool     clerk_start_pending;
    struct { uint64_t key; int pid; } ew[64];   /* event waiters in registration order */
    int      n_ew;
    double   arr_mean, srv_mean;
    tproc    proc[T_PROCS];
} tsim;

static void t_note(tsim *w, int64_t sig)
{
    if (sig != SIG_SUCCESS) {
        w->s.res->counter[7] += (uint64_t)sig;
    }
}

static bool t_is_scheduled(tsim *w, uint64_t handle)
{
    for (uint64_t k = 1u; k <= w->s.fel.count; k++) {
        if (w->s.fel.slot[k].key == handle) {
            return true;
        }
    }
    return false;
}

/* wake_event_waiters, src/cmb_event.c:200-221: the list is push-front / pop-front */
static void t_wake_event_waiters(tsim *w, uint64_t key, int64_t sig)
{
    for (int k = w->n_ew - 1; k >= 0; k--) {
        if (w->ew[k].key == key) {
            tproc *p = &w->proc[w->ew[k].pid];
            g_schedule(&w->s, ACT_WAKE_EVENT, p, sig, w->s.now, p->g.prio);
            for (int m = k; m + 1 < w->n_ew; m++) {
                w->ew[m] = w->ew[m + 1];
            }
            w->n_ew--;
        }
    }
}

/* cmi_event_remove_waiter, src/cmb_event.c:486-508: first match from the head */
static void t_event_remove_waiter(tsim *w, uint64_t key, int pid)
{
    for (int k = w->n_ew - 1; k >= 0; k--) {
        if (w->ew[k].key == key && w->ew[k].pid == pid) {
            for (int m = k; m + 1 < w->n_ew; m++) {
                w->ew[m] = w->ew[m + 1];
            }
            w->n_ew--;
            return;
        }
    }
}

/* cmb_event_cancel, src/cmb_event.c:285-302 */
static bool t_event_cancel(tsim *w, uint64_t handle)
{
    if (!heap_remove(&w->s.fel, handle)) {
        return false;
    }
    t_wake_event_waiters(w, handle, SIG_CANCELLED);
    return true;
}

/* wake_process_waiters, src/cmb_process.c:485-505 */
static void t_wake_process_waiters(tsim *w, tproc *p, int64_t sig)
{
    for (int k = 0; k < p->n_waiters; k++) {
        tproc *q = &w->proc[p->waiters[k]];
        g_schedule(&w->s, ACT_WAKE_PROCESS, q, sig, w->s.now, q->g.prio);
    }
    p->n_waiters = 0;
}

/* cmi_process_cancel_awaiteds, src/cmb_process.c:581-620 */
static void t_cancel_awaiteds(tsim *w, tproc *p)
{
    gsim *s = &w->s;
    const int pid = (int)(p - w->proc);
    while (p->g.n_awaits > 0) {
        const int type = p->g.awaits[0].type;
        const uint64_t handle = p->g.awaits[0].handle;
        void *ptr = p->g.awaits[0].ptr;
        for (int m = 0; m + 1 < p->g.n_awaits; m++) {
            p->g.awaits[m] = p->g.awaits[m + 1];
        }
        p->g.n_awaits--;
        if (type == AW_TIME) {
            (void)t_event_cancel(w, handle);
        }
        else if (type == AW_RESOURCE) {
            (void)g_guard_remove((heap *)ptr, &p->g);   /* never hits: SURVEY.md quirk 2 */
        }
        else if (type == AW_PROCESS) {
            tproc *awaited = ptr;                       /* cmi_process_remove_waiter, :529-551 */
            for (int k = 0; k < awaited->n_waiters; k++) {
                if (awaited->waiters[k] == pid) {
                    for (int m = k; m + 1 < awaited->n_waiters; m++) {
                        awaited->waiters[m] = awaited->waiters[m + 1];
                    }
                    awaited->n_waiters--;
                    break;
                }
            }
        }
        else {
            t_event_remove_waiter(w, handle, pid);
        }
    }
    uint64_t hit[64];                                   /* cmb_event_pattern_cancel(ANY, p, ANY) */
    unsigned n = 0u;
    for (uint64_t k = 1u; k <= s->fel.count && n < 64u; k++) {
        if ((void *)(intptr_t)s->fel.slot[k].item[1] == (void *)p) {
            hit[n++] = s->fel.slot[k].key;
        }
    }
    for (unsigned k = 0u; k < n; k++) {
        (void)t_event_cancel(w, hit[k]);
    }
}

/* cmb_process_timer_add, src/cmb_process.c:316-333 */
static uint64_t t_timer_add(tsim *w, tproc *p, double dur, int64_t sig)
{
    const uint64_t h = g_schedule(&w->s, ACT_WAKE_TIME, p, sig, w->s.now + dur, p->g.prio);
    aw_push(&p->g, AW_TIME, h, NULL);
    return h;
}

/* cmb_process_timer_cancel, :338-349 */
static bool t_timer_cancel(tsim *w, tproc *p, uint64_t handle)
{
    (void)aw_remove(&p->g, AW_TIME, false, handle, NULL);
    return t_event_cancel(w, handle);
}

/* cmb_process_timers_clear, :354-381: list order, other awaitables skipped */
static void t_timers_clear(tsim *w, tproc *p)
{
    int k = 0;
    while (k < p->g.n_awaits) {
        if (p->g.awaits[k].type == AW_TIME) {
            const uint64_t handle = p->g.awaits[k].handle;
            for (int m = k; m + 1 < p->g.n_awaits; m++) {
                p->g.awaits[m] = p->g.awaits[m + 1];
            }
            p->g.n_awaits--;
            (void)t_event_cancel(w, handle);
        }
        else {
            k++;
        }
    }
}

/* cmb_resourceguard_signal on the desk's guard, then on its observer (the condition's guard),
 * src/cmb_resourceguard.c:202-242; both demands are "the desk has no holder" */
static void t_desk_signal(tsim *w)
{
    g_signal(&w->s, &w->desk_guard, w->desk_holder == NULL);
    g_signal(&w->s, &w->cv, w->desk_holder == NULL);
}

/* cmb_process_wait_event up to its yield, src/cmb_process.c:461-483 */
static void t_wait_event_begin(tsim *w, tproc *p, uint64_t handle)
{
    w->ew[w->n_ew].key = handle;
    w->ew[w->n_ew].pid = (int)(p - w->proc);
    w->n_ew++;
    aw_push(&p->g, AW_EVENT, handle, NULL);
}

static void t_body(tsim *w, tproc *p, int64_t sig)
{
    gsim *s = &w->s;
    tproc *clerk = &w->proc[2];
    switch (p->g.kind * 10 + p->g.pc) {
    /* ---- patients */
    case 0:
        for (;;) {
            g_hold_begin(s, &p->g, port_exponential(&s->rng, w->arr_mean));
            p->g.pc = 1;
            return;
    case 1:
            t_note(w, g_hold_end(s, &p->g, sig));
            p->timer = t_timer_add(w, p, port_exponential(&s->rng, 2.0 * w->srv_mean), SIG_TIMEOUT);
            if (w->desk_holder == NULL) {               /* cmb_resource_acquire, src/cmb_resource.c:191-229 */
                w->desk_holder = p;
                p->holds_desk = true;
                sig = SIG_SUCCESS;
            }
            else {
                g_wait_begin(s, &w->desk_guard, &p->g);
                p->g.pc = 2;
                return;
    case 2:
                sig = g_wait_end(s, &w->desk_guard, &p->g, sig);
                if (sig == SIG_SUCCESS) {
                    w->desk_holder = p;
                    p->holds_desk = true;
                }
            }
            if (sig == SIG_SUCCESS) {
                (void)t_timer_cancel(w, p, p->timer);
                s->res->counter[0] += 1u;
                p->since = s->now;
                (void)t_timer_add(w, p, port_exponential(&s->rng, 3.0), T_SIG_ALARM);
                g_hold_begin(s, &p->g, port_exponential(&s->rng, w->srv_mean));
                p->g.pc = 3;
                return;
    case 3:
                t_note(w, g_hold_end(s, &p->g, sig));
                t_timers_clear(w, p);
                p->holds_desk = false;                  /* cmb_resource_release, :234-250 */
                w->desk_holder = NULL;
                t_desk_signal(w);
                s->res->sum_wait += s->now - p->since;
                t_timers_clear(w, p);                   /* cmb_process_timer_set = clear + add */
                (void)t_timer_add(w, p, port_exponential(&s->rng, 0.3), T_SIG_DOZE);
                p->g.pc = 4;                            /* cmb_process_yield */
                return;
    case 4:
                t_note(w, sig);
                if (sig != T_SIG_DOZE) {
                    t_timers_clear(w, p);
                }
            }
            else if (sig == SIG_TIMEOUT) {
                s->res->counter[1] += 1u;
            }
            else {
                t_note(w, sig);
                t_timers_clear(w, p);
            }
        }
    /* ---- clerk */
    case 10:
        w->clerk_start_pending = false;
        p->jobs = port_dice(&s->rng, 2, 5);
        for (p->j = 0; p->j < p->jobs; p->j++) {
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 1;
            return;
    case 11:
            t_note(w, g_hold_end(s, &p->g, sig));
            if (port_dice(&s->rng, 0, 2) == 0) {        /* cmb_process_resume, :751-760 */
                tproc *tgt = &w->proc[port_dice(&s->rng, 0, 1)];
                g_schedule(s, ACT_RESUME, tgt, T_SIG_NUDGE, s->now, tgt->g.prio);
            }
            s->res->counter[3] += 1u;
        }
        /* cmb_process_exit, :671-684: nothing held, nothing awaited */
        t_cancel_awaiteds(w, p);
        t_wake_process_waiters(w, p, SIG_SUCCESS);
        p->g.status = ST_FINISHED;
        return;
    /* ---- supervisor */
    case 20:
        for (;;) {
            if (clerk->g.status == ST_FINISHED) {       /* cmb_process_wait_process, :428-452 */
                sig = SIG_SUCCESS;
            }
            else {
                aw_push(&p->g, AW_PROCESS, 0u, clerk);
                for (int k = clerk->n_waiters; k > 0; k--) {
                    clerk->waiters[k] = clerk->waiters[k - 1];
                }
                clerk->waiters[0] = (int)(p - w->proc);
                clerk->n_waiters++;
                p->g.pc = 1;
                return;
            }
            /* fall through */
    case 21:
            if (sig == SIG_SUCCESS) {
                s->res->counter[2] += 1u;
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 0.5));
                p->g.pc = 2;
                return;
    case 22:
                t_note(w, g_hold_end(s, &p->g, sig));
                if (clerk->g.status == ST_FINISHED && !w->clerk_start_pending) {
                    w->clerk_start_pending = true;
                    g_schedule(s, ACT_START, clerk, 0, s->now, clerk->g.prio);
                }
            }
            else {
                t_note(w, sig);
            }
        }
    /* ---- ringer */
    case 30:
        for (;;) {
            {
                const double when = s->now + port_exponential(&s->rng, 2.0);
                const int64_t pri = port_dice(&s->rng, -2, 2);
                p->bell = g_schedule(s, ACT_BELL, w, 0, when, pri);
                w->bell = p->bell;
            }
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 0.7));
            p->g.pc = 1;
            return;
    case 31:
            t_note(w, g_hold_end(s, &p->g, sig));
            if (t_is_scheduled(w, p->bell)) {
                const long op = port_dice(&s->rng, 0, 3);
                uint64_t at = 0u;
                for (uint64_t k = 1u; k <= s->fel.count; k++) {
                    if (s->fel.slot[k].key == p->bell) {
                        at = k;
                    }
                }
                if (op == 0) {                          /* cmb_event_reschedule, src/cmb_event.c:308-324 */
                    const double t = s->now + port_exponential(&s->rng, 1.0);
                    heap_reprioritize(&s->fel, p->bell, t, s->fel.slot[at].i);
                    s->res->counter[5] += 1u;
                }
                else if (op == 1) {                     /* cmb_event_reprioritize, :330-344 */
                    const int64_t pri = port_dice(&s->rng, -5, 5);
                    heap_reprioritize(&s->fel, p->bell, s->fel.slot[at].d, pri);
                    s->res->counter[5] += 100u;
                }
                else if (op == 2) {
                    (void)t_event_cancel(w, p->bell);
                    s->res->counter[5] += 10000u;
                }
            }
            if (t_is_scheduled(w, p->bell)) {
                t_wait_event_begin(w, p, p->bell);
                p->g.pc = 2;
                return;
    case 32:
                t_note(w, sig);
            }
        }
    /* ---- listener */
    case 40:
        for (;;) {
            p->bell = w->bell;
            if (p->bell != 0u && t_is_scheduled(w, p->bell)) {
                t_wait_event_begin(w, p, p->bell);
                p->g.pc = 1;
                return;
    case 41:
                if (sig == SIG_SUCCESS) {
                    s->res->counter[6] += 1000u;
                }
                else {
                    t_note(w, sig);
                }
            }
            else {
                g_hold_begin(s, &p->g, port_exponential(&s->rng, 0.5));
                p->g.pc = 2;
                return;
    case 42:
                t_note(w, g_hold_end(s, &p->g, sig));
            }
        }
    /* ---- watcher: cmb_condition_wait = cmb_resourceguard_wait on the condition's guard */
    case 50:
        for (;;) {
            g_wait_begin(s, &w->cv, &p->g);
            p->g.pc = 1;
            return;
    case 51:
            sig = g_wait_end(s, &w->cv, &p->g, sig);
            if (sig == SIG_SUCCESS) {
                s->res->counter[6] += 1u;
            }
            else {
                t_note(w, sig);
            }
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 0.8));
            p->g.pc = 2;
            return;
    case 52:
            t_note(w, g_hold_end(s, &p->g, sig));
        }
    /* ---- nuisance */
    case 60:
        for (;;) {
            g_hold_begin(s, &p->g, port_exponential(&s->rng, 1.0));
            p->g.pc = 1;
            return;
    case 61:
            (void)g_hold_end(s, &p->g, sig);
            {
                tproc *victim = &w->proc[port_dice(&s->rng, 0, T_PROCS - 2)];
                const int64_t isig = port_dice(&s->rng, 1, 10);
                const int64_t ipri = port_dice(&s->rng, -5, 5);
                if (victim->g.status == ST_RUNNING) {
                    g_schedule(s, ACT_WAKE_INTERRUPT, victim, isig, s->now, ipri);
                }
            }
        }
    }
}

/* cmb_process_stop, src/cmb_process.c:698-723 */
static void t_stop(tsim *w, tproc *p)
{
    if (p->g.status != ST_RUNNING) {
        return;
    }
    p->g.status = ST_FINISHED;
    t_cancel_awaiteds(w, p);
    if (p->holds_desk) {                                /* resource_drop_holder, src/cmb_resource.c:45-56 */
        p->holds_desk = false;
        w->desk_holder = NULL;
        t_desk_signal(w);
    }
    t_wake_process_waiters(w, p, SIG_STOPPED);
}

static void run_timers(uint64_t seed, uint64_t duration, double arr_mean, double srv_mean,
                       uint64_t trace_cap, uint64_t *trace_key, double *trace_time, port_result *out)
{
    static const int kind_of[T_PROCS] = { 0, 0, 1, 2, 3, 4, 5, 6 };
    tsim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    w->arr_mean = arr_mean;
    w->srv_mean = srv_mean;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&w->desk_guard, 3u, guard_before);
    heap_init(&w->cv, 3u, guard_before);

    for (int i = 0; i < T_PROCS; i++) {
        w->proc[i].g.kind = kind_of[i];
        w->proc[i].g.prio = (i + 1 < T_PROCS) ? port_dice(&s->rng, -5, 5) : 0;
        g_schedule(s, ACT_START, &w->proc[i], 0, s->now, w->proc[i].g.prio);
    }
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        t_wake_event_waiters(w, ev.key, SIG_SUCCESS);   /* src/cmb_event.c:243-246 */
        tproc *p = (tproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->g.status = ST_RUNNING;
            p->g.pc = 0;
            t_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(&p->g, AW_TIME, false, ev.key, NULL);
            t_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
            if (p->g.status == ST_RUNNING) {
                t_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_PROCESS:                          /* src/cmb_process.c:386-410 */
            (void)aw_remove(&p->g, AW_PROCESS, true, 0u, NULL);
            if (p->g.status == ST_RUNNING) {
                t_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_EVENT:                            /* src/cmb_event.c:176-198 */
            (void)aw_remove(&p->g, AW_EVENT, true, 0u, NULL);
            if (p->g.status == ST_RUNNING) {
                t_body(w, p, ev.item[2]);
            }
            break;
        case ACT_RESUME:                                /* src/cmb_process.c:731-745 */
            t_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_INTERRUPT:
            t_cancel_awaiteds(w, p);
            t_body(w, p, ev.item[2]);
            break;
        case ACT_BELL:
            out->counter[4] += 1u;
            break;
        case ACT_USER_END:
            for (int i = 0; i < T_PROCS; i++) {
                t_stop(w, &w->proc[i]);
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->objects = out->counter[0];
    heap_free(&s->fel);
    heap_free(&w->desk_guard);
    heap_free(&w->cv);
    free(w);
}

/* ======================================== model 10: the harbor (test/test_condition.c)
 *
 * ref_driver.c model 10 restated: dynamically created ship processes, a cmb_condition with
 * user predicates signalled every hour by two processes (evaluate-all, two passes, in heap
 * array order: src/cmb_condition.c:120-167), three cmb_resourcepools with greedy partial
 * grabs (src/cmb_resourcepool.c:362-533) and their histories folded into time-weighted
 * summaries, a second condition handing finished ships to a collector process, and an end
 * event that stops everything - ships in arrival order - leaving the stale guard entries of
 * SURVEY.md quirk 2 behind.
 */
#ifndef M_PI
#define M_PI 3.14159265358979323846     /* the reference's value (glibc math.h under _POSIX_C_SOURCE) */
#endif
enum { HB_WEATHER = 0, HB_TIDE, HB_ARRIVALS, HB_DEPARTURES, HB_DOTS, HB_SHIP };

typedef struct hship {
    gproc    g;
    uint64_t id;
    unsigned size, need, held_tugs, held_berth, rem;
    double   max_wind, min_depth, t_arr, t_sys;
    bool     tugs_listed_first;         /* order of the process's resources list (push-front) */
    bool     active;
    struct hship *next_departed, *next_all;
} hship;

typedef struct {
    uint64_t cap, in_use;
    heap     guard;
    port_wsummary hist;                 /* cmb_timeseries_add fused with cmb_timeseries_summarize */
    double   x, t;
    uint64_t n;
} hpool;

typedef struct hsim {
    gsim     s;
    double   wind_magnitude, wind_direction, water_depth;
    hpool    tugs, berth[2];
    heap     harbormaster, davyjones;
    hship    fixed[5];                  /* weather, tide, arrivals, departures, dots */
    hship   *all;                       /* every ship struct still allocated */
    hship   *departed;
    uint64_t next_id, alive, most_alive;
    double   arr_mean, unload_small;
    port_summary through[2];
} hsim;

static void hb_record(hsim *w, hpool *p)                /* record_sample, src/cmb_resourcepool.c:239-247 */
{
    if (p->n > 0u) {
        (void)port_wsummary_add(&p->hist, p->x, w->s.now - p->t);
    }
    p->x = (double)p->in_use;
    p->t = w->s.now;
    p->n++;
}

static void hb_pool_signal(hsim *w, hpool *p)
{
    g_signal(&w->s, &p->guard, p->cap - p->in_use > 0u);
}

/* one pass of cmi_pool_acquire_inner's loop (no pre-emption): true when the claim is filled */
static bool hb_pool_grab(hsim *w, hpool *p, unsigned *rem, unsigned *held)
{
    const uint64_t available = p->cap - p->in_use;
    if (available >= *rem) {
        p->in_use += *rem;
        hb_record(w, p);
        *held += *rem;
        *rem = 0u;
        hb_pool_signal(w, p);
        return true;
    }
    if (available > 0u) {
        p->in_use += available;
        hb_record(w, p);
        *rem -= (unsigned)available;
        *held += (unsigned)available;
    }
    return false;
}

static void hb_pool_release(hsim *w, hpool *p, unsigned amount, unsigned *held)     /* :561-605 */
{
    *held -= amount;
    p->in_use -= amount;
    hb_record(w, p);
    hb_pool_signal(w, p);
}

static bool hb_can_dock(const hsim *w, const hship *sh)
{
    if (w->water_depth < sh->min_depth) {
        return false;
    }
    if (w->wind_magnitude > sh->max_wind) {
        return false;
    }
    if (w->tugs.cap - w->tugs.in_use < sh->need) {
        return false;
    }
    return w->berth[sh->size].cap - w->berth[sh->size].in_use >= 1u;
}

/* cmb_condition_signal, src/cmb_condition.c:120-167 */
static uint64_t hb_condition_signal(hsim *w, heap *cv, bool is_harbormaster)
{
    gsim *s = &w->s;
    uint64_t *hit = malloc((cv->count + 1u) * sizeof(*hit));
    uint64_t cnt = 0u;
    for (uint64_t k = 1u; k <= cv->count; k++) {
        hship *p = (hship *)(intptr_t)cv->slot[k].item[0];
        const bool ok = is_harbormaster ? hb_can_dock(w, p) : (w->departed != NULL);
        if (ok) {
            hit[cnt++] = cv->slot[k].key;
            g_schedule(s, ACT_WAKE_CONDITION, p, SIG_SUCCESS, s->now, p->g.prio);
        }
    }
    for (uint64_t k = 0u; k < cnt; k++) {
        heap_remove(cv, hit[k]);
    }
    free(hit);
    return cnt;
}

static double hb_pert(hsim *w, double min, double mode, double max)
{
    return port_PERT_mod(&w->s.rng, min, mode, max, 4.0);
}

static void hb_body(hsim *w, hship *p, int64_t sig)
{
    gsim *s = &w->s;
    (void)sig;
    switch (p->g.kind * 10 + p->g.pc) {
    case HB_WEATHER * 10:
        for (;;) {
            {
                const double gust = port_rayleigh(&s->rng, 5.0);
                w->wind_magnitude = 0.5 * gust + 0.5 * w->wind_magnitude;
                const double d1 = hb_pert(w, 0.0, 225.0, 360.0);
                const double d2 = hb_pert(w, 0.0, 45.0, 360.0);
                w->wind_direction = 0.75 * d1 + 0.25 * d2;
                s->res->counter[7] += hb_condition_signal(w, &w->harbormaster, true);
            }
            g_hold_begin(s, &p->g, 1.0);
            p->g.pc = 1;
            return;
    case HB_WEATHER * 10 + 1:
            (void)g_hold_end(s, &p->g, sig);
        }
    case HB_TIDE * 10:
        for (;;) {
            {
                const double half_month = 0.5 * 29.5 * 24.0;
                const double t = fmod(s->now, half_month);
                const double astro = 15.0 + 1.0 * sin(2.0 * M_PI * t / 12.4) + 0.5 * sin(2.0 * M_PI * t / 24.0)
                                   + 0.25 * sin(2.0 * M_PI * t / (0.5 * 29.5 * 24));
                const double surge = 0.5 * w->wind_magnitude
                                   - 0.5 * w->wind_magnitude * sin(w->wind_direction * M_PI / 180.0);
                w->water_depth = astro + surge;
                s->res->counter[7] += hb_condition_signal(w, &w->harbormaster, true);
            }
            g_hold_begin(s, &p->g, 1.0);
            p->g.pc = 1;
            return;
    case HB_TIDE * 10 + 1:
            (void)g_hold_end(s, &p->g, sig);
        }
    case HB_ARRIVALS * 10:
        for (;;) {
            g_hold_begin(s, &p->g, port_exponential(&s->rng, w->arr_mean));
            p->g.pc = 1;
            return;
    case HB_ARRIVALS * 10 + 1:
            (void)g_hold_end(s, &p->g, sig);
            {
                hship *sh = calloc(1, sizeof(*sh));
                sh->id = ++w->next_id;
                sh->size = port_bernoulli(&s->rng, 0.25);
                sh->max_wind = (sh->size == 0u) ? 10.0 : 12.0;
                sh->min_depth = (sh->size == 0u) ? 8.0 : 13.0;
                sh->need = (sh->size == 0u) ? 1u : 3u;
                sh->g.kind = HB_SHIP;
                sh->next_all = w->all;
                w->all = sh;
                g_schedule(s, ACT_START, sh, 0, s->now, 0);
            }
        }
    case HB_DEPARTURES * 10:
        for (;;) {
            g_wait_begin(s, &w->davyjones, &p->g);
            p->g.pc = 1;
            return;
    case HB_DEPARTURES * 10 + 1:
            (void)g_wait_end(s, &w->davyjones, &p->g, sig);
            {
                hship *sh = w->departed;
                w->departed = sh->next_departed;
                (void)port_summary_add(&w->through[sh->size], sh->t_sys);
                s->res->sum_wait += sh->t_sys;
                s->res->counter[sh->size] += 1u;
                for (hship **pp = &w->all; *pp != NULL; pp = &(*pp)->next_all) {
                    if (*pp == sh) {
                        *pp = sh->next_all;
                        break;
                    }
                }
                free(sh);
            }
        }
    case HB_DOTS * 10:
        for (;;) {
            g_hold_begin(s, &p->g, 24.0 * 7 * 52);
            p->g.pc = 1;
            return;
    case HB_DOTS * 10 + 1:
            (void)g_hold_end(s, &p->g, sig);
        }
    case HB_SHIP * 10:
        p->t_arr = s->now;
        p->active = true;
        if (++w->alive > w->most_alive) {
            w->most_alive = w->alive;
        }
        while (!hb_can_dock(w, p)) {
            g_wait_begin(s, &w->harbormaster, &p->g);
            p->g.pc = 1;
            return;
    case HB_SHIP * 10 + 1:
            (void)g_wait_end(s, &w->harbormaster, &p->g, sig);
        }
        p->rem = 1u;                                    /* both are there: the predicate just said so */
        (void)hb_pool_grab(w, &w->berth[p->size], &p->rem, &p->held_berth);
        p->rem = p->need;
        (void)hb_pool_grab(w, &w->tugs, &p->rem, &p->held_tugs);
        g_hold_begin(s, &p->g, hb_pert(w, 0.4, 0.5, 0.8));
        p->g.pc = 2;
        return;
    case HB_SHIP * 10 + 2:
        (void)g_hold_end(s, &p->g, sig);
        hb_pool_release(w, &w->tugs, p->need, &p->held_tugs);
        {
            const double tua = (p->size == 0u) ? w->unload_small : 1.5 * w->unload_small;
            g_hold_begin(s, &p->g, hb_pert(w, 0.75 * tua, tua, 2 * tua));
        }
        p->g.pc = 3;
        return;
    case HB_SHIP * 10 + 3:
        (void)g_hold_end(s, &p->g, sig);
        p->rem = p->need;
        while (!hb_pool_grab(w, &w->tugs, &p->rem, &p->held_tugs)) {
            g_wait_begin(s, &w->tugs.guard, &p->g);
            p->g.pc = 4;
            return;
    case HB_SHIP * 10 + 4:
            (void)g_wait_end(s, &w->tugs.guard, &p->g, sig);
        }
        g_hold_begin(s, &p->g, hb_pert(w, 0.4, 0.5, 0.8));
        p->g.pc = 5;
        return;
    case HB_SHIP * 10 + 5:
        (void)g_hold_end(s, &p->g, sig);
        hb_pool_release(w, &w->berth[p->size], 1u, &p->held_berth);
        hb_pool_release(w, &w->tugs, p->need, &p->held_tugs);
        p->active = false;
        w->alive--;
        p->next_departed = w->departed;
        w->departed = p;
        (void)hb_condition_signal(w, &w->davyjones, false);
        p->t_sys = s->now - p->t_arr;
        p->g.status = ST_FINISHED;                      /* return -> cmb_process_exit: nothing held or awaited */
        return;
    }
}

/* cmb_process_stop, src/cmb_process.c:698-723, incl. cmi_process_drop_resources (:507-527) */
static void hb_stop(hsim *w, hship *p)
{
    if (p->g.status != ST_RUNNING) {
        return;
    }
    p->g.status = ST_FINISHED;
    g_cancel_awaiteds(&w->s, &p->g);
    /* resources list: a pool is pushed to the front when its first unit is taken */
    if (p->held_tugs > 0u) {
        w->tugs.in_use -= p->held_tugs;
        p->held_tugs = 0u;
        hb_pool_signal(w, &w->tugs);
    }
    if (p->held_berth > 0u) {
        w->berth[p->size].in_use -= p->held_berth;
        p->held_berth = 0u;
        hb_pool_signal(w, &w->berth[p->size]);
    }
}

static void hb_pool_init(hpool *p, uint64_t cap)
{
    p->cap = cap;
    heap_init(&p->guard, 3u, guard_before);
    port_wsummary_init(&p->hist);
}

static void run_harbor(int tugs, uint64_t seed, uint64_t duration, double arr_mean, double unload_small,
                       uint64_t trace_cap, uint64_t *trace_key, double *trace_time, port_result *out)
{
    hsim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    w->arr_mean = arr_mean;
    w->unload_small = unload_small;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&w->harbormaster, 3u, guard_before);
    heap_init(&w->davyjones, 3u, guard_before);
    hb_pool_init(&w->tugs, (uint64_t)tugs);
    hb_pool_init(&w->berth[0], 6u);
    hb_pool_init(&w->berth[1], 3u);
    port_summary_init(&w->through[0]);
    port_summary_init(&w->through[1]);

    /* creation order of test/test_condition.c:523-586 fixes the event keys */
    for (int i = 0; i < 5; i++) {
        w->fixed[i].g.kind = i;
    }
    g_schedule(s, ACT_START, &w->fixed[HB_WEATHER], 0, s->now, 0);
    g_schedule(s, ACT_START, &w->fixed[HB_TIDE], 0, s->now, 0);
    hb_record(w, &w->tugs);                             /* cmb_resourcepool_start_recording */
    hb_record(w, &w->berth[0]);
    hb_record(w, &w->berth[1]);
    g_schedule(s, ACT_START, &w->fixed[HB_ARRIVALS], 0, s->now, 0);
    g_schedule(s, ACT_START, &w->fixed[HB_DEPARTURES], 0, s->now, 0);
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);
    g_schedule(s, ACT_START, &w->fixed[HB_DOTS], 0, s->now, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        hship *p = (hship *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->g.status = ST_RUNNING;
            p->g.pc = 0;
            hb_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(&p->g, AW_TIME, false, ev.key, NULL);
            hb_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
            if (p->g.status == ST_RUNNING) {
                hb_body(w, p, ev.item[2]);
            }
            break;
        case ACT_WAKE_CONDITION:
            (void)aw_remove(&p->g, AW_RESOURCE, true, 0u, NULL);
            if (p->g.status == ST_RUNNING) {
                hb_body(w, p, ev.item[2]);
            }
            break;
        case ACT_USER_END:
            for (int i = 0; i < 5; i++) {
                hb_stop(w, &w->fixed[i]);
            }
            for (;;) {                                  /* active ships in (arrival time, id) order = id order */
                hship *first = NULL;
                for (hship *q = w->all; q != NULL; q = q->next_all) {
                    if (q->active && (first == NULL || q->id < first->id)) {
                        first = q;
                    }
                }
                if (first == NULL) {
                    break;
                }
                first->active = false;
                hb_stop(w, first);
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    out->objects = out->counter[0] + out->counter[1];
    out->max_queue = w->most_alive;
    memcpy(&out->counter[2], &w->through[0].m1, 8);
    memcpy(&out->counter[3], &w->through[1].m1, 8);
    out->counter[4] = w->tugs.hist.ds.count;
    memcpy(&out->counter[5], &w->tugs.hist.ds.m1, 8);
    out->counter[6] = w->berth[0].hist.ds.count | (w->berth[1].hist.ds.count << 32);

    while (w->all != NULL) {
        hship *q = w->all;
        w->all = q->next_all;
        free(q);
    }
    heap_free(&s->fel);
    heap_free(&w->harbormaster);
    heap_free(&w->davyjones);
    heap_free(&w->tugs.guard);
    heap_free(&w->berth[0].guard);
    heap_free(&w->berth[1].guard);
    free(w);
}

/* ======================================== model 14: test/test_resource.c as it stands
 *
 * cmb_resource_acquire / release / preempt with wakeup_event_preempt and the usage history
 * (src/cmb_resource.c:45-57, 107-136, 182-320).  Workload: ref_driver.c model 14.
 */
typedef struct rsim {
    gsim     s;
    heap     guard;
    gproc   *holder;
    gproc    proc[4];
    bool     holds[4];
    double   since[4];
    uint64_t rec_n;
    double   rec_x, rec_t;
    port_wsummary hist;
} rsim;

static void r_record(rsim *w)                           /* record_sample, src/cmb_resource.c:107-118 */
{
    if (w->rec_n > 0u) {
        (void)port_wsummary_add(&w->hist, w->rec_x, w->s.now - w->rec_t);
    }
    w->rec_x = (w->holder != NULL) ? 1.0 : 0.0;
    w->rec_t = w->s.now;
    w->rec_n++;
}

static void r_grab(rsim *w, gproc *p)
{
    w->holder = p;
    w->holds[p - w->proc] = true;
}

static void r_release(rsim *w, gproc *p)                /* cmb_resource_release, :234-250 */
{
    w->holds[p - w->proc] = false;
    w->holder = NULL;
    r_record(w);
    g_signal(&w->s, &w->guard, w->holder == NULL);
}

static void r_body(rsim *w, gproc *p, int64_t sig)
{
    gsim *s = &w->s;
    const int id = (int)(p - w->proc);
    switch (p->pc) {
    case 0:
        if (id < 3) {                                   /* preemptable */
            for (;;) {
                if (w->holder == NULL) {                /* cmb_resource_acquire, :191-229 */
                    r_grab(w, p);
                    r_record(w);
                    sig = SIG_SUCCESS;
                }
                else {
                    g_wait_begin(s, &w->guard, p);
                    p->pc = 1;
                    return;
    case 1:
                    sig = g_wait_end(s, &w->guard, p, sig);
                    if (sig == SIG_SUCCESS) {
                        r_grab(w, p);
                        r_record(w);
                    }
                }
                if (sig == SIG_SUCCESS) {
                    s->res->counter[0] += 1u;
                    w->since[id] = s->now;
                    g_hold_begin(s, p, port_exponential(&s->rng, 1.0));
                    p->pc = 2;
                    return;
    case 2:
                    sig = g_hold_end(s, p, sig);
                    if (sig == SIG_SUCCESS) {
                        r_release(w, p);
                        s->res->sum_wait += s->now - w->since[id];
                    }
                    else {
                        s->res->counter[1] += 1u;
                        if (s->res->counter[5] == 0u) {
                            memcpy(&s->res->counter[4], &s->now, 8);
                            s->res->counter[5] = (uint64_t)id + 1u;
                        }
                    }
                }
                g_hold_begin(s, p, port_exponential(&s->rng, 1.0));
                p->pc = 3;
                return;
    case 3:
                (void)g_hold_end(s, p, sig);
            }
        }
        for (;;) {                                      /* preempter */
            if (w->holder == NULL) {                    /* cmb_resource_preempt, :270-320 */
                r_grab(w, p);
                r_record(w);
            }
            else if (p->prio >= w->holder->prio) {
                gproc *victim = w->holder;
                w->holds[victim - w->proc] = false;
                g_cancel_awaiteds(s, p);                /* sic: the CALLER's awaiteds */
                w->holder = NULL;
                g_schedule(s, ACT_WAKE_PREEMPT, victim, SIG_PREEMPTED, s->now, victim->prio);
                r_grab(w, p);
            }
            else {
                g_wait_begin(s, &w->guard, p);
                p->pc = 10;
                return;
    case 10:
                sig = g_wait_end(s, &w->guard, p, sig);
                if (sig == SIG_SUCCESS) {
                    r_grab(w, p);
                    r_record(w);
                }
            }
            s->res->counter[2] += 1u;
            g_hold_begin(s, p, port_exponential(&s->rng, 1.0));
            p->pc = 11;
            return;
    case 11:
            (void)g_hold_end(s, p, sig);
            r_release(w, p);
            g_hold_begin(s, p, port_exponential(&s->rng, 1.0));
            p->pc = 12;
            return;
    case 12:
            (void)g_hold_end(s, p, sig);
        }
    }
}

static void run_resource(uint64_t seed, uint64_t duration, uint64_t trace_cap, uint64_t *trace_key,
                         double *trace_time, port_result *out)
{
    rsim *w = calloc(1, sizeof(*w));
    gsim *s = &w->s;
    memset(out, 0, sizeof(*out));
    s->res = out;
    port_rng_init(&s->rng, seed);
    heap_init(&s->fel, 3u, fel_before);
    heap_init(&w->guard, 3u, guard_before);
    port_wsummary_init(&w->hist);
    r_record(w);                                        /* cmb_resource_start_recording: idle at t = 0 */
    for (int i = 0; i < 4; i++) {
        w->proc[i].prio = (i < 3) ? port_dice(&s->rng, -5, 5) : 0;
        g_schedule(s, ACT_START, &w->proc[i], 0, s->now, w->proc[i].prio);
    }
    g_schedule(s, ACT_USER_END, w, 0, (double)duration, 0);

    uint64_t n = 0u;
    for (;;) {
        if (s->fel.count > out->max_fel) {
            out->max_fel = s->fel.count;
        }
        if (!heap_pop(&s->fel)) {
            break;
        }
        const heap_tag ev = s->fel.slot[0];
        s->now = ev.d;
        if (n < trace_cap) {
            trace_key[n] = ev.key;
            trace_time[n] = s->now;
        }
        n++;
        gproc *p = (gproc *)(intptr_t)ev.item[1];
        switch ((int)ev.item[0]) {
        case ACT_START:
            p->status = ST_RUNNING;
            p->pc = 0;
            r_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_TIME:
            (void)aw_remove(p, AW_TIME, false, ev.key, NULL);
            r_body(w, p, ev.item[2]);
            break;
        case ACT_WAKE_RESOURCE:
        case ACT_WAKE_PREEMPT:
            if (p->status == ST_RUNNING) {
                r_body(w, p, ev.item[2]);
            }
            break;
        case ACT_USER_END:
            for (int i = 0; i < 4; i++) {               /* cmb_process_stop + resource_drop_holder, :45-57 */
                gproc *q = &w->proc[i];
                if (q->status != ST_RUNNING) {
                    continue;
                }
                q->status = ST_FINISHED;
                g_cancel_awaiteds(s, q);
                if (w->holds[i]) {
                    w->holds[i] = false;
                    w->holder = NULL;
                    r_record(w);
                    g_signal(s, &w->guard, true);
                }
            }
            break;
        }
    }
    out->events = n;
    out->t_end = s->now;
    r_record(w);                                        /* cmb_resource_stop_recording */
    memcpy(&out->counter[3], &w->hist.ds.m1, 8);
    out->max_queue = w->hist.ds.count;
    out->objects = out->counter[0] + out->counter[2];
    heap_free(&s->fel);
    heap_free(&w->guard);
    free(w);
}

/* ------------------------------------------------- experiment executive */

typedef struct {
    int model, servers;
    uint64_t master_seed, first, count, num_objects;
    double arr_mean, srv_mean;
    port_result *out;
    uint64_t next;          /* shared work counter, as src/cimba.c:112-118 */
} job;

static void *worker(void *arg)
{
    job *j = arg;
    for (;;) {
        const uint64_t k = __atomic_fetch_add(&j->next, 1u, __ATOMIC_RELAXED);
        if (k >= j->count) {
            break;
        }
        if (j->model == 14) {
            run_resource(port_fmix64(j->master_seed, j->first + k), j->num_objects, 0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 10) {
            run_harbor(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                       j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 8) {
            run_timers(port_fmix64(j->master_seed, j->first + k), j->num_objects,
                       j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 7) {
            run_hold(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                     j->arr_mean, 0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 6) {
            run_prioq(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                      j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 5 || j->model == 12) {
            run_buffer(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                       j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k], j->model == 12);
            continue;
        }
        if (j->model == 4) {
            run_preempt(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                        0u, NULL, NULL, &j->out[k]);
            continue;
        }
        if (j->model == 3 || j->model == 11 || j->model == 13) {
            run_guarded(j->servers, port_fmix64(j->master_seed, j->first + k), j->num_objects,
                        j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k], j->model != 3, j->model == 13);
            continue;
        }
        run_one(j->model, j->servers, port_fmix64(j->master_seed, j->first + k),
                j->num_objects, j->arr_mean, j->srv_mean, 0u, NULL, NULL, &j->out[k]);
    }
    return NULL;
}

int port_run_trials(int model, int servers, uint64_t master_seed,
                    uint64_t first, uint64_t count, uint64_t num_objects,
                    double arr_mean, double srv_mean, int threads,
                    port_result *out)
{
    job j = { model, servers, master_seed, first, count, num_objects,
              arr_mean, srv_mean, out, 0u };
    if (threads <= 1) {
        worker(&j);
        return 0;
    }
    pthread_t *tid = malloc((size_t)threads * sizeof(*tid));
    for (int t = 0; t < threads; t++) {
        pthread_create(&tid[t], NULL, worker, &j);
    }
    for (int t = 0; t < threads; t++) {
        pthread_join(tid[t], NULL);
    }
    free(tid);
    return 0;
}

int port_trace_trial(int model, int servers, uint64_t seed, uint64_t num_objects,
                     double arr_mean, double srv_mean, uint64_t trace_cap,
                     uint64_t *trace_key, double *trace_time, port_result *out)
{
    if (model == 14) {
        run_resource(seed, num_objects, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 10) {
        run_harbor(servers, seed, num_objects, arr_mean, srv_mean, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 8) {
        run_timers(seed, num_objects, arr_mean, srv_mean, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 7) {
        run_hold(servers, seed, num_objects, arr_mean, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 6) {
        run_prioq(servers, seed, num_objects, arr_mean, srv_mean, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 5 || model == 12) {
        run_buffer(servers, seed, num_objects, arr_mean, srv_mean, trace_cap, trace_key, trace_time, out, model == 12);
        return 0;
    }
    if (model == 4) {
        run_preempt(servers, seed, num_objects, trace_cap, trace_key, trace_time, out);
        return 0;
    }
    if (model == 3 || model == 11 || model == 13) {
        run_guarded(servers, seed, num_objects, arr_mean, srv_mean, trace_cap, trace_key, trace_time, out, model != 3,
                    model == 13);
        return 0;
    }
    run_one(model, servers, seed, num_objects, arr_mean, srv_mean,
            trace_cap, trace_key, trace_time, out);
    return 0;
}
