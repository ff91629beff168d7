/*
 * cimba_port.h - TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * Plain-C, CPU-only restatement of the hot path of ambonvik/cimba
 * (SURVEY.md section 8a).  Every function in cimba_port.c cites the reference
 * file:line it follows.  It is the checker for the CUDA engine; nothing under
 * cimba_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  oracle/Makefile compiles the unmodified reference
 * into oracle/_ref/ (all 10 reference golden files reproduce byte-for-byte),
 * and tests/test_oracle_*.py hold this restatement to (a) the committed golden
 * vectors in tests/golden/ generated from that build and (b) the live _ref
 * build whenever it is present.
 */
#ifndef CIMBA_PORT_H
#define CIMBA_PORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- cmb_random (src/cmb_random.c, include/cmb_random.h) ---- */
typedef struct { uint64_t a, b, c, d; } port_rng;

uint64_t port_fmix64(uint64_t seed, uint64_t nonce);
void     port_rng_init(port_rng *r, uint64_t seed);
uint64_t port_sfc64(port_rng *r);
double   port_random(port_rng *r);
double   port_uniform(port_rng *r, double lo, double hi);
double   port_std_exponential(port_rng *r);
double   port_exponential(port_rng *r, double mean);
double   port_erlang(port_rng *r, unsigned k, double m);
double   port_std_normal(port_rng *r);
double   port_normal(port_rng *r, double mu, double sigma);
unsigned port_bernoulli(port_rng *r, double p);
long     port_dice(port_rng *r, long a, long b);

/* same kinds as ref_rng_draws() in oracle/ref_build/ref_driver.c */
int port_rng_draws(uint64_t seed, int kind, double p0, double p1, uint64_t n, double *out);
/* kinds 9..33 = the rest of cmb_random, same numbering as ref_rng_draws_ex() */
int port_rng_draws_ex(uint64_t seed, int kind, const double *par, uint32_t npar, uint64_t n, double *out);

/* ---- cmb_datasummary / cmb_wtdsummary ---- */
typedef struct { uint64_t count; double min, max, m1, m2, m3, m4; } port_summary;
typedef struct { port_summary ds; double wsum; } port_wsummary;

void     port_summary_init(port_summary *s);
uint64_t port_summary_add(port_summary *s, double y);
uint64_t port_summary_merge(port_summary *tgt, const port_summary *a, const port_summary *b);
void     port_wsummary_init(port_wsummary *s);
uint64_t port_wsummary_add(port_wsummary *s, double x, double w);
uint64_t port_wsummary_merge(port_wsummary *tgt, const port_wsummary *a, const port_wsummary *b);

/* flat exports {count,min,max,m1..m4[,wsum]} matching the ref_driver helpers */
int port_datasummary_of(const double *x, uint64_t n, double *out);
int port_datasummary_split_merge(const double *x, uint64_t na, uint64_t n, double *out);
int port_wtdsummary_of(const double *x, const double *w, uint64_t n, double *out);
int port_wtdsummary_split_merge(const double *x, const double *w, uint64_t na, uint64_t n, double *out);

/* ---- models (the trial-level oracle) ---- */
typedef struct {
    uint64_t events;      /* successful cmb_event_execute_next() calls */
    uint64_t objects;     /* customers served */
    double   t_end;       /* cmb_time() when the event list ran dry */
    double   sum_wait;    /* sum of time-in-system */
    uint64_t max_fel;     /* deepest future-event list seen */
    uint64_t max_queue;   /* longest queue (models 0,1) / process structs created (model 2) */
    uint64_t counter[8];  /* model 3: puts, gets, interrupted holds/puts/gets, signal sum, final length, interrupts */
} port_result;

/* model 0 = M/M/1, 1 = G/G/1 (erlang-2 / truncated normal), 2 = M/M/c pool,
 * 3 = bounded queue with interrupts (num_objects = duration, servers = capacity) */
int port_run_trials(int model, int servers, uint64_t master_seed,
                    uint64_t first, uint64_t count, uint64_t num_objects,
                    double arr_mean, double srv_mean, int threads,
                    port_result *out);

int port_trace_trial(int model, int servers, uint64_t seed, uint64_t num_objects,
                     double arr_mean, double srv_mean, uint64_t trace_cap,
                     uint64_t *trace_key, double *trace_time, port_result *out);

/* ---- cmi_hashheap script runner (for heap-order KATs) ----
 * ops[i]: 0 = enqueue(d=vals_d[i], i=vals_i[i]) -> out_key[i] = issued key
 *         1 = dequeue -> out_key[i] = popped key (0 if empty)
 *         2 = cancel key vals_i[i] -> out_key[i] = 1 if found
 */
int port_heap_script(uint64_t n, const int *ops, const double *vals_d,
                     const int64_t *vals_i, uint64_t *out_key);

#ifdef __cplusplus
}
#endif
#endif
