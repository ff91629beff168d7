/* oracle/ref_build/awacs_driver.c - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's AWACS model (tutorial/tut_5_1.c, BASELINE config 5) compiled UNMODIFIED from where it
 * lies - the #include below is the whole of it - behind awacs_stubs/hdf5.h, which turns the HDF5 output
 * into constants and redirects four names to the hooks defined here:
 *   cmb_event_queue_execute -> awacs_counting_execute  (the same loop, counting and tracing the pops, and
 *                                                       taking a snapshot of the targets before run_trial
 *                                                       frees them)
 *   cmi_calloc              -> awacs_noting_calloc     (remembers where run_trial put its targets)
 *   printf                  -> nothing, fopen -> /dev/null, main -> awacs_reference_main (never called).
 * Entry points: awacs_ref_terrain() builds the shared terrain the way the tutorial's main() does
 * (cmb_random_initialize(seed); terrain_init), awacs_ref_trial() seeds the stream and calls the model's own
 * run_trial().  The map size and the duration are arguments (main() hard-codes 1000 x 1000 nm and 24 h;
 * the number of targets is the source's NUM_TARGETS).
 */
#include "tutorial/tut_5_1.c"

#undef cmb_event_queue_execute
#undef cmi_calloc
#undef printf
#undef fopen
#undef main

#include <stdlib.h>
#include <string.h>

struct awacs_out {
    uint64_t events;            /* cmb_event_execute_next() calls that returned true */
    double   t_end;             /* cmb_time() when the list ran dry */
    uint32_t num_found;         /* struct trial.num_found */
    uint32_t tds_count[6];      /* targets per enum target_detect_state at the end */
    uint32_t mode_count[4];     /* targets per enum target_mode at the end */
    uint32_t pad;
    double   sum_x, sum_y;      /* sum of the targets' last recorded positions */
};

/* per-thread capture area (one pointer of TLS: the reference library's own initial-exec TLS leaves little room
 * in a dlopen'ed object) */
struct awacs_capture {
    uint64_t pops;
    double t_end;
    struct target *targets;
    uint64_t trace_cap;
    uint64_t *trace_key;
    double *trace_time;
    float x[NUM_TARGETS], y[NUM_TARGETS];
    int mode[NUM_TARGETS], tds[NUM_TARGETS], det[NUM_TARGETS];
};
static _Thread_local struct awacs_capture *t_cap;
static struct terrain *g_terrain, *g_owned;   /* the map in use; the one this driver allocated itself */

static struct awacs_capture *capture(void)
{
    if (t_cap == NULL) {
        t_cap = calloc(1, sizeof(*t_cap));      /* lives as long as its thread */
        if (t_cap == NULL) abort();
    }
    return t_cap;
}

void *awacs_noting_calloc(unsigned long long n, unsigned long long sz)
{
    void *p = calloc(n, sz);
    if (p == NULL) abort();
    if (n == NUM_TARGETS && sz == sizeof(struct target)) capture()->targets = p;
    return p;
}

void awacs_counting_execute(void)
{
    struct awacs_capture *c = capture();
    while (cmb_event_execute_next()) {
        if (c->pops < c->trace_cap) {
            c->trace_key[c->pops] = cmb_event_current();
            c->trace_time[c->pops] = cmb_time();
        }
        c->pops++;
    }
    c->t_end = cmb_time();
    for (unsigned i = 0; i < NUM_TARGETS && c->targets != NULL; i++) {
        c->x[i] = c->targets[i].x_m;
        c->y[i] = c->targets[i].y_m;
        c->mode[i] = (int)c->targets[i].mode;
        c->tds[i] = (int)c->targets[i].tds;
        c->det[i] = c->targets[i].detected ? 1 : 0;
    }
}

int awacs_ref_num_targets(void) { return NUM_TARGETS; }

/* main()'s preamble (tut_5_1.c:1260-1277) with the seed and the map size as arguments */
int awacs_ref_terrain(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                      uint32_t *cols, uint32_t *rows, float *geom /* x_scale y_scale x_min x_max y_min y_max */)
{
    if (g_owned != NULL) {
        terrain_terminate(g_owned);
        terrain_destroy(g_owned);
        g_owned = NULL;
    }
    cmb_random_initialize(seed);
    g_owned = terrain_create();
    g_terrain = g_owned;
    terrain_init(g_terrain, width_nm, height_nm, ref_lat, ref_lon);
    cmb_random_terminate();
    *cols = g_terrain->cols;
    *rows = g_terrain->rows;
    geom[0] = g_terrain->x_scale; geom[1] = g_terrain->y_scale;
    geom[2] = g_terrain->x_min;   geom[3] = g_terrain->x_max;
    geom[4] = g_terrain->y_min;   geom[5] = g_terrain->y_max;
    return 0;
}

/* Use a map made elsewhere (the restatement's threaded generator produces the identical grid in a fraction of the
 * time, tests/test_oracle_awacs.py) - for timing runs only.  The caller keeps `map` alive. */
int awacs_ref_adopt_terrain(float *map, uint32_t cols, uint32_t rows, const float *geom)
{
    static struct terrain adopted;
    adopted.cols = cols;
    adopted.rows = rows;
    adopted.map = map;
    adopted.x_scale = geom[0]; adopted.y_scale = geom[1];
    adopted.x_min = geom[2];   adopted.x_max = geom[3];
    adopted.y_min = geom[4];   adopted.y_max = geom[5];
    g_terrain = &adopted;
    return 0;
}

const float *awacs_ref_map(void) { return g_terrain ? g_terrain->map : NULL; }
const int *awacs_ref_blueprint(void) { return g_terrain ? g_terrain->p : NULL; }

static void one_trial(uint64_t seed, double duration_h, uint64_t trace_cap, uint64_t *trace_key, double *trace_time,
                      struct awacs_out *out)
{
    struct awacs_capture *c = capture();
    c->pops = 0u;
    c->targets = NULL;
    c->trace_cap = trace_cap;
    c->trace_key = trace_key;
    c->trace_time = trace_time;
    struct trial trl = {};
    trl.terrain = g_terrain;
    trl.duration = duration_h;
    trl.seed_used = seed;
    cmb_random_initialize(seed);
    run_trial(&trl);                    /* ends with cmb_random_terminate() */
    memset(out, 0, sizeof(*out));
    out->events = c->pops;
    out->t_end = c->t_end;
    out->num_found = trl.num_found;
    for (unsigned i = 0; i < NUM_TARGETS; i++) {
        out->tds_count[c->tds[i]]++;
        out->mode_count[c->mode[i]]++;
        out->sum_x += c->x[i];
        out->sum_y += c->y[i];
    }
}

int awacs_ref_trial(uint64_t seed, double duration_h, uint64_t trace_cap, uint64_t *trace_key, double *trace_time,
                    struct awacs_out *out, float *x, float *y, int *mode, int *tds, int *detected)
{
    if (g_terrain == NULL) return -1;
    one_trial(seed, duration_h, trace_cap, trace_key, trace_time, out);
    const struct awacs_capture *c = capture();
    for (unsigned i = 0; i < NUM_TARGETS; i++) {
        if (x) x[i] = c->x[i];
        if (y) y[i] = c->y[i];
        if (mode) mode[i] = c->mode[i];
        if (tds) tds[i] = c->tds[i];
        if (detected) detected[i] = c->det[i];
    }
    return 0;
}

/* `count` trials through the reference's OWN executive, cimba_run_experiment (src/cimba.c:151-188: one pthread per
 * logical core pulling trial indices), seeds cmb_random_fmix64(master_seed, first + i) */
struct awacs_job { uint64_t seed; double duration_h; struct awacs_out out; };

static void awacs_job_func(void *vjob)
{
    struct awacs_job *j = vjob;
    one_trial(j->seed, j->duration_h, 0u, NULL, NULL, &j->out);
}

int awacs_ref_experiment(uint64_t master_seed, uint64_t first, uint64_t count, double duration_h, struct awacs_out *out)
{
    if (g_terrain == NULL || count == 0u) return -1;
    struct awacs_job *jobs = calloc(count, sizeof(*jobs));
    if (jobs == NULL) return -2;
    for (uint64_t i = 0; i < count; i++) {
        jobs[i].seed = cmb_random_fmix64(master_seed, first + i);
        jobs[i].duration_h = duration_h;
    }
    cimba_run_experiment(jobs, count, sizeof(*jobs), awacs_job_func);
    for (uint64_t i = 0; i < count; i++) out[i] = jobs[i].out;
    free(jobs);
    return 0;
}

/* the geometry chain on fixed inputs, for piecewise checks of a restatement (static functions of the source) */
void awacs_ref_platform_state(double t, float *six /* x y dir rol vel alt */, float *rad_eff)
{
    struct racetrack rt;
    racetrack_initialize(&rt, 0.0f, 30.0f, -10.0f, 0.0f, 50.0f, 10.0f, 310.0f, 300.0f, true);
    struct platform_state st;
    platform_state_update(&st, &rt, t);
    six[0] = st.x; six[1] = st.y; six[2] = st.dir; six[3] = st.rol; six[4] = st.vel; six[5] = st.alt;
    *rad_eff = rt.rad_eff;
}
