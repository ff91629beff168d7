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

static uint64_t g_pops;
static double g_t_end;
static struct target *g_targets;
static uint64_t g_trace_cap;
static uint64_t *g_trace_key;
static double *g_trace_time;
static float g_x[NUM_TARGETS], g_y[NUM_TARGETS], g_alt[NUM_TARGETS];
static int g_mode[NUM_TARGETS], g_tds[NUM_TARGETS], g_det[NUM_TARGETS];
static struct terrain *g_terrain;

void *awacs_noting_calloc(unsigned long long n, unsigned long long sz)
{
    void *p = calloc(n, sz);
    if (p == NULL) abort();
    if (n == NUM_TARGETS && sz == sizeof(struct target)) g_targets = p;
    return p;
}

void awacs_counting_execute(void)
{
    while (cmb_event_execute_next()) {
        if (g_pops < g_trace_cap) {
            g_trace_key[g_pops] = cmb_event_current();
            g_trace_time[g_pops] = cmb_time();
        }
        g_pops++;
    }
    g_t_end = cmb_time();
    for (unsigned i = 0; i < NUM_TARGETS && g_targets != NULL; i++) {
        g_x[i] = g_targets[i].x_m;
        g_y[i] = g_targets[i].y_m;
        g_alt[i] = g_targets[i].alt_m;
        g_mode[i] = (int)g_targets[i].mode;
        g_tds[i] = (int)g_targets[i].tds;
        g_det[i] = g_targets[i].detected ? 1 : 0;
    }
}

int awacs_ref_num_targets(void) { return NUM_TARGETS; }

/* main()'s preamble (tut_5_1.c:1260-1277) with the seed and the map size as arguments */
int awacs_ref_terrain(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                      uint32_t *cols, uint32_t *rows, float *geom /* x_scale y_scale x_min x_max y_min y_max */)
{
    if (g_terrain != NULL) {
        terrain_terminate(g_terrain);
        terrain_destroy(g_terrain);
        g_terrain = NULL;
    }
    cmb_random_initialize(seed);
    g_terrain = terrain_create();
    terrain_init(g_terrain, width_nm, height_nm, ref_lat, ref_lon);
    cmb_random_terminate();
    *cols = g_terrain->cols;
    *rows = g_terrain->rows;
    geom[0] = g_terrain->x_scale; geom[1] = g_terrain->y_scale;
    geom[2] = g_terrain->x_min;   geom[3] = g_terrain->x_max;
    geom[4] = g_terrain->y_min;   geom[5] = g_terrain->y_max;
    return 0;
}

const float *awacs_ref_map(void) { return g_terrain ? g_terrain->map : NULL; }
const int *awacs_ref_blueprint(void) { return g_terrain ? g_terrain->p : NULL; }

int awacs_ref_trial(uint64_t seed, double duration_h, uint64_t trace_cap, uint64_t *trace_key, double *trace_time,
                    struct awacs_out *out, float *x, float *y, int *mode, int *tds, int *detected)
{
    if (g_terrain == NULL) return -1;
    g_pops = 0u;
    g_targets = NULL;
    g_trace_cap = trace_cap;
    g_trace_key = trace_key;
    g_trace_time = trace_time;
    struct trial trl = {};
    trl.terrain = g_terrain;
    trl.duration = duration_h;
    trl.seed_used = seed;
    cmb_random_initialize(seed);
    run_trial(&trl);                    /* ends with cmb_random_terminate() */
    memset(out, 0, sizeof(*out));
    out->events = g_pops;
    out->t_end = g_t_end;
    out->num_found = trl.num_found;
    for (unsigned i = 0; i < NUM_TARGETS; i++) {
        out->tds_count[g_tds[i]]++;
        out->mode_count[g_mode[i]]++;
        out->sum_x += g_x[i];
        out->sum_y += g_y[i];
        if (x) x[i] = g_x[i];
        if (y) y[i] = g_y[i];
        if (mode) mode[i] = g_mode[i];
        if (tds) tds[i] = g_tds[i];
        if (detected) detected[i] = g_det[i];
    }
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
