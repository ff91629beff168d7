/*
 * awacs_port.c - TEST INFRASTRUCTURE ONLY (the parity oracle; see cimba_port.h).
 *
 * Plain-C restatement of the reference's AWACS model, tutorial/tut_5_1.c (BASELINE config 5), without the
 * coroutines and without the HDF5/VTK output: 1000 ground targets cycling hiding -> staging -> firing ->
 * driving with exponential / Erlang dwell times (:385-446), one radar ticking every second whose sweep runs a
 * five-stage detection chain over all targets in float32 (:944-1036, :548-689), an airborne platform on a
 * racetrack orbit evaluated in closed form (:724-782, :816-868), a progress-bar process that holds 100 times
 * (:1118-1137) and an end event that stops everybody (:1100-1112).  Processes are resume points in a table;
 * the future event list is a binary heap on (time, key) - every priority is 0 - and because keys are unique
 * the pop order is the total order of src/cmi_hashheap.c:55-80 whatever the heap's shape.
 *
 * Every float / double choice below follows the source expression by expression: the chain consumes the
 * trial's random stream (cmb_random_bernoulli at :688) only for targets that survive the float32 geometry, so
 * a single different rounding changes every later variate of the trial.  On the CPU this file and the
 * unmodified source (oracle/_ref/libawacs_ref.so) use the same libm and agree bit for bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "cimba_port.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define AW_TARGETS 1000                 /* NUM_TARGETS, tut_5_1.c:35 */

/* ------------------------------------------------------------------ constants, tut_5_1.c:37-66 */
static const float  k_time_step = 1.0f;
static const double k_arcsec_m = 30.87;
static const double k_nm_m = 1852.0;
static const double k_feet_m = 0.3048;
static const double k_knots_ms = (1852.0 / 3600.0);
static const double k_deg_rad = (2.0 * M_PI / 360.0);
static const double k_wgs84_a = 6378137.0;
static const double k_wgs84_f = (1.0 / 298.257223563);
#define K_WGS84_E2 (k_wgs84_f * (2.0 - k_wgs84_f))
static const float  k_terrain_max = 2500.0f;
static const unsigned k_octaves = 6u;
static const float  k_initfreq = 1.0f / 100000.0f;
static const float  k_ridginess = 1.2f;
static const float  k_peakiness = 1.7f;
static const float  k_terrain_sd = 3.0f;
static const float  k_biome_vis[3] = { 0.2f, 0.3f, 0.9f };
static const float  k_biome_elev[2] = { 400.0f, 1000.0f };

/* ------------------------------------------------------------------ terrain, tut_5_1.c:96-338 */
typedef struct {
    float x_scale, y_scale, x_min, x_max, y_min, y_max;
    uint32_t cols, rows;
    const float *map;
} aw_terrain;

static float fade(float t)                         /* :124-127 */
{
    return t * t * t * (t * (t * 6 - 15) + 10);
}

static float mix(float t, float a, float b)        /* :130-133 */
{
    return a + t * (b - a);
}

static float corner(int hash, float x, float y)    /* :136-143 */
{
    const int h = hash & 15;
    const float u = h < 8 ? x : y;
    const float v = h < 4 ? y : ((h == 12 || h == 14) ? x : 0);
    return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}

static float noise2d(const int *p, float x, float y)    /* :165-189 */
{
    const int X = (int)floorf(x) & 255;
    const int Y = (int)floorf(y) & 255;
    x -= floorf(x);
    y -= floorf(y);
    const float u = fade(x);
    const float v = fade(y);
    const int A = p[X] + Y, AA = p[A], AB = p[A + 1];
    const int B = p[X + 1] + Y, BA = p[B], BB = p[B + 1];
    return mix(v, mix(u, corner(p[AA], x, y), corner(p[BA], x - 1, y)),
                  mix(u, corner(p[AB], x, y - 1), corner(p[BB], x - 1, y - 1)));
}

/* grid size of a width_nm x height_nm map at one arc-second, :208-209 */
void port_awacs_grid(float width_nm, float height_nm, uint32_t *cols, uint32_t *rows)
{
    *cols = (unsigned int)(width_nm * k_nm_m / k_arcsec_m);
    *rows = (unsigned int)(height_nm * k_nm_m / k_arcsec_m);
}

typedef struct { const int *p; float *map; uint32_t cols, rows; float x_scale, y_scale; unsigned first, stride; } fill_job;

/* the deterministic part of terrain_init's cell loop (:241-267): six octaves of ridged noise, h * terrain_max */
static void *fill_rows(void *arg)
{
    const fill_job *j = arg;
    for (unsigned row = j->first; row < j->rows; row += j->stride) {
        const float ys = ((float)row - (float)j->rows / 2.0f) * j->y_scale;
        for (unsigned col = 0; col < j->cols; col++) {
            const float xs = ((float)col - (float)j->cols / 2.0f) * j->x_scale;
            float h = 0.0f, freq = k_initfreq, amp = 1.0f, weight = 1.0f, ampsum = 0.0f;
            for (unsigned i = 0; i < k_octaves; i++) {
                float n = noise2d(j->p, xs * freq, ys * freq);
                n = powf(1.0f - fabsf(n), k_ridginess);
                h += n * amp * weight;
                weight = n;
                freq *= 2.05f;
                ampsum += amp;
                amp *= 0.5f;
            }
            h /= ampsum;
            h = powf(h, k_peakiness);
            j->map[(size_t)row * j->cols + col] = (h * k_terrain_max);
        }
    }
    return NULL;
}

/* terrain_init, :197-294, after cmb_random_initialize(seed) as main() does (:1270-1274).
 * map: cols * rows floats owned by the caller; geom = {x_scale, y_scale, x_min, x_max, y_min, y_max} */
int port_awacs_terrain_mt(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                          float *map, float *geom, int *blueprint_out, int threads);

int port_awacs_terrain(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                       float *map, float *geom, int *blueprint_out)
{
    return port_awacs_terrain_mt(seed, width_nm, height_nm, ref_lat, ref_lon, map, geom, blueprint_out, 1);
}

/* the same map with the noise computed by `threads` workers (the tutorial's 60 000 x 60 000 grid takes a
 * quarter of an hour on one core) */
int port_awacs_terrain_mt(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                          float *map, float *geom, int *blueprint_out, int threads)
{
    (void)ref_lon;
    port_rng rng;
    port_rng_init(&rng, seed);
    uint32_t cols, rows;
    port_awacs_grid(width_nm, height_nm, &cols, &rows);
    const float ref_lat_r = (float)k_deg_rad * ref_lat;

    const double sin_lat = sinf(ref_lat_r);
    const double cos_lat = cosf(ref_lat_r);
    const double common = 1.0 - (K_WGS84_E2 * sin_lat * sin_lat);
    const double sqrt_common = sqrt(common);
    const double radius_ew = k_wgs84_a / sqrt_common;
    const double m_per_deg_ew = radius_ew * cos_lat * (M_PI / 180.0);
    const float x_scale = (float)((1.0 / 3600.0) * m_per_deg_ew);
    const double radius_ns = k_wgs84_a * (1.0 - K_WGS84_E2) / (common * sqrt_common);
    const double m_per_deg_ns = radius_ns * (M_PI / 180.0);
    const float y_scale = (float)((1.0 / 3600.0) * m_per_deg_ns);
    const float x_span = (float)(cols - 1) * x_scale;
    const float y_span = (float)(rows - 1) * y_scale;
    geom[0] = x_scale;
    geom[1] = y_scale;
    geom[2] = -(x_span / 2.0f);
    geom[3] = (x_span / 2.0f);
    geom[4] = -(y_span / 2.0f);
    geom[5] = (y_span / 2.0f);

    int p[512];                                 /* terrain_generate_blueprint, :146-162 */
    for (int i = 0; i < 256; i++) p[i] = i;
    for (int i = 255; i > 0; i--) {
        const int j = (int)port_uniform(&rng, 0, i + 1);
        const int t = p[i];
        p[i] = p[j];
        p[j] = t;
    }
    for (int i = 0; i < 256; i++) p[256 + i] = p[i];
    if (blueprint_out) memcpy(blueprint_out, p, sizeof(p));

    /* The noise part of a cell does not touch the random stream, the final cmb_random_normal does: rows are filled by
     * `threads` workers first (any order), then one pass adds the normals in the reference's cell order. */
    fill_job job = { p, map, cols, rows, x_scale, y_scale, 0u, 1u };
    if (threads <= 1) {
        fill_rows(&job);
    }
    else {
        if (threads > 256) threads = 256;
        pthread_t tid[256];
        fill_job jobs[256];
        for (int t = 0; t < threads; t++) {
            jobs[t] = job;
            jobs[t].first = (unsigned)t;
            jobs[t].stride = (unsigned)threads;
            pthread_create(&tid[t], NULL, fill_rows, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    }
    const size_t cells = (size_t)rows * cols;
    for (size_t c = 0; c < cells; c++) {
        float h_sum = map[c] + (float)port_normal(&rng, 0.0, k_terrain_sd);
        h_sum = (h_sum < 0.0f) ? 0.0f : h_sum;
        map[c] = h_sum;
    }
    return 0;
}

static float elevation(const aw_terrain *t, float x, float y)      /* terrain_index + terrain_elevation, :314-338 */
{
    const int raw_col = (int)roundf(x / t->x_scale) + (int)(t->cols / 2);
    const int raw_row = (int)roundf(y / t->y_scale) + (int)(t->rows / 2);
    const unsigned col = (unsigned)((raw_col < 0) ? 0 : (raw_col >= (int)t->cols ? (int)t->cols - 1 : raw_col));
    const unsigned row = (unsigned)((raw_row < 0) ? 0 : (raw_row >= (int)t->rows ? (int)t->rows - 1 : raw_row));
    return t->map[row * t->cols + col];
}

/* ------------------------------------------------------------------ racetrack + platform, :724-868 */
typedef struct {
    float start_time, orientation_r, length_m, turn_radius_m, altitude_m, velocity_ms;
    float turn_dist_m, orbit_dist_m, side, roll_angle_r, rad_eff;
} aw_orbit;

typedef struct { float x, y, dir, rol, vel, alt; } aw_platform;

static void orbit_init(aw_orbit *o)             /* racetrack_initialize with run_trial's arguments, :1177-1188 */
{
    const float start_time = 0.0f, anchor_lat = 30.0f, orientation = 0.0f, leg_length = 50.0f;
    const float turn_radius = 10.0f, flight_level = 310.0f, velocity = 300.0f;
    const bool clockwise = true;
    o->start_time = 3600.0f * start_time;
    const float anchor_lat_r = (float)(anchor_lat * k_deg_rad);
    o->orientation_r = (float)((90.0 - orientation) * k_deg_rad);
    o->length_m = (float)(leg_length * k_nm_m);
    o->turn_radius_m = (float)(turn_radius * k_nm_m);
    o->altitude_m = (float)(flight_level * 100.0 * k_feet_m);
    o->velocity_ms = (float)(velocity * k_knots_ms);
    o->turn_dist_m = M_PI * o->turn_radius_m;
    o->orbit_dist_m = 2.0f * (o->length_m + o->turn_dist_m);
    o->side = clockwise ? -1.0f : 1.0f;
    const double sin_lat = sinf(anchor_lat_r);
    const double common = 1.0 - (K_WGS84_E2 * sin_lat * sin_lat);
    const double sqrt_common = sqrt(common);
    const double M = k_wgs84_a * (1.0 - K_WGS84_E2) / (common * sqrt_common);
    const double N = k_wgs84_a / sqrt_common;
    const double g = 9.80665;
    const double roll_mag = atan((o->velocity_ms * o->velocity_ms) / (o->turn_radius_m * g));
    o->roll_angle_r = (float)(roll_mag * -o->side);
    const double mean_radius = sqrt(M * N);
    o->rad_eff = (float)(mean_radius * (4.0 / 3.0));
}

static void platform_at(aw_platform *st, const aw_orbit *o, double t)      /* platform_state_update, :816-868 */
{
    const double delta_t = t - o->start_time;
    double d = fmod(delta_t * o->velocity_ms, o->orbit_dist_m);
    if (d < 0) d += o->orbit_dist_m;
    double xl, yl, hdg, roll;
    if (d < o->length_m) {
        xl = d; yl = 0.0; hdg = 0.0; roll = 0.0;
    }
    else if (d < o->length_m + o->turn_dist_m) {
        const double phi = (d - o->length_m) / o->turn_radius_m - M_PI / 2.0;
        xl = o->length_m + o->turn_radius_m * cos(phi);
        yl = o->side * o->turn_radius_m * (1.0 + sin(phi));
        hdg = (phi + M_PI / 2.0) * o->side;
        roll = o->roll_angle_r;
    }
    else if (d < 2.0 * o->length_m + o->turn_dist_m) {
        const double d_seg = d - (o->length_m + o->turn_dist_m);
        xl = o->length_m - d_seg;
        yl = o->side * 2.0 * o->turn_radius_m;
        hdg = M_PI;
        roll = 0.0f;
    }
    else {
        const double phi = (d - (2.0 * o->length_m + o->turn_dist_m)) / o->turn_radius_m + M_PI / 2.0;
        xl = o->turn_radius_m * cos(phi);
        yl = o->side * o->turn_radius_m * (1.0 + sin(phi));
        hdg = M_PI + (phi - M_PI / 2.0) * o->side;
        roll = o->roll_angle_r;
    }
    const double rad_o = o->orientation_r;
    const double cos_o = cos(rad_o), sin_o = sin(rad_o);
    st->x = (float)(xl * cos_o - yl * sin_o);
    st->y = (float)(xl * sin_o + yl * cos_o);
    st->dir = (float)fmod(hdg + o->orientation_r + 2.0 * M_PI, 2.0 * M_PI);
    st->rol = (float)roll;
    st->vel = (float)o->velocity_ms;
    st->alt = (float)o->altitude_m;
}

/* six platform values + rad_eff at time t: the piece the GPU path is checked against exactly */
void port_awacs_platform_state(double t, float *six, float *rad_eff)
{
    aw_orbit o;
    aw_platform st;
    orbit_init(&o);
    platform_at(&st, &o, t);
    six[0] = st.x; six[1] = st.y; six[2] = st.dir; six[3] = st.rol; six[4] = st.vel; six[5] = st.alt;
    *rad_eff = o.rad_eff;
}

/* ------------------------------------------------------------------ targets, :343-545 */
enum { HIDING = 0, STAGING = 1, FIRING = 2, DRIVING = 3 };
enum { UNDETERMINED = 0, BEYOND_HORIZON, NADIR_HOLE, TERRAIN_SHIELDED, MISSED, DETECTED };

typedef struct {
    float rcs[4], dwell[4], height;
    int   mode, tds;
    float rcs_now, time_s, x, y, alt, dir, vel;
    bool  detected;
} aw_target;

/* ------------------------------------------------------------------ event list */
enum { EV_TARGET = 0, EV_RADAR = 1, EV_END = 2, EV_BAR = 3 };
typedef struct { double t; uint64_t key; int32_t kind, who; } aw_event;
typedef struct { aw_event *e; uint64_t n, cap, issued; } aw_list;

static bool earlier(const aw_event *a, const aw_event *b)       /* src/cmi_hashheap.c:55-80 with every priority 0 */
{
    return a->t < b->t || (a->t == b->t && a->key < b->key);
}

static uint64_t list_add(aw_list *l, double t, int kind, int who)
{
    if (l->n == l->cap) {
        l->cap = l->cap ? 2u * l->cap : 2048u;
        l->e = realloc(l->e, l->cap * sizeof(aw_event));
    }
    const aw_event ev = { t, ++l->issued, kind, who };
    uint64_t i = l->n++;
    while (i > 0u && earlier(&ev, &l->e[(i - 1u) / 2u])) {
        l->e[i] = l->e[(i - 1u) / 2u];
        i = (i - 1u) / 2u;
    }
    l->e[i] = ev;
    return ev.key;
}

static bool list_take(aw_list *l, aw_event *out)
{
    if (l->n == 0u) return false;
    *out = l->e[0];
    const aw_event last = l->e[--l->n];
    uint64_t i = 0u;
    for (;;) {
        uint64_t c = 2u * i + 1u;
        if (c >= l->n) break;
        if (c + 1u < l->n && earlier(&l->e[c + 1u], &l->e[c])) c++;
        if (!earlier(&l->e[c], &last)) break;
        l->e[i] = l->e[c];
        i = c;
    }
    if (l->n > 0u) l->e[i] = last;
    return true;
}

/* ------------------------------------------------------------------ the detection chain, :548-689 */
static bool in_swept_sector(float prev_dir, float sweep_width, float tgt_azi)
{
    float rel = tgt_azi - prev_dir;
    while (rel < 0.0f) rel += 2.0f * (float)M_PI;
    while (rel >= 2.0f * (float)M_PI) rel -= 2.0f * (float)M_PI;
    return rel <= sweep_width;
}

static bool beyond_horizon(float d_2d, float h_sensor, float h_tgt, float r_eff)
{
    const float hs = fmaxf(0.0f, h_sensor);
    const float ht = fmaxf(0.0f, h_tgt);
    const float reach = sqrtf(2.0f * r_eff * hs) + sqrtf(2.0f * r_eff * ht);
    return d_2d > reach;
}

static bool outside_vertical(float dx, float dy, float dz, float d_2d, float hdg, float roll, float lo, float hi)
{
    const float azi = atan2f(dy, dx);
    const float rel_brg = azi - hdg;
    const float geom_elev = atan2f(dz, d_2d);
    const float apparent = geom_elev - (roll * sinf(rel_brg));
    return (apparent < lo) || (apparent > hi);
}

static bool terrain_shielded(float sx, float sy, float sa, float tx, float ty, float ta, const aw_terrain *t)
{
    const float dx = tx - sx, dy = ty - sy, dz = ta - sa;
    const float d_2d = sqrtf(dx * dx + dy * dy);
    const float step = fminf(t->x_scale, t->y_scale) * 0.5f;
    const int steps = (int)(d_2d / step);
    if (steps < 1) return false;
    const float inv = 1.0f / (float)steps;
    for (int i = 1; i < steps; i++) {
        const float f = (float)i * inv;
        float cx = sx + dx * f;
        float cy = sy + dy * f;
        const float ca = sa + dz * f;
        cx = fmaxf(t->x_min, fminf(cx, t->x_max));
        cy = fmaxf(t->y_min, fminf(cy, t->y_max));
        if (ca < elevation(t, cx, cy)) return true;
    }
    return false;
}

static float detection_probability(float sa, float ref_range, float ref_rcs, float ta, float tcx, float d_3d)
{
    const float r = fmaxf(1.0f, d_3d);
    const float snr = powf(ref_range / r, 4.0f) * (tcx / ref_rcs);
    float bv = k_biome_vis[2];
    for (unsigned b = 0; b < 2; b++) {
        if (ta < k_biome_elev[b]) {
            bv = k_biome_vis[b];
            break;
        }
    }
    const float dz = sa - ta;
    float sin_grazing = 0.0f;
    if (dz > 0.0f) sin_grazing = fminf(1.0f, dz / r);
    const float clutter = 1.0f - (sin_grazing * 0.8f);
    const float sinr = snr * bv * clutter;
    return 1.0f / (1.0f + expf(-0.5f * (sinr - 10.0f)));
}

/* ------------------------------------------------------------------ one trial, run_trial :1139-1255 */
typedef struct {
    uint64_t events;
    double   t_end;
    uint32_t num_found;
    uint32_t tds_count[6];
    uint32_t mode_count[4];
    uint32_t pad;
    double   sum_x, sum_y;
} port_awacs_out;

static void target_hold(aw_list *fel, port_rng *rng, aw_target *g, int who, double now)
{
    /* the part of target_proc's loop from its top to the next cmb_process_hold, :404-444 */
    if (g->mode == HIDING) {
        g->rcs_now = g->rcs[HIDING];
        g->time_s = (float)now;
        g->vel = 0.0f;
        list_add(fel, now + port_exponential(rng, g->dwell[HIDING]), EV_TARGET, who);
    }
    else {
        g->rcs_now = g->rcs[DRIVING];
        g->dir = (float)port_uniform(rng, 0.0, 2.0 * M_PI);
        g->vel = (float)port_uniform(rng, 5.0, 20.0);
        const double t_m = g->dwell[DRIVING] / (float)5u;
        g->time_s = (float)now;
        list_add(fel, now + port_erlang(rng, 5u, t_m), EV_TARGET, who);
    }
}

int port_awacs_trial(uint64_t seed, double duration_h, const float *map, uint32_t cols, uint32_t rows,
                     const float *geom, uint64_t trace_cap, uint64_t *trace_key, double *trace_time,
                     port_awacs_out *out, float *xs, float *ys, int *modes, int *tdss, int *dets)
{
    const aw_terrain ter = { geom[0], geom[1], geom[2], geom[3], geom[4], geom[5], cols, rows, map };
    port_rng rng;
    port_rng_init(&rng, seed);
    aw_list fel = { NULL, 0u, 0u, 0u };
    aw_target *tg = calloc(AW_TARGETS, sizeof(aw_target));
    unsigned char *phase = calloc(AW_TARGETS, 1);       /* 0 = not started; else the hold the target sleeps in */
    double now = 0.0;

    for (int i = 0; i < AW_TARGETS; i++) {              /* target_initialize, :455-490 with :1153-1166 */
        aw_target *g = &tg[i];
        g->height = 2.0f;
        g->rcs[0] = 5.0f; g->rcs[1] = 100.0f; g->rcs[2] = 1000.0f; g->rcs[3] = 50.0f;
        g->dwell[0] = 3.0f * 3600.0f;
        g->dwell[1] = 5.0f * 60.0f;
        g->dwell[2] = 30.0f;
        g->dwell[3] = 1.0f * 3600.0f;
        g->tds = UNDETERMINED;
        list_add(&fel, now, EV_TARGET, i);              /* cmb_process_start */
    }
    aw_orbit orbit;
    orbit_init(&orbit);
    aw_platform host = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
    /* sensor_initialize, :1045-1065 with :1200-1206 */
    const float rpm = 6.0f;
    const float max_elev = (float)(60.0 * k_deg_rad);
    const float min_elev = (float)(-20.0 * k_deg_rad);
    const float ref_range_m = 150.0f * (float)k_nm_m;
    const float ref_rcs = 1.0f;
    float cur_dir = (float)(M_PI / 2.0);
    const float rot_inc = (float)(rpm * (k_time_step / 60.0f) * (2.0f * M_PI));
    bool radar_started = false;
    list_add(&fel, now, EV_RADAR, 0);                   /* cmb_process_start(radar) */
    double t_end_s = duration_h * 3600.0;
    list_add(&fel, t_end_s, EV_END, 0);                 /* cmb_event_schedule(end_sim, ...) */
    list_add(&fel, now, EV_BAR, 0);                     /* the progress-bar process */
    unsigned bar_cycles = 0u;
    double bar_incr = 0.0;

    uint64_t pops = 0u;
    aw_event ev;
    while (list_take(&fel, &ev)) {
        now = ev.t;
        if (pops < trace_cap) {
            trace_key[pops] = ev.key;
            trace_time[pops] = now;
        }
        pops++;
        if (ev.kind == EV_TARGET) {
            aw_target *g = &tg[ev.who];
            if (phase[ev.who] == 0u) {                  /* first entry of target_proc, :393-402 */
                g->time_s = (float)now;
                g->x = (float)port_uniform(&rng, ter.x_min, ter.x_max);
                g->y = (float)port_uniform(&rng, ter.y_min, ter.y_max);
                g->alt = elevation(&ter, g->x, g->y) + g->height;
                const double ph = g->dwell[HIDING] / (g->dwell[HIDING] + g->dwell[DRIVING]);
                g->mode = port_bernoulli(&rng, ph) ? HIDING : DRIVING;
                g->tds = UNDETERMINED;
                phase[ev.who] = 1u;
                target_hold(&fel, &rng, g, ev.who, now);
            }
            else if (g->mode == HIDING) {               /* unmask, :414-420 */
                g->mode = STAGING;
                g->rcs_now = g->rcs[STAGING];
                g->time_s = (float)now;
                const double t_m = g->dwell[STAGING] / (float)10u;
                list_add(&fel, now + port_erlang(&rng, 10u, t_m), EV_TARGET, ev.who);
            }
            else if (g->mode == STAGING) {              /* shoot, :422-429 */
                g->mode = FIRING;
                g->rcs_now = g->rcs[FIRING];
                g->time_s = (float)now;
                const double t_m = g->dwell[FIRING] / (float)20u;
                list_add(&fel, now + port_erlang(&rng, 20u, t_m), EV_TARGET, ev.who);
            }
            else if (g->mode == FIRING) {               /* scoot: :430, then the driving branch */
                g->mode = DRIVING;
                target_hold(&fel, &rng, g, ev.who, now);
            }
            else {                                      /* done driving: :443, then the hiding branch */
                g->mode = HIDING;
                target_hold(&fel, &rng, g, ev.who, now);
            }
        }
        else if (ev.kind == EV_RADAR) {
            if (!radar_started) {                       /* :954-957 */
                radar_started = true;
                platform_at(&host, &orbit, now);
                list_add(&fel, now + (double)k_time_step, EV_RADAR, 0);
                continue;
            }
            const float prev_hdg = host.dir;            /* :962-979 */
            const float prev_sensor_dir = cur_dir;
            platform_at(&host, &orbit, now);
            const float ddir = host.dir - prev_hdg;
            float sweep_width = rot_inc + ddir;
            cur_dir += sweep_width;
            while (cur_dir >= 2.0f * (float)M_PI) cur_dir -= 2.0f * (float)M_PI;
            while (cur_dir < 0.0f) cur_dir += 2.0f * (float)M_PI;
            if (sweep_width < 0.0f) sweep_width = 0.01f;

            for (int i = 0; i < AW_TARGETS; i++) {      /* :982-1031 */
                aw_target *g = &tg[i];
                if (g->vel > 0.0f) {                    /* target_position_update, :507-545 */
                    const double dt = now - g->time_s;
                    float x = g->x + (float)(dt * g->vel * cosf(g->dir));
                    if (x > ter.x_max) x = ter.x_min + (x - ter.x_max);
                    else if (x < ter.x_min) x = ter.x_max - (ter.x_min - x);
                    float y = g->y + (float)(dt * g->vel * sinf(g->dir));
                    if (y > ter.y_max) y = ter.y_min + (y - ter.y_max);
                    else if (y < ter.y_min) y = ter.y_max - (ter.y_min - y);
                    const float alt = elevation(&ter, x, y) + g->height;
                    g->time_s = (float)now;
                    g->x = x;
                    g->y = y;
                    g->alt = alt;
                }
                const float sx = host.x, sy = host.y, sa = host.alt;
                const float tx = g->x, ty = g->y, ta = g->alt;
                const float dx = tx - sx, dy = ty - sy, dz = ta - sa;
                const float d_2d = sqrtf(dx * dx + dy * dy);
                const float d_3d = sqrtf(d_2d * d_2d + dz * dz);
                const float azi = atan2f(dy, dx);
                if (!in_swept_sector(prev_sensor_dir, sweep_width, azi)) continue;
                if (beyond_horizon(d_2d, sa, ta, orbit.rad_eff)) { g->tds = BEYOND_HORIZON; continue; }
                if (outside_vertical(dx, dy, dz, d_2d, host.dir, host.rol, min_elev, max_elev)) { g->tds = NADIR_HOLE; continue; }
                if (terrain_shielded(sx, sy, sa, tx, ty, ta, &ter)) { g->tds = TERRAIN_SHIELDED; continue; }
                const float pd = detection_probability(sa, ref_range_m, ref_rcs, ta, g->rcs_now, d_3d);
                if (port_bernoulli(&rng, pd)) {
                    g->tds = DETECTED;
                    if (g->mode != FIRING) g->detected = true;
                }
                else {
                    g->tds = MISSED;
                }
            }
            list_add(&fel, now + (double)k_time_step, EV_RADAR, 0);
        }
        else if (ev.kind == EV_END) {                   /* end_sim, :1100-1112: every target and the radar sleep in a hold */
            uint64_t kept = 0u;
            aw_event *rest = malloc((fel.n + 1u) * sizeof(aw_event));
            for (uint64_t i = 0; i < fel.n; i++) {
                if (fel.e[i].kind == EV_BAR) rest[kept++] = fel.e[i];
            }
            fel.n = 0u;
            const uint64_t issued = fel.issued;
            for (uint64_t i = 0; i < kept; i++) {       /* re-insert keeping the keys */
                fel.issued = rest[i].key - 1u;
                list_add(&fel, rest[i].t, rest[i].kind, rest[i].who);
            }
            fel.issued = issued;
            free(rest);
        }
        else {                                          /* ent_proc, :1118-1137 */
            if (bar_cycles == 0u) {
                bar_incr = (t_end_s - now) / 100u;
            }
            if (bar_cycles < 100u) {
                bar_cycles++;
                list_add(&fel, now + bar_incr, EV_BAR, 0);
            }
        }
    }

    memset(out, 0, sizeof(*out));
    out->events = pops;
    out->t_end = now;
    for (int i = 0; i < AW_TARGETS; i++) {
        if (tg[i].detected) out->num_found++;
        out->tds_count[tg[i].tds]++;
        out->mode_count[tg[i].mode]++;
        out->sum_x += tg[i].x;
        out->sum_y += tg[i].y;
        if (xs) xs[i] = tg[i].x;
        if (ys) ys[i] = tg[i].y;
        if (modes) modes[i] = tg[i].mode;
        if (tdss) tdss[i] = tg[i].tds;
        if (dets) dets[i] = tg[i].detected ? 1 : 0;
    }
    free(fel.e);
    free(tg);
    free(phase);
    return 0;
}
