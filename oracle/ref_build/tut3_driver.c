/* oracle/ref_build/tut3_driver.c - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's third tutorial (tutorial/tut_3_1.c: a theme park - M/G/n attractions with batch rides, priority queues, and
 * visitors that balk, jockey and renege on patience timers) compiled UNMODIFIED from where it lies - the #include below is the
 * whole of it - with four names redirected to the hooks defined here:
 *   cmb_random_hwseed       -> tut3_seed_hook          (the trial is seeded by the caller instead of the hardware)
 *   cmb_event_queue_execute -> tut3_counting_execute   (the same loop, counting the pops and noting the final clock)
 *   printf / stdout         -> nothing / a stream on /dev/null (the tutorial prints its reports)
 *   main                    -> tut3_reference_main (never called).
 * Entry point: tut3_ref_trial() calls the tutorial's own run_trial().
 */
#include <cimba.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdint.h>

static _Thread_local uint64_t t_seed;
static _Thread_local uint64_t t_pops;
static _Thread_local double t_end;
static FILE *g_sink;

static uint64_t tut3_seed_hook(void)
{
    return t_seed;
}

static void tut3_counting_execute(void)
{
    uint64_t n = 0u;
    while (cmb_event_execute_next()) {
        n++;
    }
    t_pops = n;
    t_end = cmb_time();
}

static int tut3_quiet_printf(const char *fmt, ...)
{
    (void)fmt;
    return 0;
}

static FILE *tut3_sink(void)
{
    if (g_sink == NULL) {
        g_sink = fopen("/dev/null", "w");
    }
    return g_sink;
}

#define cmb_random_hwseed tut3_seed_hook
#define cmb_event_queue_execute tut3_counting_execute
#define printf tut3_quiet_printf
#undef stdout
#define stdout (tut3_sink())
#define main tut3_reference_main

#include "tutorial/tut_3_1.c"

#undef cmb_random_hwseed
#undef cmb_event_queue_execute
#undef printf
#undef main

struct tut3_out {
    uint64_t events;
    double   t_end;
    double   avg_time_in_park, avg_time_riding, avg_time_waiting, avg_time_walking, avg_num_rides;
};

/* one trial of the tutorial with this seed */
int tut3_ref_trial(uint64_t seed, struct tut3_out *out)
{
    struct trial trl = { 0 };
    load_params(&trl);
    t_seed = seed;
    t_pops = 0u;
    t_end = 0.0;
    cmb_logger_flags_off(LOGFLAG_ALL & 0x00ffffffu);   /* the user-level chatter; it would only go to the sink */
    run_trial(&trl);
    out->events = t_pops;
    out->t_end = t_end;
    out->avg_time_in_park = trl.avg_time_in_park;
    out->avg_time_riding = trl.avg_time_riding;
    out->avg_time_waiting = trl.avg_time_waiting;
    out->avg_time_walking = trl.avg_time_walking;
    out->avg_num_rides = trl.avg_num_rides;
    return trl.seed_used == seed ? 0 : -1;
}
