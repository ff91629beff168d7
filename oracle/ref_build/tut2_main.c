/* oracle/ref_build/tut2_main.c - TEST INFRASTRUCTURE ONLY.
 *
 * The reference's second tutorial (tutorial/tut_2_1.c: mice acquiring, rats pre-empting, a cat interrupting - a cmb_resourcepool
 * under every kind of process interaction) compiled UNMODIFIED from where it lies, as a PROGRAM: the pool's pre-emption order
 * breaks priority ties by process ADDRESS (src/cmb_resourcepool.c:82-89) and cmb_random_flip keeps cached bits between trials, so
 * a trial is only a function of its seed in a fresh process - one run of this program per trial.
 *   usage: tut2_ref <seed>      prints "<events executed> <final clock as a hex float> <the next raw output of the random stream>"
 * Redirected names: cmb_random_hwseed -> the seed on the command line, cmb_event_queue_execute -> a counting loop, printf ->
 * nothing, main -> tut2_reference_main (never called).
 */
#include <cimba.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

static uint64_t g_seed, g_pops, g_next;
static double g_end;

static uint64_t tut2_seed_hook(void)
{
    return g_seed;
}

static void tut2_counting_execute(void)
{
    uint64_t n = 0u;
    while (cmb_event_execute_next()) {
        n++;
    }
    g_pops = n;
    g_end = cmb_time();
    g_next = cmb_random_sfc64();        /* where the random stream stands after the run: a fingerprint of every draw made */
}

static int tut2_quiet_printf(const char *fmt, ...)
{
    (void)fmt;
    return 0;
}

#define cmb_random_hwseed tut2_seed_hook
#define cmb_event_queue_execute tut2_counting_execute
#define printf tut2_quiet_printf
#define main tut2_reference_main

#include "tutorial/tut_2_1.c"

#undef cmb_random_hwseed
#undef cmb_event_queue_execute
#undef printf
#undef main

int main(int argc, char **argv)
{
    if (argc < 2) {
        return 2;
    }
    g_seed = strtoull(argv[1], NULL, 0);
    cmb_logger_flags_off(USERFLAG1);
    run_trial(NULL);
    printf("%llu %a %llu\n", (unsigned long long)g_pops, g_end, (unsigned long long)g_next);
    return 0;
}
