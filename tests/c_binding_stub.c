/* tests/c_binding_stub.c - what INTEGRATION.md shows a maintainer of the reference: plain C including
 * include/cimba_b200.h and linking libcimba_b200.so.  tests/test_capi_abi.py compiles it with gcc (C11, -Wall -Wextra
 * -Werror) and runs it: without a GPU every compute call must come back with CIMBA_B200_ENODEVICE and a message. */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cimba_b200.h"

struct trial { double arr_mean; double srv_mean; uint64_t obj_cnt; double sum_wait; double avg_wait; };   /* benchmark/MM1_multi.c:39-45 */

static void *on_thread_start(void *usrarg, uint64_t tid) { (void)tid; return usrarg; }
static void on_thread_end(void *ctx) { (void)ctx; }

int main(void)
{
    enum { TRIALS = 8 };
    struct trial *experiment = calloc(TRIALS, sizeof(*experiment));
    if (experiment == NULL) return 2;
    for (unsigned i = 0; i < TRIALS; i++) {
        experiment[i].arr_mean = 1.0 / 0.9;
        experiment[i].srv_mean = 1.0;
    }
    const cimba_b200_experiment desc = {
        .model = CIMBA_B200_MODEL_MM1, .device = -1, .master_seed = 0x34f05c64d7ad598fULL, .first_trial = 0,
        .num_objects = 1000,
        .off_arr_mean = offsetof(struct trial, arr_mean), .off_srv_mean = offsetof(struct trial, srv_mean),
        .off_obj_cnt = offsetof(struct trial, obj_cnt), .off_sum_wait = offsetof(struct trial, sum_wait),
        .off_avg_wait = offsetof(struct trial, avg_wait),
        .off_events = CIMBA_B200_NO_FIELD, .off_t_end = CIMBA_B200_NO_FIELD, .off_status = CIMBA_B200_NO_FIELD,
        .off_max_queue = CIMBA_B200_NO_FIELD, .off_counters = CIMBA_B200_NO_FIELD,
    };
    cimba_b200_set_thread_hooks(on_thread_start, experiment, on_thread_end);
    const int devices = cimba_b200_device_count();
    const int rc = cimba_b200_run_experiment(experiment, TRIALS, sizeof(*experiment), &desc);
    const int rc_all = cimba_b200_run_experiment_all_gpus(experiment, TRIALS, sizeof(*experiment), &desc, 0);

    cimba_b200_datasummary s;
    cimba_b200_datasummary_initialize(&s);
    for (unsigned i = 0; i < 5; i++) cimba_b200_datasummary_add(&s, (double)i);
    printf("version %s devices %d rc %d rc_all %d msg \"%s\" fmix %llx mean %.1f ctx %s\n", cimba_b200_version(), devices, rc, rc_all,
           cimba_b200_last_error(), (unsigned long long)cimba_b200_fmix64(0x34f05c64d7ad598fULL, 0), cimba_b200_datasummary_mean(&s),
           cimba_b200_thread_context() == NULL ? "null" : "set");
    cimba_b200_datasummary_print(&s, stdout, 1);
    if (devices > 0) {
        for (unsigned i = 0; i < TRIALS; i++) if (experiment[i].obj_cnt != 1000) return 3;
        return rc == CIMBA_B200_OK && rc_all == CIMBA_B200_OK ? 0 : 4;
    }
    free(experiment);
    return rc == CIMBA_B200_ENODEVICE && rc_all == CIMBA_B200_ENODEVICE && strlen(cimba_b200_last_error()) > 0 ? 0 : 5;
}
