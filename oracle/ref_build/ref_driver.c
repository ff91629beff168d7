/*
 * ref_driver.c - TEST INFRASTRUCTURE ONLY.
 *
 * Seeded model drivers written against the UNMODIFIED reference library
 * (ambonvik/cimba, compiled from /root/reference by oracle/Makefile into
 * oracle/_ref/).  They exist so that (a) the C restatement in oracle/port and
 * the CUDA engine in cimba_b200/csrc can be compared trial-for-trial with the
 * real reference, and (b) bench.py can time the reference's own pthread
 * executive (cimba_run_experiment, src/cimba.c:151) on the GPU box's host
 * cores.  Nothing here is linked into the product library.
 *
 * Workloads:
 *   model 0  M/M/1   = benchmark/MM1_multi.c:52-125, but seeded with
 *                      cmb_random_fmix64(master, trial index) exactly like
 *                      test/test_cimba.c:396 instead of the hardware seed.
 *   model 1  G/G/1   = same structure; inter-arrival cmb_random_erlang(2, m/2),
 *                      service = normal(mean, mean/4) redrawn while negative.
 *   model 2  M/M/c   = generator process starting one customer process per
 *                      arrival (recycled process structs), customers contend
 *                      for a cmb_resourcepool of capacity c.
 *
 * An "event" is one successful cmb_event_execute_next() (src/cmb_event.c:229).
 */
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cimba.h>
#include <cmb_priorityqueue.h>
#include <cmb_condition.h>

#include "cmi_mempool.h"
#include "cmi_hashheap.h"
#include <math.h>

struct ref_trial {
    /* in */
    uint64_t seed;
    uint64_t num_objects;
    double arr_mean;
    double srv_mean;
    int32_t model;
    int32_t servers;
    int32_t stock;              /* 1 = timed as the reference's benchmark runs it: no per-event bookkeeping (bench.py) */
    int32_t pad_;
    /* out */
    uint64_t events;
    uint64_t objects;
    double t_end;
    double sum_wait;
    uint64_t max_fel;
    uint64_t max_queue;
    uint64_t counter[8];        /* model 3: see struct ref_counters below */
    /* optional pop trace (single-trial calls only) */
    uint64_t trace_cap;
    uint64_t *trace_key;
    double *trace_time;
};

/* ---------------------------------------------------------------- queues */

static CMB_THREAD_LOCAL struct cmi_mempool stamp_pool = CMI_MEMPOOL_STATIC_INIT(8u, 512u);

struct q_world {
    struct ref_trial *trl;
    struct cmb_objectqueue *queue;
    struct cmb_process *source;
    struct cmb_process *server;
};

static double draw_interarrival(const struct ref_trial *t)
{
    if (t->model == 1) {
        return cmb_random_erlang(2u, 0.5 * t->arr_mean);
    }
    return cmb_random_exponential(t->arr_mean);
}

static double draw_service(const struct ref_trial *t)
{
    if (t->model == 1) {
        double s;
        do {
            s = cmb_random_normal(t->srv_mean, 0.25 * t->srv_mean);
        } while (s < 0.0);
        return s;
    }
    return cmb_random_exponential(t->srv_mean);
}

static void *q_source_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct q_world *w = vw;
    for (uint64_t i = 0u; i < w->trl->num_objects; i++) {
        cmb_process_hold(draw_interarrival(w->trl));
        double *stamp = cmi_mempool_alloc(&stamp_pool);
        *stamp = cmb_time();
        cmb_objectqueue_put(w->queue, stamp);
        const uint64_t len = cmb_objectqueue_length(w->queue);
        if (len > w->trl->max_queue) {
            w->trl->max_queue = len;
        }
    }
    return NULL;
}

static void *q_server_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct q_world *w = vw;
    for (;;) {
        void *obj = NULL;
        cmb_objectqueue_get(w->queue, &obj);
        cmb_process_hold(draw_service(w->trl));
        w->trl->sum_wait += cmb_time() - *(double *)obj;
        w->trl->objects += 1u;
        cmi_mempool_free(&stamp_pool, obj);
    }
}

/* The arrival process exactly as benchmark/MM1_multi.c:52-68 has it - no queue-length bookkeeping per put.
 * bench.py times THIS body (stock = 1); the instrumented one above serves the parity tests. */
static void *q_source_body_stock(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct q_world *w = vw;
    for (uint64_t i = 0u; i < w->trl->num_objects; i++) {
        cmb_process_hold(draw_interarrival(w->trl));
        double *stamp = cmi_mempool_alloc(&stamp_pool);
        *stamp = cmb_time();
        cmb_objectqueue_put(w->queue, stamp);
    }
    return NULL;
}

/* ----------------------------------------------------------------- M/M/c */

struct c_world;

struct c_customer {
    struct cmb_process proc;        /* parent "class" first, as the reference's tutorials do */
    struct c_world *world;
    double t_arrival;
    struct c_customer *next_free;
};

struct c_world {
    struct ref_trial *trl;
    struct cmb_resourcepool *pool;
    struct cmb_process *source;
    struct c_customer *free_list;
    struct c_customer *all_list[4096];
    unsigned all_count;
};

static void *c_customer_body(struct cmb_process *me, void *vw)
{
    struct c_customer *cu = (struct c_customer *)me;
    struct c_world *w = vw;
    cmb_resourcepool_acquire(w->pool, 1u);
    cmb_process_hold(cmb_random_exponential(w->trl->srv_mean));
    cmb_resourcepool_release(w->pool, 1u);
    w->trl->sum_wait += cmb_time() - cu->t_arrival;
    w->trl->objects += 1u;
    cu->next_free = w->free_list;
    w->free_list = cu;
    return NULL;
}

static void *c_source_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct c_world *w = vw;
    for (uint64_t i = 0u; i < w->trl->num_objects; i++) {
        cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        struct c_customer *cu = w->free_list;
        if (cu != NULL) {
            w->free_list = cu->next_free;
        }
        else {
            cu = calloc(1, sizeof(*cu));
            cmb_process_initialize(&cu->proc, "Customer", c_customer_body, w, 0);
            cu->world = w;
            if (w->all_count >= 4096u) {
                fprintf(stderr, "ref_driver: more than 4096 live customers\n");
                abort();
            }
            w->all_list[w->all_count++] = cu;
        }
        cu->t_arrival = cmb_time();
        cmb_process_start(&cu->proc);
    }
    return NULL;
}


/* ------------------------------------------------- model 3: guarded queue under fire
 *
 * The reference's own object-queue torture test (test/test_objectqueue.c:40-170)
 * restated with counters instead of log lines: three putters and three getters
 * with random priorities on a BOUNDED cmb_objectqueue (both guards in play), a
 * nuisance process that interrupts a random victim with a random signal and a
 * random event priority, and an end event that stops everybody.  It exercises
 * cmb_process_interrupt / cmi_process_cancel_awaiteds / cmb_event_cancel /
 * cmb_event_pattern_cancel / cmi_hashheap_remove on the guard, process and event
 * priorities in the event list and in the guard order, and cmb_process_stop.
 *
 *   num_objects = duration (time units), servers = queue capacity,
 *   arr_mean = putter hold mean, srv_mean = getter hold mean, nuisance hold mean 1.
 * counters: [0] successful puts [1] successful gets [2] interrupted holds
 *           [3] interrupted puts [4] interrupted gets [5] sum of signals received
 *           [6] final queue length [7] interrupts issued
 * sum_wait = sum over successful gets of (get time - put time of that object)
 */
static void pump_events(struct ref_trial *t);
#define G_PUTTERS 3u
#define G_GETTERS 3u

struct g_world {
    struct ref_trial *trl;
    struct cmb_objectqueue *queue;
    struct cmb_priorityqueue *pq;       /* model 13 (test/test_priorityqueue.c): the same workload on a priority queue */
    struct cmb_process *worker[G_PUTTERS + G_GETTERS];
    struct cmb_process *nuisance;
};

static void note_signal(struct ref_trial *t, int64_t sig, unsigned which)
{
    if (sig != CMB_PROCESS_SUCCESS) {
        t->counter[which] += 1u;
        t->counter[5] += (uint64_t)sig;
    }
}

static void *g_putter_body(struct cmb_process *me, void *vw)
{
    struct g_world *w = vw;
    for (;;) {
        int64_t sig = cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        note_signal(w->trl, sig, 2u);
        double *stamp = cmi_mempool_alloc(&stamp_pool);
        *stamp = cmb_time();
        sig = (w->pq != NULL) ? cmb_priorityqueue_put(w->pq, stamp, cmb_process_priority(me), NULL)
                              : cmb_objectqueue_put(w->queue, stamp);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[0] += 1u;
        }
        else {
            note_signal(w->trl, sig, 3u);
            cmi_mempool_free(&stamp_pool, stamp);
        }
    }
}

static void *g_getter_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct g_world *w = vw;
    for (;;) {
        int64_t sig = cmb_process_hold(cmb_random_exponential(w->trl->srv_mean));
        note_signal(w->trl, sig, 2u);
        void *obj = NULL;
        sig = (w->pq != NULL) ? cmb_priorityqueue_get(w->pq, &obj) : cmb_objectqueue_get(w->queue, &obj);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[1] += 1u;
            w->trl->sum_wait += cmb_time() - *(double *)obj;
            cmi_mempool_free(&stamp_pool, obj);
        }
        else {
            note_signal(w->trl, sig, 4u);
        }
    }
}

static void *g_nuisance_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct g_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)(G_PUTTERS + G_GETTERS - 1u));
        const int64_t sig = cmb_random_dice(1, 10);
        const int64_t pri = cmb_random_dice(-5, 5);
        w->trl->counter[7] += 1u;
        cmb_process_interrupt(w->worker[victim], sig, pri);
    }
}

static void g_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct g_world *w = subject;
    for (unsigned i = 0u; i < G_PUTTERS + G_GETTERS; i++) {
        cmb_process_stop(w->worker[i], NULL);
    }
    cmb_process_stop(w->nuisance, NULL);
}

static void run_guarded_trial(struct ref_trial *t)
{
    struct g_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->queue = cmb_objectqueue_create();
    cmb_objectqueue_initialize(w->queue, "Queue", (uint64_t)t->servers);
    if (t->model == 11) {
        /* model 11 = model 3 with the queue's history on, exactly test/test_objectqueue.c:187-191: with
         * capacity 10, both means 1, duration 1e6 and seed 0x34f05c64d7ad598f the time-weighted queue length
         * is test/reference/objectqueue.txt's "N 5689021 Mean 5.008" */
        cmb_objectqueue_recording_start(w->queue);
    }
    if (t->model == 13) {
        /* model 13 = test/test_priorityqueue.c: the same seven processes on a cmb_priorityqueue, objects put
         * with the putter's own priority, history on -> test/reference/priorityqueue.txt (the same "N 5689021
         * Mean 5.008": which object a get returns does not change the event sequence; sum_wait does differ).
         * The test itself passes the address of a local pointer as the object; here it is the stamp. */
        w->pq = cmb_priorityqueue_create();
        cmb_priorityqueue_initialize(w->pq, "Queue", (uint64_t)t->servers);
        cmb_priorityqueue_recording_start(w->pq);
    }
    for (unsigned i = 0u; i < G_PUTTERS + G_GETTERS; i++) {
        w->worker[i] = cmb_process_create();
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_initialize(w->worker[i], (i < G_PUTTERS) ? "Putter" : "Getter",
                               (i < G_PUTTERS) ? g_putter_body : g_getter_body, w, pri);
        cmb_process_start(w->worker[i]);
    }
    w->nuisance = cmb_process_create();
    cmb_process_initialize(w->nuisance, "Nuisance", g_nuisance_body, w, 0);
    cmb_process_start(w->nuisance);
    (void)cmb_event_schedule(g_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->counter[6] = (w->pq != NULL) ? cmb_priorityqueue_length(w->pq) : cmb_objectqueue_length(w->queue);
    t->objects = t->counter[1];
    if (t->model == 13) {
        cmb_priorityqueue_recording_stop(w->pq);
        struct cmb_wtdsummary ws;
        cmb_wtdsummary_initialize(&ws);
        (void)cmb_timeseries_summarize(cmb_priorityqueue_history(w->pq), &ws);
        const double mean = cmb_wtdsummary_mean(&ws);
        memcpy(&t->counter[6], &mean, 8);
        t->max_queue = cmb_wtdsummary_count(&ws);
        cmb_priorityqueue_destroy(w->pq);
    }
    if (t->model == 11) {
        /* counter[6] = mean queue length (bits), max_queue = samples with a duration */
        cmb_objectqueue_recording_stop(w->queue);
        struct cmb_wtdsummary ws;
        cmb_wtdsummary_initialize(&ws);
        (void)cmb_timeseries_summarize(cmb_objectqueue_history(w->queue), &ws);
        const double mean = cmb_wtdsummary_mean(&ws);
        memcpy(&t->counter[6], &mean, 8);
        t->max_queue = cmb_wtdsummary_count(&ws);
    }
    for (unsigned i = 0u; i < G_PUTTERS + G_GETTERS; i++) {
        cmb_process_terminate(w->worker[i]);
        cmb_process_destroy(w->worker[i]);
    }
    cmb_process_terminate(w->nuisance);
    cmb_process_destroy(w->nuisance);
    cmb_objectqueue_destroy(w->queue);
    free(w);
}


/* ------------------------------------------------- model 4: resource pool with pre-emption
 *
 * The reference's pool torture test (test/test_resourcepool.c:72-330) restated with
 * counters: three "mice" that change their own priority and acquire politely, two
 * "rats" that pre-empt (cmb_resourcepool_preempt), one "cat" that interrupts a random
 * rodent, an end event that stops everybody (dropping what they hold).  Exercises
 * cmi_pool_acquire_inner in full (partial grabs, the holders heap, pre-emption,
 * roll-back on interrupt), cmb_resourcepool_release (partial), cmb_process_priority_set
 * -> reprioritize_holder -> cmi_hashheap_reprioritize, resourcepool_drop_holder.
 *
 * All processes live in ONE array so that their addresses ascend with their index:
 * the holders heap breaks priority ties by larger address first
 * (src/cmb_resourcepool.c:82-89), and the device uses the index.
 *
 *   num_objects = duration, servers = pool capacity.
 * counters: [0] acquires ok [1] pre-empts ok [2] PREEMPTED received [3] other signals
 *           received [4] sum of signals (two's complement) [5] units released
 *           [6] final in_use [7] own-vs-library holding mismatches (must stay 0)
 * sum_wait = sum over releases of cmb_time() * units released
 */
#define P_MICE 3u
#define P_RATS 2u
#define P_RODENTS (P_MICE + P_RATS)

struct p_world {
    struct ref_trial *trl;
    struct cmb_resourcepool *pool;
    struct cmb_process *proc;           /* P_RODENTS + 1 contiguous process structs */
};

static void p_check(struct p_world *w, struct cmb_process *me, uint64_t held)
{
    if (cmb_resourcepool_held_by_process(w->pool, me) != held) {
        w->trl->counter[7] += 1u;
    }
}

static void p_signal(struct p_world *w, int64_t sig, uint64_t *held)
{
    if (sig == CMB_PROCESS_PREEMPTED) {
        w->trl->counter[2] += 1u;
        *held = 0u;
    }
    else if (sig != CMB_PROCESS_SUCCESS) {
        w->trl->counter[3] += 1u;
    }
    w->trl->counter[4] += (uint64_t)sig;
}

static void *p_rodent_body(struct cmb_process *me, void *vw)
{
    struct p_world *w = vw;
    const bool rat = (me >= &w->proc[P_MICE]);
    uint64_t held = 0u;
    for (;;) {
        p_check(w, me, held);
        const uint64_t req = (uint64_t)cmb_random_dice(1, 5);
        int64_t sig;
        if (rat) {
            sig = cmb_resourcepool_preempt(w->pool, req);
        }
        else {
            cmb_process_priority_set(me, cmb_random_dice(-5, 5));
            sig = cmb_resourcepool_acquire(w->pool, req);
        }
        if (sig == CMB_PROCESS_SUCCESS) {
            held += req;
            w->trl->counter[rat ? 1 : 0] += 1u;
            p_check(w, me, held);
            sig = cmb_process_hold(cmb_random_exponential(1.0));
            if (sig == CMB_PROCESS_SUCCESS) {
                uint64_t rel = (uint64_t)cmb_random_dice(1, 5);
                if (rel > held || cmb_random_dice(0, 1) == 1) {
                    rel = held;                         /* every other time: give everything back */
                }
                cmb_resourcepool_release(w->pool, rel);
                held -= rel;
                w->trl->counter[5] += rel;
                w->trl->sum_wait += cmb_time() * (double)rel;
            }
            else {
                p_signal(w, sig, &held);
            }
        }
        else {
            p_signal(w, sig, &held);
        }
        p_check(w, me, held);
        sig = cmb_process_hold(cmb_random_exponential(1.0));
        if (sig != CMB_PROCESS_SUCCESS) {
            p_signal(w, sig, &held);
        }
    }
}

static void *p_cat_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct p_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)P_RODENTS - 1);
        const int64_t loud = cmb_random_dice(10, 100);
        const int64_t sig = (cmb_random_dice(0, 1) == 1) ? CMB_PROCESS_INTERRUPTED : loud;
        cmb_process_interrupt(&w->proc[victim], sig, 0);
    }
}

static void p_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct p_world *w = subject;
    for (unsigned i = 0u; i <= P_RODENTS; i++) {
        cmb_process_stop(&w->proc[i], NULL);
    }
}

static void run_preempt_trial(struct ref_trial *t)
{
    struct p_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->pool = cmb_resourcepool_create();
    cmb_resourcepool_initialize(w->pool, "Cheese", (uint64_t)t->servers);
    w->proc = calloc(P_RODENTS + 1u, sizeof(struct cmb_process));
    for (unsigned i = 0u; i < P_RODENTS; i++) {
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_initialize(&w->proc[i], (i < P_MICE) ? "Mouse" : "Rat", p_rodent_body, w, pri);
        cmb_process_start(&w->proc[i]);
    }
    cmb_process_initialize(&w->proc[P_RODENTS], "Cat", p_cat_body, w, 0);
    cmb_process_start(&w->proc[P_RODENTS]);
    (void)cmb_event_schedule(p_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->counter[6] = cmb_resourcepool_in_use(w->pool);
    t->objects = t->counter[0] + t->counter[1];
    for (unsigned i = 0u; i <= P_RODENTS; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    cmb_resourcepool_destroy(w->pool);
    free(w);
}


/* ------------------------------------------------- model 5: buffer + binary resource
 *
 * cmb_buffer with partial fulfilment (src/cmb_buffer.c:194-346) and cmb_resource with
 * pre-emption (src/cmb_resource.c:182-320), driven like the reference's own
 * test/test_buffer.c and test/test_resource.c: two fillers and two drainers moving
 * random amounts through a buffer of capacity `servers`, one polite and one
 * pre-empting worker sharing a binary resource, a nuisance interrupting all six,
 * an end event stopping everybody.
 * counters: [0] units put [1] units got [2] interrupted puts [3] interrupted gets
 *           [4] resource acquisitions [5] PREEMPTED received [6] sum of signals
 *           [7] final buffer level
 * sum_wait = sum over completed resource tenures of their length
 */
#define B_PROCS 6u

struct b_world {
    struct ref_trial *trl;
    struct cmb_buffer *buffer;
    struct cmb_resource *tool;
    struct cmb_process *proc;           /* B_PROCS + 1 contiguous */
    long amount_max;                    /* model 5: 8; model 12 (test/test_buffer.c as it stands): 15 */
};

static void b_note(struct b_world *w, int64_t sig)
{
    if (sig != CMB_PROCESS_SUCCESS) {
        w->trl->counter[6] += (uint64_t)sig;
    }
}

static void *b_filler_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct b_world *w = vw;
    for (;;) {
        b_note(w, cmb_process_hold(cmb_random_exponential(w->trl->arr_mean)));
        const uint64_t want = (uint64_t)cmb_random_dice(1, w->amount_max);
        uint64_t amount = want;
        const int64_t sig = cmb_buffer_put(w->buffer, &amount);
        w->trl->counter[0] += want - amount;
        if (sig != CMB_PROCESS_SUCCESS) {
            w->trl->counter[2] += 1u;
            b_note(w, sig);
        }
    }
}

static void *b_drainer_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct b_world *w = vw;
    for (;;) {
        b_note(w, cmb_process_hold(cmb_random_exponential(w->trl->srv_mean)));
        uint64_t amount = (uint64_t)cmb_random_dice(1, w->amount_max);
        const int64_t sig = cmb_buffer_get(w->buffer, &amount);
        w->trl->counter[1] += amount;
        if (sig != CMB_PROCESS_SUCCESS) {
            w->trl->counter[3] += 1u;
            b_note(w, sig);
        }
    }
}

static void *b_worker_body(struct cmb_process *me, void *vw)
{
    struct b_world *w = vw;
    const bool pushy = (me == &w->proc[5]);
    for (;;) {
        int64_t sig = pushy ? cmb_resource_preempt(w->tool) : cmb_resource_acquire(w->tool);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[4] += 1u;
            const double since = cmb_time();
            sig = cmb_process_hold(cmb_random_exponential(1.0));
            if (sig == CMB_PROCESS_PREEMPTED) {
                w->trl->counter[5] += 1u;
                b_note(w, sig);
            }
            else {
                b_note(w, sig);
                cmb_resource_release(w->tool);
                w->trl->sum_wait += cmb_time() - since;
            }
        }
        else {
            b_note(w, sig);
        }
        b_note(w, cmb_process_hold(cmb_random_exponential(1.0)));
    }
}

static void *b_nuisance_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct b_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)B_PROCS - 1);
        const int64_t sig = cmb_random_dice(1, 10);
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_interrupt(&w->proc[victim], sig, pri);
    }
}

static void b_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct b_world *w = subject;
    for (unsigned i = 0u; i <= B_PROCS; i++) {
        cmb_process_stop(&w->proc[i], NULL);
    }
}

static void run_buffer_trial(struct ref_trial *t)
{
    struct b_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->buffer = cmb_buffer_create();
    cmb_buffer_initialize(w->buffer, "Buffer", (uint64_t)t->servers);
    /* model 12 = test/test_buffer.c as it stands: three putters and three getters moving 1..15 units, the
     * buffer's level history on.  With capacity 10, means 1, duration 10000 and seed 0x34f05c64d7ad598f the
     * time-weighted level is test/reference/buffer.txt's "N 41876 Mean 4.980". */
    const bool plain = t->model == 12;
    const unsigned fillers = plain ? 3u : 2u, drainers = plain ? 3u : 2u;
    w->amount_max = plain ? 15 : 8;
    if (plain) {
        cmb_buffer_recording_start(w->buffer);
    }
    w->tool = cmb_resource_create();
    cmb_resource_initialize(w->tool, "Tool");
    w->proc = calloc(B_PROCS + 1u, sizeof(struct cmb_process));
    for (unsigned i = 0u; i < B_PROCS; i++) {
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_initialize(&w->proc[i], "Proc",
                               (i < fillers) ? b_filler_body : (i < fillers + drainers) ? b_drainer_body : b_worker_body,
                               w, pri);
        cmb_process_start(&w->proc[i]);
    }
    cmb_process_initialize(&w->proc[B_PROCS], "Nuisance", b_nuisance_body, w, 0);
    cmb_process_start(&w->proc[B_PROCS]);
    (void)cmb_event_schedule(b_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->counter[7] = cmb_buffer_level(w->buffer);
    t->objects = t->counter[1];
    if (plain) {
        /* counter[4] = time-weighted mean level (bits), max_queue = history samples with a duration */
        cmb_buffer_recording_stop(w->buffer);
        struct cmb_wtdsummary ws;
        cmb_wtdsummary_initialize(&ws);
        (void)cmb_timeseries_summarize(cmb_buffer_history(w->buffer), &ws);
        const double mean = cmb_wtdsummary_mean(&ws);
        memcpy(&t->counter[4], &mean, 8);
        t->max_queue = cmb_wtdsummary_count(&ws);
    }
    for (unsigned i = 0u; i <= B_PROCS; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    cmb_resource_destroy(w->tool);
    cmb_buffer_destroy(w->buffer);
    free(w);
}


/* ------------------------------------------------- model 6: priority queue + condition
 *
 * cmb_priorityqueue put/get/position/cancel/reprioritize (src/cmb_priorityqueue.c:189-320,
 * include/cmb_priorityqueue.h:152-185) and cmb_condition wait/signal with user
 * predicates (src/cmb_condition.c:63-167), in the manner of test/test_priorityqueue.c
 * and test/test_condition.c: two producers and a consumer on a bounded priority
 * queue, a shuffler that looks up / re-ranks / withdraws queued items by handle, a
 * tide process that changes a level and signals a condition two waiters watch with
 * different thresholds, a nuisance interrupting all seven, an end event.
 * Objects are small integers ("weights") smuggled through the void * payload.
 * counters: [0] puts ok [1] weight received [2] interrupted puts + gets
 *           [3] sum of positions + 1000 * items withdrawn [4] condition wake-ups issued
 *           [5] waiter passes [6] sum of signals [7] final queue length
 * sum_wait = sum over gets of cmb_time() * weight
 */
#define C_PROCS 7u

struct c6_world {
    struct ref_trial *trl;
    struct cmb_priorityqueue *pq;
    struct cmb_condition *cv;
    struct cmb_process *proc;           /* C_PROCS + 1 contiguous */
    uint64_t last_handle[2];
    long level;
    long threshold[2];
};

static void c6_note(struct c6_world *w, int64_t sig)
{
    if (sig != CMB_PROCESS_SUCCESS) {
        w->trl->counter[6] += (uint64_t)sig;
    }
}

static void *c6_producer_body(struct cmb_process *me, void *vw)
{
    struct c6_world *w = vw;
    const unsigned self = (unsigned)(me - w->proc);
    for (;;) {
        c6_note(w, cmb_process_hold(cmb_random_exponential(w->trl->arr_mean)));
        const long weight = cmb_random_dice(1, 9);
        const int64_t pri = cmb_random_dice(-3, 3);
        uint64_t handle = 0u;
        const int64_t sig = cmb_priorityqueue_put(w->pq, (void *)(uintptr_t)weight, pri, &handle);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[0] += 1u;
            w->last_handle[self] = handle;
        }
        else {
            w->trl->counter[2] += 1u;
            c6_note(w, sig);
        }
    }
}

static void *c6_consumer_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct c6_world *w = vw;
    for (;;) {
        c6_note(w, cmb_process_hold(cmb_random_exponential(w->trl->srv_mean)));
        void *obj = NULL;
        const int64_t sig = cmb_priorityqueue_get(w->pq, &obj);
        if (sig == CMB_PROCESS_SUCCESS) {
            const uint64_t weight = (uint64_t)(uintptr_t)obj;
            w->trl->counter[1] += weight;
            w->trl->sum_wait += cmb_time() * (double)weight;
        }
        else {
            w->trl->counter[2] += 1u;
            c6_note(w, sig);
        }
    }
}

static void *c6_shuffler_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct c6_world *w = vw;
    for (;;) {
        c6_note(w, cmb_process_hold(cmb_random_exponential(1.5)));
        const uint64_t handle = w->last_handle[cmb_random_dice(0, 1)];
        if (handle == 0u) {
            continue;
        }
        const uint64_t pos = cmb_priorityqueue_position(w->pq, handle);
        w->trl->counter[3] += pos;
        if (pos > 0u) {
            if (cmb_random_dice(0, 1) == 1) {
                cmb_priorityqueue_reprioritize(w->pq, handle, cmb_random_dice(-3, 3));
            }
            else {
                (void)cmb_priorityqueue_cancel(w->pq, handle);
                w->trl->counter[3] += 1000u;
            }
        }
    }
}

static void *c6_tide_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct c6_world *w = vw;
    for (;;) {
        c6_note(w, cmb_process_hold(cmb_random_exponential(1.0)));
        w->level = cmb_random_dice(0, 5);
        w->trl->counter[4] += cmb_condition_signal(w->cv);
    }
}

static bool c6_high_enough(const struct cmb_condition *cvp, const struct cmb_process *pp, const void *ctx)
{
    cmb_unused(pp);
    cmb_unused(cvp);
    const struct c6_world *w = *(struct c6_world *const *)ctx;
    const long *thr = ((const long *const *)ctx)[1];
    return w->level >= *thr;
}

static void *c6_waiter_body(struct cmb_process *me, void *vw)
{
    struct c6_world *w = vw;
    const unsigned self = (unsigned)(me - w->proc) - 5u;
    const void *ctx[2] = { w, &w->threshold[self] };
    for (;;) {
        bool through = true;
        while (w->level < w->threshold[self]) {
            const int64_t sig = cmb_condition_wait(w->cv, c6_high_enough, ctx);
            if (sig != CMB_PROCESS_SUCCESS) {
                c6_note(w, sig);
                through = false;
                break;
            }
        }
        if (through) {
            w->trl->counter[5] += 1u;
        }
        c6_note(w, cmb_process_hold(cmb_random_exponential(1.0)));
    }
}

static void *c6_nuisance_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct c6_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)C_PROCS - 1);
        const int64_t sig = cmb_random_dice(1, 10);
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_interrupt(&w->proc[victim], sig, pri);
    }
}

static void c6_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct c6_world *w = subject;
    for (unsigned i = 0u; i <= C_PROCS; i++) {
        cmb_process_stop(&w->proc[i], NULL);
    }
}

static void run_prioq_trial(struct ref_trial *t)
{
    struct c6_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->pq = cmb_priorityqueue_create();
    cmb_priorityqueue_initialize(w->pq, "PQ", (uint64_t)t->servers);
    w->cv = cmb_condition_create();
    cmb_condition_initialize(w->cv, "Tide");
    w->threshold[0] = 2;
    w->threshold[1] = 4;
    w->proc = calloc(C_PROCS + 1u, sizeof(struct cmb_process));
    cmb_process_func *body[C_PROCS] = { c6_producer_body, c6_producer_body, c6_consumer_body,
                                        c6_shuffler_body, c6_tide_body, c6_waiter_body, c6_waiter_body };
    for (unsigned i = 0u; i < C_PROCS; i++) {
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_initialize(&w->proc[i], "Proc", body[i], w, pri);
        cmb_process_start(&w->proc[i]);
    }
    cmb_process_initialize(&w->proc[C_PROCS], "Nuisance", c6_nuisance_body, w, 0);
    cmb_process_start(&w->proc[C_PROCS]);
    (void)cmb_event_schedule(c6_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->counter[7] = cmb_priorityqueue_length(w->pq);
    t->objects = t->counter[0];
    for (unsigned i = 0u; i <= C_PROCS; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    cmb_condition_destroy(w->cv);
    cmb_priorityqueue_destroy(w->pq);
    free(w);
}


/* ------------------------------------------------- model 7: the hold model (large event list)
 *
 * `servers` worker processes that do nothing but hold for an exponential time, plus a
 * ticker that holds exactly 1.0 (the shape of tutorial/tut_5_1.c: 1000 target processes
 * in cmb_process_hold loops and a sensor ticking once a second, SURVEY.md section 8d-5,
 * without its float32 physics), and an end event at t = num_objects that stops
 * everybody.  The future-event list holds servers + 2 entries for the whole run, so
 * every event is a pop and a push on a deep heap (src/cmi_hashheap.c:428-524).
 * counters: [0] worker wake-ups [1] ticks
 * sum_wait = sum over worker wake-ups of cmb_time()
 */
struct h_world {
    struct ref_trial *trl;
    struct cmb_process *proc;           /* servers + 1 contiguous */
    unsigned count;
};

static void *h_worker_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct h_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        w->trl->counter[0] += 1u;
        w->trl->sum_wait += cmb_time();
    }
}

static void *h_ticker_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct h_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(1.0);
        w->trl->counter[1] += 1u;
    }
}

static void h_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct h_world *w = subject;
    for (unsigned i = 0u; i < w->count; i++) {
        cmb_process_stop(&w->proc[i], NULL);
    }
}

static void run_hold_trial(struct ref_trial *t)
{
    struct h_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->count = (unsigned)t->servers + 1u;
    w->proc = calloc(w->count, sizeof(struct cmb_process));
    for (unsigned i = 0u; i + 1u < w->count; i++) {
        cmb_process_initialize(&w->proc[i], "Worker", h_worker_body, w, 0);
        cmb_process_start(&w->proc[i]);
    }
    cmb_process_initialize(&w->proc[w->count - 1u], "Ticker", h_ticker_body, w, 0);
    cmb_process_start(&w->proc[w->count - 1u]);
    (void)cmb_event_schedule(h_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->objects = t->counter[0];
    for (unsigned i = 0u; i < w->count; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    free(w);
}

/* ------------------------------------------------- model 16: impatient customers (many processes, timers, cancels)
 *
 * The oracle of cimba_b200/models/renege_model.cuh, in the manner of tutorial/tut_3_1.c's reneging: `servers`
 * customer processes with priorities drawn from 0..3 think (exponential, mean arr_mean), then ask a
 * cmb_resourcepool of (servers + 7) / 8 clerks for one unit with a patience timer running
 * (cmb_process_timer_add, exponential with mean par0, signal 17).  Served in time: cmb_process_timers_clear (the
 * timer event is cancelled by handle), service hold (mean srv_mean), release.  Not served in time: the timer resumes
 * the waiter, cmb_resourcepool_acquire unwinds and returns 17.  An end event at t = num_objects stops everybody.
 * counters: [0] served [1] reneged [2] other signals [3] clerks in use at the end; sum_wait = time in line of the served.
 */
#define RN_TIMER_RENEGING 17

struct rn_world {
    struct ref_trial *trl;
    struct cmb_resourcepool *clerks;
    struct rn_customer *cust;
    unsigned count;
    double patience_mean;
};

struct rn_customer {
    struct cmb_process proc;
    struct rn_world *world;
    double t_join;
};

static void *rn_customer_body(struct cmb_process *me, void *vw)
{
    struct rn_customer *cu = (struct rn_customer *)me;
    struct rn_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        cu->t_join = cmb_time();
        (void)cmb_process_timer_add(me, cmb_random_exponential(w->patience_mean), RN_TIMER_RENEGING);
        const int64_t sig = cmb_resourcepool_acquire(w->clerks, 1u);
        if (sig == CMB_PROCESS_SUCCESS) {
            cmb_process_timers_clear(me);
            w->trl->sum_wait += cmb_time() - cu->t_join;
            (void)cmb_process_hold(cmb_random_exponential(w->trl->srv_mean));
            cmb_resourcepool_release(w->clerks, 1u);
            w->trl->counter[0] += 1u;
        }
        else if (sig == RN_TIMER_RENEGING) {
            w->trl->counter[1] += 1u;
        }
        else {
            w->trl->counter[2] += 1u;
        }
    }
}

static void rn_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct rn_world *w = subject;
    for (unsigned i = 0u; i < w->count; i++) {
        cmb_process_stop(&w->cust[i].proc, NULL);
    }
}

static double g_renege_patience = 0.0;          /* par0, set by ref_set_param before the trials run */
static double g_warmup_time = 0.0;              /* par0 of model 19 */

void ref_set_param(int index, double value)
{
    if (index == 0) {
        g_renege_patience = value;
        g_warmup_time = value;
    }
}

static void run_renege_trial(struct ref_trial *t)
{
    struct rn_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->count = (unsigned)t->servers;
    w->patience_mean = g_renege_patience > 0.0 ? g_renege_patience : t->srv_mean;
    w->clerks = cmb_resourcepool_create();
    cmb_resourcepool_initialize(w->clerks, "Clerks", (uint64_t)((w->count + 7u) / 8u));
    w->cust = calloc(w->count, sizeof(struct rn_customer));
    for (unsigned i = 0u; i < w->count; i++) {
        const int64_t prio = cmb_random_dice(0, 3);
        w->cust[i].world = w;
        cmb_process_initialize(&w->cust[i].proc, "Customer", rn_customer_body, w, prio);
        cmb_process_start(&w->cust[i].proc);
    }
    (void)cmb_event_schedule(rn_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->objects = t->counter[0];
    t->counter[3] = cmb_resourcepool_in_use(w->clerks);
    for (unsigned i = 0u; i < w->count; i++) {
        cmb_process_terminate(&w->cust[i].proc);
    }
    free(w->cust);
    cmb_resourcepool_destroy(w->clerks);
    free(w);
}

/* ------------------------------------------------- model 17: two stations in tandem, bounded buffer between them
 *
 * The oracle of examples/tandem_user_model.cu (a model that exists only as a user-built library on the device side):
 * a source puts stamped objects into an unlimited queue; station 1 takes one, serves it (exponential, srv_mean) and
 * puts it into a queue of `servers` places - blocking while that is full; station 2 takes it and serves it
 * (uniform on 0.5 .. 1.5 srv_mean).  sum_wait = time from arrival to the end of station 2.
 */
struct td_world {
    struct ref_trial *trl;
    struct cmb_objectqueue *first, *second;
    struct cmb_process *proc[3];
};

static void *td_source_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct td_world *w = vw;
    for (uint64_t i = 0u; i < w->trl->num_objects; i++) {
        (void)cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        double *stamp = cmi_mempool_alloc(&stamp_pool);
        *stamp = cmb_time();
        (void)cmb_objectqueue_put(w->first, stamp);
    }
    return NULL;
}

static void *td_station1_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct td_world *w = vw;
    for (;;) {
        void *obj = NULL;
        (void)cmb_objectqueue_get(w->first, &obj);
        (void)cmb_process_hold(cmb_random_exponential(w->trl->srv_mean));
        (void)cmb_objectqueue_put(w->second, obj);
    }
}

static void *td_station2_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct td_world *w = vw;
    for (;;) {
        void *obj = NULL;
        (void)cmb_objectqueue_get(w->second, &obj);
        (void)cmb_process_hold(cmb_random_uniform(0.5 * w->trl->srv_mean, 1.5 * w->trl->srv_mean));
        w->trl->sum_wait += cmb_time() - *(double *)obj;
        w->trl->objects += 1u;
        cmi_mempool_free(&stamp_pool, obj);
    }
}

static void run_tandem_trial(struct ref_trial *t)
{
    struct td_world w = { .trl = t };
    w.first = cmb_objectqueue_create();
    cmb_objectqueue_initialize(w.first, "First", CMB_UNLIMITED);
    w.second = cmb_objectqueue_create();
    cmb_objectqueue_initialize(w.second, "Second", (uint64_t)t->servers);
    cmb_process_func *body[3] = { td_source_body, td_station1_body, td_station2_body };
    for (int i = 0; i < 3; i++) {
        w.proc[i] = cmb_process_create();
        cmb_process_initialize(w.proc[i], "Tandem", body[i], &w, 0);
        cmb_process_start(w.proc[i]);
    }

    pump_events(t);

    cmb_process_stop(w.proc[1], NULL);
    cmb_process_stop(w.proc[2], NULL);
    for (int i = 0; i < 3; i++) {
        cmb_process_terminate(w.proc[i]);
        cmb_process_destroy(w.proc[i]);
    }
    cmb_objectqueue_destroy(w.first);
    cmb_objectqueue_destroy(w.second);
}

/* ------------------------------------------------- model 18: test/test_resourcepool.c as it stands
 *
 * The reference's own pool test (test/test_resourcepool.c:50-305) with counters instead of log lines and the pool's usage
 * history on, as the test has it: three mice (priority_set + acquire 1..10), two rats (pre-empt 1..10), a cat interrupting
 * a random rodent with INTERRUPTED or a signal in 10..100 (cmb_random_flip decides), on a pool of `servers` units; end event
 * at t = num_objects.  With 20 units, 100 time units and cmb_random_initialize(0x34f05c64d7ad598f) this is
 * test/reference/resourcepool.txt: "N 120  Mean 19.77  StdDev 1.147 ...".  The six process structs are contiguous here
 * (the test mallocs them one after the other): the holders' tie-break by address is then by index.
 * counters[0..7] = the usage history's cmb_wtdsummary {count, min, max, m1, m2, m3, m4, wsum} (bit patterns);
 * objects = successful acquisitions + pre-emptions; sum_wait = sum of cmb_time() over them.
 */
#define CH_MICE 3u
#define CH_RATS 2u
#define CH_RODENTS (CH_MICE + CH_RATS)

struct ch_world {
    struct ref_trial *trl;
    struct cmb_resourcepool *cheese;
    struct cmb_process *proc;           /* mice, rats, cat: contiguous */
};

static void *ch_mouse_body(struct cmb_process *me, void *vw)
{
    struct ch_world *w = vw;
    uint64_t held = 0u;
    for (;;) {
        const uint64_t req = (uint64_t)cmb_random_dice(1, 10);
        const int64_t pri = cmb_random_dice(-10, 10);
        cmb_process_priority_set(me, pri);
        int64_t sig = cmb_resourcepool_acquire(w->cheese, req);
        if (sig == CMB_PROCESS_SUCCESS) {
            held += req;
            w->trl->objects += 1u;
            w->trl->sum_wait += cmb_time();
            sig = cmb_process_hold(cmb_random_exponential(1.0));
            if (sig == CMB_PROCESS_SUCCESS) {
                uint64_t rel = (uint64_t)cmb_random_dice(1, 10);
                if (rel > held) {
                    rel = held;
                }
                cmb_resourcepool_release(w->cheese, rel);
                held -= rel;
            }
            else if (sig == CMB_PROCESS_PREEMPTED) {
                held = 0u;
            }
        }
        else if (sig == CMB_PROCESS_PREEMPTED) {
            held = 0u;
        }
        sig = cmb_process_hold(cmb_random_exponential(1.0));
        if (sig == CMB_PROCESS_PREEMPTED) {
            held = 0u;
        }
    }
}

static void *ch_rat_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct ch_world *w = vw;
    uint64_t held = 0u;
    for (;;) {
        const uint64_t req = (uint64_t)cmb_random_dice(1, 10);
        int64_t sig = cmb_resourcepool_preempt(w->cheese, req);
        if (sig == CMB_PROCESS_SUCCESS) {
            held += req;
            w->trl->objects += 1u;
            w->trl->sum_wait += cmb_time();
            sig = cmb_process_hold(cmb_random_exponential(1.0));
            if (sig == CMB_PROCESS_SUCCESS) {
                uint64_t rel = (uint64_t)cmb_random_dice(1, 10);
                if (rel > held) {
                    rel = held;
                }
                cmb_resourcepool_release(w->cheese, rel);
                held -= rel;
            }
            else if (sig == CMB_PROCESS_PREEMPTED) {
                held = 0u;
            }
        }
        else if (sig == CMB_PROCESS_PREEMPTED) {
            held = 0u;
        }
        sig = cmb_process_hold(cmb_random_exponential(1.0));
        if (sig == CMB_PROCESS_PREEMPTED) {
            held = 0u;
        }
    }
}

/* cmb_random_flip keeps 64 cached bits in a function-static thread-local that survives cmb_random_initialize
 * (src/cmb_random.c:541-552): in the reference a trial's coin flips depend on what the previous trial on the same pthread
 * left in the cache.  The device gives every trial a fresh cache - what the reference does for the first trial of a thread,
 * and what its own golden run (one trial per process) sees.  To be an oracle for THAT, this driver counts its flips and
 * uses up the leftover bits before a trial starts (no generator draw happens while the cache is non-empty). */
static CMB_THREAD_LOCAL uint64_t ch_flips_taken = 0u;

static int ch_flip(void)
{
    ch_flips_taken++;
    return cmb_random_flip();
}

static void ch_flip_align(void)
{
    while ((ch_flips_taken & 63u) != 0u) {
        (void)ch_flip();
    }
}

static void *ch_cat_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct ch_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)CH_RODENTS - 1);
        const int64_t loud = cmb_random_dice(10, 100);
        const int64_t sig = ch_flip() ? CMB_PROCESS_INTERRUPTED : loud;
        cmb_process_interrupt(&w->proc[victim], sig, 0);
    }
}

static void ch_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct ch_world *w = subject;
    for (unsigned i = 0u; i <= CH_RODENTS; i++) {
        cmb_process_stop(&w->proc[i], NULL);
    }
}

static void run_cheese_trial(struct ref_trial *t)
{
    ch_flip_align();
    struct ch_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->cheese = cmb_resourcepool_create();
    cmb_resourcepool_initialize(w->cheese, "Cheese", (uint64_t)t->servers);
    cmb_resourcepool_start_recording(w->cheese);
    w->proc = calloc(CH_RODENTS + 1u, sizeof(struct cmb_process));
    for (unsigned i = 0u; i <= CH_RODENTS; i++) {
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_func *body = (i < CH_MICE) ? ch_mouse_body : ((i < CH_RODENTS) ? ch_rat_body : ch_cat_body);
        cmb_process_initialize(&w->proc[i], "Rodent", body, w, pri);
        cmb_process_start(&w->proc[i]);
    }
    (void)cmb_event_schedule(ch_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    cmb_resourcepool_stop_recording(w->cheese);
    struct cmb_wtdsummary ws;
    cmb_wtdsummary_initialize(&ws);
    (void)cmb_timeseries_summarize(cmb_resourcepool_get_history(w->cheese), &ws);
    const struct cmb_datasummary *ds = (const struct cmb_datasummary *)&ws;
    const double v[7] = { ds->min, ds->max, ds->m1, ds->m2, ds->m3, ds->m4, ws.wsum };
    t->counter[0] = ds->count;
    memcpy(&t->counter[1], v, sizeof(v));
    for (unsigned i = 0u; i <= CH_RODENTS; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    cmb_resourcepool_destroy(w->cheese);
    free(w);
}

/* ------------------------------------------------- model 8: timers, waits, observers
 *
 * The remaining asynchronous calls of cmb_process / cmb_event / cmb_resourceguard in one
 * workload, in the manner of tutorial/tut_3_1.c (reneging with cmb_process_timer_add /
 * timer_set / timers_clear, cmb_process_yield + cmb_process_resume) and
 * test/test_process.c:104-129 (cmb_process_wait_event, cmb_process_wait_process,
 * cmb_process_exit) and test/test_event.c:175-181 (cmb_event_reschedule / reprioritize):
 *   0,1 patients   hold; patience timer; cmb_resource_acquire(desk) (renege on TIMEOUT);
 *                  timer_cancel; alarm timer; service hold; timers_clear; release;
 *                  timer_set + cmb_process_yield
 *   2   clerk      a few holds, now and then cmb_process_resume(patient, 9), then
 *                  cmb_process_exit
 *   3   supervisor cmb_process_wait_process(clerk); hold; restart the FINISHED clerk
 *   4   ringer     schedules a bell event, holds, then reschedules / reprioritizes /
 *                  cancels it (waiters get CANCELLED), then cmb_process_wait_event
 *   5   listener   cmb_process_wait_event(current bell)
 *   6   watcher    cmb_condition_wait on a condition whose guard is registered as an
 *                  OBSERVER of the desk's guard (cmb_resourceguard_register,
 *                  src/cmb_resourceguard.c:231-239 forwards every signal)
 *   7   nuisance   interrupts one of 0..6 with a random signal at a random priority
 * and an end event at t = num_objects stopping all eight (bells still pending ring later).
 * counters: [0] desk acquisitions [1] reneges [2] supervisor saw the clerk finish
 *           [3] clerk jobs [4] bell rings [5] ringer ops (1 resched, 100 reprio, 10000 cancel)
 *           [6] watcher passes + 1000 * bells heard by the listener [7] sum of signals
 * sum_wait = sum of desk occupation times
 */
#define T_PROCS 8u
#define T_SIG_ALARM 77
#define T_SIG_DOZE 55
#define T_SIG_NUDGE 9

struct t_world {
    struct ref_trial *trl;
    struct cmb_resource *desk;
    struct cmb_condition *cv;
    struct cmb_process *proc;           /* T_PROCS contiguous */
    uint64_t bell;
    bool clerk_start_pending;
};

static void t_note(struct t_world *w, int64_t sig)
{
    if (sig != CMB_PROCESS_SUCCESS) {
        w->trl->counter[7] += (uint64_t)sig;
    }
}

static void *t_patient_body(struct cmb_process *me, void *vw)
{
    struct t_world *w = vw;
    for (;;) {
        t_note(w, cmb_process_hold(cmb_random_exponential(w->trl->arr_mean)));
        const uint64_t patience = cmb_process_timer_add(me, cmb_random_exponential(2.0 * w->trl->srv_mean),
                                                        CMB_PROCESS_TIMEOUT);
        int64_t sig = cmb_resource_acquire(w->desk);
        if (sig == CMB_PROCESS_SUCCESS) {
            (void)cmb_process_timer_cancel(me, patience);
            w->trl->counter[0] += 1u;
            const double since = cmb_time();
            (void)cmb_process_timer_add(me, cmb_random_exponential(3.0), T_SIG_ALARM);
            t_note(w, cmb_process_hold(cmb_random_exponential(w->trl->srv_mean)));
            cmb_process_timers_clear(me);
            cmb_resource_release(w->desk);
            w->trl->sum_wait += cmb_time() - since;
            (void)cmb_process_timer_set(me, cmb_random_exponential(0.3), T_SIG_DOZE);
            sig = cmb_process_yield();
            t_note(w, sig);
            if (sig != T_SIG_DOZE) {
                cmb_process_timers_clear(me);
            }
        }
        else if (sig == CMB_PROCESS_TIMEOUT) {
            w->trl->counter[1] += 1u;
        }
        else {
            t_note(w, sig);
            cmb_process_timers_clear(me);
        }
    }
}

static void *t_clerk_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    w->clerk_start_pending = false;
    const long jobs = cmb_random_dice(2, 5);
    for (long j = 0; j < jobs; j++) {
        t_note(w, cmb_process_hold(cmb_random_exponential(1.0)));
        if (cmb_random_dice(0, 2) == 0) {
            cmb_process_resume(&w->proc[cmb_random_dice(0, 1)], T_SIG_NUDGE);
        }
        w->trl->counter[3] += 1u;
    }
    cmb_process_exit((void *)(uintptr_t)jobs);
    return NULL;
}

static void *t_supervisor_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    struct cmb_process *clerk = &w->proc[2];
    for (;;) {
        const int64_t sig = cmb_process_wait_process(clerk);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[2] += 1u;
            t_note(w, cmb_process_hold(cmb_random_exponential(0.5)));
            if (cmb_process_status(clerk) == CMB_PROCESS_FINISHED && !w->clerk_start_pending) {
                w->clerk_start_pending = true;
                cmb_process_start(clerk);
            }
        }
        else {
            t_note(w, sig);
        }
    }
}

static void t_bell_event(void *subject, void *object)
{
    cmb_unused(object);
    struct t_world *w = subject;
    w->trl->counter[4] += 1u;
}

static void *t_ringer_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    for (;;) {
        const double when = cmb_time() + cmb_random_exponential(2.0);
        const uint64_t h = cmb_event_schedule(t_bell_event, w, NULL, when, cmb_random_dice(-2, 2));
        w->bell = h;
        t_note(w, cmb_process_hold(cmb_random_exponential(0.7)));
        if (cmb_event_is_scheduled(h)) {
            const long op = cmb_random_dice(0, 3);
            if (op == 0) {
                (void)cmb_event_reschedule(h, cmb_time() + cmb_random_exponential(1.0));
                w->trl->counter[5] += 1u;
            }
            else if (op == 1) {
                (void)cmb_event_reprioritize(h, cmb_random_dice(-5, 5));
                w->trl->counter[5] += 100u;
            }
            else if (op == 2) {
                (void)cmb_event_cancel(h);
                w->trl->counter[5] += 10000u;
            }
        }
        if (cmb_event_is_scheduled(h)) {
            t_note(w, cmb_process_wait_event(h));
        }
    }
}

static void *t_listener_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    for (;;) {
        const uint64_t h = w->bell;
        if (h != 0u && cmb_event_is_scheduled(h)) {
            const int64_t sig = cmb_process_wait_event(h);
            if (sig == CMB_PROCESS_SUCCESS) {
                w->trl->counter[6] += 1000u;
            }
            else {
                t_note(w, sig);
            }
        }
        else {
            t_note(w, cmb_process_hold(cmb_random_exponential(0.5)));
        }
    }
}

static bool t_desk_is_free(const struct cmb_condition *cvp, const struct cmb_process *pp, const void *ctx)
{
    cmb_unused(cvp);
    cmb_unused(pp);
    const struct t_world *w = ctx;
    return w->desk->holder == NULL;
}

static void *t_watcher_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    for (;;) {
        const int64_t sig = cmb_condition_wait(w->cv, t_desk_is_free, w);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[6] += 1u;
        }
        else {
            t_note(w, sig);
        }
        t_note(w, cmb_process_hold(cmb_random_exponential(0.8)));
    }
}

static void *t_nuisance_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct t_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        const long victim = cmb_random_dice(0, (long)T_PROCS - 2);
        const int64_t sig = cmb_random_dice(1, 10);
        const int64_t pri = cmb_random_dice(-5, 5);
        /* the clerk may have exited: interrupting a FINISHED process would resume a dead coroutine */
        if (cmb_process_status(&w->proc[victim]) == CMB_PROCESS_RUNNING) {
            cmb_process_interrupt(&w->proc[victim], sig, pri);
        }
    }
}

static void t_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct t_world *w = subject;
    for (unsigned i = 0u; i < T_PROCS; i++) {
        if (cmb_process_status(&w->proc[i]) == CMB_PROCESS_RUNNING) {
            cmb_process_stop(&w->proc[i], NULL);
        }
    }
}

static void run_timers_trial(struct ref_trial *t)
{
    struct t_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->desk = cmb_resource_create();
    cmb_resource_initialize(w->desk, "Desk");
    w->cv = cmb_condition_create();
    cmb_condition_initialize(w->cv, "DeskFree");
    cmb_resourceguard_register(&w->desk->guard, &w->cv->guard);
    w->proc = calloc(T_PROCS, sizeof(struct cmb_process));
    cmb_process_func *body[T_PROCS] = { t_patient_body, t_patient_body, t_clerk_body, t_supervisor_body,
                                        t_ringer_body, t_listener_body, t_watcher_body, t_nuisance_body };
    for (unsigned i = 0u; i < T_PROCS; i++) {
        const int64_t pri = (i + 1u < T_PROCS) ? cmb_random_dice(-5, 5) : 0;
        cmb_process_initialize(&w->proc[i], "Proc", body[i], w, pri);
        cmb_process_start(&w->proc[i]);
    }
    (void)cmb_event_schedule(t_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    t->objects = t->counter[0];
    for (unsigned i = 0u; i < T_PROCS; i++) {
        cmb_process_terminate(&w->proc[i]);
    }
    free(w->proc);
    (void)cmb_resourceguard_unregister(&w->desk->guard, &w->cv->guard);
    cmb_condition_destroy(w->cv);
    cmb_resource_destroy(w->desk);
    free(w);
}


/* ------------------------------------------------- model 10: the harbor (test/test_condition.c)
 *
 * The reference's own condition-variable test model (also tutorial/tut_4_1.c), the
 * "cmb_condition + divergent process" workload of BASELINE config 5 and SURVEY.md 8f-2,
 * restated against the reference API with the same process creation order, the same draws
 * in the same order and the same parameters, so that seed 0x34f05c64d7ad598f and a duration
 * of 100 years reproduce test/reference/condition.txt (N 328781 small / 109454 large ships,
 * berth and tug utilisation).  A weather and a tide process update the state once an hour
 * and signal the harbormaster condition; an arrival process creates one ship PROCESS per
 * arrival; a ship waits on the harbormaster until depth, wind, tugs and a berth all suit
 * it, takes a berth and tugs from three cmb_resourcepools, docks, unloads, undocks, joins
 * the departed list and signals Davy Jones, whose departure process collects the exit
 * value; an idle "entertainment" process ticks once a year; an end event stops them all.
 * num_objects = duration in hours, servers = tugs, arr_mean = mean inter-arrival time,
 * srv_mean = mean unloading time of a small ship (a large one takes 1.5 x that).
 * counters: [0] small ships through [1] large ships through [2] mean system time small (bits)
 *           [3] ... large (bits) [4] tug-history samples with a duration [5] time-weighted mean
 *           tugs in use (bits) [6] berth-history samples, small | large << 32 [7] ships
 *           reactivated by the harbormaster
 * sum_wait = sum of all system times in departure order; max_queue = most ships alive at once
 */
struct hb_world;

struct hb_ship {
    struct cmb_process core;            /* a ship IS a process (first member) */
    uint64_t id;
    double max_wind, min_depth;
    unsigned tugs, size;
    double t_sys;
    struct hb_ship *next_departed;
    struct hb_world *world;
};

struct hb_world {
    struct ref_trial *trl;
    double wind_magnitude, wind_direction, water_depth;
    struct cmb_process *weather, *tide, *arrivals, *departures, *dots;
    struct cmb_resourcepool *tugs, *berths[2];
    struct cmb_condition *harbormaster, *davyjones;
    struct cmi_hashheap *active;
    struct hb_ship *departed;           /* LIFO, like the test's cmi_slist */
    struct cmb_datasummary through[2];
    uint64_t alive, most_alive;
};

static void *hb_weather_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct hb_world *w = vw;
    for (;;) {
        const double gust = cmb_random_rayleigh(5.0);
        w->wind_magnitude = 0.5 * gust + 0.5 * w->wind_magnitude;
        const double d1 = cmb_random_PERT(0.0, 225.0, 360.0);
        const double d2 = cmb_random_PERT(0.0, 45.0, 360.0);
        w->wind_direction = 0.75 * d1 + 0.25 * d2;
        w->trl->counter[7] += cmb_condition_signal(w->harbormaster);
        (void)cmb_process_hold(1.0);
    }
}

static void *hb_tide_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct hb_world *w = vw;
    for (;;) {
        const double half_month = 0.5 * 29.5 * 24.0;
        const double t = fmod(cmb_time(), half_month);
        const double astro = 15.0 + 1.0 * sin(2.0 * M_PI * t / 12.4) + 0.5 * sin(2.0 * M_PI * t / 24.0)
                           + 0.25 * sin(2.0 * M_PI * t / (0.5 * 29.5 * 24));
        const double surge = 0.5 * w->wind_magnitude
                           - 0.5 * w->wind_magnitude * sin(w->wind_direction * M_PI / 180.0);
        w->water_depth = astro + surge;
        w->trl->counter[7] += cmb_condition_signal(w->harbormaster);
        (void)cmb_process_hold(1.0);
    }
}

static bool hb_can_dock(const struct cmb_condition *cvp, const struct cmb_process *pp, const void *ctx)
{
    cmb_unused(cvp);
    cmb_unused(ctx);
    const struct hb_ship *s = (const struct hb_ship *)pp;
    const struct hb_world *w = s->world;
    if (w->water_depth < s->min_depth) {
        return false;
    }
    if (w->wind_magnitude > s->max_wind) {
        return false;
    }
    if (cmb_resourcepool_available(w->tugs) < s->tugs) {
        return false;
    }
    return cmb_resourcepool_available(w->berths[s->size]) >= 1u;
}

static void *hb_ship_body(struct cmb_process *me, void *vw)
{
    struct hb_world *w = vw;
    struct hb_ship *s = (struct hb_ship *)me;
    const double t_arr = cmb_time();
    cmi_hashheap_enqueue(w->active, s, NULL, NULL, NULL, s->id, t_arr, 0u);
    if (++w->alive > w->most_alive) {
        w->most_alive = w->alive;
    }
    while (!hb_can_dock(NULL, me, NULL)) {
        (void)cmb_condition_wait(w->harbormaster, hb_can_dock, NULL);
    }
    (void)cmb_resourcepool_acquire(w->berths[s->size], 1u);
    (void)cmb_resourcepool_acquire(w->tugs, s->tugs);
    (void)cmb_process_hold(cmb_random_PERT(0.4, 0.5, 0.8));
    cmb_resourcepool_release(w->tugs, s->tugs);
    const double tua = (s->size == 0u) ? w->trl->srv_mean : 1.5 * w->trl->srv_mean;
    (void)cmb_process_hold(cmb_random_PERT(0.75 * tua, tua, 2 * tua));
    (void)cmb_resourcepool_acquire(w->tugs, s->tugs);
    (void)cmb_process_hold(cmb_random_PERT(0.4, 0.5, 0.8));
    cmb_resourcepool_release(w->berths[s->size], 1u);
    cmb_resourcepool_release(w->tugs, s->tugs);
    (void)cmi_hashheap_remove(w->active, s->id);
    w->alive--;
    s->next_departed = w->departed;
    w->departed = s;
    (void)cmb_condition_signal(w->davyjones);
    s->t_sys = cmb_time() - t_arr;
    return &s->t_sys;
}

static void *hb_arrivals_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct hb_world *w = vw;
    uint64_t cnt = 0u;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        struct hb_ship *s = calloc(1, sizeof(*s));
        s->id = ++cnt;
        s->size = cmb_random_bernoulli(0.25);
        s->max_wind = (s->size == 0u) ? 10.0 : 12.0;
        s->min_depth = (s->size == 0u) ? 8.0 : 13.0;
        s->tugs = (s->size == 0u) ? 1u : 3u;
        s->world = w;
        cmb_process_initialize(&s->core, "Ship", hb_ship_body, w, 0);
        cmb_process_start(&s->core);
    }
}

static bool hb_somebody_left(const struct cmb_condition *cvp, const struct cmb_process *pp, const void *ctx)
{
    cmb_unused(cvp);
    cmb_unused(pp);
    const struct hb_world *w = ctx;
    return w->departed != NULL;
}

static void *hb_departures_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct hb_world *w = vw;
    for (;;) {
        (void)cmb_condition_wait(w->davyjones, hb_somebody_left, w);
        struct hb_ship *s = w->departed;
        w->departed = s->next_departed;
        const double *t_sys = cmb_process_exit_value(&s->core);
        (void)cmb_datasummary_add(&w->through[s->size], *t_sys);
        w->trl->sum_wait += *t_sys;
        w->trl->counter[s->size] += 1u;
        cmb_process_terminate(&s->core);
        free(s);
    }
}

static void *hb_dots_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    cmb_unused(vw);
    for (;;) {
        (void)cmb_process_hold(24.0 * 7 * 52);          /* one (unprinted) dot per simulated year */
    }
}

static void hb_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct hb_world *w = subject;
    cmb_process_stop(w->weather, NULL);
    cmb_process_stop(w->tide, NULL);
    cmb_process_stop(w->arrivals, NULL);
    cmb_process_stop(w->departures, NULL);
    cmb_process_stop(w->dots, NULL);
    while (cmi_hashheap_count(w->active) > 0u) {
        void **item = cmi_hashheap_dequeue(w->active);
        struct hb_ship *s = item[0];
        (void)cmb_process_stop(&s->core, NULL);
        cmb_process_terminate(&s->core);
        free(s);
    }
}

static uint64_t hb_history_summary(struct cmb_resourcepool *rp, double *mean)
{
    struct cmb_wtdsummary ws;
    cmb_wtdsummary_initialize(&ws);
    const struct cmb_timeseries *hist = cmb_resourcepool_get_history(rp);
    if (cmb_timeseries_count(hist) > 0u) {
        (void)cmb_timeseries_summarize(hist, &ws);
    }
    *mean = cmb_wtdsummary_mean(&ws);
    return cmb_wtdsummary_count(&ws);
}

static void run_harbor_trial(struct ref_trial *t)
{
    struct hb_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    cmb_datasummary_initialize(&w->through[0]);
    cmb_datasummary_initialize(&w->through[1]);

    w->weather = cmb_process_create();
    cmb_process_initialize(w->weather, "Wind", hb_weather_body, w, 0);
    cmb_process_start(w->weather);
    w->tide = cmb_process_create();
    cmb_process_initialize(w->tide, "Depth", hb_tide_body, w, 0);
    cmb_process_start(w->tide);

    w->tugs = cmb_resourcepool_create();
    cmb_resourcepool_initialize(w->tugs, "Tugs", (uint64_t)t->servers);
    cmb_resourcepool_start_recording(w->tugs);
    for (int i = 0; i < 2; i++) {
        w->berths[i] = cmb_resourcepool_create();
        cmb_resourcepool_initialize(w->berths[i], (i == 0) ? "Small berth" : "Large berth", (i == 0) ? 6u : 3u);
        cmb_resourcepool_start_recording(w->berths[i]);
    }
    w->harbormaster = cmb_condition_create();
    cmb_condition_initialize(w->harbormaster, "Harbormaster");
    w->davyjones = cmb_condition_create();
    cmb_condition_initialize(w->davyjones, "Davy Jones");

    w->arrivals = cmb_process_create();
    cmb_process_initialize(w->arrivals, "Arrivals", hb_arrivals_body, w, 0);
    cmb_process_start(w->arrivals);
    w->departures = cmb_process_create();
    cmb_process_initialize(w->departures, "Departures", hb_departures_body, w, 0);
    cmb_process_start(w->departures);

    w->active = cmi_hashheap_create();
    cmi_hashheap_initialize(w->active, 3u, NULL);
    (void)cmb_event_schedule(hb_end_event, w, NULL, (double)t->num_objects, 0);

    w->dots = cmb_process_create();
    cmb_process_initialize(w->dots, "Dot", hb_dots_body, NULL, 0);
    cmb_process_start(w->dots);

    pump_events(t);

    t->objects = t->counter[0] + t->counter[1];
    t->max_queue = w->most_alive;
    const double mean_small = cmb_datasummary_mean(&w->through[0]);
    const double mean_large = cmb_datasummary_mean(&w->through[1]);
    memcpy(&t->counter[2], &mean_small, 8);
    memcpy(&t->counter[3], &mean_large, 8);
    double tug_mean, berth_mean;
    t->counter[4] = hb_history_summary(w->tugs, &tug_mean);
    memcpy(&t->counter[5], &tug_mean, 8);
    t->counter[6] = hb_history_summary(w->berths[0], &berth_mean);
    t->counter[6] |= hb_history_summary(w->berths[1], &berth_mean) << 32;

    while (w->departed != NULL) {                       /* left the harbor in the very last instant */
        struct hb_ship *s = w->departed;
        w->departed = s->next_departed;
        cmb_process_terminate(&s->core);
        free(s);
    }
    cmb_process_terminate(w->weather);  cmb_process_destroy(w->weather);
    cmb_process_terminate(w->tide);     cmb_process_destroy(w->tide);
    cmb_process_terminate(w->arrivals); cmb_process_destroy(w->arrivals);
    cmb_process_terminate(w->departures); cmb_process_destroy(w->departures);
    cmb_process_terminate(w->dots);     cmb_process_destroy(w->dots);
    cmi_hashheap_terminate(w->active);
    cmi_hashheap_destroy(w->active);
    cmb_condition_destroy(w->harbormaster);
    cmb_condition_destroy(w->davyjones);
    cmb_resourcepool_destroy(w->tugs);
    cmb_resourcepool_destroy(w->berths[0]);
    cmb_resourcepool_destroy(w->berths[1]);
    free(w);
}

/* ------------------------------------------------- model 14: test/test_resource.c as it stands
 *
 * Three "preemptable" processes with random priorities and one "preempter" (priority 0) competing for one
 * cmb_resource whose usage history is recorded; an end event stops the four.  No nuisance.  With duration 25 and
 * seed 0x34f05c64d7ad598f this is test/reference/resource.txt: history "N 30 Mean 0.9816", and Target_3 loses
 * the resource to the preempter at t = 6.3280.
 * counters: [0] acquisitions by the targets [1] PREEMPTED received [2] acquisitions by the preempter
 *           [3] time-weighted mean utilisation (bits) [4] time of the first pre-emption (bits) [5] its victim + 1
 * sum_wait = sum of the targets' completed tenures; max_queue = history samples with a duration
 */
struct r_world {
    struct ref_trial *trl;
    struct cmb_resource *res;
    struct cmb_process *proc[4];
};

static void *r_target_body(struct cmb_process *me, void *vw)
{
    struct r_world *w = vw;
    for (;;) {
        int64_t sig = cmb_resource_acquire(w->res);
        if (sig == CMB_PROCESS_SUCCESS) {
            w->trl->counter[0] += 1u;
            const double since = cmb_time();
            sig = cmb_process_hold(cmb_random_exponential(1.0));
            if (sig == CMB_PROCESS_SUCCESS) {
                cmb_resource_release(w->res);
                w->trl->sum_wait += cmb_time() - since;
            }
            else {
                w->trl->counter[1] += 1u;
                if (w->trl->counter[5] == 0u) {
                    const double when = cmb_time();
                    memcpy(&w->trl->counter[4], &when, 8);
                    for (unsigned i = 0u; i < 3u; i++) {
                        if (w->proc[i] == me) {
                            w->trl->counter[5] = i + 1u;
                        }
                    }
                }
            }
        }
        (void)cmb_process_hold(cmb_random_exponential(1.0));
    }
}

static void *r_preempter_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct r_world *w = vw;
    for (;;) {
        (void)cmb_resource_preempt(w->res);
        w->trl->counter[2] += 1u;
        (void)cmb_process_hold(cmb_random_exponential(1.0));
        cmb_resource_release(w->res);
        (void)cmb_process_hold(cmb_random_exponential(1.0));
    }
}

static void r_end_event(void *subject, void *object)
{
    cmb_unused(object);
    struct r_world *w = subject;
    for (unsigned i = 0u; i < 4u; i++) {
        cmb_process_stop(w->proc[i], NULL);
    }
}

static void run_resource_trial(struct ref_trial *t)
{
    struct r_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->res = cmb_resource_create();
    cmb_resource_initialize(w->res, "Resource_1");
    cmb_resource_start_recording(w->res);
    for (unsigned i = 0u; i < 3u; i++) {
        w->proc[i] = cmb_process_create();
        const int64_t pri = cmb_random_dice(-5, 5);
        cmb_process_initialize(w->proc[i], "Target", r_target_body, w, pri);
        cmb_process_start(w->proc[i]);
    }
    w->proc[3] = cmb_process_create();
    cmb_process_initialize(w->proc[3], "Preempter", r_preempter_body, w, 0);
    cmb_process_start(w->proc[3]);
    (void)cmb_event_schedule(r_end_event, w, NULL, (double)t->num_objects, 0);

    pump_events(t);

    cmb_resource_stop_recording(w->res);
    struct cmb_wtdsummary ws;
    cmb_wtdsummary_initialize(&ws);
    (void)cmb_timeseries_summarize(cmb_resource_history(w->res), &ws);
    const double mean = cmb_wtdsummary_mean(&ws);
    memcpy(&t->counter[3], &mean, 8);
    t->max_queue = cmb_wtdsummary_count(&ws);
    t->objects = t->counter[0] + t->counter[2];
    for (unsigned i = 0u; i < 4u; i++) {
        cmb_process_terminate(w->proc[i]);
        cmb_process_destroy(w->proc[i]);
    }
    cmb_resource_destroy(w->res);
    free(w);
}

/* ------------------------------------------------- model 19: the trial of the reference's first tutorial
 *
 * tutorial/tut_1_7.c run_MM1_trial (:155-222) and its two processes and three events (:69-151): an M/M/1 queue held in a
 * cmb_buffer (amounts of 1), the level history switched on by an event at the warm-up time and off by an event at warm-up +
 * duration, where an end event of priority -100 stops both processes.  The tutorial's result is the time-weighted mean level.
 *   arr_mean / srv_mean = 1 / arr_rate, 1 / srv_rate; num_objects = duration; par0 = warm-up time.
 * counters[0..7] = the eight words of the history's cmb_wtdsummary; objects = units put; sum_wait = units got.
 */
struct u_world {
    struct ref_trial *trl;
    struct cmb_buffer *que;
    struct cmb_process *arr, *srv;
};

static void *u_arrival_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct u_world *w = vw;
    for (;;) {
        (void)cmb_process_hold(cmb_random_exponential(w->trl->arr_mean));
        uint64_t n = 1u;
        (void)cmb_buffer_put(w->que, &n);
        w->trl->objects += 1u;
    }
}

static void *u_service_body(struct cmb_process *me, void *vw)
{
    cmb_unused(me);
    struct u_world *w = vw;
    for (;;) {
        uint64_t n = 1u;
        (void)cmb_buffer_get(w->que, &n);
        w->trl->sum_wait += 1.0;
        (void)cmb_process_hold(cmb_random_exponential(w->trl->srv_mean));
    }
}

static void u_start_rec(void *subject, void *object)
{
    cmb_unused(subject);
    cmb_buffer_recording_start(((struct u_world *)object)->que);
}

static void u_stop_rec(void *subject, void *object)
{
    cmb_unused(subject);
    cmb_buffer_recording_stop(((struct u_world *)object)->que);
}

static void u_end_sim(void *subject, void *object)
{
    cmb_unused(subject);
    struct u_world *w = object;
    cmb_process_stop(w->arr, NULL);
    cmb_process_stop(w->srv, NULL);
}

static void run_tutorial1_trial(struct ref_trial *t)
{
    struct u_world w = { .trl = t };
    w.que = cmb_buffer_create();
    cmb_buffer_initialize(w.que, "Queue", CMB_UNLIMITED);
    w.arr = cmb_process_create();
    cmb_process_initialize(w.arr, "Arrival", u_arrival_body, &w, 0);
    cmb_process_start(w.arr);
    w.srv = cmb_process_create();
    cmb_process_initialize(w.srv, "Service", u_service_body, &w, 0);
    cmb_process_start(w.srv);
    double when = g_warmup_time;
    (void)cmb_event_schedule(u_start_rec, NULL, &w, when, 0);
    when += (double)t->num_objects;
    (void)cmb_event_schedule(u_stop_rec, NULL, &w, when, 0);
    (void)cmb_event_schedule(u_end_sim, NULL, &w, when, -100);

    pump_events(t);

    struct cmb_wtdsummary ws;
    cmb_wtdsummary_initialize(&ws);
    (void)cmb_timeseries_summarize(cmb_buffer_history(w.que), &ws);
    const struct cmb_datasummary *ds = (const struct cmb_datasummary *)&ws;
    const double v[7] = { ds->min, ds->max, ds->m1, ds->m2, ds->m3, ds->m4, ws.wsum };
    t->counter[0] = ds->count;
    memcpy(&t->counter[1], v, sizeof(v));
    cmb_process_terminate(w.srv);
    cmb_process_destroy(w.srv);
    cmb_process_terminate(w.arr);
    cmb_process_destroy(w.arr);
    cmb_buffer_terminate(w.que);
    cmb_buffer_destroy(w.que);
}

/* -------------------------------------------------------------- dispatcher */

static void stock_noop_event(void *subject, void *object)
{
    cmb_unused(subject);
    cmb_unused(object);
}

static void pump_events(struct ref_trial *t)
{
    if (t->stock) {
        /* the benchmark's own dispatcher call, benchmark/MM1_multi.c:113.  The pop count comes for free
         * afterwards: event handles are the event list's running enqueue counter (src/cmi_hashheap.c:449-453),
         * the list ran dry, and models 0-2 never cancel an event, so a probe event's handle minus one is the
         * number of events executed. */
        cmb_event_queue_execute();
        const uint64_t probe = cmb_event_schedule(stock_noop_event, NULL, NULL, cmb_time(), 0);
        (void)cmb_event_cancel(probe);
        t->events = probe - 1u;
        t->t_end = cmb_time();
        return;
    }
    uint64_t n = 0u;
    for (;;) {
        const uint64_t depth = cmb_event_queue_count();
        if (depth > t->max_fel) {
            t->max_fel = depth;
        }
        if (!cmb_event_execute_next()) {
            break;
        }
        if (n < t->trace_cap) {
            t->trace_key[n] = cmb_event_current();
            t->trace_time[n] = cmb_time();
        }
        n++;
    }
    t->events = n;
    t->t_end = cmb_time();
}

static void run_queue_trial(struct ref_trial *t)
{
    struct q_world w = { .trl = t };
    w.queue = cmb_objectqueue_create();
    cmb_objectqueue_initialize(w.queue, "Queue", CMB_UNLIMITED);
    if (t->model == 9) {
        /* model 9 = model 0 with the queue's history switched on, as tutorial/tut_1_*.c and
         * test/test_cimba.c do: every put/get appends (length, cmb_time()) to a cmb_timeseries */
        cmb_objectqueue_recording_start(w.queue);
    }
    w.source = cmb_process_create();
    cmb_process_initialize(w.source, "Arrival", t->stock ? q_source_body_stock : q_source_body, &w, 0);
    cmb_process_start(w.source);
    w.server = cmb_process_create();
    cmb_process_initialize(w.server, "Service", q_server_body, &w, 0);
    cmb_process_start(w.server);

    pump_events(t);

    if (t->model == 9) {
        /* time-weighted queue length: cmb_timeseries_summarize -> cmb_wtdsummary_add per sample
         * (src/cmb_timeseries.c:167-188); exported as bit patterns in counter[0..7] */
        cmb_objectqueue_recording_stop(w.queue);
        struct cmb_wtdsummary ws;
        cmb_wtdsummary_initialize(&ws);
        (void)cmb_timeseries_summarize(cmb_objectqueue_history(w.queue), &ws);
        const struct cmb_datasummary *ds = (const struct cmb_datasummary *)&ws;
        const double v[7] = { ds->min, ds->max, ds->m1, ds->m2, ds->m3, ds->m4, ws.wsum };
        t->counter[0] = ds->count;
        memcpy(&t->counter[1], v, sizeof(v));
    }
    cmb_process_stop(w.server, NULL);
    cmb_process_terminate(w.source);
    cmb_process_terminate(w.server);
    cmb_process_destroy(w.source);
    cmb_process_destroy(w.server);
    cmb_objectqueue_destroy(w.queue);
}

static void run_pool_trial(struct ref_trial *t)
{
    struct c_world *w = calloc(1, sizeof(*w));
    w->trl = t;
    w->pool = cmb_resourcepool_create();
    cmb_resourcepool_initialize(w->pool, "Servers", (uint64_t)t->servers);
    w->source = cmb_process_create();
    cmb_process_initialize(w->source, "Arrival", c_source_body, w, 0);
    cmb_process_start(w->source);

    pump_events(t);

    cmb_process_terminate(w->source);
    cmb_process_destroy(w->source);
    for (unsigned i = 0u; i < w->all_count; i++) {
        cmb_process_terminate(&w->all_list[i]->proc);
        free(w->all_list[i]);
    }
    t->max_queue = w->all_count;
    cmb_resourcepool_destroy(w->pool);
    free(w);
}

static void run_trial(void *vt)
{
    struct ref_trial *t = vt;
    t->events = 0u;
    t->objects = 0u;
    t->sum_wait = 0.0;
    t->t_end = 0.0;
    t->max_fel = 0u;
    t->max_queue = 0u;

    cmb_logger_flags_off(CMB_LOGGER_INFO);
    cmb_random_initialize(t->seed);
    cmb_event_queue_initialize(0.0);
    memset(t->counter, 0, sizeof(t->counter));
    if (t->model == 19) {
        run_tutorial1_trial(t);
    }
    else if (t->model == 18) {
        run_cheese_trial(t);
    }
    else if (t->model == 17) {
        run_tandem_trial(t);
    }
    else if (t->model == 16) {
        run_renege_trial(t);
    }
    else if (t->model == 14) {
        run_resource_trial(t);
    }
    else if (t->model == 10) {
        run_harbor_trial(t);
    }
    else if (t->model == 8) {
        run_timers_trial(t);
    }
    else if (t->model == 7) {
        run_hold_trial(t);
    }
    else if (t->model == 6) {
        run_prioq_trial(t);
    }
    else if (t->model == 5 || t->model == 12) {
        run_buffer_trial(t);
    }
    else if (t->model == 4) {
        run_preempt_trial(t);
    }
    else if (t->model == 3 || t->model == 11 || t->model == 13) {
        run_guarded_trial(t);
    }
    else if (t->model == 2) {
        run_pool_trial(t);
    }
    else {
        run_queue_trial(t);
    }
    cmb_event_queue_terminate();
}

/* ------------------------------------------------------------------ C ABI */

struct ref_result {
    uint64_t events;
    uint64_t objects;
    double t_end;
    double sum_wait;
    uint64_t max_fel;
    uint64_t max_queue;
    uint64_t counter[8];
};

/*
 * Run trials [first, first + count) of an experiment whose per-trial seed is
 * cmb_random_fmix64(master_seed, global trial index).
 * parallel != 0: through the reference's own executive, cimba_run_experiment()
 * (one pthread per logical core, src/cimba.c:171); parallel == 0: serially on
 * the calling thread.
 */
int ref_run_trials(int model, int servers, uint64_t master_seed,
                   uint64_t first, uint64_t count, uint64_t num_objects,
                   double arr_mean, double srv_mean, int parallel,
                   struct ref_result *out)
{
    struct ref_trial *exp = calloc(count, sizeof(*exp));
    if (exp == NULL) {
        return -1;
    }
    for (uint64_t i = 0u; i < count; i++) {
        exp[i].seed = cmb_random_fmix64(master_seed, first + i);
        exp[i].num_objects = num_objects;
        exp[i].arr_mean = arr_mean;
        exp[i].srv_mean = srv_mean;
        exp[i].model = model;
        exp[i].servers = servers;
    }
    if (parallel) {
        cimba_run_experiment(exp, count, sizeof(*exp), run_trial);
    }
    else {
        for (uint64_t i = 0u; i < count; i++) {
            run_trial(&exp[i]);
        }
    }
    for (uint64_t i = 0u; i < count; i++) {
        out[i].events = exp[i].events;
        out[i].objects = exp[i].objects;
        out[i].t_end = exp[i].t_end;
        out[i].sum_wait = exp[i].sum_wait;
        out[i].max_fel = exp[i].max_fel;
        out[i].max_queue = exp[i].max_queue;
        memcpy(out[i].counter, exp[i].counter, sizeof(out[i].counter));
    }
    free(exp);
    return 0;
}

/*
 * The timed leg of bench.py: trials [first, first + count) of model 0 (M/M/1), 1 (G/G/1) or 2 (M/M/c) run the way
 * the reference's benchmarks run them - stock process bodies, cmb_event_queue_execute() - through
 * cimba_run_experiment() (threads == 0: one pthread per logical core, src/cimba.c:171) or serially on the calling
 * thread (threads == 1: the benchmark/MM1_single.c row).  Results as ref_run_trials; max_fel / max_queue stay 0.
 */
int ref_bench_trials(int model, int servers, uint64_t master_seed,
                     uint64_t first, uint64_t count, uint64_t num_objects,
                     double arr_mean, double srv_mean, int threads,
                     struct ref_result *out)
{
    if (model < 0 || model > 2) {
        return -2;
    }
    struct ref_trial *exp = calloc(count, sizeof(*exp));
    if (exp == NULL) {
        return -1;
    }
    for (uint64_t i = 0u; i < count; i++) {
        exp[i].seed = cmb_random_fmix64(master_seed, first + i);
        exp[i].num_objects = num_objects;
        exp[i].arr_mean = arr_mean;
        exp[i].srv_mean = srv_mean;
        exp[i].model = model;
        exp[i].servers = servers;
        exp[i].stock = 1;
    }
    if (threads == 1) {
        for (uint64_t i = 0u; i < count; i++) {
            run_trial(&exp[i]);
        }
    }
    else {
        cimba_run_experiment(exp, count, sizeof(*exp), run_trial);
    }
    for (uint64_t i = 0u; i < count; i++) {
        out[i].events = exp[i].events;
        out[i].objects = exp[i].objects;
        out[i].t_end = exp[i].t_end;
        out[i].sum_wait = exp[i].sum_wait;
        out[i].max_fel = 0u;
        out[i].max_queue = 0u;
        memset(out[i].counter, 0, sizeof(out[i].counter));
    }
    free(exp);
    return 0;
}

/* One trial with an explicit seed, recording (key, clock) of the first pops. */
int ref_trace_trial(int model, int servers, uint64_t seed, uint64_t num_objects,
                    double arr_mean, double srv_mean, uint64_t trace_cap,
                    uint64_t *trace_key, double *trace_time,
                    struct ref_result *out)
{
    struct ref_trial t = { .seed = seed, .num_objects = num_objects,
                           .arr_mean = arr_mean, .srv_mean = srv_mean,
                           .model = model, .servers = servers,
                           .trace_cap = trace_cap, .trace_key = trace_key,
                           .trace_time = trace_time };
    run_trial(&t);
    out->events = t.events;
    out->objects = t.objects;
    out->t_end = t.t_end;
    out->sum_wait = t.sum_wait;
    out->max_fel = t.max_fel;
    out->max_queue = t.max_queue;
    memcpy(out->counter, t.counter, sizeof(out->counter));
    return 0;
}

uint64_t ref_fmix64(uint64_t seed, uint64_t nonce)
{
    return cmb_random_fmix64(seed, nonce);
}

extern uint32_t cmi_cpu_cores(void);   /* src/port/x86-64/linux/cmi_cpu_cores.c:23 */

int ref_cpu_cores(void)
{
    return (int)cmi_cpu_cores();
}

/*
 * Draw n variates after cmb_random_initialize(seed).
 * kind 0: raw sfc64 (bit pattern stored in the double slot)
 *      1: cmb_random_exponential(p0)     2: cmb_random_std_normal()
 *      3: cmb_random()                   4: cmb_random_normal(p0, p1)
 *      5: cmb_random_erlang((unsigned)p0, p1)
 *      6: cmb_random_uniform(p0, p1)     7: cmb_random_dice((long)p0,(long)p1) as double
 *      8: cmb_random_bernoulli(p0) as 0/1
 */
int ref_rng_draws(uint64_t seed, int kind, double p0, double p1, uint64_t n, double *out)
{
    cmb_random_initialize(seed);
    for (uint64_t i = 0u; i < n; i++) {
        switch (kind) {
        case 0: { uint64_t u = cmb_random_sfc64(); memcpy(&out[i], &u, 8); break; }
        case 1: out[i] = cmb_random_exponential(p0); break;
        case 2: out[i] = cmb_random_std_normal(); break;
        case 3: out[i] = cmb_random(); break;
        case 4: out[i] = cmb_random_normal(p0, p1); break;
        case 5: out[i] = cmb_random_erlang((unsigned)p0, p1); break;
        case 6: out[i] = cmb_random_uniform(p0, p1); break;
        case 7: out[i] = (double)cmb_random_dice((long)p0, (long)p1); break;
        case 8: out[i] = (double)cmb_random_bernoulli(p0); break;
        default: return -1;
        }
    }
    return 0;
}

/* The rest of cmb_random (include/cmb_random.h:189-940, src/cmb_random.c:299-766): kinds 9..33,
 * parameters in par[] (arrays inline: a count, then the values).  The same kind numbers are
 * implemented by oracle/port (port_rng_draws_ex) and the CUDA library (cimba_b200_rng_draws_ex). */
int ref_rng_draws_ex(uint64_t seed, int kind, const double *par, uint32_t npar, uint64_t n, double *out)
{
    cmb_unused(npar);
    cmb_random_initialize(seed);
    struct cmb_random_alias *alias = NULL;
    if (kind == 30) {
        alias = cmb_random_alias_create((unsigned)par[0], &par[1]);
    }
    for (uint64_t i = 0u; i < n; i++) {
        switch (kind) {
        case 9:  out[i] = cmb_random_triangular(par[0], par[1], par[2]); break;
        case 10: out[i] = cmb_random_lognormal(par[0], par[1]); break;
        case 11: out[i] = cmb_random_logistic(par[0], par[1]); break;
        case 12: out[i] = cmb_random_cauchy(par[0], par[1]); break;
        case 13: out[i] = cmb_random_hypoexponential((unsigned)par[0], &par[1]); break;
        case 14: out[i] = cmb_random_hyperexponential((unsigned)par[0], &par[1], &par[1 + (unsigned)par[0]]); break;
        case 15: out[i] = cmb_random_gamma(par[0], par[1]); break;
        case 16: out[i] = cmb_random_beta(par[0], par[1], par[2], par[3]); break;
        case 17: out[i] = cmb_random_PERT(par[0], par[1], par[2]); break;
        case 18: out[i] = cmb_random_weibull(par[0], par[1]); break;
        case 19: out[i] = cmb_random_pareto(par[0], par[1]); break;
        case 20: out[i] = cmb_random_chisquared(par[0]); break;
        case 21: out[i] = cmb_random_F_dist(par[0], par[1]); break;
        case 22: out[i] = cmb_random_t_dist(par[0], par[1], par[2]); break;
        case 23: out[i] = cmb_random_rayleigh(par[0]); break;
        case 24: out[i] = (double)cmb_random_flip(); break;
        case 25: out[i] = (double)cmb_random_geometric(par[0]); break;
        case 26: out[i] = (double)cmb_random_binomial((unsigned)par[0], par[1]); break;
        case 27: out[i] = (double)cmb_random_negative_binomial((unsigned)par[0], par[1]); break;
        case 28: out[i] = (double)cmb_random_poisson(par[0]); break;
        case 29: out[i] = (double)cmb_random_loaded_dice((unsigned)par[0], &par[1]); break;
        case 30: out[i] = (double)cmb_random_alias_sample(alias); break;
        case 31: out[i] = cmb_random_std_gamma(par[0]); break;
        case 32: out[i] = cmb_random_PERT_mod(par[0], par[1], par[2], par[3]); break;
        case 33: out[i] = (double)cmb_random_pascal((unsigned)par[0], par[1]); break;
        default: return -1;
        }
    }
    if (alias != NULL) {
        cmb_random_alias_destroy(alias);
    }
    return 0;
}

/* Summary parity helpers: out = {count, min, max, m1, m2, m3, m4} */
static void export_summary(const struct cmb_datasummary *ds, double *out)
{
    out[0] = (double)ds->count;
    out[1] = ds->min;
    out[2] = ds->max;
    out[3] = ds->m1;
    out[4] = ds->m2;
    out[5] = ds->m3;
    out[6] = ds->m4;
}

int ref_datasummary_of(const double *x, uint64_t n, double *out)
{
    struct cmb_datasummary ds;
    cmb_datasummary_initialize(&ds);
    for (uint64_t i = 0u; i < n; i++) {
        cmb_datasummary_add(&ds, x[i]);
    }
    export_summary(&ds, out);
    return 0;
}

/* Summarise x[0..na) and x[na..n) separately, merge (src/cmb_datasummary.c:93) */
int ref_datasummary_split_merge(const double *x, uint64_t na, uint64_t n, double *out)
{
    struct cmb_datasummary a, b, m;
    cmb_datasummary_initialize(&a);
    cmb_datasummary_initialize(&b);
    cmb_datasummary_initialize(&m);
    for (uint64_t i = 0u; i < na; i++) {
        cmb_datasummary_add(&a, x[i]);
    }
    for (uint64_t i = na; i < n; i++) {
        cmb_datasummary_add(&b, x[i]);
    }
    cmb_datasummary_merge(&m, &a, &b);
    export_summary(&m, out);
    return 0;
}

/* Weighted twins: out = {count, min, max, m1, m2, m3, m4, wsum} */
static void export_wsummary(const struct cmb_wtdsummary *ws, double *out)
{
    export_summary(&ws->ds, out);
    out[7] = ws->wsum;
}

int ref_wtdsummary_of(const double *x, const double *w, uint64_t n, double *out)
{
    struct cmb_wtdsummary ws;
    cmb_wtdsummary_initialize(&ws);
    for (uint64_t i = 0u; i < n; i++) {
        cmb_wtdsummary_add(&ws, x[i], w[i]);
    }
    export_wsummary(&ws, out);
    return 0;
}

int ref_wtdsummary_split_merge(const double *x, const double *w, uint64_t na,
                               uint64_t n, double *out)
{
    struct cmb_wtdsummary a, b, m;
    cmb_wtdsummary_initialize(&a);
    cmb_wtdsummary_initialize(&b);
    cmb_wtdsummary_initialize(&m);
    for (uint64_t i = 0u; i < na; i++) {
        cmb_wtdsummary_add(&a, x[i], w[i]);
    }
    for (uint64_t i = na; i < n; i++) {
        cmb_wtdsummary_add(&b, x[i], w[i]);
    }
    cmb_wtdsummary_merge(&m, &a, &b);
    export_wsummary(&m, out);
    return 0;
}
