// general.cuh - the general (cancel / interrupt / stop / priority) half of the
// simulation kernel, device side, for models that need more than "every process
// owns one pending event".
//
// Reference pieces restated here:
//   cmi_hashheap        src/cmi_hashheap.c:277-370 (sift loops), :428-478 (enqueue),
//                       :486-524 (dequeue), :529-579 (remove by key)
//   cmb_event           src/cmb_event.c:123-140 (schedule), :285-302 (cancel),
//                       :385-425 (pattern cancel by subject)
//   cmb_process         src/cmb_process.c:220-260 (awaitables), :262-285 + 316-349
//                       (hold and its interrupted branch), :581-620 (cancel awaiteds),
//                       :628-666 (interrupt), :698-723 (stop)
//   cmb_resourceguard   src/cmb_resourceguard.c:71-90 (order), :125-163 (wait),
//                       :202-226 (signal), :270-290 (remove)
//
// Unlike the event-order-only containers in engine.cuh, the heaps here keep the
// reference's physical layout (1-based binary heap, slot 0 = last popped, same
// sift and remove steps).  That matters for the guard wait lists: their
// comparator is not a strict weak order when priorities differ (SURVEY.md quirk
// 1), so which waiter reaches the head depends on the exact sequence of swaps.
//
// Storage: one GeneralState per trial in HBM (lane per trial, L1/L2-cached).
// This path is about coverage and bit parity, not peak throughput.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "distributions.cuh"
#include "engine.cuh"
#include "rng.cuh"

namespace cimba_b200 {

enum : uint32_t { ACT_WAKE_INTERRUPT = 4u, ACT_USER = 5u,
                  ACT_WAKE_PREEMPT = 6u,         // wakeup_event_preempt, src/cmb_resource.c:256-268
                  ACT_WAKE_CONDITION = 7u,       // wakeup_event_condition, src/cmb_condition.c:85-103
                  ACT_WAKE_PROCESS = 8u,         // wakeup_event_process, src/cmb_process.c:386-410
                  ACT_WAKE_EVENT = 9u,           // wakeup_event_event, src/cmb_event.c:176-198
                  ACT_RESUME = 10u,              // resume_event, src/cmb_process.c:731-745
                  ACT_BELL = 11u };              // a user event with no process behind it (model 8)
enum : uint32_t { AWAIT_TIME = 0u, AWAIT_RESOURCE = 1u, AWAIT_PROCESS = 2u, AWAIT_EVENT = 3u };
enum : uint32_t { PROC_CREATED = 0u, PROC_RUNNING = 1u, PROC_FINISHED = 2u };

struct HeapTag {                // cmi_heap_tag (src/cmi_hashheap.h:53-59) packed to 24 bytes
    double   d;                 // rank_d64: event time / guard entry time
    uint32_t key;               // hash_key
    int32_t  prio;              // rank_i64
    uint16_t act;               // item[0]: action id
    uint16_t subj;              // item[1]: process index (0xffff = the model itself)
    int32_t  arg;               // item[2]: signal
};

// default_compare, src/cmi_hashheap.c:55-80
struct EventOrder {
    static __device__ __forceinline__ bool before(const HeapTag &a, const HeapTag &b)
    {
        if (a.d < b.d) return true;
        if (a.d > b.d) return false;
        if (a.prio > b.prio) return true;
        if (a.prio < b.prio) return false;
        return a.key < b.key;
    }
};

// guard_queue_check, src/cmb_resourceguard.c:71-90 - INCLUDING the fall-through when
// a has the lower priority
struct GuardOrder {
    static __device__ __forceinline__ bool before(const HeapTag &a, const HeapTag &b)
    {
        if (a.prio > b.prio) return true;
        if (a.d < b.d) return true;
        if (a.key < b.key) return true;
        return false;
    }
};

// holder_queue_check, src/cmb_resourcepool.c:75-92: lowest priority first (the likeliest
// victim of a pre-emption), then the LARGER key first (the reference
// keys holders by process address; here by process index + 1)
struct HolderOrder {
    static __device__ __forceinline__ bool before(const HeapTag &a, const HeapTag &b)
    {
        if (a.prio < b.prio) return true;
        if (a.prio == b.prio && a.key > b.key) return true;
        return false;
    }
};

// compare_func of cmb_priorityqueue, src/cmb_priorityqueue.c:43-54: priority desc, then FIFO
struct PrioOrder {
    static __device__ __forceinline__ bool before(const HeapTag &a, const HeapTag &b)
    {
        if (a.prio != b.prio) return a.prio > b.prio;
        return a.key < b.key;
    }
};

template <int CAP, class Order>
struct BinHeap {
    HeapTag  slot[CAP + 1];     // 1-based; slot[0] = last popped
    uint32_t count;
    uint32_t issued;            // item_counter

    __device__ void clear()
    {
        count = 0u;
        issued = 0u;
    }

    // The sift / remove routines are kept out of line: the general-path kernels call them from dozens of
    // places, and inlined copies made those kernels 270-390 KB of code - every warp then waits on the
    // instruction cache (profiles/r01_harbor.md).
    __device__ __noinline__ void sift_up(uint32_t k)    // heap_up, :277-316
    {
        const HeapTag moving = slot[k];
        uint32_t parent;
        while ((parent = (k >> 1)) > 0u) {
            if (!Order::before(moving, slot[parent])) {
                break;
            }
            slot[k] = slot[parent];
            k = parent;
        }
        slot[k] = moving;
    }

    __device__ __noinline__ void sift_down(uint32_t k)  // heap_down, :321-370
    {
        const HeapTag moving = slot[k];
        const uint32_t last_parent = count >> 1;
        while (k <= last_parent) {
            uint32_t child = k << 1;
            if (child + 1u <= count && Order::before(slot[child + 1u], slot[child])) {
                child++;
            }
            if (Order::before(moving, slot[child])) {
                break;
            }
            slot[k] = slot[child];
            k = child;
        }
        slot[k] = moving;
    }

    // cmi_hashheap_enqueue; key 0 = issue the next one.  Returns 0 on overflow.
    __device__ __noinline__ uint32_t push(uint32_t key, double d, int32_t prio, uint32_t act, uint32_t subj, int32_t arg)
    {
        issued += 1u;
        if (key == 0u) {
            key = issued;
        }
        if (count >= (uint32_t)CAP) {
            return 0u;
        }
        const uint32_t at = ++count;
        HeapTag t;
        t.d = d;
        t.key = key;
        t.prio = prio;
        t.act = (uint16_t)act;
        t.subj = (uint16_t)subj;
        t.arg = arg;
        slot[at] = t;
        sift_up(at);
        return key;
    }

    __device__ __noinline__ bool pop()                  // cmi_hashheap_dequeue
    {
        if (count == 0u) {
            return false;
        }
        slot[0] = slot[1];
        if (count > 1u) {
            slot[1] = slot[count];
            count--;
            if (count > 1u) {
                sift_down(1u);
            }
        }
        else {
            count = 0u;
        }
        return true;
    }

    __device__ uint32_t find(uint32_t key) const        // cmi_hash_find_index (by scan): 0 = absent
    {
        for (uint32_t k = 1u; k <= count; k++) {
            if (slot[k].key == key) {
                return k;
            }
        }
        return 0u;
    }

    // cmi_hashheap_reprioritize, src/cmi_hashheap.c:679-711
    __device__ __noinline__ void reprioritize(uint32_t key, double d, int32_t prio)
    {
        const uint32_t at = find(key);
        if (at == 0u) {
            return;
        }
        const HeapTag old = slot[at];
        slot[at].d = d;
        slot[at].prio = prio;
        if (Order::before(old, slot[at])) {
            sift_down(at);
        }
        else {
            sift_up(at);
        }
    }

    __device__ __noinline__ bool remove(uint32_t key)   // cmi_hashheap_remove (lookup by scan)
    {
        uint32_t at = 0u;
        for (uint32_t k = 1u; k <= count; k++) {
            if (slot[k].key == key) {
                at = k;
                break;
            }
        }
        if (at == 0u) {
            return false;
        }
        if (at == count) {
            count--;
            return true;
        }
        const bool down = Order::before(slot[at], slot[count]);
        slot[at] = slot[count];
        count--;
        if (down) {
            sift_down(at);
        }
        else {
            sift_up(at);
        }
        return true;
    }
};

constexpr int GEN_FEL_CAP = 31;
constexpr int GEN_GUARD_CAP = 15;
constexpr int GEN_MAX_PROCS = 8;
constexpr int GEN_MAX_AWAITS = 4;
constexpr int GEN_EVENT_WAITERS = 16;

using EventHeap = BinHeap<GEN_FEL_CAP, EventOrder>;
using GuardHeap = BinHeap<GEN_GUARD_CAP, GuardOrder>;
using HolderHeap = BinHeap<GEN_MAX_PROCS, HolderOrder>;     // tag.subj = holder, tag.arg = amount held
using PrioHeap = BinHeap<GEN_GUARD_CAP, PrioOrder>;         // tag.arg = the queued object (a small integer)

struct GenProc {                // struct cmb_process (include/cmb_process.h:116-123), the parts that act
    uint32_t pc, status, kind;
    int32_t  prio;
    uint32_t n_awaits;
    uint32_t await_type[GEN_MAX_AWAITS];    // most recent first (cmi_slist push-front)
    uint32_t await_ref[GEN_MAX_AWAITS];     // event handle or guard index
    uint32_t hold_handle, guard_key;
    double   stamp;
    // resource-pool bookkeeping of the process body (model 4)
    uint32_t holds_pool;        // a cmi_process_holdable tag for the pool is on its resources list
    uint32_t holds_tool;        // ... for the binary cmb_resource (model 5)
    uint32_t held, req, rem, initially_held;
    // processes waiting for this one to finish (cmb_process::waiters), most recent first
    uint8_t  waiters[GEN_MAX_PROCS];
    uint32_t n_waiters;
    // body-local variables of the model-8 processes
    uint32_t timer, bell;
    int32_t  jobs, job;
};

struct GeneralState {
    EventHeap fel;
    GuardHeap guard[3];         // 0 = front (getters wait here), 1 = rear (putters), 2 = binary resource
    GenProc   proc[GEN_MAX_PROCS];
    double    ring[16];
    uint32_t  ring_cap, ring_head, ring_len;
    uint32_t  guard_seq;        // enqueue_seq, src/cmb_resourceguard.c:64
    uint32_t  status;
    // cmb_resourcepool (model 4): guard[0] is its guard
    HolderHeap holders;
    uint32_t  pool_cap, pool_in_use;
    // cmb_buffer (guards 0/1) and cmb_resource (guard 2), model 5
    uint32_t  buf_cap, buf_level;
    uint32_t  tool_holder;      // process index, or NO_HOLDER
    // cmb_priorityqueue (guards 0/1) and cmb_condition (guard 2), model 6
    PrioHeap  pq;
    uint32_t  pq_cap;
    uint32_t  last_handle[2];
    int32_t   level;
    // processes waiting for events (the waiters list riding in item[3] of an event's heap
    // tag, src/cmb_event.c:51-58), here a side table keyed by event handle, in registration order
    uint32_t  ew_key[GEN_EVENT_WAITERS];
    uint8_t   ew_pid[GEN_EVENT_WAITERS];
    uint32_t  n_ew;
    uint32_t  tool_observer;    // 1 + index of the guard registered as observer of guard[2], 0 = none
    uint32_t  bell;             // model 8: handle of the latest bell event
    uint32_t  clerk_start_pending;
};

constexpr uint32_t NO_HOLDER = 0xffffffffu;

struct GeneralSim {
    GeneralState *st;
    Sfc64 rng;
    double now;
    const ZigHot *hot;

    // ---- awaitables (src/cmb_process.c:220-260)
    __device__ __noinline__ void await_push(GenProc &p, uint32_t type, uint32_t ref)
    {
        if (p.n_awaits >= (uint32_t)GEN_MAX_AWAITS) {
            st->status |= TRIAL_ERR_PROC_OVERFLOW;
            return;
        }
        for (uint32_t k = p.n_awaits; k > 0u; k--) {
            p.await_type[k] = p.await_type[k - 1u];
            p.await_ref[k] = p.await_ref[k - 1u];
        }
        p.await_type[0] = type;
        p.await_ref[0] = ref;
        p.n_awaits++;
    }

    // cmi_process_remove_awaitable(pp, type, NULL): first entry of that type, any value
    __device__ __noinline__ bool await_remove_any(GenProc &p, uint32_t type)
    {
        for (uint32_t k = 0u; k < p.n_awaits; k++) {
            if (p.await_type[k] == type) {
                for (uint32_t m = k; m + 1u < p.n_awaits; m++) {
                    p.await_type[m] = p.await_type[m + 1u];
                    p.await_ref[m] = p.await_ref[m + 1u];
                }
                p.n_awaits--;
                return true;
            }
        }
        return false;
    }

    __device__ __noinline__ bool await_remove(GenProc &p, uint32_t type, uint32_t ref)
    {
        for (uint32_t k = 0u; k < p.n_awaits; k++) {
            if (p.await_type[k] == type && p.await_ref[k] == ref) {
                for (uint32_t m = k; m + 1u < p.n_awaits; m++) {
                    p.await_type[m] = p.await_type[m + 1u];
                    p.await_ref[m] = p.await_ref[m + 1u];
                }
                p.n_awaits--;
                return true;
            }
        }
        return false;
    }

    // ---- events
    __device__ uint32_t schedule(uint32_t act, uint32_t subj, int32_t arg, double t, int32_t prio)
    {
        const uint32_t key = st->fel.push(0u, t, prio, act, subj, arg);
        if (key == 0u) {
            st->status |= TRIAL_ERR_FEL_OVERFLOW;
        }
        return key;
    }

    // wake_event_waiters, src/cmb_event.c:200-221 (the list is push-front / pop-front)
    __device__ __noinline__ void wake_event_waiters(uint32_t key, int32_t sig)
    {
        for (uint32_t k = st->n_ew; k > 0u; k--) {
            if (st->ew_key[k - 1u] == key) {
                const uint32_t pid = st->ew_pid[k - 1u];
                schedule(ACT_WAKE_EVENT, pid, sig, now, st->proc[pid].prio);
                for (uint32_t m = k - 1u; m + 1u < st->n_ew; m++) {
                    st->ew_key[m] = st->ew_key[m + 1u];
                    st->ew_pid[m] = st->ew_pid[m + 1u];
                }
                st->n_ew--;
            }
        }
    }

    // cmi_event_remove_waiter, src/cmb_event.c:486-508: first match from the head
    __device__ void event_remove_waiter(uint32_t key, uint32_t pid)
    {
        for (uint32_t k = st->n_ew; k > 0u; k--) {
            if (st->ew_key[k - 1u] == key && st->ew_pid[k - 1u] == pid) {
                for (uint32_t m = k - 1u; m + 1u < st->n_ew; m++) {
                    st->ew_key[m] = st->ew_key[m + 1u];
                    st->ew_pid[m] = st->ew_pid[m + 1u];
                }
                st->n_ew--;
                return;
            }
        }
    }

    __device__ bool event_cancel(uint32_t handle)       // src/cmb_event.c:285-302
    {
        if (!st->fel.remove(handle)) {
            return false;
        }
        if (st->n_ew != 0u) {
            wake_event_waiters(handle, (int32_t)SIG_CANCELLED);
        }
        return true;
    }

    __device__ bool event_is_scheduled(uint32_t handle) const   // src/cmb_event.c:145-150
    {
        return st->fel.find(handle) != 0u;
    }

    __device__ bool event_reschedule(uint32_t handle, double t)         // :308-324
    {
        const uint32_t at = st->fel.find(handle);
        if (at == 0u) {
            return false;
        }
        st->fel.reprioritize(handle, t, st->fel.slot[at].prio);
        return true;
    }

    __device__ bool event_reprioritize(uint32_t handle, int32_t prio)   // :330-344
    {
        const uint32_t at = st->fel.find(handle);
        if (at == 0u) {
            return false;
        }
        st->fel.reprioritize(handle, st->fel.slot[at].d, prio);
        return true;
    }

    __device__ __noinline__ void cancel_events_of(uint32_t subj)     // cmb_event_pattern_cancel(ANY, subj, ANY)
    {
        uint32_t hit[GEN_FEL_CAP];
        uint32_t n = 0u;
        for (uint32_t k = 1u; k <= st->fel.count; k++) {        // first pass, array order
            if (st->fel.slot[k].subj == subj) {
                hit[n++] = st->fel.slot[k].key;
            }
        }
        for (uint32_t k = 0u; k < n; k++) {                     // second pass
            event_cancel(hit[k]);
        }
    }

    // ---- process layer
    __device__ __noinline__ void cancel_awaiteds(uint32_t pid)       // src/cmb_process.c:581-620
    {
        GenProc &p = st->proc[pid];
        while (p.n_awaits > 0u) {
            const uint32_t type = p.await_type[0];
            const uint32_t ref = p.await_ref[0];
            for (uint32_t m = 0u; m + 1u < p.n_awaits; m++) {
                p.await_type[m] = p.await_type[m + 1u];
                p.await_ref[m] = p.await_ref[m + 1u];
            }
            p.n_awaits--;
            if (type == AWAIT_TIME) {
                (void)event_cancel(ref);
            }
            else if (type == AWAIT_PROCESS) {                   // cmi_process_remove_waiter, :529-551
                GenProc &awaited = st->proc[ref];
                for (uint32_t k = 0u; k < awaited.n_waiters; k++) {
                    if (awaited.waiters[k] == pid) {
                        for (uint32_t m = k; m + 1u < awaited.n_waiters; m++) {
                            awaited.waiters[m] = awaited.waiters[m + 1u];
                        }
                        awaited.n_waiters--;
                        break;
                    }
                }
            }
            else if (type == AWAIT_EVENT) {
                event_remove_waiter(ref, pid);
            }
            // AWAIT_RESOURCE: cmb_resourceguard_remove(guard, process) searches for a key
            // equal to the process ADDRESS; entries are keyed by sequence number, so it
            // never finds one.  Nothing to do - the waiter removes its
            // own entry with the right key when it resumes (wait_end).
        }
        cancel_events_of(pid);
    }

    __device__ __noinline__ void hold_begin(uint32_t pid, double dur)        // :262-273, 316-333
    {
        GenProc &p = st->proc[pid];
        if (dur < 0.0) {
            st->status |= TRIAL_ERR_NEGATIVE_HOLD;
        }
        p.hold_handle = schedule(ACT_WAKE_TIME, pid, (int32_t)SIG_SUCCESS, __dadd_rn(now, dur), p.prio);
        await_push(p, AWAIT_TIME, p.hold_handle);
    }

    __device__ int32_t hold_end(uint32_t pid, int32_t sig)      // :274-284, 338-349
    {
        GenProc &p = st->proc[pid];
        if (sig != (int32_t)SIG_SUCCESS) {
            (void)await_remove(p, AWAIT_TIME, p.hold_handle);
            (void)event_cancel(p.hold_handle);
        }
        return sig;
    }

    __device__ __noinline__ void wait_begin(uint32_t g, uint32_t pid)        // src/cmb_resourceguard.c:125-152
    {
        GenProc &p = st->proc[pid];
        p.guard_key = ++st->guard_seq;
        if (st->guard[g].push(p.guard_key, now, p.prio, 0u, pid, 0) == 0u) {
            st->status |= TRIAL_ERR_GUARD_OVERFLOW;
        }
        await_push(p, AWAIT_RESOURCE, g);
    }

    __device__ int32_t wait_end(uint32_t g, uint32_t pid, int32_t sig)      // :153-162
    {
        GenProc &p = st->proc[pid];
        if (sig != (int32_t)SIG_SUCCESS) {
            (void)st->guard[g].remove(p.guard_key);
        }
        (void)await_remove(p, AWAIT_RESOURCE, g);
        return sig;
    }

    __device__ __noinline__ void signal(uint32_t g, bool demand_holds)       // :202-226
    {
        GuardHeap &h = st->guard[g];
        if (h.count > 0u && demand_holds) {
            const uint32_t pid = h.slot[1].subj;
            h.pop();
            schedule(ACT_WAKE_RESOURCE, pid, (int32_t)SIG_SUCCESS, now, st->proc[pid].prio);
        }
    }

    __device__ void interrupt(uint32_t pid, int32_t sig, int32_t pri)       // src/cmb_process.c:653-666
    {
        schedule(ACT_WAKE_INTERRUPT, pid, sig, now, pri);
    }

    __device__ __noinline__ void stop(uint32_t pid)                          // :698-723
    {
        GenProc &p = st->proc[pid];
        if (p.status != PROC_RUNNING) {
            return;
        }
        p.status = PROC_FINISHED;
        cancel_awaiteds(pid);
        if (p.holds_pool) {                                     // cmi_process_drop_resources, :507-527
            p.holds_pool = 0u;
            pool_drop_holder(pid);
        }
        if (p.holds_tool) {                                     // resource_drop_holder, src/cmb_resource.c:45-56
            p.holds_tool = 0u;
            st->tool_holder = NO_HOLDER;
            tool_signal();
        }
        wake_process_waiters(pid, (int32_t)SIG_STOPPED);
    }

    // cmb_resourceguard_signal on the binary resource's guard, forwarded to its observer
    // (src/cmb_resourceguard.c:202-242); both demands are "the resource has no holder"
    __device__ void tool_signal()
    {
        signal(2u, st->tool_holder == NO_HOLDER);
        if (st->tool_observer != 0u) {
            signal(st->tool_observer - 1u, st->tool_holder == NO_HOLDER);
        }
    }

    __device__ __noinline__ void wake_process_waiters(uint32_t pid, int32_t sig)         // src/cmb_process.c:485-505
    {
        GenProc &p = st->proc[pid];
        for (uint32_t k = 0u; k < p.n_waiters; k++) {
            const uint32_t q = p.waiters[k];
            schedule(ACT_WAKE_PROCESS, q, sig, now, st->proc[q].prio);
        }
        p.n_waiters = 0u;
    }

    // cmb_process_exit for a process that holds nothing, :671-684
    __device__ void exit(uint32_t pid)
    {
        cancel_awaiteds(pid);
        wake_process_waiters(pid, (int32_t)SIG_SUCCESS);
        st->proc[pid].status = PROC_FINISHED;
    }

    __device__ uint32_t timer_add(uint32_t pid, double dur, int32_t sig)    // :316-333
    {
        GenProc &p = st->proc[pid];
        const uint32_t h = schedule(ACT_WAKE_TIME, pid, sig, __dadd_rn(now, dur), p.prio);
        await_push(p, AWAIT_TIME, h);
        return h;
    }

    __device__ bool timer_cancel(uint32_t pid, uint32_t handle)             // :338-349
    {
        (void)await_remove(st->proc[pid], AWAIT_TIME, handle);
        return event_cancel(handle);
    }

    __device__ __noinline__ void timers_clear(uint32_t pid)                              // :354-381
    {
        GenProc &p = st->proc[pid];
        uint32_t k = 0u;
        while (k < p.n_awaits) {
            if (p.await_type[k] == AWAIT_TIME) {
                const uint32_t handle = p.await_ref[k];
                for (uint32_t m = k; m + 1u < p.n_awaits; m++) {
                    p.await_type[m] = p.await_type[m + 1u];
                    p.await_ref[m] = p.await_ref[m + 1u];
                }
                p.n_awaits--;
                (void)event_cancel(handle);
            }
            else {
                k++;
            }
        }
    }

    __device__ void wait_process_begin(uint32_t pid, uint32_t awaited)      // :428-452
    {
        GenProc &a = st->proc[awaited];
        await_push(st->proc[pid], AWAIT_PROCESS, awaited);
        for (uint32_t k = a.n_waiters; k > 0u; k--) {
            a.waiters[k] = a.waiters[k - 1u];
        }
        a.waiters[0] = (uint8_t)pid;
        a.n_waiters++;
    }

    __device__ void wait_event_begin(uint32_t pid, uint32_t handle)         // :461-483
    {
        if (st->n_ew >= (uint32_t)GEN_EVENT_WAITERS) {
            st->status |= TRIAL_ERR_PROC_OVERFLOW;
            return;
        }
        st->ew_key[st->n_ew] = handle;
        st->ew_pid[st->n_ew] = (uint8_t)pid;
        st->n_ew++;
        await_push(st->proc[pid], AWAIT_EVENT, handle);
    }

    __device__ void resume(uint32_t pid, int32_t sig)                       // :751-760
    {
        schedule(ACT_RESUME, pid, sig, now, st->proc[pid].prio);
    }

    // ---- cmb_resourcepool (src/cmb_resourcepool.c); guard index 0
    __device__ void pool_signal()
    {
        signal(0u, st->pool_cap - st->pool_in_use > 0u);        // is_available, :198-211
    }

    __device__ uint32_t pool_held_by(uint32_t pid) const        // :302-318
    {
        const uint32_t k = st->holders.find(pid + 1u);
        return k ? (uint32_t)st->holders.slot[k].arg : 0u;
    }

    __device__ __noinline__ void pool_update_record(uint32_t pid, uint32_t amount)       // :324-355
    {
        const uint32_t k = st->holders.find(pid + 1u);
        if (k != 0u) {
            st->holders.slot[k].arg += (int32_t)amount;
        }
        else {
            st->proc[pid].holds_pool = 1u;
            if (st->holders.push(pid + 1u, 0.0, st->proc[pid].prio, 0u, pid, (int32_t)amount) == 0u) {
                st->status |= TRIAL_ERR_PROC_OVERFLOW;
            }
        }
    }

    __device__ __noinline__ void pool_release(uint32_t pid, uint32_t amount)             // :561-605
    {
        const uint32_t k = st->holders.find(pid + 1u);
        if (k != 0u && (uint32_t)st->holders.slot[k].arg == amount) {
            st->holders.remove(pid + 1u);
            st->proc[pid].holds_pool = 0u;
        }
        else if (k != 0u) {
            st->holders.slot[k].arg -= (int32_t)amount;
        }
        st->pool_in_use -= amount;
        pool_signal();
    }

    __device__ __noinline__ void pool_drop_holder(uint32_t pid)                          // :98-121
    {
        const uint32_t k = st->holders.find(pid + 1u);
        if (k != 0u) {
            st->pool_in_use -= (uint32_t)st->holders.slot[k].arg;
            st->holders.remove(pid + 1u);
            pool_signal();
        }
    }

    // cmb_process_priority_set for the RUNNING process (its awaits list is empty),
    // src/cmb_process.c:150-198 -> reprioritize_holder, src/cmb_resourcepool.c:127-137
    __device__ void priority_set_self(uint32_t pid, int32_t pri)
    {
        st->proc[pid].prio = pri;
        if (st->proc[pid].holds_pool) {
            st->holders.reprioritize(pid + 1u, 0.0, pri);
        }
    }
};

}  // namespace cimba_b200
