// engine.cuh - device-side simulation kernel pieces shared by all models.
//
// What the reference does per event:
//   cmb_event_execute_next (src/cmb_event.c:229-252) pops the future-event list
//   (cmi_hashheap_dequeue, src/cmi_hashheap.c:486-524), advances the clock and
//   calls the event's action, which resumes a stackful coroutine
//   (src/cmi_coroutine.c:296) that runs user code until its next blocking call
//   (cmb_process_hold -> cmb_event_schedule -> cmi_hashheap_enqueue) and yields.
//
// What this engine does instead (B200-first, not a translation):
//   * a process is a resume-point index; a blocking call is a *command* the
//     process body hands back to the dispatcher (hold for a duration, hold for a
//     random variate, wait on a guard, exit).  The dispatcher executes commands
//     in warp-converged code, so lanes that are in different processes still
//     share the expensive part (variate generation, event-list insert);
//   * event records are 16-24 bytes, not 64 (time, issue key, action, owner);
//   * ordering is the reference's strict total order - time ascending, priority
//     descending, issue key ascending (default_compare, src/cmi_hashheap.c:55-80)
//     - and keys are issued 1,2,3... per trial in program order
//     (src/cmi_hashheap.c:449-453), so pop order is bit-identical whatever the
//     container's internal layout ("Event ordering").
#pragma once

#include <cstdint>
#ifndef CMB_HOST_BUILD     // tests/cmb_engine_host.cpp compiles this text for the CPU
#include <cuda_runtime.h>
#endif

namespace cimba_b200 {

// Signals a blocking call returns (include/cmb_process.h:59-99) - same values.
enum : int64_t {
    SIG_SUCCESS = 0,
    SIG_PREEMPTED = -1,
    SIG_INTERRUPTED = -2,
    SIG_STOPPED = -3,
    SIG_CANCELLED = -4,
    SIG_TIMEOUT = -5,
};

// Event actions = the reference's event functions.
enum : uint32_t {
    ACT_NONE = 0u,
    ACT_START = 1u,            // start_event,            src/cmb_process.c:115-122
    ACT_WAKE_TIME = 2u,        // wakeup_event_time,      src/cmb_process.c:292-308
    ACT_WAKE_RESOURCE = 3u,    // wakeup_event_resource,  src/cmb_resourceguard.c:168-180
};

// Per-trial status word written back with the results (SURVEY.md section 8b,
// "Error convention": the host turns non-zero into the reference's fatal path).
enum : uint32_t {
    TRIAL_OK = 0u,
    TRIAL_ERR_QUEUE_OVERFLOW = 1u,     // object queue outgrew window + spill ring
    TRIAL_ERR_FEL_OVERFLOW = 2u,       // more pending events than the container holds
    TRIAL_ERR_KEY_OVERFLOW = 4u,       // more than 2^32-1 events issued in one trial
    TRIAL_ERR_GUARD_OVERFLOW = 8u,     // more waiters than the guard holds
    TRIAL_ERR_PROC_OVERFLOW = 16u,     // more live processes than the pool holds
    TRIAL_ERR_NEGATIVE_HOLD = 32u,     // cmb_process_hold(dur < 0), src/cmb_process.c:264
};

// ---------------------------------------------------------------------------
// SlotFel<N>: future-event list for models in which every process owns at most
// one pending event (start, hold wake-up or resource wake-up) - true for any
// model without timers/interrupts.  Slot p belongs to process p, so an insert
// is a register write and pop-min is an N-way compare; the whole list lives in
// registers.  Equal priorities only (all shipped workloads use priority 0).
// ---------------------------------------------------------------------------
template <int N>
struct SlotFel {
    double   t[N];
    uint32_t key[N];
    uint32_t act[N];
    uint32_t issued;           // item_counter, src/cmi_hashheap.c:449-453

    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < N; i++) {
            t[i] = __longlong_as_double(0x7ff0000000000000LL);   // +inf = empty
            key[i] = 0u;
            act[i] = ACT_NONE;
        }
        issued = 0u;
    }

    // cmb_event_schedule (src/cmb_event.c:123-140) for owner process p.
    // Returns false if p already has a pending event (capacity violation).
    __device__ __forceinline__ bool schedule(int p, uint32_t action, double time)
    {
        const uint32_t k = ++issued;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i == p) {
                ok = (act[i] == ACT_NONE);
                t[i] = time;
                key[i] = k;
                act[i] = action;
            }
        }
        return ok;
    }

    __device__ __forceinline__ int count() const
    {
        int n = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            n += (act[i] != ACT_NONE) ? 1 : 0;
        }
        return n;
    }

    // cmi_hashheap_dequeue (src/cmi_hashheap.c:486-524): remove the entry that is
    // first under (time asc, key asc).  Returns false when the list is empty.
    __device__ __forceinline__ bool pop(int &p, uint32_t &action, double &time, uint32_t &k)
    {
        int best = 0;
        double bt = t[0];
        uint32_t bk = key[0];
#pragma unroll
        for (int i = 1; i < N; i++) {
            const bool before = (t[i] < bt) || (t[i] == bt && key[i] < bk);
            if (before) {
                best = i;
                bt = t[i];
                bk = key[i];
            }
        }
        uint32_t ba = ACT_NONE;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i == best) {
                ba = act[i];
                act[i] = ACT_NONE;
                t[i] = __longlong_as_double(0x7ff0000000000000LL);
            }
        }
        p = best;
        action = ba;
        time = bt;
        k = bk;
        return ba != ACT_NONE;
    }
};

// ---------------------------------------------------------------------------
// StampRing: the cmb_objectqueue of the queueing workloads.  The reference
// keeps a linked list of 16-byte tags pointing at pooled 8-byte objects
// (src/cmb_objectqueue.c:45-52, benchmark/MM1_multi.c:61-64); the only payload
// is the arrival timestamp, so the queue is a FIFO ring of doubles.
//
// The oldest WINDOW entries live in shared memory (one column per thread,
// element i of thread tid at win[i * stride + tid]: with 8-byte elements and a
// stride that is a multiple of 16 the 32 lanes of a warp hit 32 distinct bank
// pairs whatever their row, so accesses are conflict-free); anything beyond
// spills to a per-trial ring in HBM and is pulled back one element per get.
// ---------------------------------------------------------------------------
template <int WINDOW>
struct StampRing {
    static_assert((WINDOW & (WINDOW - 1)) == 0, "window must be a power of two");
    double  *win;              // shared memory, already offset by threadIdx.x
    double  *spill;            // global memory, this trial's ring (spill_cap doubles)
    uint32_t stride;           // blockDim.x
    uint32_t spill_mask;       // spill_cap - 1 (spill_cap is a power of two, may be 0)
    uint32_t head;             // index of the oldest entry (monotone)
    uint32_t len;

    __device__ __forceinline__ void init(double *w, uint32_t s, double *sp, uint32_t spill_cap)
    {
        win = w;
        stride = s;
        spill = sp;
        spill_mask = spill_cap - 1u;
        head = 0u;
        len = 0u;
    }

    // false = overflow (entry dropped, trial must be flagged)
    __device__ __forceinline__ bool put(double v)
    {
        const uint32_t pos = head + len;
        if (len < (uint32_t)WINDOW) {
            win[(pos & (WINDOW - 1)) * stride] = v;
        }
        else {
            if (spill == nullptr || len - WINDOW > spill_mask) {
                return false;
            }
            spill[pos & spill_mask] = v;
        }
        len++;
        return true;
    }

    __device__ __forceinline__ double take()
    {
        const double v = win[(head & (WINDOW - 1)) * stride];
        head++;
        len--;
        if (len >= (uint32_t)WINDOW) {
            const uint32_t pos = head + WINDOW - 1u;
            win[(pos & (WINDOW - 1)) * stride] = spill[pos & spill_mask];
        }
        return v;
    }
};

}  // namespace cimba_b200

#ifndef CMB_HOST_BUILD     // shared-memory containers with inline PTX: device only
namespace cimba_b200 {

// ---------------------------------------------------------------------------
// EventList<CAP>: general future-event list for lane-per-trial models whose
// processes come and go (any number of pending events per process, up to CAP in
// total).  The reference keeps a binary heap of 64-byte tags
// (src/cmi_hashheap.c); with the 5-12 live entries of the pooled-server models
// an unsorted array with a full scan per pop is cheaper on a SIMT machine: the
// scan is a fixed instruction sequence that all lanes of a warp execute
// together (sift loops would diverge per lane), inserts are O(1), and removal
// moves the last entry into the hole.  Pop order is the reference's total order
// (time asc, key asc at equal priority), which is all that parity needs.
//
// Layout: shared memory, one column per thread; entry i of thread tid is the
// 16-byte record head[i * stride + tid] = {time, (key << 2 | action), tag} plus
// an 8-byte payload pay[i * stride + tid] (the process-local state that travels
// with the continuation, e.g. a customer's arrival time).  16-byte accesses at a
// 16-byte lane pitch are conflict-free whatever row each lane addresses.
// ---------------------------------------------------------------------------
#ifndef EVENTLIST_BATCH
#define EVENTLIST_BATCH 2      // measured on B200 (M/M/c, c = 8): 2 -> 2.97e10, 4 -> 2.90e10, 8 -> 2.74e10 events/s at 32 768 trials
#endif

struct __align__(16) EventHead {
    double   time;
    uint32_t keyact;           // (issue key << 2) | action ; 0 = none
    uint32_t tag;              // model-defined (process kind / id)
};

template <int CAP>
struct EventList {
    uint32_t   head;           // shared-window byte address of this thread's column of EventHead records
    uint32_t   pay;            // ... of its column of payloads
    uint32_t   hstride;        // bytes between consecutive rows of head (blockDim.x * 16)
    uint32_t   pstride;        // ... of pay (blockDim.x * 8)
    uint32_t   count;
    uint32_t   issued;         // item_counter, src/cmi_hashheap.c:449-453

    __device__ __forceinline__ void init(EventHead *h, double *p, uint32_t threads)
    {
        head = (uint32_t)__cvta_generic_to_shared(h);
        pay = (uint32_t)__cvta_generic_to_shared(p);
        hstride = threads * 16u;
        pstride = threads * 8u;
        count = 0u;
        issued = 0u;
    }

    static __device__ __forceinline__ void ld_head(uint32_t addr, double &time, uint32_t &keyact, uint32_t &tag)
    {
        uint32_t lo, hi;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(hi), "=r"(keyact), "=r"(tag) : "r"(addr) : "memory");
        time = __hiloint2double((int)hi, (int)lo);
    }

    static __device__ __forceinline__ void st_head(uint32_t addr, double time, uint32_t keyact, uint32_t tag)
    {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                     :: "r"(addr), "r"((uint32_t)__double2loint(time)), "r"((uint32_t)__double2hiint(time)),
                        "r"(keyact), "r"(tag) : "memory");
    }

    // cmb_event_schedule: false if the list is full (entry dropped; flag the trial)
    __device__ __forceinline__ bool schedule(uint32_t action, uint32_t tag, double time, double payload)
    {
        const uint32_t k = ++issued;
        if (count >= (uint32_t)CAP) {
            return false;
        }
        st_head(head + count * hstride, time, (k << 2) | action, tag);
        asm volatile("st.shared.f64 [%0], %1;" :: "r"(pay + count * pstride), "d"(payload) : "memory");
        count++;
        return true;
    }

    // Pop the first entry under (time asc, key asc).  `scan` = number of rows to
    // look at; pass the warp-wide maximum of `count` so the loop trip count is
    // uniform.  Returns false if this lane's list is empty.
    __device__ __forceinline__ bool pop(uint32_t scan, EventHead &out, double &payload)
    {
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        double bt = INF;
        uint32_t bk = 0xffffffffu, bi = 0u, btag = 0u;
        // EVENTLIST_BATCH rows at a time: their loads are independent (one shared-memory latency, not
        // one per row) and the comparisons form a tournament, so the dependent chain of a scan over n
        // rows is n / EVENTLIST_BATCH steps.  Rows beyond `count` enter as (+inf, key ~0) and never
        // win.  The order is a strict total order (keys are unique), so the winner does not depend on
        // the bracket.
        for (uint32_t base = 0u; base < scan; base += EVENTLIST_BATCH) {
            double et[EVENTLIST_BATCH];
            uint32_t ek[EVENTLIST_BATCH], etag[EVENTLIST_BATCH], ei[EVENTLIST_BATCH];
#pragma unroll
            for (int j = 0; j < EVENTLIST_BATCH; j++) {
                et[j] = INF;
                ek[j] = 0xffffffffu;
                etag[j] = 0u;
                ei[j] = base + j;
                if (base + j < count) {
                    ld_head(head + (base + j) * hstride, et[j], ek[j], etag[j]);
                }
            }
#pragma unroll
            for (int width = EVENTLIST_BATCH / 2; width >= 1; width /= 2) {
#pragma unroll
                for (int j = 0; j < width; j++) {
                    const bool up = (et[j + width] < et[j]) | ((et[j + width] == et[j]) & (ek[j + width] < ek[j]));
                    et[j] = up ? et[j + width] : et[j];
                    ek[j] = up ? ek[j + width] : ek[j];
                    etag[j] = up ? etag[j + width] : etag[j];
                    ei[j] = up ? ei[j + width] : ei[j];
                }
            }
            const bool before = (et[0] < bt) | ((et[0] == bt) & (ek[0] < bk));
            bt = before ? et[0] : bt;
            bk = before ? ek[0] : bk;
            btag = before ? etag[0] : btag;
            bi = before ? ei[0] : bi;
        }
        if (count == 0u) {
            return false;
        }
        out.time = bt;
        out.keyact = bk;
        out.tag = btag;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(payload) : "r"(pay + bi * pstride) : "memory");
        count--;
        if (bi != count) {                              // move the last entry into the hole
            double lt, lp;
            uint32_t lk, ltag;
            ld_head(head + count * hstride, lt, lk, ltag);
            st_head(head + bi * hstride, lt, lk, ltag);
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(lp) : "r"(pay + count * pstride) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" :: "r"(pay + bi * pstride), "d"(lp) : "memory");
        }
        return true;
    }
};

}  // namespace cimba_b200
#endif  // CMB_HOST_BUILD
