// examples/tandem_static_user_model.cu - the tandem model (tandem_model.cuh) exported on the STATIC tier: its three processes and
// two queues are all the model ever has, so the event list and the process records fit in registers and the queues in shared
// memory (cimba_b200/csrc/cmb_static.cuh).  Same model text as tandem_user_model.cu, one different line here; a trial that
// needs more than the tier holds (a queue beyond 32 + queue_spill_cap entries) is re-run by the general engine inside the launch.
//
//     python scripts/build_model.py examples/tandem_static_user_model.cu
#include "../cimba_b200/csrc/cmb_launch.cuh"
#include "tandem_model.cuh"

CMB_EXPORT_STATIC_MODEL(tandem_example::TandemT, 3, 2, "tandem queue with a bounded buffer (static tier)")
