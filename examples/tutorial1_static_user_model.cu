// examples/tutorial1_static_user_model.cu - the reference's first tutorial (cimba_b200/models/tutorial1_model.cuh) as a user-built
// library on the static tier WITH events of the model's own: two processes, no object queue (the tutorial queues in a cmb_buffer),
// three event slots (start recording, stop recording, end of simulation at priority -100).
//
//     python scripts/build_model.py examples/tutorial1_static_user_model.cu
#include "../cimba_b200/csrc/cmb_launch.cuh"
#include "../cimba_b200/models/tutorial1_model.cuh"

CMB_EXPORT_STATIC_MODEL_EVENTS(cimba_b200::models::Tutorial1T, 2, 0, 3, "tutorial 1 (tut_1_7.c), static tier with three events of its own")
