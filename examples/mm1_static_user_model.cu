// examples/mm1_static_user_model.cu - benchmark/MM1_multi.c (cimba_b200/models/mm1_model.cuh, 50 lines of model) as a user-built
// library on the static tier: python scripts/build_model.py examples/mm1_static_user_model.cu
#include "../cimba_b200/csrc/cmb_launch.cuh"
#include "../cimba_b200/models/mm1_model.cuh"

CMB_EXPORT_STATIC_MODEL(cimba_b200::models::MM1T, 2, 1, "M/M/1 written against cmb_device.cuh, static tier")
