// examples/mm1_user_model.cu - a model library of one's own: benchmark/MM1_multi.c written against the device authoring
// surface (cimba_b200/models/mm1_model.cuh is the model: two process bodies and a run_trial, about fifty lines), exported so
// that the C-ABI library can load it:
//
//     python scripts/build_model.py examples/mm1_user_model.cu            -> cimba_b200/lib/models/libmm1_user_model.so
//     id = cimba_b200_model_load("cimba_b200/lib/models/libmm1_user_model.so");
//     ... cimba_b200_experiment.model = id; cimba_b200_run_experiment(array, n, stride, &desc);
#include "../cimba_b200/csrc/cmb_launch.cuh"
#include "../cimba_b200/models/mm1_model.cuh"

CMB_EXPORT_MODEL(cimba_b200::models::MM1, "mm1 (user build)")
