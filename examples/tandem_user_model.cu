// examples/tandem_user_model.cu - a model library of one's own, start to finish:
//
//     examples/tandem_model.cuh            the model: three process bodies and a run_trial against cmb_device.cuh
//     python scripts/build_model.py examples/tandem_user_model.cu     -> cimba_b200/lib/models/libtandem_user_model.so
//     id = cimba_b200_model_load(".../libtandem_user_model.so");      then use id as cimba_b200_experiment.model
//
// tests/test_gpu_cmb_engine.py does exactly that and holds the result to the reference (ref_driver.c model 17).
#include "../cimba_b200/csrc/cmb_launch.cuh"
#include "tandem_model.cuh"

CMB_EXPORT_MODEL(tandem_example::Tandem, "tandem queue with a bounded buffer")
