/* oracle/ref_build/awacs_stubs/hdf5.h - TEST INFRASTRUCTURE ONLY.
 *
 * tutorial/tut_5_1.c (the AWACS model, BASELINE config 5) includes "hdf5.h" and writes ParaView
 * files from inside the simulation; HDF5 is not installed here.  This header lets the UNMODIFIED
 * source compile where it lies: every H5* call becomes a constant (the model never reads anything
 * back from HDF5), and - because tut_5_1.c includes this file after <stdio.h> and <cimba.h> - it is
 * also the place to silence the console output, to keep racetrack.vtp out of the working directory,
 * to count future-event-list pops and to remember where run_trial put its targets (see awacs_driver.c).
 */
#ifndef AWACS_STUB_HDF5_H
#define AWACS_STUB_HDF5_H

typedef long long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;

#define H5P_DEFAULT 0
#define H5S_ALL 0
#define H5S_SELECT_SET 0
#define H5S_UNLIMITED ((hsize_t)-1)
#define H5S_SCALAR 0
#define H5P_DATASET_CREATE 0
#define H5P_FILE_ACCESS 0
#define H5F_LIBVER_LATEST 0
#define H5F_ACC_TRUNC 0
#define H5F_SCOPE_GLOBAL 0
#define H5T_STR_NULLPAD 0
#define H5T_C_S1 0
#define H5T_NATIVE_INT32 0
#define H5T_NATIVE_INT64 0
#define H5T_NATIVE_UINT64 0
#define H5T_NATIVE_FLOAT 0
#define H5T_NATIVE_DOUBLE 0

#define H5Fcreate(...) ((hid_t)1)
#define H5Fclose(...) 0
#define H5Fflush(...) 0
#define H5Gcreate2(...) ((hid_t)1)
#define H5Gopen2(...) ((hid_t)1)
#define H5Gclose(...) 0
#define H5Pcreate(...) ((hid_t)1)
#define H5Pclose(...) 0
#define H5Pset_chunk(...) 0
#define H5Pset_libver_bounds(...) 0
#define H5Screate(...) ((hid_t)1)
#define H5Screate_simple(...) ((hid_t)1)
#define H5Sclose(...) 0
#define H5Sselect_hyperslab(...) 0
#define H5Dcreate2(...) ((hid_t)1)
#define H5Dclose(...) 0
#define H5Dwrite(...) 0
#define H5Dset_extent(...) 0
#define H5Dget_space(...) ((hid_t)1)
#define H5Acreate2(...) ((hid_t)1)
#define H5Awrite(...) 0
#define H5Aclose(...) 0
#define H5Tcopy(...) ((hid_t)1)
#define H5Tset_size(...) 0
#define H5Tset_strpad(...) 0
#define H5Tclose(...) 0

/* hooks of awacs_driver.c */
extern void awacs_counting_execute(void);
extern void *awacs_noting_calloc(unsigned long long n, unsigned long long sz);
#define cmb_event_queue_execute awacs_counting_execute
#define cmi_calloc(n, sz) awacs_noting_calloc((n), (sz))
#define printf(...) ((void)0)
#define fopen(name, mode) fopen("/dev/null", (mode))
#define main awacs_reference_main

#endif
