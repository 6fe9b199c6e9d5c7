// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Builds the reference's own CPython extension `MolEmb` (C_API/MolEmb.cpp, C_API/SH.hpp)
// *from the sources where they lie* under $TM_REFERENCE (default /root/reference) by
// #including them; nothing of the reference is copied into this repository.
//
// The only thing this shim adds is NumPy-2 compatibility for one OUT-OF-SCOPE routine
// (Overlap_RBFS, C_API/MolEmb.cpp:1705-1760) which passes a PyObject* to PyArray_DIM /
// PyArray_DATA and uses the removed NPY_IN_ARRAY flag.  The hot-path routines used as
// oracles (Make_NListNaive :1180-1247, Make_ANI1_Sym :1913-1988, Make_ANI1_Sym_deri
// :1844-1911) are compiled unmodified.
#include <Python.h>
#include <numpy/arrayobject.h>
#ifndef NPY_IN_ARRAY
#define NPY_IN_ARRAY NPY_ARRAY_IN_ARRAY
#endif
static inline npy_intp tm_shim_dim(const void* a, int i) { return PyArray_DIM((const PyArrayObject*)a, i); }
static inline void* tm_shim_data(const void* a) { return PyArray_DATA((PyArrayObject*)a); }
#define PyArray_DIM(a, i) tm_shim_dim((const void*)(a), (i))
#define PyArray_DATA(a) tm_shim_data((const void*)(a))
#include TM_REFERENCE_MOLEMB
