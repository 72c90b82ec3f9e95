# cython: language_level=3
# core.pyx -- the reference-side native binding (INTEGRATION.md section 2).
#
# Same name, same typed-memoryview signature and the same argument meaning as the reference's native entry
# `maximum_path_c` (monotonic_align/core.pyx:38); the body is one call into libvits_mas.so
# (include/vits_mas.h: mas_maximum_path_c_host), so the reference's own wrapper
# (monotonic_align/__init__.py:7-20) runs unmodified on top of the sm_100a kernels.
# Built by tools/install_reference.py into baseline/_ref/binding/ and exercised by
# tests/test_binding.py under the reference's unmodified __init__.py.
cimport cython

cdef extern from "vits_mas.h":
    int mas_maximum_path_c_host(int* paths, const float* values, const int* t_ys, const int* t_xs,
                                int B, int T_y, int T_x) nogil
    const char* mas_error_string(int code) nogil


@cython.boundscheck(False)
@cython.wraparound(False)
cpdef void maximum_path_c(int[:, :, ::1] paths, float[:, :, ::1] values, int[::1] t_ys, int[::1] t_xs) nogil:
    cdef int rc
    if paths.shape[0] == 0 or paths.shape[1] == 0 or paths.shape[2] == 0:
        return
    rc = mas_maximum_path_c_host(&paths[0, 0, 0], &values[0, 0, 0], &t_ys[0], &t_xs[0],
                                 <int>paths.shape[0], <int>paths.shape[1], <int>paths.shape[2])
    # rc > 0 with a zero low byte = MAS_STATUS_* bits << 8: lengths the reference leaves undefined
    # (t_x > t_y, empty); those utterances keep their all-zero path, like an untouched np.zeros.
    if rc != 0 and not (rc > 0 and (rc & 0xFF) == 0):
        with gil:
            raise RuntimeError("mas_maximum_path_c_host failed: %s (%d)" % (mas_error_string(rc).decode(), rc))
