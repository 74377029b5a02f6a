"""ctypes bindings of include/ibmfoam_b200.h.

The library is built in-tree (ibmfoam_b200/lib/libibmfoam_b200.so).  Loading fails loudly when it is
missing; there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libibmfoam_b200.so")

IBM_OK = 0
ERRORS = {-1: "IBM_E_ARG", -2: "IBM_E_CUDA", -3: "IBM_E_STATE", -4: "IBM_E_UNSUPPORTED", -5: "IBM_E_COMM", -6: "IBM_E_CAPACITY"}

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class PatchDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("side", C.c_int), ("neighbPatch", C.c_int), ("nFaces", C.c_int),
                ("Cf", c_double_p), ("Sf", c_double_p), ("magSf", c_double_p)]


class MeshDesc(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("origin", C.c_double * 3), ("delta", C.c_double * 3),
                ("nBlocks", C.c_int), ("blocks", c_int_p), ("latticeToCell", c_int_p), ("nPatches", C.c_int),
                ("patches", C.POINTER(PatchDesc))]


class CloudCfg(C.Structure):
    _fields_ = [("moving", C.c_int), ("rotating", C.c_int), ("coupled", C.c_int), ("e", C.c_double),
                ("UInf", C.c_double * 3), ("rhoRef", C.c_double), ("pRef", C.c_double), ("g", C.c_double * 3),
                ("cyclicWithoutWalls", C.c_int), ("maxInstances", C.c_int)]


class ParticleInit(C.Structure):
    _fields_ = [("kind", C.c_int), ("position", C.c_double * 3), ("d", C.c_double), ("a", C.c_double),
                ("b", C.c_double), ("nTheta", C.c_int), ("nk", C.c_int), ("delta", C.c_double),
                ("theta", C.c_double * 3), ("v", C.c_double * 3), ("omega", C.c_double * 3), ("rho", C.c_double)]


class ParticleState(C.Structure):
    _fields_ = [("index", C.c_int), ("copy", C.c_int), ("active", C.c_int), ("onMesh", C.c_int),
                ("center", C.c_double * 3), ("theta", C.c_double * 3), ("v", C.c_double * 3),
                ("omega", C.c_double * 3), ("rho", C.c_double), ("age", C.c_double), ("F", C.c_double * 3),
                ("T", C.c_double * 3), ("nShell", C.c_int), ("N", C.c_int)]


class SolveStats(C.Structure):
    _fields_ = [("nSubSteps", C.c_int), ("nInstances", C.c_int), ("avgV", C.c_double * 3), ("avgOmega", C.c_double * 3),
                ("avgP", C.c_double * 3), ("totK", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("update_ms", C.c_double), ("forcing_ms", C.c_double), ("stress_ms", C.c_double), ("advance_ms", C.c_double),
                ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("layout_ms", C.c_double),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("n_kernel_launches", C.c_longlong),
                ("n_forcing_calls", C.c_longlong), ("n_updates", C.c_longlong)]


# every symbol include/ibmfoam_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "ibm_create": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(CloudCfg), C.c_int, C.POINTER(_VP)]),
    "ibm_set_particles": (C.c_int, [_VP, C.c_int, C.POINTER(ParticleInit)]),
    "ibm_destroy": (C.c_int, [_VP]),
    "ibm_last_error": (C.c_char_p, [_VP]),
    "ibm_abi_version": (C.c_int, []),
    "ibm_update": (C.c_int, [_VP]),
    "ibm_forcing": (C.c_int, [_VP, _VP, _VP, _VP, C.c_double, _VP]),
    "ibm_forcing_device": (C.c_int, [_VP, _VP, _VP, _VP, C.c_double, _VP]),
    "ibm_integrate_stress": (C.c_int, [_VP, _VP, _VP, _VP]),
    "ibm_integrate_stress_device": (C.c_int, [_VP, _VP, _VP, _VP]),
    "ibm_combine_forces": (C.c_int, [_VP]),
    "ibm_advance": (C.c_int, [_VP, C.c_double, C.POINTER(SolveStats)]),
    "ibm_solve": (C.c_int, [_VP, _VP, _VP, C.c_double, C.POINTER(SolveStats)]),
    "ibm_solve_device": (C.c_int, [_VP, _VP, _VP, C.c_double, C.POINTER(SolveStats)]),
    "ibm_substep": (C.c_int, [_VP, C.c_double]),
    "ibm_n_instances": (C.c_int, [_VP, c_int_p]),
    "ibm_get_particles": (C.c_int, [_VP, C.POINTER(ParticleState)]),
    "ibm_set_force_torque": (C.c_int, [_VP, _VP]),
    "ibm_cd": (C.c_int, [_VP, _VP]),
    "ibm_max_co_num": (C.c_int, [_VP, c_double_p]),
    "ibm_particle_mass": (C.c_int, [_VP, C.c_int, c_double_p, c_double_p]),
    "ibm_substep_schedule": (C.c_int, [C.c_double, _VP, C.c_int]),
    "ibm_get_shell": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ibm_get_points": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ibm_get_sf": (C.c_int, [_VP, C.c_int, _VP]),
    "ibm_get_indicator": (C.c_int, [_VP, _VP]),
    "ibm_get_solid_fraction": (C.c_int, [_VP, C.c_int, _VP]),
    "ibm_device_field_bytes": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_size_t)]),
    "ibm_upload_field": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "ibm_download_field": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "ibm_restore_particles": (C.c_int, [_VP, C.c_int, _VP]),
    "ibm_host_register": (C.c_int, [_VP, _VP, C.c_size_t]),
    "ibm_host_unregister": (C.c_int, [_VP, _VP]),
    "ibm_sync": (C.c_int, [_VP]),
    "ibm_set_kernel_variant": (C.c_int, [_VP, C.c_int, C.c_int]),
    "ibm_set_option": (C.c_int, [_VP, C.c_char_p, C.c_int]),
    "ibm_set_real_option": (C.c_int, [_VP, C.c_char_p, C.c_double]),
    "ibm_get_counter": (C.c_int, [_VP, C.c_char_p, C.POINTER(C.c_longlong)]),
    "ibm_timings": (C.c_int, [_VP, C.POINTER(Timing)]),
    "ibm_timings_reset": (C.c_int, [_VP]),
    "ibm_comm_unique_id_bytes": (C.c_int, []),
    "ibm_comm_get_unique_id": (C.c_int, [_VP]),
    "ibm_comm_init": (C.c_int, [_VP, _VP, C.c_int, C.c_int]),
    "ibm_lattice_from_arrays": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ibm_set_slab": (C.c_int, [_VP, C.c_int, C.c_int]),
    "ibm_set_slab_halo": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ibm_comm_info": (C.c_int, [_VP, c_int_p, c_int_p]),
    "ibm_comm_local_group_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "ibm_comm_local_group_destroy": (C.c_int, [_VP]),
    "ibm_comm_init_local": (C.c_int, [_VP, _VP, C.c_int]),
}

_lib = None


class IbmError(RuntimeError):
    pass


def load():
    """Load libibmfoam_b200.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IbmError(f"{LIB_PATH} is missing: build it with `python -m ibmfoam_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback for the immersed-boundary path.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int, ctx=None):
    if rc == IBM_OK:
        return
    msg = load().ibm_last_error(ctx)
    raise IbmError(f"{ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")
