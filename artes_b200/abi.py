"""ctypes mirror of include/artes_gpu.h (struct layouts and constants)."""
import ctypes as C

ABI_VERSION = 3
MODE_FAITHFUL = 0
MODE_FAST = 1
ERR_SLOTS = 64
NCCL_ID_BYTES = 128


class Launch(C.Structure):
    """artes_launch_t: the scalar program-scope inputs of radiative_transfer (src/ARTES.f90:19-55)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("mode", C.c_int32),
        ("n_photons", C.c_uint64), ("photon_id_base", C.c_uint64), ("seed", C.c_uint64),
        ("photon_source", C.c_int32), ("photon_scattering", C.c_int32),
        ("photon_emission", C.c_int32), ("stellar_direction", C.c_int32),
        ("limb_emission", C.c_int32), ("flow_global", C.c_int32), ("flow_theta", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("wl_index", C.c_int32),
        ("fstop", C.c_double), ("photon_minimum", C.c_double), ("photon_bias", C.c_double),
        ("surface_albedo", C.c_double), ("theta_star", C.c_double), ("phi_star", C.c_double),
        ("det_theta", C.c_double), ("det_phi", C.c_double), ("x_max", C.c_double), ("y_max", C.c_double),
    ]


class Stats(C.Structure):
    """artes_stats_t"""
    _fields_ = [
        ("n_emit", C.c_uint64), ("n_cell_face", C.c_uint64), ("n_scatter", C.c_uint64),
        ("n_peel", C.c_uint64), ("n_surface", C.c_uint64), ("n_draws", C.c_uint64),
        ("n_error", C.c_uint64), ("reserved", C.c_uint64),
        ("kernel_ms", C.c_double), ("reduce_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def make_launch(**kw):
    """Launch with the defaults of `initialize` (src/ARTES.f90:283-314) for the fields it covers."""
    import math
    d = dict(mode=MODE_FAITHFUL, n_photons=100000, photon_id_base=0, seed=1, photon_source=1,
             photon_scattering=1, photon_emission=1, stellar_direction=0, limb_emission=0,
             flow_global=0, flow_theta=0, nx=25, ny=25, fstop=1e-5, photon_minimum=1e-20,
             photon_bias=0.8, surface_albedo=0.0, theta_star=math.pi / 2.0, phi_star=0.0,
             det_theta=math.pi / 2.0, det_phi=math.pi / 2.0, x_max=0.0, y_max=0.0)
    d.update(kw)
    L = Launch()
    L.struct_size = C.sizeof(Launch)
    for k, v in d.items():
        setattr(L, k, v)
    return L
