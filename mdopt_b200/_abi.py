"""ctypes mirror of ``include/mdopt_b200.h`` (struct layout and prototypes); no CUDA needed to import."""
from __future__ import annotations

import ctypes as C


class DecodeDesc(C.Structure):
    _fields_ = [
        ("n_sites", C.c_int32), ("n_logical", C.c_int32), ("chi_max", C.c_int32), ("mode", C.c_int32),
        ("renormalise", C.c_int32), ("trace_stride", C.c_int32), ("trace_max", C.c_int32), ("n_ops", C.c_int32),
        ("cut_sweep", C.c_double), ("cut_move", C.c_double), ("rank_tol", C.c_double), ("bias_p", C.c_double),
        ("n_head", C.c_int32), ("kept_mask", C.c_uint32),
        ("readout_rel", C.c_double), ("readout_abs", C.c_double),
    ]


ST_OK, ST_NOT_CONVERGED, ST_CAPACITY, ST_ZERO_NORM, ST_NOT_MONOMIAL = 0, 1, 2, 3, 4
E_BADARG, E_CAPACITY, E_NOGPU = -1, -2, -3
DEFAULT_RANK_TOL = 1e-14
CUT_MOVE = 1e-12  # default cut of split_two_site_tensor, used by move_orth_centre (mdopt/utils/utils.py:293)

_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_sz = C.c_size_t
_desc = C.POINTER(DecodeDesc)

# name -> (restype, argtypes); every symbol include/mdopt_b200.h declares
PROTOTYPES = {
    "mdopt_b200_abi_version": (_i, []),
    "mdopt_device_sm_count": (_i, []),
    "mdopt_chi_capacity": (_i, [_i]),
    "mdopt_decode_num_ctas": (_i, [_i]),
    "mdopt_decode_workspace_bytes": (_sz, [_i, _i, _i]),
    "mdopt_decode_batch_device": (_i, [_desc, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mdopt_decode_batch_host": (_i, [_desc, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                     _sz, _p]),
    "mdopt_mono_chi_capacity": (_i, [_i]),
    "mdopt_mono_num_warps": (_i, [_i]),
    "mdopt_mono_workspace_bytes": (_sz, [_i, _i, _i]),
    "mdopt_mono_decode_batch_device": (_i, [_desc, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mdopt_mono_decode_batch_host": (_i, [_desc, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                          _sz, _p]),
    "mdopt_site_stride": (_sz, [_i]),
    "mdopt_mps_program_workspace_bytes": (_sz, [_i, _i]),
    "mdopt_mps_program_device": (_i, [_desc, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mdopt_svd_batch_device": (_i, [_i, _i, _i, _i, _p, _i, _d, _i, _p, _p, _p, _p, _p, _p, _p]),
    "mdopt_site_apply_batch_device": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "mdopt_readout_batch_device": (_i, [_i, _i, _p, _i, _p, _p, _p]),
}


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def make_desc(n_sites, n_logical, chi_max, mode, renormalise, n_ops, cut_sweep, bias_p, trace_stride=0,
              trace_max=0, cut_move=CUT_MOVE, rank_tol=DEFAULT_RANK_TOL, kept_mask=0, readout_rel=1e-9,
              readout_abs=1e-12) -> DecodeDesc:
    kept_mask = int(kept_mask)
    n_head = kept_mask.bit_length() if kept_mask else 0
    return DecodeDesc(int(n_sites), int(n_logical), int(chi_max), int(mode), int(bool(renormalise)),
                      int(trace_stride), int(trace_max), int(n_ops), float(cut_sweep), float(cut_move),
                      float(rank_tol), float(bias_p), n_head, kept_mask, float(readout_rel), float(readout_abs))


CAPACITIES = (8, 16, 32, 64, 128, 256)  # <= 64: work matrices in shared memory; 128/256: global-memory tier
MAX_CAPACITY = CAPACITIES[-1]


def readout_min_capacity(n_logical: int) -> int:
    """Smallest capacity whose read-out pages hold the dense stage for ``n_logical`` kept sites."""
    k = int(n_logical)
    return (1 << (k - 1)) if k <= 6 else (1 << ((k + 1) // 2))


MONO_CAPACITIES = (8, 16, 32, 64, 128, 256, 512)  # the monomial warp engine


def mono_chi_capacity(chi: int) -> int:
    for cap in MONO_CAPACITIES:
        if chi <= cap:
            return cap
    return 0


def chi_capacity(chi: int) -> int:
    for cap in CAPACITIES:
        if chi <= cap:
            return cap
    return 0
