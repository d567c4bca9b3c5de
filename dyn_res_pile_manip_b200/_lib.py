"""ctypes binding of libpilegnn.so (C ABI in include/pile_gnn.h).

There is deliberately NO fallback: if the shared library is missing or a call returns a
non-zero status the caller gets an exception.  Build with `python __graft_entry__.py` or
`make -C dyn_res_pile_manip_b200/csrc`.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("PILE_GNN_LIB") or os.path.join(CSRC, "libpilegnn.so")   # override: a prebuilt library

_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_LL = C.c_longlong
_D = C.c_double


class PusherStruct(C.Structure):
    """`pile_pusher` of include/pile_gnn.h (HOST struct describing the frame of the push)."""
    _fields_ = [("kind", C.c_int), ("cam_m12", C.c_float * 12), ("global_scale", C.c_float),
                ("s2r_scale", C.c_float), ("wkspc_center_x", C.c_float), ("wkspc_center_y", C.c_float)]


_PU = C.POINTER(PusherStruct)

# name -> (restype, argtypes); mirrors include/pile_gnn.h one to one
SIGNATURES = {
    "pile_abi_version": (_I, []),
    "pile_nf_effect": (_I, []),
    "pile_max_relations": (_I, []),
    "pile_error_string": (C.c_char_p, [_I]),
    "pile_set_tensor_cores": (_I, [_I]),
    "pile_get_tensor_cores": (_I, []),
    "pile_debug_set_trace": (_I, [_P, _I, _I]),
    "pile_wpack_num_slots": (_I, []),
    "pile_wpack_slot_offset": (_LL, [_I]),
    "pile_wpack_slot_size": (_LL, [_I]),
    "pile_wpack_total": (_LL, []),
    "pile_gen_s_delta": (_I, [_P, _P, _I, _PU, _I, _I, _P, _P]),
    "pile_gen_s_delta_backward": (_I, [_P, _P, _I, _PU, _I, _I, _P, _P, _P, _I, _P]),
    "pile_build_relations": (_I, [_P, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "pile_step_scratch_bytes": (_LL, [_I, _I]),
    "pile_tape_step_bytes": (_LL, [_I, _I]),
    "pile_predict_step": (_I, [_P, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P]),
    "pile_forward_relations": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "pile_relations_view": (_I, [_P, _I, _I, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "pile_rollout_forward": (_I, [_P, _P, _P, _P, _P, _PU, _F, _I, _I, _I, _P, _P, _P, _P]),
    "pile_profile_step": (_I, [_P, _P, _P, _P, _P, _I, _PU, _F, _I, _I, _P, _P, _I, _P, _P]),
    "pile_bwd_scratch_bytes": (_LL, [_I, _I]),
    "pile_step_backward": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "pile_rollout_backward": (_I, [_P, _P, _P, _P, _PU, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pile_reward": (_I, [_P, _LL, _LL, _I, _P, _I, _I, _P, _I, _P, _F, _F, _I, _P, _P, _P]),
    "pile_reward_backward": (_I, [_P, _LL, _LL, _I, _P, _I, _I, _P, _I, _P, _F, _F, _I, _P, _P, _P, _LL, _I, _P]),
    "pile_adam_clamp": (_I, [_P, _P, _P, _P, _LL, _I, _F, _F, _F, _F, _P, _P, _P]),
    "pile_adam_clamp_dev": (_I, [_P, _P, _P, _P, _LL, _P, _F, _F, _F, _F, _P, _P, _P]),
    "pile_counter_add": (_I, [_P, _I, _P]),
    "pile_gd_track": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "pile_fps": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pile_fps_sets": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "pile_depth_counts_len": (_I, [_I, _I]),
    "pile_depth_to_points": (_I, [_P, _I, _I, _P, _F, _P, _I, _P, _P, _P]),
    "pile_voxel_downsample_bytes": (_LL, [_I]),
    "pile_voxel_downsample": (_I, [_P, _I, _D, _P, _P, _P, _P]),
    "pile_cover_radius": (_I, [_P, _I, _P, _I, _I, _P, _P]),
    "pile_recenter": (_I, [_P, _I, _P, _I, _I, _P, _D, _D, _P, _P]),
    "pile_train_tape_bytes": (_LL, [_I, _I]),
    "pile_train_scratch_bytes": (_LL, [_I, _I]),
    "pile_train_grad_offset": (_LL, [_I]),
    "pile_train_forward": (_I, [_P, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P]),
    "pile_train_forward_relations": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "pile_train_backward": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pile_train_relations_view": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "pile_general_wpack_slot_offset": (_LL, [_I, _I]),
    "pile_general_tape_bytes": (_LL, [_I, _I, _I]),
    "pile_general_scratch_bytes": (_LL, [_I, _I, _I]),
    "pile_general_grad_offset": (_LL, [_I, _I]),
    "pile_general_forward": (_I, [_P, _I, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P]),
    "pile_general_forward_inference": (_I, [_P, _I, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P]),
    "pile_general_forward_relations": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "pile_general_backward": (_I, [_P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pile_general_relations_view": (_I, [_P, _I, _I, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "pile_rgr_param_offset": (_LL, [_I]),
    "pile_rgr_workspace_bytes": (_LL, [_I, _I, _I]),
    "pile_rgr_forward": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "pile_mppi_num_chunks": (_I, [_I]),
    "pile_mppi_partials": (_I, [_P, _P, _I, _I, _F, _P, _P]),
    "pile_mppi_combine": (_I, [_P, _I, _I, _P, _P]),
}

ABI_VERSION = 4          # PILE_ABI_VERSION of include/pile_gnn.h
_lib = None


class PileLibraryError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libpilegnn.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise PileLibraryError("building libpilegnn.so failed")
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise PileLibraryError(
            "%s is missing -- the CUDA path is the only path (no CPU fallback). "
            "Build it with `python __graft_entry__.py` or `make -C %s`." % (LIB_PATH, CSRC))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pile_abi_version() != ABI_VERSION:
        raise PileLibraryError("libpilegnn ABI version mismatch")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().pile_error_string(code)
        raise PileLibraryError("%s failed: CUDA error %d (%s)" % (what, code, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def host_floats(values):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr


def host_doubles(values):
    return (C.c_double * len(values))(*[float(v) for v in values])
