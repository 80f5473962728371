"""ctypes binding of ``include/caviar_b200.h``.  No CPU fallback: if the library is missing, importing
this module raises; if there is no B200, plan creation raises."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH


class CaviarError(RuntimeError):
    """A C-ABI call returned a negative status; the message is caviar_last_error()."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"caviar_b200 error {code}: {msg}")
        self.code = code


CV_OK, CV_E_INVALID, CV_E_CUDA, CV_E_NO_DEVICE, CV_E_UNSUPPORTED, CV_E_NOMEM = 0, -1, -2, -3, -4, -5
BOX_NONE, BOX_ORTHO, BOX_TRICLINIC = 0, 1, 2
PLAN_FORCE_DIRECT, PLAN_FORCE_STAGED = 1, 2
MODE_DIRECT, MODE_ZEROCOPY, MODE_STAGED, MODE_ALLPAIRS = 0, 1, 2, 3
ORDER_CALLER, ORDER_PLAN = 0, 1
MODE_NAMES = {MODE_DIRECT: "direct", MODE_ZEROCOPY: "zerocopy", MODE_STAGED: "staged", MODE_ALLPAIRS: "allpairs"}


class CvPlanDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_atoms", C.c_int32),
        ("n_pairs", C.c_int64), ("pairs", C.c_void_p),
        ("n_triplets", C.c_int64), ("triplets", C.c_void_p),
        ("box_kind", C.c_int32), ("flags", C.c_int32),
        ("pair_cutoff", C.c_double), ("hbond_dist_cutoff", C.c_double), ("hbond_angle_cutoff", C.c_double),
        ("device", C.c_int32), ("n_ctas", C.c_int32),
    ]


class CvPlanInfo(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("n_tiles", C.c_int32), ("n_ctas", C.c_int32), ("threads_per_cta", C.c_int32),
        ("stages", C.c_int32), ("frames_per_stage", C.c_int32), ("smem_bytes", C.c_int32),
        ("partial_slots", C.c_int32),
        ("n_features", C.c_int64), ("n_flag_words", C.c_int64), ("n_distinct_atoms", C.c_int64),
        ("staged_floats_per_frame", C.c_int64), ("boxrec_floats_per_frame", C.c_int64),
        ("workspace_bytes", C.c_int64), ("algorithmic_bytes_per_frame", C.c_int64),
        ("staged_fixedpoint_floats", C.c_int64), ("gather_bank_conflicts", C.c_int64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["mode_name"] = MODE_NAMES.get(self.mode, "?")
        return d


class CvFrameOutputs(C.Structure):
    _fields_ = [("dist", C.c_void_p), ("angle", C.c_void_p), ("flags", C.c_void_p), ("score", C.c_void_p)]


class CvAggregates(C.Structure):
    _fields_ = [("occupancy", C.c_void_p), ("sum", C.c_void_p), ("sumsq", C.c_void_p), ("occupancy_f64", C.c_void_p)]


class CvHostOutputs(C.Structure):
    _fields_ = [("dist", C.c_void_p), ("angle", C.c_void_p), ("flags", C.c_void_p), ("score", C.c_void_p),
                ("occupancy", C.c_void_p), ("sum", C.c_void_p), ("sumsq", C.c_void_p),
                ("weights_q32", C.c_void_p), ("score_q32", C.c_void_p)]


# every symbol include/caviar_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "caviar_plan_describe": (C.c_int, [C.POINTER(CvPlanDesc), C.c_int32, C.POINTER(CvPlanInfo)]),
    "caviar_plan_chunks": (C.c_int, [C.POINTER(CvPlanDesc), C.c_int32, C.c_int64, C.c_void_p]),
    "caviar_plan_layout": (C.c_int, [C.POINTER(CvPlanDesc), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "caviar_plan_allpairs_check": (C.c_int64, [C.POINTER(CvPlanDesc), C.c_int32, C.c_void_p]),
    "caviar_plan_create": (C.c_int, [C.POINTER(CvPlanDesc), C.POINTER(C.c_void_p)]),
    "caviar_plan_destroy": (None, [C.c_void_p]),
    "caviar_plan_info": (C.c_int, [C.c_void_p, C.POINTER(CvPlanInfo)]),
    "caviar_prepare_box": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "caviar_plan_atom_order": (C.c_int64, [C.POINTER(CvPlanDesc), C.c_int32, C.c_void_p, C.c_int64]),
    "caviar_stage_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "caviar_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                             C.POINTER(CvFrameOutputs), C.POINTER(CvAggregates), C.c_void_p, C.c_void_p]),
    "caviar_last_launch_count": (C.c_int64, []),
    "caviar_residue_components": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double,
                                            C.c_void_p, C.c_void_p]),
    "caviar_cov_padded_columns": (C.c_int64, [C.c_int64]),
    "caviar_cov_split": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "caviar_cov_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_float,
                                  C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "caviar_weighted_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "caviar_features_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                       C.POINTER(CvHostOutputs), C.c_int64, C.c_int32]),
    "caviar_pair_distances_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "caviar_release_cache": (None, []),
    "caviar_pair_distances_host_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "caviar_compact_frames_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                           C.c_void_p]),
    "caviar_compact_frames_soa_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                               C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "caviar_mdcrd_decode_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_void_p]),
    "caviar_xtc_index": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    "caviar_xtc_decode_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_float,
                                         C.c_void_p, C.c_void_p]),
    "caviar_format_csv": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_void_p, C.c_int64]),
    "caviar_last_error": (C.c_char_p, []),
    "caviar_abi_version": (C.c_int, []),
}


def _load():
    global LIB_PATH
    LIB_PATH = os.environ.get("CAVIAR_B200_LIB", LIB_PATH)      # A/B testing of alternative builds
    if not os.path.exists(LIB_PATH):
        # The library is a build artefact (git-ignored).  If it is missing but nvcc is at hand, compile the SAME
        # sm_100a sources in-tree once; there is still no CPU or PyTorch code path -- without the library, no import.
        try:
            from . import build as _build
            _build.build()
        except Exception as exc:
            raise ImportError(
                f"{LIB_PATH} is missing and could not be built ({exc}). caviar_b200 has no CPU or PyTorch fallback: "
                "build the sm_100a library first (python -c 'import __graft_entry__ as g; g.build()'  or  "
                "python caviar_b200/build.py).") from exc
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != CV_OK:
        raise CaviarError(rc, (lib.caviar_last_error() or b"").decode("utf-8", "replace"))
