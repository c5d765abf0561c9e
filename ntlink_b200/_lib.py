"""ctypes binding of libntlink_b200.so (include/ntlink_b200.h). No CPU fallback: a missing library or a missing
CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libntlink_b200.so")

NTL_OK = 0
NTL_ERR = {-1: "NTL_ERR_CUDA", -2: "NTL_ERR_ARG", -3: "NTL_ERR_WORKSPACE", -4: "NTL_ERR_STATE", -5: "NTL_ERR_ASSERT"}
STRAND_BIT = 0x80000000
POS_MASK = 0x7FFFFFFF
T_NAMES = ["pack", "dense", "select", "gap", "emit", "lookup", "chain", "tally", "index", "total"]


class NtlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{NTL_ERR.get(code, code)}: {msg}")
        self.code = code


class SketchOut(C.Structure):
    _fields_ = [("n_mx", C.c_uint64), ("nseq", C.c_uint32), ("reserved", C.c_uint32),
                ("hash", C.POINTER(C.c_uint64)), ("pos_strand", C.POINTER(C.c_uint32)),
                ("seq_off", C.POINTER(C.c_uint64))]


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("w", C.c_int32), ("z", C.c_int32), ("f", C.c_int32), ("x", C.c_double),
                ("sensitive", C.c_int32), ("repeat_filter", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("ctg", C.c_uint32), ("ctg_pos_strand", C.c_uint32), ("read_pos_strand", C.c_uint32)]


class Run(C.Structure):
    _fields_ = [("ctg", C.c_uint32), ("start", C.c_uint32), ("count", C.c_uint32)]


class Event(C.Structure):
    _fields_ = [("read", C.c_uint32), ("ord", C.c_uint32), ("src", C.c_uint32), ("tgt", C.c_uint32),
                ("gap", C.c_int32), ("flags", C.c_uint32)]


class MapOut(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("reserved", C.c_uint32), ("n_mx", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_runs", C.c_uint64), ("n_events", C.c_uint64),
                ("hit_off", C.POINTER(C.c_uint32)), ("nruns", C.POINTER(C.c_uint32)),
                ("runs", C.POINTER(Run)), ("hits", C.POINTER(Hit)),
                ("ev_off", C.POINTER(C.c_uint32)), ("ev_cnt", C.POINTER(C.c_uint32)), ("events", C.POINTER(Event))]


class MappingsOut(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_slots", C.c_uint32), ("n_contigs", C.c_uint32), ("reserved", C.c_uint32),
                ("hit_off", C.POINTER(C.c_uint32)), ("nruns", C.POINTER(C.c_uint32)), ("read_len", C.POINTER(C.c_uint32)),
                ("runs", C.POINTER(C.c_uint32)), ("hits", C.POINTER(C.c_uint32)),
                ("read_names", C.c_void_p), ("read_name_off", C.POINTER(C.c_uint64)),
                ("ctg_names", C.c_void_p), ("ctg_name_off", C.POINTER(C.c_uint64))]


class Pair(C.Structure):
    _fields_ = [("src", C.c_uint32), ("tgt", C.c_uint32), ("flags", C.c_uint32), ("n", C.c_uint32),
                ("anchor", C.c_uint32), ("reserved", C.c_uint32), ("gap_off", C.c_uint64), ("first_key", C.c_uint64)]


class PairsOut(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_gaps", C.c_uint64), ("pairs", C.POINTER(Pair)),
                ("gaps", C.POINTER(C.c_int32))]


class TsvOut(C.Structure):
    _fields_ = [("n_seq", C.c_uint32), ("reserved", C.c_uint32), ("n_mx", C.c_uint64),
                ("hash", C.POINTER(C.c_uint64)), ("pos_strand", C.POINTER(C.c_uint32)), ("mx_off", C.POINTER(C.c_uint64)),
                ("seq_len", C.POINTER(C.c_uint32)), ("names", C.c_void_p), ("name_off", C.POINTER(C.c_uint64))]


# every symbol include/ntlink_b200.h declares: name -> (restype, argtypes)
_VP, _U64P, _U32P = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
SIGNATURES = {
    "ntl_init": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "ntl_destroy": (None, [_VP]),
    "ntl_last_error": (C.c_char_p, [_VP]),
    "ntl_version": (C.c_int, []),
    "ntl_set_option": (C.c_int, [_VP, C.c_char_p, C.c_double]),
    "ntl_sketch": (C.c_int, [_VP, _VP, _VP, C.c_uint32, C.c_int, C.c_int, C.POINTER(SketchOut)]),
    "ntl_index_build": (C.c_int, [_VP, _VP, _VP, _VP, C.c_uint64, _VP, _VP, C.c_uint32]),
    "ntl_index_build_from_sequences": (C.c_int, [_VP, _VP, _VP, C.c_uint32, C.c_int, C.c_int, _VP, C.POINTER(SketchOut)]),
    "ntl_index_stats": (C.c_int, [_VP, _U64P, _U64P, _U64P]),
    "ntl_device_sketch_arrays": (C.c_int, [_VP, _U64P, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "ntl_index_build_device": (C.c_int, [_VP, _VP, _VP, _VP, C.c_uint64, _VP, _VP, C.c_uint32]),
    "ntl_map_reads": (C.c_int, [_VP, _VP, _VP, C.c_uint32, C.c_uint64, C.POINTER(Params), C.POINTER(MapOut)]),
    "ntl_map_sketch": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_uint32, C.c_uint64, C.POINTER(Params), C.POINTER(MapOut)]),
    "ntl_tally_mappings": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_uint32, C.c_uint64, C.POINTER(Params), _U64P]),
    "ntl_map_groups": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP, _VP, _VP, _VP, _VP, C.c_uint32, C.POINTER(Params), C.POINTER(MapOut)]),
    "ntl_liftover_mappings": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP, C.c_uint32, C.c_int, C.POINTER(MapOut)]),
    "ntl_stream": (C.c_int, [_VP, C.POINTER(C.c_void_p)]),
    "ntl_events_export_async": (C.c_int, [_VP, _VP, C.c_uint64, _U64P]),
    "ntl_events_import_counts": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint64, _VP]),
    "ntl_events_import_device": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint64]),
    "ntl_verbose_open": (C.c_int, [C.c_char_p, _VP, _VP, C.c_uint32, C.POINTER(C.c_void_p)]),
    "ntl_verbose_read": (C.c_int, [_VP, C.c_uint64, C.c_int, C.POINTER(MappingsOut)]),
    "ntl_verbose_error": (C.c_char_p, [_VP]),
    "ntl_verbose_close": (None, [_VP]),
    "ntl_tsv_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ntl_tsv_read": (C.c_int, [_VP, C.c_uint64, C.POINTER(TsvOut)]),
    "ntl_tsv_error": (C.c_char_p, [_VP]),
    "ntl_tsv_close": (None, [_VP]),
    "ntl_events_reset": (C.c_int, [_VP]),
    "ntl_events_append": (C.c_int, [_VP, _VP, C.c_uint64]),
    "ntl_events_count": (C.c_int, [_VP, _U64P]),
    "ntl_events_device": (C.c_int, [_VP, _U64P, C.POINTER(_VP)]),
    "ntl_events_append_device": (C.c_int, [_VP, _VP, C.c_uint64]),
    "ntl_events_export": (C.c_int, [_VP, _VP, C.c_uint64, _U64P]),
    "ntl_events_import_gathered": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint64, C.POINTER(C.c_int)]),
    "ntl_pairs_finish": (C.c_int, [_VP, C.POINTER(PairsOut)]),
    "ntl_format_sketch_tsv": (C.c_int64, [C.POINTER(SketchOut), _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.POINTER(_VP)]),
    "ntl_format_verbose": (C.c_int64, [C.POINTER(MapOut), _VP, _VP, _VP, _VP, C.c_int, C.POINTER(_VP)]),
    "ntl_format_paf": (C.c_int64, [C.POINTER(MapOut), _VP, _VP, _VP, _VP, _VP, _VP, C.c_int, C.c_int, C.POINTER(_VP)]),
    "ntl_buf_free": (None, [_VP]),
    "ntl_seqfile_open": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
    "ntl_seqfile_read": (C.c_int, [_VP, C.c_uint64, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), _U32P]),
    "ntl_seqfile_close": (None, [_VP]),
    "ntl_free": (None, [_VP]),
    "ntl_reads_upload": (C.c_int, [_VP, _VP, _VP, C.c_uint32]),
    "ntl_map_resident": (C.c_int, [_VP, C.c_uint64, C.POINTER(Params), C.POINTER(MapOut)]),
    "ntl_target_upload": (C.c_int, [_VP, _VP, _VP, C.c_uint32, _VP]),
    "ntl_index_build_resident": (C.c_int, [_VP, C.c_int, C.c_int]),
    "ntl_timing_reset": (C.c_int, [_VP]),
    "ntl_get_stat": (C.c_int, [_VP, C.c_char_p, C.POINTER(C.c_double)]),
    "ntl_timing": (C.c_int, [_VP, C.POINTER(C.c_double), _U64P, _U64P, _U64P]),
    "ntl_timing_dense": (C.c_int, [_VP, C.POINTER(C.c_double), _U64P, _U64P]),
    "ntl_device_sync": (C.c_int, [_VP]),
    "ntl_mark": (C.c_int, [_VP, C.c_int]),
    "ntl_mark_elapsed": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "ntl_copy_device": (C.c_int, [_VP, _VP, _VP, C.c_uint64]),
    "ntl_target_sketch_resident": (C.c_int, [_VP, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _U64P, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "ntl_target_resident_meta": (C.c_int, [_VP, _VP, _VP]),
    "ntl_synth_target_resident": (C.c_int, [_VP, C.c_uint64, _VP, C.c_uint32, _VP]),
    "ntl_synth_reads_resident": (C.c_int, [_VP, C.c_uint64, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _U64P]),
    "ntl_resident_info": (C.c_int, [_VP, C.c_int, _U32P, _U64P]),
    "ntl_resident_download": (C.c_int, [_VP, C.c_int, C.c_uint32, C.c_uint32, _VP, _VP]),
    "ntl_synth_host_contigs": (C.c_int, [C.c_uint64, _VP, C.c_uint32, _VP, _VP, C.c_int]),
    "ntl_synth_host_reads": (C.c_int, [C.c_uint64, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _VP, _VP, C.c_int]),
}

_lib = None


def load():
    """Load libntlink_b200.so (built in-tree by ntlink_b200/build.py). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m ntlink_b200.build` (needs nvcc). "
                              "ntlink_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
