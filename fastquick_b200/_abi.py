"""ctypes / numpy mirror of include/fastquick_b200.h (struct layouts and prototypes)."""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libfastquick_b200.so")

FQB_MAX_CIGAR = 24
FQB_MAX_MULTI = 11
FQB_BATCH_PAIRS = 0x40000


class GapOpt(C.Structure):
    _fields_ = [
        ("s_mm", C.c_int32), ("s_gapo", C.c_int32), ("s_gape", C.c_int32),
        ("mode", C.c_int32),
        ("indel_end_skip", C.c_int32), ("max_del_occ", C.c_int32), ("max_entries", C.c_int32),
        ("fnr", C.c_double),
        ("max_diff", C.c_int32), ("max_gapo", C.c_int32), ("max_gape", C.c_int32),
        ("max_seed_diff", C.c_int32), ("seed_len", C.c_int32),
        ("max_top2", C.c_int32),
        ("trim_qual", C.c_int32),
        ("flank_len", C.c_int32), ("flank_long_len", C.c_int32),
        ("read_len", C.c_int32),
        ("kmer_thresh", C.c_int32),
        ("is_il13", C.c_int32),
    ]


class PeOpt(C.Structure):
    _fields_ = [
        ("max_isize", C.c_int32), ("force_isize", C.c_int32),
        ("max_occ", C.c_uint32),
        ("n_multi", C.c_int32), ("N_multi", C.c_int32),
        ("type", C.c_int32), ("is_sw", C.c_int32),
        ("ap_prior", C.c_double),
    ]


class ISize(C.Structure):
    _fields_ = [("avg", C.c_double), ("std", C.c_double), ("ap_prior", C.c_double),
                ("low", C.c_uint32), ("high", C.c_uint32), ("high_bayesian", C.c_uint32), ("pad_", C.c_uint32)]


class SynthRefCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_long", C.c_int32), ("n_short", C.c_int32), ("n_x", C.c_int32),
                ("n_y", C.c_int32), ("flank_short", C.c_int32), ("flank_long", C.c_int32), ("spacing", C.c_int32), ("n_dup", C.c_int32)]


class SynthReadCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_int32), ("max_indel_len", C.c_int32),
                ("f_on", C.c_double), ("sub_rate", C.c_double), ("ins_rate", C.c_double), ("del_rate", C.c_double),
                ("n_rate", C.c_double), ("isize_mean", C.c_double), ("isize_sd", C.c_double), ("bad_tail_rate", C.c_double)]


ALN_DTYPE = np.dtype([("k", "<u4"), ("l", "<u4"), ("score", "<i4"),
                      ("n_mm", "u1"), ("n_gapo", "u1"), ("n_gape", "u1"), ("a", "u1")])
assert ALN_DTYPE.itemsize == 16

READ_DTYPE = np.dtype([
    ("pos", "<u4"), ("sa", "<u4"), ("c1", "<u4"), ("c2", "<u4"), ("score", "<i4"),
    ("len", "<i4"), ("full_len", "<i4"), ("clip_len", "<i4"),
    ("type", "u1"), ("strand", "u1"), ("filtered", "u1"), ("extra_flag", "u1"),
    ("n_mm", "u1"), ("n_gapo", "u1"), ("n_gape", "u1"), ("mapQ", "u1"),
    ("seQ", "u1"), ("n_cigar", "u1"), ("n_multi", "u1"), ("has_cigar", "u1"),
    ("nm", "<u2"), ("n_aln", "<u2"),
    ("cigar", "<u2", (FQB_MAX_CIGAR,)),
])
assert READ_DTYPE.itemsize == 96, READ_DTYPE.itemsize

# every extern "C" symbol include/fastquick_b200.h declares
EXPORTED_SYMBOLS = [
    "fqb_gap_opt_default", "fqb_pe_opt_default", "fqb_create", "fqb_destroy", "fqb_last_error",
    "fqb_index_info", "fqb_kmer_tables_origin", "fqb_kmer_tables_fetch", "fqb_align_pairs", "fqb_prefetch_pairs", "fqb_prefetch_hits", "fqb_submit_pairs", "fqb_collect_pairs", "fqb_packed_stride", "fqb_pack_reads", "fqb_align_pairs_packed", "fqb_stage_load_packed", "fqb_submit_pairs_packed", "fqb_comm_ring_handle", "fqb_comm_unique_id", "fqb_comm_init", "fqb_comm_init_local", "fqb_collect_pairs_sharded", "fqb_comm_merge_stats", "fqb_bam_open", "fqb_bam_emit", "fqb_bam_emit2", "fqb_bam_attach", "fqb_bam_close",
    "fqb_stage_load", "fqb_stage_align", "fqb_get_stream_state", "fqb_drand48_zero_index", "fqb_set_stream_state", "fqb_set_pair_base", "fqb_stats_group_bytes", "fqb_stats_export", "fqb_stats_import",
    "fqb_stats_set_target_region", "fqb_stats_open", "fqb_stats_begin_file", "fqb_stats_reset", "fqb_stats_file_counters", "fqb_stage_stats", "fqb_stats_emit", "fqb_stats_emit2", "fqb_emit_sync", "fqb_stats_finish", "fqb_isize_adjusted_file", "fqb_infer_isize_hist", "fqb_isize_penalty", "fqb_host_tables", "fqb_search_buckets", "fqb_stats_var_count", "fqb_stats_var_export", "fqb_stats_var_import", "fqb_stats_close_table", "fqb_stats_merge_tables",
    "fqb_stage_pair", "fqb_stage_sw_refine", "fqb_stage_fetch_rows", "fqb_stage_fetch_rows_async", "fqb_rows_wait", "fqb_reset_stream", "fqb_stage_fetch_prep", "fqb_stage_fetch_aln", "fqb_stage_counters", "fqb_launch_count", "fqb_rank_query_time", "fqb_stream", "fqb_measure_l2", "fqb_feeder_open", "fqb_feeder_format", "fqb_feeder_fill", "fqb_feeder_fill_packed", "fqb_feeder_close", "fqb_gunzip", "fqb_bgzf_compress", "fqb_bgzf_write_file", "fqb_host_alloc", "fqb_host_free",
    "fqb_index_from_flank_fasta", "fqb_synth_ref_cfg_default", "fqb_synth_read_cfg_default", "fqb_synth_create", "fqb_synth_destroy",
    "fqb_create_from_synth", "fqb_synth_write_inputs", "fqb_synth_write_index", "fqb_synth_reads", "fqb_write_fastq_gz",
]


def u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def load_library(path=None):
    """Load the C-ABI library.  There is no CPU fallback: a missing library is an error.
    FQB_LIB_PATH (development only) names another build of the same library, e.g. `make kstats`."""
    if path is None:
        path = os.environ.get("FQB_LIB_PATH", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is the product path; there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.fqb_last_error.restype = C.c_char_p
    return lib
