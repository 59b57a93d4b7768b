"""ctypes binding of libabides_b200.so (the C ABI declared in include/abides_b200.h).

There is no CPU fallback: if the CUDA library has not been built (``python -c "import __graft_entry__ as g;
g.build()"``) loading fails loudly.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libabides_b200.so")

ABX_VERSION = 2
ABX_OK = 0
RNG_PHILOX, RNG_TAPE = 0, 1
LAT_MATRIX_NOISE, LAT_CUBIC = 0, 1

F_DONE, F_QUEUE_OVERFLOW, F_LEVEL_OVERFLOW, F_ORDER_OVERFLOW, F_AGENT_ORDERS_OVERFLOW = 0x1, 0x2, 0x4, 0x8, 0x10
F_THETA_INDEX, F_TAPE_UNDERRUN, F_TAPE_KIND, F_TRACE_OVERFLOW, F_TIME_RANGE = 0x20, 0x40, 0x80, 0x100, 0x200
F_UNSUPPORTED, F_OBS_INVALID = 0x400, 0x800
F_HISTORY_OVERFLOW, F_REF_EXCEPTION, F_ID_RANGE = 0x1000, 0x2000, 0x4000
F_ERROR_MASK = 0x7FFE

MSG_KINDS = [
    "NONE", "WHEN_MKT_OPEN", "WHEN_MKT_CLOSE", "QUERY_SPREAD", "LIMIT_ORDER", "CANCEL_ORDER", "MODIFY_ORDER",
    "ORDER_ACCEPTED", "ORDER_EXECUTED", "ORDER_CANCELLED", "MKT_CLOSED", "QUERY_LAST_TRADE",
    "QUERY_TRANSACTED_VOLUME", "ORDER_MODIFIED", "QUERY_ORDER_STREAM", "MARKET_DATA",
    "MARKET_DATA_SUBSCRIPTION_REQUEST", "MARKET_DATA_SUBSCRIPTION_CANCELLATION",
]


class ZiGroup(C.Structure):
    _fields_ = [("count", C.c_int32), ("r_min", C.c_int32), ("r_max", C.c_int32), ("_pad", C.c_int32),
                ("eta", C.c_double)]


class SimConfig(C.Structure):
    """abx_sim_config (include/abides_b200.h)."""
    _fields_ = [
        ("version", C.c_int32), ("n_agents", C.c_int32), ("n_groups", C.c_int32), ("q_max", C.c_int32),
        ("groups", ZiGroup * 8),
        ("start_ns", C.c_int64), ("stop_ns", C.c_int64), ("mkt_open_ns", C.c_int64), ("mkt_close_ns", C.c_int64),
        ("default_computation_delay_ns", C.c_int64), ("exchange_computation_delay_ns", C.c_int64),
        ("exchange_pipeline_delay_ns", C.c_int64), ("starting_cash", C.c_int64),
        ("order_size", C.c_int32), ("stream_history", C.c_int32),
        ("r_bar", C.c_double), ("kappa", C.c_double), ("fund_vol", C.c_double), ("megashock_lambda_a", C.c_double),
        ("megashock_mean", C.c_double), ("megashock_var", C.c_double),
        ("sigma_n", C.c_double), ("agent_kappa", C.c_double), ("sigma_s", C.c_double), ("sigma_pv", C.c_double),
        ("lambda_a", C.c_double),
        ("latency_model", C.c_int32), ("n_noise", C.c_int32), ("latency_mirrored", C.c_int32), ("_pad0", C.c_int32),
        ("latency_lo", C.c_double), ("latency_hi", C.c_double),
        ("jitter", C.c_double), ("jitter_clip", C.c_double), ("jitter_unit", C.c_double),
        ("queue_cap", C.c_int32), ("level_cap", C.c_int32), ("order_cap", C.c_int32), ("rng_mode", C.c_int32),
        ("trace_cap", C.c_int32), ("hash_pops", C.c_int32),
        ("population", C.c_int32), ("n_noise_agents", C.c_int32), ("n_value_agents", C.c_int32), ("n_mm_agents", C.c_int32),
        ("n_momentum_agents", C.c_int32),
        ("size_lo", C.c_int32), ("size_hi", C.c_int32), ("value_depth_spread", C.c_int32), ("value_percent_aggr", C.c_double),
        ("noise_wake_lo_ns", C.c_int64), ("noise_wake_hi_ns", C.c_int64),
        ("mom_min_size", C.c_int32), ("mom_max_size", C.c_int32), ("mom_wake_ns", C.c_int64),
        ("mm_pov", C.c_double), ("mm_min_order_size", C.c_int32), ("mm_window_size", C.c_int32), ("mm_num_ticks", C.c_int32),
        ("_pad1", C.c_int32), ("mm_wake_ns", C.c_int64),
        ("n_pov_exec", C.c_int32), ("pov_exec_is_buy", C.c_int32), ("pov_exec_pov", C.c_double), ("pov_exec_quantity", C.c_int64),
        ("pov_exec_start_ns", C.c_int64), ("pov_exec_end_ns", C.c_int64), ("pov_exec_freq_ns", C.c_int64), ("pov_exec_lookback_ns", C.c_int64),
        ("draw_log_cap", C.c_int32), ("event_ring_cap", C.c_int32),
        ("hbl_L", C.c_int32), ("mkm_min_size", C.c_int32), ("mkm_max_size", C.c_int32), ("mkm_num_levels", C.c_int32), ("mkm_wake_ns", C.c_int64),
        ("mkm_subscribe", C.c_int32), ("mom_subscribe", C.c_int32), ("mkm_sub_freq_ns", C.c_int64), ("mom_sub_freq_ns", C.c_int64),
        ("exec_kind", C.c_int32), ("exec_limit_price", C.c_int32),
        ("hist_log_cap", C.c_int32), ("hbl_table_rows", C.c_int32),
    ]


class EnvConfig(C.Structure):
    """abx_env_config (include/abides_b200.h)."""
    _fields_ = [
        ("version", C.c_int32), ("order_level", C.c_int32), ("is_buy", C.c_int32), ("n_horizon", C.c_int32),
        ("start_ns", C.c_int64), ("stop_ns", C.c_int64), ("mkt_open_ns", C.c_int64), ("mkt_close_ns", C.c_int64),
        ("horizon_start_ns", C.c_int64), ("horizon_step_ns", C.c_int64), ("quantity", C.c_double), ("steep", C.c_double),
        ("stream_history", C.c_int32), ("queue_cap", C.c_int32), ("level_cap", C.c_int32), ("order_cap", C.c_int32),
        ("trace_cap", C.c_int32), ("hash_pops", C.c_int32),
    ]


class DqConfig(C.Structure):
    """abx_dq_config (include/abides_b200.h)."""
    _fields_ = [
        ("version", C.c_int32), ("n_momentum", C.c_int32), ("n_twap", C.c_int32), ("has_ddqn", C.c_int32), ("is_buy", C.c_int32),
        ("n_horizon", C.c_int32), ("quantity", C.c_int64), ("start_ns", C.c_int64), ("stop_ns", C.c_int64),
        ("mkt_open_ns", C.c_int64), ("mkt_close_ns", C.c_int64), ("horizon_start_ns", C.c_int64), ("horizon_step_ns", C.c_int64),
        ("mom_wake_ns", C.c_int64), ("mom_min_size", C.c_int32), ("mom_max_size", C.c_int32), ("stream_history", C.c_int32),
        ("queue_cap", C.c_int32), ("level_cap", C.c_int32), ("order_cap", C.c_int32), ("trace_cap", C.c_int32), ("hash_pops", C.c_int32),
    ]


class EnvStats(C.Structure):
    """abx_env_stats (include/abides_b200.h)."""
    _fields_ = [
        ("messages", C.c_int64), ("now_ns", C.c_int64), ("pop_hash", C.c_uint64),
        ("limit_orders", C.c_uint32), ("cancels", C.c_uint32), ("fills", C.c_uint32), ("spread_queries", C.c_uint32),
        ("max_queue", C.c_uint32), ("n_bid_levels", C.c_uint32), ("n_ask_levels", C.c_uint32), ("n_resting", C.c_uint32),
        ("best_bid", C.c_int32), ("best_bid_qty", C.c_int32), ("best_ask", C.c_int32), ("best_ask_qty", C.c_int32),
        ("last_trade", C.c_int32), ("fundamental", C.c_int32), ("flags", C.c_uint32), ("trace_len", C.c_uint32),
        ("uniq", C.c_uint32), ("orders_allocated", C.c_uint32), ("sum_shares", C.c_int64), ("sum_cash", C.c_int64),
    ]


class TraceRec(C.Structure):
    _fields_ = [("tag", C.c_int32), ("a", C.c_int32), ("t", C.c_int64), ("v", C.c_int32 * 16)]


STATS_DTYPE = [
    ("messages", "<i8"), ("now_ns", "<i8"), ("pop_hash", "<u8"), ("limit_orders", "<u4"), ("cancels", "<u4"),
    ("fills", "<u4"), ("spread_queries", "<u4"), ("max_queue", "<u4"), ("n_bid_levels", "<u4"),
    ("n_ask_levels", "<u4"), ("n_resting", "<u4"), ("best_bid", "<i4"), ("best_bid_qty", "<i4"), ("best_ask", "<i4"),
    ("best_ask_qty", "<i4"), ("last_trade", "<i4"), ("fundamental", "<i4"), ("flags", "<u4"), ("trace_len", "<u4"),
    ("uniq", "<u4"), ("orders_allocated", "<u4"), ("sum_shares", "<i8"), ("sum_cash", "<i8"),
]
TRACE_DTYPE = [("tag", "<i4"), ("a", "<i4"), ("t", "<i8"), ("v", "<i4", (16,))]
DRAW_DTYPE = [("stream_kind", "<u4"), ("bits", "<u8"), ("_pad", "<u4")]          # abx_draw_rec (bits_lo, bits_hi read as one little-endian u64 at offset 4)

assert C.sizeof(EnvStats) == 112 and C.sizeof(TraceRec) == 80


class AbxError(RuntimeError):
    pass


def _bind(L):
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("abx_strerror", C.c_char_p, i32)
    if hasattr(L, "abx_selftest_log_unit"):            # device self-test: not part of the host emulation used by the CPU tests
        sig("abx_selftest_log_unit", i32, P(C.c_double), P(C.c_double), i32, i32)
    if hasattr(L, "abx_selftest_exp"):
        sig("abx_selftest_exp", i32, P(C.c_double), P(C.c_double), i32, i32)
    sig("abx_last_cuda_error", C.c_char_p)
    sig("abx_device_count", i32)
    sig("abx_config_sparse_zi", i32, i32, P(SimConfig))
    sig("abx_config_rmsc03", i32, P(SimConfig))
    sig("abx_config_rmsc03_pov", i32, P(SimConfig))
    sig("abx_config_rmsc01", i32, P(SimConfig))
    sig("abx_config_rmsc02", i32, P(SimConfig))
    sig("abx_dq_set_schedule", i32, vp, i32, P(i32), i32)
    sig("abx_sim_pov_exec", i32, vp, i32, P(i64), vp)
    sig("abx_sim_create", i32, P(SimConfig), i32, i32, P(vp))
    sig("abx_sim_destroy", i32, vp)
    sig("abx_sim_device_bytes", i64, vp)
    sig("abx_sim_reset_philox", i32, vp, P(C.c_uint64), vp)
    sig("abx_sim_reset_tape", i32, vp, P(C.c_uint64), P(C.c_uint8), P(i64), P(C.c_double), P(C.c_double), vp)
    sig("abx_sim_reset_tape_shared", i32, vp, i32, P(C.c_uint64), P(C.c_uint8), P(i64), P(C.c_double), P(C.c_double), vp)
    sig("abx_sim_run", i32, vp, i64, vp)
    sig("abx_sim_run_each", i32, vp, vp, vp)
    sig("abx_sim_finalize", i32, vp, vp)
    sig("abx_sim_stats", i32, vp, vp, vp)
    sig("abx_sim_stats_device", i32, vp, vp, vp)
    sig("abx_sim_holdings", i32, vp, i32, P(i64), vp)
    sig("abx_sim_book_snapshot", i32, vp, i32, i32, i32, P(i32), P(i32), vp)
    sig("abx_sim_trace", i32, vp, i32, vp, i32, P(i32), vp)
    sig("abx_sim_draw_log", i32, vp, i32, vp, i32, P(i32), vp)
    sig("abx_sim_agent_init", i32, vp, i32, vp, vp, vp, vp, vp, vp)
    sig("abx_sim_events_device", i32, vp, vp, vp, vp)
    sig("abx_sim_launch_count", i64, vp)
    sig("abx_env_config_default", i32, P(EnvConfig))
    sig("abx_env_create", i32, P(EnvConfig), P(i64), i64, i32, i32, P(vp))
    sig("abx_env_create_days", i32, P(EnvConfig), P(i64), P(i64), i32, i32, i32, P(vp))
    sig("abx_dq_create_days", i32, P(DqConfig), P(i64), P(i64), i32, i32, i32, P(vp))
    sig("abx_env_reset", i32, vp, vp)
    sig("abx_env_reset_mask", i32, vp, vp, i32, vp)
    sig("abx_env_set_auto_reset", i32, vp, i32)
    sig("abx_env_step", i32, vp, vp, vp, vp, vp, vp)
    sig("abx_env_step_host", i32, vp, vp, vp, vp, vp, vp)
    sig("abx_book_create", i32, i32, i32, i32, i32, i32, i32, P(vp))
    sig("abx_book_replay", i32, vp, P(i64), i64, vp)
    sig("abx_dq_config_default", i32, P(DqConfig))
    sig("abx_dq_create", i32, P(DqConfig), P(i64), i64, i32, i32, P(vp))
    sig("abx_dq_reset", i32, vp, vp, vp, vp)
    sig("abx_dq_step", i32, vp, vp, vp, vp, vp, vp, vp)
    sig("abx_dq_step_host", i32, vp, vp, vp, vp, vp, vp, vp)
    sig("abx_dq_holdings", i32, vp, i32, vp, vp, vp)
    if hasattr(L, "abx_qnet_create"):                 # the host emulation harness of the CPU test-suite has no tensor-core kernels
        sig("abx_qnet_last_error", C.c_char_p)
        sig("abx_qnet_param_count", i32, P(i32), i32)
        sig("abx_qnet_create", i32, P(i32), i32, P(C.c_float), i32, P(vp))
        sig("abx_qnet_set_params", i32, vp, P(C.c_float), vp)
        sig("abx_qnet_set_params_device", i32, vp, vp, vp)
        sig("abx_qnet_destroy", i32, vp)
        sig("abx_qnet_launch_count", i64, vp)
        sig("abx_qnet_forward", i32, vp, vp, i32, i32, i32, vp, vp, C.c_double, C.c_uint64, C.c_uint64, vp)
    return L


_cache = {}


def load(path=None):
    """Load the C-ABI library.  `path` is only overridden by the CPU test-suite's host emulation harness."""
    path = path or os.environ.get("ABX_LIB_PATH") or LIB_PATH          # ABX_LIB_PATH: kernel A/B experiments (tools/), never set in production
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise AbxError(
            "%s is missing: the sm_100a CUDA library has not been built (run __graft_entry__.build()). "
            "There is no CPU fallback." % path)
    _cache[path] = _bind(C.CDLL(path))
    return _cache[path]


def check(L, status, what):
    if status != ABX_OK:
        msg = L.abx_strerror(status).decode()
        if status == -2:
            msg += ": " + L.abx_last_cuda_error().decode()
        raise AbxError("%s failed: %s (status %d)" % (what, msg, status))
