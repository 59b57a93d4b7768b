/*
 * abides_b200.h -- C ABI of the B200-native batched ABIDES market simulator (libabides_b200.so).
 *
 * The reference (yutiansut/marl-optimal-execution, an ABIDES fork) is 100 % Python and has no FFI; its
 * hot path is entered through three plain call surfaces.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference root).  All functions are
 * extern "C", take plain pointers and sizes (no torch / C++ types), return an int32 status
 * (ABX_OK == 0, negative == error, see abx_strerror) and never throw.  A handle is owned by one host
 * thread at a time (the reference's callers are single threaded: Kernel.py:190, ABIDESEnv.py:30).
 * `stream` arguments are a cudaStream_t passed as void* (NULL == default stream).
 *
 * There is no CPU fallback: every entry point that computes runs hand-written sm_100a kernels and
 * returns ABX_ERR_CUDA when no usable device is present.
 */
#ifndef ABIDES_B200_H
#define ABIDES_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABX_VERSION 2

enum abx_status {
  ABX_OK = 0,
  ABX_ERR_ARG = -1,       /* bad argument / config */
  ABX_ERR_CUDA = -2,      /* CUDA runtime error (abx_last_cuda_error has the text) */
  ABX_ERR_STATE = -3,     /* call sequence error (e.g. run before reset) */
  ABX_ERR_CAPACITY = -4   /* a fixed-capacity structure overflowed in some environment (see abx_env_stats.flags) */
};

/* msg.body["msg"] strings of the reference protocol (agent/ExchangeAgent.py:129-340,
 * agent/TradingAgent.py:181-268). */
enum abx_msg_kind {
  ABX_NONE = 0, ABX_WHEN_MKT_OPEN, ABX_WHEN_MKT_CLOSE, ABX_QUERY_SPREAD, ABX_LIMIT_ORDER, ABX_CANCEL_ORDER,
  ABX_MODIFY_ORDER, ABX_ORDER_ACCEPTED, ABX_ORDER_EXECUTED, ABX_ORDER_CANCELLED, ABX_MKT_CLOSED,
  ABX_QUERY_LAST_TRADE, ABX_QUERY_TRANSACTED_VOLUME, ABX_ORDER_MODIFIED, ABX_QUERY_ORDER_STREAM, ABX_MARKET_DATA,
  ABX_MARKET_DATA_SUBSCRIPTION_REQUEST, ABX_MARKET_DATA_SUBSCRIPTION_CANCELLATION
};
/* queue entry types, message/Message.py:5-10 (tie-break order MESSAGE < WAKEUP < CANCEL_ORDER) */
enum abx_event_type { ABX_T_MESSAGE = 1, ABX_T_WAKEUP = 2, ABX_T_CANCEL_ORDER = 3 };

enum abx_rng_mode {
  ABX_RNG_PHILOX = 0, /* counter-based Philox4x32-10 streams keyed by (env seed, stream id) */
  ABX_RNG_TAPE = 1    /* replay of recorded standard variates, one tape per RandomState of the reference */
};
enum abx_latency_model {
  ABX_LAT_MATRIX_NOISE = 0, /* Kernel.py:410-412: pairwise latency + uniform integer noise in [0, n_noise) */
  ABX_LAT_CUBIC = 1,        /* model/LatencyModel.py:109-140 */
  ABX_LAT_ZERO = 2          /* zero latency matrix with noise [1.0] (ABIDESEnv.py:91-92): delivery == send time, no draw */
};

/* per-environment flag bits in abx_env_stats.flags */
#define ABX_F_DONE            0x001u /* event loop ended (queue empty or past stop time, Kernel.py:190) */
#define ABX_F_QUEUE_OVERFLOW  0x002u
#define ABX_F_LEVEL_OVERFLOW  0x004u
#define ABX_F_ORDER_OVERFLOW  0x008u
#define ABX_F_AGENT_ORDERS_OVERFLOW 0x010u
#define ABX_F_THETA_INDEX     0x020u /* the reference would have raised IndexError (ZeroIntelligenceAgent.py:268) */
#define ABX_F_TAPE_UNDERRUN   0x040u
#define ABX_F_TAPE_KIND       0x080u /* tape entry kind differs from the draw the simulator asked for */
#define ABX_F_TRACE_OVERFLOW  0x100u
#define ABX_F_TIME_RANGE      0x200u
#define ABX_F_UNSUPPORTED     0x400u /* a reference behaviour this build does not model was hit (no code path sets it since round 2; kept for ABI stability) */
#define ABX_F_OBS_INVALID     0x800u /* get_observation would have raised in the reference (empty side / no stored LOB) */
#define ABX_F_HISTORY_OVERFLOW 0x1000u /* the ring of transaction tuples behind get_transacted_volume overwrote one that could still count */
#define ABX_F_REF_EXCEPTION   0x2000u /* the reference would have raised: KeyError of MarketReplayAgent.wakeup off a stream timestamp, IndexError of take_action below 4 levels */
#define ABX_F_ID_RANGE        0x4000u /* generated order ids reached the replayed stream's explicit ORDER_IDs (the range abx_*_create checked was exceeded) */

/* One ZeroIntelligenceAgent strategy group: config/sparse_zi_1000.py:196-204 tuples (n, R_min, R_max, eta). */
typedef struct abx_zi_group {
  int32_t count, r_min, r_max, _pad;
  double eta;
} abx_zi_group;

/* Replaces the config-as-script surface (config/sparse_zi_100.py, config/sparse_zi_1000.py) plus the
 * Kernel.runner arguments (Kernel.py:50-64).  Times are int64 ns since midnight of the simulated date. */
typedef struct abx_sim_config {
  int32_t version;                 /* ABX_VERSION */
  int32_t n_agents;                /* exchange (id 0) + traders */
  int32_t n_groups;                /* <= 8 */
  int32_t q_max;                   /* ZeroIntelligenceAgent q_max (theta has 2*q_max entries, <= 20) */
  abx_zi_group groups[8];
  int64_t start_ns, stop_ns;       /* Kernel.runner startTime / stopTime */
  int64_t mkt_open_ns, mkt_close_ns;
  int64_t default_computation_delay_ns; /* Kernel.runner defaultComputationDelay */
  int64_t exchange_computation_delay_ns, exchange_pipeline_delay_ns; /* ExchangeAgent.py:56-59 */
  int64_t starting_cash;           /* cents */
  int32_t order_size;              /* ZeroIntelligenceAgent.py:308 (100) */
  int32_t stream_history;          /* ExchangeAgent stream_history: OrderBook.history keeps this many trade-delimited buckets besides the open one (read by MODIFY_ORDER fan-out,
                                    * get_transacted_volume and QUERY_ORDER_STREAM; the sparse_zi populations never read it) */
  /* util/oracle/SparseMeanRevertingOracle.py symbol parameters */
  double r_bar, kappa, fund_vol, megashock_lambda_a, megashock_mean, megashock_var;
  /* ZeroIntelligenceAgent parameters */
  double sigma_n, agent_kappa, sigma_s, sigma_pv, lambda_a;
  /* latency */
  int32_t latency_model;           /* abx_latency_model */
  int32_t n_noise;                 /* len(latencyNoise) for ABX_LAT_MATRIX_NOISE */
  int32_t latency_mirrored;        /* 1: latency[i][0] == latency[0][i] (config/sparse_zi_1000.py:264-277) */
  int32_t _pad0;
  double latency_lo, latency_hi;   /* U(lo, hi) pairwise (min) latency, ns (Philox mode draws it on device) */
  double jitter, jitter_clip, jitter_unit;
  /* capacities (per environment) */
  int32_t queue_cap;               /* event slots, multiple of 32, <= 4096 */
  int32_t level_cap;               /* price levels per side, <= 2048 */
  int32_t order_cap;               /* resting orders, <= 65535 */
  int32_t rng_mode;                /* abx_rng_mode */
  int32_t trace_cap;               /* trace records per environment (0 = tracing off) */
  int32_t hash_pops;               /* 1: maintain the FNV-1a hash of the pop sequence (parity runs) */
  /* population layout.  0: ZeroIntelligence groups (above).  1: config/rmsc03.py -- ids 1..n_noise_agents NoiseAgents, then n_value_agents
   * ValueAgents (they use sigma_n / agent_kappa / sigma_s / lambda_a above), n_mm_agents POVMarketMakerAgents (0 or 1), n_momentum_agents
   * MomentumAgents; zero latency (latency_model ABX_LAT_ZERO).  3: config/rmsc01.py (see hbl_L below). */
  int32_t population, n_noise_agents, n_value_agents, n_mm_agents, n_momentum_agents;
  int32_t size_lo, size_hi;        /* Noise/Value order size = np.random.randint(lo, hi) (agent/NoiseAgent.py:34, ValueAgent.py:55) */
  int32_t value_depth_spread;      /* ValueAgent.depth_spread (2) */
  double value_percent_aggr;       /* ValueAgent.percent_aggr (0.1) */
  int64_t noise_wake_lo_ns, noise_wake_hi_ns;  /* util.get_wake_time(noise_mkt_open, noise_mkt_close) window (config/rmsc03.py:115-116) */
  int32_t mom_min_size, mom_max_size; int64_t mom_wake_ns;            /* MomentumAgent min_size, max_size, wake_up_freq */
  double mm_pov; int32_t mm_min_order_size, mm_window_size, mm_num_ticks, _pad1; int64_t mm_wake_ns;   /* POVMarketMakerAgent */
  /* population 1, optional: one POVExecutionAgent (agent/execution/baselines/pov_agent.py; config/execution_iabs_plots.py:200-226) as the LAST
   * agent id -- BASELINE.json configs[2] "rmsc03 ... with POV execution agent".  Every `freq` it cancels its orders, asks for the whole book
   * (depth sys.maxsize) and the transacted volume of the last `lookback`, and sends round(pov * volume) as a client-side market order. */
  int32_t n_pov_exec, pov_exec_is_buy;
  double pov_exec_pov; int64_t pov_exec_quantity, pov_exec_start_ns, pov_exec_end_ns, pov_exec_freq_ns, pov_exec_lookback_ns;
  /* parity instrumentation (like trace_cap / hash_pops): > 0 keeps, per environment, a log of every standard variate the Philox streams
   * hand out (abx_sim_draw_log), so that the reference's algorithm can be re-run on exactly those draws (oracle external tapes) */
  int32_t draw_log_cap;
  /* > 0: keep, per environment, a ring of the last event_ring_cap exchange events the reference logs for its realism tooling (abx_sim_events):
   * order arrivals and the BEST_BID / BEST_ASK / LAST_TRADE lines of util/OrderBook.py:114-141 */
  int32_t event_ring_cap;
  /* population 3, config/rmsc01.py: ids 1..n_mm_agents MarketMakerAgents (agent/market_makers/MarketMakerAgent.py, polling mode), then groups[0].count
   * ZeroIntelligenceAgents, groups[1].count HeuristicBeliefLearningAgents (agent/HeuristicBeliefLearningAgent.py, history length hbl_L), then n_momentum_agents
   * MomentumAgents; zero latency.  hist_log_cap: entries of the per-environment order-history log behind QUERY_ORDER_STREAM (util/OrderBook.py:52-60). */
  int32_t hbl_L, mkm_min_size, mkm_max_size, mkm_num_levels;
  int64_t mkm_wake_ns;
  /* config/rmsc02.py: the same population with the market maker and / or the momentum agents in SUBSCRIPTION mode (MARKET_DATA_SUBSCRIPTION_REQUEST at their first
   * wake-up, agent/TradingAgent.py:160-172; the exchange publishes MARKET_DATA after every book operation to subscribers whose `freq` ns have passed,
   * agent/ExchangeAgent.py:342-387).  mkm_num_levels levels a side for the market maker (subscribe_num_levels), 1 for momentum agents. */
  int32_t mkm_subscribe, mom_subscribe;
  int64_t mkm_sub_freq_ns, mom_sub_freq_ns;   /* subscribe_freq (10e9) / MomentumAgent.py:58 (10e9) */
  /* population 1, the optional execution agent (n_pov_exec = 1) as one of the other baselines instead of the POV agent: exec_kind 1 = PassiveAgent
   * (agent/execution/baselines/passive_agent.py: one LIMIT order of pov_exec_quantity at pov_exec_start_ns, at exec_limit_price or, when that is 0, at the best bid (BUY) /
   * ask (SELL) of a QUERY_SPREAD it sends then), 2 = AggressiveAgent (aggressive_agent.py: getCurrentSpread(depth=100) at pov_exec_start_ns, then a market order of
   * pov_exec_quantity walked over the levels it was told, TradingAgent.py:351-397).  0 = POVExecutionAgent. */
  int32_t exec_kind, exec_limit_price;
  int32_t hist_log_cap, hbl_table_rows;   /* hbl_table_rows: price rows of the HBL belief table held per environment, 0 = hist_log_cap / 4 (the maximum); wider price spans take a slower exact form */
} abx_sim_config;

/* Per-environment counters; replaces the "Event Queue elapsed ..., messages: N" line (Kernel.py:321-327). */
typedef struct abx_env_stats {
  int64_t messages;      /* ttl_messages, Kernel.py:211 (requeued pops included) */
  int64_t now_ns;        /* Kernel.currentTime */
  uint64_t pop_hash;     /* FNV-1a over (t, recipient, type, uniq) of every pop when hash_pops */
  uint32_t limit_orders, cancels, fills, spread_queries;
  uint32_t max_queue, n_bid_levels, n_ask_levels, n_resting;
  int32_t best_bid, best_bid_qty, best_ask, best_ask_qty; /* getInsideBids(1)/getInsideAsks(1); qty 0 == empty */
  int32_t last_trade, fundamental;
  uint32_t flags, trace_len;
  uint32_t uniq, orders_allocated;
  int64_t sum_shares, sum_cash; /* filled by abx_sim_finalize: conservation checksums over traders */
} abx_env_stats;

/* One trace record (80 bytes), written when trace_cap > 0; layouts mirror tools/record_reference.py rows.
 *   tag 0  queue pop (Kernel.py:192):        a = recipient, t, v = {type, uniq (-1 for msg None), kind}
 *   tag 1  exchange outbound message          a = recipient, t, v = {kind, order_id, is_buy, qty, limit_price,
 *          (agent/ExchangeAgent.py:471-485):      fill_price | last_trade, bid, bid_qty, ask, ask_qty, mkt_closed}
 *   tag 2  book state after a book op:        t, v = {n_bid_levels, n_ask_levels, n_resting,
 *          (util/OrderBook.py:38,284)             bid p,q x3, ask p,q x3, last_trade}                         */
typedef struct abx_trace_rec {
  int32_t tag;
  int32_t a;
  int64_t t;
  int32_t v[16];
} abx_trace_rec;

typedef struct abx_sim abx_sim; /* opaque */

/* Self-test of the device-side logarithm used by the Philox-mode variate transforms (Box-Muller, exponential inter-arrival times; the reference draws them
 * with numpy's RandomState.normal / .exponential, e.g. agent/ZeroIntelligenceAgent.py:349-350, util/oracle/SparseMeanRevertingOracle.py:105-107):
 * y[i] = log(x[i]) for x in (0, 1], host buffers.  Tests compare it with libm (relative error < 1e-11). */
int32_t abx_selftest_log_unit(const double *x_host, double *y_host, int32_t n, int32_t device);
/* Device self-test of the exp the kernels use for exp(-kappa d) of the OU step (util/oracle/SparseMeanRevertingOracle.py:105-106) and the (1 - kappa) ** x
 * powers of the ZI belief update (agent/ZeroIntelligenceAgent.py:229-256): y[i] = exp(x[i]), host buffers.  Tests compare it with libm (<= 1 ulp, equal on > 99.9 %). */
int32_t abx_selftest_exp(const double *x_host, double *y_host, int32_t n, int32_t device);
const char *abx_strerror(int32_t status);
const char *abx_last_cuda_error(void);
int32_t abx_device_count(void);

/* Fill `cfg` with the population/parameters of config/sparse_zi_100.py (variant 100) or
 * config/sparse_zi_1000.py (variant 1000), capacities sized from the measured maxima. */
int32_t abx_config_sparse_zi(int32_t variant, abx_sim_config *cfg);
/* config/rmsc03.py:49-232: 1 exchange + 50 Noise + 10 Value + 1 POV market maker + 2 Momentum agents, 09:30 -> 09:45 (+1 min). */
int32_t abx_config_rmsc03(abx_sim_config *cfg);
/* The same population plus one POVExecutionAgent (id 64): BUY 120 000 at 50 % of the volume transacted in the last 30 s, every 30 s, 09:32 -> 09:43. */
int32_t abx_config_rmsc03_pov(abx_sim_config *cfg);
/* config/rmsc01.py:60-262: 1 exchange + 1 MarketMakerAgent + 50 ZeroIntelligenceAgents + 25 HeuristicBeliefLearningAgents (L = 2, served by the exchange's
 * QUERY_ORDER_STREAM, agent/ExchangeAgent.py:251-279) + 24 MomentumAgents, 09:30 -> 16:00 (+1 min); zero latency, zero computation delay. */
int32_t abx_config_rmsc01(abx_sim_config *cfg);
/* config/rmsc02.py: the rmsc01 population with the market maker and the momentum agents in subscription mode (MARKET_DATA), pairwise latency U(21 us, 13 ms) + 6-entry
 * noise, midnight -> 17:00. */
int32_t abx_config_rmsc02(abx_sim_config *cfg);
/* POVExecutionAgent of one environment: out HOST int64 [3] = remaining quantity, executed orders, open orders. */
int32_t abx_sim_pov_exec(abx_sim *h, int32_t env, int64_t *out, void *stream);

/* Replaces: Kernel(...) construction + agent list construction (config/sparse_zi_1000.py:146-251).
 * Allocates all per-environment state for n_envs independent simulations on CUDA device `device`. */
int32_t abx_sim_create(const abx_sim_config *cfg, int32_t n_envs, int32_t device, abx_sim **out);
int32_t abx_sim_destroy(abx_sim *h);
int64_t abx_sim_device_bytes(const abx_sim *h);

/* Replaces: np.random.seed(seed) and the per-object RandomState cascade, agent __init__
 * (ZeroIntelligenceAgent.py:65-70 theta draw), latency matrix draw, oracle __init__, Kernel.runner :154-175
 * (kernelInitializing / kernelStarting: one WAKEUP per agent at start_ns).  seeds: host array [n_envs]. */
int32_t abx_sim_reset_philox(abx_sim *h, const uint64_t *seeds, void *stream);

/* Tape-mode reset of ALL environments.  Host arrays:
 *   tape_bits  [total]            standard variates: fp64 bits for kinds 'n','e','u', integer offset for 'i'
 *   tape_kinds [total]            'n' normal, 'e' exponential, 'u' uniform, 'i' randint (value - low)
 *   tape_offsets [n_envs*(n_agents+3)+1]  per env, streams in order: 0 symbol, 1 kernel, 2 latency model,
 *                                 3 global (megashock gaps, kind 'e'), 3+a agent a (a = 1..n_agents-1)
 *   lat_to_exchange / lat_from_exchange [n_envs*n_agents]  fp64 ns: latency[a][0] / latency[0][a]
 *       (population 1 has zero latency: the two arrays instead carry what the config script drew from the global stream
 *        before the kernel started -- lat_to = NoiseAgent/ValueAgent.size, lat_from = NoiseAgent.wakeup_time in ns)
 * Replays what Kernel/agents/oracle drew in a recorded reference run (rng_mode must be ABX_RNG_TAPE). */
int32_t abx_sim_reset_tape(abx_sim *h, const uint64_t *tape_bits, const uint8_t *tape_kinds,
                           const int64_t *tape_offsets, const double *lat_to_exchange,
                           const double *lat_from_exchange, void *stream);

/* Tape-mode reset with SHARED tapes: n_tapes recorded runs, environment e replays run e % n_tapes (parity at production occupancy: thousands of
 * environments without thousands of copies of the tapes).  tape_offsets [n_tapes*(n_agents+3)+1], lat_* [n_tapes*n_agents]. */
int32_t abx_sim_reset_tape_shared(abx_sim *h, int32_t n_tapes, const uint64_t *tape_bits, const uint8_t *tape_kinds, const int64_t *tape_offsets,
                                  const double *lat_to_exchange, const double *lat_from_exchange, void *stream);

/* Replaces: the hot loop of Kernel.runner (Kernel.py:190-292) for every environment at once.  Pops events
 * while the next event time <= until_ns (and the reference loop condition holds).  Asynchronous on `stream`. */
int32_t abx_sim_run(abx_sim *h, int64_t until_ns, void *stream);

/* Same, with a per-environment horizon: until_ns_host is a HOST array [n_envs] (pinned memory makes the copy
 * asynchronous); it is copied to the device on `stream` before the launch.  This is the vectorised form of the
 * gym-style stepping surface (GymKernel.stepRunner, GymKernel.py:158-306), where every environment stops at its
 * own next decision time. */
int32_t abx_sim_run_each(abx_sim *h, const int64_t *until_ns_host, void *stream);

/* Replaces: Kernel.runner :310-311 (kernelStopping of every agent: mark to market, ZI surplus valuation,
 * which advances the shared fundamental, ZeroIntelligenceAgent.py:80-123). */
int32_t abx_sim_finalize(abx_sim *h, void *stream);

/* Device -> host copy of per-environment counters (synchronises `stream`).  out: host [n_envs]. */
int32_t abx_sim_stats(abx_sim *h, abx_env_stats *out, void *stream);
/* Same counters left on the device: out_dev is a DEVICE pointer [n_envs] (e.g. a torch tensor's data_ptr). */
int32_t abx_sim_stats_device(abx_sim *h, abx_env_stats *out_dev, void *stream);

/* Replaces: the "Final holdings for ..." lines (agent/TradingAgent.py:124-126) for one environment.
 * out: host int64 [(n_agents-1) * 5] rows (agent id, shares, cash, marked_to_market, surplus). */
int32_t abx_sim_holdings(abx_sim *h, int32_t env, int64_t *out, void *stream);

/* Replaces: OrderBook.getInsideBids(depth) / getInsideAsks(depth) (util/OrderBook.py:377-398) for one env.
 * out: host int32 [2*depth] (price, qty) pairs best first; returns the number of levels in *n_levels. */
int32_t abx_sim_book_snapshot(abx_sim *h, int32_t env, int32_t is_bid, int32_t depth, int32_t *out,
                              int32_t *n_levels, void *stream);

/* One entry of the draw log (draw_log_cap > 0, ABX_RNG_PHILOX): stream (0 symbol, 1 kernel, 2 latency model, 3 global, 3 + a agent a -- the
 * stream numbering of abx_sim_reset_tape), kind ('n' 'e' 'u' 'i', as on a tape) and the standard variate's bits, in the order drawn. */
typedef struct abx_draw_rec { uint32_t stream_kind; /* stream | kind << 24 */ uint32_t bits_lo, bits_hi, _pad; } abx_draw_rec;
/* Copy the draw log of one environment to the host.  out: host [max_recs]; *n_recs = entries written. */
int32_t abx_sim_draw_log(abx_sim *h, int32_t env, abx_draw_rec *out, int32_t max_recs, int32_t *n_recs, void *stream);
/* Start-of-run state of one environment's traders as the reset left it (what the config script / agent constructors draw before the kernel
 * starts): theta HOST int32 [n_agents][20] (ZeroIntelligenceAgent private values, sorted), lat_to / lat_from HOST fp64 [n_agents]
 * (latency[a][0], latency[0][a]), sizes HOST int32 [n_agents] (Noise / Value / Momentum order size), wakes HOST int64 [n_agents] (NoiseAgent
 * wake-up time, ns).  Row 0 (the exchange) is zero.  Any pointer may be NULL.  Call after reset, before the first run. */
int32_t abx_sim_agent_init(abx_sim *h, int32_t env, int32_t *theta, double *lat_to, double *lat_from, int32_t *sizes, int64_t *wakes, void *stream);

/* One record of the exchange event ring (event_ring_cap > 0), 16 bytes.  kind: ABX_EV_ORDER = a LIMIT_ORDER reached the book (a = limit price, b = quantity,
 * signed: > 0 buy) -- the order stream behind realism/order_flow_stylized_facts.py; ABX_EV_BEST_BID / ABX_EV_BEST_ASK (a = price, b = total quantity of the
 * level) and ABX_EV_LAST_TRADE (a = int(round(average price)), b = traded quantity) -- the logEvent lines of util/OrderBook.py:114-141.
 * t_kind = t_ns | kind << 60. */
enum abx_event_kind { ABX_EV_ORDER = 0, ABX_EV_BEST_BID = 1, ABX_EV_BEST_ASK = 2, ABX_EV_LAST_TRADE = 3 };
typedef struct abx_event_rec { uint64_t t_kind; int32_t a, b; } abx_event_rec;
/* Copy every environment's ring and event count into DEVICE buffers (e.g. torch tensors): out_dev [n_envs][event_ring_cap] records (slot i holds event
 * number i mod cap), counts_dev uint32 [n_envs] = events logged since the reset.  Asynchronous on `stream`. */
int32_t abx_sim_events_device(abx_sim *h, abx_event_rec *out_dev, uint32_t *counts_dev, void *stream);

/* Copy the trace of one environment to the host.  out: host [max_recs]; *n_recs = records written. */
int32_t abx_sim_trace(abx_sim *h, int32_t env, abx_trace_rec *out, int32_t max_recs, int32_t *n_recs, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * ABIDESEnv shape: Exchange (id 0) + MarketReplayAgent (id 1) + DummyRLExecutionAgent (id 2) under GymKernel.
 * Replaces the gym surface ABIDESEnv.py:7-57 (reset / step) for n_envs environments at once.  Handles are abx_sim*;
 * abx_sim_stats / abx_sim_trace / abx_sim_book_snapshot / abx_sim_destroy / abx_sim_launch_count work on them too
 * (traces report the stream's original ORDER_IDs).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct abx_env_config {
  int32_t version;                 /* ABX_VERSION */
  int32_t order_level;             /* DummyRLExecutionAgent order_level (1 or 2): action has order_level + 1 entries */
  int32_t is_buy;                  /* direction == "BUY" */
  int32_t n_horizon;               /* len(execution_time_horizon) (agent_config.py:134-136: 761) */
  int64_t start_ns, stop_ns;       /* ABIDESEnv.py:86-88: midnight .. 16:10 */
  int64_t mkt_open_ns, mkt_close_ns;
  int64_t horizon_start_ns, horizon_step_ns; /* 09:40:00, 30 s */
  double quantity, steep;          /* parent order size, a_q_map_steep_factor */
  int32_t stream_history;          /* ExchangeAgent stream_history (10): bounds the ORDER_MODIFIED fan-out */
  int32_t queue_cap, level_cap, order_cap;   /* per-environment capacities */
  int32_t trace_cap, hash_pops;
} abx_env_config;

/* Defaults of ABIDESEnv.initAgents / agent_config.py (BUY 1e5 shares, "30S", 09:40 -> 16:00, order_level 2, steep 0.5). */
int32_t abx_env_config_default(abx_env_config *cfg);

/* Replaces: ABIDESEnv.__init__ (ABIDESEnv.py:8-26) + LOBSTEROrdersProcessor (agent/examples/MarketReplayAgent.py:162-220
 * output: the parsed stream).  stream5: HOST int64 [n_rows][5] rows (t_ns since midnight, ORDER_ID, PRICE cents, SIZE,
 * is_buy) sorted by time; every environment replays this stream. */
int32_t abx_env_create(const abx_env_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out);
/* Several replayed days in one handle (the reference runs one process per date, e.g. config/execution/marketreplay/..._parallel.py): stream5 is the
 * concatenation of the days' rows, row_offsets HOST int64 [n_days + 1]; environment e replays day e % n_days. */
int32_t abx_env_create_days(const abx_env_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out);
/* Replaces: ABIDESEnv.reset() (ABIDESEnv.py:51-57): initAgents + GymKernel.initRunner. */
int32_t abx_env_reset(abx_sim *h, void *stream);
/* ABIDESEnv.reset() of SOME environments (the reference's reset is per environment object, ABIDESEnv.py:51-57): mask_dev is a DEVICE uint8 [n_envs],
 * environments with a non-zero entry start a fresh episode; advance_day != 0 moves each of them on to its next replayed day (environment e replays
 * day (e + resets so far) % n_days -- the date sweep of config/execution/marketreplay/..._parallel.py), 0 restarts the same day.  Works on
 * abx_env_create* and abx_dq_create* handles (the DDQN shape re-draws its MomentumAgent sizes from seed + episode unless sizes were supplied). */
int32_t abx_env_reset_mask(abx_sim *h, const uint8_t *mask_dev, int32_t advance_day, void *stream);
/* Auto-reset: after every abx_env_step / abx_dq_step the environments whose event loop has ended (done == 1 in that step's output) are reset
 * before the next step.  mode 0 off (default), 1 restart the same day, 2 move on to the next day. */
int32_t abx_env_set_auto_reset(abx_sim *h, int32_t mode);
/* Replaces: ABIDESEnv.step(action) (ABIDESEnv.py:30-49) -> GymKernel.stepRunner (GymKernel.py:158-306) for every environment.
 * DEVICE pointers: actions fp64 [n_envs][3] (x_hat, o_hat_1, o_hat_2), obs fp64 [n_envs][9] (zeros when the reference
 * would return []), reward fp64 [n_envs] (0: the reference's get_reward returns None), done uint8 [n_envs]. */
int32_t abx_env_step(abx_sim *h, const double *actions_dev, double *obs_dev, double *reward_dev, uint8_t *done_dev, void *stream);
/* Same call with HOST buffers (pinned memory keeps the copies asynchronous); synchronises `stream`. */
int32_t abx_env_step_host(abx_sim *h, const double *actions, double *obs, double *reward, uint8_t *done, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * DDQN execution shape: config/execution/marketreplay/execution_marketreplay_ddqn.py (-a rl) -- Exchange (id 0) +
 * MarketReplayAgent (1) + n_momentum MomentumAgents (2..) + n_twap TWAPExecutionAgents + DDQLearningExecutionAgent (last id)
 * under Kernel.runner, for n_envs environments at once.  The reference agent calls its Keras Q-network once per decision tick,
 * batch 1, inside place_order (agent/execution/qlearning/ddqlearning_execution_agent.py:245,339-365); here one abx_dq_step runs
 * every environment up to that call, the caller evaluates the Q-network for the whole batch (abx_qnet_forward below) and the
 * next abx_dq_step resumes place_order with the chosen actions.  Handles are abx_sim* (stats / trace / snapshot / destroy work).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct abx_dq_config {
  int32_t version;                 /* ABX_VERSION */
  int32_t n_momentum;              /* MomentumAgents (config: 7), <= 8 */
  int32_t n_twap;                  /* TWAPExecutionAgents (config -a rl: 1), 0..2 */
  int32_t has_ddqn;                /* DDQLearningExecutionAgent present (the last agent id) */
  int32_t is_buy;                  /* --direction BUY */
  int32_t n_horizon;               /* len(execution_time_horizon): horizon_length / freq + 1 (config: 661) */
  int64_t quantity;                /* --parent_qty */
  int64_t start_ns, stop_ns;       /* Kernel.runner startTime (midnight) / stopTime (horizon end + 10 min, :317-318) */
  int64_t mkt_open_ns, mkt_close_ns;
  int64_t horizon_start_ns, horizon_step_ns;  /* --start_hour, --freq (the agent hard-codes 30 s intervals, :373) */
  int64_t mom_wake_ns; int32_t mom_min_size, mom_max_size;   /* MomentumAgent wake_up_freq "20s", min_size 1, max_size 10 (:150-163) */
  int32_t stream_history;
  int32_t queue_cap, level_cap, order_cap;
  int32_t trace_cap, hash_pops;
} abx_dq_config;
int32_t abx_dq_config_default(abx_dq_config *cfg);
/* stream5 as abx_env_create. */
int32_t abx_dq_create(const abx_dq_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out);
/* Several days, as abx_env_create_days. */
int32_t abx_dq_create_days(const abx_dq_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out);
/* Agent construction + Kernel.runner start-up (Kernel.py:154-175).  seeds: HOST uint64 [n_envs] keying the MomentumAgent size draws
 * (random_state.randint(min_size, max_size), MomentumAgent.py:42); mom_sizes: optional HOST int32 [n_envs][n_momentum] that
 * overrides the draws (replay of a recorded reference run); either may be NULL (seed 0). */
int32_t abx_dq_reset(abx_sim *h, const uint64_t *seeds, const int32_t *mom_sizes, void *stream);
/* Baseline execution agent `k` (0 .. n_twap-1) becomes a VWAPExecutionAgent (agent/execution/baselines/vwap_agent.py:48-62): qty HOST int32 [n] are its schedule's
 * per-bin child quantities, round(volume_profile[bin] * quantity) for the n_horizon - 1 bins of the horizon (what generate_schedule builds; the agent class differs from the
 * TWAP agent in nothing else).  Entries < 0 and bins beyond n keep the TWAP quantity.  Takes effect from the next abx_dq_reset / step; same schedule in every environment. */
int32_t abx_dq_set_schedule(abx_sim *h, int32_t k, const int32_t *qty, int32_t n);
/* One decision step for every environment.  DEVICE pointers: actions int32 [n_envs] in 0..23 (ACTIONS, :24-37; ignored by
 * environments with no decision pending, i.e. on the first call), obs fp64 [n_envs][8] = the 6 observation features (:332) + the 2
 * digitised state entries the network sees (:334, util.py:23-42), trans fp64 [n_envs][6] = the finalised experience entry of the
 * previous tick (s0, s1, a, s'0, s'1, r; r NaN == None; all NaN before the first tick), reward fp64 [n_envs] = sum of the
 * step_reward_hist entries since the previous decision (:535), done uint8 [n_envs] (the event loop ended). */
int32_t abx_dq_step(abx_sim *h, const int32_t *actions_dev, double *obs_dev, double *trans_dev, double *reward_dev, uint8_t *done_dev, void *stream);
int32_t abx_dq_step_host(abx_sim *h, const int32_t *actions, double *obs, double *trans, double *reward, uint8_t *done, void *stream);
/* Final state of one environment.  out: HOST int64 [(n_agents-1) * 5] rows (agent id, shares, cash, last_trade, open orders or -1);
 * exec_out: HOST double [n_exec * 5] rows (remaining quantity, arrival price, executed orders, remaining_time, t). */
int32_t abx_dq_holdings(abx_sim *h, int32_t env, int64_t *out, double *exec_out, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batched Q-network forward of the DDQN agent on the tensor cores (tcgen05 / TMEM; csrc/abx_qnet.cu).
 * Replaces: eval_model.predict(s_array_2d) + np.argmax / the epsilon branch of choose_action
 * (agent/execution/qlearning/ddqlearning_execution_agent.py:339-365) and the network of util/model/QNets.py:7-27,55-60
 * (Dense + ReLU hidden layers, linear output; Dropout is the identity at inference), for a whole batch of states per launch.
 * dims: n_layers + 1 sizes (reference: 2, 32, 64, 128, 128, 64, 32, 24); dims[0] <= 16, every other size <= 128, n_layers <= 8.
 * params: HOST fp32, per layer W[out][in] row major (the transpose of a Keras Dense kernel) followed by b[out].
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct abx_qnet abx_qnet; /* opaque */
const char *abx_qnet_last_error(void);
int32_t abx_qnet_param_count(const int32_t *dims, int32_t n_layers);
int32_t abx_qnet_create(const int32_t *dims, int32_t n_layers, const float *params, int32_t device, abx_qnet **out);
/* New weights (the learner's update / the target-network sync of train_neural_nets :486-490); synchronises `stream`. */
int32_t abx_qnet_set_params(abx_qnet *q, const float *params, void *stream);
/* Same from a DEVICE fp32 vector (the learner's parameters), asynchronous on `stream`: the hi/lo operand image is rebuilt by a kernel. */
int32_t abx_qnet_set_params_device(abx_qnet *q, const float *params_dev, void *stream);
int32_t abx_qnet_destroy(abx_qnet *q);
int64_t abx_qnet_launch_count(const abx_qnet *q);
/* x_dev: DEVICE fp64 [n][x_stride]; the state of row i is x_dev[i * x_stride + x_offset ...+ dims[0]) (abx_dq_step's obs with
 * x_stride 8, x_offset 6).  q_out_dev: DEVICE fp32 [n][dims[n_layers]] or NULL; action_out_dev: DEVICE int32 [n] or NULL: the
 * first maximum of the Q row (np.argmax) with probability greedy_prob, otherwise uniform over the actions (Philox keyed by
 * (seed, counter, row); greedy_prob >= 1 draws nothing). */
int32_t abx_qnet_forward(abx_qnet *q, const double *x_dev, int32_t x_stride, int32_t x_offset, int32_t n, float *q_out_dev,
                         int32_t *action_out_dev, double greedy_prob, uint64_t seed, uint64_t counter, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Book surface: n_envs bare limit-order books driven by an operation tape recorded at the reference's
 * exchange boundary (what agent/ExchangeAgent.py:311,324,339 passes to OrderBook.handleLimitOrder / cancelOrder / modifyOrder,
 * util/OrderBook.py:38,284,341) -- "order streams recorded from the reference are replayed through the GPU books".
 * Handles are abx_sim*: after a replay abx_sim_trace returns, per operation, every notification the book sent (tag 1: ORDER_EXECUTED
 * pairs, ORDER_ACCEPTED, ORDER_CANCELLED, ORDER_MODIFIED; the reference's owner.sendMessage calls) followed by the book state (tag 2:
 * level counts, resting orders, three best levels per side, last_trade); abx_sim_book_snapshot is getInsideBids/Asks(depth);
 * abx_sim_stats has the counters.
 * ------------------------------------------------------------------------------------------------------------ */
int32_t abx_book_create(int32_t stream_history, int32_t level_cap, int32_t order_cap, int32_t trace_cap, int32_t n_envs, int32_t device, abx_sim **out);
/* ops9: HOST int64 [n_ops][9] rows (t_ns, op: 0 handleLimitOrder / 1 cancelOrder / 2 modifyOrder, agent id, order_id, is_buy, limit
 * price, quantity, new price, new quantity) -- the rows tools/record_reference.py records.  Every book of the handle replays the tape
 * (from an empty book; a second call continues on the current books).  A modify row with order_id 0 is a no-op: in the reference its
 * new_order carries a freshly generated id and isSameOrder fails (util/order/Order.py:27, util/OrderBook.py:343). */
int32_t abx_book_replay(abx_sim *h, const int64_t *ops9, int64_t n_ops, void *stream);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t abx_sim_launch_count(const abx_sim *h);

#ifdef __cplusplus
}
#endif
#endif /* ABIDES_B200_H */
