// abx_host_common.h -- host-side helpers shared by the CUDA C-ABI (abx_capi.cu) and the host emulation harness
// (tests/emu): config presets mirroring config/sparse_zi_100.py / config/sparse_zi_1000.py, validation, and the
// derived constants that the reference computes with libm on the host side of every call.
#pragma once
#include <math.h>
#include <string.h>
#include <vector>
#include <unordered_map>
#include <algorithm>
#include <utility>
#include "abx_core.cuh"

namespace abx {

static const int64_t NS = 1000000000LL;

// config/sparse_zi_1000.py:196-204 and config/sparse_zi_100.py:204-212: (n, R_min, R_max, eta)
static inline int config_sparse_zi(int variant, abx_sim_config *c) {
  if (!c || (variant != 100 && variant != 1000)) return ABX_ERR_ARG;
  memset(c, 0, sizeof(*c));
  static const int n1000[7] = {143, 143, 143, 143, 143, 143, 142}, n100[7] = {15, 15, 14, 14, 14, 14, 14};
  static const int rmin[7] = {0, 0, 0, 0, 0, 250, 250}, rmax[7] = {250, 500, 1000, 1000, 2000, 500, 500};
  static const double eta[7] = {1, 1, 0.8, 1, 0.8, 0.8, 1};
  c->version = ABX_VERSION; c->n_groups = 7; c->q_max = 10; c->n_agents = 1;
  for (int g = 0; g < 7; g++) {
    c->groups[g].count = variant == 1000 ? n1000[g] : n100[g]; c->groups[g].r_min = rmin[g]; c->groups[g].r_max = rmax[g]; c->groups[g].eta = eta[g];
    c->n_agents += c->groups[g].count;
  }
  c->start_ns = 0; c->stop_ns = 17 * 3600 * NS;                               // :86-88 midnight .. 17:00
  c->mkt_open_ns = (9 * 3600 + 30 * 60) * NS; c->mkt_close_ns = 16 * 3600 * NS; // :163-164
  c->default_computation_delay_ns = NS;                                        // :90 one second
  c->exchange_computation_delay_ns = 0; c->exchange_pipeline_delay_ns = 0;    // :186-187
  c->starting_cash = 10000000; c->order_size = 100; c->stream_history = 10;
  c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4;                    // :130-142
  c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;
  c->sigma_n = 1000000.0; c->agent_kappa = 1.67e-15; c->sigma_s = 1e-4; c->sigma_pv = 5e6; c->lambda_a = 1e-12; // :232-250
  if (variant == 1000) {                                                      // :264-286 old latency model
    c->latency_model = ABX_LAT_MATRIX_NOISE; c->n_noise = 6; c->latency_mirrored = 1; c->latency_lo = 21000; c->latency_hi = 13000000;
    c->queue_cap = 2240; c->level_cap = 428; c->order_cap = 2048;   // measured maxima over 128 seeds: queue 2 005, levels 357 / 349 (mean 311, sd ~15), resting 716
  } else {                                                                    // sparse_zi_100.py:305-318 cubic model
    c->latency_model = ABX_LAT_CUBIC; c->n_noise = 1; c->latency_mirrored = 0; c->latency_lo = 21000; c->latency_hi = 100000;
    c->jitter = 0.3; c->jitter_clip = 0.05; c->jitter_unit = 5.0;
    c->queue_cap = 384; c->level_cap = 128; c->order_cap = 512;
  }
  c->rng_mode = ABX_RNG_PHILOX; c->trace_cap = 0; c->hash_pops = 0;
  return ABX_OK;
}

// config/rmsc03.py:49-232
static inline int config_rmsc03(abx_sim_config *c) {
  if (!c) return ABX_ERR_ARG;
  memset(c, 0, sizeof(*c));
  c->version = ABX_VERSION; c->population = 1; c->n_noise_agents = 50; c->n_value_agents = 10; c->n_mm_agents = 1; c->n_momentum_agents = 2; c->n_agents = 64; c->n_groups = 0; c->q_max = 10;
  c->mkt_open_ns = (9 * 3600 + 1800) * NS; c->mkt_close_ns = (9 * 3600 + 2700) * NS;       // :69-70 09:30 .. 09:45
  c->start_ns = c->mkt_open_ns; c->stop_ns = c->mkt_close_ns + 60 * NS;                      // :205-207
  c->default_computation_delay_ns = 0; c->exchange_computation_delay_ns = 0; c->exchange_pipeline_delay_ns = 0;
  c->starting_cash = 10000000; c->order_size = 0; c->stream_history = 10;
  c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4; c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;
  c->sigma_n = 1e5 / 10; c->agent_kappa = 1.67e-15; c->sigma_s = 100000; c->sigma_pv = 0; c->lambda_a = 7e-11;   // :77-80; sigma_s: ValueAgent default
  c->latency_model = ABX_LAT_ZERO; c->n_noise = 1;                                         // np.zeros latency, noise [0.0] :209-210
  c->size_lo = 20; c->size_hi = 50; c->value_depth_spread = 2; c->value_percent_aggr = 0.1;   // NoiseAgent.py:34, ValueAgent.py:53-56
  c->noise_wake_lo_ns = 9 * 3600 * NS; c->noise_wake_hi_ns = 16 * 3600 * NS;                // :115-116
  c->mom_min_size = 1; c->mom_max_size = 10; c->mom_wake_ns = 20 * NS;                      // :190-192
  c->mm_pov = 0.05; c->mm_min_order_size = 20; c->mm_window_size = 5; c->mm_num_ticks = 20; c->mm_wake_ns = NS;   // :41-45
  c->queue_cap = 256; c->level_cap = 128; c->order_cap = 512; c->rng_mode = ABX_RNG_PHILOX; c->trace_cap = 0; c->hash_pops = 0;
  return ABX_OK;
}
// rmsc03 population + one POVExecutionAgent (BASELINE.json configs[2]; parameters in the style of config/execution_iabs_plots.py:200-226, scaled to the 15-minute session)
static inline int config_rmsc03_pov(abx_sim_config *c) {
  int st = config_rmsc03(c); if (st != ABX_OK) return st;
  c->n_pov_exec = 1; c->n_agents += 1; c->pov_exec_is_buy = 1; c->pov_exec_pov = 0.5; c->pov_exec_quantity = 120000;
  c->pov_exec_start_ns = (9 * 3600 + 32 * 60) * NS; c->pov_exec_end_ns = (9 * 3600 + 43 * 60) * NS; c->pov_exec_freq_ns = 30 * NS; c->pov_exec_lookback_ns = 30 * NS;
  return ABX_OK;
}
// config/rmsc01.py:60-262: 1 MarketMakerAgent, 50 ZeroIntelligenceAgents, 25 HeuristicBeliefLearningAgents (L = 2), 24 MomentumAgents, 09:30-16:00
static inline int config_rmsc01(abx_sim_config *c) {
  if (!c) return ABX_ERR_ARG;
  memset(c, 0, sizeof(*c));
  c->version = ABX_VERSION; c->population = 3; c->n_mm_agents = 1; c->n_groups = 2; c->q_max = 10; c->n_momentum_agents = 24;
  c->groups[0].count = 50; c->groups[0].r_min = 0; c->groups[0].r_max = 100; c->groups[0].eta = 1;                  // ZI :134-157
  c->groups[1].count = 25; c->groups[1].r_min = 0; c->groups[1].r_max = 100; c->groups[1].eta = 1; c->hbl_L = 2;   // HBL :163-187
  c->n_agents = 1 + 1 + 50 + 25 + 24;
  c->mkt_open_ns = (9 * 3600 + 1800) * NS; c->mkt_close_ns = 16 * 3600 * NS; c->start_ns = c->mkt_open_ns; c->stop_ns = (16 * 3600 + 60) * NS;   // :72-73, :246-247
  c->default_computation_delay_ns = 0; c->exchange_computation_delay_ns = 0; c->exchange_pipeline_delay_ns = 0;     // :85-86, :249
  c->starting_cash = 10000000; c->order_size = 100; c->stream_history = 10;
  c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4; c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;
  c->sigma_n = 10000; c->agent_kappa = 1.67e-15; c->sigma_s = 1e-4; c->sigma_pv = 5e4; c->lambda_a = 1e-12;          // :141-153
  c->latency_model = ABX_LAT_ZERO; c->n_noise = 1;                                                                  // :250-251
  c->mom_min_size = 1; c->mom_max_size = 10; c->mom_wake_ns = 60 * NS;                                              // MomentumAgent defaults, wake_up_freq "60s"
  c->mkm_min_size = 500; c->mkm_max_size = 1000; c->mkm_num_levels = 5; c->mkm_wake_ns = NS;                        // MarketMakerAgent :100-110 + defaults
  c->queue_cap = 512; c->level_cap = 256; c->order_cap = 2048; c->hist_log_cap = 32768;                             // 09:30-09:45 of seed 123456789 peaks at 31 levels a side, 67 resting orders
  c->rng_mode = ABX_RNG_PHILOX; c->trace_cap = 0; c->hash_pops = 0;
  return ABX_OK;
}
// config/rmsc02.py: rmsc01 with the market maker and the momentum agents in subscription mode, the sparse_zi_1000 latency matrix + noise, midnight .. 17:00
static inline int config_rmsc02(abx_sim_config *c) {
  int st = config_rmsc01(c); if (st != ABX_OK) return st;
  c->mkm_subscribe = 1; c->mom_subscribe = 1; c->mkm_sub_freq_ns = 10 * NS; c->mom_sub_freq_ns = 10 * NS;   // :108,205; subscribe_freq = 10e9, MomentumAgent.py:58
  c->start_ns = 0; c->stop_ns = 17 * 3600 * NS;                                                              // :264-265
  c->latency_model = ABX_LAT_MATRIX_NOISE; c->n_noise = 6; c->latency_mirrored = 0; c->latency_lo = 21000; c->latency_hi = 13000000;   // :268-269
  c->hist_log_cap = 4096;                                                                                   // the whole day is ~32 000 book operations with ~4 000 fills: buckets are short
  return ABX_OK;
}
static inline int config_validate(const abx_sim_config *c) {
  if (!c || c->version != ABX_VERSION) return ABX_ERR_ARG;
  if (c->n_agents < 2 || c->n_agents > 32767 || c->n_groups < 0 || c->n_groups > 8 || c->q_max < 1 || c->q_max > 10) return ABX_ERR_ARG;
  int n = 1;
  if (c->population == 0) { if (c->n_groups < 1) return ABX_ERR_ARG; for (int g = 0; g < c->n_groups; g++) { if (c->groups[g].count < 0 || c->groups[g].r_max < c->groups[g].r_min) return ABX_ERR_ARG; n += c->groups[g].count; } }
  else if (c->population == 1) {
    if (c->n_noise_agents < 0 || c->n_value_agents < 0 || c->n_mm_agents < 0 || c->n_mm_agents > 1 || c->n_momentum_agents < 0 || c->n_momentum_agents > 4) return ABX_ERR_ARG;
    if (c->latency_model != ABX_LAT_ZERO || c->size_hi <= c->size_lo || c->mom_max_size <= c->mom_min_size || 2 * (c->mm_num_ticks + 1) > MM_ORDER_CAP / 2 || c->mm_wake_ns <= 0 || c->mom_wake_ns <= 0) return ABX_ERR_ARG;
    if (c->n_pov_exec < 0 || c->n_pov_exec > 1 || (c->n_pov_exec && (!(c->pov_exec_pov > 0) || c->pov_exec_quantity <= 0 || c->pov_exec_quantity > 0x3fffffffLL || c->pov_exec_freq_ns <= 0 || c->pov_exec_lookback_ns <= 0))) return ABX_ERR_ARG;
    if (c->exec_kind < 0 || c->exec_kind > 2 || c->exec_limit_price < 0 || (c->exec_kind && (!c->n_pov_exec || c->pov_exec_start_ns < c->mkt_open_ns))) return ABX_ERR_ARG;
    n += c->n_noise_agents + c->n_value_agents + c->n_mm_agents + c->n_momentum_agents + c->n_pov_exec;
  } else if (c->population == 3) {
    if (c->n_groups != 2 || c->n_mm_agents < 0 || c->n_mm_agents > 1 || c->n_momentum_agents < 0 || c->n_momentum_agents > 60 || (c->latency_model != ABX_LAT_ZERO && c->latency_model != ABX_LAT_MATRIX_NOISE)) return ABX_ERR_ARG;
    if ((c->mkm_subscribe && (c->mkm_sub_freq_ns < 0 || c->mkm_num_levels > SUB_LEVELS)) || (c->mom_subscribe && c->mom_sub_freq_ns < 0)) return ABX_ERR_ARG;
    // one MARKET_DATA body per subscriber is in flight at a time (its levels wait in the subscriber's snapshot slot): the update period must exceed the longest delivery
    if (c->latency_model != ABX_LAT_ZERO && ((c->mkm_subscribe && (double)c->mkm_sub_freq_ns <= c->latency_hi + c->n_noise) || (c->mom_subscribe && (double)c->mom_sub_freq_ns <= c->latency_hi + c->n_noise))) return ABX_ERR_ARG;
    if (c->latency_model == ABX_LAT_ZERO && ((c->mkm_subscribe && c->mkm_sub_freq_ns == 0) || (c->mom_subscribe && c->mom_sub_freq_ns == 0))) return ABX_ERR_ARG;
    for (int g = 0; g < 2; g++) { if (c->groups[g].count < 0 || c->groups[g].r_max < c->groups[g].r_min) return ABX_ERR_ARG; n += c->groups[g].count; }
    if (c->hbl_L < 1 || c->hbl_L > 16 || c->stream_history < 1 || c->hist_log_cap < 64 || c->hist_log_cap > 65536 || (c->hist_log_cap & (c->hist_log_cap - 1)) || c->hbl_table_rows < 0 || c->hbl_table_rows > c->hist_log_cap / 4) return ABX_ERR_ARG;
    if (c->mom_max_size <= c->mom_min_size || c->mom_wake_ns <= 0) return ABX_ERR_ARG;
    if (c->n_mm_agents && (c->mkm_max_size <= c->mkm_min_size || c->mkm_num_levels < 1 || 4 * c->mkm_num_levels > MM_ORDER_CAP / 2 || c->mkm_wake_ns <= 0)) return ABX_ERR_ARG;
    n += c->n_mm_agents + c->n_momentum_agents;
  } else return ABX_ERR_ARG;
  if (n != c->n_agents) return ABX_ERR_ARG;
  if (c->queue_cap < 32 || c->queue_cap % 32 || c->queue_cap > 4096) return ABX_ERR_ARG;
  if (c->level_cap < 8 || c->level_cap > 2048 || c->level_cap % 4 || c->order_cap < 8 || c->order_cap > 65535) return ABX_ERR_ARG;   // level_cap % 4: the ladders are searched with 128-bit shared-memory loads
  if (c->stop_ns >= KEY_T_MAX || c->start_ns < 0 || c->mkt_close_ns >= KEY_T_MAX) return ABX_ERR_ARG;
  if (c->latency_model != ABX_LAT_MATRIX_NOISE && c->latency_model != ABX_LAT_CUBIC && c->latency_model != ABX_LAT_ZERO) return ABX_ERR_ARG;
  if (c->latency_model == ABX_LAT_MATRIX_NOISE && c->n_noise < 1) return ABX_ERR_ARG;
  if (c->rng_mode != ABX_RNG_PHILOX && c->rng_mode != ABX_RNG_TAPE) return ABX_ERR_ARG;
  if (c->trace_cap < 0 || !(c->kappa > 0) || !(c->lambda_a > 0) || !(c->megashock_lambda_a > 0)) return ABX_ERR_ARG;
  return ABX_OK;
}

// Entries of the transaction-tuple ring behind get_transacted_volume (util/OrderBook.py:400-436): the tuples of stream_history + 1 surviving history
// buckets; 64 per bucket (an order sweeping more than that many resting orders between two trades raises ABX_F_HISTORY_OVERFLOW), a power of two >= 512.
static inline int tv_ring_for(int stream_history) { int need = 64 * (stream_history + 2), r = TV_RING_MIN; while (r < need) r <<= 1; return r; }
// Constants the reference evaluates with CPython/libm on every call; evaluated once here with the same libm.
static inline void derive_params(SimParams &P) {
  const abx_sim_config &c = P.c;
  P.n_qgroups = c.queue_cap / 32; P.n_streams = c.n_agents + 3;
  P.one_minus_kappa_a = 1 - c.agent_kappa;
  P.log_base_a = log(P.one_minus_kappa_a);                         // exact argument: 1 - kappa as CPython rounds it
  P.sigma_denom = 1 - pow(1 - c.agent_kappa, 2.0);                 // ZeroIntelligenceAgent.py:234
  P.sqrt_sigma_n = sqrt(c.sigma_n); P.sqrt_sigma_pv = sqrt(c.sigma_pv); P.sqrt_megashock_var = sqrt(c.megashock_var);
  P.inv_lambda_a = 1.0 / c.lambda_a; P.inv_megashock_lambda = 1.0 / c.megashock_lambda_a;
  P.ou_scale = pow(c.fund_vol, 2.0) / (2 * c.kappa);               // SparseMeanRevertingOracle.py:106
  P.tv_ring = tv_ring_for(c.stream_history);
}

// ---- ABIDESEnv shape ----
static inline int env_config_default(abx_env_config *c) {                       // ABIDESEnv.py:59-103, agent_config.py:42-154
  if (!c) return ABX_ERR_ARG;
  memset(c, 0, sizeof(*c));
  c->version = ABX_VERSION; c->order_level = 2; c->is_buy = 1; c->n_horizon = 761;
  c->start_ns = 0; c->stop_ns = (16 * 3600 + 600) * NS; c->mkt_open_ns = (9 * 3600 + 1800) * NS; c->mkt_close_ns = 16 * 3600 * NS;
  c->horizon_start_ns = (9 * 3600 + 2400) * NS; c->horizon_step_ns = 30 * NS; c->quantity = 1e5; c->steep = 0.5;
  c->stream_history = 10; c->queue_cap = 64; c->level_cap = 256; c->order_cap = 16384; c->trace_cap = 0; c->hash_pops = 0;
  return ABX_OK;
}
static inline int env_config_validate(const abx_env_config *c) {
  if (!c || c->version != ABX_VERSION || c->order_level < 0 || c->order_level > 2 || c->n_horizon < 2) return ABX_ERR_ARG;
  if (c->queue_cap < 32 || c->queue_cap % 32 || c->queue_cap > 4096 || c->level_cap < 8 || c->level_cap > 2048 || c->level_cap % 4) return ABX_ERR_ARG;
  if (c->order_cap < 8 || c->order_cap > 65535 || c->stop_ns >= KEY_T_MAX || c->start_ns < 0 || c->horizon_step_ns <= 0) return ABX_ERR_ARG;
  if (c->stream_history < 0 || c->stream_history > 14 || !(c->quantity > 0) || c->trace_cap < 0) return ABX_ERR_ARG;
  return ABX_OK;
}
// ABIDESEnv shape proper (abx_env_create*): the on-chip event queue of that shape always spans ENV_QUEUE_MIN slots per environment, so a smaller
// queue_cap would overrun the next environment's slots (the bare-book surface uses the grouped queue and may go down to 32)
constexpr int ENV_QUEUE_MIN = 64;
static inline int env_shape_validate(const abx_env_config *c) { int st = env_config_validate(c); if (st != ABX_OK) return st; return c->queue_cap < ENV_QUEUE_MIN ? ABX_ERR_ARG : ABX_OK; }
// the generic (abx_sim_config) part of the parameter block for the ABIDESEnv shape: 3 agents, zero delays, no oracle
static inline void env_fill_params(const abx_env_config &e, SimParams &P) {
  abx_sim_config &c = P.c; memset(&c, 0, sizeof(c));
  c.version = ABX_VERSION; c.n_agents = e.order_level > 0 ? 3 : 2; c.n_groups = 1; c.q_max = 1; c.groups[0].count = c.n_agents - 1;
  c.start_ns = e.start_ns; c.stop_ns = e.stop_ns; c.mkt_open_ns = e.mkt_open_ns; c.mkt_close_ns = e.mkt_close_ns;
  c.default_computation_delay_ns = 0; c.exchange_computation_delay_ns = 0; c.exchange_pipeline_delay_ns = 0;   // ABIDESEnv.py:89, agent_config.py:49-50
  c.stream_history = e.stream_history; c.latency_model = ABX_LAT_ZERO; c.n_noise = 1;
  c.queue_cap = e.queue_cap; c.level_cap = e.level_cap; c.order_cap = e.order_cap; c.rng_mode = ABX_RNG_PHILOX; c.trace_cap = e.trace_cap; c.hash_pops = e.hash_pops;
  P.n_qgroups = c.queue_cap / 32; P.n_streams = 0;
  P.n_h = e.n_horizon; P.h0_ns = e.horizon_start_ns; P.h_step_ns = e.horizon_step_ns; P.h_step_inv = e.horizon_step_ns > 0 ? 1.0 / (double)e.horizon_step_ns : 0.0; P.rl_quantity = e.quantity; P.rl_steep = e.steep;
  P.order_level = e.order_level; P.rl_is_buy = e.is_buy;
}
// ---- DDQN execution shape (config/execution/marketreplay/execution_marketreplay_ddqn.py) ----
static inline int dq_config_default(abx_dq_config *c) {
  if (!c) return ABX_ERR_ARG;
  memset(c, 0, sizeof(*c));
  c->version = ABX_VERSION; c->n_momentum = 7; c->n_twap = 1; c->has_ddqn = 1; c->is_buy = 1; c->n_horizon = 661; c->quantity = 500000;   // :140-258, scripts: BUY 5e5 from 10:00 over 330 min at 30 s
  c->start_ns = 0; c->mkt_open_ns = (9 * 3600 + 1800) * NS; c->mkt_close_ns = 16 * 3600 * NS;                                     // :93-97
  c->horizon_start_ns = 10 * 3600 * NS; c->horizon_step_ns = 30 * NS; c->stop_ns = c->horizon_start_ns + 660 * c->horizon_step_ns + 600 * NS;   // :317-318
  c->mom_wake_ns = 20 * NS; c->mom_min_size = 1; c->mom_max_size = 10; c->stream_history = 10;
  c->queue_cap = 1536;             // the closing market order walks up to DQ_DEPTH levels: <= 500 LIMIT_ORDERs in flight + 2 ORDER_EXECUTED each (recorded maximum 1 066)
  c->level_cap = 256; c->order_cap = 16384; c->trace_cap = 0; c->hash_pops = 0;
  return ABX_OK;
}
static inline int dq_config_validate(const abx_dq_config *c) {
  if (!c || c->version != ABX_VERSION || c->n_momentum < 0 || c->n_momentum > 8 || c->n_twap < 0 || c->n_twap > 2 || c->n_horizon < 3) return ABX_ERR_ARG;
  if (c->queue_cap < 32 || c->queue_cap % 32 || c->queue_cap > 4096 || c->level_cap < 8 || c->level_cap > 2048 || c->level_cap % 4) return ABX_ERR_ARG;
  if (c->order_cap < 8 || c->order_cap > 65535 || c->stop_ns >= KEY_T_MAX || c->start_ns < 0 || c->horizon_step_ns <= 0 || c->mom_wake_ns <= 0) return ABX_ERR_ARG;
  if (c->queue_cap < 128) return ABX_ERR_ARG;                            // two-tier queue: the first two groups' slots belong to the on-chip tier
  if (c->stream_history < 0 || c->stream_history > 14 || c->quantity <= 0 || c->quantity > 0x3fffffffLL || c->trace_cap < 0 || c->mom_max_size <= c->mom_min_size) return ABX_ERR_ARG;
  return ABX_OK;
}
static inline void dq_fill_params(const abx_dq_config &e, SimParams &P) {
  abx_sim_config &c = P.c; memset(&c, 0, sizeof(c));
  c.version = ABX_VERSION; c.population = 2; c.n_momentum_agents = e.n_momentum; c.n_mm_agents = e.has_ddqn ? 1 : 0;   // population 2: n_mm_agents == "has a DDQN agent" (agent_type_of)
  c.n_agents = 2 + e.n_momentum + e.n_twap + (e.has_ddqn ? 1 : 0); c.n_groups = 0; c.q_max = 1;
  c.start_ns = e.start_ns; c.stop_ns = e.stop_ns; c.mkt_open_ns = e.mkt_open_ns; c.mkt_close_ns = e.mkt_close_ns;
  c.default_computation_delay_ns = 0; c.exchange_computation_delay_ns = 0; c.exchange_pipeline_delay_ns = 0; c.starting_cash = 0;           // :113-114,327 ; starting_cash=0 :141,163,...
  c.stream_history = e.stream_history; c.latency_model = ABX_LAT_ZERO; c.n_noise = 1; c.mom_wake_ns = e.mom_wake_ns; c.mom_min_size = e.mom_min_size; c.mom_max_size = e.mom_max_size;
  c.queue_cap = e.queue_cap; c.level_cap = e.level_cap; c.order_cap = e.order_cap; c.rng_mode = ABX_RNG_PHILOX; c.trace_cap = e.trace_cap; c.hash_pops = e.hash_pops;
  P.n_qgroups = c.queue_cap / 32; P.n_streams = 0;
  P.n_h = e.n_horizon; P.h0_ns = e.horizon_start_ns; P.h_step_ns = e.horizon_step_ns; P.h_step_inv = e.horizon_step_ns > 0 ? 1.0 / (double)e.horizon_step_ns : 0.0; P.rl_quantity = (double)e.quantity; P.rl_steep = 0.5; P.order_level = 0; P.rl_is_buy = e.is_buy;
  P.dq_n_mom = e.n_momentum; P.dq_n_twap = e.n_twap; P.dq_has_ddqn = e.has_ddqn ? 1 : 0; P.dq_quantity = e.quantity;
}
// ids the generator can hand out in one run: momentum orders + execution-agent orders (market orders walk <= DQ_DEPTH levels)
static inline int64_t dq_max_generated_ids(const abx_dq_config &e) {
  int64_t wakes = (e.stop_ns - e.mkt_open_ns) / e.mom_wake_ns + 2;
  return (int64_t)e.n_momentum * wakes + (int64_t)(e.n_twap + 1) * ((int64_t)e.n_horizon * 8 + 2 * DQ_DEPTH) + 16;
}
struct EnvStreamHost { std::vector<int64_t> ts, id_orig; std::vector<int32_t> first, xid, xfirst; std::vector<int4> rows; int64_t min_id, max_id; };   // xid: the explicit ORDER_IDs sorted, xfirst: first row of each
// LOBSTER ORDER_IDs -> dense indices; rows grouped by identical timestamp (orders_dict of MarketReplayAgent.py:214)
static inline int env_build_stream(const int64_t *s5, int64_t n, int64_t max_rl_ids, EnvStreamHost &o) {
  if (!s5 || n < 1 || n > 0x3fffffff) return ABX_ERR_ARG;
  std::unordered_map<int64_t, int32_t> dense; int64_t n_zero = 0, min_id = 0x7fffffffffffLL;
  o.rows.resize(n);
  for (int64_t i = 0; i < n; i++) {
    const int64_t *r = s5 + 5 * i;
    if (i > 0 && r[0] < s5[5 * (i - 1)]) return ABX_ERR_ARG;                                     // must be time sorted
    if (r[1] < 0 || r[1] > 0x3fffffffLL || r[2] < 0 || r[2] > 0x3fffffffLL || r[3] < 0 || r[3] > 0x7fffffffLL) return ABX_ERR_ARG;
    int32_t d;
    if (r[1] == 0) { d = -1; n_zero++; }                                                          // ORDER_ID 0: "unset", the agent gets generated ids
    else { if (r[1] < min_id) min_id = r[1];
      auto it = dense.find(r[1]);
      if (it == dense.end()) { d = (int32_t)o.id_orig.size(); dense.emplace(r[1], d); o.id_orig.push_back(r[1]); } else d = it->second; }
    if (i == 0 || r[0] != s5[5 * (i - 1)]) { o.ts.push_back(r[0]); o.first.push_back((int32_t)i); }
    int4 row; row.x = d; row.y = (int32_t)r[2]; row.z = (int32_t)r[3]; row.w = r[4] ? 1 : 0; o.rows[i] = row;
  }
  o.first.push_back((int32_t)n); o.min_id = min_id; o.max_id = 0; (void)max_rl_ids; (void)n_zero;
  // util/order/Order.py:35-42: generateOrderId skips every id already used, explicit ORDER_IDs of the stream included once their first row was replayed.
  // The device keeps the explicit ids sorted with the first row of each, so the generator can ask "has this id been used yet" (Sim::gen_id).
  { std::vector<std::pair<int64_t, int32_t>> v; v.reserve(dense.size());
    std::vector<int32_t> first_row(o.id_orig.size(), -1);
    for (int64_t i = 0; i < n; i++) { int32_t d = o.rows[i].x; if (d >= 0 && first_row[d] < 0) first_row[d] = (int32_t)i; }
    for (size_t d = 0; d < o.id_orig.size(); d++) v.push_back(std::make_pair(o.id_orig[d], first_row[d]));
    std::sort(v.begin(), v.end());
    for (auto &pr : v) { o.xid.push_back((int32_t)pr.first); o.xfirst.push_back(pr.second); if (pr.first > o.max_id) o.max_id = pr.first; } }
  if (o.id_orig.empty()) { o.id_orig.push_back(0); o.min_id = 0x7fffffffLL; }
  return ABX_OK;
}
// Book surface: recorded op rows -> device rows (order ids become dense device ids; a modify of order id 0 is a no-op, see abides_b200.h)
static inline int book_ops_to_device(const int64_t *ops9, int64_t n, std::unordered_map<int64_t, int32_t> &dense, std::vector<int64_t> &id_orig, std::vector<int64_t> &out) {
  out.assign(ops9, ops9 + 9 * n);
  for (int64_t i = 0; i < n; i++) {
    int64_t *r = out.data() + 9 * i;
    if (r[1] < 0 || r[1] > 2 || r[2] < 0 || r[2] > 0xffff || r[0] < 0 || r[0] >= KEY_T_MAX || r[5] > 0x3fffffffLL || r[6] > 0x7fffffffLL || r[5] < -0x3fffffffLL) return ABX_ERR_ARG;
    if (r[1] == 2 && r[3] == 0) { r[3] = 0xffffffffLL; continue; }
    auto it = dense.find(r[3]); int32_t d;
    if (it == dense.end()) { if (id_orig.size() >= 0x3fffffffu) return ABX_ERR_ARG; d = (int32_t)id_orig.size(); dense.emplace(r[3], d); id_orig.push_back(r[3]); } else d = it->second;
    r[3] = (int64_t)REPLAY_ID_BASE + d;
  }
  return ABX_OK;
}
// Several replayed days in one handle: environment e replays day e % n_days.  Each day has its own dense order-id space; the device arrays
// are the concatenation of the days' arrays plus one {first timestamp, timestamps, first `first` entry, -} record per day.
struct EnvDaysHost { std::vector<EnvStreamHost> days; std::vector<int64_t> ts; std::vector<int32_t> first, xid, xfirst; std::vector<int4> rows, day_tab, day_tab2; int max_ids; int64_t min_id; };
static inline int env_build_days(const int64_t *s5, const int64_t *row_off, int n_days, int64_t max_gen_ids, EnvDaysHost &o) {
  if (!s5 || !row_off || n_days < 1 || n_days > 4096) return ABX_ERR_ARG;
  o.days.resize(n_days); o.max_ids = 1; o.min_id = 0x7fffffffffffLL;
  for (int d = 0; d < n_days; d++) {
    int64_t n = row_off[d + 1] - row_off[d]; EnvStreamHost &sd = o.days[d];
    int st = env_build_stream(s5 + 5 * row_off[d], n, max_gen_ids, sd); if (st != ABX_OK) return st;
    if (o.rows.size() + (size_t)n > 0x3fffffffu) return ABX_ERR_ARG;
    int4 rec; rec.x = (int32_t)o.ts.size(); rec.y = (int32_t)sd.ts.size(); rec.z = (int32_t)o.first.size(); rec.w = 0; o.day_tab.push_back(rec);
    int4 r2; r2.x = (int32_t)o.xid.size(); r2.y = (int32_t)sd.xid.size(); r2.z = (int32_t)(sd.min_id > 0x7fffffffLL ? 0x7fffffffLL : sd.min_id); r2.w = (int32_t)sd.max_id; o.day_tab2.push_back(r2);
    o.xid.insert(o.xid.end(), sd.xid.begin(), sd.xid.end()); o.xfirst.insert(o.xfirst.end(), sd.xfirst.begin(), sd.xfirst.end());
    int32_t row_base = (int32_t)o.rows.size();
    o.ts.insert(o.ts.end(), sd.ts.begin(), sd.ts.end());
    for (int32_t f : sd.first) o.first.push_back(f + row_base);
    o.rows.insert(o.rows.end(), sd.rows.begin(), sd.rows.end());
    if ((int)sd.id_orig.size() > o.max_ids) o.max_ids = (int)sd.id_orig.size();
    if (sd.min_id < o.min_id) o.min_id = sd.min_id;
  }
  return ABX_OK;
}
ABX_HD void init_envx(const SimParams &P, EnvX &x) {
  x.ra_time = x.rl_time = P.c.start_ns; x.ra_cash = x.rl_cash = 0; x.rem_quantity = P.rl_quantity; x.executed_sum = 0.0;   // starting_cash 0 (agent_config.py:73,133)
  x.ra_shares = x.rl_shares = 0; x.ra_last_trade = x.rl_last_trade = 0; x.ra_flags = 0;
  x.rl_flags = RLF_TRADE | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT); x.wt_cursor = 0; x.n_executed = 0; x.rl_n_orders = 0; x.n_lobs = 0; x.lob_head = 0; x.p0 = 0;
  x.rem_time = P.n_h - 1; x.obs_len = 0; x.g0_qty = 0; x.steps = 0;
  for (int i = 0; i < RL_ORDER_CAP; i++) { x.rl_oid[i] = 0; x.rl_oprice[i] = 0; x.rl_oqty[i] = 0; }
  for (int i = 0; i < 9; i++) x.obs[i] = 0.0; x.g0_pq = 0; x.rows_done = 0;
}

// abx_dq_holdings rows from the trader records of one environment (agent/TradingAgent.py:124-126 final holdings)
static inline void dq_holdings_rows(const SimParams &P, const ZiAgent *ag, const EnvX &x, int64_t *out, double *exec_out) {
  int n = P.c.n_agents, ke = 0;
  for (int id = 1; id < n; id++) { int64_t *r = out + 5 * (id - 1); r[0] = id;
    if (id == 1) { r[1] = x.ra_shares; r[2] = x.ra_cash; r[3] = (x.ra_flags & AF_HAS_LAST) ? x.ra_last_trade : 0; r[4] = -1; continue; }
    const ZiAgent &z = ag[id]; int type = (int)((z.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT);
    r[1] = z.shares; r[2] = z.cash; r[3] = (z.flags & AF_HAS_LAST) ? z.last_trade : 0; r[4] = type == AT_MOMENTUM ? -1 : z.n_orders;
    if (type != AT_MOMENTUM && exec_out) { const ExecAux *ex = reinterpret_cast<const ExecAux *>(z.oid); double *e = exec_out + 5 * ke++;
      e[0] = ex->rem_qty; e[1] = (double)ex->arr2 / 2; e[2] = ex->n_executed; e[3] = ex->rem_time; e[4] = ex->t; } }
}

// abx_sim_agent_init rows from the trader records of one freshly reset environment
static inline void agent_init_rows(const SimParams &P, const ZiAgent *ag, int32_t *theta, double *lat_to, double *lat_from, int32_t *sizes, int64_t *wakes) {
  int n = P.c.n_agents;
  for (int id = 0; id < n; id++) {
    const ZiAgent &z = ag[id]; int type = id == 0 ? -1 : agent_type_of(P.c, id);
    if (theta) for (int i = 0; i < 20; i++) theta[(size_t)id * 20 + i] = (type == AT_ZI || type == AT_HBL) ? z.theta[i] : 0;
    if (lat_to) lat_to[id] = id ? z.lat_to : 0.0;
    if (lat_from) lat_from[id] = id ? z.lat_from : 0.0;
    if (sizes) sizes[id] = (type == AT_NOISE || type == AT_VALUE || type == AT_MOMENTUM || type == AT_MKM) ? reinterpret_cast<const AgentAux *>(z.theta)->size : 0;
    if (wakes) wakes[id] = type == AT_NOISE ? z.prev_wake : 0;
  }
}

static inline const char *status_string(int32_t st) {
  switch (st) {
    case ABX_OK: return "ok";
    case ABX_ERR_ARG: return "invalid argument or configuration";
    case ABX_ERR_CUDA: return "CUDA runtime error";
    case ABX_ERR_STATE: return "call sequence error";
    case ABX_ERR_CAPACITY: return "fixed-capacity structure overflowed";
    default: return "unknown status";
  }
}

// stats record from an EnvState plus the top of the two ladders
ABX_HD void fill_stats(const EnvState &s, int32_t bb, int32_t bbq, int32_t ba, int32_t baq, abx_env_stats *o) {
  o->messages = s.ttl; o->now_ns = s.now; o->pop_hash = s.pop_hash; o->limit_orders = s.c_limit; o->cancels = s.c_cancel; o->fills = s.c_fills;
  o->spread_queries = s.c_query; o->max_queue = s.max_q; o->n_bid_levels = s.n_bid_lv; o->n_ask_levels = s.n_ask_lv; o->n_resting = s.n_resting;
  o->best_bid = bb; o->best_bid_qty = bbq; o->best_ask = ba; o->best_ask_qty = baq; o->last_trade = s.last_trade; o->fundamental = s.or_v;
  o->flags = s.flags; o->trace_len = s.trace_n; o->uniq = s.uniq; o->orders_allocated = s.next_order_id; o->sum_shares = s.sum_shares; o->sum_cash = s.sum_cash;
}

}  // namespace abx
