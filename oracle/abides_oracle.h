/*
 * abides_oracle.h -- CPU restatement (plain C) of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product path (marl_optimal_execution_b200/) never links or calls it.
 *
 * Parity status: PINNED.  From the seed alone this restatement reproduces the reference's recorded
 * golden run tests/sparse_zi_1000.txt (185 200 messages, 1000 final-holdings lines) and the traces
 * recorded from the live reference by tools/record_reference*.py (tests/golden/ .npz files): event-queue pop
 * order, exchange-boundary ops, every outbound exchange message, book snapshots and every RNG draw -- for
 * sparse_zi_100 / sparse_zi_1000 (two seeds each), rmsc03 (two seeds) and rmsc03 with the reference's
 * POVExecutionAgent appended, two ABIDESEnv episodes, the GOOG market-replay day, and three runs of the DDQN
 * execution config (BUY, SELL, and one whose agent was driven by a real fp32 network).  NOT pinned: the
 * arithmetic of the Keras Q-network (TensorFlow is absent from the image; the oracle takes the action of
 * every decision tick as an input).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef ABIDES_ORACLE_H
#define ABIDES_ORACLE_H
#include <stdint.h>
#include "../include/abides_b200.h"   /* abx_sim_config only: the parameter struct both sides take, so parity tests change a field once for both */

#ifdef __cplusplus
extern "C" {
#endif

/* message kinds: msg.body["msg"] strings of the reference protocol */
enum abo_kind {
  ABO_NONE = 0, ABO_WHEN_MKT_OPEN, ABO_WHEN_MKT_CLOSE, ABO_QUERY_SPREAD, ABO_LIMIT_ORDER, ABO_CANCEL_ORDER,
  ABO_MODIFY_ORDER, ABO_ORDER_ACCEPTED, ABO_ORDER_EXECUTED, ABO_ORDER_CANCELLED, ABO_MKT_CLOSED,
  ABO_QUERY_LAST_TRADE, ABO_QUERY_TRANSACTED_VOLUME, ABO_ORDER_MODIFIED, ABO_QUERY_ORDER_STREAM, ABO_MARKET_DATA,
  ABO_MARKET_DATA_SUBSCRIPTION_REQUEST, ABO_MARKET_DATA_SUBSCRIPTION_CANCELLATION
};
/* queue entry types: message/Message.py:5-10 */
enum abo_type { ABO_T_MESSAGE = 1, ABO_T_WAKEUP = 2, ABO_T_CANCEL_ORDER = 3 };

enum abo_trace_flags { ABO_TRACE_POPS = 1, ABO_TRACE_OPS = 2, ABO_TRACE_NOTES = 4, ABO_TRACE_SNAPS = 8, ABO_TRACE_TAPES = 16 };

/* ---------------- numpy legacy RandomState (MT19937) ---------------- */
typedef struct abo_rng abo_rng;
abo_rng *abo_rng_new(uint32_t seed);
void abo_rng_free(abo_rng *);
uint32_t abo_rng_u32(abo_rng *);
double abo_rng_double(abo_rng *);                 /* random_sample() */
double abo_rng_gauss(abo_rng *);                  /* standard_normal() (legacy polar, cached 2nd variate) */
double abo_rng_std_exponential(abo_rng *);        /* standard_exponential() */
int64_t abo_rng_randint(abo_rng *, int64_t low, int64_t high); /* randint(low, high) masked rejection */

/* ---------------- standalone order book (util/OrderBook.py) ---------------- */
typedef struct abo_book abo_book;
/* notification rows written by the book: 13 int64 each, same layout tools/record_reference.py records:
 * (t, recipient, kind, order_id, is_buy, qty, limit_price, fill_price, 0,0,0,0,0) */
abo_book *abo_book_new(int stream_history);
void abo_book_free(abo_book *);
void abo_book_set_time(abo_book *, int64_t now_ns);
void abo_book_limit(abo_book *, int64_t agent, int64_t order_id, int is_buy, int64_t price, int64_t qty);
void abo_book_cancel(abo_book *, int64_t agent, int64_t order_id, int is_buy, int64_t price);
void abo_book_modify(abo_book *, int64_t agent, int64_t order_id, int is_buy, int64_t price,
                     int64_t new_order_id, int64_t new_price, int64_t new_qty);
int abo_book_inside(abo_book *, int is_bid, int depth, int64_t *out_price_qty /* 2*depth */);
int64_t abo_book_last_trade(abo_book *); /* -1 == None */
int64_t abo_book_transacted_volume(abo_book *, int64_t lookback_ns);
int abo_book_n_levels(abo_book *, int is_bid);
int abo_book_n_resting(abo_book *);
int64_t abo_book_n_notes(abo_book *);
const int64_t *abo_book_notes(abo_book *);
void abo_book_clear_notes(abo_book *);
/* level dump: for level i of a side writes (order_id, qty, limit_price) triples; returns count */
int abo_book_level_orders(abo_book *, int is_bid, int level, int64_t *out, int max_orders);

/* ---------------- full simulation: config/sparse_zi_100.py, config/sparse_zi_1000.py ---------------- */
typedef struct abo_sim abo_sim;
/* variant: 100 or 1000.  seed: the -s argument.  trace_flags: OR of abo_trace_flags. */
abo_sim *abo_sim_new_sparse_zi(int variant, uint32_t seed, int trace_flags);
/* config/rmsc03.py: 1 exchange + 50 NoiseAgents + 10 ValueAgents + 1 POVMarketMakerAgent + 2 MomentumAgents, 09:30-09:45 */
abo_sim *abo_sim_new_rmsc03(uint32_t seed, int trace_flags);
/* the same population plus one POVExecutionAgent (agent/execution/baselines/pov_agent.py) as the last agent; times ns since midnight */
abo_sim *abo_sim_new_rmsc03_pov(uint32_t seed, int trace_flags, double pov, int64_t quantity, int is_buy, int64_t start_ns, int64_t end_ns,
                                int64_t freq_ns, int64_t lookback_ns);
void abo_sim_pov_exec(abo_sim *, int64_t *out3); /* rem_quantity, executed orders, open orders */
/* The config scripts' numbers as an abx_sim_config (variant 100 / 1000 / 3 = rmsc03 / 4 = rmsc03 + POV execution agent), and a simulation built from
 * any such struct: the same seed cascade with other agent counts / parameters (parametrised parity). */
int abo_default_config(int variant, abx_sim_config *cfg);
abo_sim *abo_sim_new_config(const abx_sim_config *cfg, uint32_t seed, int trace_flags);
/* External tapes (see abides_oracle.c): standard variates per stream in the product's stream order + the start-of-run state drawn with them. */
int abo_sim_set_external(abo_sim *, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const int32_t *theta, const double *lat_to,
                         const double *lat_from, const int32_t *sizes, const int64_t *wakes);
int abo_sim_rng_error(abo_sim *);            /* 1 underrun | 2 kind mismatch on an external tape */
int64_t abo_sim_external_unread(abo_sim *);
void abo_sim_free(abo_sim *);
/* runtime draws on the GLOBAL np.random stream (kinds 'e','u','i') */
int64_t abo_sim_global_tape(abo_sim *, const uint8_t **kinds, const uint64_t **bits);
void abo_sim_agent_info(abo_sim *, int id, int64_t *out4); /* type, size, noise wakeup_time, MM order_size */
/* Kernel.runner (Kernel.py:50-345): start, event loop, kernelStopping.  Returns ttl_messages. */
int64_t abo_sim_run(abo_sim *);
/* Event loop only, until the next pop would be later than `until_ns` or the loop ends (used to compare
 * intermediate state with the GPU path).  Returns pops made by this call; *done set when the loop ended. */
int64_t abo_sim_run_until(abo_sim *, int64_t until_ns, int *done);
void abo_sim_start(abo_sim *);    /* kernelInitializing + kernelStarting (Kernel.py:154-175) */
void abo_sim_stop(abo_sim *);     /* kernelStopping (Kernel.py:310-311) */

int abo_sim_n_agents(abo_sim *);
int64_t abo_sim_n_pops(abo_sim *);
/* per trading agent (ids 1..n-1): rows (id, shares, cash, marked_to_market, surplus) */
void abo_sim_holdings(abo_sim *, int64_t *out5);
uint64_t abo_sim_pop_hash(abo_sim *);
int64_t abo_sim_n_hash_ckpt(abo_sim *);
const uint64_t *abo_sim_hash_ckpt(abo_sim *);
uint64_t abo_sim_note_hash(abo_sim *);
uint64_t abo_sim_snap_hash(abo_sim *);
/* traces (rows of int64): pops x5, ops x9, notes x13, snaps x16 -- layouts in tools/record_reference.py */
int64_t abo_sim_trace(abo_sim *, int which /*0 pops,1 ops,2 notes,3 snaps*/, const int64_t **rows);
/* RNG tapes: stream s in creation order (symbol, kernel, [latency model], exchange, agents...) */
int abo_sim_n_streams(abo_sim *);
int64_t abo_sim_tape(abo_sim *, int stream, const uint8_t **kinds, const uint64_t **bits);
uint32_t abo_sim_stream_seed(abo_sim *, int stream);
int64_t abo_sim_global_exp_tape(abo_sim *, const double **vals);
/* initial state the GPU path needs in tape mode */
void abo_sim_theta(abo_sim *, int agent, int32_t *out /* 2*q_max */);
void abo_sim_latency_vectors(abo_sim *, double *to_exchange, double *from_exchange); /* [n_agents] each */
void abo_sim_zi_params(abo_sim *, int agent, double *out4 /* R_min, R_max, eta, group */);
int64_t abo_sim_counter(abo_sim *, int which); /* 0 limit, 1 cancel, 2 fills, 3 spread queries, 4 max queue, 5 max bid lv, 6 max ask lv, 7 max resting, 8 n_orders_alloc, 9 uniq */
void abo_sim_book_l1(abo_sim *, int64_t *out5); /* bid, bid_qty, ask, ask_qty, last_trade */
int64_t abo_sim_fundamental(abo_sim *);          /* last oracle value r[symbol][1] */

/* ---------------- ABIDESEnv: Exchange + MarketReplayAgent + DummyRLExecutionAgent under GymKernel ---------------- */
typedef struct abo_env abo_env;
/* stream5: rows (t_ns since midnight, ORDER_ID, PRICE cents, SIZE, is_buy) as LOBSTEROrdersProcessor yields them
 * (agent/examples/MarketReplayAgent.py:162-220), sorted by time.  quantity / order_level: agent_config.py:132-154. */
abo_env *abo_env_new(const int64_t *stream5, int64_t n_rows, double quantity, int order_level, int trace_flags);
/* order_level 0: no RL agent == config/marketreplay.py (Exchange + MarketReplayAgent under Kernel.runner); one step() runs to the end */
abo_env *abo_env_new2(const int64_t *stream5, int64_t n_rows, double quantity, int order_level, int trace_flags, int64_t stop_ns);
void abo_env_free(abo_env *);
/* ABIDESEnv.step (ABIDESEnv.py:30-49): returns len(obs) (0 or 9), fills obs_out[9] and *done */
int abo_env_step(abo_env *, const double *action, double *obs_out, int *done);
int64_t abo_env_n_pops(abo_env *);
uint64_t abo_env_pop_hash(abo_env *);
uint64_t abo_env_note_hash(abo_env *);
uint64_t abo_env_snap_hash(abo_env *);
int64_t abo_env_n_hash_ckpt(abo_env *);
const uint64_t *abo_env_hash_ckpt(abo_env *);
int64_t abo_env_trace(abo_env *, int which, const int64_t **rows);
void abo_env_final(abo_env *, double *out8); /* rl rem_quantity, shares, cash, n_executed, replay shares, cash, open orders, now */
int64_t abo_env_counter(abo_env *, int which); /* 0 max queue, 1 max bid levels, 2 max ask levels, 3 max resting, 4 uniq, 5 next order id */

/* ---------------- DDQN execution config: config/execution/marketreplay/execution_marketreplay_ddqn.py (-a rl) ----------------
 * Exchange + MarketReplayAgent + n_mom MomentumAgents + n_twap TWAPExecutionAgents + DDQLearningExecutionAgent under Kernel.runner.
 * The Q-network is an input (the action of every decision tick); handles are abo_env (abo_env_free / n_pops / hashes / trace work). */
abo_env *abo_dq_new(const int64_t *stream5, int64_t n_rows, int n_mom, const int64_t *mom_sizes, int n_twap, int has_ddqn, int is_buy,
                    int64_t quantity, int64_t h0_ns, int64_t h_step_ns, int n_h, int64_t mom_wake_ns, int trace_flags);
int abo_dq_set_schedule(abo_env *, int k, const int64_t *qty, int n);   /* per-bin child quantities of baseline execution agent k (VWAP) */
int abo_dq_step(abo_env *, int action, double *out8, double *trans6, double *reward, int *done);
int abo_dq_error(abo_env *);
int64_t abo_dq_series(abo_env *, int which /*0 price_path, 1 experience x6, 2 step_reward_hist, 3 action_hist*/, const double **v);
void abo_dq_holdings(abo_env *, int64_t *out5);
void abo_dq_exec_final(abo_env *, int k, double *out5);

#ifdef __cplusplus
}
#endif
#endif
