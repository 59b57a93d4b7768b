// abx_core.cuh -- warp-uniform simulation logic of the batched ABIDES simulator (one warp == one environment).
//
// Everything in this header is *uniform* code: on the GPU all 32 lanes of the environment's warp execute it
// redundantly with identical register values (so ballots/shuffles inside the cooperative primitives of the
// context type never diverge), and only the context's primitives (abx_warp.cuh: event queue, price ladders,
// agent-record staging) split work across lanes.  The same header compiles as plain C++ for the host
// emulation harness under tests/emu (CPU CI of this logic; it is not part of libabides_b200.so).
//
// Reference semantics restated here (paths relative to the reference root):
//   Kernel.py:190-292,347-462            event loop, requeue rule, send/wakeup
//   agent/ExchangeAgent.py:129-340,471   exchange protocol
//   util/OrderBook.py:38-398             matching engine
//   agent/TradingAgent.py, agent/ZeroIntelligenceAgent.py   trader state machine, belief update, pricing
//   util/oracle/SparseMeanRevertingOracle.py:88-227        sparse OU fundamental with megashocks
//   model/LatencyModel.py:109-140        cubic latency
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/abides_b200.h"
#include "abx_exp_table.h"

#if defined(__CUDACC__)
#define ABX_HD __host__ __device__ __forceinline__
#define ABX_D __device__ __forceinline__
#define ABX_NI __host__ __device__ __noinline__     // one copy of the long libm / Philox bodies instead of one per call site
#else
#define ABX_HD inline
#define ABX_NI static inline
#endif

#if !defined(__CUDACC__)
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(8) int2 { int32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) int4 { int32_t x, y, z, w; };
inline float cospif(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
#endif

namespace abx {

// ---------------------------------------------------------------------------------------------------
// state structs (HBM layout; all 16-byte aligned so a warp moves them with 128-bit loads/stores)
// ---------------------------------------------------------------------------------------------------
struct alignas(16) EnvState {           // 192 B per environment
  int64_t now, ttl;                     // Kernel.currentTime, ttl_messages
  int64_t exch_time, exch_comp_delay;   // agentCurrentTimes[0], agentComputationDelays[0]
  int64_t or_t, ms_t;                   // oracle r[symbol][0], next megashock time
  double ms_v; uint64_t pop_hash;       // next megashock value; FNV-1a of the pop sequence
  uint64_t seed; int32_t or_v, last_trade; // oracle r[symbol][1]; OrderBook.last_trade
  uint32_t uniq, next_order_id, q_count, max_q; // Message.uniq, Order.order_id counters; queue fill
  int32_t n_bid_lv, n_ask_lv; int32_t n_resting; uint32_t free_head; // ladder sizes (0 bids, 1 asks); order-node free list
  uint32_t pool_top, flags, trace_n, c_limit;
  uint32_t c_cancel, c_fills, c_query, ctr_symbol;
  uint32_t ctr_kernel, ctr_latency, ctr_global, trade_epoch;   // trade_epoch: OrderBook.history rotations (util/OrderBook.py:146)
  int64_t sum_shares, sum_cash;
  uint32_t kblk[4];                     // cached Philox block of the kernel (latency-noise) stream
  uint32_t draw_n, evt_n, book_flags, episode; // entries in the draw log (parity runs) / in the LAST_TRADE, BEST_BID, BEST_ASK event ring; BKF_*; resets of this environment that moved on to the next replayed day
  uint32_t hist_n, n_subs; int64_t book_update; // orders registered in the order-history log (population 3: what QUERY_ORDER_STREAM hands out); subscribers (ExchangeAgent.subscription_dict); OrderBook.last_update_ts
  int64_t next_pub, pad_p;                      // earliest book-update time at which some subscriber is due a MARKET_DATA body (min over subscribers of last update + period)
};
static_assert(sizeof(EnvState) == 240, "EnvState layout");

enum : uint32_t {
  AF_HAS_OPEN = 1u, AF_HAS_CLOSE = 2u, AF_MKT_CLOSED = 4u, AF_HAS_LAST = 8u, AF_HAS_DAILY = 16u, AF_HAS_PREV = 32u,
  AF_HAS_BID = 64u, AF_HAS_ASK = 128u, AF_STATE_SHIFT = 8, AF_STATE_MASK = 3u << 8, AF_GROUP_SHIFT = 12, AF_GROUP_MASK = 7u << 12
};
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE = 1, ST_AWAITING_SPREAD = 2, ST_AWAITING_TV = 3, ST_AWAITING_STREAM = 3 };   // 3: AWAITING_TRANSACTED_VOLUME (POV execution agent) or AWAITING_STREAM (HBL agent): no class uses both // ZeroIntelligenceAgent.state; POVExecutionAgent AWAITING_TRANSACTED_VOLUME
enum : uint32_t { AF_TYPE_SHIFT = 16, AF_TYPE_MASK = 15u << 16, AF_SUB_REQUESTED = 1u << 20 };   // AF_SUB_REQUESTED: subscription mode, request sent (the agent's state is AWAITING_MARKET_DATA from then on)
enum { AT_ZI = 0, AT_NOISE = 1, AT_VALUE = 2, AT_MOMENTUM = 3, AT_POVMM = 4, AT_TWAP = 5, AT_DDQN = 6, AT_POVEXEC = 7, AT_MKM = 8, AT_HBL = 9 };   // agent class (rmsc03 / DDQN execution populations)
constexpr uint32_t EPOCH_START = 16;    // EnvState.trade_epoch of a fresh book (so that epoch differences of up to 16 never wrap below zero)
constexpr int AGENT_ORDER_CAP = 4;      // open orders tracked per trader (ZI holds <= 2)

struct alignas(16) ZiAgent {            // 192 B per trader: TradingAgent + ZeroIntelligenceAgent state
  int64_t agent_time;                   // Kernel.agentCurrentTimes[id]
  int64_t prev_wake;                    // ZI.prev_wake_time (valid with AF_HAS_PREV)
  double r_t, sigma_t;                  // ZI belief
  int64_t cash; int32_t shares, last_trade;     // holdings["CASH"], holdings[symbol], last_trade[symbol]
  int32_t daily_close, bid, bid_q, ask;         // daily_close_price, known_bids[0], known_asks[0]
  int32_t ask_q; uint32_t flags, rng_ctr; int32_t n_orders;
  uint32_t oid[AGENT_ORDER_CAP]; int32_t oprice[AGENT_ORDER_CAP]; int32_t oqty[AGENT_ORDER_CAP]; // self.orders (qty > 0 buy, < 0 sell)
  int16_t theta[20];                    // private values, sorted descending
  double lat_to, lat_from;              // latency[id][0], latency[0][id]  (min_latency for the cubic model)
  int64_t surplus;                      // FINAL_VALUATION (kernelStopping)
};
static_assert(sizeof(ZiAgent) == 192, "ZiAgent layout");

// rmsc03 agents overlay these 40 bytes on ZiAgent.theta (only ZI agents have private values)
struct AgentAux { int32_t size, order_size, last_mid, tv; uint32_t mmflags; int32_t n_mids; double avg20, avg50; };
static_assert(sizeof(AgentAux) == 40, "AgentAux overlays ZiAgent.theta");
// Execution agents (TWAPExecutionAgent / DDQLearningExecutionAgent of the DDQN config, POVExecutionAgent of the rmsc03 population) overlay these 88 bytes on ZiAgent.oid .. surplus
// (their open orders live in a per-environment HBM table: a market order walks up to DQ_DEPTH levels).
struct ExecAux {
  int32_t rem_qty, executed_sum, n_executed, arr2;      // remaining_qty / rem_quantity, sum of fills, len(executed_orders), 2 * arrival_price (0 == None)
  int32_t t, rem_time, n_pp, pp0_2;                     // DDQN: self.t, self.remaining_time, len(price_path), 2 * price_path[0]
  int32_t pp_last2, child_qty, exflags, e_a;            // 2 * price_path[-1]; schedule quantity; EXF_*; experience[t-1] action
  int16_t cur_s[2], sp[2];                              // self.s; s' of the pending decision
  int16_t e_s[2], e_sp[2];                              // experience[t-1] = (s, a, s', r)
  int32_t tv, snap_n;                                   // POVExecutionAgent: transacted_volume[symbol]; levels in the agent's cached book: bids | asks << 16 (known_bids / known_asks of the last QUERY_SPREAD reply)
  double e_r, step_reward;                              // r (valid with EXF_E_R); sum of step_reward_hist entries since the last decision
};
static_assert(sizeof(ExecAux) == 88 && sizeof(ExecAux) <= 112, "ExecAux overlays ZiAgent.oid .. surplus");
enum : uint32_t { EXF_TRADE = 1u, EXF_E_VALID = 2u, EXF_E_R = 4u };
constexpr int EXEC_ORDER_CAP = 512;     // self.orders of one execution agent
constexpr int DQ_DEPTH = 500;           // getCurrentSpread(depth=500) execution_agent.py:77, ddqlearning_execution_agent.py:152
enum : uint32_t { MMF_AW_SPREAD = 1u, MMF_AW_VOL = 2u, MMF_HAS_MID = 4u, MOF_HAS20 = 8u, MOF_HAS50 = 16u };
constexpr int SUB_CAP = 64;             // market-data subscribers per exchange (config/rmsc02.py has 25); their table follows the market maker's orders in idtab
constexpr int SUB_LEVELS = 8;           // levels a side a MARKET_DATA body can carry (MarketMakerAgent subscribes to 5)
constexpr int MM_ORDER_CAP = 128;       // POV market maker: 2 * (num_ticks + 1) = 42 orders placed per wake, cancelled at the next
constexpr int TV_RING_MIN = 512;        // recent (time, qty) transaction tuples kept for get_transacted_volume: SimParams.tv_ring >= this, 64 per surviving history bucket
constexpr int MOM_MIDS = 64;            // MomentumAgent: last 50 mid prices are all ma(20)/ma(50) need

struct Event {                          // one PriorityQueue entry, unpacked
  int64_t t; int32_t recipient, type; uint32_t uniq; int32_t kind, sender;
  int32_t p[6];                         // order: {order_id, limit_price, qty, fill_price, is_buy, -}
                                        // spread reply: {bid, bid_qty, ask, ask_qty, last_trade, has_bid|has_ask<<1|mkt_closed<<2}
  double lat_back;                      // trader -> exchange messages carry latency[0][sender] for the reply
  int32_t x0, x1;                       // exchange -> agent QUERY_SPREAD replies with depth > 1: level-2 bid / ask price (shares the lat_back words)
};

// 64-bit sort key: [ time : 47 | recipient : 15 | type : 2 ], ties broken by uniq (message/Message.py:39-45)
constexpr int KEY_T_SHIFT = 17;
constexpr int64_t KEY_T_MAX = (int64_t(1) << 46);   // times are saturated below 2^46 ns (19.5 h) -- far past any stop time
ABX_HD uint64_t key_pack(int64_t t, int recipient, int type) { return (uint64_t(t) << KEY_T_SHIFT) | (uint64_t(recipient) << 2) | uint64_t(type); }
ABX_HD int64_t key_time(uint64_t k) { return int64_t(k >> KEY_T_SHIFT); }
ABX_HD int key_recipient(uint64_t k) { return int((k >> 2) & 0x7fff); }
ABX_HD int key_type(uint64_t k) { return int(k & 3); }
constexpr uint64_t KEY_EMPTY = ~uint64_t(0);

struct EnvX;
// One resting order (16 B in HBM).  Replay shapes (ENV / DQ / BOOK), whose books see MODIFY_ORDER, store the order's own limit price
// {id, qty, price, next | agent << 16}: the reference reads a level's price from its slot 0 (util/OrderBook.py:381,393), so after a re-pricing
// modify the level takes the price of whichever order becomes its head.  The background populations never modify and keep
// {id, qty, agent | history epoch << 16, next}.
struct NodeRec { uint32_t id; int32_t qty; uint32_t agent; uint32_t next; int32_t price; };
template <bool PRICED> ABX_HD uint4 node_pack(const NodeRec &r) {
  uint4 v; v.x = r.id; v.y = (uint32_t)r.qty;
  if (PRICED) { v.z = (uint32_t)r.price; v.w = (r.next & 0xffffu) | (r.agent << 16); } else { v.z = r.agent; v.w = r.next; }
  return v;
}
template <bool PRICED> ABX_HD NodeRec node_unpack(const uint4 &v) {
  NodeRec r; r.id = v.x; r.qty = (int32_t)v.y;
  if (PRICED) { r.price = (int32_t)v.z; r.next = (v.w & 0xffffu) == 0xffffu ? 0xffffffffu : (v.w & 0xffffu); r.agent = v.w >> 16; } else { r.agent = v.z; r.next = v.w; r.price = 0; }
  return r;
}
constexpr uint32_t NIL = 0xffffffffu;
enum : uint32_t { BKF_REPRICED = 1u };   // a MODIFY changed a head order's price: level prices follow their heads from now on, ladders may be unsorted, the per-id census no longer names levels

// Device-side parameter block (config + derived constants + HBM base pointers).
struct SimParams {
  abx_sim_config c;
  int32_t n_envs, n_qgroups, n_streams, n_tapes;   // n_tapes > 0: tape mode with shared tapes, environment e replays tape e % n_tapes
  double one_minus_kappa_a;     // 1 - agent_kappa
  double log_base_a;            // log(1 - agent_kappa)
  double sigma_denom;           // 1 - (1 - agent_kappa) ** 2      (host libm pow, ZeroIntelligenceAgent.py:234)
  double sqrt_sigma_n, sqrt_sigma_pv, sqrt_megashock_var;
  double inv_lambda_a, inv_megashock_lambda;
  double ou_scale;              // (fund_vol ** 2) / (2 * kappa)     SparseMeanRevertingOracle.py:106
  // HBM arrays, environment-major
  uint4 *qkey, *qpay0, *qpay1;  // [n_envs][queue_cap]
  uint4 *qcache;                // [n_envs][n_qgroups]   {min key lo, hi, min uniq, occupancy mask}
  ZiAgent *agents;              // [n_envs][n_agents]    (index 0 unused: the exchange lives in EnvState)
  int32_t *lv_price, *lv_qty; uint32_t *lv_ht;  // [n_envs][2][level_cap]
  uint4 *nodes;                 // [n_envs][order_cap]
  EnvState *env;                // [n_envs]
  abx_trace_rec *trace;         // [n_envs][trace_cap]
  uint4 *draw_log;              // [n_envs][draw_log_cap] {stream | kind << 24, bits lo, bits hi, -}  (parity runs under Philox)
  const int32_t *sched;         // [dq_n_twap][n_h] per-bin child quantities of the baseline execution agents (VWAP schedule), < 0 = the TWAP quantity; shared by all environments
  uint4 *hlog;                  // [n_envs][hist_stride_of(c)] {order id, limit price, history epoch at registration, is_buy | has transactions << 1}  (population 3)
  uint4 *evt;                   // [n_envs][event_ring_cap] {t lo, t hi | kind << 28, a, b}: order arrivals, BEST_BID / BEST_ASK / LAST_TRADE (realism tooling)
  const uint64_t *tape_bits; const uint8_t *tape_kinds; const int64_t *tape_off; // tape mode
  // ---- ABIDESEnv shape (exchange + MarketReplayAgent + RL execution agent); zero for the sparse_zi shape ----
  int32_t n_ts, n_rows, n_ids, n_h;                 // replayed stream: timestamps, rows, distinct order ids; horizon length
  int64_t h0_ns, h_step_ns;                         // execution_time_horizon = h0 + k * step, k < n_h (agent_config.py:134-136)
  double h_step_inv;                                // 1 / h_step_ns (0 when there is no horizon): Sim::hdiv
  double rl_quantity, rl_steep; int32_t order_level, rl_is_buy;
  const int64_t *st_ts;                             // [n_ts]      distinct timestamps (ns), ascending
  const int32_t *st_first;                          // [n_ts + 1]  first row of each timestamp
  const int4 *st_rows;                              // [n_rows]    {dense id, PRICE cents, SIZE, is_buy}
  const int4 *day_tab; int32_t n_days, pad_days;    // [n_days] {first entry of st_ts, timestamps, first entry of st_first, -}: environment e replays day e % n_days
  const int4 *day_tab2;                             // [n_days] {first entry of st_xid, explicit ids, smallest, largest explicit ORDER_ID}
  const int32_t *st_xid, *st_xfirst;                // per day: the explicit ORDER_IDs sorted ascending, and the first row of each
  struct EnvX *envx;                                // [n_envs]
  uint4 *idtab;                                     // [n_envs][n_ids] {agent-view qty, price<<1|is_buy, last registration epoch, epoch mask}
  uint2 *idbook;                                    // [n_envs][n_ids] {nodes in the book carrying the id | several-levels flag << 31, price<<1|side of their level}
  int4 *lobs;                                       // [n_envs][LOB_CAP][3] stored QUERY_SPREAD replies (ABIDESEnvMetrics.data)
  // ---- DDQN execution config (population 2): ids 0 exchange, 1 replay, 2.. momentum, then TWAP agents, then the DDQN agent ----
  int32_t dq_n_mom, dq_n_twap, dq_has_ddqn, dq_order_base;   // dq_order_base: first idtab entry of the execution agents' order tables
  int64_t dq_quantity;                              // parent order size
  int64_t dq_id_limit;                              // smallest explicit ORDER_ID of the replayed stream: generated ids must stay below it
  // ---- deep QUERY_SPREAD replies (depth 500 / sys.maxsize: the execution agents): the exchange copies the first snap_depth levels of both sides into the
  // asking agent's area when it PROCESSES the query (agent/ExchangeAgent.py:231-245); the agent reads that copy, not the live ladders ----
  int2 *snap;                                       // [n_envs][n_snap][2][snap_depth] {price, qty}, best level first
  int32_t snap_depth, n_snap, tv_ring, pad_s;       // tv_ring: entries of the transaction-tuple ring (rmsc03 population), a power of two sized from stream_history
};
constexpr int LOB_CAP = 100;                        // ABIDESEnvMetrics(maxlen = 100) dummy_rl_execution_agent.py:125
constexpr int RL_ORDER_CAP = 8;
constexpr uint32_t REPLAY_ID_BASE = 0x40000000u;    // device order id of replayed order k = REPLAY_ID_BASE + dense id k
enum : uint32_t { RLF_TRADE = 1u << 16, RLF_METRICS_INIT = 1u << 17, DQF_PENDING = 1u << 18 };   // DQF_PENDING: the DDQN agent waits inside place_order for its action

struct alignas(16) EnvX {                           // per-environment state of the two ABIDESEnv traders (shared memory while stepping)
  int64_t ra_time, rl_time;                         // Kernel.agentCurrentTimes[1], [2]
  int64_t ra_cash, rl_cash;
  double rem_quantity, executed_sum;                // ExecutionAgent.rem_quantity, sum of executed quantities
  int32_t ra_shares, rl_shares, ra_last_trade, rl_last_trade;
  uint32_t ra_flags, rl_flags; int32_t wt_cursor, n_executed;
  int32_t rl_n_orders, n_lobs, lob_head, p0;
  int32_t rem_time, obs_len, g0_qty, steps;          // g0_*: the replay agent's order under GENERATED id 0 (rows with ORDER_ID 0)
  uint32_t rl_oid[RL_ORDER_CAP]; int32_t rl_oprice[RL_ORDER_CAP]; int32_t rl_oqty[RL_ORDER_CAP];
  double obs[9]; int32_t g0_pq, rows_done;           // rows_done: rows of the stream the replay agent has processed (which explicit ORDER_IDs count as used, Order.py:35-42)
};
static_assert(sizeof(EnvX) % 16 == 0, "EnvX layout");


// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
ABX_HD int64_t py_round_i64(double x) { return (int64_t)rint(x); }   // Python int(round(x)): half-to-even
ABX_HD uint64_t fnv_mix(uint64_t h, int64_t sv) {
  uint64_t v = (uint64_t)sv;
#pragma unroll
  for (int i = 0; i < 8; i++) { h = (h ^ (v & 0xFF)) * 0x100000001B3ULL; v >>= 8; }
  return h;
}
ABX_HD int __popc_compat(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
ABX_HD uint64_t dbl_bits(double d) { union { double d; uint64_t u; } x; x.d = d; return x.u; }
ABX_HD double bits_dbl(uint64_t u) { union { double d; uint64_t u; } x; x.u = u; return x.d; }
// fp64 arithmetic that must not be contracted into FMAs (the reference is CPython: every op rounds)
#if defined(__CUDA_ARCH__)
ABX_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
ABX_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
ABX_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
ABX_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
ABX_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
ABX_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
#endif

// queue slot encoding: key {hi.lo, hi.hi, uniq, kind | sender << 8}, pay0 {p0..p3}, pay1 {p4, p5, lat_back bits}
ABX_HD void event_pack(const Event &e, uint4 &k, uint4 &a, uint4 &b) {
  uint64_t hi = key_pack(e.t, e.recipient, e.type);
  uint64_t lb = e.sender == 0 ? ((uint64_t)(uint32_t)e.x0 | ((uint64_t)(uint32_t)e.x1 << 32)) : dbl_bits(e.lat_back);
  k.x = (uint32_t)hi; k.y = (uint32_t)(hi >> 32); k.z = e.uniq; k.w = (uint32_t)e.kind | ((uint32_t)e.sender << 8);
  a.x = (uint32_t)e.p[0]; a.y = (uint32_t)e.p[1]; a.z = (uint32_t)e.p[2]; a.w = (uint32_t)e.p[3];
  b.x = (uint32_t)e.p[4]; b.y = (uint32_t)e.p[5]; b.z = (uint32_t)lb; b.w = (uint32_t)(lb >> 32);
}
ABX_HD void event_unpack(const uint4 &k, const uint4 &a, const uint4 &b, Event &e) {
  uint64_t hi = (uint64_t)k.x | ((uint64_t)k.y << 32);
  e.t = key_time(hi); e.recipient = key_recipient(hi); e.type = key_type(hi); e.uniq = k.z; e.kind = (int)(k.w & 0xffu); e.sender = (int)(k.w >> 8);
  e.p[0] = (int32_t)a.x; e.p[1] = (int32_t)a.y; e.p[2] = (int32_t)a.z; e.p[3] = (int32_t)a.w; e.p[4] = (int32_t)b.x; e.p[5] = (int32_t)b.y;
  e.lat_back = bits_dbl((uint64_t)b.z | ((uint64_t)b.w << 32)); e.x0 = (int32_t)b.z; e.x1 = (int32_t)b.w;
}
ABX_HD bool key_less(uint64_t ah, uint32_t au, uint64_t bh, uint32_t bu) { return ah < bh || (ah == bh && au < bu); }

// ---------------------------------------------------------------------------------------------------
// RNG: Philox4x32-10 counter streams, or replay of recorded standard variates (tape)
// ---------------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
#ifndef ABX_PHILOX_ROUNDS
#define ABX_PHILOX_ROUNDS 10
#endif
#ifndef ABX_PHILOX_UNROLL
#define ABX_PHILOX_UNROLL 1
#endif
ABX_NI U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr int UNROLL = ABX_PHILOX_UNROLL;
#pragma unroll UNROLL                                                       // rolled: the unrolled rounds are 2 KB of a hot loop body that lives on the edge of the 32 KB instruction cache (+6 % msgs/s, profiles/r02_optimisation_log.md)
  for (int r = 0; r < ABX_PHILOX_ROUNDS; r++) {
    uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
    uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0, n1 = uint32_t(p1), n2 = uint32_t(p0 >> 32) ^ c3 ^ k1, n3 = uint32_t(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}
// exp(x) for the OU step of the fundamental and the (1 - kappa) ** x powers of the belief update (three per order placement).  The kernel is bound by dependent
// fp64 chains (profiles/r02_optimisation_log.md: replacing this body by a two-instruction stand-in gained 10.7 %), and the libm body is a ~20-deep Horner chain.
// Device: the table method of current libms (x = (128 k + j) ln2/128 + r, exp(x) = 2^k * H[j] * (1 + T[j] + r + r^2 (1/2 + r/6) + r^4 (1/24 + r/120))), nine
// dependent operations, maximum error 0.51 ulp over 2e8 points against long-double expl and equal to glibc's exp on 99.94 % of them (tools/check_exp.c) -- closer to
// the oracle's libm than the 1-ulp CUDA exp it replaces.  abx_selftest_exp exposes it to the tests.
#if defined(__CUDACC__) && !defined(ABX_LIBM_EXP)
__device__ __constant__ double2 abx_exp_tab[128] = { ABX_EXP_TABLE };
#endif
#if defined(__CUDA_ARCH__) && !defined(ABX_LIBM_EXP)
ABX_NI double exp_ni(double x) {
  const bool far = !(fabs(x) < 700.0);                                                   // results outside the normal range (or NaN): scaled in two exact steps below
  if (far) { if (x != x) return x; x = x < -800.0 ? -800.0 : (x > 800.0 ? 800.0 : x); }
  const double t = fma(x, ABX_EXP_INV_LN2N, 0x1.8p52);                                   // round(x * 128 / ln2) lands in the low mantissa bits
  const int k = __double2loint(t);
  const double kd = t - 0x1.8p52;
  double r = fma(kd, -ABX_EXP_LN2N_HI, x); r = fma(kd, -ABX_EXP_LN2N_LO, r);
  const double2 ht = abx_exp_tab[k & 127];
  const double r2 = r * r;
  const double p = fma(fma(1.0 / 120, r, 1.0 / 24), r2, fma(1.0 / 6, r, 0.5));
  const double q = fma(p, r2, ht.y + r);
  const double y = fma(ht.x, q, ht.x);
  const int e = k >> 7;
  if (far) { const int e1 = e / 2; return y * __hiloint2double((1023 + e1) << 20, 0) * __hiloint2double((1023 + e - e1) << 20, 0); }   // |e| <= 1155: both factors are normal powers of two, the last product rounds once (inf / subnormal / 0)
  return __hiloint2double(__double2hiint(y) + (e << 20), __double2loint(y));
}
#else
ABX_NI double exp_ni(double x) { return exp(x); }
#endif
ABX_NI double log_ni(double x) { return log(x); }
// log(x), x in (0, 1] and normal, for the Philox-mode variate transforms only (no parity constraint): fp64 throughout, relative error
// < 1e-11 (checked against libm by tests/test_gpu_philox.py through abx_selftest_log_unit).  x = 2^e * m with m in [sqrt(1/2), sqrt(2)],
// log(m) = 2 atanh(s), s = (m - 1) / (m + 1), seven terms of the series (|s| <= 0.1716), the quotient through a refined MUFU reciprocal.
// ~30 instructions inline; the libm log body is 237 instructions of instruction-cache footprint and 3 of them run per order.
ABX_HD double log_unit(double x) {
#if defined(__CUDA_ARCH__)
  uint64_t b = dbl_bits(x);
  int e = (int)(b >> 52) - 1023;
  double m = bits_dbl((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);            // [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  double d = m + 1.0, r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r); r = fma(r, fma(-d, r, 1.0), r);                      // two Newton steps: 1 / d to ~1 ulp
  double s = (m - 1.0) * r, s2 = s * s;
  double p = fma(s2, 1.0 / 13.0, 1.0 / 11.0); p = fma(p, s2, 1.0 / 9.0); p = fma(p, s2, 1.0 / 7.0); p = fma(p, s2, 1.0 / 5.0); p = fma(p, s2, 1.0 / 3.0); p = fma(p, s2, 1.0);
  return fma((double)e, 0.6931471805599453, 2.0 * s * p);
#else
  return log(x);
#endif
}
ABX_NI double box_muller(uint32_t a, uint32_t b, uint32_t cw) {                        // one body for every normal draw (value in, value out: nothing address-taken)
  double u1 = 1.0 - ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0;           // u1 in (0,1]
  float ang = (float)(cw >> 8) * (2.0f / 16777216.0f);                                 // angle / pi in [0, 2): 24 random bits, fp32 cos
  return sqrt(-2.0 * log_unit(u1)) * (double)cospif(ang);
}
ABX_NI double pow_ni(double x, double y) { return pow(x, y); }
enum { S_SYMBOL = 0, S_KERNEL = 1, S_LATENCY = 2, S_GLOBAL = 3, S_AGENT0 = 3 };  // agent a uses stream S_AGENT0 + a

// MODE: ABX_RNG_PHILOX / ABX_RNG_TAPE fixed at compile time (one kernel instantiation per mode), or -1 = read P->c.rng_mode
// REC: keep a log of every standard variate the Philox streams hand out (parity instrumentation, compiled out of the production kernels)
template <int MODE, bool REC = false>
struct RngT {
  const SimParams *P; int env; uint64_t seed; uint32_t err; uint4 *lg; uint32_t lg_n, lg_cap;
  ABX_HD bool tape() const { return MODE < 0 ? P->c.rng_mode == ABX_RNG_TAPE : MODE == ABX_RNG_TAPE; }
  ABX_HD void rec(int stream, uint8_t kind, uint64_t bits) {                            // uniform code: every lane stores the same entry
    if (!REC || !lg) return;
    if (lg_n >= lg_cap) { err |= ABX_F_TRACE_OVERFLOW; return; }
    uint4 v; v.x = (uint32_t)stream | ((uint32_t)kind << 24); v.y = (uint32_t)bits; v.z = (uint32_t)(bits >> 32); v.w = 0u; lg[lg_n++] = v;
  }
  ABX_HD uint64_t tape_next(int stream, uint32_t &ctr, uint8_t kind) {
    const int64_t *off = P->tape_off + (int64_t)env * P->n_streams + stream;
    int64_t i = off[0] + ctr;
    if (i >= off[1]) { err |= ABX_F_TAPE_UNDERRUN; return 0; }
    ctr++;
    if (P->tape_kinds[i] != kind) err |= ABX_F_TAPE_KIND;
    return P->tape_bits[i];
  }
  ABX_HD U4 philox(int stream, uint32_t &ctr) { U4 o = philox4x32_10(ctr, (uint32_t)stream, 0x41424958u, 0, (uint32_t)seed, (uint32_t)(seed >> 32)); ctr++; return o; }
  static ABX_HD double u53(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0; }
  ABX_HD double std_normal(int stream, uint32_t &ctr) {
    if (tape()) return bits_dbl(tape_next(stream, ctr, 'n'));
    U4 o = philox(stream, ctr);
    double z = box_muller(o.x, o.y, o.z); rec(stream, 'n', dbl_bits(z)); return z;
  }
  ABX_HD double std_exponential(int stream, uint32_t &ctr) {
    if (tape()) return bits_dbl(tape_next(stream, ctr, 'e'));
    U4 o = philox(stream, ctr); double e = -log_unit(1.0 - u53(o.x, o.y)); rec(stream, 'e', dbl_bits(e)); return e;
  }
  ABX_HD double u01(int stream, uint32_t &ctr) {
    if (tape()) return bits_dbl(tape_next(stream, ctr, 'u'));
    U4 o = philox(stream, ctr); double u = u53(o.x, o.y); rec(stream, 'u', dbl_bits(u)); return u;
  }
  // integer in [0, range] (numpy randint(low, high) with range = high - 1 - low; no draw when range == 0)
  ABX_HD int64_t randint(int stream, uint32_t &ctr, uint32_t range) {
    if (range == 0) { if (tape()) return (int64_t)tape_next(stream, ctr, 'i'); rec(stream, 'i', 0); return 0; }
    if (tape()) return (int64_t)tape_next(stream, ctr, 'i');
    U4 o = philox(stream, ctr); int64_t v = (int64_t)((uint64_t(o.x) * (uint64_t(range) + 1)) >> 32); rec(stream, 'i', (uint64_t)v); return v;
  }
  // Latency-noise draws (one per sendMessage, Kernel.py:411) take successive 32-bit words of the kernel stream's Philox
  // blocks; the current block is cached in EnvState so only every 4th send runs the 10 rounds.
  ABX_HD int64_t randint_cached(int stream, uint32_t &ctr, uint32_t range, uint32_t blk[4]) {
    if (tape()) return (int64_t)tape_next(stream, ctr, 'i');
    uint32_t i = ctr++;
    if ((i & 3u) == 0u) { U4 o = philox4x32_10(i >> 2, (uint32_t)stream, 0x4b424958u, 0, (uint32_t)seed, (uint32_t)(seed >> 32)); blk[0] = o.x; blk[1] = o.y; blk[2] = o.z; blk[3] = o.w; }
    uint32_t w = (i & 3u) == 0u ? blk[0] : ((i & 3u) == 1u ? blk[1] : ((i & 3u) == 2u ? blk[2] : blk[3]));
    int64_t v = (int64_t)((uint64_t(w) * (uint64_t(range) + 1)) >> 32); rec(stream, 'i', (uint64_t)v); return v;
  }
  ABX_HD double normal(int stream, uint32_t &ctr, double loc, double scale) { return dadd(loc, dmul(scale, std_normal(stream, ctr))); }
};
typedef RngT<-1> Rng;


// ---------------------------------------------------------------------------------------------------
// reset: agent construction (config/sparse_zi_1000.py:211-251, ZeroIntelligenceAgent.__init__ :65-70) -- one
// thread per (environment, trader) on the GPU.  In tape mode lat_to / lat_from were preloaded by the host.
// ---------------------------------------------------------------------------------------------------
// int4 entries per environment in SimParams.lobs: the ABIDESEnv LOB ring, or the momentum agents' mid-price rings (population 3 has 24 of them)
ABX_HD int lob_stride_of(const abx_sim_config &c) { int need = c.population == 3 ? (c.n_momentum_agents * MOM_MIDS + 3) / 4 : 0; return need > LOB_CAP * 3 ? need : LOB_CAP * 3; }
// uint4 entries per environment in SimParams.hlog: the ring of hist_log_cap order records + hist_log_cap / 4 scratch rows (8 B each) for the HBL belief table
ABX_HD size_t hist_stride_of(const abx_sim_config &c) { return (size_t)c.hist_log_cap + (size_t)c.hist_log_cap / 8; }
ABX_HD int agent_type_of(const abx_sim_config &c, int id) {
  if (c.population == 0) return AT_ZI;
  if (c.population == 3) {                                                              // config/rmsc01.py: market maker(s), ZI, HBL, momentum
    if (id <= c.n_mm_agents) return AT_MKM;
    if (id <= c.n_mm_agents + c.groups[0].count) return AT_ZI;
    if (id <= c.n_mm_agents + c.groups[0].count + c.groups[1].count) return AT_HBL;
    return AT_MOMENTUM;
  }
  if (c.population == 2) return id < 2 + c.n_momentum_agents ? AT_MOMENTUM : (id == c.n_agents - 1 && c.n_mm_agents ? AT_DDQN : AT_TWAP);   // n_mm_agents doubles as has_ddqn
  if (id <= c.n_noise_agents) return AT_NOISE;
  if (id <= c.n_noise_agents + c.n_value_agents) return AT_VALUE;
  if (id <= c.n_noise_agents + c.n_value_agents + c.n_mm_agents) return AT_POVMM;
  if (c.n_pov_exec && id == c.n_agents - 1) return AT_POVEXEC;
  return AT_MOMENTUM;
}
// util.get_wake_time (util/util.py:35-58): U-quadratic inverse CDF on [0, 1]
ABX_HD double u_quadratic_inverse_cdf(double y) {
  double n = dsub(dmul(3.0 / 12.0, y), 0.125);
  double c = n < 0 ? -pow(-n, 1.0 / 3.0) : pow(n, 1.0 / 3.0);
  return dadd(c, 0.5);
}
ABX_HD void init_agent_record_r3(const SimParams &P, int env, int id, uint64_t seed, ZiAgent *z, uint32_t *err) {
  RngT<-1> rng; rng.P = &P; rng.env = P.n_tapes > 0 ? env % P.n_tapes : env; rng.seed = seed; rng.err = 0;
  int type = agent_type_of(P.c, id); uint32_t ctr = 0, c2 = 0; int cs = P.n_streams + id;
  int32_t size = (int32_t)z->lat_to; int64_t wake = (int64_t)z->lat_from;        // tape mode: preloaded by the host (drawn by the config script)
  if (P.c.rng_mode == ABX_RNG_PHILOX) {
    if (type == AT_NOISE) { double m = u_quadratic_inverse_cdf(rng.u01(cs, c2)); wake = P.c.noise_wake_lo_ns + (int64_t)dmul(m, (double)(P.c.noise_wake_hi_ns - P.c.noise_wake_lo_ns)); }
    if (type == AT_NOISE || type == AT_VALUE) size = P.c.size_lo + (int32_t)rng.randint(cs, c2, (uint32_t)(P.c.size_hi - P.c.size_lo - 1));
  }
  if (type == AT_MOMENTUM) size = P.c.mom_min_size + (int32_t)rng.randint(S_AGENT0 + id, ctr, (uint32_t)(P.c.mom_max_size - P.c.mom_min_size - 1));   // MomentumAgent.py:42 (own stream)
  z->agent_time = P.c.start_ns; z->prev_wake = type == AT_NOISE ? wake : 0; z->r_t = P.c.r_bar; z->sigma_t = 0.0; z->cash = P.c.starting_cash; z->shares = 0; z->last_trade = 0;
  z->daily_close = 0; z->bid = 0; z->bid_q = 0; z->ask = 0; z->ask_q = 0;
  z->flags = ((uint32_t)type << AF_TYPE_SHIFT) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT); z->rng_ctr = ctr; z->n_orders = 0;
  for (int i = 0; i < AGENT_ORDER_CAP; i++) { z->oid[i] = 0; z->oprice[i] = 0; z->oqty[i] = 0; }
  AgentAux ax; ax.size = size; ax.order_size = P.c.mm_min_order_size; ax.last_mid = 0; ax.tv = 0; ax.mmflags = type == AT_POVMM ? (MMF_AW_SPREAD | MMF_AW_VOL) : 0u; ax.n_mids = 0; ax.avg20 = 0.0; ax.avg50 = 0.0;
  *reinterpret_cast<AgentAux *>(z->theta) = ax;
  z->lat_to = 0.0; z->lat_from = 0.0; z->surplus = 0;
  if (type == AT_POVEXEC) {                                                           // POVExecutionAgent.__init__ (pov_agent.py:34-48)
    ExecAux ex; ex.rem_qty = (int32_t)P.c.pov_exec_quantity; ex.executed_sum = 0; ex.n_executed = 0; ex.arr2 = 0; ex.t = 0; ex.rem_time = 0; ex.n_pp = 0; ex.pp0_2 = 0; ex.pp_last2 = 0; ex.child_qty = 0;
    ex.exflags = EXF_TRADE; ex.e_a = 0; for (int i = 0; i < 2; i++) { ex.cur_s[i] = 0; ex.sp[i] = 0; ex.e_s[i] = 0; ex.e_sp[i] = 0; } ex.tv = 0; ex.snap_n = 0; ex.e_r = 0.0; ex.step_reward = 0.0;
    *reinterpret_cast<ExecAux *>(z->oid) = ex;
  }
  if (rng.err) *err |= rng.err;
}
// round(x / 2) of MarketMakerAgent.size (agent/market_makers/MarketMakerAgent.py:55,99): Python's round half to even on the float quotient
ABX_HD int32_t mkm_half(int64_t v) { return (int32_t)rint((double)v / 2); }
// the trader's pair of entries of the config's latency matrix (to / from the exchange): zero (ABX_LAT_ZERO), what abx_sim_reset_tape handed in (tape mode), or
// drawn per environment from the matrix' own Philox stream
template <class R> ABX_HD void init_latency_pair(const SimParams &P, R &rng, int id, const ZiAgent *z, double &lat_to, double &lat_from) {
  lat_to = z->lat_to; lat_from = z->lat_from;
  if (P.c.latency_model == ABX_LAT_ZERO) { lat_to = 0.0; lat_from = 0.0; }              // np.zeros latency matrix (config/rmsc01.py:250)
  else if (P.c.rng_mode == ABX_RNG_PHILOX) {
    uint32_t c2 = 0; int cs = P.n_streams + id;
    lat_to = dadd(P.c.latency_lo, dmul(dsub(P.c.latency_hi, P.c.latency_lo), rng.u01(cs, c2)));
    lat_from = P.c.latency_mirrored ? lat_to : dadd(P.c.latency_lo, dmul(dsub(P.c.latency_hi, P.c.latency_lo), rng.u01(cs, c2)));
  }
}
ABX_HD void init_agent_record(const SimParams &P, int env, int id, uint64_t seed, ZiAgent *z, uint32_t *err) {
  if (P.c.population == 1) { init_agent_record_r3(P, env, id, seed, z, err); return; }
  Rng rng; rng.P = &P; rng.env = P.n_tapes > 0 ? env % P.n_tapes : env; rng.seed = seed; rng.err = 0;
  int grp = 0, acc = 1, type3 = P.c.population == 3 ? agent_type_of(P.c, id) : AT_ZI;
  if (P.c.population == 3) acc += P.c.n_mm_agents;
  if (type3 == AT_MKM || type3 == AT_MOMENTUM) {                                        // population 3: MarketMakerAgent.__init__ :55 / MomentumAgent.__init__ :42 draw their size from the agent's stream
    uint32_t ctr = 0; int32_t size;
    if (type3 == AT_MKM) size = mkm_half(P.c.mkm_min_size + rng.randint(S_AGENT0 + id, ctr, (uint32_t)(P.c.mkm_max_size - P.c.mkm_min_size - 1)));
    else size = P.c.mom_min_size + (int32_t)rng.randint(S_AGENT0 + id, ctr, (uint32_t)(P.c.mom_max_size - P.c.mom_min_size - 1));
    z->agent_time = P.c.start_ns; z->prev_wake = 0; z->r_t = P.c.r_bar; z->sigma_t = 0.0; z->cash = P.c.starting_cash; z->shares = 0; z->last_trade = 0;
    z->daily_close = 0; z->bid = 0; z->bid_q = 0; z->ask = 0; z->ask_q = 0;
    z->flags = ((uint32_t)type3 << AF_TYPE_SHIFT) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT); z->rng_ctr = ctr; z->n_orders = 0;
    for (int i = 0; i < AGENT_ORDER_CAP; i++) { z->oid[i] = 0; z->oprice[i] = 0; z->oqty[i] = 0; }
    AgentAux ax; ax.size = size; ax.order_size = 0; ax.last_mid = 0; ax.tv = 0; ax.mmflags = 0; ax.n_mids = 0; ax.avg20 = 0.0; ax.avg50 = 0.0;
    *reinterpret_cast<AgentAux *>(z->theta) = ax;
    double lt, lf; init_latency_pair(P, rng, id, z, lt, lf);
    z->lat_to = lt; z->lat_from = lf; z->surplus = 0;
    if (rng.err) *err |= rng.err;
    return;
  }
  for (int g = 0; g < P.c.n_groups; g++) { if (id >= acc && id < acc + P.c.groups[g].count) grp = g; acc += P.c.groups[g].count; }
  uint32_t ctr = 0; int stream = S_AGENT0 + id; int m = 2 * P.c.q_max;
  double th[20];
  for (int i = 0; i < 20; i++) th[i] = 0;
  for (int i = 0; i < m; i++) th[i] = rint(rng.normal(stream, ctr, 0.0, P.sqrt_sigma_pv));   // np.round(normal(0, sqrt(sigma_pv)))
  for (int i = 1; i < m; i++) { double x = th[i]; int j = i - 1; while (j >= 0 && th[j] < x) { th[j + 1] = th[j]; j--; } th[j + 1] = x; }  // sorted(reverse=True)
  double lat_to, lat_from; init_latency_pair(P, rng, id, z, lat_to, lat_from);
  z->agent_time = P.c.start_ns; z->prev_wake = 0; z->r_t = P.c.r_bar; z->sigma_t = 0.0; z->cash = P.c.starting_cash; z->shares = 0; z->last_trade = 0;
  z->daily_close = 0; z->bid = 0; z->bid_q = 0; z->ask = 0; z->ask_q = 0; z->flags = ((uint32_t)grp << AF_GROUP_SHIFT) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT) | ((uint32_t)type3 << AF_TYPE_SHIFT);
  z->rng_ctr = ctr; z->n_orders = 0;
  for (int i = 0; i < AGENT_ORDER_CAP; i++) { z->oid[i] = 0; z->oprice[i] = 0; z->oqty[i] = 0; }
  for (int i = 0; i < 20; i++) { double v = th[i]; if (v > 32767.0) { v = 32767.0; rng.err |= ABX_F_THETA_INDEX; } if (v < -32768.0) { v = -32768.0; rng.err |= ABX_F_THETA_INDEX; } z->theta[i] = (int16_t)v; }
  z->lat_to = lat_to; z->lat_from = lat_from; z->surplus = 0;
  if (rng.err) *err |= rng.err;
}
// DDQN execution config (population 2): MomentumAgent.__init__ (agent/examples/MomentumAgent.py:39-46), ExecutionAgent.__init__
// (agent/execution/baselines/execution_agent.py:34-48), TWAP / DDQN generate_schedule (twap_agent.py:51-63, ddqlearning_execution_agent.py:496-505).
// mom_size < 0: draw MomentumAgent.size = randint(min_size, max_size) from the agent's Philox stream.
ABX_HD void init_agent_record_dq(const SimParams &P, int env, int id, uint64_t seed, int32_t mom_size, ZiAgent *z) {
  int type = agent_type_of(P.c, id); uint32_t ctr = 0;
  z->agent_time = P.c.start_ns; z->prev_wake = 0; z->r_t = 0.0; z->sigma_t = 0.0; z->cash = P.c.starting_cash; z->shares = 0; z->last_trade = 0;
  z->daily_close = 0; z->bid = 0; z->bid_q = 0; z->ask = 0; z->ask_q = 0; z->n_orders = 0;
  for (int i = 0; i < AGENT_ORDER_CAP; i++) { z->oid[i] = 0; z->oprice[i] = 0; z->oqty[i] = 0; }
  for (int i = 0; i < 20; i++) z->theta[i] = 0;
  z->lat_to = 0.0; z->lat_from = 0.0; z->surplus = 0;
  if (type == AT_MOMENTUM) {
    if (mom_size < 0) { RngT<ABX_RNG_PHILOX> rng; rng.P = &P; rng.env = env; rng.seed = seed; rng.err = 0; mom_size = P.c.mom_min_size + (int32_t)rng.randint(S_AGENT0 + id, ctr, (uint32_t)(P.c.mom_max_size - P.c.mom_min_size - 1)); }
    AgentAux ax; ax.size = mom_size; ax.order_size = 0; ax.last_mid = 0; ax.tv = 0; ax.mmflags = 0; ax.n_mids = 0; ax.avg20 = 0.0; ax.avg50 = 0.0;
    *reinterpret_cast<AgentAux *>(z->theta) = ax;
  } else {
    ExecAux ex; ex.rem_qty = (int32_t)P.dq_quantity; ex.executed_sum = 0; ex.n_executed = 0; ex.arr2 = 0; ex.t = 0; ex.rem_time = P.n_h - 1; ex.n_pp = 0; ex.pp0_2 = 0; ex.pp_last2 = 0;
    ex.child_qty = type == AT_DDQN ? (int32_t)((double)P.dq_quantity / (double)(P.n_h - 1)) : (int32_t)((double)P.dq_quantity / (double)P.n_h);
    ex.exflags = EXF_TRADE; ex.e_a = 0; for (int i = 0; i < 2; i++) { ex.cur_s[i] = 0; ex.sp[i] = 0; ex.e_s[i] = 0; ex.e_sp[i] = 0; } ex.tv = 0; ex.snap_n = 0; ex.e_r = 0.0; ex.step_reward = 0.0;
    *reinterpret_cast<ExecAux *>(z->oid) = ex;
  }
  z->flags = ((uint32_t)type << AF_TYPE_SHIFT) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT); z->rng_ctr = ctr;
}
ABX_HD void init_env_state(const SimParams &P, uint64_t seed, EnvState &s) {
  s.now = P.c.start_ns; s.ttl = 0; s.exch_time = P.c.start_ns; s.exch_comp_delay = P.c.default_computation_delay_ns;   // Kernel.py:97,105
  s.or_t = P.c.mkt_open_ns; s.ms_t = 0; s.ms_v = 0.0; s.pop_hash = 0xCBF29CE484222325ULL; s.seed = seed;
  s.or_v = (int32_t)P.c.r_bar; s.last_trade = (int32_t)P.c.r_bar;     // SparseMeanRevertingOracle.py:58; ExchangeAgent.kernelInitializing :91-102
  s.uniq = 0; s.next_order_id = 0; s.q_count = 0; s.max_q = 0; s.n_bid_lv = s.n_ask_lv = 0; s.n_resting = 0; s.free_head = NIL;
  s.pool_top = 0; s.flags = 0; s.trace_n = 0; s.c_limit = s.c_cancel = s.c_fills = s.c_query = 0;
  s.ctr_symbol = s.ctr_kernel = s.ctr_latency = s.ctr_global = 0; s.trade_epoch = EPOCH_START; s.sum_shares = 0; s.sum_cash = 0;
  s.kblk[0] = s.kblk[1] = s.kblk[2] = s.kblk[3] = 0; s.draw_n = 0; s.evt_n = 0; s.book_flags = 0; s.episode = 0; s.hist_n = 0; s.n_subs = 0; s.book_update = 0; s.next_pub = KEY_T_MAX; s.pad_p = 0;
}

// ---------------------------------------------------------------------------------------------------
// The simulation, templated on the context type that supplies the cooperative primitives:
//   queue:   bool q_min(uint64_t&hi, uint32_t&uniq, int&group); void q_fetch(group, Event&); void q_remove(); void q_requeue(int64_t t);
//            bool q_push(const Event&, now)       (false == overflow)
//   ladders: void lv_find(side, price, int n, int&pos, bool&found); lv_insert/lv_remove; lv_price/lv_qty/lv_head/lv_tail get/set
//   nodes:   NodeRec node_load(i); void node_store(i, rec)
//   agents:  ZiAgent* agent_stage(id)  -- 128-bit cooperative copy HBM -> on-chip staging record, returned pointer is
//            readable by every lane; void agent_commit(id) -- staging record -> HBM; double agent_lat_from(id)
//   outbox:  uint32_t* outbox()  -- on-chip array of OUT_CAP * OUT_WORDS words
//   misc:    bool onchip_writer(); void trace(const abx_trace_rec&, uint32_t idx); void sync();
// Rule for on-chip memory shared by the lanes: uniform code lets EVERY lane store the same value (one STS, no
// divergent branch), and c.sync() -- a compiler fence on the GPU, where a converged warp issues its memory
// instructions in program order -- separates a write from later reads.
// ---------------------------------------------------------------------------------------------------
constexpr int OUT_CAP = 24, OUT_WORDS = 12;     // 24 x 48 B in the replay shapes (the DDQN agent's closing market order flushes in batches); the sparse_zi shape, whose handlers emit at most three messages, keeps 8 (Ctx::OUTN)
enum : uint32_t { OF_WAKEUP = 1u << 24, OF_BUMP_UNIQ = 1u << 25, OF_FROM_EXCH = 1u << 26, OF_CANCEL_EVT = 1u << 27 };

struct AgentRegs {                      // scalar part of ZiAgent held in registers while an event is handled
  int64_t agent_time, prev_wake, cash; double r_t, sigma_t, lat_to, lat_from;
  int32_t shares, last_trade, daily_close, bid, bid_q, ask, ask_q, n_orders; uint32_t flags, rng_ctr;
};
ABX_HD void regs_load(AgentRegs &a, const ZiAgent *z) {
  a.agent_time = z->agent_time; a.prev_wake = z->prev_wake; a.cash = z->cash; a.r_t = z->r_t; a.sigma_t = z->sigma_t;
  a.lat_to = z->lat_to; a.lat_from = z->lat_from; a.shares = z->shares; a.last_trade = z->last_trade; a.daily_close = z->daily_close;
  a.bid = z->bid; a.bid_q = z->bid_q; a.ask = z->ask; a.ask_q = z->ask_q; a.n_orders = z->n_orders; a.flags = z->flags; a.rng_ctr = z->rng_ctr;
}
ABX_HD void regs_store(ZiAgent *z, const AgentRegs &a) {
  z->agent_time = a.agent_time; z->prev_wake = a.prev_wake; z->cash = a.cash; z->r_t = a.r_t; z->sigma_t = a.sigma_t;
  z->shares = a.shares; z->last_trade = a.last_trade; z->daily_close = a.daily_close;
  z->bid = a.bid; z->bid_q = a.bid_q; z->ask = a.ask; z->ask_q = a.ask_q; z->n_orders = a.n_orders; z->flags = a.flags; z->rng_ctr = a.rng_ctr;
}

// RNG_MODE / LAT_MODEL: compile-time copies of cfg.rng_mode / cfg.latency_model (-1 = decide at run time);
// INSTR: parity instrumentation (pop hash + trace records) compiled in or out.
enum { SHAPE_ZI = 0, SHAPE_ENV = 1, SHAPE_R3 = 2, SHAPE_DQ = 3, SHAPE_BOOK = 4, SHAPE_P3 = 5 };   // sparse_zi population | ABIDESEnv / marketreplay | rmsc03 population | DDQN execution config | bare order books (op-tape replay)
template <class Ctx, int RNG_MODE = -1, int LAT_MODEL = -1, bool INSTR = true, int SHAPE = SHAPE_ZI>
struct Sim {
  static constexpr bool DQ = SHAPE == SHAPE_DQ, BOOK = SHAPE == SHAPE_BOOK, ENV = SHAPE == SHAPE_ENV || DQ || BOOK, P3 = SHAPE == SHAPE_P3, R3 = SHAPE == SHAPE_R3 || P3;   // P3: config/rmsc01.py population (runs the rmsc03 loop with more agent classes)
  Ctx &c; const SimParams &P; EnvState s; RngT<RNG_MODE, INSTR> rng; int64_t addl_delay; int n_out; int self_id; int env_id; bool cancel_found = false; uint32_t hidx = 0;   // cancel_found: the last book_cancel removed an order (it moves OrderBook.last_update_ts); hidx: order-history log index of the order handleLimitOrder is working on (population 3)
  AgentRegs a; ZiAgent *z;                // the trader whose event is being handled (registers + staged record)
  uint32_t kb_win = 0xffffffffu;          // device: which 32-block window of the kernel stream the lanes' s.kblk hold (kernel_noise)

  ABX_HD Sim(Ctx &c_, const SimParams &P_, const EnvState &s_, int env) : c(c_), P(P_), s(s_), addl_delay(0), n_out(0), self_id(0), z(nullptr) {
    rng.P = &P; rng.env = P.n_tapes > 0 ? env % P.n_tapes : env; env_id = env; rng.seed = s.seed; rng.err = 0; a.lat_from = 0.0; a.lat_to = 0.0;
    rng.lg = (INSTR && P.draw_log) ? P.draw_log + (size_t)env * (size_t)P.c.draw_log_cap : nullptr; rng.lg_n = s.draw_n; rng.lg_cap = (uint32_t)P.c.draw_log_cap;
  }
  // fold the generator's error bits and draw-log cursor back into the environment state (end of every entry point that may have drawn)
  ABX_HD void rng_sync() { s.flags |= rng.err; if (INSTR) s.draw_n = rng.lg_n; }

  ABX_HD bool lat_zero() const { return LAT_MODEL < 0 ? P.c.latency_model == ABX_LAT_ZERO : LAT_MODEL == ABX_LAT_ZERO; }
  ABX_HD int n_lv(int side) const { return side ? s.n_ask_lv : s.n_bid_lv; }
  ABX_HD void set_n_lv(int side, int n) { if (side) s.n_ask_lv = n; else s.n_bid_lv = n; }

  // ---- tracing (parity runs only) ----
  ABX_HD void trace_rec(const abx_trace_rec &r) {
    if (!INSTR) return;
    if (s.trace_n >= (uint32_t)P.c.trace_cap) { s.flags |= ABX_F_TRACE_OVERFLOW; return; }
    c.trace(r, s.trace_n); s.trace_n++;
  }
  ABX_HD void trace_note(int recipient, int kind, const int32_t p[6]) {
    abx_trace_rec r; r.tag = 1; r.a = recipient; r.t = s.now;
    for (int i = 0; i < 16; i++) r.v[i] = 0;
    r.v[0] = kind;
    if (kind == ABX_ORDER_ACCEPTED || kind == ABX_ORDER_EXECUTED || kind == ABX_ORDER_CANCELLED || kind == ABX_ORDER_MODIFIED) {
      r.v[1] = p[0]; r.v[2] = p[4]; r.v[3] = p[2]; r.v[4] = p[1]; r.v[5] = p[3];
    } else if (kind == ABX_QUERY_SPREAD) {
      r.v[5] = p[4];
      if (p[5] & 1) { r.v[6] = p[0]; r.v[7] = p[1]; }
      if (p[5] & 2) { r.v[8] = p[2]; r.v[9] = p[3]; }
      r.v[10] = (p[5] >> 2) & 1;
    } else if (P3 && kind == ABX_MARKET_DATA) {                                         // level counts, position-weighted price / size sums over the levels, top of book, last trade (what tools/record_reference.py stores)
      int k = p[2]; r.v[1] = p[0]; r.v[2] = p[1]; r.v[5] = p[4];
      for (int i = 0; i < p[0]; i++) { int2 lv = c.snap_load(k, 0, i); r.v[3] += (i + 1) * lv.x; if (i == 0) { r.v[6] = lv.x; r.v[7] = lv.y; } }
      for (int i = 0; i < p[1]; i++) { int2 lv = c.snap_load(k, 1, i); r.v[4] += (i + 1) * lv.x; if (i == 0) { r.v[8] = lv.x; r.v[9] = lv.y; } }
      for (int i = 0; i < p[0] && i < p[1]; i++) r.v[10] += (i + 1) * (c.snap_load(k, 0, i).y + 3 * c.snap_load(k, 1, i).y);
    }
    trace_rec(r);
  }
  ABX_HD void trace_snap() {
    if (!INSTR || P.c.trace_cap <= 0) return;
    abx_trace_rec r; r.tag = 2; r.a = 0; r.t = s.now;
    for (int i = 0; i < 16; i++) r.v[i] = 0;
    r.v[0] = s.n_bid_lv; r.v[1] = s.n_ask_lv; r.v[2] = s.n_resting;
    for (int side = 0; side < 2; side++) for (int k = 0; k < 3; k++) {
      int n = n_lv(side); if (k < n) { r.v[3 + side * 6 + 2 * k] = c.lv_price(side, n - 1 - k); r.v[4 + side * 6 + 2 * k] = c.lv_qty(side, n - 1 - k); }
    }
    r.v[15] = s.last_trade;
    trace_rec(r);
  }

  // ---- outbox: messages and wakeups produced while handling one event, delivered in order by flush() ----
  // entry words: [0] recipient | kind<<16 | flags, [1..6] payload, [7,8] pair latency (fp64 bits), [9,10] int64:
  // send offset (computation delay + additional + pipeline delay) for messages, absolute time for wakeups.  (The device keeps them as three quads in another
  // order, see emit; the word numbers are those of the host layout and of flush's local copy.)
  ABX_HD void emit(uint32_t w0, const int32_t p[6], double lat, int64_t off) {
    if (n_out >= Ctx::OUTN) { s.flags |= ABX_F_QUEUE_OVERFLOW; return; }
    if (c.onchip_writer()) {
      uint32_t *o = c.outbox() + n_out * OUT_WORDS; uint64_t lb = dbl_bits(lat);
#if defined(__CUDA_ARCH__)                                                             // 128-bit stores instead of eleven words (emit is inlined at every send site); device layout {w0, off} {lat, p0, p1} {p2..p5}:
      uint4 *o4 = reinterpret_cast<uint4 *>(o);                                         // a wakeup is its first quad alone (w0's flags are compile-time constants at every site, the test folds away)
      o4[0] = make_uint4(w0, (uint32_t)(uint64_t)off, (uint32_t)((uint64_t)off >> 32), 0u);
      if (!(w0 & OF_WAKEUP)) { o4[1] = make_uint4((uint32_t)lb, (uint32_t)(lb >> 32), (uint32_t)p[0], (uint32_t)p[1]); o4[2] = make_uint4((uint32_t)p[2], (uint32_t)p[3], (uint32_t)p[4], (uint32_t)p[5]); }
#else
      o[0] = w0; for (int i = 0; i < 6; i++) o[1 + i] = (uint32_t)p[i];
      o[7] = (uint32_t)lb; o[8] = (uint32_t)(lb >> 32);
      o[9] = (uint32_t)(uint64_t)off; o[10] = (uint32_t)((uint64_t)off >> 32);
#endif
    }
    n_out++;
  }
  // Kernel.setWakeup (Kernel.py:435-462)
  ABX_HD void set_wakeup(int sender, int64_t t) { int32_t p[6] = {0, 0, 0, 0, 0, 0}; emit((uint32_t)sender | OF_WAKEUP, p, 0.0, t); }
  // ExchangeAgent.sendMessage (agent/ExchangeAgent.py:471-485) -> Kernel.sendMessage
  ABX_HD void exch_send(int recipient, int kind, const int32_t p[6], double lat_to_recipient) {
    int64_t delay = (kind == ABX_ORDER_ACCEPTED || kind == ABX_ORDER_CANCELLED || kind == ABX_ORDER_EXECUTED) ? P.c.exchange_pipeline_delay_ns : 0;
    emit((uint32_t)recipient | ((uint32_t)kind << 16) | OF_FROM_EXCH, p, lat_to_recipient, s.exch_comp_delay + addl_delay + delay);
  }
  ABX_HD void exch_send_order(int recipient, int kind, uint32_t oid, int32_t price, int32_t qty, int32_t fill, int is_buy, double lat) {
    int32_t p[6] = {(int32_t)oid, price, qty, fill, is_buy, 0}; exch_send(recipient, kind, p, lat);
  }
  // Agent.sendMessage (agent/Agent.py:148-149) from the trader being handled, always to the exchange
  ABX_HD void ta_send(int kind, const int32_t p[6], bool bump_uniq) {
    emit(0u | ((uint32_t)kind << 16) | (bump_uniq ? OF_BUMP_UNIQ : 0u), p, a.lat_to, P.c.default_computation_delay_ns + addl_delay);
  }
  // Latency-noise draw of one sendMessage (Kernel.py:410-412): word (i & 3) of block (i >> 2) of the kernel stream, i = the draw's index.
  ABX_HD int64_t kernel_noise() {
#if defined(__CUDA_ARCH__) && !defined(ABX_SEQ_DRAWS)
    // The 32 lanes hold 32 CONSECUTIVE blocks (lane L: block (i >> 7) * 32 + L), computed in one pass of the Philox rounds -- every lane runs the rounds anyway, each
    // on its own counter -- so the ten rounds run once per 128 sends instead of once per 4 (same blocks, same words: the stream is unchanged).  The window lives in
    // s.kblk of each lane and is recomputed after a launch boundary (kb_win).
    if (!rng.tape()) {
      uint32_t i = s.ctr_kernel++, win = i >> 7;
      if (win != kb_win) {
        U4 o = philox4x32_10((win << 5) + (threadIdx.x & 31u), (uint32_t)S_KERNEL, 0x4b424958u, 0, (uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
        s.kblk[0] = o.x; s.kblk[1] = o.y; s.kblk[2] = o.z; s.kblk[3] = o.w; kb_win = win;
      }
      uint32_t w = (i & 3u) == 0u ? s.kblk[0] : ((i & 3u) == 1u ? s.kblk[1] : ((i & 3u) == 2u ? s.kblk[2] : s.kblk[3]));
      w = __shfl_sync(0xffffffffu, w, (int)((i >> 2) & 31u));
      int64_t v = (int64_t)((uint64_t(w) * (uint64_t)(uint32_t)P.c.n_noise) >> 32); rng.rec(S_KERNEL, 'i', (uint64_t)v); return v;
    }
#endif
    return rng.randint_cached(S_KERNEL, s.ctr_kernel, (uint32_t)(P.c.n_noise - 1), s.kblk);
  }
  // Kernel.sendMessage (Kernel.py:347-433) for every queued entry, in emission order
  ABX_HD void flush() {
    c.sync();
#pragma unroll 1
    for (int i = 0; i < n_out; i++) {
      const uint32_t *op = c.outbox() + i * OUT_WORDS; uint32_t o[OUT_WORDS];
#if defined(__CUDA_ARCH__)
      { const uint4 *o4 = reinterpret_cast<const uint4 *>(op); uint4 v0 = o4[0], v1 = make_uint4(0u, 0u, 0u, 0u), v2 = make_uint4(0u, 0u, 0u, 0u);
        if (!(v0.x & OF_WAKEUP)) { v1 = o4[1]; v2 = o4[2]; }
        o[0] = v0.x; o[9] = v0.y; o[10] = v0.z; o[11] = 0u; o[7] = v1.x; o[8] = v1.y; o[1] = v1.z; o[2] = v1.w; o[3] = v2.x; o[4] = v2.y; o[5] = v2.z; o[6] = v2.w; }
#else
      for (int k = 0; k < OUT_WORDS; k++) o[k] = op[k];
#endif
      uint32_t w0 = o[0]; Event e;
      for (int k = 0; k < 6; k++) e.p[k] = (int32_t)o[1 + k];
      double lat = bits_dbl((uint64_t)o[7] | ((uint64_t)o[8] << 32));
      int64_t off = (int64_t)((uint64_t)o[9] | ((uint64_t)o[10] << 32));
      e.recipient = (int)(w0 & 0xffffu); e.kind = (int)((w0 >> 16) & 0xffu);
      if (w0 & OF_WAKEUP) {
        e.t = off < KEY_T_MAX ? off : KEY_T_MAX; e.type = (w0 & OF_CANCEL_EVT) ? ABX_T_CANCEL_ORDER : ABX_T_WAKEUP; e.uniq = 0; e.sender = e.recipient; e.lat_back = 0.0; e.x0 = e.x1 = 0;
      } else {
        e.uniq = s.uniq++;                                                            // Message() construction order (message/Message.py:33-34)
        if (w0 & OF_BUMP_UNIQ) s.uniq++;                                              // TradingAgent.getCurrentSpread's never-sent msg_copy (:281)
        bool from_exch = (w0 & OF_FROM_EXCH) != 0;
        if (INSTR && from_exch && P.c.trace_cap > 0) trace_note(e.recipient, e.kind, e.p);
        int64_t sent = s.now + off;                                                   // Kernel.py:391-393
        int64_t deliver;
        if (LAT_MODEL < 0 ? P.c.latency_model == ABX_LAT_ZERO : LAT_MODEL == ABX_LAT_ZERO) { deliver = sent; }                                // zero latency matrix, noise [1.0]: no draw (ABIDESEnv.py:91-92)
        else if (LAT_MODEL < 0 ? P.c.latency_model == ABX_LAT_CUBIC : LAT_MODEL == ABX_LAT_CUBIC) {   // model/LatencyModel.py:133-138
          double u = rng.u01(S_LATENCY, s.ctr_latency);
          double x = dadd(P.c.jitter_clip, dmul(dsub(1.0, P.c.jitter_clip), u));      // uniform(low=clip, high=1.0)
          double latency = dadd(lat, dmul(P.c.jitter / pow_ni(x, 3.0), lat / P.c.jitter_unit));
          deliver = sent + (int64_t)latency;                                          // pd.Timedelta(float) truncates
        } else {                                                                      // Kernel.py:410-412
          int64_t noise = kernel_noise();
          deliver = sent + (int64_t)dadd(lat, (double)noise);
        }
        e.t = deliver < KEY_T_MAX ? deliver : KEY_T_MAX; e.type = ABX_T_MESSAGE;
        e.sender = from_exch ? 0 : self_id; e.lat_back = from_exch ? 0.0 : a.lat_from;
        e.x0 = (int32_t)o[7]; e.x1 = (int32_t)o[8];                                   // exchange replies: level-2 prices ride in the latency words
      }
      if (BOOK) continue;                                                             // bare-book replay: the notification is the output (traced above), there is no kernel queue
      if (SHAPE == SHAPE_ZI && e.recipient != 0 && e.type == ABX_T_MESSAGE) c.agent_prefetch(e.recipient);   // the reply is popped within microseconds of simulated time: have the record in L2 by then
      if (!c.q_push(e, s.now)) s.flags |= ABX_F_QUEUE_OVERFLOW;                              // Kernel.py:425 / :462
      else { s.q_count++; if (s.q_count > s.max_q) s.max_q = s.q_count; }
    }
    n_out = 0;
    c.sync();
  }

  // ---- SparseMeanRevertingOracle ----
  ABX_HD void oracle_new_megashock(int64_t from) {                                      // :67-73, :168-171
    double gap = dmul(rng.std_exponential(S_GLOBAL, s.ctr_global), P.inv_megashock_lambda);
    s.ms_t = from + (int64_t)gap;                                                       // Timedelta("{}ns".format(float)) truncates
    double msv = rng.normal(S_SYMBOL, s.ctr_symbol, P.c.megashock_mean, P.sqrt_megashock_var);
    s.ms_v = rng.randint(S_SYMBOL, s.ctr_symbol, 1) == 0 ? msv : -msv;
  }
  ABX_HD int32_t oracle_compute(int64_t ts, double v_adj, int64_t pt, int32_t pv) {     // :88-125
    double d = (double)(ts - pt);
    // exp(-2 kappa d) as the square of exp(-kappa d): it only enters through 1 - e^{-2 kappa d} ~ 2 kappa d, where an ulp of the square moves the scale of the draw by < 1e-7
    // relative -- far below what int(round(.)) of a value ~1e5 resolves -- and saves one of the two exp evaluations per oracle step (+3.7 % msgs/s, profiles/r02_optimisation_log.md)
    double ek = exp_ni(dmul(-P.c.kappa, d));
    return oracle_finish(ts, v_adj, pv, ek, rng.std_normal(S_SYMBOL, s.ctr_symbol));
  }
  // the rest of one OU step once exp(-kappa d) and the step's standard normal are known
  ABX_HD int32_t oracle_finish(int64_t ts, double v_adj, int32_t pv, double ek, double z) {
    double mu = P.c.r_bar, scale = dmul(P.ou_scale, dsub(1.0, dmul(ek, ek)));           // variance formula passed as scale
    double loc = dadd(mu, dmul(dsub((double)pv, mu), ek));
    double v = dadd(loc, dmul(scale, z));                                               // normal(loc, scale)
    v = dadd(v, v_adj); if (!(v > 0)) v = 0;
    int32_t iv = (int32_t)py_round_i64(v); s.or_t = ts; s.or_v = iv; return iv;
  }
  ABX_HD int32_t oracle_advance(int64_t t) {                                            // :131-181
    int64_t pt = s.or_t; int32_t pv = s.or_v;
    if (t <= pt) return pv;
#pragma unroll 1
    for (;;) {                                                                          // megashocks strictly before t, then t itself
      bool shock = s.ms_t < t;
      int64_t ts = shock ? s.ms_t : t;
      pv = oracle_compute(ts, shock ? s.ms_v : 0.0, pt, pv); pt = ts;
      if (!shock) break;
      oracle_new_megashock(pt);
    }
    return pv;
  }

  // ---- order book (util/OrderBook.py).  Ladders are sorted so that the BEST level is the LAST element. ----
  // Order.generateOrderId (util/order/Order.py:35-42): the smallest id >= the class counter that no Order has used yet.  Generated ids are dense; the
  // replayed stream's explicit ORDER_IDs count as used from the row that first carries them.  An explicit id the generator reaches BEFORE its first row
  // would be handed out now and collide later in the reference (two orders, one id): flagged ABX_F_ID_RANGE, the run continues with separate ids.
  ABX_HD uint32_t gen_id() {
    uint32_t id = s.next_order_id;
    if (ENV && !BOOK) {
      int4 d2 = c.day_rec2();
      if ((int32_t)id >= d2.z && (int32_t)id <= d2.w) {
        int rows_done = c.envx()->rows_done;
#pragma unroll 1
        for (;;) { int fr = c.xid_first_row(d2, (int32_t)id); if (fr < 0) break; if (fr >= rows_done) { s.flags |= ABX_F_ID_RANGE; break; } id++; }
      }
    }
    s.next_order_id = id + 1; return id;
  }
  // the exchange's event log for the realism tooling (instrumented kernels only): slot = event number mod capacity
  ABX_HD void evt_log(int kind, int32_t a_, int32_t b_) {
    if (!INSTR || !P.evt) return;
    uint4 v; v.x = (uint32_t)(uint64_t)s.now; v.y = (uint32_t)((uint64_t)s.now >> 32) | ((uint32_t)kind << 28); v.z = (uint32_t)a_; v.w = (uint32_t)b_;
    P.evt[(size_t)env_id * (size_t)P.c.event_ring_cap + (s.evt_n % (uint32_t)P.c.event_ring_cap)] = v; s.evt_n++;
  }
  ABX_HD NodeRec nload(uint32_t i) { return c.template node_load<ENV>(i); }
  ABX_HD void nstore(uint32_t i, const NodeRec &r) { c.template node_store<ENV>(i, r); }
  ABX_HD uint32_t node_alloc() {
    uint32_t n;
    if (s.free_head != NIL) { n = s.free_head; s.free_head = nload(n).next; }
    else if (s.pool_top < (uint32_t)P.c.order_cap) { n = s.pool_top++; }
    else { s.flags |= ABX_F_ORDER_OVERFLOW; return NIL; }
    return n;
  }
  ABX_HD void node_free(uint32_t n) { NodeRec r; r.id = 0; r.qty = 0; r.agent = 0; r.next = s.free_head; r.price = 0; nstore(n, r); s.free_head = n; }
  // A level's head changed to node `h`: the level now shows that order's price (util/OrderBook.py:381,393 read book[i][0].limit_price).  Only after a
  // re-pricing MODIFY can that differ from the price the level had.
  ABX_HD void level_follow_head(int side, int pos, uint32_t h) { if (ENV && (s.book_flags & BKF_REPRICED)) c.lv_setp(side, pos, nload(h).price); }

  // Per-order-id census of the book (replayed orders only): how many resting nodes carry the id and in which level.  modifyOrder's live scan
  // of a whole price level (:350-351; 220 orders on average, up to 1 380, on an IBM 2003 day -- 3.5 M dependent node loads per environment-day
  // against 165 k messages) only needs the NUMBER of nodes carrying the id in that level, which the census answers in one access unless
  // copies of the id rest in several levels (an id re-placed at another price while a head-slot copy of it survives: flag, fall back to the scan).
  ABX_HD void ib_inc(uint32_t oid, int side, int32_t price) {
    if (!ENV || oid < REPLAY_ID_BASE) return;
    uint2 v = c.ib_load((int)(oid - REPLAY_ID_BASE)); uint32_t ps = ((uint32_t)price << 1) | (uint32_t)side, cnt = v.x & 0x7fffffffu, multi = v.x >> 31;
    if (cnt == 0) { v.y = ps; multi = 0; } else if (v.y != ps) multi = 1;
    v.x = (cnt + 1) | (multi << 31); c.ib_store((int)(oid - REPLAY_ID_BASE), v);
  }
  ABX_HD void ib_dec(uint32_t oid) {
    if (!ENV || oid < REPLAY_ID_BASE) return;
    uint2 v = c.ib_load((int)(oid - REPLAY_ID_BASE)); uint32_t cnt = v.x & 0x7fffffffu;
    v.x = cnt <= 1 ? 0u : ((cnt - 1) | (v.x & 0x80000000u)); c.ib_store((int)(oid - REPLAY_ID_BASE), v);
  }
  // enterOrder :256-282: first level from the best whose slot-0 price the order beats (new level before it) or equals (joins its FIFO); worse
  // than every level -> new worst level.  lv_find returns exactly that position (for sorted ladders the ordinary sorted insert).
  ABX_HD void book_enter(int side, uint32_t oid, int agent, int32_t price, int32_t qty) {
    int n = n_lv(side); int pos; bool found;
    c.lv_find(side, price, n, pos, found);
    uint32_t node = node_alloc(); if (node == NIL) return;
    NodeRec r; r.id = oid; r.qty = qty; r.agent = (uint32_t)agent | (P3 ? (hidx << 16) : R3 ? (s.trade_epoch << 16) : 0u); r.next = NIL; r.price = price; nstore(node, r);   // rmsc03: + registration epoch; population 3: + index of its order-history log entry
    if (found) {
      uint32_t tail = c.lv_tail(side, pos);
      NodeRec tr = nload(tail); tr.next = node; nstore(tail, tr);
      c.lv_set(side, pos, c.lv_qty(side, pos) + qty, c.lv_head(side, pos), node);
    } else {
      if (n >= P.c.level_cap) { s.flags |= ABX_F_LEVEL_OVERFLOW; node_free(node); return; }
      c.lv_insert(side, pos, n, price, qty, node, node); set_n_lv(side, n + 1);
    }
    s.n_resting++; ib_inc(oid, side, price);
  }
  // handleLimitOrder :38-170 (+ executeOrder :172-240).  lat_in = latency[0][incoming agent]
  ABX_HD void book_handle_limit(uint32_t oid, int agent, int is_buy, int32_t price, int32_t qty, double lat_in) {
    if (qty <= 0) return;                                                               // :47-49
    if (ENV && oid >= REPLAY_ID_BASE) {                                                 // :52-60 history[0][order_id] = {...}
      uint4 t = c.ord_load((int)(oid - REPLAY_ID_BASE)); uint32_t d = s.trade_epoch - t.z;
      t.w = (t.w == 0 || d >= 16) ? 1u : (((t.w << d) | 1u) & 0xffffu); t.z = s.trade_epoch; c.ord_store((int)(oid - REPLAY_ID_BASE), t);
    }
    evt_log(ABX_EV_ORDER, price, is_buy ? qty : -qty);
    if (P3) hidx = s.hist_n++ & 0xffffu;                                                // :52-60 history[0][order_id] = {...}: the entry is written below, once "transactions" is known
    int opp = is_buy ? 1 : 0;                                                           // a buy matches asks (side 1)
    uint32_t epoch0 = s.trade_epoch;                                                    // the incoming order's history bucket (:52-60)
    int64_t trade_qty = 0, trade_px = 0;
    bool matching = true;
#pragma unroll 1
    while (matching) {                                                                  // :68-110
      int n = n_lv(opp); bool matched = false;
      if (n > 0) {
        int32_t bp = c.lv_price(opp, n - 1);
        if (is_buy ? price >= bp : price <= bp) {                                       // isMatch :242-254, head of best level only
          uint32_t h = c.lv_head(opp, n - 1); NodeRec hr = nload(h);
          int32_t fq;
          if (ENV && hr.id >= REPLAY_ID_BASE) c.id_prefetch((int)(hr.id - REPLAY_ID_BASE));   // its owner updates the record when ORDER_EXECUTED arrives
          if (qty >= hr.qty) {                                                          // :204-210 whole resting order consumed
            fq = hr.qty;
            if (hr.next == NIL) set_n_lv(opp, n - 1);                                    // level emptied: it is the last element
            else { c.lv_set(opp, n - 1, c.lv_qty(opp, n - 1) - fq, hr.next, c.lv_tail(opp, n - 1)); level_follow_head(opp, n - 1, hr.next); }
            node_free(h); s.n_resting--; ib_dec(hr.id);
          } else {                                                                      // :212-217 partial
            fq = qty; hr.qty -= fq; nstore(h, hr); c.lv_set(opp, n - 1, c.lv_qty(opp, n - 1) - fq, h, c.lv_tail(opp, n - 1));
          }
          if (P3) {                                                                     // :230-237 the resting order's record gets a transaction if one of the retained buckets still holds it
            uint32_t sl = (hr.agent >> 16) & (uint32_t)(P.c.hist_log_cap - 1); uint4 h4 = c.hist_load((int)sl);
            if (h4.x == hr.id && epoch0 - h4.z <= (uint32_t)P.c.stream_history && !(h4.w & 2u)) { h4.w |= 2u; c.hist_store((int)sl, h4); }
          } else if (R3) {                                                              // :227,230-237 history "transactions" tuples read back by get_transacted_volume
            tv_record(qty, epoch0);                                                     //   incoming order: its PRE-fill remaining quantity
            uint32_t re = hr.agent >> 16; if (((epoch0 - re) & 0xffffu) <= (uint32_t)P.c.stream_history) tv_record(fq, (epoch0 & 0xffff0000u) | re);   // resting order, if its bucket survives
          }
          qty -= fq;                                                                    // :77
          exch_send_order(agent, ABX_ORDER_EXECUTED, oid, price, fq, bp, is_buy, lat_in);             // :88 (incoming copy)
          exch_send_order((int)(hr.agent & 0xffffu), ABX_ORDER_EXECUTED, hr.id, bp, fq, bp, !is_buy, (ENV || (R3 && (!P3 || lat_zero()))) ? 0.0 : c.agent_lat_from(P3 ? (int)(hr.agent & 0xffffu) : (int)hr.agent)); // :89-91
          trade_qty += fq; trade_px += (int64_t)bp * fq; s.c_fills++; matched = true;
          if (qty <= 0) matching = false;
          if (n_out >= Ctx::OUTN - 3) flush(); else c.sync();
        }
      }
      if (!matched) {
        book_enter(is_buy ? 0 : 1, oid, agent, price, qty);                             // :101
        exch_send_order(agent, ABX_ORDER_ACCEPTED, oid, price, qty, 0, is_buy, lat_in); // :108
        matching = false;
      }
    }
    if (P3) { uint4 h4; h4.x = oid; h4.y = (uint32_t)price; h4.z = epoch0; h4.w = (is_buy ? 1u : 0u) | (trade_qty > 0 ? 2u : 0u); c.hist_store((int)(hidx & (uint32_t)(P.c.hist_log_cap - 1)), h4); }   // :52-60 + :227
    if (INSTR && P.evt) {                                                               // :114-128 BEST_BID / BEST_ASK of the book as the order left it
      if (s.n_bid_lv > 0) evt_log(ABX_EV_BEST_BID, c.lv_price(0, s.n_bid_lv - 1), c.lv_qty(0, s.n_bid_lv - 1));
      if (s.n_ask_lv > 0) evt_log(ABX_EV_BEST_ASK, c.lv_price(1, s.n_ask_lv - 1), c.lv_qty(1, s.n_ask_lv - 1));
    }
    if (trade_qty > 0) { s.last_trade = (int32_t)py_round_i64((double)trade_px / (double)trade_qty); s.trade_epoch++; evt_log(ABX_EV_LAST_TRADE, s.last_trade, (int32_t)trade_qty); } // :131-149 (history.insert(0, {}))
  }
  // cancelOrder :284-339: levels from the best whose slot-0 price equals the request's, the first that holds the id
  ABX_HD void book_cancel(uint32_t oid, int agent, int is_buy, int32_t price, double lat_in) {
    cancel_found = false;
    int side = is_buy ? 0 : 1; int n = n_lv(side); if (n == 0) return;
    int limit = n;
#pragma unroll 1
    for (;;) {
      int cnt; int pos = c.lv_find_eq(side, price, limit, cnt);
      if (pos < 0) return;
      uint32_t prev = NIL, cur = c.lv_head(side, pos);
#pragma unroll 1
      while (cur != NIL) {
        NodeRec r = nload(cur);
        if (r.id == oid) {
          if (prev == NIL) {
            if (r.next == NIL) { c.lv_remove(side, pos, n); set_n_lv(side, n - 1); }
            else { c.lv_set(side, pos, c.lv_qty(side, pos) - r.qty, r.next, c.lv_tail(side, pos)); level_follow_head(side, pos, r.next); }
          } else {
            NodeRec pr = nload(prev); pr.next = r.next; nstore(prev, pr);
            c.lv_set(side, pos, c.lv_qty(side, pos) - r.qty, c.lv_head(side, pos), r.next == NIL ? prev : c.lv_tail(side, pos));
          }
          node_free(cur); s.n_resting--; ib_dec(r.id); cancel_found = true;
          exch_send_order(agent, ABX_ORDER_CANCELLED, r.id, ENV ? r.price : price, r.qty, 0, is_buy, lat_in);   // :334-336 the BOOK's copy of the order, to the REQUEST's agent
          return;
        }
        prev = cur; cur = r.next;
      }
      if (cnt <= 1) return;                                                             // no other level shows that price (always, while the ladder is sorted)
      limit = pos;
    }
  }

  // ---- ExchangeAgent.receiveMessage :129-340 ----
  ABX_HD void exch_receive(const Event &m) {
    s.exch_comp_delay = P.c.exchange_computation_delay_ns;                              // :139
    bool t_closed = s.now > P.c.mkt_close_ns;
    double lat = m.lat_back;                                                            // latency[0][sender]
    int32_t p[6] = {0, 0, 0, 0, 0, 0}; int reply = ABX_NONE;                            // the direct reply, sent from ONE site below (every inlined send is ~12 instructions of loop body)
    bool is_order = m.kind == ABX_LIMIT_ORDER || m.kind == ABX_CANCEL_ORDER || m.kind == ABX_MODIFY_ORDER;
    bool is_query = m.kind == ABX_QUERY_SPREAD || m.kind == ABX_QUERY_LAST_TRADE || m.kind == ABX_QUERY_TRANSACTED_VOLUME || m.kind == ABX_QUERY_ORDER_STREAM;
    if (t_closed && (is_order || !is_query)) reply = ABX_MKT_CLOSED;                    // :142-160
    else if (m.kind == ABX_WHEN_MKT_OPEN || m.kind == ABX_WHEN_MKT_CLOSE) { s.exch_comp_delay = 0; reply = m.kind; }   // :175-192
    else if (m.kind == ABX_QUERY_SPREAD) {                                              // :215-245 (depth 1)
      s.c_query++; int f = 0;
      int nb = s.n_bid_lv, na = s.n_ask_lv;
      if (nb > 0) { p[0] = c.lv_price(0, nb - 1); p[1] = c.lv_qty(0, nb - 1); f |= 1; }
      if (na > 0) { p[2] = c.lv_price(1, na - 1); p[3] = c.lv_qty(1, na - 1); f |= 2; }
      if (t_closed) f |= 4;
      p[4] = s.last_trade; p[5] = f; reply = ABX_QUERY_SPREAD;
    } else if (m.kind == ABX_LIMIT_ORDER) {                                             // :304-312
      s.c_limit++; book_handle_limit((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], m.p[2], lat); c.sync(); trace_snap();
    } else if (m.kind == ABX_CANCEL_ORDER) {                                            // :313-325
      s.c_cancel++; book_cancel((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], lat); c.sync(); trace_snap();
    }
    if (reply != ABX_NONE) exch_send(m.sender, reply, p, lat);
  }

  // ---- TradingAgent / ZeroIntelligenceAgent (the trader is in `a` / `z`) ----
  ABX_HD void ta_get_spread() { int32_t p[6] = {0, 0, 0, 0, 0, 0}; ta_send(ABX_QUERY_SPREAD, p, true); }   // TradingAgent.getCurrentSpread :277-282
  // ZeroIntelligenceAgent.wakeup :125-187 (+ TradingAgent.wakeup :142-158)
  ABX_HD void zi_wakeup(int id) {
    if (!(a.flags & AF_HAS_OPEN)) {                                                     // TradingAgent.py:149-153
      int32_t p[6] = {0, 0, 0, 0, 0, 0};
      ta_send(ABX_WHEN_MKT_OPEN, p, false); ta_send(ABX_WHEN_MKT_CLOSE, p, false);
    }
    uint32_t st = ST_INACTIVE;
    if ((a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE) && !((a.flags & AF_MKT_CLOSED) && (a.flags & AF_HAS_DAILY))) {   // :130-147
      double delta_time = dmul(rng.std_exponential(S_AGENT0 + id, a.rng_ctr), P.inv_lambda_a);       // :157
      set_wakeup(id, s.now + py_round_i64(delta_time));                                              // :158
      if (!((a.flags & AF_MKT_CLOSED) && !(a.flags & AF_HAS_DAILY))) {                  // :162-166 skip cancels when waiting for the close price
#pragma unroll 1
        for (int i = 0; i < a.n_orders; i++) {                                          // cancelOrders :336-344
          int32_t q = z->oqty[i];
          int32_t p[6] = {(int32_t)z->oid[i], z->oprice[i], q < 0 ? -q : q, 0, q > 0, 0};
          ta_send(ABX_CANCEL_ORDER, p, false);
        }
      }
      if (P3 && ((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT) == AT_HBL && !((a.flags & AF_MKT_CLOSED) && !(a.flags & AF_HAS_DAILY))) {   // HeuristicBeliefLearningAgent.wakeup :61-74: the ZI wakeup left it "ACTIVE"
        int32_t p[6] = {P.c.hbl_L, 0, 0, 0, 0, 0}; ta_send(ABX_QUERY_ORDER_STREAM, p, false); st = ST_AWAITING_STREAM;                    // getOrderStream(length = L)
      } else { ta_get_spread(); st = ST_AWAITING_SPREAD; }                              // :164 / :183-185
    }
    a.flags = (a.flags & ~AF_STATE_MASK) | (st << AF_STATE_SHIFT);
  }
  // ZeroIntelligenceAgent.placeOrder :277-309 (+ updateEstimates :189-275, TradingAgent.placeLimitOrder :309-349)
  ABX_HD void zi_place_order(int id, bool hbl = false) {                                // hbl: HeuristicBeliefLearningAgent.placeOrder with a full order stream (population 3)
    int stream = S_AGENT0 + id;
    const int64_t t_obs = s.now >= P.c.mkt_close_ns ? P.c.mkt_close_ns - 1 : s.now;     // observePrice :210-227
    // Philox mode: the three draws of one order placement (noisy observation, side, surplus R) come from ONE Philox block of the agent's stream
    // (words 0-2 the normal, bit 0 of word 2 the side, word 3 R) instead of three blocks; tape mode replays the reference's draws one by one.
    bool one_block = !rng.tape(); U4 blk; blk.x = blk.y = blk.z = blk.w = 0;
    int32_t r_now; double z_obs;
#if defined(__CUDA_ARCH__) && !defined(ABX_SEQ_DRAWS)
    double pw0_l = 0.0, pw2_l = 0.0;
    if (one_block) {
      // Lane-parallel evaluation: every lane of the warp runs the same scalar code anyway, so the independent chains of one order placement go to different lanes and
      // run ONCE -- lane 0 the fundamental's OU step (Philox block of the symbol stream -> Box-Muller, exp(-kappa d)), lane 1 the agent's block (noisy observation,
      // side, R) and (1 - kappa) ** delta, lane 2 (1 - kappa) ** d2: one pass of the Philox rounds, of Box-Muller and of exp instead of two, two and three.  Same
      // functions on the same arguments as the sequential form.  A megashock due before the observation is one more trip through the same code (lane 0 steps to
      // the shock; the other lanes' results are simply recomputed on the last trip).
      const uint32_t ln = threadIdx.x & 31u; const bool need = t_obs > s.or_t;        // oracle_advance returns the stored value for t <= r[symbol][0]
      const int64_t pwk = (a.flags & AF_HAS_PREV) ? a.prev_wake : P.c.mkt_open_ns;
      const double dl = (double)(s.now - pwk); double dc = (double)(P.c.mkt_close_ns - s.now); if (!(dc > 0)) dc = 0;
      r_now = s.or_v; uint32_t zlo, zhi, elo, ehi; U4 o;
#pragma unroll 1
      for (;;) {
        const bool shock = need && s.ms_t < t_obs; const int64_t ts = shock ? s.ms_t : t_obs;
        o = philox4x32_10(ln == 0u ? s.ctr_symbol : a.rng_ctr, ln == 0u ? (uint32_t)S_SYMBOL : (uint32_t)stream, 0x41424958u, 0, (uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
        const uint64_t zb = dbl_bits(box_muller(o.x, o.y, o.z));
        const uint64_t eb = dbl_bits(exp_ni(ln == 0u ? dmul(-P.c.kappa, (double)(ts - s.or_t)) : dmul(ln == 1u ? dl : dc, P.log_base_a)));
        zlo = (uint32_t)zb; zhi = (uint32_t)(zb >> 32); elo = (uint32_t)eb; ehi = (uint32_t)(eb >> 32);
        if (need) {
          const double z_or = bits_dbl((uint64_t)__shfl_sync(0xffffffffu, zlo, 0) | ((uint64_t)__shfl_sync(0xffffffffu, zhi, 0) << 32));
          const double ek = bits_dbl((uint64_t)__shfl_sync(0xffffffffu, elo, 0) | ((uint64_t)__shfl_sync(0xffffffffu, ehi, 0) << 32));
          s.ctr_symbol++; rng.rec(S_SYMBOL, 'n', dbl_bits(z_or));
          r_now = oracle_finish(ts, shock ? s.ms_v : 0.0, s.or_v, ek, z_or);
        }
        if (!shock) break;
        oracle_new_megashock(ts);
      }
      a.rng_ctr++;
      z_obs = bits_dbl((uint64_t)__shfl_sync(0xffffffffu, zlo, 1) | ((uint64_t)__shfl_sync(0xffffffffu, zhi, 1) << 32));
      pw0_l = bits_dbl((uint64_t)__shfl_sync(0xffffffffu, elo, 1) | ((uint64_t)__shfl_sync(0xffffffffu, ehi, 1) << 32));
      pw2_l = bits_dbl((uint64_t)__shfl_sync(0xffffffffu, elo, 2) | ((uint64_t)__shfl_sync(0xffffffffu, ehi, 2) << 32));
      blk.z = __shfl_sync(0xffffffffu, o.z, 1); blk.w = __shfl_sync(0xffffffffu, o.w, 1);
    } else
#endif
    {
      r_now = oracle_advance(t_obs);
      if (one_block) blk = rng.philox(stream, a.rng_ctr);
      z_obs = one_block ? box_muller(blk.x, blk.y, blk.z) : rng.std_normal(stream, a.rng_ctr);
    }
    if (one_block) rng.rec(stream, 'n', dbl_bits(z_obs));                               // draw log: the same three entries a tape would hold
    int32_t obs_t = (int32_t)py_round_i64(dadd((double)r_now, dmul(P.sqrt_sigma_n, z_obs)));
    int q = a.shares / 100;                                                             // :203 int(x / 100): the correctly rounded fp64 quotient of two ints truncates to the C quotient (|x| < 2^31, remainder / 100 <= 0.99)
    int q_max = P.c.q_max; bool buy;
    if (q >= q_max) buy = false; else if (q <= -q_max) buy = true; else { buy = one_block ? (blk.z & 1u) != 0 : rng.randint(stream, a.rng_ctr, 1) != 0; if (one_block) rng.rec(stream, 'i', buy ? 1u : 0u); } // :205-213
    if (!(a.flags & AF_HAS_PREV)) { a.prev_wake = P.c.mkt_open_ns; a.flags |= AF_HAS_PREV; }         // :217-218
    double r_bar = P.c.r_bar, sigma_n = P.c.sigma_n;
    double delta = (double)(s.now - a.prev_wake);                                       // :221
    double d2 = (double)(P.c.mkt_close_ns - s.now); if (!(d2 > 0)) d2 = 0;              // :251
    // (1 - kappa) ** x evaluated as exp(x * log(1 - kappa)) with the logarithm precomputed on the host (log1p): both are
    // within an ulp of the true power; the result only feeds int(round(.)) of values ~1e5.
    // (1 - kappa) ** (2 delta) as the square of (1 - kappa) ** delta: it only enters through 1 - pw1 ~ 2 delta kappa ~ 1e-3 (sigma_t stays 0, :242), so an
    // ulp of pw1 moves r_t by ~1e-11 -- and saves one of the seven exp evaluations per order (12 % of all executed instructions were exp).
#if defined(__CUDA_ARCH__) && !defined(ABX_SEQ_DRAWS)
    double pw0 = one_block ? pw0_l : exp_ni(dmul(delta, P.log_base_a)), pw1 = dmul(pw0, pw0), pw2 = one_block ? pw2_l : exp_ni(dmul(d2, P.log_base_a));
#else
    double pw0 = exp_ni(dmul(delta, P.log_base_a)), pw1 = dmul(pw0, pw0), pw2 = exp_ni(dmul(d2, P.log_base_a));
#endif
    double r_tprime = dmul(dsub(1.0, pw0), r_bar);                                    // :229
    r_tprime = dadd(r_tprime, dmul(pw0, a.r_t));                                      // :230
    double sigma_tprime = dmul(pw1, a.sigma_t);                                       // :233
    sigma_tprime = dadd(sigma_tprime, dmul(dsub(1.0, pw1) / P.sigma_denom, P.c.sigma_s)); // :234
    double den = dadd(sigma_n, sigma_tprime);
    double r_t = dmul(sigma_n / den, r_tprime);                                         // :239
    r_t = dadd(r_t, dmul(sigma_tprime / den, (double)obs_t));                           // :240
    a.r_t = r_t;
    if (a.sigma_t != 0.0)                                                               // sigma_t starts at 0 and 0 * sigma_n / (sigma_n + 0) is exactly 0: the division only runs if a caller seeds it otherwise
    a.sigma_t = dmul(sigma_n, a.sigma_t) / dadd(sigma_n, a.sigma_t);                    // :242
    double r_T = dmul(dsub(1.0, pw2), r_bar);                                         // :255
    r_T = dadd(r_T, dmul(pw2, a.r_t));                                                // :256
    int32_t r_Ti = (int32_t)py_round_i64(r_T);                                          // :259
    a.prev_wake = s.now;                                                                // :262
    q += q_max - 1;                                                                     // :267
    int idx = buy ? q + 1 : q, n = 2 * q_max;                                           // :268 (Python list indexing)
    if (idx < 0) idx += n;
    if (idx < 0 || idx >= n) { s.flags |= ABX_F_THETA_INDEX; idx = idx < 0 ? 0 : n - 1; }
    int32_t v = r_Ti + (int32_t)z->theta[idx];                                          // :270
    if (P3 && hbl) { hbl_place_order(v, buy); return; }
    // placeOrder
    int grp = (a.flags & AF_GROUP_MASK) >> AF_GROUP_SHIFT;
    int32_t r_min = P.c.groups[grp].r_min, r_max = P.c.groups[grp].r_max; double eta = P.c.groups[grp].eta;
    int32_t R = r_min + (one_block ? (int32_t)((uint64_t(blk.w) * (uint64_t)(uint32_t)(r_max - r_min + 1)) >> 32) : (int32_t)rng.randint(stream, a.rng_ctr, (uint32_t)(r_max - r_min)));   // :284
    if (one_block) rng.rec(stream, 'i', (uint64_t)(uint32_t)(R - r_min));
    int32_t p = buy ? v - R : v + R;                                                    // :287
    int32_t ask_vol = (a.flags & AF_HAS_ASK) ? a.ask_q : 0, bid_vol = (a.flags & AF_HAS_BID) ? a.bid_q : 0;
    if (buy && ask_vol > 0) { int32_t R_ask = v - a.ask; if ((double)R_ask >= dmul(eta, (double)R)) p = a.ask; }             // :291-297
    else if (!buy && bid_vol > 0) { int32_t R_bid = a.bid - v; if ((double)R_bid >= dmul(eta, (double)R)) p = a.bid; }       // :298-305
    zi_submit(buy, p);
  }
  // TradingAgent.placeLimitOrder :309-349 for the staged ZI / HBL trader
  ABX_HD void zi_submit(bool buy, int32_t p) {
    uint32_t oid = s.next_order_id++;                                                   // util/order/Order.py:27,35-42
    int32_t size = P.c.order_size;
    if (size > 0) {
      if (a.n_orders < AGENT_ORDER_CAP) {
        if (c.onchip_writer()) { int k = a.n_orders; z->oid[k] = oid; z->oprice[k] = p; z->oqty[k] = buy ? size : -size; }
        a.n_orders++;
      } else s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW;
      int32_t pl[6] = {(int32_t)oid, p, size, 0, buy, 0};
      ta_send(ABX_LIMIT_ORDER, pl, false);                                              // :343
    }
  }
  // HeuristicBeliefLearningAgent.placeOrder :98-185 once v and the side are known: the belief Pr(an order at price p transacts) from the orders in the
  // `n` history buckets the exchange handed out (epochs E_q - n .. E_q - 1 of the order-history log, read NOW: the reference holds references to the
  // exchange's dicts), limit price = first argmax of Pr(p) * surplus(p) over every price between the lowest and the highest order in them.
  ABX_HD void hbl_place_order(int32_t v, bool buy) {
    uint32_t E_q = (uint32_t)(uint64_t)z->surplus, n = (uint32_t)((uint64_t)z->surplus >> 32);
    int32_t best_p = 0; uint32_t err = 0;
    bool place = c.hbl_best(s.hist_n, E_q - n, E_q - 1, buy, v, best_p, err);
    s.flags |= err;
    if (place) zi_submit(buy, best_p);
  }
  // index of order_id in self.orders, or -1
  ABX_HD int orders_find(uint32_t oid) { int f = -1; for (int i = 0; i < AGENT_ORDER_CAP; i++) if (i < a.n_orders && f < 0 && z->oid[i] == oid) f = i; return f; }
  ABX_HD void orders_remove(int i) {
    c.sync();
    if (c.onchip_writer()) for (int k = i; k + 1 < AGENT_ORDER_CAP; k++) { z->oid[k] = z->oid[k + 1]; z->oprice[k] = z->oprice[k + 1]; z->oqty[k] = z->oqty[k + 1]; }
    a.n_orders--;
    c.sync();
  }
  // TradingAgent.receiveMessage :181-268 + ZeroIntelligenceAgent.receiveMessage :311-334
  ABX_HD void zi_receive(int id, const Event &m) {
    bool had = (a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE);
    if (m.kind == ABX_WHEN_MKT_OPEN) a.flags |= AF_HAS_OPEN;
    else if (m.kind == ABX_WHEN_MKT_CLOSE) a.flags |= AF_HAS_CLOSE;
    else if (m.kind == ABX_ORDER_EXECUTED) {                                            // orderExecuted :422-462
      int32_t q = m.p[2]; int32_t sq = m.p[4] ? q : -q;
      a.shares += sq; a.cash -= (int64_t)sq * m.p[3];
      int i = orders_find((uint32_t)m.p[0]);
      if (i >= 0) {
        int32_t oq0 = z->oqty[i]; int32_t oq = oq0 < 0 ? -oq0 : oq0;
        if (q >= oq) orders_remove(i);
        else { c.sync(); if (c.onchip_writer()) z->oqty[i] = oq0 < 0 ? -(oq - q) : (oq - q); c.sync(); }
      }
    } else if (m.kind == ABX_ORDER_CANCELLED) {                                         // orderCancelled :476-489
      int i = orders_find((uint32_t)m.p[0]); if (i >= 0) orders_remove(i);
    } else if (m.kind == ABX_MKT_CLOSED) a.flags |= AF_MKT_CLOSED;                      // marketClosed :492-499
    else if (m.kind == ABX_QUERY_SPREAD) {                                              // :232-238, querySpread :514-537
      if (m.p[5] & 4) a.flags |= AF_MKT_CLOSED;
      a.last_trade = m.p[4]; a.flags |= AF_HAS_LAST;
      if (a.flags & AF_MKT_CLOSED) { a.daily_close = a.last_trade; a.flags |= AF_HAS_DAILY; }
      a.flags &= ~(AF_HAS_BID | AF_HAS_ASK);
      if (m.p[5] & 1) { a.flags |= AF_HAS_BID; a.bid = m.p[0]; a.bid_q = m.p[1]; } else { a.bid = 0; a.bid_q = 0; }
      if (m.p[5] & 2) { a.flags |= AF_HAS_ASK; a.ask = m.p[2]; a.ask_q = m.p[3]; } else { a.ask = 0; a.ask_q = 0; }
    } else if (P3 && m.kind == ABX_QUERY_ORDER_STREAM) {                                // :240-246, queryOrderStream :549-554: self.stream_history[symbol] = the references
      if (m.p[5] & 4) a.flags |= AF_MKT_CLOSED;
      c.sync(); if (c.onchip_writer()) z->surplus = (int64_t)((uint64_t)(uint32_t)m.p[0] | ((uint64_t)(uint32_t)m.p[1] << 32)); c.sync();
    }
    if ((a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE) && !had) {                  // :258-268
      int64_t off = rng.randint(S_AGENT0 + id, a.rng_ctr, 99);                          // ZI.getWakeFrequency :349-350
      set_wakeup(id, P.c.mkt_open_ns + off);
    }
    uint32_t st = (a.flags & AF_STATE_MASK) >> AF_STATE_SHIFT;
    if (st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD && !(a.flags & AF_MKT_CLOSED)) {      // ZI :319-334
      if (P3) zi_place_order(id, ((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT) == AT_HBL && (uint32_t)((uint64_t)z->surplus >> 32) >= (uint32_t)P.c.hbl_L);   // HBL :83-87: fewer than L buckets -> the ZI order
      else zi_place_order(id);
      a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT);
    }
    if (P3 && st == ST_AWAITING_STREAM && m.kind == ABX_QUERY_ORDER_STREAM && ((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT) == AT_HBL && !(a.flags & AF_MKT_CLOSED)) {   // HBL.receiveMessage :176-195
      ta_get_spread(); a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_SPREAD << AF_STATE_SHIFT);
    }
  }

  // ---- reset of one environment: oracle __init__ (first megashock) and Kernel.runner :154-175 (one WAKEUP per agent) ----
  ABX_HD void reset_env() {
    oracle_new_megashock(P.c.mkt_open_ns);                                              // SparseMeanRevertingOracle.py:67-73
#pragma unroll 1
    for (int id = 0; id < P.c.n_agents; id++) {                                         // Agent.kernelStarting :78
      set_wakeup(id, P.c.start_ns);
      if (n_out >= Ctx::OUTN - 1) flush();
    }
    flush();
    rng_sync();
  }

  // ---- Kernel.runner hot loop :190-292 ----
  ABX_HD void run(int64_t until) {
#pragma unroll 1
    while (!(s.flags & ABX_F_DONE)) {
      uint64_t khi; uint32_t kuniq; int grp;
      bool any = c.q_min(khi, kuniq, grp);
      if (!any || !(s.now <= P.c.stop_ns) || (rng.tape() && (rng.err & ABX_F_TAPE_UNDERRUN))) { s.flags |= ABX_F_DONE; break; }   // an exhausted tape hands out zeros: with zero delays an agent would re-wake itself at the same instant for ever            // :190 tested BEFORE the pop
      if (key_time(khi) > until) break;
      { int rid = key_recipient(khi); if (rid != 0) c.agent_load_issue(rid); }          // the recipient's record is on its way while the event is unpacked
      Event ev; c.q_fetch(grp, ev);                                                     // :192
      s.now = ev.t; s.ttl++;                                                            // :211
      if (INSTR && P.c.hash_pops) s.pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s.pop_hash, ev.t), ev.recipient), ev.type), ev.type == ABX_T_MESSAGE ? (int64_t)ev.uniq : -1);
      if (INSTR && P.c.trace_cap > 0) {
        abx_trace_rec r; r.tag = 0; r.a = ev.recipient; r.t = ev.t; for (int i = 0; i < 16; i++) r.v[i] = 0;
        r.v[0] = ev.type; r.v[1] = ev.type == ABX_T_MESSAGE ? (int32_t)ev.uniq : -1; r.v[2] = ev.kind; trace_rec(r);
      }
      addl_delay = 0;                                                                   // :214
      int id = ev.recipient;
      if (id == 0) {
        if (s.exch_time > s.now) { c.q_requeue(s.exch_time); continue; }                // :224-230 / :258-264 (same uniq)
        c.q_remove(); s.q_count--;
        s.exch_time = s.now;                                                            // :234 / :268
        self_id = 0;
        if (ev.type == ABX_T_MESSAGE) exch_receive(ev);                                 // exchange WAKEUP: Agent.wakeup is a no-op
        s.exch_time += s.exch_comp_delay + addl_delay;                                  // :240-242 / :274-276
      } else {
        z = c.agent_stage_issued(id); regs_load(a, z);
        if (a.agent_time > s.now) { c.q_requeue(a.agent_time); continue; }
        c.q_remove(); s.q_count--;
        a.agent_time = s.now; self_id = id;
        if (ev.type == ABX_T_WAKEUP) zi_wakeup(id); else zi_receive(id, ev);
        a.agent_time += P.c.default_computation_delay_ns + addl_delay;
        c.sync();
        if (c.onchip_writer()) regs_store(z, a);
        c.sync();
        c.agent_commit(id);
      }
      flush();                                                                          // deliver what this event sent
    }
    rng_sync();
  }


  // =================================================================================================
  // ABIDESEnv shape: GymKernel.stepRunner (GymKernel.py:158-306) over Exchange (id 0), MarketReplayAgent (id 1,
  // agent/examples/MarketReplayAgent.py) and DummyRLExecutionAgent (id 2, agent/execution/rl/dummy_rl_execution_agent.py,
  // agent/execution/baselines/execution_agent.py, ABIDESEnvMetrics.py).  Latency 0, computation delay 0, no oracle.
  // =================================================================================================
  ABX_HD void env_send(int kind, const int32_t p[6], bool bump) {                       // Agent.sendMessage from trader self_id to the exchange
    emit(0u | ((uint32_t)kind << 16) | (bump ? OF_BUMP_UNIQ : 0u), p, P3 ? a.lat_to : 0.0, P.c.default_computation_delay_ns + addl_delay);   // population 3 may run under a latency matrix (config/rmsc02.py)
  }
  ABX_HD void env_set_cancel(int sender, int64_t t) { int32_t p[6] = {0, 0, 0, 0, 0, 0}; emit((uint32_t)sender | OF_WAKEUP | OF_CANCEL_EVT, p, 0.0, t); }   // GymKernel.setCancelOrder :364-389
  ABX_HD bool ta_wakeup(uint32_t flags) {                                               // TradingAgent.wakeup :142-158 -> can_trade
    if (!(flags & AF_HAS_OPEN)) { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_WHEN_MKT_OPEN, p, false); env_send(ABX_WHEN_MKT_CLOSE, p, false); }
    return (flags & AF_HAS_OPEN) && (flags & AF_HAS_CLOSE) && !(flags & AF_MKT_CLOSED);
  }
  // TradingAgent.receiveMessage :181-268 common part; true when the market hours just became known
  ABX_HD bool ta_receive(const Event &m, uint32_t &flags, int32_t &shares, int64_t &cash, int32_t &last_trade) {
    bool had = (flags & AF_HAS_OPEN) && (flags & AF_HAS_CLOSE);
    if (m.kind == ABX_WHEN_MKT_OPEN) flags |= AF_HAS_OPEN;
    else if (m.kind == ABX_WHEN_MKT_CLOSE) flags |= AF_HAS_CLOSE;
    else if (m.kind == ABX_ORDER_EXECUTED) { int32_t sq = m.p[4] ? m.p[2] : -m.p[2]; shares += sq; cash -= (int64_t)sq * m.p[3]; }
    else if (m.kind == ABX_MKT_CLOSED) flags |= AF_MKT_CLOSED;
    else if (m.kind == ABX_QUERY_SPREAD) { if (m.p[5] & 4) flags |= AF_MKT_CLOSED; last_trade = m.p[4]; flags |= AF_HAS_LAST; }
    return (flags & AF_HAS_OPEN) && (flags & AF_HAS_CLOSE) && !had;
  }
  // ---- exchange additions: MODIFY_ORDER (util/OrderBook.py:341-372) ----
  ABX_HD void book_modify(uint32_t oid, int agent, int is_buy, int32_t price, int32_t new_price, int32_t new_qty) {
    int side = is_buy ? 0 : 1; int n = n_lv(side); if (n == 0) return;                  // :345-347
    int buckets = 0;                                                                    // :353-367 one ORDER_MODIFIED per history bucket holding the id
    if (oid >= REPLAY_ID_BASE) { uint4 t = c.ord_load((int)(oid - REPLAY_ID_BASE)); uint32_t d = s.trade_epoch - t.z; if (t.w != 0 && d <= (uint32_t)P.c.stream_history) buckets = __popc_compat(t.w & ((1u << (P.c.stream_history + 1 - d)) - 1u)); }
    int limit = n;
#pragma unroll 1
    for (;;) {                                                                          // :348-349 every level, best first, whose slot-0 price equals the OLD order's price
      int cnt; int pos = c.lv_find_eq(side, price, limit, cnt);
      if (pos < 0) return;
      uint32_t head = c.lv_head(side, pos); int matches = 0; bool scan = true;
      if (oid >= REPLAY_ID_BASE && !(s.book_flags & BKF_REPRICED)) {                    // :350-351 "every node of the level carrying the id": from the census when it can tell
        uint2 v = c.ib_load((int)(oid - REPLAY_ID_BASE)); uint32_t k = v.x & 0x7fffffffu;
        if (k == 0) return;
        if (!(v.x >> 31)) { if (v.y != (((uint32_t)price << 1) | (uint32_t)side)) return; matches = (int)k; scan = false; }
      }
      if (scan) {
        uint32_t cur = head;
#pragma unroll 1
        while (cur != NIL) { NodeRec r = nload(cur); if (r.id == oid) matches++; cur = r.next; }   // live scan of the level
      }
      if (matches > 0) {
        NodeRec hr = nload(head);
        c.lv_set(side, pos, c.lv_qty(side, pos) - hr.qty + new_qty, head, c.lv_tail(side, pos));
        if (hr.id != oid) { ib_dec(hr.id); ib_inc(oid, side, price); }                  // the head slot changes identity
        hr.id = oid; hr.qty = new_qty; hr.agent = (uint32_t)agent; hr.price = new_price; nstore(head, hr);   // :352 book[i][0] = new_order  (slot 0, App. A-13)
        if (new_price != price) { s.book_flags |= BKF_REPRICED; c.lv_setp(side, pos, new_price); }          // the level now shows the new price where it stands: no re-sorting (:381,393)
#pragma unroll 1
        for (int k = 0; k < matches * buckets; k++) {
          exch_send_order(agent, ABX_ORDER_MODIFIED, oid, new_price, new_qty, 0, is_buy, 0.0);
          if (n_out >= Ctx::OUTN - 3) flush();
        }
      }
      if (cnt <= 1) return;
      limit = pos;
    }
  }
  ABX_HD void env_exch_receive(const Event &m) {                                        // ExchangeAgent.receiveMessage :129-340
    s.exch_comp_delay = P.c.exchange_computation_delay_ns;
    bool t_closed = s.now > P.c.mkt_close_ns;
    int32_t p[6] = {0, 0, 0, 0, 0, 0};
    if (t_closed && m.kind != ABX_QUERY_SPREAD) { exch_send(m.sender, ABX_MKT_CLOSED, p, 0.0); return; }       // :142-160
    if (m.kind == ABX_WHEN_MKT_OPEN) { s.exch_comp_delay = 0; exch_send(m.sender, ABX_WHEN_MKT_OPEN, p, 0.0); }
    else if (m.kind == ABX_WHEN_MKT_CLOSE) { s.exch_comp_delay = 0; exch_send(m.sender, ABX_WHEN_MKT_CLOSE, p, 0.0); }
    else if (m.kind == ABX_QUERY_SPREAD) {                                              // :215-245, depth 500: the agent reads 2 levels
      s.c_query++; int nb = s.n_bid_lv, na = s.n_ask_lv; int f = 0; int32_t b2 = 0, a2 = 0;
      if (nb > 0) { p[0] = c.lv_price(0, nb - 1); p[1] = c.lv_qty(0, nb - 1); f |= 1; } if (nb > 1) b2 = c.lv_price(0, nb - 2);
      if (na > 0) { p[2] = c.lv_price(1, na - 1); p[3] = c.lv_qty(1, na - 1); f |= 2; } if (na > 1) a2 = c.lv_price(1, na - 2);
      if (t_closed) f |= 4;
      f |= (nb > 2 ? 2 : nb) << 3; f |= (na > 2 ? 2 : na) << 5;
      p[4] = s.last_trade; p[5] = f;
      if (DQ && m.sender >= 2 + P.dq_n_mom) b2 = snap_take(m.sender - (2 + P.dq_n_mom));  // execution agents (depth 500): the levels are copied NOW; the reply carries their counts
      exch_send(m.sender, ABX_QUERY_SPREAD, p, bits_dbl((uint64_t)(uint32_t)b2 | ((uint64_t)(uint32_t)a2 << 32)));
    } else if (m.kind == ABX_LIMIT_ORDER) { s.c_limit++; s.ctr_kernel++; book_handle_limit((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], m.p[2], 0.0); c.sync(); trace_snap(); }
    else if (m.kind == ABX_CANCEL_ORDER) { s.c_cancel++; s.ctr_kernel++; book_cancel((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], 0.0); c.sync(); trace_snap(); }
    else if (m.kind == ABX_MODIFY_ORDER) { s.ctr_kernel++; if (!(m.p[4] & 2)) book_modify((uint32_t)m.p[0], m.sender, m.p[4] & 1, m.p[1], m.p[3], m.p[5]); c.sync(); trace_snap(); }   // :326-340 (bit 1: ids differ, :343)
  }
  // ---- MarketReplayAgent ----
  ABX_HD void replay_place(EnvX *x, int r) {                                            // placeOrder :69-96 for one row
    int4 row = c.row_load(r);
    { int r0 = r - c.day_row0(); c.sync(); if (c.onchip_writer()) x->rows_done = r0 + (row.x >= 0 ? 1 : 0); c.sync(); }   // rows before this one are used ids; an explicit id counts from its own row on
    if (row.x < 0) {                                                                    // ORDER_ID 0 == "unset" (util/order/Order.py:27): orders.get(0) finds
      int32_t gq = x->g0_qty, gp = x->g0_pq;                                            // only the order that received GENERATED id 0
      if (gq == 0 && row.z > 0) {
        uint32_t oid = gen_id();
        if (oid == 0) { c.sync(); if (c.onchip_writer()) { x->g0_qty = row.z; x->g0_pq = (row.y << 1) | (row.w & 1); } c.sync(); }
        int32_t p[6] = {(int32_t)oid, row.y, row.z, 0, row.w, 0}; env_send(ABX_LIMIT_ORDER, p, false);
      } else if (gq != 0 && row.z == 0) {
        int32_t p[6] = {0, gp >> 1, gq, 0, gp & 1, 0}; env_send(ABX_CANCEL_ORDER, p, false);
      } else if (gq != 0) {                                                             // new_order gets a fresh id != 0: isSameOrder fails at the exchange
        gen_id();
        int32_t p[6] = {0, gp >> 1, gq, row.y, (gp & 1) | 2, row.z}; env_send(ABX_MODIFY_ORDER, p, false);
      }
      if (n_out >= Ctx::OUTN - 3) flush();
      return;
    }
    uint32_t oid = REPLAY_ID_BASE + (uint32_t)row.x;
    uint4 t = c.ord_load(row.x); bool existing = t.x != 0;
    c.ib_prefetch(row.x);                                                               // the exchange reads the order's census when the message arrives
    if (!existing && row.z > 0) {                                                       // placeLimitOrder(order_id=ORDER_ID)
      t.x = (uint32_t)row.z; t.y = ((uint32_t)row.y << 1) | (uint32_t)(row.w & 1); c.ord_store(row.x, t);
      int32_t p[6] = {(int32_t)oid, row.y, row.z, 0, row.w, 0}; env_send(ABX_LIMIT_ORDER, p, false);
    } else if (existing && row.z == 0) {                                                // cancelOrder(existing_order)
      int32_t p[6] = {(int32_t)oid, (int32_t)(t.y >> 1), (int32_t)t.x, 0, (int32_t)(t.y & 1u), 0}; env_send(ABX_CANCEL_ORDER, p, false);
    } else if (existing) {                                                              // modifyOrder(existing_order, LimitOrder(new SIZE, PRICE))
      int32_t p[6] = {(int32_t)oid, (int32_t)(t.y >> 1), (int32_t)t.x, row.y, (int32_t)(t.y & 1u), row.z}; env_send(ABX_MODIFY_ORDER, p, false);
    }
    if (n_out >= Ctx::OUTN - 3) flush();
  }
  ABX_HD void replay_wakeup(EnvX *x) {                                                  // wakeup :50-60
    ta_wakeup(x->ra_flags);
    if (!(x->ra_flags & AF_HAS_OPEN) || !(x->ra_flags & AF_HAS_CLOSE)) return;
    int cur = x->wt_cursor;
    if (cur >= c.n_ts()) return;                                                        // wakeup_times[0] -> IndexError: nothing placed
    set_wakeup(1, c.ts_load(cur)); if (c.onchip_writer()) x->wt_cursor = cur + 1;       // setWakeup(wakeup_times[0]); pop(0)
    int k = cur == 0 ? 0 : cur - 1;                                                     // orders_dict[currentTime]: the list's first entry is woken twice
    if (c.ts_load(k) != s.now) { s.flags |= ABX_F_REF_EXCEPTION; return; }              // orders_dict[currentTime] of a wakeup that is not on a stream timestamp: KeyError in the reference (:57)
    int r0 = c.first_load(k), r1 = c.first_load(k + 1);
    int4 nxt = c.row_load(k + 1 < c.n_ts() ? r1 : r0);                                  // first row of the next wakeup: its per-order records are cold
#pragma unroll 1
    for (int r = r0; r < r1; r++) replay_place(x, r);
    if (nxt.x >= 0) { c.id_prefetch(nxt.x); c.ib_prefetch(nxt.x); }
  }
  ABX_HD void replay_receive(EnvX *x, const Event &m) {
    uint32_t fl = x->ra_flags; int32_t sh = x->ra_shares, lt = x->ra_last_trade; int64_t cash = x->ra_cash;
    bool newly = ta_receive(m, fl, sh, cash, lt);
    if (m.kind == ABX_ORDER_EXECUTED || m.kind == ABX_ORDER_CANCELLED) {                // orderExecuted :422-462 / orderCancelled :476-489
      uint32_t oid = (uint32_t)m.p[0];
      if (oid >= REPLAY_ID_BASE) { uint4 t = c.ord_load((int)(oid - REPLAY_ID_BASE));
        if (t.x != 0) { if (m.kind == ABX_ORDER_CANCELLED || (uint32_t)m.p[2] >= t.x) t.x = 0; else t.x -= (uint32_t)m.p[2]; c.ord_store((int)(oid - REPLAY_ID_BASE), t); } }
      else if (oid == 0 && x->g0_qty != 0) { int32_t gq = x->g0_qty; gq = (m.kind == ABX_ORDER_CANCELLED || m.p[2] >= gq) ? 0 : gq - m.p[2]; c.sync(); if (c.onchip_writer()) x->g0_qty = gq; c.sync(); }
      if (m.kind == ABX_ORDER_EXECUTED) { lt = m.p[3]; fl |= AF_HAS_LAST; }             // MarketReplayAgent.receiveMessage :62-67
    }
    if (newly) set_wakeup(1, c.ts_load(0));                                             // mkt_open + getWakeFrequency() == first_wakeup
    if (c.onchip_writer()) { x->ra_flags = fl; x->ra_shares = sh; x->ra_last_trade = lt; x->ra_cash = cash; }
  }
  // ---- DummyRLExecutionAgent ----
  ABX_HD void rl_wakeup(EnvX *x) {                                                      // wakeup :184-218
    uint32_t fl = x->rl_flags;
    if (!ta_wakeup(fl)) return;
    if (fl & RLF_TRADE) {                                                               // first horizon time > now -> CANCEL_ORDER event (Timedelta(0.5) == 0)
      int64_t k = s.now < P.h0_ns ? 0 : hdiv(s.now - P.h0_ns) + 1;
      if (k < P.n_h) env_set_cancel(2, P.h0_ns + k * P.h_step_ns); else fl &= ~RLF_TRADE;
    }
    if (fl & RLF_TRADE) {                                                               // effective horizon = horizon[:-1]
      int64_t k = s.now < P.h0_ns ? 0 : hdiv(s.now - P.h0_ns) + 1;
      if (k < P.n_h - 1) set_wakeup(2, P.h0_ns + k * P.h_step_ns); else fl &= ~RLF_TRADE;
    }
    { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); }         // getCurrentSpread(depth=500)
    fl = (fl & ~AF_STATE_MASK) | (ST_AWAITING_SPREAD << AF_STATE_SHIFT);
    if (c.onchip_writer()) x->rl_flags = fl;
  }
  ABX_HD void rl_orders_remove(EnvX *x, int i) {
    c.sync();
    if (c.onchip_writer()) { for (int k = i; k + 1 < RL_ORDER_CAP; k++) { x->rl_oid[k] = x->rl_oid[k + 1]; x->rl_oprice[k] = x->rl_oprice[k + 1]; x->rl_oqty[k] = x->rl_oqty[k + 1]; } x->rl_n_orders = x->rl_n_orders - 1; }
    c.sync();
  }
  ABX_HD int rl_orders_find(EnvX *x, uint32_t oid) { int f = -1; int n = x->rl_n_orders; for (int i = 0; i < RL_ORDER_CAP; i++) if (i < n && f < 0 && x->rl_oid[i] == oid) f = i; return f; }
  ABX_HD void rl_receive(EnvX *x, const Event &m) {                                     // receiveMessage :230-245
    uint32_t fl = x->rl_flags; int32_t sh = x->rl_shares, lt = x->rl_last_trade; int64_t cash = x->rl_cash;
    bool newly = ta_receive(m, fl, sh, cash, lt);
    double rem = x->rem_quantity;
    if (m.kind == ABX_ORDER_EXECUTED) {                                                 // orderExecuted + ExecutionAgent.handleOrderExecution :88-99
      int i = rl_orders_find(x, (uint32_t)m.p[0]);
      if (i >= 0) { int32_t oq = x->rl_oqty[i]; if (m.p[2] >= oq) rl_orders_remove(x, i); else { c.sync(); if (c.onchip_writer()) x->rl_oqty[i] = oq - m.p[2]; c.sync(); } }
      double ex = x->executed_sum + (double)m.p[2]; rem = P.rl_quantity - ex;
      if (c.onchip_writer()) { x->executed_sum = ex; x->n_executed = x->n_executed + 1; x->rem_quantity = rem; }
    } else if (m.kind == ABX_ORDER_CANCELLED) { int i = rl_orders_find(x, (uint32_t)m.p[0]); if (i >= 0) rl_orders_remove(x, i); }
    if (newly) set_wakeup(2, P.h0_ns);                                                  // mkt_open + (horizon[0] - mkt_open)
    uint32_t st = (fl & AF_STATE_MASK) >> AF_STATE_SHIFT;
    if (rem > 0 && st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) {            // metrics.addLOB :62-84
      fl = (fl & ~AF_STATE_MASK) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT);
      int head = (x->lob_head + LOB_CAP - 1) % LOB_CAP; int n = x->n_lobs;
      int32_t w[12] = {m.p[0], m.p[1], m.x0, m.p[2], m.p[3], m.x1, m.p[4], (m.p[5] >> 3) & 3, (m.p[5] >> 5) & 3, 0, 0, 0};
      c.lob_store(head, w);
      if (c.onchip_writer()) { if (!(fl & RLF_METRICS_INIT)) x->p0 = m.p[4]; x->lob_head = head; x->n_lobs = n < LOB_CAP ? n + 1 : n; }
      fl |= RLF_METRICS_INIT;
    }
    if (c.onchip_writer()) { x->rl_flags = fl; x->rl_shares = sh; x->rl_last_trade = lt; x->rl_cash = cash; }
    c.sync();
  }
  ABX_HD void rl_observe(EnvX *x) {                                                     // get_observation :294-315
    int64_t curr = hdiv(s.now) * P.h_step_ns; int rem = P.n_h;                      // get_remaining_time :282-292 (Timestamp.floor(freq))
    if (curr >= P.h0_ns) { int64_t kh = hdiv(curr - P.h0_ns); if (kh < P.n_h) rem = P.n_h - 1 - (int)kh; }
    int n = x->n_lobs, head = x->lob_head; int obs_len = 0; double o[9];
    for (int i = 0; i < 9; i++) o[i] = 0.0;
    if (n > 0) {
      int32_t l[12]; c.lob_load(head, l);
      if (l[7] > 0 && l[8] > 0) {
        double p0 = (double)x->p0, pt = (double)l[6];
        double mid = ((double)l[0] + (double)l[3]) / 2;
        bool bad = false; double vol = c.lob_midvol(n, head, p0, bad);                  // getMidPriceVolatility: np.std over the stored LOBs
        int d;                                                                          // getTradeDirection :197-216
        if (pt > mid) d = 1; else if (pt < mid) d = -1;
        else { int32_t lo[12]; c.lob_load((head + n - 1) % LOB_CAP, lo); if (lo[7] == 0 || lo[8] == 0) bad = true; double lm = ((double)lo[0] + (double)lo[3]) / 2; d = mid > lm ? 1 : -1; }
        if (!bad) {
          o[0] = rem; o[1] = x->rem_quantity; o[2] = log_ni(pt / p0); o[3] = (double)(l[3] - l[0]);
          o[4] = ((double)l[1] - (double)l[4]) / ((double)l[1] + (double)l[4]);
          o[5] = tanh((double)l[3] / (double)l[4] - (double)l[0] / (double)l[1]);
          o[6] = vol; o[7] = d; o[8] = 2 * d * (pt - mid) / mid; obs_len = 9;
        }
      }
    }
    if (obs_len == 0) s.flags |= ABX_F_OBS_INVALID;                                     // the reference raises ValueError here
    if (c.onchip_writer()) { x->rem_time = rem; x->obs_len = obs_len; for (int i = 0; i < 9; i++) x->obs[i] = o[i]; }
    c.sync();
  }
  ABX_HD void rl_place_orders(EnvX *x, double a0, double a1, double a2) {               // place_orders :159-181 + process_action :138-157
    int n = P.order_level; double q0 = P.rl_quantity, q = P.rl_quantity;                // metrics.rem_quantity is never updated (App. A-10)
    double act[2] = {a1, a2}; double sum = 0; for (int i = 0; i < 2; i++) if (i < n) sum = dadd(sum, act[i]);
    double q_hat = q / q0;
    double o_total = rint(dmul(dmul(q0, q_hat), pow_ni(a0, pow_ni(q_hat, P.rl_steep))));
    double o[2]; double part = 0;
    for (int i = 0; i < 2; i++) { double oh = sum == 0.0 ? 1.0 / n : act[i] / sum; o[i] = i < n ? rint(dmul(o_total, oh)) : 0.0; }
    for (int i = 0; i < 2; i++) if (i < n - 1) part = dadd(part, o[i]);
    if (n == 1) o[0] = dsub(o_total, part); else o[1] = dsub(o_total, part);
    int nl = x->n_lobs; if (nl == 0) return;                                            // exceptions are swallowed (bare except)
    int32_t l[12]; c.lob_load(x->lob_head, l);
    if (l[7] == 0 || l[8] == 0) return;
    self_id = 2;
    for (int lv = 0; lv < 2; lv++) if (lv < n) {
      if (lv >= l[7] || lv >= l[8]) continue;                                           // IndexError swallowed
      int32_t price = P.rl_is_buy ? (lv == 0 ? l[0] : l[2]) : (lv == 0 ? l[3] : l[5]);
      uint32_t oid = gen_id();                                                          // LimitOrder() built before the quantity test
      if (!(o[lv] > 0)) continue;
      int k = x->rl_n_orders;
      if (k < RL_ORDER_CAP) { c.sync(); if (c.onchip_writer()) { x->rl_oid[k] = oid; x->rl_oprice[k] = price; x->rl_oqty[k] = (int32_t)o[lv]; x->rl_n_orders = k + 1; } c.sync(); }
      else s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW;
      int32_t p[6] = {(int32_t)oid, price, (int32_t)o[lv], 0, P.rl_is_buy, 0}; env_send(ABX_LIMIT_ORDER, p, false);
    }
  }
  ABX_HD void rl_cancel_all(EnvX *x) {                                                  // cancelAllOrders :247-255
    self_id = 2; int n = x->rl_n_orders;
#pragma unroll 1
    for (int i = 0; i < n; i++) { int32_t p[6] = {(int32_t)x->rl_oid[i], x->rl_oprice[i], x->rl_oqty[i], 0, P.rl_is_buy, 0}; env_send(ABX_CANCEL_ORDER, p, false); }
  }
  // reset: Kernel/GymKernel.initRunner :139-146 -- one WAKEUP per agent at start
  ABX_HD void env_reset() { for (int id = 0; id < P.c.n_agents; id++) set_wakeup(id, P.c.start_ns); flush(); }
  // GymKernel.stepRunner :158-306.  Returns done as ABIDESEnv.step computes it (ABIDESEnv.py:42-46).
  ABX_HD bool env_step(double a0, double a1, double a2) {
    EnvX *x = c.envx();
    if (P.order_level > 0) { rl_place_orders(x, a0, a1, a2); flush(); }                 // order_level 0: no RL agent (config/marketreplay.py)
    bool end_step = false, more = true;
#pragma unroll 1
    while (!end_step) {
      uint64_t khi; uint32_t kuniq; int grp;
      bool any = c.q_min(khi, kuniq, grp);
      if (!any || !(s.now <= P.c.stop_ns)) { more = false; break; }
      Event ev; c.q_fetch(grp, ev);
      s.now = ev.t; s.ttl++;
      if (INSTR && P.c.hash_pops) s.pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s.pop_hash, ev.t), ev.recipient), ev.type), ev.type == ABX_T_MESSAGE ? (int64_t)ev.uniq : -1);
      if (INSTR && P.c.trace_cap > 0) {
        abx_trace_rec r; r.tag = 0; r.a = ev.recipient; r.t = ev.t; for (int i = 0; i < 16; i++) r.v[i] = 0;
        r.v[0] = ev.type; r.v[1] = ev.type == ABX_T_MESSAGE ? (int32_t)ev.uniq : -1; r.v[2] = ev.kind; trace_rec(r);
      }
      addl_delay = 0;
      int id = ev.recipient;
      if (ev.type == ABX_T_CANCEL_ORDER) { c.q_remove(); s.q_count--; rl_cancel_all(x); flush(); continue; }   // :241-245 (get_reward is None)
      int64_t at = id == 0 ? s.exch_time : (id == 1 ? x->ra_time : x->rl_time);
      if (at > s.now) { c.q_requeue(at); continue; }
      c.q_remove(); s.q_count--;
      self_id = id; int64_t delay = P.c.default_computation_delay_ns;
      if (id == 0) { if (ev.type == ABX_T_MESSAGE) env_exch_receive(ev); delay = s.exch_comp_delay; s.exch_time = s.now + delay + addl_delay; }
      else if (id == 1) { if (ev.type == ABX_T_WAKEUP) replay_wakeup(x); else replay_receive(x, ev); if (c.onchip_writer()) x->ra_time = s.now + delay + addl_delay; }
      else {
        if (ev.type == ABX_T_WAKEUP) rl_wakeup(x); else rl_receive(x, ev);
        if (c.onchip_writer()) x->rl_time = s.now + delay + addl_delay;
        if (ev.type == ABX_T_MESSAGE && ev.kind == ABX_QUERY_SPREAD) { rl_observe(x); end_step = true; }       // :286-289
      }
      flush();
    }
    if (more) { uint64_t khi; uint32_t kuniq; int grp; more = c.q_min(khi, kuniq, grp) && (s.now <= P.c.stop_ns); }
    if (!more) s.flags |= ABX_F_DONE;
    if (c.onchip_writer()) x->steps = x->steps + 1;
    c.sync();
    return !more;
  }


  // =================================================================================================
  // Book surface: what ExchangeAgent calls on util/OrderBook.py -- handleLimitOrder :38-170, cancelOrder :284-339,
  // modifyOrder :341-372 -- driven by an operation tape recorded at the exchange boundary instead of by agents.  ops: rows of 9 int64
  // (t_ns, op 0 limit / 1 cancel / 2 modify, agent, device order id, is_buy, price, qty, new_price, new_qty).  Every notification the book
  // sends (ORDER_EXECUTED pairs, ORDER_ACCEPTED, ORDER_CANCELLED, ORDER_MODIFIED) and the book state after every operation go to the trace.
  // =================================================================================================
  ABX_HD void book_replay(const int64_t *ops, int64_t n_ops) {
    self_id = 0;
#pragma unroll 1
    for (int64_t i = 0; i < n_ops; i++) {
      const int64_t *r = ops + 9 * i;
      s.now = r[0]; s.ttl++; addl_delay = 0;
      int op = (int)r[1], agent = (int)r[2], is_buy = (int)r[4]; uint32_t oid = (uint32_t)r[3]; int32_t price = (int32_t)r[5], qty = (int32_t)r[6];
      if (op == 0) { s.c_limit++; book_handle_limit(oid, agent, is_buy, price, qty, 0.0); }
      else if (op == 1) { s.c_cancel++; book_cancel(oid, agent, is_buy, price, 0.0); }
      else if (op == 2 && oid != 0xffffffffu) book_modify(oid, agent, is_buy, price, (int32_t)r[7], (int32_t)r[8]);   // id 0xffffffff: isSameOrder fails (:343)
      c.sync(); trace_snap();
      flush();
    }
    s.flags |= ABX_F_DONE;
  }

  // =================================================================================================
  // DDQN execution config (config/execution/marketreplay/execution_marketreplay_ddqn.py, -a rl): Exchange (0) + MarketReplayAgent (1,
  // EnvX) + MomentumAgents (2.., staged records, r3_* handlers) + TWAPExecutionAgent(s) + DDQLearningExecutionAgent (last id) under
  // Kernel.runner; zero latency and computation delay.  agent/execution/baselines/execution_agent.py:66-130, twap_agent.py:51-63,
  // agent/execution/qlearning/ddqlearning_execution_agent.py:20-37,141-185,228-447,507-611, agent/execution/util.py:6-42,
  // agent/TradingAgent.py:351-397 (placeMarketOrder).  The Q-network is outside: one launch runs every environment up to the DDQN
  // agent's next choose_action (:245) and the following launch resumes place_order with the action.
  // Depth-500 QUERY_SPREAD replies: the exchange copies the first 500 levels of both sides into the asking agent's snapshot area in HBM when it
  // PROCESSES the query (ExchangeAgent.py:231-245) and the reply carries the level counts; the agent's handlers (market-order walk,
  // level prices of take_action) read that copy -- known_bids / known_asks -- never the live ladders.  L1 is cached in the record.
  // =================================================================================================
  ABX_HD ExecAux *exaux() { return reinterpret_cast<ExecAux *>(z->oid); }
  ABX_HD int exec_index(int id) const { return R3 ? 0 : id - (2 + P.dq_n_mom); }        // which snapshot area / order table an execution agent owns
  // ExchangeAgent.py:231-245 for a deep query: getInsideBids(depth) / getInsideAsks(depth) copied at processing time.  Returns bids | asks << 16.
  ABX_HD int32_t snap_take(int k, int depth = 0x7fffffff) {
    if (depth > P.snap_depth) depth = P.snap_depth;
    int nb = s.n_bid_lv < depth ? s.n_bid_lv : depth, na = s.n_ask_lv < depth ? s.n_ask_lv : depth;
    c.snap_store(k, 0, s.n_bid_lv, nb); c.snap_store(k, 1, s.n_ask_lv, na);
    return (int32_t)((uint32_t)nb | ((uint32_t)na << 16));
  }
  ABX_HD int dq_order_base(int id) const { return R3 ? P.dq_order_base : P.dq_order_base + (id - (2 + P.dq_n_mom)) * EXEC_ORDER_CAP; }   // rmsc03 population: one POV execution agent
  // floor(a / h_step_ns) for 0 <= a < 2^53 (nanoseconds of one day): a software 64-bit division is ~55 instructions per use; the product with the
  // host-computed reciprocal is off by at most one, which the remainder corrects exactly.
  ABX_HD int64_t hdiv(int64_t a) const {
    int64_t q = (int64_t)((double)a * P.h_step_inv), r = a - q * P.h_step_ns;
    if (r < 0) q--; else if (r >= P.h_step_ns) q++;
    return q;
  }
  ABX_HD int dq_horizon_index(int64_t t) const { if (t < P.h0_ns) return -1; int64_t k = hdiv(t - P.h0_ns); return (k * P.h_step_ns == t - P.h0_ns && k < P.n_h) ? (int)k : -1; }
  ABX_HD void dq_place_limit(int id, int32_t size, bool buy, int32_t price) {           // TradingAgent.placeLimitOrder :309-349
    uint32_t oid = gen_id();
    if (size <= 0) return;
    if (a.n_orders < EXEC_ORDER_CAP) { uint4 v; v.x = oid; v.y = (uint32_t)price; v.z = (uint32_t)(buy ? size : -size); v.w = 0; c.id_store(dq_order_base(id) + a.n_orders, v); a.n_orders++; }
    else s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW;
    int32_t pl[6] = {(int32_t)oid, price, size, 0, buy, 0}; env_send(ABX_LIMIT_ORDER, pl, false);
    if (n_out >= Ctx::OUTN - 3) flush();
  }
  ABX_HD void dq_cancel_all(int id) {                                                   // ExecutionAgent.cancelOrders :126-128 / DDQN cancel_orders :578-585
    int base = dq_order_base(id);
#pragma unroll 1
    for (int i = 0; i < a.n_orders; i++) {
      uint4 v = c.id_load(base + i); int32_t q = (int32_t)v.z;
      int32_t p[6] = {(int32_t)v.x, (int32_t)v.y, q < 0 ? -q : q, 0, q > 0, 0}; env_send(ABX_CANCEL_ORDER, p, false);
      if (n_out >= Ctx::OUTN - 3) flush();
    }
  }
  ABX_HD void dq_order_update(int id, uint32_t oid, int32_t fill, bool cancel) {         // TradingAgent.orderExecuted :445-452, orderCancelled :480-483
    int base = dq_order_base(id), f = c.tab_find(base, a.n_orders, oid);
    if (f < 0) return;
    uint4 v = c.id_load(base + f); int32_t q = (int32_t)v.z, aq = q < 0 ? -q : q;
    if (!cancel && fill < aq) { v.z = (uint32_t)(q < 0 ? -(aq - fill) : (aq - fill)); c.id_store(base + f, v); return; }
    c.tab_remove(base, a.n_orders, f);                                                  // dict deletion keeps the order of the others
    a.n_orders--;
  }
  ABX_HD void dq_place_market(int id, int32_t quantity) { dq_place_market(id, quantity, P.rl_is_buy != 0); }
  ABX_HD void dq_place_market(int id, int32_t quantity, bool is_buy) {                  // TradingAgent.placeMarketOrder :351-397 over the cached opposite side
    if (quantity <= 0) return;
    int opp = is_buy ? 1 : 0, k = exec_index(id); int32_t sn = exaux()->snap_n; int depth = opp ? (sn >> 16) & 0xffff : sn & 0xffff;
    if (depth == 0) { s.flags |= ABX_F_OBS_INVALID; return; }                           // the reference iterates None
#pragma unroll 1
    for (int i = 0; i < depth; i++) {
      int2 lv = c.snap_load(k, opp, i); int32_t price = lv.x, sz = lv.y;
      bool last = quantity <= sz;
      dq_place_limit(id, last ? quantity : sz, is_buy, price);
      if (last) break;
      quantity -= sz;
    }
  }
  // np.digitize(x, np.linspace(0, 1, 201)[1:-1]): number of split points k * (1 / 200), k = 1 .. 199, that are <= x
  ABX_HD int dq_digitize(double x) const {
    int k = x > 0.0 ? (x < 1.0 ? (int)(x * 200.0) : 199) : 0; if (k > 199) k = 199;
    while (k < 199 && dmul((double)(k + 1), 1.0 / 200) <= x) k++;
    while (k > 0 && dmul((double)k, 1.0 / 200) > x) k--;
    return k;
  }
  // DDQLearningExecutionAgent.get_observation :280-336 from the cached L1 (known_bids[0] / known_asks[0])
  ABX_HD void dq_get_observation(ExecAux &ex, double obs[6], int disc[2]) {
    int64_t curr = hdiv(s.now) * P.h_step_ns; int hi = dq_horizon_index(curr);
    ex.rem_time = hi >= 0 ? P.n_h - 1 - hi : P.n_h;
    for (int i = 0; i < 6; i++) obs[i] = 0.0; disc[0] = disc[1] = 0;
    if (!(a.flags & AF_HAS_BID) || !(a.flags & AF_HAS_ASK)) { s.flags |= ABX_F_OBS_INVALID; return; }
    obs[0] = dsub(dmul(2.0, (double)ex.rem_time / (double)P.n_h), 1.0);
    obs[1] = dsub(dmul(2.0, (double)ex.rem_qty / (double)P.dq_quantity), 1.0);
    obs[2] = (double)(a.ask - a.bid);
    obs[3] = (double)(a.ask_q - a.bid_q) / (double)(a.ask_q + a.bid_q);
    int32_t mid2 = a.bid + a.ask, prev2 = ex.pp_last2; double mid = (double)mid2 / 2;
    if (ex.n_pp == 0) ex.pp0_2 = mid2;
    ex.pp_last2 = mid2; ex.n_pp++;                                                      // price_path.append(mid_p)
    obs[4] = s.now == P.h0_ns ? 0.0 : log_ni(mid / ((double)prev2 / 2));
    obs[5] = log_ni(mid / ((double)ex.pp0_2 / 2));
    disc[0] = dq_digitize(obs[0]); disc[1] = dq_digitize(obs[1]);                       // the grid has two dimensions: zip() drops the other four features
  }
  ABX_HD void exaux_store(const ExecAux &ex) { c.sync(); if (c.onchip_writer()) *exaux() = ex; c.sync(); }
  ABX_HD void dq_exec_wakeup(int id, int type) {
    if (!ta_wakeup(a.flags)) return;
    ExecAux ex = *exaux();
    int64_t k = s.now < P.h0_ns ? 0 : hdiv(s.now - P.h0_ns) + 1;              // first horizon time > now
    bool query = false;
    if (type == AT_DDQN) {                                                              // ddqlearning_execution_agent.py:141-153
      if (ex.exflags & EXF_TRADE) { if (k < P.n_h) set_wakeup(id, P.h0_ns + k * P.h_step_ns); else ex.exflags &= ~EXF_TRADE; }
      query = true;
    } else if (ex.exflags & EXF_TRADE) {                                                // execution_agent.py:66-78
      if (k < P.n_h) set_wakeup(id, P.h0_ns + k * P.h_step_ns);
      query = true;
    }
    if (query) { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_SPREAD << AF_STATE_SHIFT); }
    exaux_store(ex);
  }
  ABX_HD void dq_exec_receive(int id, int type, const Event &m, EnvX *x) {
    bool had = (a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE);
    ExecAux ex = *exaux();
    if (m.kind == ABX_WHEN_MKT_OPEN) a.flags |= AF_HAS_OPEN;
    else if (m.kind == ABX_WHEN_MKT_CLOSE) a.flags |= AF_HAS_CLOSE;
    else if (m.kind == ABX_ORDER_EXECUTED) {                                            // TradingAgent.orderExecuted :422-462
      int32_t q = m.p[2]; int32_t sq = m.p[4] ? q : -q; a.shares += sq; a.cash -= (int64_t)sq * m.p[3];
      dq_order_update(id, (uint32_t)m.p[0], q, false);
      ex.executed_sum += q; ex.n_executed++; ex.rem_qty = (int32_t)P.dq_quantity - ex.executed_sum;     // handleOrderExecution :88-92 / handle_order_execution :517-521
    } else if (m.kind == ABX_ORDER_CANCELLED) dq_order_update(id, (uint32_t)m.p[0], 0, true);
    else if (m.kind == ABX_MKT_CLOSED) a.flags |= AF_MKT_CLOSED;
    else if (m.kind == ABX_QUERY_SPREAD) {                                              // querySpread :514-537
      if (m.p[5] & 4) a.flags |= AF_MKT_CLOSED;
      a.last_trade = m.p[4]; a.flags |= AF_HAS_LAST;
      a.flags &= ~(AF_HAS_BID | AF_HAS_ASK);
      if (m.p[5] & 1) { a.flags |= AF_HAS_BID; a.bid = m.p[0]; a.bid_q = m.p[1]; } else { a.bid = 0; a.bid_q = 0; }
      if (m.p[5] & 2) { a.flags |= AF_HAS_ASK; a.ask = m.p[2]; a.ask_q = m.p[3]; } else { a.ask = 0; a.ask_q = 0; }
      ex.snap_n = m.x0;                                                                 // known_bids / known_asks = the lists the exchange copied when it processed the query
    }
    if ((a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE) && !had) set_wakeup(id, P.h0_ns);         // mkt_open + (start_time - mkt_open)
    uint32_t st = (a.flags & AF_STATE_MASK) >> AF_STATE_SHIFT;
    bool both = (a.flags & AF_HAS_BID) && (a.flags & AF_HAS_ASK);
    if (type == AT_TWAP) {                                                              // ExecutionAgent.receiveMessage :80-86
      if (ex.rem_qty > 0 && st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) {
        dq_cancel_all(id);
        int hi = dq_horizon_index(s.now);                                               // placeOrders :107-124
        if (hi == P.n_h - 2) dq_place_market(id, ex.rem_qty);
        else if (hi >= 0 && hi < P.n_h - 2) {
          if (!both) s.flags |= ABX_F_OBS_INVALID;
          else { if (hi == 0) ex.arr2 = a.bid + a.ask; int32_t sq = P.sched ? P.sched[exec_index(id) * P.n_h + hi] : -1;      // self.schedule[pd.Interval(now, now + 30 s)] (execution_agent.py:121): TWAP's one quantity or the VWAP agent's per-bin one
                 dq_place_limit(id, sq >= 0 ? sq : ex.child_qty, P.rl_is_buy != 0, P.rl_is_buy ? a.ask : a.bid); }
        }
      }
    } else if (m.kind == ABX_ORDER_ACCEPTED || m.kind == ABX_ORDER_EXECUTED) {         // handle_order_acceptance :550-576 / handle_order_execution :507-548
      int64_t curr = hdiv(s.now) * P.h_step_ns;
      if (dq_horizon_index(curr) >= 0) {
        double o6[6]; int sp[2]; dq_get_observation(ex, o6, sp);
        ex.e_sp[0] = (int16_t)sp[0]; ex.e_sp[1] = (int16_t)sp[1]; ex.cur_s[0] = (int16_t)sp[0]; ex.cur_s[1] = (int16_t)sp[1];
        if (m.kind == ABX_ORDER_EXECUTED) {                                             // compute_reward :411-447
          double fp = (double)m.p[3], ar = (double)ex.arr2 / 2;
          double slip = P.rl_is_buy ? dsub(fp, ar) : dsub(ar, fp);
          double r = dmul(dmul(dsub(1.0, slip / ar), (double)m.p[2]) / (double)P.dq_quantity, 10000.0);
          ex.e_r = r; ex.step_reward = dadd(ex.step_reward, r);
        } else ex.e_r = 0.0;
        ex.exflags |= EXF_E_R;
      }
    } else {
      int hi = dq_horizon_index(s.now);
      if (hi >= 0 && hi < P.n_h - 1 && ex.rem_qty > 0 && st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) {   // :164-172
        dq_cancel_all(id);
        if (!both) s.flags |= ABX_F_OBS_INVALID;
        else {                                                                          // place_order :228-245 up to choose_action
          double o6[6]; int sp[2];
          if (hi == 0) { ex.arr2 = a.bid + a.ask; dq_get_observation(ex, o6, sp); ex.cur_s[0] = (int16_t)sp[0]; ex.cur_s[1] = (int16_t)sp[1]; }
          dq_get_observation(ex, o6, sp); ex.sp[0] = (int16_t)sp[0]; ex.sp[1] = (int16_t)sp[1];
          c.sync();
          if (c.onchip_writer()) { for (int i = 0; i < 6; i++) x->obs[i] = o6[i]; x->obs[6] = sp[0]; x->obs[7] = sp[1]; x->obs_len = 8; x->rl_flags = x->rl_flags | DQF_PENDING; }
          c.sync();
        }
      }
    }
    exaux_store(ex);
  }
  static ABX_HD double dq_alloc_frac(int alloc, int lv) {                               // SIZE_ALLOCATION :20
    if (alloc == 1) return lv == 0 ? 1.0 : 0.0;
    if (alloc == 2) return lv < 2 ? 0.5 : 0.0;
    return lv == 0 ? 0.34 : (lv < 3 ? 0.33 : 0.0);
  }
  // place_order :245-278 from choose_action's return on, then receiveMessage :172 (self.t += 1); the DDQN agent's record is staged
  ABX_HD void dq_resume(int id, int action) {
    ExecAux ex = *exaux();
    if (action < 0) action = 0; if (action > 23) action = 23;
    int alloc = action / 6, j = action % 6; int32_t qty = ex.child_qty;                 // ACTIONS[a] = (allocation a / 6, SIZE_SCALE[a % 6]) :24-37
    if (ex.rem_time == 1) { qty = ex.rem_qty; alloc = 0; }                              // take_action :380-384
    else { double sc = j == 0 ? 0.1 : dmul(0.5, (double)j); qty = (int32_t)py_round_i64(dmul(sc, (double)qty)); if (qty < 0) qty = 0; }
    ex.e_s[0] = ex.cur_s[0]; ex.e_s[1] = ex.cur_s[1]; ex.e_a = action; ex.e_sp[0] = ex.sp[0]; ex.e_sp[1] = ex.sp[1]; ex.e_r = 0.0;   // experience[t] = (self.s, a, s', None)
    ex.exflags = (ex.exflags | EXF_E_VALID) & ~EXF_E_R; ex.cur_s[0] = ex.sp[0]; ex.cur_s[1] = ex.sp[1]; ex.t++; ex.step_reward = 0.0;
    exaux_store(ex);
    if (alloc == 0) dq_place_market(id, qty);
    else {
      int own = P.rl_is_buy ? 0 : 1, n = own ? (ex.snap_n >> 16) & 0xffff : ex.snap_n & 0xffff, k = exec_index(id);
#pragma unroll 1
      for (int lv = 0; lv < 4; lv++) {                                                  // :395-409 (the price is read before the size test: IndexError below 4 levels)
        int32_t size = (int32_t)py_round_i64(dmul(dq_alloc_frac(alloc, lv), (double)qty));
        if (lv >= n) { s.flags |= ABX_F_REF_EXCEPTION; break; }                         // bids[3] / asks[3] of a thinner book: the reference raises IndexError
        int32_t price = c.snap_load(k, own, lv).x;
        if (size != 0) dq_place_limit(id, size, P.rl_is_buy != 0, price);
      }
    }
  }
  // One decision step: finish the pending place_order with `action`, then Kernel.runner's loop (Kernel.py:190-292) until the DDQN agent
  // reaches choose_action again (returns true) or the loop ends (ABX_F_DONE, returns false).
  ABX_HD bool dq_step(int action) {
    EnvX *x = c.envx(); int ddqn_id = P.dq_has_ddqn ? P.c.n_agents - 1 : -1;
    if (x->rl_flags & DQF_PENDING) {
      z = c.agent_stage(ddqn_id); regs_load(a, z); self_id = ddqn_id; addl_delay = 0;
      dq_resume(ddqn_id, action);
      c.sync(); if (c.onchip_writer()) { regs_store(z, a); x->rl_flags = x->rl_flags & ~DQF_PENDING; x->obs_len = 0; } c.sync();
      c.agent_commit(ddqn_id); flush();
    }
    bool paused = false;
#pragma unroll 1
    while (!paused) {
      uint64_t khi; uint32_t kuniq; int grp;
      bool any = c.q_min(khi, kuniq, grp);
      if (!any || !(s.now <= P.c.stop_ns) || (rng.tape() && (rng.err & ABX_F_TAPE_UNDERRUN))) { s.flags |= ABX_F_DONE; break; }   // an exhausted tape hands out zeros: with zero delays an agent would re-wake itself at the same instant for ever
      Event ev; c.q_fetch(grp, ev);
      s.now = ev.t; s.ttl++;
      if (INSTR && P.c.hash_pops) s.pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s.pop_hash, ev.t), ev.recipient), ev.type), ev.type == ABX_T_MESSAGE ? (int64_t)ev.uniq : -1);
      if (INSTR && P.c.trace_cap > 0) {
        abx_trace_rec r; r.tag = 0; r.a = ev.recipient; r.t = ev.t; for (int i = 0; i < 16; i++) r.v[i] = 0;
        r.v[0] = ev.type; r.v[1] = ev.type == ABX_T_MESSAGE ? (int32_t)ev.uniq : -1; r.v[2] = ev.kind; trace_rec(r);
      }
      addl_delay = 0;
      int id = ev.recipient;
      if (id <= 1) {
        int64_t at = id == 0 ? s.exch_time : x->ra_time;
        if (at > s.now) { c.q_requeue(at); continue; }
        c.q_remove(); s.q_count--; self_id = id;
        if (id == 0) { if (ev.type == ABX_T_MESSAGE) env_exch_receive(ev); s.exch_time = s.now + s.exch_comp_delay + addl_delay; }
        else { if (ev.type == ABX_T_WAKEUP) replay_wakeup(x); else replay_receive(x, ev); if (c.onchip_writer()) x->ra_time = s.now + P.c.default_computation_delay_ns + addl_delay; }
      } else {
        z = c.agent_stage(id); regs_load(a, z);
        if (a.agent_time > s.now) { c.q_requeue(a.agent_time); continue; }
        c.q_remove(); s.q_count--; self_id = id;
        int type = (int)((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT);
        if (type == AT_MOMENTUM) { if (ev.type == ABX_T_WAKEUP) r3_wakeup(id); else r3_receive(id, ev); }
        else if (ev.type == ABX_T_WAKEUP) dq_exec_wakeup(id, type); else dq_exec_receive(id, type, ev, x);
        a.agent_time = s.now + P.c.default_computation_delay_ns + addl_delay;
        c.sync(); if (c.onchip_writer()) regs_store(z, a); c.sync();
        c.agent_commit(id);
        paused = (x->rl_flags & DQF_PENDING) != 0;
      }
      flush();
    }
    rng_sync();
    c.sync();
    return paused;
  }

  // =================================================================================================
  // rmsc03 population (config/rmsc03.py): NoiseAgent, ValueAgent, MomentumAgent, POVMarketMakerAgent + the exchange's
  // QUERY_TRANSACTED_VOLUME (util/OrderBook.py:400-436).  Zero latency, computation delay 0.
  // Per-environment tables (HBM): idtab[0 .. MM_ORDER_CAP) = the market maker's open orders {id, price, signed qty, -},
  // idtab[MM_ORDER_CAP .. +tv_ring) = ring of transaction tuples {time lo, time hi, qty, record epoch | add epoch << 16},
  // lobs[k * MOM_MIDS/4 ...] = last MOM_MIDS doubled mid prices of momentum agent k.
  // =================================================================================================
  ABX_HD void tv_record(int32_t qty, uint32_t rec_epoch) {
    uint32_t i = s.ctr_latency++;                                                       // ring cursor (the latency stream is unused with zero latency)
    uint4 v; v.x = (uint32_t)(uint64_t)s.now; v.y = (uint32_t)((uint64_t)s.now >> 32); v.z = (uint32_t)qty; v.w = (rec_epoch & 0xffffu) | (s.trade_epoch << 16);
    c.id_store(MM_ORDER_CAP + (int)(i & (uint32_t)(P.tv_ring - 1)), v);
  }
  // get_transacted_volume :400-436: distinct (time, qty) tuples on surviving history records with time >= now - lookback
  ABX_HD int32_t transacted_volume(int64_t lookback) {
    uint32_t RING = (uint32_t)P.tv_ring, n = s.ctr_latency < RING ? s.ctr_latency : RING; int64_t start = s.now - lookback; int64_t sum = 0;
    uint32_t E = s.trade_epoch;
    if (s.ctr_latency > RING) { uint4 o = c.id_load(MM_ORDER_CAP + (int)(s.ctr_latency & (RING - 1))); if (((E - (o.w >> 16)) & 0xffffu) <= (uint32_t)P.c.stream_history) s.flags |= ABX_F_HISTORY_OVERFLOW; }  // a tuple that could still count was overwritten: capacity, like the queue / ladder / order pools
#pragma unroll 1
    for (uint32_t k = 0; k < n; k++) {                                                  // newest first
      uint4 v = c.id_load(MM_ORDER_CAP + (int)((s.ctr_latency - 1 - k) & (RING - 1)));
      int64_t t = (int64_t)((uint64_t)v.x | ((uint64_t)v.y << 32));
      if (((E - (v.w >> 16)) & 0xffffu) > (uint32_t)P.c.stream_history + 1) break;      // added more than stream_history + 1 rotations ago: nothing older can survive
      if (t < start || ((E - (v.w & 0xffffu)) & 0xffffu) > (uint32_t)P.c.stream_history) continue;
      bool dup = false;
#pragma unroll 1
      for (uint32_t j = 0; j < k && !dup; j++) {
        uint4 w = c.id_load(MM_ORDER_CAP + (int)((s.ctr_latency - 1 - j) & (RING - 1)));
        if (w.x == v.x && w.y == v.y && w.z == v.z && ((E - (w.w & 0xffffu)) & 0xffffu) <= (uint32_t)P.c.stream_history) dup = true;
      }
      if (!dup) sum += (int32_t)v.z;
    }
    return (int32_t)sum;
  }
  ABX_HD void r3_exch_receive(const Event &m) {                                         // ExchangeAgent.receiveMessage :129-340
    s.exch_comp_delay = P.c.exchange_computation_delay_ns;
    bool t_closed = s.now > P.c.mkt_close_ns;
    int32_t p[6] = {0, 0, 0, 0, 0, 0};
    bool is_query = m.kind == ABX_QUERY_SPREAD || m.kind == ABX_QUERY_TRANSACTED_VOLUME || (P3 && m.kind == ABX_QUERY_ORDER_STREAM);
    double lat = P3 ? m.lat_back : 0.0;                                                 // latency[0][sender] (zero in the zero-latency configs)
    if (t_closed && !is_query) { exch_send(m.sender, ABX_MKT_CLOSED, p, lat); return; }
    if (P3 && m.kind == ABX_MARKET_DATA_SUBSCRIPTION_REQUEST) {                         // updateSubscriptionDict :342-357: dict[agent] = [levels, freq, now]; a repeated request keeps its place
      int k = c.tab_find(MM_ORDER_CAP, (int)s.n_subs, (uint32_t)m.sender);
      if (k < 0) { if (s.n_subs >= (uint32_t)SUB_CAP) { s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW; return; } k = (int)s.n_subs++; }
      uint4 e4; e4.x = (uint32_t)m.sender; e4.y = (uint32_t)m.p[0]; e4.z = (uint32_t)(uint64_t)s.now; e4.w = (uint32_t)((uint64_t)s.now >> 32); c.id_store(MM_ORDER_CAP + k, e4);
      int64_t due = s.now + (agent_type_of(P.c, m.sender) == AT_MKM ? P.c.mkm_sub_freq_ns : P.c.mom_sub_freq_ns); if (due < s.next_pub) s.next_pub = due;
      return;
    }
    if (P3 && m.kind == ABX_QUERY_ORDER_STREAM) {                                       // :251-279 history[1 : length + 1]: REFERENCES to the bucket dicts of the last `length` trades.  The reply names them
      uint32_t E = s.trade_epoch, n = (uint32_t)m.p[0];                                 // by epoch {current epoch, how many}; the agent reads the log when it places its order, so it sees what the dicts hold THEN
      if (n > E - EPOCH_START) n = E - EPOCH_START; if (n > (uint32_t)P.c.stream_history) n = (uint32_t)P.c.stream_history;   // len(history) - 1 buckets exist besides the open one
      p[0] = (int32_t)E; p[1] = (int32_t)n; p[5] = t_closed ? 4 : 0; exch_send(m.sender, ABX_QUERY_ORDER_STREAM, p, lat); return;
    }
    if (m.kind == ABX_WHEN_MKT_OPEN) { s.exch_comp_delay = 0; exch_send(m.sender, ABX_WHEN_MKT_OPEN, p, lat); }
    else if (m.kind == ABX_WHEN_MKT_CLOSE) { s.exch_comp_delay = 0; exch_send(m.sender, ABX_WHEN_MKT_CLOSE, p, lat); }
    else if (m.kind == ABX_QUERY_TRANSACTED_VOLUME) {                                   // :280-303
      int64_t lookback = (int64_t)((uint64_t)(uint32_t)m.p[0] | ((uint64_t)(uint32_t)m.p[1] << 32));
      p[0] = transacted_volume(lookback); p[5] = t_closed ? 4 : 0; exch_send(m.sender, ABX_QUERY_TRANSACTED_VOLUME, p, lat);
    } else if (m.kind == ABX_QUERY_SPREAD) {
      s.c_query++; int f = 0; int nb = s.n_bid_lv, na = s.n_ask_lv;
      if (nb > 0) { p[0] = c.lv_price(0, nb - 1); p[1] = c.lv_qty(0, nb - 1); f |= 1; }
      if (na > 0) { p[2] = c.lv_price(1, na - 1); p[3] = c.lv_qty(1, na - 1); f |= 2; }
      if (t_closed) f |= 4;
      int32_t sn = (P.c.n_pov_exec && m.sender == P.c.n_agents - 1) ? snap_take(0, P.c.exec_kind == 2 ? 100 : 0x7fffffff) : 0;   // POVExecutionAgent asks for depth sys.maxsize (AggressiveAgent: 100): the levels are copied now
      p[4] = s.last_trade; p[5] = f; exch_send(m.sender, ABX_QUERY_SPREAD, p, P3 ? lat : bits_dbl((uint64_t)(uint32_t)sn));
    } else if (m.kind == ABX_LIMIT_ORDER) { s.c_limit++; if (!P3) s.ctr_kernel++; book_handle_limit((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], m.p[2], lat); c.sync(); trace_snap(); if (P3 && m.p[2] > 0) { s.book_update = s.now; exch_publish(); } }   // ctr_kernel doubles as the book-operation counter where the kernel stream is unused
    else if (m.kind == ABX_CANCEL_ORDER) { s.c_cancel++; if (!P3) s.ctr_kernel++; book_cancel((uint32_t)m.p[0], m.sender, m.p[4], m.p[1], lat); c.sync(); trace_snap(); if (P3) { if (cancel_found) s.book_update = s.now; exch_publish(); } }
  }
  // ExchangeAgent.publishOrderBookData :359-387: after every book operation, MARKET_DATA to each subscriber (in subscription order) whose `freq` ns have passed since its
  // last update.  The body's `levels` levels a side are copied into the subscriber's snapshot area NOW; the message carries the counts, the last trade and the slot.
  ABX_HD void exch_publish() {
    if (s.book_update < s.next_pub) return;                                             // nobody is due (a subscriber is served at most once per period; the book changes far more often)
    int64_t next = KEY_T_MAX;
#pragma unroll 1
    for (int k = 0; k < (int)s.n_subs; k++) {
      uint4 e4 = c.id_load(MM_ORDER_CAP + k); int agent = (int)e4.x, levels = (int)e4.y; int64_t last = (int64_t)((uint64_t)e4.z | ((uint64_t)e4.w << 32));
      int64_t freq = agent_type_of(P.c, agent) == AT_MKM ? P.c.mkm_sub_freq_ns : P.c.mom_sub_freq_ns;
      if (!(freq == 0 || (s.book_update > last && s.book_update - last >= freq))) { if (last + freq < next) next = last + freq; continue; }
      int nb = s.n_bid_lv < levels ? s.n_bid_lv : levels, na = s.n_ask_lv < levels ? s.n_ask_lv : levels;
      c.snap_store(k, 0, s.n_bid_lv, nb); c.snap_store(k, 1, s.n_ask_lv, na);
      int32_t p[6] = {nb, na, k, 0, s.last_trade, 0};
      exch_send(agent, ABX_MARKET_DATA, p, lat_zero() ? 0.0 : c.agent_lat_from(agent));
      if (n_out >= Ctx::OUTN - 3) flush();
      e4.z = (uint32_t)(uint64_t)s.book_update; e4.w = (uint32_t)((uint64_t)s.book_update >> 32); c.id_store(MM_ORDER_CAP + k, e4);
      if (s.book_update + freq < next) next = s.book_update + freq;
    }
    s.next_pub = next;
  }
  ABX_HD AgentAux *aux() { return reinterpret_cast<AgentAux *>(z->theta); }
  // TradingAgent.placeLimitOrder :309-349 for the staged trader; `track`: the agent later iterates self.orders (Value, market maker)
  ABX_HD void r3_place_limit(int id, int32_t size, bool buy, int32_t price, bool track) {
    uint32_t oid = gen_id();
    if (size <= 0) return;
    if (track) {
      int type = (int)((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT);
      if (type == AT_POVMM || (P3 && type == AT_MKM)) { if (a.n_orders < MM_ORDER_CAP) { uint4 v; v.x = oid; v.y = (uint32_t)price; v.z = (uint32_t)(buy ? size : -size); v.w = 0; c.id_store(a.n_orders, v); a.n_orders++; } else s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW; }
      else if (a.n_orders < AGENT_ORDER_CAP) { c.sync(); if (c.onchip_writer()) { int k = a.n_orders; z->oid[k] = oid; z->oprice[k] = price; z->oqty[k] = buy ? size : -size; } c.sync(); a.n_orders++; }
      else s.flags |= ABX_F_AGENT_ORDERS_OVERFLOW;
    }
    int32_t pl[6] = {(int32_t)oid, price, size, 0, buy, 0}; env_send(ABX_LIMIT_ORDER, pl, false);
    if (n_out >= Ctx::OUTN - 3) flush();
    (void)id;
  }
  ABX_HD void r3_cancel_all(int type) {                                                 // cancelOrders / cancelAllOrders
#pragma unroll 1
    for (int i = 0; i < a.n_orders; i++) {
      uint32_t oid; int32_t price, q;
      if (type == AT_POVMM || (P3 && type == AT_MKM)) { uint4 v = c.id_load(i); oid = v.x; price = (int32_t)v.y; q = (int32_t)v.z; } else { oid = z->oid[i]; price = z->oprice[i]; q = z->oqty[i]; }
      int32_t p[6] = {(int32_t)oid, price, q < 0 ? -q : q, 0, q > 0, 0}; env_send(ABX_CANCEL_ORDER, p, false);
      if (n_out >= Ctx::OUTN - 3) flush();
    }
  }
  ABX_HD void r3_wakeup(int id) {
    int type = (int)((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT);
    if (P3 && (type == AT_ZI || type == AT_HBL)) { zi_wakeup(id); return; }
    bool can_trade = ta_wakeup(a.flags);                                                // TradingAgent.wakeup :142-158
    uint32_t st = (a.flags & AF_STATE_MASK) >> AF_STATE_SHIFT;
    bool hours = (a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE);
    bool closed_done = (a.flags & AF_MKT_CLOSED) && (a.flags & AF_HAS_DAILY), closed_wait = (a.flags & AF_MKT_CLOSED) && !(a.flags & AF_HAS_DAILY);
    if (type == AT_NOISE) {                                                             // agent/NoiseAgent.py:82-112
      st = ST_INACTIVE;
      if (hours && !closed_done) { if (a.prev_wake > s.now) set_wakeup(id, a.prev_wake); { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); } st = ST_AWAITING_SPREAD; }
    } else if (type == AT_VALUE) {                                                      // agent/ValueAgent.py:100-138
      st = ST_INACTIVE;
      if (hours && !closed_done) {
        double delta_time = dmul(rng.std_exponential(S_AGENT0 + id, a.rng_ctr), P.inv_lambda_a);
        set_wakeup(id, s.now + py_round_i64(delta_time));
        if (!closed_wait) r3_cancel_all(type);
        { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); } st = ST_AWAITING_SPREAD;
      }
    } else if (type == AT_MOMENTUM) {                                                   // agent/examples/MomentumAgent.py:53-63
      if (P3 && P.c.mom_subscribe) { if (!(a.flags & AF_SUB_REQUESTED)) { int32_t p[6] = {1, 0, 0, 0, 0, 0}; env_send(ABX_MARKET_DATA_SUBSCRIPTION_REQUEST, p, false); a.flags |= AF_SUB_REQUESTED; } }   // :56-59 levels = 1, requested at the very first wake-up; wake-ups do nothing else
      else if (can_trade) { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); st = ST_AWAITING_SPREAD; }
    } else if (P3 && type == AT_MKM) {                                                  // agent/market_makers/MarketMakerAgent.py:66-77
      if (P.c.mkm_subscribe) { if (!(a.flags & AF_SUB_REQUESTED)) { int32_t p[6] = {P.c.mkm_num_levels, 0, 0, 0, 0, 0}; env_send(ABX_MARKET_DATA_SUBSCRIPTION_REQUEST, p, false); a.flags |= AF_SUB_REQUESTED; } }   // :69-72
      else if (can_trade) { r3_cancel_all(type); int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); st = ST_AWAITING_SPREAD; }
    } else if (type == AT_POVEXEC && P.c.exec_kind != 0) {                              // PassiveAgent.wakeup (passive_agent.py:40-58) / AggressiveAgent.wakeup (aggressive_agent.py:26-32): one action at the timestamp
      if (can_trade && s.now == P.c.pov_exec_start_ns) {
        if (P.c.exec_kind == 1 && P.c.exec_limit_price != 0) dq_place_limit(id, (int32_t)P.c.pov_exec_quantity, P.c.pov_exec_is_buy != 0, P.c.exec_limit_price);
        else { int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true); st = ST_AWAITING_SPREAD; }
      }
    } else if (type == AT_POVEXEC) {                                                    // agent/execution/baselines/pov_agent.py:55-64
      if (can_trade && exaux()->rem_qty > 0 && s.now < P.c.pov_exec_end_ns) {
        set_wakeup(id, s.now + P.c.pov_exec_freq_ns); dq_cancel_all(id);
        int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true);         // depth = sys.maxsize
        int32_t q[6] = {(int32_t)(uint32_t)(uint64_t)P.c.pov_exec_lookback_ns, (int32_t)(uint32_t)((uint64_t)P.c.pov_exec_lookback_ns >> 32), 0, 0, 0, 0}; env_send(ABX_QUERY_TRANSACTED_VOLUME, q, false);
        st = ST_AWAITING_TV;
      }
    } else {                                                                            // POVMarketMakerAgent.py:83-100 (with the getTransactedVolume alias)
      if (can_trade) {
        int32_t p[6] = {0, 0, 0, 0, 0, 0}; env_send(ABX_QUERY_SPREAD, p, true);
        int32_t q[6] = {(int32_t)(uint32_t)(uint64_t)P.c.mm_wake_ns, (int32_t)(uint32_t)((uint64_t)P.c.mm_wake_ns >> 32), 0, 0, 0, 0}; env_send(ABX_QUERY_TRANSACTED_VOLUME, q, false);
      }
    }
    a.flags = (a.flags & ~AF_STATE_MASK) | (st << AF_STATE_SHIFT);
  }
  ABX_HD void r3_value_place(int id) {                                                  // ValueAgent.updateEstimates :140-205 + placeOrder :207-243
    int stream = S_AGENT0 + id; AgentAux ax = *aux();
    int32_t r_now = oracle_advance(s.now >= P.c.mkt_close_ns ? P.c.mkt_close_ns - 1 : s.now);
    int32_t obs_t = (int32_t)py_round_i64(rng.normal(stream, a.rng_ctr, (double)r_now, P.sqrt_sigma_n));
    if (!(a.flags & AF_HAS_PREV)) { a.prev_wake = P.c.mkt_open_ns; a.flags |= AF_HAS_PREV; }
    double r_bar = P.c.r_bar, sigma_n = P.c.sigma_n;
    double delta = (double)(s.now - a.prev_wake);
    double d2 = (double)(P.c.mkt_close_ns - s.now); if (!(d2 > 0)) d2 = 0;
    double pw0 = exp_ni(dmul(delta, P.log_base_a)), pw1 = exp_ni(dmul(dmul(2.0, delta), P.log_base_a)), pw2 = exp_ni(dmul(d2, P.log_base_a));
    double r_tprime = dmul(dsub(1.0, pw0), r_bar); r_tprime = dadd(r_tprime, dmul(pw0, a.r_t));
    double sigma_tprime = dmul(pw1, a.sigma_t); sigma_tprime = dadd(sigma_tprime, dmul(dsub(1.0, pw1) / P.sigma_denom, P.c.sigma_s));
    double den = dadd(sigma_n, sigma_tprime);
    double r_t = dmul(sigma_n / den, r_tprime); r_t = dadd(r_t, dmul(sigma_tprime / den, (double)obs_t)); a.r_t = r_t;
    if (a.sigma_t != 0.0) a.sigma_t = dmul(sigma_n, a.sigma_t) / dadd(sigma_n, a.sigma_t);
    double r_T = dmul(dsub(1.0, pw2), r_bar); r_T = dadd(r_T, dmul(pw2, a.r_t));
    int32_t r_Ti = (int32_t)py_round_i64(r_T); a.prev_wake = s.now;
    bool buy; int32_t p;
    bool has_bid = (a.flags & AF_HAS_BID) && a.bid != 0, has_ask = (a.flags & AF_HAS_ASK) && a.ask != 0;
    if (has_bid && has_ask) {
      int32_t mid = (int32_t)((double)(a.ask + a.bid) / 2); int32_t spread = a.ask > a.bid ? a.ask - a.bid : a.bid - a.ask; int32_t adjust;
      if (rng.u01(S_GLOBAL, s.ctr_global) < P.c.value_percent_aggr) adjust = 0;          // np.random.rand() < percent_aggr (GLOBAL stream)
      else adjust = (int32_t)rng.randint(S_GLOBAL, s.ctr_global, (uint32_t)(P.c.value_depth_spread * spread - 1));
      if (r_Ti < mid) { buy = false; p = a.bid + adjust; } else { buy = true; p = a.ask - adjust; }
    } else { buy = rng.randint(S_GLOBAL, s.ctr_global, 1) != 0; p = r_Ti; }
    r3_place_limit(id, ax.size, buy, p, true);
  }
  ABX_HD void r3_momentum_place(int id) {                                               // MomentumAgent.placeOrders :78-93, ma :95-99
    bool has_bid = (a.flags & AF_HAS_BID) && a.bid != 0, has_ask = (a.flags & AF_HAS_ASK) && a.ask != 0;
    if (!has_bid || !has_ask) return;
    AgentAux ax = *aux();
    int k = id - (P3 ? 1 + P.c.n_mm_agents + P.c.groups[0].count + P.c.groups[1].count : 1 + P.c.n_noise_agents + P.c.n_value_agents + P.c.n_mm_agents);   // momentum agent index
    c.mid_store(k, ax.n_mids % MOM_MIDS, a.bid + a.ask);                                // 2 * mid: exact integer
    ax.n_mids++;
    int L = ax.n_mids;
    if (L > 20) { int64_t sum2 = c.mid_sum(k, L, 20); ax.avg20 = rint(dmul(((double)sum2 / 2) / 20, 100.0)) / 100.0; ax.mmflags |= MOF_HAS20; }
    if (L > 50) { int64_t sum2 = c.mid_sum(k, L, 50); ax.avg50 = rint(dmul(((double)sum2 / 2) / 50, 100.0)) / 100.0; ax.mmflags |= MOF_HAS50; }
    c.sync(); if (c.onchip_writer()) *aux() = ax; c.sync();
    if ((ax.mmflags & MOF_HAS20) && (ax.mmflags & MOF_HAS50)) { if (ax.avg20 >= ax.avg50) r3_place_limit(id, ax.size, true, a.ask, false); else r3_place_limit(id, ax.size, false, a.bid, false); }
  }
  ABX_HD void r3_receive(int id, const Event &m) {
    int type = (int)((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT);
    if (P3 && (type == AT_ZI || type == AT_HBL)) { zi_receive(id, m); return; }
    bool had = (a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE);
    if (m.kind == ABX_WHEN_MKT_OPEN) a.flags |= AF_HAS_OPEN;
    else if (m.kind == ABX_WHEN_MKT_CLOSE) a.flags |= AF_HAS_CLOSE;
    else if (m.kind == ABX_ORDER_EXECUTED) {                                            // orderExecuted :422-462
      int32_t q = m.p[2]; int32_t sq = m.p[4] ? q : -q; a.shares += sq; a.cash -= (int64_t)sq * m.p[3];
      if (type == AT_VALUE) { int i = orders_find((uint32_t)m.p[0]); if (i >= 0) { int32_t oq0 = z->oqty[i]; int32_t oq = oq0 < 0 ? -oq0 : oq0; if (q >= oq) orders_remove(i); else { c.sync(); if (c.onchip_writer()) z->oqty[i] = oq0 < 0 ? -(oq - q) : (oq - q); c.sync(); } } }
      else if (type == AT_POVMM || (P3 && type == AT_MKM)) r3_mm_order_update((uint32_t)m.p[0], q, false);
      else if (type == AT_POVEXEC) { dq_order_update(id, (uint32_t)m.p[0], q, false); ExecAux ex = *exaux(); ex.executed_sum += q; ex.n_executed++; ex.rem_qty = (int32_t)P.c.pov_exec_quantity - ex.executed_sum; exaux_store(ex); }   // handleOrderExecution :103-107
    } else if (m.kind == ABX_ORDER_CANCELLED) {
      if (type == AT_VALUE) { int i = orders_find((uint32_t)m.p[0]); if (i >= 0) orders_remove(i); }
      else if (type == AT_POVMM || (P3 && type == AT_MKM)) r3_mm_order_update((uint32_t)m.p[0], 0, true);
      else if (type == AT_POVEXEC) dq_order_update(id, (uint32_t)m.p[0], 0, true);
    } else if (m.kind == ABX_MKT_CLOSED) a.flags |= AF_MKT_CLOSED;
    else if (m.kind == ABX_QUERY_TRANSACTED_VOLUME) {                                   // :248-251,556-558
      if (m.p[5] & 4) a.flags |= AF_MKT_CLOSED;
      if (type == AT_POVEXEC) { ExecAux ex = *exaux(); ex.tv = m.p[0]; exaux_store(ex); }
      else { AgentAux ax = *aux(); ax.tv = m.p[0]; c.sync(); if (c.onchip_writer()) *aux() = ax; c.sync(); }
    } else if (m.kind == ABX_QUERY_SPREAD) {
      if (m.p[5] & 4) a.flags |= AF_MKT_CLOSED;
      a.last_trade = m.p[4]; a.flags |= AF_HAS_LAST;
      if (a.flags & AF_MKT_CLOSED) { a.daily_close = a.last_trade; a.flags |= AF_HAS_DAILY; }
      a.flags &= ~(AF_HAS_BID | AF_HAS_ASK);
      if (m.p[5] & 1) { a.flags |= AF_HAS_BID; a.bid = m.p[0]; a.bid_q = m.p[1]; } else { a.bid = 0; a.bid_q = 0; }
      if (m.p[5] & 2) { a.flags |= AF_HAS_ASK; a.ask = m.p[2]; a.ask_q = m.p[3]; } else { a.ask = 0; a.ask_q = 0; }
    } else if (P3 && m.kind == ABX_MARKET_DATA) {                                       // handleMarketData :539-546: known_bids / known_asks = the body's lists (read from the snapshot slot), last_trade
      a.last_trade = m.p[4]; a.flags |= AF_HAS_LAST; a.flags &= ~(AF_HAS_BID | AF_HAS_ASK);
      if (m.p[0] > 0) { int2 lv = c.snap_load(m.p[2], 0, 0); a.flags |= AF_HAS_BID; a.bid = lv.x; a.bid_q = lv.y; } else { a.bid = 0; a.bid_q = 0; }
      if (m.p[1] > 0) { int2 lv = c.snap_load(m.p[2], 1, 0); a.flags |= AF_HAS_ASK; a.ask = lv.x; a.ask_q = lv.y; } else { a.ask = 0; a.ask_q = 0; }
    }
    if ((a.flags & AF_HAS_OPEN) && (a.flags & AF_HAS_CLOSE) && !had) {                  // :258-268 getWakeFrequency per class
      int64_t off = type == AT_MOMENTUM ? P.c.mom_wake_ns : (type == AT_POVMM ? P.c.mm_wake_ns : (type == AT_POVEXEC ? (P.c.exec_kind != 0 ? P.c.pov_exec_start_ns - P.c.mkt_open_ns : P.c.pov_exec_freq_ns) : ((P3 && type == AT_MKM) ? P.c.mkm_wake_ns : rng.randint(S_AGENT0 + id, a.rng_ctr, 99))));
      set_wakeup(id, P.c.mkt_open_ns + off);
    }
    uint32_t st = (a.flags & AF_STATE_MASK) >> AF_STATE_SHIFT;
    if (type == AT_POVEXEC) {                                                           // POVExecutionAgent.receiveMessage :69-99
      if (m.kind == ABX_QUERY_SPREAD) { ExecAux e0 = *exaux(); e0.snap_n = m.x0; exaux_store(e0); }     // known_bids / known_asks
      if (P.c.exec_kind != 0) {                                                         // PassiveAgent.receiveMessage :60-70 / AggressiveAgent.receiveMessage :34-37 (the state is never left)
        if (st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) {
          bool buy = P.c.pov_exec_is_buy != 0;
          if (P.c.exec_kind == 1) { if (!(a.flags & (buy ? AF_HAS_BID : AF_HAS_ASK))) s.flags |= ABX_F_OBS_INVALID; else dq_place_limit(id, (int32_t)P.c.pov_exec_quantity, buy, buy ? a.bid : a.ask); }   // an empty side: limit price None in the reference
          else dq_place_market(id, (int32_t)P.c.pov_exec_quantity, buy);
        }
        return;
      }
      if (s.now > P.c.pov_exec_end_ns) return;
      ExecAux ex = *exaux();
      if (ex.rem_qty > 0 && st == ST_AWAITING_TV && m.kind == ABX_QUERY_TRANSACTED_VOLUME && s.now > P.c.pov_exec_start_ns) {
        int32_t qty = (int32_t)py_round_i64(dmul(P.c.pov_exec_pov, (double)ex.tv));
        dq_cancel_all(id);
        dq_place_market(id, qty, P.c.pov_exec_is_buy != 0);
      }
      return;
    }
    if (P3 && type == AT_MKM && P.c.mkm_subscribe) {                                    // MarketMakerAgent.receiveMessage :108-118 + placeOrders :120-139 (subscription mode)
      if ((a.flags & AF_SUB_REQUESTED) && m.kind == ABX_MARKET_DATA) {
        r3_cancel_all(type);
        int num_levels = 1 + (int)rng.randint(S_AGENT0 + id, a.rng_ctr, 3);             // randint(1, len(levels_quote_dict)): 1..4 levels
        int nb = m.p[0], na = m.p[1], k = m.p[2];
        if (nb > 0 && na > 0) {
          int32_t size = mkm_half(P.c.mkm_min_size + rng.randint(S_AGENT0 + id, a.rng_ctr, (uint32_t)(P.c.mkm_max_size - P.c.mkm_min_size - 1)));
          double w0 = num_levels == 1 ? 1.0 : num_levels == 2 ? 0.5 : num_levels == 3 ? 0.34 : 0.25, w1 = num_levels == 2 ? 0.5 : num_levels == 3 ? 0.33 : 0.25;   // DEFAULT_LEVELS_QUOTE_DICT :6-12
          int32_t qp[2][4], qv[2][4]; int nq[2] = {0, 0};                                 // buy_quotes / sell_quotes: dicts keyed by price (a repeated key keeps its place and takes the new volume)
          for (int i = 0; i < num_levels; i++) {
            int32_t vol = (int32_t)py_round_i64(dmul(i == 0 ? w0 : w1, (double)size));
            for (int side = 0; side < 2; side++) {
              int n = side ? na : nb; int32_t px = i < n ? c.snap_load(k, side, i).x : c.snap_load(k, side, n - 1).x + (side ? 1 : -1);   // IndexError branch: one cent beyond the last level
              int f = 0; while (f < nq[side] && qp[side][f] != px) f++;
              if (f == nq[side]) nq[side]++;
              qp[side][f] = px; qv[side][f] = vol;
            }
          }
          for (int side = 0; side < 2; side++) for (int f = 0; f < nq[side]; f++) r3_place_limit(id, qv[side][f], side == 0, qp[side][f], true);
        }
      }
      return;
    }
    if (P3 && type == AT_MKM) {                                                         // MarketMakerAgent.receiveMessage :79-108 (polling mode)
      if (st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) {
        r3_cancel_all(type);
        int32_t mid = a.last_trade, spread = 10;                                        // self.last_spread = 10 is never updated
        if ((a.flags & AF_HAS_BID) && a.bid != 0 && (a.flags & AF_HAS_ASK) && a.ask != 0) { mid = (int32_t)((double)(a.ask + a.bid) / 2); int32_t d = a.ask > a.bid ? a.ask - a.bid : a.bid - a.ask; spread = (int32_t)((double)d / 2); }
#pragma unroll 1
        for (int i = 0; i < 2 * P.c.mkm_num_levels; i++) {                              // :98-103 a fresh size per level
          int32_t size = mkm_half(P.c.mkm_min_size + rng.randint(S_AGENT0 + id, a.rng_ctr, (uint32_t)(P.c.mkm_max_size - P.c.mkm_min_size - 1)));
          r3_place_limit(id, size, true, mid - spread - i, true); r3_place_limit(id, size, false, mid + spread + i, true);
        }
        set_wakeup(id, s.now + P.c.mkm_wake_ns);
        a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT);
      }
      return;
    }
    if (type == AT_POVMM) {                                                             // POVMarketMakerAgent.receiveMessage :102-150
      AgentAux ax = *aux();
      if (m.kind == ABX_QUERY_TRANSACTED_VOLUME && (ax.mmflags & MMF_AW_VOL)) {         // updateOrderSize :152-156
        int32_t qty = (int32_t)py_round_i64(dmul(P.c.mm_pov, (double)ax.tv)); ax.order_size = qty >= P.c.mm_min_order_size ? qty : P.c.mm_min_order_size; ax.mmflags &= ~MMF_AW_VOL; }
      if (m.kind == ABX_QUERY_SPREAD && (ax.mmflags & MMF_AW_SPREAD)) {
        bool has_bid = (a.flags & AF_HAS_BID) && a.bid != 0, has_ask = (a.flags & AF_HAS_ASK) && a.ask != 0;
        if (has_bid && has_ask) { ax.last_mid = (int32_t)((double)(a.ask + a.bid) / 2); ax.mmflags |= MMF_HAS_MID; ax.mmflags &= ~MMF_AW_SPREAD; } }
      bool go = !(ax.mmflags & MMF_AW_SPREAD) && !(ax.mmflags & MMF_AW_VOL);
      if (go) ax.mmflags |= MMF_AW_SPREAD | MMF_AW_VOL;                                 // self.state = self.initialiseState()
      c.sync(); if (c.onchip_writer()) *aux() = ax; c.sync();
      if (go) {
        r3_cancel_all(type);                                                            // the cancelled orders stay in self.orders until ORDER_CANCELLED arrives
        int32_t mid = ax.last_mid, highest_bid = mid - 1, lowest_ask = mid + P.c.mm_window_size;   // computeOrdersToPlace :158-177 (anchor bottom)
        int32_t lowest_bid = highest_bid - P.c.mm_num_ticks, highest_ask = lowest_ask + P.c.mm_num_ticks;
#pragma unroll 1
        for (int32_t px = lowest_bid; px <= highest_bid; px++) r3_place_limit(id, ax.order_size, true, px, true);
#pragma unroll 1
        for (int32_t px = lowest_ask; px <= highest_ask; px++) r3_place_limit(id, ax.order_size, false, px, true);
        set_wakeup(id, s.now + P.c.mm_wake_ns);
      }
      return;
    }
    if (type == AT_MOMENTUM) {                                                          // MomentumAgent.receiveMessage :65-76
      if (P3 && P.c.mom_subscribe) { if ((a.flags & AF_SUB_REQUESTED) && m.kind == ABX_MARKET_DATA && m.p[0] > 0 && m.p[1] > 0) r3_momentum_place(id); return; }   // :71-75 placeOrders(bids[0][0], asks[0][0])
      if (st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD) { r3_momentum_place(id); set_wakeup(id, s.now + P.c.mom_wake_ns); a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT); }
      return;
    }
    if (st == ST_AWAITING_SPREAD && m.kind == ABX_QUERY_SPREAD && !(a.flags & AF_MKT_CLOSED)) {      // NoiseAgent :123-129 / ValueAgent :245-251
      if (type == AT_NOISE) {                                                           // NoiseAgent.placeOrder :114-121
        bool buy = rng.randint(S_GLOBAL, s.ctr_global, 1) != 0; int32_t size = aux()->size;
        bool has_bid = (a.flags & AF_HAS_BID) && a.bid != 0, has_ask = (a.flags & AF_HAS_ASK) && a.ask != 0;
        if (buy && has_ask) r3_place_limit(id, size, true, a.ask, false); else if (!buy && has_bid) r3_place_limit(id, size, false, a.bid, false);
      } else r3_value_place(id);
      a.flags = (a.flags & ~AF_STATE_MASK) | (ST_AWAITING_WAKEUP << AF_STATE_SHIFT);
    }
  }
  // the market maker's self.orders: fill / cancel bookkeeping (TradingAgent.orderExecuted :445-452, orderCancelled :480-483)
  ABX_HD void r3_mm_order_update(uint32_t oid, int32_t fill, bool cancel) {
    int f = c.tab_find(0, a.n_orders, oid);
    if (f < 0) return;
    uint4 v = c.id_load(f); int32_t q = (int32_t)v.z, aq = q < 0 ? -q : q;
    if (!cancel && fill < aq) { v.z = (uint32_t)(q < 0 ? -(aq - fill) : (aq - fill)); c.id_store(f, v); return; }
    c.tab_remove(0, a.n_orders, f);                                                     // dict deletion keeps the order of the others
    a.n_orders--;
  }
  // Kernel.runner hot loop for the rmsc03 population
  ABX_HD void r3_run(int64_t until) {
#pragma unroll 1
    while (!(s.flags & ABX_F_DONE)) {
      uint64_t khi; uint32_t kuniq; int grp;
      bool any = c.q_min(khi, kuniq, grp);
      if (!any || !(s.now <= P.c.stop_ns) || (rng.tape() && (rng.err & ABX_F_TAPE_UNDERRUN))) { s.flags |= ABX_F_DONE; break; }   // an exhausted tape hands out zeros: with zero delays an agent would re-wake itself at the same instant for ever
      if (key_time(khi) > until) break;
      Event ev; c.q_fetch(grp, ev);
      s.now = ev.t; s.ttl++;
      if (INSTR && P.c.hash_pops) s.pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s.pop_hash, ev.t), ev.recipient), ev.type), ev.type == ABX_T_MESSAGE ? (int64_t)ev.uniq : -1);
      if (INSTR && P.c.trace_cap > 0) {
        abx_trace_rec r; r.tag = 0; r.a = ev.recipient; r.t = ev.t; for (int i = 0; i < 16; i++) r.v[i] = 0;
        r.v[0] = ev.type; r.v[1] = ev.type == ABX_T_MESSAGE ? (int32_t)ev.uniq : -1; r.v[2] = ev.kind; trace_rec(r);
      }
      addl_delay = 0;
      int id = ev.recipient; int64_t at = s.exch_time;
      if (id != 0) { z = c.agent_stage(id); regs_load(a, z); at = a.agent_time; }
      bool in_future = at > s.now;
      c.q_settle(in_future, at);
      if (in_future) continue;
      s.q_count--; self_id = id;
      if (id == 0) {
        if (ev.type == ABX_T_MESSAGE) r3_exch_receive(ev);
        s.exch_time = s.now + s.exch_comp_delay + addl_delay;
      } else {
        if (ev.type == ABX_T_WAKEUP) r3_wakeup(id); else r3_receive(id, ev);
        a.agent_time = s.now + P.c.default_computation_delay_ns + addl_delay;
        c.sync(); if (c.onchip_writer()) regs_store(z, a); c.sync();
        c.agent_commit(id);
      }
      flush();
    }
    rng_sync();
  }
  // kernelStopping: ValueAgent.kernelStopping :49-61 observes the fundamental (advances the oracle); holdings are read by the host
  ABX_HD void r3_finalize() {
    if (P3) { finalize(); return; }
    int64_t sum_sh = 0, sum_cash = 0;
#pragma unroll 1
    for (int id = 1; id < P.c.n_agents; id++) {
      z = c.agent_stage(id); regs_load(a, z);
      if (agent_type_of(P.c, id) == AT_VALUE) { int64_t cur = a.agent_time - P.c.default_computation_delay_ns; oracle_advance(cur >= P.c.mkt_close_ns ? P.c.mkt_close_ns - 1 : cur); }
      sum_sh += a.shares; sum_cash += a.cash;
    }
    s.sum_shares = sum_sh; s.sum_cash = sum_cash; rng_sync();
  }

  // ---- Kernel.runner :310-311 kernelStopping for every trader, in id order (ZeroIntelligenceAgent.py:80-123) ----
  ABX_HD void finalize() {
    int64_t sum_sh = 0, sum_cash = 0; int q_max = P.c.q_max;
#pragma unroll 1
    for (int id = 1; id < P.c.n_agents; id++) {
      z = c.agent_stage(id); regs_load(a, z);
      if (P3) { int t = (int)((a.flags & AF_TYPE_MASK) >> AF_TYPE_SHIFT); if (t != AT_ZI && t != AT_HBL) { sum_sh += a.shares; sum_cash += a.cash; continue; } }   // market maker / momentum agents have no closing valuation
      double hr = dmul(rint((double)a.shares / 100.0), 100.0);                          // round(int, -2): half-even on hundreds
      int H = (int)(hr / 100.0);
      int64_t cur = a.agent_time - P.c.default_computation_delay_ns;                    // Agent.currentTime of the trader's last event
      int32_t rT = oracle_advance(cur >= P.c.mkt_close_ns ? P.c.mkt_close_ns - 1 : cur);                    // observePrice(sigma_n=0)
      int64_t surplus = 0;
      if (H > 0) { for (int x = 1; x <= H; x++) { int k = x + q_max - 1; if (k < 2 * q_max) surplus += z->theta[k]; } }
      else if (H < 0) { for (int x = H + 1; x <= 0; x++) { int k = x + q_max - 1; if (k >= 0) surplus += z->theta[k]; } surplus = -surplus; }
      surplus += (int64_t)rT * H; surplus += a.cash - P.c.starting_cash;
      sum_sh += a.shares; sum_cash += a.cash;
      c.sync();
      if (c.onchip_writer()) z->surplus = surplus;
      c.sync();
      c.agent_commit(id);
    }
    s.sum_shares = sum_sh; s.sum_cash = sum_cash;
    rng_sync();
  }
};

}  // namespace abx
