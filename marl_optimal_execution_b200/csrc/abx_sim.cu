// abx_sim.cu -- kernels and C ABI of the batched ABIDES simulator (sm_100a).  See include/abides_b200.h.
//
// Launch geometry: one warp per environment, ABX_WARPS_PER_CTA warps per CTA, each warp owning a private slice of
// dynamic shared memory (abx_warp.cuh).  A warp runs its environment's whole event loop (Kernel.py:190-292) up to
// the requested simulated time in one launch; environments never communicate.
#include <type_traits>
#include <cstddef>
#include <cuda_runtime.h>
#include <stdio.h>
#include <new>
#include <vector>
#include <unordered_map>
#include "abx_warp.cuh"
#include "abx_host_common.h"

using namespace abx;

#ifndef ABX_WARPS_PER_CTA
#define ABX_WARPS_PER_CTA 1
#endif

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
__global__ void abx_init_agents_kernel(SimParams P, const uint64_t *__restrict__ seeds, uint32_t *__restrict__ init_err) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)P.n_envs * P.c.n_agents;
  if (t >= total) return;
  int env = (int)(t / P.c.n_agents), id = (int)(t % P.c.n_agents);
  if (id == 0) return;
  uint32_t err = 0;
  init_agent_record(P, env, id, seeds ? seeds[env] : 0, P.agents + t, &err);
  if (err) atomicOr(init_err + env, err);
}

__device__ __forceinline__ EnvState env_load(const EnvState *g) {
  EnvState s; const uint4 *src = reinterpret_cast<const uint4 *>(g); uint4 *dst = reinterpret_cast<uint4 *>(&s);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(EnvState) / 16); i++) dst[i] = __ldcg(src + i);
  return s;
}
// SKIP: bit i set = the i-th 16-byte quad of the record cannot have changed in this kernel and is not written back -- its fields are then dead after the load and
// free their registers for the whole event loop (the run kernels live at the 128-register limit of 16 one-warp CTAs per SM)
template <uint32_t SKIP = 0u>
__device__ __forceinline__ void env_store(EnvState *g, const EnvState &s, int lane) {
  if (lane == 0) { const uint4 *src = reinterpret_cast<const uint4 *>(&s); uint4 *dst = reinterpret_cast<uint4 *>(g);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(EnvState) / 16); i++) if (!((SKIP >> i) & 1u)) __stcg(dst + i, src[i]); }
}
// quads of EnvState the sparse_zi event loop never writes without instrumentation: {sum_shares, sum_cash} (kernelStopping), {draw_n, evt_n, book_flags, episode},
// {hist_n, n_subs, book_update}, {next_pub, pad}.  (Also skipping the lone cold words pop_hash and trace_n frees three more registers and measures 1.8 % SLOWER:
// profiles/r02_optimisation_log.md.)
constexpr uint32_t ENV_QUADS_COLD_ZI = (1u << 10) | (1u << 12) | (1u << 13) | (1u << 14);
#ifndef ABX_NO_COLD_QUADS_MORE
// the other production kernels: the rmsc03 population runs at zero latency (no kernel-stream block, quad 11) and has neither order-stream history nor subscriptions
// (quads 13, 14); the rmsc01 / rmsc02 population keeps those two; the replay shapes (ABIDESEnv, DDQN config) have no fundamental either ({or_t, ms_t}, {ms_v, pop_hash})
// but follow re-priced levels (book_flags, quad 12)
constexpr uint32_t ENV_QUADS_COLD_R3 = (1u << 10) | (1u << 11) | (1u << 12) | (1u << 13) | (1u << 14);
constexpr uint32_t ENV_QUADS_COLD_P3 = (1u << 10) | (1u << 12), ENV_QUADS_COLD_P3_ZEROLAT = ENV_QUADS_COLD_P3 | (1u << 11);
constexpr uint32_t ENV_QUADS_COLD_REPLAY = (1u << 2) | (1u << 3) | (1u << 10) | (1u << 11) | (1u << 13) | (1u << 14);
#else
constexpr uint32_t ENV_QUADS_COLD_R3 = 0, ENV_QUADS_COLD_P3 = 0, ENV_QUADS_COLD_P3_ZEROLAT = 0, ENV_QUADS_COLD_REPLAY = 0;
#endif
static_assert(offsetof(EnvState, or_t) == 32 && offsetof(EnvState, ms_v) == 48 && offsetof(EnvState, kblk) == 176, "ENV_QUADS_COLD_* follow the EnvState layout");
static_assert(offsetof(EnvState, sum_shares) == 160 && offsetof(EnvState, draw_n) == 192 && offsetof(EnvState, hist_n) == 208 && offsetof(EnvState, next_pub) == 224, "ENV_QUADS_COLD_ZI follows the EnvState layout");

template <class Ctx>
__device__ void reset_env_with(const SimParams &P, EnvState &s, int env, unsigned char *smem) {
  Ctx ctx(P, env, smem);
  ctx.q_clear();
  Sim<Ctx> sim(ctx, P, s, env);
  sim.reset_env();
  ctx.store_onchip(sim.s);
  env_store(P.env + env, sim.s, ctx.lane);
}
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA)
abx_reset_env_kernel(SimParams P, const uint64_t *__restrict__ seeds, const uint32_t *__restrict__ init_err, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  EnvState s; init_env_state(P, seeds ? seeds[env] : 0, s); s.flags |= init_err[env];
  if (P.c.population == 0) reset_env_with<WarpCtxNearQ>(P, s, env, smem + warp * smem_per_warp);     // sparse_zi: near/far event queue
  else reset_env_with<WarpCtx>(P, s, env, smem + warp * smem_per_warp);
}

// One instantiation per (RNG mode, latency model, instrumentation): the production path (Philox, no instrumentation)
// carries neither the tape-replay branches nor the parity hash/trace code.
#ifndef ABX_RUN_MINB
#define ABX_RUN_MINB 16       // one-warp CTAs per SM the run kernel is compiled for
#endif
template <int RNG, int LAT, bool INSTR, int SHAPE = SHAPE_ZI>
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA, ABX_RUN_MINB)
abx_run_kernel(SimParams P, int64_t until_ns, const int64_t *__restrict__ until_each, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  typedef typename std::conditional<SHAPE == SHAPE_ZI, WarpCtxNearQ, WarpCtx>::type Ctx;
  Ctx ctx(P, env, smem + warp * smem_per_warp);
  EnvState s = env_load(P.env + env);
  if (s.flags & ABX_F_DONE) return;
  ctx.load_onchip(s);
  Sim<Ctx, RNG, LAT, INSTR, SHAPE> sim(ctx, P, s, env);
  if (SHAPE == SHAPE_R3 || SHAPE == SHAPE_P3) sim.r3_run(until_each ? until_each[env] : until_ns); else sim.run(until_each ? until_each[env] : until_ns);
  ctx.store_onchip(sim.s);
#ifndef ABX_NO_COLD_QUADS
  env_store<INSTR ? 0u : SHAPE == SHAPE_ZI ? ENV_QUADS_COLD_ZI : SHAPE == SHAPE_R3 ? ENV_QUADS_COLD_R3 : SHAPE == SHAPE_P3 ? (LAT == ABX_LAT_ZERO ? ENV_QUADS_COLD_P3_ZEROLAT : ENV_QUADS_COLD_P3) : 0u>(P.env + env, sim.s, ctx.lane);
#else
  env_store(P.env + env, sim.s, ctx.lane);
#endif
}
typedef void (*run_kernel_fn)(SimParams, int64_t, const int64_t *, size_t);
static run_kernel_fn run_kernel_for(const abx_sim_config &c) {
  bool instr = c.trace_cap > 0 || c.hash_pops != 0 || c.draw_log_cap > 0 || c.event_ring_cap > 0; int r = c.rng_mode, l = c.latency_model;
  if (c.population == 1) {                       // config/rmsc03.py population: zero latency
    if (r == ABX_RNG_PHILOX) return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_R3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_ZERO, false, SHAPE_R3>;
    return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_ZERO, true, SHAPE_R3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_ZERO, false, SHAPE_R3>;
  }
  if (c.population == 3 && l == ABX_LAT_MATRIX_NOISE) {   // config/rmsc02.py: the same population under the pairwise latency matrix + noise
    if (r == ABX_RNG_PHILOX) return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_MATRIX_NOISE, true, SHAPE_P3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_MATRIX_NOISE, false, SHAPE_P3>;
    return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_MATRIX_NOISE, true, SHAPE_P3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_MATRIX_NOISE, false, SHAPE_P3>;
  }
  if (c.population == 3) {                       // config/rmsc01.py population: the rmsc03 loop over more agent classes
    if (r == ABX_RNG_PHILOX) return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_P3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_PHILOX, ABX_LAT_ZERO, false, SHAPE_P3>;
    return instr ? (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_ZERO, true, SHAPE_P3> : (run_kernel_fn)abx_run_kernel<ABX_RNG_TAPE, ABX_LAT_ZERO, false, SHAPE_P3>;
  }
#define PICK(R, L) (instr ? (run_kernel_fn)abx_run_kernel<R, L, true> : (run_kernel_fn)abx_run_kernel<R, L, false>)
  if (r == ABX_RNG_PHILOX) return l == ABX_LAT_CUBIC ? PICK(ABX_RNG_PHILOX, ABX_LAT_CUBIC) : PICK(ABX_RNG_PHILOX, ABX_LAT_MATRIX_NOISE);
  return l == ABX_LAT_CUBIC ? PICK(ABX_RNG_TAPE, ABX_LAT_CUBIC) : PICK(ABX_RNG_TAPE, ABX_LAT_MATRIX_NOISE);
#undef PICK
}

__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA)
abx_finalize_kernel(SimParams P, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  WarpCtx ctx(P, env, smem + warp * smem_per_warp);
  EnvState s = env_load(P.env + env);
  if (P.c.population == 1) { Sim<WarpCtx, -1, ABX_LAT_ZERO, true, SHAPE_R3> sim(ctx, P, s, env); sim.r3_finalize(); env_store(P.env + env, sim.s, ctx.lane); return; }   // INSTR: the ValueAgents' closing observations go to the draw log
  if (P.c.population == 3) { Sim<WarpCtx, -1, -1, true, SHAPE_P3> sim(ctx, P, s, env); sim.r3_finalize(); env_store(P.env + env, sim.s, ctx.lane); return; }
  Sim<WarpCtx> sim(ctx, P, s, env);
  sim.finalize();
  env_store(P.env + env, sim.s, ctx.lane);
}

// ---- ABIDESEnv shape: reset and step (GymKernel.initRunner / stepRunner) ----
// Which environments a reset touches: all (mask NULL, mode RESET_ALL), those with mask[e] != 0, or those whose event loop has ended (RESET_DONE: the
// auto-reset pass after a step).  advance_day: the environment moves on to its next replayed day (episode + 1), otherwise it restarts the same day.
enum { RESET_ALL = 0, RESET_MASK = 1, RESET_DONE = 2 };
__device__ __forceinline__ bool reset_selected(const SimParams &P, int env, const uint8_t *mask, int mode, int advance_day, uint32_t &episode) {
  episode = 0;
  if (mode == RESET_ALL) return true;
  if (mode == RESET_MASK && !mask[env]) return false;
  EnvState old = env_load(P.env + env);
  if (mode == RESET_DONE && !(old.flags & ABX_F_DONE)) return false;
  episode = old.episode + (advance_day ? 1u : 0u);
  return true;
}
template <bool SMALLQ>
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA)
abx_env_reset_kernel(SimParams P, const uint8_t *__restrict__ mask, int mode, int advance_day, size_t smem_per_warp) {
  typedef Sim<WarpCtxT<SMALLQ ? 1 : 0>, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_ENV> EnvSimInstr;
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  uint32_t episode; if (!reset_selected(P, env, mask, mode, advance_day, episode)) return;
  WarpCtxT<SMALLQ ? 1 : 0> ctx(P, env, smem + warp * smem_per_warp);
  ctx.set_episode(episode); if (mode != RESET_ALL) ctx.clear_tables();
  EnvState s; init_env_state(P, 0, s); s.last_trade = -1; s.episode = episode;   // no oracle: OrderBook.last_trade stays None (ExchangeAgent.py:97-102)
  init_envx(P, *ctx.envx()); ctx.sync();
  ctx.q_clear();
  EnvSimInstr sim(ctx, P, s, env);
  sim.env_reset();
  ctx.store_onchip(sim.s); ctx.envx_store();
  env_store(P.env + env, sim.s, ctx.lane);
}

#ifndef ABX_STEP_MINB
#define ABX_STEP_MINB 16      // one-warp CTAs per SM the step kernels are compiled for (register budget = 65536 / (32 * ABX_STEP_MINB))
#endif
template <bool INSTR, bool SMALLQ>
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA, ABX_STEP_MINB)
abx_env_step_kernel(SimParams P, const double *__restrict__ actions, double *__restrict__ obs, double *__restrict__ reward,
                    uint8_t *__restrict__ done, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  WarpCtxT<SMALLQ ? 1 : 0> ctx(P, env, smem + warp * smem_per_warp);
  EnvState s = env_load(P.env + env);
  bool was_done = (s.flags & ABX_F_DONE) != 0;
  if (!was_done) {
    ctx.set_episode(s.episode); ctx.envx_load(); ctx.load_onchip(s);
    Sim<WarpCtxT<SMALLQ ? 1 : 0>, ABX_RNG_PHILOX, ABX_LAT_ZERO, INSTR, SHAPE_ENV> sim(ctx, P, s, env);
    sim.env_step(actions[3 * env], actions[3 * env + 1], actions[3 * env + 2]);
    ctx.store_onchip(sim.s); ctx.envx_store();
    env_store<INSTR ? 0u : ENV_QUADS_COLD_REPLAY>(P.env + env, sim.s, ctx.lane);
    s.flags = sim.s.flags;
  }
  // obs (9 x fp64), reward, done: lanes 0..8 write one observation value each (72 contiguous bytes per environment)
  const EnvX *x = ctx.envx();
  if (ctx.lane < 9) obs[9 * env + ctx.lane] = (!was_done && x->obs_len) ? x->obs[ctx.lane] : 0.0;
  if (ctx.lane == 0) { if (reward) reward[env] = 0.0; done[env] = (s.flags & ABX_F_DONE) ? 1 : 0; }
}

// ---- DDQN execution shape: reset and decision step ----
typedef Sim<WarpCtxHybridQ, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_DQ> DqSimInstr;

__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA)
abx_dq_reset_kernel(SimParams P, const uint64_t *__restrict__ seeds, const int32_t *__restrict__ mom_sizes, const uint8_t *__restrict__ mask, int mode, int advance_day, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  uint32_t episode; if (!reset_selected(P, env, mask, mode, advance_day, episode)) return;
  WarpCtxHybridQ ctx(P, env, smem + warp * smem_per_warp);
  ctx.set_episode(episode); if (mode != RESET_ALL) ctx.clear_tables();
  uint64_t seed = (seeds ? seeds[env] : 0) + episode;                   // a later episode of the environment draws its MomentumAgent sizes afresh
  EnvState s; init_env_state(P, seed, s); s.last_trade = -1; s.episode = episode;   // no oracle: OrderBook.last_trade stays None until the first trade
  init_envx(P, *ctx.envx()); ctx.sync();
  ctx.q_clear();
  for (int id = 2 + ctx.lane; id < P.c.n_agents; id += 32)             // one lane per trader record
    init_agent_record_dq(P, env, id, seed, (mom_sizes && id < 2 + P.dq_n_mom) ? mom_sizes[(size_t)env * P.dq_n_mom + id - 2] : -1, P.agents + (size_t)env * P.c.n_agents + id);
  __syncwarp();
  DqSimInstr sim(ctx, P, s, env);
  sim.env_reset();
  ctx.store_onchip(sim.s); ctx.envx_store();
  env_store(P.env + env, sim.s, ctx.lane);
}

template <bool INSTR>
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA, ABX_STEP_MINB)      // 16 one-warp CTAs per SM: 128 registers (unbounded ptxas takes 168 -> 12 per SM)
abx_dq_step_kernel(SimParams P, const int32_t *__restrict__ actions, double *__restrict__ obs, double *__restrict__ trans, double *__restrict__ reward,
                   uint8_t *__restrict__ done, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  WarpCtxHybridQ ctx(P, env, smem + warp * smem_per_warp);
  EnvState s = env_load(P.env + env);
  bool was_done = (s.flags & ABX_F_DONE) != 0, paused = false;
  if (!was_done) {
    ctx.set_episode(s.episode); ctx.envx_load(); ctx.load_onchip(s);
    Sim<WarpCtxHybridQ, ABX_RNG_PHILOX, ABX_LAT_ZERO, INSTR, SHAPE_DQ> sim(ctx, P, s, env);
    paused = sim.dq_step(actions ? actions[env] : 0);
    ctx.store_onchip(sim.s); ctx.envx_store();
    env_store<INSTR ? 0u : ENV_QUADS_COLD_REPLAY>(P.env + env, sim.s, ctx.lane);
    s.flags = sim.s.flags;
  }
  // outputs: 8 + 6 + 1 doubles and the done byte per environment
  const EnvX *x = ctx.envx();
  if (ctx.lane < 8) obs[8 * env + ctx.lane] = paused ? x->obs[ctx.lane] : 0.0;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double tv = qnan, rw = 0.0;
  if (!was_done && P.dq_has_ddqn) {
    const ZiAgent *z = ctx.agent_stage(P.c.n_agents - 1); const ExecAux *ex = reinterpret_cast<const ExecAux *>(z->oid);
    if (ex->exflags & EXF_E_VALID) {
      int l = ctx.lane;
      tv = l == 0 ? (double)ex->e_s[0] : l == 1 ? (double)ex->e_s[1] : l == 2 ? (double)ex->e_a : l == 3 ? (double)ex->e_sp[0] : l == 4 ? (double)ex->e_sp[1] : ((ex->exflags & EXF_E_R) ? ex->e_r : qnan);
    }
    rw = ex->step_reward;
  }
  if (ctx.lane < 6) trans[6 * env + ctx.lane] = tv;
  if (ctx.lane == 0) { if (reward) reward[env] = rw; done[env] = (s.flags & ABX_F_DONE) ? 1 : 0; }
}

// ---- Book surface: op-tape replay through bare books ----
__global__ void __launch_bounds__(32 * ABX_WARPS_PER_CTA)
abx_book_replay_kernel(SimParams P, const int64_t *__restrict__ ops, int64_t n_ops, int fresh, size_t smem_per_warp) {
  extern __shared__ __align__(16) unsigned char smem[];
  int warp = threadIdx.x >> 5, env = blockIdx.x * ABX_WARPS_PER_CTA + warp;
  if (env >= P.n_envs) return;
  WarpCtx ctx(P, env, smem + warp * smem_per_warp);
  EnvState s;
  if (fresh) { init_env_state(P, 0, s); s.last_trade = -1; ctx.q_clear(); }      // OrderBook.__init__: last_trade None
  else { s = env_load(P.env + env); s.flags &= ~ABX_F_DONE; ctx.load_onchip(s); }
  Sim<WarpCtx, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_BOOK> sim(ctx, P, s, env);
  sim.book_replay(ops, n_ops);
  ctx.store_onchip(sim.s);
  env_store(P.env + env, sim.s, ctx.lane);
}

__global__ void abx_stats_kernel(SimParams P, abx_env_stats *__restrict__ out) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= P.n_envs) return;
  const EnvState &s = P.env[env];
  size_t l = (size_t)env * 2 * P.c.level_cap; int nb = s.n_bid_lv, na = s.n_ask_lv;
  abx_env_stats o;
  fill_stats(s, nb ? P.lv_price[l + nb - 1] : 0, nb ? P.lv_qty[l + nb - 1] : 0,
             na ? P.lv_price[l + P.c.level_cap + na - 1] : 0, na ? P.lv_qty[l + P.c.level_cap + na - 1] : 0, &o);
  out[env] = o;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static thread_local char g_cuda_err[512] = "";
#define CU(call)                                                                                         \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); return ABX_ERR_CUDA; } } while (0)

// inside the create functions, once the handle exists: a failing CUDA call must not leak it
#define CUH(call)                                                                                        \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); abx_sim_destroy(h); return ABX_ERR_CUDA; } } while (0)

struct abx_sim {
  SimParams P; int n_envs, device; bool reset_done; size_t smem_per_warp; int64_t bytes, launches;
  uint64_t *d_seeds; uint32_t *d_init_err; abx_env_stats *d_stats; int64_t *d_until;
  uint64_t *d_tbits; uint8_t *d_tkinds; int64_t *d_toff;
  bool is_env, is_dq, have_seeds, have_msizes; EnvStreamHost *st; EnvDaysHost *dh; int4 *d_daytab, *d_daytab2; int32_t *d_xid, *d_xfirst; int64_t *d_ts; int32_t *d_first; int4 *d_rows; double *d_act, *d_obs, *d_rew; uint8_t *d_done;
  int32_t *d_iact, *d_msizes, *d_sched; double *d_trans;
  int auto_reset;                       // 0 off, 1 restart the same day, 2 move on to the next day: applied to finished environments after every step
  bool is_book; int64_t *d_ops; int64_t ops_cap; std::unordered_map<int64_t, int32_t> *book_ids;
};

template <class T> static int dalloc(T **p, size_t n, int64_t *acc) {
  size_t b = n * sizeof(T); if (b == 0) { *p = nullptr; return ABX_OK; }
  CU(cudaMalloc((void **)p, b)); *acc += (int64_t)b; return ABX_OK;
}
static inline unsigned grid_for(int n_envs) { return (unsigned)((n_envs + ABX_WARPS_PER_CTA - 1) / ABX_WARPS_PER_CTA); }

extern "C" {

__global__ void abx_selftest_log_unit_kernel(const double *__restrict__ x, double *__restrict__ y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = log_unit(x[i]);
}
int32_t abx_selftest_log_unit(const double *x_host, double *y_host, int32_t n, int32_t device) {
  if (!x_host || !y_host || n < 1) return ABX_ERR_ARG;
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present", device); return ABX_ERR_CUDA; }
  CU(cudaSetDevice(device));
  double *dx = nullptr, *dy = nullptr; int32_t st = ABX_OK;
  if (cudaMalloc(&dx, sizeof(double) * n) != cudaSuccess || cudaMalloc(&dy, sizeof(double) * n) != cudaSuccess) st = ABX_ERR_CUDA;
  if (st == ABX_OK && cudaMemcpy(dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) st = ABX_ERR_CUDA;
  if (st == ABX_OK) { abx_selftest_log_unit_kernel<<<(n + 127) / 128, 128>>>(dx, dy, n); if (cudaMemcpy(y_host, dy, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) st = ABX_ERR_CUDA; }
  if (st != ABX_OK) snprintf(g_cuda_err, sizeof(g_cuda_err), "%s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(dx); cudaFree(dy); return st;
}
__global__ void abx_selftest_exp_kernel(const double *__restrict__ x, double *__restrict__ y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = exp_ni(x[i]);
}
int32_t abx_selftest_exp(const double *x_host, double *y_host, int32_t n, int32_t device) {
  if (!x_host || !y_host || n < 1) return ABX_ERR_ARG;
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present", device); return ABX_ERR_CUDA; }
  CU(cudaSetDevice(device));
  double *dx = nullptr, *dy = nullptr; int32_t st = ABX_OK;
  if (cudaMalloc(&dx, sizeof(double) * n) != cudaSuccess || cudaMalloc(&dy, sizeof(double) * n) != cudaSuccess) st = ABX_ERR_CUDA;
  if (st == ABX_OK && cudaMemcpy(dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) st = ABX_ERR_CUDA;
  if (st == ABX_OK) { abx_selftest_exp_kernel<<<(n + 127) / 128, 128>>>(dx, dy, n); if (cudaMemcpy(y_host, dy, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) st = ABX_ERR_CUDA; }
  if (st != ABX_OK) snprintf(g_cuda_err, sizeof(g_cuda_err), "%s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(dx); cudaFree(dy); return st;
}
const char *abx_strerror(int32_t st) { return status_string(st); }
const char *abx_last_cuda_error(void) { return g_cuda_err; }
int32_t abx_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int32_t abx_config_sparse_zi(int32_t variant, abx_sim_config *cfg) { return config_sparse_zi(variant, cfg); }
int32_t abx_config_rmsc03(abx_sim_config *cfg) { return config_rmsc03(cfg); }
int32_t abx_config_rmsc03_pov(abx_sim_config *cfg) { return config_rmsc03_pov(cfg); }
int32_t abx_config_rmsc01(abx_sim_config *cfg) { return config_rmsc01(cfg); }
int32_t abx_config_rmsc02(abx_sim_config *cfg) { return config_rmsc02(cfg); }

int32_t abx_sim_destroy(abx_sim *h) {
  if (!h) return ABX_OK;
  cudaSetDevice(h->device);
  void *ptrs[] = {h->P.qkey, h->P.qpay0, h->P.qpay1, h->P.qcache, h->P.agents, h->P.lv_price, h->P.lv_qty, h->P.lv_ht, h->P.nodes, h->P.env,
                  h->P.trace, h->P.draw_log, h->P.evt, h->P.hlog, h->P.snap, h->d_seeds, h->d_init_err, h->d_stats, h->d_until, h->d_tbits, h->d_tkinds, h->d_toff,
                  h->P.envx, h->P.idtab, h->P.idbook, h->P.lobs, h->d_ts, h->d_first, h->d_rows, h->d_act, h->d_obs, h->d_rew, h->d_done, h->d_iact, h->d_msizes, h->d_sched, h->d_trans, h->d_ops, h->d_daytab, h->d_daytab2, h->d_xid, h->d_xfirst};
  for (void *p : ptrs) if (p) cudaFree(p);
  delete h->st; delete h->dh; delete h->book_ids; delete h; return ABX_OK;
}

int32_t abx_sim_create(const abx_sim_config *cfg, int32_t n_envs, int32_t device, abx_sim **out) {
  if (!out || n_envs < 1 || config_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  int ndev = 0; CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present (%d visible)", device, ndev); return ABX_ERR_CUDA; }
  CU(cudaSetDevice(device));
  abx_sim *h = new (std::nothrow) abx_sim(); if (!h) return ABX_ERR_ARG;
  memset(h, 0, sizeof(*h)); h->P.c = *cfg; h->P.n_envs = n_envs; h->n_envs = n_envs; h->device = device; derive_params(h->P);
  h->smem_per_warp = (warp_smem_bytes(*cfg, false, false, false, cfg->population == 0) + 15) & ~(size_t)15;
  size_t smem_cta = h->smem_per_warp * ABX_WARPS_PER_CTA;
  if (smem_cta > 227 * 1024) { delete h; return ABX_ERR_ARG; }
  const abx_sim_config &c = *cfg; size_t E = (size_t)n_envs; int st;
#define DA(ptr, n) if ((st = dalloc(&(ptr), (n), &h->bytes)) != ABX_OK) { abx_sim_destroy(h); return st; }
  DA(h->P.qkey, E * c.queue_cap) DA(h->P.qpay0, E * c.queue_cap) DA(h->P.qpay1, E * c.queue_cap) DA(h->P.qcache, E * h->P.n_qgroups)
  DA(h->P.agents, E * c.n_agents) DA(h->P.lv_price, E * 2 * c.level_cap) DA(h->P.lv_qty, E * 2 * c.level_cap) DA(h->P.lv_ht, E * 2 * c.level_cap)
  DA(h->P.nodes, E * c.order_cap) DA(h->P.env, E) DA(h->P.trace, E * (size_t)c.trace_cap) DA(h->P.draw_log, E * (size_t)c.draw_log_cap) DA(h->P.evt, E * (size_t)c.event_ring_cap)
  DA(h->d_seeds, E) DA(h->d_init_err, E) DA(h->d_stats, E) DA(h->d_until, E)
  if (c.population == 1) { h->P.dq_order_base = MM_ORDER_CAP + h->P.tv_ring; h->P.n_ids = h->P.dq_order_base + (c.n_pov_exec ? EXEC_ORDER_CAP : 0);       // market maker orders + transaction ring [+ POV execution agent orders]; momentum mids
    DA(h->P.idtab, E * h->P.n_ids) DA(h->P.lobs, E * LOB_CAP * 3)
    if (c.n_pov_exec) { h->P.n_snap = 1; h->P.snap_depth = c.level_cap; DA(h->P.snap, E * 2 * (size_t)c.level_cap) } }      // POVExecutionAgent asks for depth sys.maxsize
  if (c.population == 3) { h->P.n_ids = MM_ORDER_CAP + SUB_CAP; h->P.n_snap = SUB_CAP; h->P.snap_depth = SUB_LEVELS; DA(h->P.snap, E * (size_t)SUB_CAP * 2 * SUB_LEVELS) DA(h->P.idtab, E * h->P.n_ids) DA(h->P.lobs, E * (size_t)lob_stride_of(c)) DA(h->P.hlog, E * hist_stride_of(c)) }   // market maker orders; momentum mids; order-history log
#undef DA
  if (smem_cta > 48 * 1024) {
    CUH(cudaFuncSetAttribute((const void *)run_kernel_for(*cfg), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute(abx_reset_env_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute(abx_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
  }
  CUH(cudaMemset(h->P.agents, 0, E * c.n_agents * sizeof(ZiAgent)));
  *out = h; return ABX_OK;
}
int64_t abx_sim_device_bytes(const abx_sim *h) { return h ? h->bytes : 0; }
int64_t abx_sim_launch_count(const abx_sim *h) { return h ? h->launches : 0; }

static int32_t do_reset(abx_sim *h, bool have_seeds, cudaStream_t st) {
  const abx_sim_config &c = h->P.c;
  CU(cudaMemsetAsync(h->d_init_err, 0, sizeof(uint32_t) * h->n_envs, st));
  int64_t total = (int64_t)h->n_envs * c.n_agents;
  abx_init_agents_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(h->P, have_seeds ? h->d_seeds : nullptr, h->d_init_err);
  abx_reset_env_kernel<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, st>>>(h->P, have_seeds ? h->d_seeds : nullptr, h->d_init_err, h->smem_per_warp);
  h->launches += 2;
  CU(cudaGetLastError());
  h->reset_done = true; return ABX_OK;
}

int32_t abx_sim_reset_philox(abx_sim *h, const uint64_t *seeds, void *stream) {
  if (!h || !seeds || h->P.c.rng_mode != ABX_RNG_PHILOX) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(h->d_seeds, seeds, sizeof(uint64_t) * h->n_envs, cudaMemcpyHostToDevice, st));
  return do_reset(h, true, st);
}

static int32_t reset_tape_impl(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to,
                               const double *lat_from, void *stream);
int32_t abx_sim_reset_tape(abx_sim *h, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to,
                           const double *lat_from, void *stream) { return reset_tape_impl(h, 0, bits, kinds, off, lat_to, lat_from, stream); }
int32_t abx_sim_reset_tape_shared(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to,
                                  const double *lat_from, void *stream) {
  if (!h || n_tapes < 1 || n_tapes > h->n_envs || !lat_to || !lat_from) return ABX_ERR_ARG;
  size_t na = (size_t)h->P.c.n_agents, E = (size_t)h->n_envs;
  std::vector<double> lt(E * na), lf(E * na);                                   // per-environment copies of the shared runs' latency vectors
  for (size_t e = 0; e < E; e++) { memcpy(&lt[e * na], lat_to + (e % n_tapes) * na, sizeof(double) * na); memcpy(&lf[e * na], lat_from + (e % n_tapes) * na, sizeof(double) * na); }
  int32_t rc = reset_tape_impl(h, n_tapes, bits, kinds, off, lt.data(), lf.data(), stream); if (rc != ABX_OK) return rc;
  CU(cudaStreamSynchronize((cudaStream_t)stream));                              // lt / lf are local staging buffers
  return ABX_OK;
}
static int32_t reset_tape_impl(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to,
                               const double *lat_from, void *stream) {
  if (!h || !bits || !kinds || !off || !lat_to || !lat_from || h->P.c.rng_mode != ABX_RNG_TAPE) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  h->P.n_tapes = n_tapes;
  size_t nS = (size_t)(n_tapes > 0 ? n_tapes : h->n_envs) * h->P.n_streams; int64_t total = off[nS]; if (total < 0) return ABX_ERR_ARG;
  if (h->d_tbits) { cudaFree(h->d_tbits); h->d_tbits = nullptr; } if (h->d_tkinds) { cudaFree(h->d_tkinds); h->d_tkinds = nullptr; } if (h->d_toff) { cudaFree(h->d_toff); h->d_toff = nullptr; }
  int64_t dummy = 0;
  if (dalloc(&h->d_tbits, (size_t)total + 1, &dummy) || dalloc(&h->d_tkinds, (size_t)total + 1, &dummy) || dalloc(&h->d_toff, nS + 1, &dummy)) return ABX_ERR_CUDA;
  CU(cudaMemcpyAsync(h->d_tbits, bits, sizeof(uint64_t) * total, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_tkinds, kinds, (size_t)total, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_toff, off, sizeof(int64_t) * (nS + 1), cudaMemcpyHostToDevice, st));
  h->P.tape_bits = h->d_tbits; h->P.tape_kinds = h->d_tkinds; h->P.tape_off = h->d_toff;
  // latency[a][0] / latency[0][a] into the trader records (strided 8-byte columns of the 192-byte records)
  size_t n = (size_t)h->n_envs * h->P.c.n_agents;
  CU(cudaMemcpy2DAsync((char *)h->P.agents + offsetof(ZiAgent, lat_to), sizeof(ZiAgent), lat_to, sizeof(double), sizeof(double), n, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpy2DAsync((char *)h->P.agents + offsetof(ZiAgent, lat_from), sizeof(ZiAgent), lat_from, sizeof(double), sizeof(double), n, cudaMemcpyHostToDevice, st));
  return do_reset(h, false, st);
}

int32_t abx_sim_run(abx_sim *h, int64_t until_ns, void *stream) {
  if (!h) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  run_kernel_for(h->P.c)<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, until_ns, nullptr, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}

int32_t abx_sim_run_each(abx_sim *h, const int64_t *until_ns_host, void *stream) {
  if (!h || !until_ns_host) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->d_until, until_ns_host, sizeof(int64_t) * h->n_envs, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  run_kernel_for(h->P.c)<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, 0, h->d_until, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}

int32_t abx_sim_finalize(abx_sim *h, void *stream) {
  if (!h) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  abx_finalize_kernel<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}

int32_t abx_sim_stats_device(abx_sim *h, abx_env_stats *out_dev, void *stream) {
  if (!h || !out_dev) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  abx_stats_kernel<<<(h->n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->P, out_dev);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}
int32_t abx_sim_stats(abx_sim *h, abx_env_stats *out, void *stream) {
  if (!h || !out) return ABX_ERR_ARG;
  int32_t st = abx_sim_stats_device(h, h->d_stats, stream); if (st != ABX_OK) return st;
  CU(cudaMemcpyAsync(out, h->d_stats, sizeof(abx_env_stats) * h->n_envs, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return ABX_OK;
}

int32_t abx_sim_holdings(abx_sim *h, int32_t env, int64_t *out, void *stream) {
  if (!h || h->is_env || !out || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device));
  int n = h->P.c.n_agents; ZiAgent *tmp = (ZiAgent *)malloc(sizeof(ZiAgent) * n); if (!tmp) return ABX_ERR_ARG;
  cudaError_t e = cudaMemcpyAsync(tmp, h->P.agents + (size_t)env * n, sizeof(ZiAgent) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) { free(tmp); snprintf(g_cuda_err, sizeof(g_cuda_err), "holdings copy: %s", cudaGetErrorString(e)); return ABX_ERR_CUDA; }
  for (int id = 1; id < n; id++) { const ZiAgent &z = tmp[id]; int64_t *r = out + 5 * (id - 1);      // agent/TradingAgent.py:124-126, markToMarket :609-633
    r[0] = id; r[1] = z.shares; r[2] = z.cash; r[3] = z.cash + (int64_t)z.shares * ((z.flags & AF_HAS_LAST) ? z.last_trade : 0); r[4] = z.surplus; }
  free(tmp); return ABX_OK;
}

int32_t abx_sim_pov_exec(abx_sim *h, int32_t env, int64_t *out, void *stream) {
  if (!h || h->is_env || !out || env < 0 || env >= h->n_envs || h->P.c.population != 1 || !h->P.c.n_pov_exec) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); ZiAgent z;
  CU(cudaMemcpyAsync(&z, h->P.agents + (size_t)env * h->P.c.n_agents + (h->P.c.n_agents - 1), sizeof(z), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  const ExecAux *ex = reinterpret_cast<const ExecAux *>(z.oid); out[0] = ex->rem_qty; out[1] = ex->n_executed; out[2] = z.n_orders; return ABX_OK;
}

int32_t abx_sim_book_snapshot(abx_sim *h, int32_t env, int32_t is_bid, int32_t depth, int32_t *out, int32_t *n_levels, void *stream) {
  if (!h || !out || !n_levels || env < 0 || env >= h->n_envs || depth < 0) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  EnvState s; CU(cudaMemcpyAsync(&s, h->P.env + env, sizeof(s), cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  int side = is_bid ? 0 : 1, n = side ? s.n_ask_lv : s.n_bid_lv, m = depth < n ? depth : n; *n_levels = m; if (m == 0) return ABX_OK;
  int32_t *p = (int32_t *)malloc(sizeof(int32_t) * 2 * m); if (!p) return ABX_ERR_ARG;
  size_t base = ((size_t)env * 2 + side) * h->P.c.level_cap + (n - m);
  cudaError_t e = cudaMemcpyAsync(p, h->P.lv_price + base, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p + m, h->P.lv_qty + base, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { free(p); snprintf(g_cuda_err, sizeof(g_cuda_err), "snapshot copy: %s", cudaGetErrorString(e)); return ABX_ERR_CUDA; }
  for (int k = 0; k < m; k++) { out[2 * k] = p[m - 1 - k]; out[2 * k + 1] = p[m + m - 1 - k]; }   // best level is stored last
  free(p); return ABX_OK;
}

int32_t abx_sim_trace(abx_sim *h, int32_t env, abx_trace_rec *out, int32_t max_recs, int32_t *n_recs, void *stream) {
  if (!h || !out || !n_recs || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  EnvState s; CU(cudaMemcpyAsync(&s, h->P.env + env, sizeof(s), cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  int n = (int)s.trace_n; if (n > max_recs) n = max_recs; if (n > h->P.c.trace_cap) n = h->P.c.trace_cap;
  if (n > 0) { CU(cudaMemcpyAsync(out, h->P.trace + (size_t)env * h->P.c.trace_cap, sizeof(abx_trace_rec) * n, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); }
  if (h->is_env) { const std::vector<int64_t> &ido = h->dh ? h->dh->days[(int)(((uint32_t)env + s.episode) % (uint32_t)h->dh->days.size())].id_orig : h->st->id_orig;
    for (int i = 0; i < n; i++) if (out[i].tag == 1 && (uint32_t)out[i].v[1] >= REPLAY_ID_BASE) out[i].v[1] = (int32_t)ido[(uint32_t)out[i].v[1] - REPLAY_ID_BASE]; }
  *n_recs = n; return ABX_OK;
}

int32_t abx_sim_draw_log(abx_sim *h, int32_t env, abx_draw_rec *out, int32_t max_recs, int32_t *n_recs, void *stream) {
  if (!h || !out || !n_recs || env < 0 || env >= h->n_envs || h->is_env) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  EnvState s; CU(cudaMemcpyAsync(&s, h->P.env + env, sizeof(s), cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  int n = (int)s.draw_n; if (n > max_recs) n = max_recs; if (n > h->P.c.draw_log_cap) n = h->P.c.draw_log_cap;
  static_assert(sizeof(abx_draw_rec) == sizeof(uint4), "draw log entry layout");
  if (n > 0) { CU(cudaMemcpyAsync(out, h->P.draw_log + (size_t)env * h->P.c.draw_log_cap, sizeof(uint4) * n, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); }
  *n_recs = n; return ABX_OK;
}

__global__ void abx_event_counts_kernel(SimParams P, uint32_t *__restrict__ out) { int e = blockIdx.x * blockDim.x + threadIdx.x; if (e < P.n_envs) out[e] = P.env[e].evt_n; }
int32_t abx_sim_events_device(abx_sim *h, abx_event_rec *out_dev, uint32_t *counts_dev, void *stream) {
  if (!h || h->is_env || !out_dev || !counts_dev || h->P.c.event_ring_cap <= 0) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  static_assert(sizeof(abx_event_rec) == sizeof(uint4), "event record layout");
  CU(cudaMemcpyAsync(out_dev, h->P.evt, sizeof(uint4) * (size_t)h->n_envs * h->P.c.event_ring_cap, cudaMemcpyDeviceToDevice, st));
  abx_event_counts_kernel<<<(h->n_envs + 127) / 128, 128, 0, st>>>(h->P, counts_dev);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}
int32_t abx_sim_agent_init(abx_sim *h, int32_t env, int32_t *theta, double *lat_to, double *lat_from, int32_t *sizes, int64_t *wakes, void *stream) {
  if (!h || h->is_env || env < 0 || env >= h->n_envs) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  int n = h->P.c.n_agents; std::vector<ZiAgent> tmp(n);
  CU(cudaMemcpyAsync(tmp.data(), h->P.agents + (size_t)env * n, sizeof(ZiAgent) * n, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st));
  agent_init_rows(h->P, tmp.data(), theta, lat_to, lat_from, sizes, wakes);
  return ABX_OK;
}

static int32_t dq_reset_launch(abx_sim *h, const uint8_t *mask_dev, int mode, int advance_day, cudaStream_t st) {
  abx_dq_reset_kernel<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, st>>>(h->P, h->have_seeds ? h->d_seeds : nullptr, h->have_msizes ? h->d_msizes : nullptr,
                                                                                                                  mask_dev, mode, advance_day, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}

// ---------------- ABIDESEnv shape ----------------
int32_t abx_env_config_default(abx_env_config *cfg) { return env_config_default(cfg); }

int32_t abx_env_create_days(const abx_env_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out) {
  static_assert(ENV_QUEUE_MIN == SMALLQ_CAP, "queue_cap floor of the ABIDESEnv shape == slots of its on-chip queue");
  if (!out || n_envs < 1 || env_shape_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  EnvDaysHost *dh = new (std::nothrow) EnvDaysHost(); if (!dh) return ABX_ERR_ARG;
  if (env_build_days(stream5, row_offsets, n_days, 4LL * cfg->n_horizon + 16, *dh) != ABX_OK) { delete dh; return ABX_ERR_ARG; }
  int ndev = 0; cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || device < 0 || device >= ndev) { delete dh; snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present (%d visible): %s", device, ndev, cudaGetErrorString(ce)); return ABX_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) { delete dh; return ABX_ERR_CUDA; }
  abx_sim *h = new (std::nothrow) abx_sim(); if (!h) { delete dh; return ABX_ERR_ARG; }
  memset(h, 0, sizeof(*h)); h->is_env = true; h->dh = dh; h->n_envs = n_envs; h->device = device;
  env_fill_params(*cfg, h->P); h->P.n_envs = n_envs;
  h->P.n_ts = (int)dh->ts.size(); h->P.n_rows = (int)dh->rows.size(); h->P.n_ids = dh->max_ids; h->P.n_days = n_days;
  h->smem_per_warp = (warp_smem_bytes(h->P.c, true, warp_small_queue(h->P.c)) + 15) & ~(size_t)15;
  size_t smem_cta = h->smem_per_warp * ABX_WARPS_PER_CTA;
  if (smem_cta > 227 * 1024) { abx_sim_destroy(h); return ABX_ERR_ARG; }
  const abx_sim_config &c = h->P.c; size_t E = (size_t)n_envs; int stt;
#define DA(ptr, n) if ((stt = dalloc(&(ptr), (n), &h->bytes)) != ABX_OK) { abx_sim_destroy(h); return stt; }
  DA(h->P.qkey, E * c.queue_cap) DA(h->P.qpay0, E * c.queue_cap) DA(h->P.qpay1, E * c.queue_cap) DA(h->P.qcache, E * h->P.n_qgroups)
  DA(h->P.lv_price, E * 2 * c.level_cap) DA(h->P.lv_qty, E * 2 * c.level_cap) DA(h->P.lv_ht, E * 2 * c.level_cap)
  DA(h->P.nodes, E * c.order_cap) DA(h->P.env, E) DA(h->P.trace, E * (size_t)c.trace_cap) DA(h->d_stats, E)
  DA(h->P.envx, E) DA(h->P.idtab, E * h->P.n_ids) DA(h->P.idbook, E * h->P.n_ids) DA(h->P.lobs, E * LOB_CAP * 3)
  DA(h->d_ts, dh->ts.size()) DA(h->d_first, dh->first.size()) DA(h->d_rows, dh->rows.size()) DA(h->d_daytab, dh->day_tab.size())
  DA(h->d_daytab2, dh->day_tab2.size()) DA(h->d_xid, dh->xid.size() + 1) DA(h->d_xfirst, dh->xfirst.size() + 1)
  DA(h->d_act, E * 3) DA(h->d_obs, E * 9) DA(h->d_rew, E) DA(h->d_done, E)
#undef DA
  CUH(cudaMemcpy(h->d_ts, dh->ts.data(), sizeof(int64_t) * dh->ts.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_first, dh->first.data(), sizeof(int32_t) * dh->first.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_rows, dh->rows.data(), sizeof(int4) * dh->rows.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_daytab, dh->day_tab.data(), sizeof(int4) * dh->day_tab.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_daytab2, dh->day_tab2.data(), sizeof(int4) * dh->day_tab2.size(), cudaMemcpyHostToDevice));
  if (!dh->xid.empty()) { CUH(cudaMemcpy(h->d_xid, dh->xid.data(), sizeof(int32_t) * dh->xid.size(), cudaMemcpyHostToDevice)); CUH(cudaMemcpy(h->d_xfirst, dh->xfirst.data(), sizeof(int32_t) * dh->xfirst.size(), cudaMemcpyHostToDevice)); }
  h->P.st_ts = h->d_ts; h->P.st_first = h->d_first; h->P.st_rows = h->d_rows; h->P.day_tab = h->d_daytab; h->P.day_tab2 = h->d_daytab2; h->P.st_xid = h->d_xid; h->P.st_xfirst = h->d_xfirst;
  if (smem_cta > 48 * 1024) {
    CUH(cudaFuncSetAttribute(abx_env_reset_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute(abx_env_reset_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute((abx_env_step_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute((abx_env_step_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute((abx_env_step_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute((abx_env_step_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
  }
  *out = h; return ABX_OK;
}
int32_t abx_env_create(const abx_env_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out) {
  int64_t off[2] = {0, n_rows}; return abx_env_create_days(cfg, stream5, off, 1, n_envs, device, out);
}

static int32_t env_reset_launch(abx_sim *h, const uint8_t *mask_dev, int mode, int advance_day, cudaStream_t st) {
  if (warp_small_queue(h->P.c)) abx_env_reset_kernel<true><<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, st>>>(h->P, mask_dev, mode, advance_day, h->smem_per_warp);
  else abx_env_reset_kernel<false><<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, st>>>(h->P, mask_dev, mode, advance_day, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  return ABX_OK;
}
int32_t abx_env_reset(abx_sim *h, void *stream) {
  if (!h || !h->is_env || h->is_dq || h->is_book) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemsetAsync(h->P.idtab, 0, sizeof(uint4) * (size_t)h->n_envs * h->P.n_ids, st));
  CU(cudaMemsetAsync(h->P.idbook, 0, sizeof(uint2) * (size_t)h->n_envs * h->P.n_ids, st));
  int32_t rc = env_reset_launch(h, nullptr, RESET_ALL, 0, st); if (rc != ABX_OK) return rc;
  h->reset_done = true; return ABX_OK;
}
int32_t abx_env_reset_mask(abx_sim *h, const uint8_t *mask_dev, int32_t advance_day, void *stream) {
  if (!h || !h->is_env || h->is_book || !mask_dev) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  return h->is_dq ? dq_reset_launch(h, mask_dev, RESET_MASK, advance_day ? 1 : 0, (cudaStream_t)stream) : env_reset_launch(h, mask_dev, RESET_MASK, advance_day ? 1 : 0, (cudaStream_t)stream);
}
int32_t abx_env_set_auto_reset(abx_sim *h, int32_t mode) {
  if (!h || !h->is_env || h->is_book || mode < 0 || mode > 2) return ABX_ERR_ARG;
  h->auto_reset = mode; return ABX_OK;
}

int32_t abx_env_step(abx_sim *h, const double *actions_dev, double *obs_dev, double *reward_dev, uint8_t *done_dev, void *stream) {
  if (!h || !h->is_env || !actions_dev || !obs_dev || !done_dev) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  bool instr = h->P.c.trace_cap > 0 || h->P.c.hash_pops != 0;
  bool smallq = warp_small_queue(h->P.c);
#define ENV_STEP(I, Q) abx_env_step_kernel<I, Q><<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, actions_dev, obs_dev, reward_dev, done_dev, h->smem_per_warp)
  if (instr) { if (smallq) ENV_STEP(true, true); else ENV_STEP(true, false); }
  else { if (smallq) ENV_STEP(false, true); else ENV_STEP(false, false); }
#undef ENV_STEP
  h->launches += 1;
  CU(cudaGetLastError());
  if (h->auto_reset) return env_reset_launch(h, nullptr, RESET_DONE, h->auto_reset == 2, (cudaStream_t)stream);   // finished episodes start over before the next step
  return ABX_OK;
}

int32_t abx_env_step_host(abx_sim *h, const double *actions, double *obs, double *reward, uint8_t *done, void *stream) {
  if (!h || !h->is_env || !actions || !obs || !done) return ABX_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream; size_t E = (size_t)h->n_envs;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->d_act, actions, sizeof(double) * 3 * E, cudaMemcpyHostToDevice, st));
  int32_t rc = abx_env_step(h, h->d_act, h->d_obs, h->d_rew, h->d_done, stream); if (rc != ABX_OK) return rc;
  CU(cudaMemcpyAsync(obs, h->d_obs, sizeof(double) * 9 * E, cudaMemcpyDeviceToHost, st));
  if (reward) CU(cudaMemcpyAsync(reward, h->d_rew, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(done, h->d_done, E, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ABX_OK;
}

// ---------------- DDQN execution shape ----------------
int32_t abx_dq_config_default(abx_dq_config *cfg) { return dq_config_default(cfg); }

int32_t abx_dq_create_days(const abx_dq_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out) {
  if (!out || n_envs < 1 || dq_config_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  EnvDaysHost *dh = new (std::nothrow) EnvDaysHost(); if (!dh) return ABX_ERR_ARG;
  if (env_build_days(stream5, row_offsets, n_days, dq_max_generated_ids(*cfg), *dh) != ABX_OK) { delete dh; return ABX_ERR_ARG; }
  int ndev = 0; cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || device < 0 || device >= ndev) { delete dh; snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present (%d visible): %s", device, ndev, cudaGetErrorString(ce)); return ABX_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) { delete dh; return ABX_ERR_CUDA; }
  abx_sim *h = new (std::nothrow) abx_sim(); if (!h) { delete dh; return ABX_ERR_ARG; }
  memset(h, 0, sizeof(*h)); h->is_env = true; h->is_dq = true; h->dh = dh; h->n_envs = n_envs; h->device = device;
  dq_fill_params(*cfg, h->P); h->P.n_envs = n_envs;
  h->P.n_ts = (int)dh->ts.size(); h->P.n_rows = (int)dh->rows.size(); h->P.dq_order_base = dh->max_ids; h->P.dq_id_limit = dh->min_id; h->P.n_days = n_days;
  int n_exec = cfg->n_twap + (cfg->has_ddqn ? 1 : 0);
  h->P.n_ids = h->P.dq_order_base + n_exec * EXEC_ORDER_CAP;
  h->smem_per_warp = (warp_smem_bytes(h->P.c, true, false, true) + 15) & ~(size_t)15;
  size_t smem_cta = h->smem_per_warp * ABX_WARPS_PER_CTA;
  if (smem_cta > 227 * 1024) { abx_sim_destroy(h); return ABX_ERR_ARG; }
  const abx_sim_config &c = h->P.c; size_t E = (size_t)n_envs; int stt;
#define DA(ptr, n) if ((stt = dalloc(&(ptr), (n), &h->bytes)) != ABX_OK) { abx_sim_destroy(h); return stt; }
  DA(h->P.qkey, E * c.queue_cap) DA(h->P.qpay0, E * c.queue_cap) DA(h->P.qpay1, E * c.queue_cap) DA(h->P.qcache, E * h->P.n_qgroups)
  DA(h->P.agents, E * c.n_agents) DA(h->P.lv_price, E * 2 * c.level_cap) DA(h->P.lv_qty, E * 2 * c.level_cap) DA(h->P.lv_ht, E * 2 * c.level_cap)
  DA(h->P.nodes, E * c.order_cap) DA(h->P.env, E) DA(h->P.trace, E * (size_t)c.trace_cap) DA(h->d_stats, E) DA(h->d_seeds, E)
  DA(h->P.envx, E) DA(h->P.idtab, E * h->P.n_ids) DA(h->P.idbook, E * h->P.n_ids) DA(h->P.lobs, E * LOB_CAP * 3)
  DA(h->d_ts, dh->ts.size()) DA(h->d_first, dh->first.size()) DA(h->d_rows, dh->rows.size()) DA(h->d_daytab, dh->day_tab.size())
  DA(h->d_daytab2, dh->day_tab2.size()) DA(h->d_xid, dh->xid.size() + 1) DA(h->d_xfirst, dh->xfirst.size() + 1)
  DA(h->d_iact, E) DA(h->d_obs, E * 8) DA(h->d_trans, E * 6) DA(h->d_rew, E) DA(h->d_done, E) DA(h->d_msizes, E * (cfg->n_momentum > 0 ? cfg->n_momentum : 1))
  if (cfg->n_twap > 0) { DA(h->d_sched, (size_t)cfg->n_twap * cfg->n_horizon) }
  h->P.n_snap = n_exec > 0 ? n_exec : 1; h->P.snap_depth = DQ_DEPTH; DA(h->P.snap, E * (size_t)h->P.n_snap * 2 * DQ_DEPTH)                      // getCurrentSpread(depth=500) copies
#undef DA
  if (h->d_sched) CUH(cudaMemset(h->d_sched, 0xFF, sizeof(int32_t) * (size_t)cfg->n_twap * cfg->n_horizon));     // -1: no per-bin schedule, the TWAP quantity
  CUH(cudaMemcpy(h->d_ts, dh->ts.data(), sizeof(int64_t) * dh->ts.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_first, dh->first.data(), sizeof(int32_t) * dh->first.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_rows, dh->rows.data(), sizeof(int4) * dh->rows.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_daytab, dh->day_tab.data(), sizeof(int4) * dh->day_tab.size(), cudaMemcpyHostToDevice));
  CUH(cudaMemcpy(h->d_daytab2, dh->day_tab2.data(), sizeof(int4) * dh->day_tab2.size(), cudaMemcpyHostToDevice));
  if (!dh->xid.empty()) { CUH(cudaMemcpy(h->d_xid, dh->xid.data(), sizeof(int32_t) * dh->xid.size(), cudaMemcpyHostToDevice)); CUH(cudaMemcpy(h->d_xfirst, dh->xfirst.data(), sizeof(int32_t) * dh->xfirst.size(), cudaMemcpyHostToDevice)); }
  CUH(cudaMemset(h->P.agents, 0, E * c.n_agents * sizeof(ZiAgent)));
  h->P.st_ts = h->d_ts; h->P.st_first = h->d_first; h->P.st_rows = h->d_rows; h->P.day_tab = h->d_daytab; h->P.day_tab2 = h->d_daytab2; h->P.st_xid = h->d_xid; h->P.st_xfirst = h->d_xfirst;
  if (smem_cta > 48 * 1024) {
    CUH(cudaFuncSetAttribute(abx_dq_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute(abx_dq_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
    CUH(cudaFuncSetAttribute(abx_dq_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
  }
  *out = h; return ABX_OK;
}
int32_t abx_dq_create(const abx_dq_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out) {
  int64_t off[2] = {0, n_rows}; return abx_dq_create_days(cfg, stream5, off, 1, n_envs, device, out);
}

int32_t abx_dq_set_schedule(abx_sim *h, int32_t k, const int32_t *qty, int32_t n) {
  if (!h || !h->is_env || h->P.c.population != 2 || k < 0 || k >= h->P.dq_n_twap || !qty || n < 1 || !h->d_sched) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device));
  std::vector<int32_t> row((size_t)h->P.n_h, -1);
  for (int i = 0; i < n && i < h->P.n_h; i++) row[i] = qty[i];
  CU(cudaMemcpy(h->d_sched + (size_t)k * h->P.n_h, row.data(), sizeof(int32_t) * row.size(), cudaMemcpyHostToDevice));
  h->P.sched = h->d_sched; return ABX_OK;
}
int32_t abx_dq_reset(abx_sim *h, const uint64_t *seeds, const int32_t *mom_sizes, void *stream) {
  if (!h || !h->is_dq) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  if (seeds) CU(cudaMemcpyAsync(h->d_seeds, seeds, sizeof(uint64_t) * h->n_envs, cudaMemcpyHostToDevice, st));
  if (mom_sizes && h->P.dq_n_mom > 0) CU(cudaMemcpyAsync(h->d_msizes, mom_sizes, sizeof(int32_t) * (size_t)h->n_envs * h->P.dq_n_mom, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(h->P.idtab, 0, sizeof(uint4) * (size_t)h->n_envs * h->P.n_ids, st));
  CU(cudaMemsetAsync(h->P.idbook, 0, sizeof(uint2) * (size_t)h->n_envs * h->P.n_ids, st));
  CU(cudaMemsetAsync(h->P.lobs, 0, sizeof(int4) * (size_t)h->n_envs * LOB_CAP * 3, st));
  h->have_seeds = seeds != nullptr; h->have_msizes = mom_sizes && h->P.dq_n_mom > 0;
  int32_t rc = dq_reset_launch(h, nullptr, RESET_ALL, 0, st); if (rc != ABX_OK) return rc;
  h->reset_done = true; return ABX_OK;
}

int32_t abx_dq_step(abx_sim *h, const int32_t *actions_dev, double *obs_dev, double *trans_dev, double *reward_dev, uint8_t *done_dev, void *stream) {
  if (!h || !h->is_dq || !obs_dev || !trans_dev || !done_dev) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  CU(cudaSetDevice(h->device));
  bool instr = h->P.c.trace_cap > 0 || h->P.c.hash_pops != 0;
  if (instr) abx_dq_step_kernel<true><<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, actions_dev, obs_dev, trans_dev, reward_dev, done_dev, h->smem_per_warp);
  else abx_dq_step_kernel<false><<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, (cudaStream_t)stream>>>(h->P, actions_dev, obs_dev, trans_dev, reward_dev, done_dev, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  if (h->auto_reset) return dq_reset_launch(h, nullptr, RESET_DONE, h->auto_reset == 2, (cudaStream_t)stream);    // finished episodes start over: the next step runs them to their first decision tick
  return ABX_OK;
}

int32_t abx_dq_step_host(abx_sim *h, const int32_t *actions, double *obs, double *trans, double *reward, uint8_t *done, void *stream) {
  if (!h || !h->is_dq || !obs || !trans || !done) return ABX_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream; size_t E = (size_t)h->n_envs;
  CU(cudaSetDevice(h->device));
  if (actions) CU(cudaMemcpyAsync(h->d_iact, actions, sizeof(int32_t) * E, cudaMemcpyHostToDevice, st));
  int32_t rc = abx_dq_step(h, actions ? h->d_iact : nullptr, h->d_obs, h->d_trans, h->d_rew, h->d_done, stream); if (rc != ABX_OK) return rc;
  CU(cudaMemcpyAsync(obs, h->d_obs, sizeof(double) * 8 * E, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(trans, h->d_trans, sizeof(double) * 6 * E, cudaMemcpyDeviceToHost, st));
  if (reward) CU(cudaMemcpyAsync(reward, h->d_rew, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(done, h->d_done, E, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ABX_OK;
}

int32_t abx_dq_holdings(abx_sim *h, int32_t env, int64_t *out, double *exec_out, void *stream) {
  if (!h || !h->is_dq || !out || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  int n = h->P.c.n_agents; std::vector<ZiAgent> tmp(n); EnvX x;
  CU(cudaMemcpyAsync(tmp.data(), h->P.agents + (size_t)env * n, sizeof(ZiAgent) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&x, h->P.envx + env, sizeof(EnvX), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  dq_holdings_rows(h->P, tmp.data(), x, out, exec_out);
  return ABX_OK;
}

// ---------------- Book surface ----------------
int32_t abx_book_create(int32_t stream_history, int32_t level_cap, int32_t order_cap, int32_t trace_cap, int32_t n_envs, int32_t device, abx_sim **out) {
  abx_env_config ec; env_config_default(&ec);
  ec.order_level = 0; ec.stream_history = stream_history; ec.queue_cap = 32; ec.level_cap = level_cap; ec.order_cap = order_cap; ec.trace_cap = trace_cap; ec.hash_pops = 0;
  if (!out || n_envs < 1 || env_config_validate(&ec) != ABX_OK) return ABX_ERR_ARG;
  int ndev = 0; cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || device < 0 || device >= ndev) { snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d not present (%d visible): %s", device, ndev, cudaGetErrorString(ce)); return ABX_ERR_CUDA; }
  CU(cudaSetDevice(device));
  abx_sim *h = new (std::nothrow) abx_sim(); if (!h) return ABX_ERR_ARG;
  memset(h, 0, sizeof(*h)); h->is_env = true; h->is_book = true; h->n_envs = n_envs; h->device = device;
  h->st = new (std::nothrow) EnvStreamHost(); h->book_ids = new (std::nothrow) std::unordered_map<int64_t, int32_t>();
  if (!h->st || !h->book_ids) { abx_sim_destroy(h); return ABX_ERR_ARG; }
  env_fill_params(ec, h->P); h->P.n_envs = n_envs; h->P.c.n_agents = 2;
  h->smem_per_warp = (warp_smem_bytes(h->P.c, true) + 15) & ~(size_t)15;
  size_t smem_cta = h->smem_per_warp * ABX_WARPS_PER_CTA;
  if (smem_cta > 227 * 1024) { abx_sim_destroy(h); return ABX_ERR_ARG; }
  const abx_sim_config &c = h->P.c; size_t E = (size_t)n_envs; int stt;
#define DA(ptr, n) if ((stt = dalloc(&(ptr), (n), &h->bytes)) != ABX_OK) { abx_sim_destroy(h); return stt; }
  DA(h->P.qkey, E * c.queue_cap) DA(h->P.qpay0, E * c.queue_cap) DA(h->P.qpay1, E * c.queue_cap) DA(h->P.qcache, E * h->P.n_qgroups)
  DA(h->P.lv_price, E * 2 * c.level_cap) DA(h->P.lv_qty, E * 2 * c.level_cap) DA(h->P.lv_ht, E * 2 * c.level_cap)
  DA(h->P.nodes, E * c.order_cap) DA(h->P.env, E) DA(h->P.trace, E * (size_t)c.trace_cap) DA(h->d_stats, E) DA(h->P.envx, E)
#undef DA
  if (smem_cta > 48 * 1024) CUH(cudaFuncSetAttribute(abx_book_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
  *out = h; return ABX_OK;
}

int32_t abx_book_replay(abx_sim *h, const int64_t *ops9, int64_t n_ops, void *stream) {
  if (!h || !h->is_book || !ops9 || n_ops < 1) return ABX_ERR_ARG;
  CU(cudaSetDevice(h->device)); cudaStream_t st = (cudaStream_t)stream;
  std::vector<int64_t> dev_ops; int st_rc = book_ops_to_device(ops9, n_ops, *h->book_ids, h->st->id_orig, dev_ops);
  if (st_rc != ABX_OK) return st_rc;
  int n_ids = (int)h->st->id_orig.size(); bool fresh = !h->reset_done;
  if (n_ids > h->P.n_ids) {                                            // grow the per-order history table (one 16-byte record per distinct order id), keeping its contents
    uint4 *old = h->P.idtab; uint2 *oldb = h->P.idbook; int old_n = h->P.n_ids; int new_n = n_ids + n_ids / 2 + 64; uint4 *nw = nullptr; uint2 *nb = nullptr;
    CU(cudaMalloc((void **)&nw, sizeof(uint4) * (size_t)h->n_envs * new_n)); CU(cudaMemsetAsync(nw, 0, sizeof(uint4) * (size_t)h->n_envs * new_n, st));
    CU(cudaMalloc((void **)&nb, sizeof(uint2) * (size_t)h->n_envs * new_n)); CU(cudaMemsetAsync(nb, 0, sizeof(uint2) * (size_t)h->n_envs * new_n, st));
    if (old && !fresh) {
      CU(cudaMemcpy2DAsync(nw, sizeof(uint4) * new_n, old, sizeof(uint4) * old_n, sizeof(uint4) * old_n, h->n_envs, cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpy2DAsync(nb, sizeof(uint2) * new_n, oldb, sizeof(uint2) * old_n, sizeof(uint2) * old_n, h->n_envs, cudaMemcpyDeviceToDevice, st));
    }
    CU(cudaStreamSynchronize(st)); if (old) cudaFree(old); if (oldb) cudaFree(oldb);
    h->P.idtab = nw; h->P.idbook = nb; h->P.n_ids = new_n;
  }
  if (n_ops > h->ops_cap) { if (h->d_ops) { CU(cudaStreamSynchronize(st)); cudaFree(h->d_ops); h->d_ops = nullptr; } CU(cudaMalloc((void **)&h->d_ops, sizeof(int64_t) * 9 * (size_t)n_ops)); h->ops_cap = n_ops; }
  CU(cudaMemcpyAsync(h->d_ops, dev_ops.data(), sizeof(int64_t) * 9 * (size_t)n_ops, cudaMemcpyHostToDevice, st));
  abx_book_replay_kernel<<<grid_for(h->n_envs), 32 * ABX_WARPS_PER_CTA, h->smem_per_warp * ABX_WARPS_PER_CTA, st>>>(h->P, h->d_ops, n_ops, fresh ? 1 : 0, h->smem_per_warp);
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));                                       // dev_ops is a local staging buffer
  h->reset_done = true; return ABX_OK;
}

}  // extern "C"
