// abx_warp.cuh -- warp-cooperative context for abx::Sim (one 32-lane warp == one environment), sm_100a.
//
// On-chip (shared memory, per warp):   staged trader record (192 B, moved with 128-bit loads/stores),
//                                      outbox, event-queue group cache, both price ladders.
// HBM (per environment):               event slots (SoA: 16 B key + 2 x 16 B payload), trader records, order nodes.
//
// Event queue = unordered slots in groups of 32 + a cached (min key, occupancy mask) per group:
//   pop  : REDUX-min over the group cache (on chip) -> ONE coalesced 3 x 512 B load of the winning group -> REDUX-min
//   push : ballot for a group with a free slot -> three 16 B stores -> cache update
// so a pop costs one dependent HBM round trip instead of the ~11 of a binary heap with 2 000 entries, and the
// (time, recipient, type, uniq) order of the reference's heapq (Kernel.py:192,425; message/Message.py:39-45)
// is reproduced exactly because the minimum is recomputed from full keys.
//
// Price ladders = two sorted arrays (best level LAST) of {price, total qty, FIFO head|tail} searched with one
// pass of 32 lanes + ballot/REDUX, shifted in 32-wide chunks on insert/remove; FIFO order within a level is a
// linked list of 16 B order nodes in HBM (util/OrderBook.py:24-25 bids/asks[level][fifo]).
#pragma once
#include "abx_core.cuh"

namespace abx {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint4 ldcg4(const uint4 *p) { return __ldcg(p); }

// lane index of the smallest (hi, uniq) among lanes with valid == true, or -1
__device__ __forceinline__ int warp_argmin(uint64_t hi, uint32_t uniq, bool valid) {
  uint32_t h = valid ? (uint32_t)(hi >> 32) : 0xffffffffu;
  uint32_t m = __reduce_min_sync(FULL, h);
  bool c = valid && h == m;
  uint32_t l = c ? (uint32_t)hi : 0xffffffffu;
  uint32_t m2 = __reduce_min_sync(FULL, l);
  c = c && l == m2;
  uint32_t u = c ? uniq : 0xffffffffu;
  uint32_t m3 = __reduce_min_sync(FULL, u);
  c = c && u == m3;
  uint32_t b = __ballot_sync(FULL, c);
  return b ? __ffs(b) - 1 : -1;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint4 shfl4(uint4 v, int src) {
  uint4 r; r.x = __shfl_sync(FULL, v.x, src); r.y = __shfl_sync(FULL, v.y, src); r.z = __shfl_sync(FULL, v.z, src); r.w = __shfl_sync(FULL, v.w, src); return r;
}

// bytes of shared memory one environment needs
constexpr int SMALLQ_CAP = 64;          // queue capacities up to this keep every key on chip (WarpCtxT<true>)
constexpr int NEAR_OUT_CAP = 8;
constexpr int NEARQ_CAP = 32;           // QMODE 3: on-chip tier of the events due soon
constexpr int64_t NEAR_NS = 4000000000LL;   // QMODE 3: an event due within 4 simulated seconds is 'near' (messages travel 1 s + latency; wakeups are ~1000 s away)
__host__ __device__ inline bool warp_small_queue(const abx_sim_config &c) { return c.queue_cap <= SMALLQ_CAP; }
constexpr int OC_N = 16;                 // entries of the on-chip cache of replayed orders' records
__host__ __device__ inline size_t warp_smem_bytes(const abx_sim_config &c, bool env_shape = false, bool small_queue = false, bool hybrid_queue = false, bool near_queue = false) {
  size_t q = near_queue ? (size_t)NEARQ_CAP * 48 + (size_t)(c.queue_cap / 32) * 16 + 16 : hybrid_queue ? (size_t)SMALLQ_CAP * 48 + (size_t)(c.queue_cap / 32) * 16 : (small_queue ? (size_t)SMALLQ_CAP * 48 : (size_t)(c.queue_cap / 32) * 16);   // on-chip tier: key + 32-byte payload per slot
  return sizeof(ZiAgent) + (near_queue ? NEAR_OUT_CAP : OUT_CAP) * OUT_WORDS * 4 + q + (size_t)c.level_cap * 2 * 12 + (env_shape ? sizeof(EnvX) + OC_N * 32 : 0);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { uint64_t b = dbl_bits(v); uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)b, o), hi = __shfl_xor_sync(FULL, (uint32_t)(b >> 32), o); v += bits_dbl((uint64_t)lo | ((uint64_t)hi << 32)); }
  return v;
}

// SMALLQ: event queues of at most SMALLQ_CAP entries (ABIDESEnv shape: <= 40 pending events) keep all keys {key.lo, key.hi, uniq, kind|sender} in
// shared memory: a pop is two conflict-free key reads per lane + one warp arg-min + one 32-byte payload fetch, a push one key store + the payload;
// no group cache, no second arg-min over a fetched group, no recomputation after a removal.  Larger queues use the grouped layout (header comment).
// QMODE 0: grouped queue (header comment).  QMODE 1 (SMALLQ): all keys on chip.  QMODE 2: both -- the first SMALLQ_CAP events live on chip and only an
// QMODE 3 (sparse_zi): like 2 with NEARQ_CAP on-chip slots, but only events due within NEAR_NS go on chip (the messages in flight; the ~1000 pending
// wakeups stay in the grouped tier) and the minimum of the grouped tier is cached, so a message pop or push touches no HBM and scans no group cache.
// overflow (the DDQN config's closing market order: up to 500 orders in flight at once, one tick in 660) goes to the grouped structure, whose groups
// 0 and 1 are left unused because their HBM slots hold the on-chip tier's payloads; a pop takes the smaller of the two tiers' minima.
template <int QMODE>
struct WarpCtxT {
  static constexpr bool SMALLQ = QMODE >= 1, HYBRID = QMODE >= 2, NEARQ = QMODE == 3; static constexpr int NQ = NEARQ ? NEARQ_CAP : SMALLQ_CAP, G0 = HYBRID ? NQ / 32 : 0;
  static constexpr int OUTN = NEARQ ? NEAR_OUT_CAP : OUT_CAP;             // outbox entries (sparse_zi handlers emit at most three messages)
  const SimParams &P; int env, lane;
  // HBM bases of this environment
  uint4 *qkey, *qpay0, *qpay1; ZiAgent *agents; uint4 *nodes; abx_trace_rec *tr;
  // shared memory of this warp
  ZiAgent *staged; uint32_t *obox; uint4 *qc, *qs, *qp0, *qp1, *ovm; int32_t *lvp, *lvq; uint32_t *lvht; EnvX *ex;   // qc: group cache, qs: on-chip keys
  uint4 *oc_t; uint2 *oc_b; int32_t *oc_tag;   // on-chip cache of replayed orders' records (direct mapped, write-through)
  uint4 *idt; int4 *lob; uint2 *idb;   // ABIDESEnv shape: replay agent's per-order table, stored LOBs, per-order book census (HBM)
  int2 *snp;                           // deep QUERY_SPREAD replies: the execution agents' book snapshots (HBM)
  uint4 *hl;                           // population 3: the exchange's order-history log (HBM ring)
  // registers describing the group fetched by q_fetch
  uint64_t my_hi; uint32_t my_uniq; uint32_t cur_mask; int cur_group, cur_lane; int n_ovf; bool cur_t2; int day;   // day: the replayed day of this environment   // n_ovf: events in the overflow tier

  __device__ WarpCtxT(const SimParams &P_, int env_, unsigned char *smem) : P(P_), env(env_), lane(threadIdx.x & 31) {
    // unsigned 32 x 32 -> 64 products: the compiler recomputes these bases inside the loop body rather than hold them in registers, and the signed form is three instructions longer each time
    size_t q = (size_t)((uint64_t)(uint32_t)env * (uint32_t)P.c.queue_cap); qkey = P.qkey + q; qpay0 = P.qpay0 + q; qpay1 = P.qpay1 + q;
    agents = P.agents + (size_t)((uint64_t)(uint32_t)env * (uint32_t)P.c.n_agents); nodes = P.nodes + (size_t)((uint64_t)(uint32_t)env * (uint32_t)P.c.order_cap);
    tr = P.trace ? P.trace + (size_t)env * P.c.trace_cap : nullptr;
    staged = reinterpret_cast<ZiAgent *>(smem); smem += sizeof(ZiAgent);
    obox = reinterpret_cast<uint32_t *>(smem); smem += OUTN * OUT_WORDS * 4;
    qs = reinterpret_cast<uint4 *>(smem); smem += SMALLQ ? (size_t)NQ * 16 : 0;
    qp0 = reinterpret_cast<uint4 *>(smem); smem += SMALLQ ? (size_t)NQ * 16 : 0;
    qp1 = reinterpret_cast<uint4 *>(smem); smem += SMALLQ ? (size_t)NQ * 16 : 0;
    ovm = reinterpret_cast<uint4 *>(smem); smem += NEARQ ? 16 : 0;                   // cached minimum of the grouped tier {key lo, hi, uniq, group | ~0 = unknown}
    qc = reinterpret_cast<uint4 *>(smem); smem += (SMALLQ && !HYBRID) ? 0 : (size_t)P.n_qgroups * 16;
    lvp = reinterpret_cast<int32_t *>(smem); smem += (size_t)P.c.level_cap * 2 * 4;
    lvq = reinterpret_cast<int32_t *>(smem); smem += (size_t)P.c.level_cap * 2 * 4;
    lvht = reinterpret_cast<uint32_t *>(smem); smem += (size_t)P.c.level_cap * 2 * 4;
    ex = reinterpret_cast<EnvX *>(smem); smem += sizeof(EnvX);
    oc_t = reinterpret_cast<uint4 *>(smem); oc_b = reinterpret_cast<uint2 *>(smem + OC_N * 16); oc_tag = reinterpret_cast<int32_t *>(smem + OC_N * 24);
    if (P.idbook && lane < OC_N) oc_tag[lane] = -1;
    idb = P.idbook ? P.idbook + (size_t)env * P.n_ids : nullptr;
    snp = P.snap ? P.snap + (size_t)env * P.n_snap * 2 * P.snap_depth : nullptr;
    idt = P.idtab ? P.idtab + (size_t)env * P.n_ids : nullptr; lob = P.lobs ? P.lobs + (size_t)env * lob_stride_of(P.c) : nullptr;
    hl = P.hlog ? P.hlog + (size_t)env * hist_stride_of(P.c) : nullptr;
    cur_group = cur_lane = -1; cur_mask = 0; my_hi = KEY_EMPTY; my_uniq = 0xffffffffu; n_ovf = 0; cur_t2 = false;
    day = P.n_days > 1 ? env % P.n_days : 0;                               // once per launch: an integer modulo is ~170 instructions
  }
  // environment e replays day (e + episode) % n_days: every day-advancing reset moves it on to its next day
  __device__ __forceinline__ void set_episode(uint32_t episode) { if (P.n_days > 1 && episode) day = (int)(((uint32_t)env + episode) % (uint32_t)P.n_days); }
  // per-order tables of one environment back to zero (masked resets; a whole-batch reset uses cudaMemsetAsync instead)
  __device__ void clear_tables() {
    uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    if (idt) for (int i = lane; i < P.n_ids; i += 32) __stcg(idt + i, z4);
    if (idb) for (int i = lane; i < P.n_ids; i += 32) __stcg(idb + i, make_uint2(0u, 0u));
    if (lob) for (int i = lane; i < LOB_CAP * 3; i += 32) __stcg(lob + i, make_int4(0, 0, 0, 0));
    if (P.idbook && lane < OC_N) oc_tag[lane] = -1;
    __syncwarp();
  }
  // Uniform code stores on-chip state from every lane (same value, same address: one STS, no branch).
  __device__ __forceinline__ bool onchip_writer() const { return true; }
  // sync(): between UNIFORM stores (every lane writes the same value to the same address) and later reads a COMPILER fence is
  // enough -- each lane reads back what it stored itself.  Real barriers (xsync(), __syncwarp()) stand wherever lanes hand
  // DIFFERENT data to each other: ladder shifts, the staging copies, and everything that goes through HBM (q_push, node_store,
  // agent_commit).  -DABX_STRICT_SYNC turns every sync() into __syncwarp() as well (the A/B build of tools/ab_sync_variants.sh;
  // the GPU parity suite runs on both).
#ifdef ABX_STRICT_SYNC
  __device__ __forceinline__ void sync() const { __syncwarp(); }
#else
  __device__ __forceinline__ void sync() const { asm volatile("" ::: "memory"); }
#endif
  // xsync(): ALWAYS a real warp barrier.  Used wherever one lane's shared-memory store is read by a DIFFERENT lane that did not
  // store the same value itself (ladder shifts, staging copies where lane i moves element i): a compiler fence is not enough there
  // under independent thread scheduling.  Uniform code (every lane stores the same value to the same address, then reads it back)
  // only needs sync(): a thread always observes its own store.
#ifdef ABX_FENCE_ONLY                     // A/B build only (tools/): the round-1 behaviour, never the shipped library
  __device__ __forceinline__ void xsync() const { asm volatile("" ::: "memory"); }
#else
  __device__ __forceinline__ void xsync() const { __syncwarp(); }
#endif
  __device__ __forceinline__ uint32_t *outbox() const { return obox; }
  __device__ __forceinline__ void trace(const abx_trace_rec &r, uint32_t i) { if (lane == 0) tr[i] = r; }

  // ---- staging of the on-chip structures between launches ----
  __device__ void load_onchip(const EnvState &s) {
    if (SMALLQ) {                                                            // keys and payloads of the on-chip tier (HBM only between launches)
#pragma unroll
      for (int j = 0; j < NQ / 32; j++) { int i = lane + 32 * j; uint4 k = ldcg4(qkey + i); qs[i] = k;
        if ((k.x & k.y) != 0xffffffffu) { qp0[i] = ldcg4(qpay0 + i); qp1[i] = ldcg4(qpay1 + i); } }
      if (NEARQ) ovm[0] = make_uint4(0u, 0u, 0u, 0xffffffffu);
    }
    if (!SMALLQ || HYBRID) {
      const uint4 *gc = P.qcache + (size_t)env * P.n_qgroups; int cnt = 0;
#pragma unroll 1
      for (int g = lane; g < P.n_qgroups; g += 32) { uint4 v = ldcg4(gc + g); qc[g] = v; if (HYBRID && g >= G0) cnt += __popc(v.w); }
      if (HYBRID) n_ovf = __reduce_add_sync(FULL, cnt);
    }
    size_t l = (size_t)env * 2 * P.c.level_cap;
#pragma unroll 1
    for (int side = 0; side < 2; side++)
#pragma unroll 1
      for (int i = lane; i < (side ? s.n_ask_lv : s.n_bid_lv); i += 32) {
      int k = side * P.c.level_cap + i; lvp[k] = P.lv_price[l + k]; lvq[k] = P.lv_qty[l + k]; lvht[k] = P.lv_ht[l + k];
    }
    xsync();                                                                 // lane i staged element i; every lane reads all of them from here on
  }
  __device__ void store_onchip(const EnvState &s) {
    xsync();
    if (SMALLQ) {
#pragma unroll
      for (int j = 0; j < NQ / 32; j++) { int i = lane + 32 * j; uint4 k = qs[i]; __stcg(qkey + i, k);
        if ((k.x & k.y) != 0xffffffffu) { __stcg(qpay0 + i, qp0[i]); __stcg(qpay1 + i, qp1[i]); } }
    }
    if (!SMALLQ || HYBRID) {
      uint4 *gc = P.qcache + (size_t)env * P.n_qgroups;
#pragma unroll 1
      for (int g = lane; g < P.n_qgroups; g += 32) gc[g] = qc[g];
    }
    size_t l = (size_t)env * 2 * P.c.level_cap;
#pragma unroll 1
    for (int side = 0; side < 2; side++)
#pragma unroll 1
      for (int i = lane; i < (side ? s.n_ask_lv : s.n_bid_lv); i += 32) {
      int k = side * P.c.level_cap + i; P.lv_price[l + k] = lvp[k]; P.lv_qty[l + k] = lvq[k]; P.lv_ht[l + k] = lvht[k];
    }
  }
  __device__ void q_clear() {
    if (SMALLQ) { qs[lane] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u); if (NQ > 32) qs[lane + 32] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u); n_ovf = 0;
      if (NEARQ) ovm[0] = make_uint4(0u, 0u, 0u, 0xffffffffu); }
    if (!SMALLQ || HYBRID) for (int g = lane; g < P.n_qgroups; g += 32) qc[g] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
    xsync(); }

  // ---- event queue ----
  // grp identifies the winner for q_fetch: a slot of the on-chip tier (SMALLQ: 0 .. SMALLQ_CAP-1), otherwise SMALLQ_CAP + group index
  __device__ __forceinline__ bool q_min(uint64_t &hi, uint32_t &uniq, int &grp) {
    bool have1 = false;
    if (SMALLQ) {                                                          // empty slots hold KEY_EMPTY
      uint4 k0 = qs[lane]; uint64_t h = (uint64_t)k0.x | ((uint64_t)k0.y << 32); uint32_t u = k0.z; bool second = false;
      if (NQ > 32) { uint4 k1 = qs[lane + 32]; uint64_t h1 = (uint64_t)k1.x | ((uint64_t)k1.y << 32); second = key_less(h1, k1.z, h, u); if (second) { h = h1; u = k1.z; } }
      int wl = warp_argmin(h, u, h != KEY_EMPTY);
      if (wl >= 0) { have1 = true; grp = __shfl_sync(FULL, second ? lane + 32 : lane, wl); hi = shfl64(h, wl); uniq = __shfl_sync(FULL, u, wl); }
      if (!HYBRID || n_ovf == 0) return have1;
      if (NEARQ) {                                                         // the grouped tier's minimum is cached until that tier loses an event
        uint4 m = ovm[0];
        if (m.w != 0xffffffffu) {
          uint64_t h2 = (uint64_t)m.x | ((uint64_t)m.y << 32);
          if (!have1 || key_less(h2, m.z, hi, uniq)) { grp = NQ + (int)m.w; hi = h2; uniq = m.z; }
          return true;
        }
      }
    }
    uint64_t bh = KEY_EMPTY; uint32_t bu = 0xffffffffu; int bg = -1;
    for (int g = G0 + lane; g < P.n_qgroups; g += 32) {
      uint4 cc = qc[g];
      if (cc.w) { uint64_t h = (uint64_t)cc.x | ((uint64_t)cc.y << 32); if (bg < 0 || key_less(h, cc.z, bh, bu)) { bh = h; bu = cc.z; bg = g; } }
    }
    int wl = warp_argmin(bh, bu, bg >= 0);
    if (wl < 0) return have1;
    int g2 = __shfl_sync(FULL, bg, wl); uint64_t h2 = shfl64(bh, wl); uint32_t u2 = __shfl_sync(FULL, bu, wl);
    if (NEARQ) { sync(); ovm[0] = make_uint4((uint32_t)h2, (uint32_t)(h2 >> 32), u2, (uint32_t)g2); sync(); }
    if (!have1 || key_less(h2, u2, hi, uniq)) { grp = (SMALLQ ? NQ : 0) + g2; hi = h2; uniq = u2; }
    return true;
  }
  __device__ __forceinline__ void q_fetch(int g, Event &e) {
    uint4 k, a, b;                                                 // one unpack for both tiers (code size: the loop body lives on the edge of the instruction cache)
    if (SMALLQ && g < NQ) {                                        // a slot of the on-chip tier: key and payload in shared memory (no L2 round trip per pop)
      cur_group = g; cur_t2 = false;
      k = qs[g]; a = qp0[g]; b = qp1[g];
    } else {
      if (SMALLQ) { g -= NQ; cur_t2 = true; }
      int slot = g * 32 + lane;
      k = ldcg4(qkey + slot); a = ldcg4(qpay0 + slot); b = ldcg4(qpay1 + slot);       // 3 x 512 B coalesced
      cur_mask = qc[g].w; cur_group = g;
      bool occ = (cur_mask >> lane) & 1u;
      my_hi = (uint64_t)k.x | ((uint64_t)k.y << 32); my_uniq = k.z;
      int w = warp_argmin(my_hi, my_uniq, occ);
      cur_lane = w;
      k = shfl4(k, w); a = shfl4(a, w); b = shfl4(b, w);
    }
    event_unpack(k, a, b, e);
  }
  __device__ __forceinline__ void group_writeback() {     // recompute the cached minimum of cur_group from the keys in registers
    bool occ = (cur_mask >> lane) & 1u;
    int w2 = warp_argmin(my_hi, my_uniq, occ);
    uint64_t nh = KEY_EMPTY; uint32_t nu = 0xffffffffu;
    if (w2 >= 0) { nh = shfl64(my_hi, w2); nu = __shfl_sync(FULL, my_uniq, w2); }
    qc[cur_group] = make_uint4((uint32_t)nh, (uint32_t)(nh >> 32), nu, cur_mask);
    sync();
  }
  // The fetched event either leaves the queue or goes back in at time t (Kernel.py:226,260: same entry, new time, same uniq).  ONE body for both, so
  // that the grouped tier's cache recomputation (an arg-min over the fetched group) exists once per loop.
  __device__ __forceinline__ void q_settle(bool requeue, int64_t t) {
    if (SMALLQ && !cur_t2) {
      uint4 k = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
      if (requeue) { k = qs[cur_group]; uint64_t h = (uint64_t)k.x | ((uint64_t)k.y << 32);
        h = key_pack(t < KEY_T_MAX ? t : KEY_T_MAX, key_recipient(h), key_type(h)); k.x = (uint32_t)h; k.y = (uint32_t)(h >> 32); }
      sync(); qs[cur_group] = k; sync(); return;
    }
    if (requeue) {
      if (lane == cur_lane) {
        my_hi = key_pack(t < KEY_T_MAX ? t : KEY_T_MAX, key_recipient(my_hi), key_type(my_hi));
        uint2 *kp = reinterpret_cast<uint2 *>(qkey + cur_group * 32 + lane); *kp = make_uint2((uint32_t)my_hi, (uint32_t)(my_hi >> 32));
      }
    } else { if (HYBRID) n_ovf--; cur_mask &= ~(1u << cur_lane); }
    if (NEARQ) ovm[0] = make_uint4(0u, 0u, 0u, 0xffffffffu);
    group_writeback();
  }
  __device__ __forceinline__ void q_remove() { q_settle(false, 0); }
  __device__ __forceinline__ void q_requeue(int64_t t) { q_settle(true, t); }
  __device__ __forceinline__ bool q_push(const Event &e, int64_t now) {
    uint4 k, a, b; event_pack(e, k, a, b);                                 // packed once for both tiers
    if (SMALLQ && (!NEARQ || e.t - now < NEAR_NS)) {
      bool f0 = qs[lane].y == 0xffffffffu && qs[lane].x == 0xffffffffu, f1 = NQ > 32 && qs[NQ > 32 ? lane + 32 : lane].y == 0xffffffffu && qs[NQ > 32 ? lane + 32 : lane].x == 0xffffffffu;
      uint32_t b0 = __ballot_sync(FULL, f0), b1 = NQ > 32 ? __ballot_sync(FULL, f1) : 0u;
      if (b0 | b1) {
        int slot = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
        sync(); qs[slot] = k; qp0[slot] = a; qp1[slot] = b; sync();
        return true;
      }
      if (!HYBRID) return false;
    }
    int fg = -1;
    for (int g = G0 + lane; g < P.n_qgroups; g += 32) if (qc[g].w != 0xffffffffu) { fg = g; break; }
    uint32_t bal = __ballot_sync(FULL, fg >= 0);
    if (!bal) return false;
    int g = __shfl_sync(FULL, fg, __ffs(bal) - 1);
    uint4 cc = qc[g]; int i = __ffs(~cc.w) - 1;
    int slot = g * 32 + i;
    if (lane < 3) { uint4 *dst = lane == 0 ? qkey : (lane == 1 ? qpay0 : qpay1); dst[slot] = lane == 0 ? k : (lane == 1 ? a : b); }
    {
      uint64_t h = (uint64_t)k.x | ((uint64_t)k.y << 32), ch = (uint64_t)cc.x | ((uint64_t)cc.y << 32);
      if (cc.w == 0 || key_less(h, k.z, ch, cc.z)) { cc.x = k.x; cc.y = k.y; cc.z = k.z; }
      cc.w |= 1u << i; qc[g] = cc;
      if (NEARQ) { uint4 m = ovm[0]; if (m.w != 0xffffffffu && (n_ovf == 0 || key_less(h, k.z, (uint64_t)m.x | ((uint64_t)m.y << 32), m.z))) ovm[0] = make_uint4(k.x, k.y, k.z, (uint32_t)g); }
    }
    if (HYBRID) n_ovf++;
    __syncwarp();
    return true;
  }

  // ---- ladders (side 0 bids ascending, side 1 asks descending: best level last) ----
  __device__ __forceinline__ int32_t lv_price(int side, int i) const { return lvp[side * P.c.level_cap + i]; }
  __device__ __forceinline__ int32_t lv_qty(int side, int i) const { return lvq[side * P.c.level_cap + i]; }
  __device__ __forceinline__ uint32_t lv_head(int side, int i) const { return lvht[side * P.c.level_cap + i] & 0xffffu; }
  __device__ __forceinline__ uint32_t lv_tail(int side, int i) const { return lvht[side * P.c.level_cap + i] >> 16; }
  __device__ __forceinline__ void lv_set(int side, int i, int32_t qty, uint32_t head, uint32_t tail) {
    sync();
    lvq[side * P.c.level_cap + i] = qty; lvht[side * P.c.level_cap + i] = head | (tail << 16);
    sync();
  }
  // enterOrder's scan (util/OrderBook.py:271-282) on the best-last arrays: the level nearest the best whose price the order beats or equals.
  // found: it equals (the order joins that level); otherwise pos is where the new level goes.  On a sorted ladder this is the sorted insert.
  __device__ __forceinline__ void lv_find(int side, int32_t price, int n, int &pos, bool &found) {
    const int32_t *p = lvp + side * P.c.level_cap; int best = -1;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) { int32_t v = p[i]; if (side == 0 ? v <= price : v >= price) best = i; }
    best = __reduce_max_sync(FULL, best);
    if (n > 0 && (side == 0 ? p[0] > price : p[0] < price)) best = -1;        // :267-270 tested first: worse than the LAST level -> new last level, whatever stands before it
    found = best >= 0 && p[best] == price; pos = found ? best : best + 1;
  }
  // cancelOrder / modifyOrder (:306, :349): the level nearest the best, below index `limit`, whose price EQUALS `price` (-1: none);
  // cnt = how many such levels there are (1 on a sorted ladder)
  __device__ __forceinline__ int lv_find_eq(int side, int32_t price, int limit, int &cnt) {
    const int32_t *p = lvp + side * P.c.level_cap; int best = -1, k = 0;
#pragma unroll 1
    for (int i = lane; i < limit; i += 32) if (p[i] == price) { best = i; k++; }
    cnt = __reduce_add_sync(FULL, k);
    return __reduce_max_sync(FULL, best);
  }
  __device__ __forceinline__ void lv_setp(int side, int i, int32_t price) { sync(); lvp[side * P.c.level_cap + i] = price; sync(); }
  __device__ __forceinline__ void lv_insert(int side, int pos, int n, int32_t price, int32_t qty, uint32_t head, uint32_t tail) {
    int b = side * P.c.level_cap;
    sync();
    for (int hi = n; hi > pos; hi -= 32) {                  // shift [pos, n) up by one, top chunk first
      int i = hi - 1 - lane; bool act = i >= pos; int32_t x = 0, y = 0; uint32_t z = 0;
      if (act) { x = lvp[b + i]; y = lvq[b + i]; z = lvht[b + i]; }
      xsync();                                                               // lane L overwrites the slot lane L-1 has just read
      if (act) { lvp[b + i + 1] = x; lvq[b + i + 1] = y; lvht[b + i + 1] = z; }
      xsync();
    }
    lvp[b + pos] = price; lvq[b + pos] = qty; lvht[b + pos] = head | (tail << 16);
    sync();
  }
  __device__ __forceinline__ void lv_remove(int side, int pos, int n) {
    int b = side * P.c.level_cap;
    sync();
    for (int lo = pos + 1; lo < n; lo += 32) {              // shift (pos, n) down by one, bottom chunk first
      int i = lo + lane; bool act = i < n; int32_t x = 0, y = 0; uint32_t z = 0;
      if (act) { x = lvp[b + i]; y = lvq[b + i]; z = lvht[b + i]; }
      xsync();
      if (act) { lvp[b + i - 1] = x; lvq[b + i - 1] = y; lvht[b + i - 1] = z; }
      xsync();
    }
  }

  // getInsideBids(depth) / getInsideAsks(depth) of a deep query, copied when the exchange processes it: the first n_copy of the side's n_total levels,
  // best first, into execution agent k's area (32 levels per pass; a real barrier before anyone reads them back)
  __device__ __forceinline__ void snap_store(int k, int side, int n_total, int n_copy) {
    int2 *dst = snp + (size_t)(k * 2 + side) * P.snap_depth; int b = side * P.c.level_cap;
#pragma unroll 1
    for (int i = lane; i < n_copy; i += 32) __stcg(dst + i, make_int2(lvp[b + n_total - 1 - i], lvq[b + n_total - 1 - i]));
    __syncwarp();
  }
  __device__ __forceinline__ int2 snap_load(int k, int side, int i) const { return __ldcg(snp + (size_t)(k * 2 + side) * P.snap_depth + i); }
  // ---- order nodes (HBM, 16 B each) ----
  template <bool PRICED> __device__ __forceinline__ NodeRec node_load(uint32_t i) const { return node_unpack<PRICED>(ldcg4(nodes + i)); }
  template <bool PRICED> __device__ __forceinline__ void node_store(uint32_t i, const NodeRec &r) { if (lane == 0) __stcg(nodes + i, node_pack<PRICED>(r)); __syncwarp(); }

  // ---- ABIDESEnv shape ----
  __device__ __forceinline__ EnvX *envx() const { return ex; }
  __device__ void envx_load() { const uint4 *src = reinterpret_cast<const uint4 *>(P.envx + env); for (int i = lane; i < (int)(sizeof(EnvX) / 16); i += 32) reinterpret_cast<uint4 *>(ex)[i] = ldcg4(src + i); xsync(); }
  __device__ void envx_store() { xsync(); uint4 *dst = reinterpret_cast<uint4 *>(P.envx + env); for (int i = lane; i < (int)(sizeof(EnvX) / 16); i += 32) __stcg(dst + i, reinterpret_cast<const uint4 *>(ex)[i]); }
  __device__ __forceinline__ uint4 id_load(int i) const { return ldcg4(idt + i); }
  __device__ __forceinline__ void id_store(int i, uint4 v) { if (lane == 0) __stcg(idt + i, v); __syncwarp(); }
  // An agent's self.orders as a dense table of n 16-byte entries {order id, price, signed qty, -} at idt[base..]: index of `oid` (insertion order kept, so the
  // first match is the dict's entry) and deletion of entry f.  32 entries per pass, coalesced, instead of one dependent L2 round trip per entry.
  __device__ int tab_find(int base, int n, uint32_t oid) const {
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane; uint32_t x = i < n ? ldcg4(idt + base + i).x : 0u;
      uint32_t m = __ballot_sync(FULL, i < n && x == oid);
      if (m) return i0 + __ffs(m) - 1;
    }
    return -1;
  }
  __device__ void tab_remove(int base, int n, int f) {
#pragma unroll 1
    for (int i0 = f; i0 + 1 < n; i0 += 32) {                               // entries i0+1 .. i0+32 move down by one: all reads of a pass precede its writes
      int i = i0 + lane; bool on = i + 1 < n; uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (on) v = ldcg4(idt + base + i + 1);
      __syncwarp();
      if (on) __stcg(idt + base + i, v);
      __syncwarp();
    }
  }
  // Replayed orders' records {agent-side entry of idtab, census entry of idbook}: one replayed order is touched by three to five consecutive events
  // (placement, the exchange's registration and census, the acknowledgement), each a dependent L2/HBM round trip.  A miss fetches both halves at once;
  // stores write through, so HBM is always current and the cache needs no flush.  Same value, same address from every lane: one transaction.
  __device__ __forceinline__ int oc_slot(int i) {
    int sl = i & (OC_N - 1);
    if (oc_tag[sl] != i) { uint4 t = ldcg4(idt + i); uint2 b = __ldcg(idb + i); sync(); oc_tag[sl] = i; oc_t[sl] = t; oc_b[sl] = b; sync(); }
    return sl;
  }
  __device__ __forceinline__ uint4 ord_load(int i) { return oc_t[oc_slot(i)]; }
  __device__ __forceinline__ void ord_store(int i, uint4 v) { __stcg(idt + i, v); int sl = i & (OC_N - 1); sync(); if (oc_tag[sl] == i) oc_t[sl] = v; sync(); }
  __device__ __forceinline__ uint2 ib_load(int i) { return oc_b[oc_slot(i)]; }
  // the per-order records are cold HBM sectors (one per order, ~1 MB per environment): ask L2 for them an event or more before they are needed
  __device__ __forceinline__ void id_prefetch(int i) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(idt + i)); }
  __device__ __forceinline__ void ib_prefetch(int i) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(idb + i)); }
  __device__ __forceinline__ void ib_store(int i, uint2 v) { __stcg(idb + i, v); int sl = i & (OC_N - 1); sync(); if (oc_tag[sl] == i) oc_b[sl] = v; sync(); }
  __device__ __forceinline__ int4 row_load(int r) const { return __ldg(P.st_rows + r); }
  __device__ __forceinline__ int4 day_rec() const { return __ldg(P.day_tab + day); }
  __device__ __forceinline__ int n_ts() const { return day_rec().y; }
  __device__ __forceinline__ int64_t ts_load(int k) const { return __ldg(P.st_ts + day_rec().x + k); }
  __device__ __forceinline__ int first_load(int k) const { return __ldg(P.st_first + day_rec().z + k); }
  __device__ __forceinline__ int day_row0() const { return first_load(0); }               // rows of all days are stored back to back
  __device__ __forceinline__ int4 day_rec2() const { return __ldg(P.day_tab2 + day); }
  // first row (day-local) that carries explicit ORDER_ID `id`, or -1 when the day's stream never uses it: binary search over the sorted ids
  __device__ int xid_first_row(int4 d2, int32_t id) const {
    int lo = 0, hi = d2.y - 1;
#pragma unroll 1
    while (lo <= hi) { int mid = (lo + hi) >> 1; int32_t v = __ldg(P.st_xid + d2.x + mid); if (v == id) return __ldg(P.st_xfirst + d2.x + mid); if (v < id) lo = mid + 1; else hi = mid - 1; }
    return -1;
  }
  __device__ __forceinline__ void lob_store(int slot, const int32_t w[12]) {
    if (lane < 3) __stcg(lob + slot * 3 + lane, make_int4(w[lane * 4], w[lane * 4 + 1], w[lane * 4 + 2], w[lane * 4 + 3]));
    __syncwarp();
  }
  __device__ __forceinline__ void lob_load(int slot, int32_t w[12]) const {
#pragma unroll
    for (int k = 0; k < 3; k++) { int4 v = __ldcg(lob + slot * 3 + k); w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
  }
  // ---- population 3 (config/rmsc01.py): the exchange's order history as a ring of {order id, limit price, history epoch, is_buy | transacted << 1} ----
  __device__ __forceinline__ uint4 hist_load(int slot) const { return ldcg4(hl + slot); }
  __device__ __forceinline__ void hist_store(int slot, uint4 v) { if (lane == 0) __stcg(hl + slot, v); __syncwarp(); }
  // HeuristicBeliefLearningAgent.placeOrder :98-168 over the log entries of epochs e_lo .. e_hi (a contiguous run: epochs never decrease along the log).
  // The reference fills one row per price between the lowest and the highest of those orders with cumulative counts of successful / unsuccessful asks and
  // bids and takes the first argmax of Pr * surplus.  Pr is piecewise constant and only changes at p_i (counts "at or below") or p_i + 1 ("at or above"), and
  // within a piece Pr * surplus is strictly monotonic, so the argmax over all rows is the argmax over {low, high, p_i - 1, p_i, p_i + 1}.  Two forms:
  // while the price span fits the scratch rows (hist_log_cap / 4), the table is built as the reference builds it (histogram + prefix scan over the rows,
  // 32 rows per pass); otherwise lanes take candidate prices 32 at a time and count the run for each.  Integer counts, the reference's fp64 quotient and
  // product, ties to the lower price -- both forms are exact.
  __device__ bool hbl_best(uint32_t hist_n, uint32_t e_lo, uint32_t e_hi, bool buy, int32_t v, int32_t &best_p, uint32_t &err) const {
    uint32_t cap = (uint32_t)P.c.hist_log_cap, avail = hist_n < cap ? hist_n : cap, k_first = avail, k_end = avail; bool stop = false;
#pragma unroll 1
    for (uint32_t k0 = 0; k0 < avail && !stop; k0 += 32) {                  // k counts back from the newest entry
      uint32_t k = k0 + lane; bool on = k < avail; uint32_t ep = on ? ldcg4(hl + ((hist_n - 1 - k) & (cap - 1))).z : 0u;
      uint32_t inr = __ballot_sync(FULL, on && ep >= e_lo && ep <= e_hi), older = __ballot_sync(FULL, on && ep < e_lo);
      if (inr && k_first == avail) k_first = k0 + __ffs(inr) - 1;
      if (older) { k_end = k0 + __ffs(older) - 1; stop = true; }
    }
    if (!stop && hist_n > cap) err |= ABX_F_HISTORY_OVERFLOW;               // the ring no longer reaches back to bucket e_lo
    if (k_first >= k_end) { err |= ABX_F_REF_EXCEPTION; return false; }     // empty buckets: np.zeros((negative, 8)) raises in the reference (cannot happen: every bucket holds the order whose trade closed it)
    uint32_t N = k_end - k_first, top = hist_n - 1 - k_first;               // entries top, top - 1, .. top - N + 1
    int32_t lo = 0x7fffffff, hi = -0x7fffffff - 1;
#pragma unroll 1
    for (uint32_t j = lane; j < N; j += 32) { int32_t pr = (int32_t)ldcg4(hl + ((top - j) & (cap - 1))).y; lo = pr < lo ? pr : lo; hi = pr > hi ? pr : hi; }
    lo = __reduce_min_sync(FULL, lo); hi = __reduce_max_sync(FULL, hi);
    double bes = -1.0e300; int32_t bp = 0x7fffffff;
    uint32_t rows_cap = P.c.hbl_table_rows > 0 ? (uint32_t)P.c.hbl_table_rows : cap / 4; uint2 *scr = reinterpret_cast<uint2 *>(hl + cap);   // one {forward class, reverse class} counter pair per price row, behind the ring
    if ((int64_t)hi - (int64_t)lo < (int64_t)rows_cap) {
      // the reference's table itself (:112-165): histogram of the run over the price rows, one prefix scan, Pr * surplus per row -- O(N + rows)
      uint32_t R = (uint32_t)(hi - lo) + 1u, atot = 0, btot = 0;
#pragma unroll 1
      for (uint32_t r = lane; r < R; r += 32) __stcg(scr + r, make_uint2(0u, 0u));
      __syncwarp();
#pragma unroll 1
      for (uint32_t j0 = 0; j0 < N; j0 += 32) {
        uint32_t j = j0 + lane; bool on = j < N; uint4 r = on ? ldcg4(hl + ((top - j) & (cap - 1))) : make_uint4(0u, 0u, 0u, 0u);
        bool isb = r.w & 1u, tx = (r.w & 2u) != 0, fwd = on && (tx || (buy ? !isb : isb));   // counted towards num; every other order of the run only towards the denominator
        if (on) atomicAdd(fwd ? &scr[(int32_t)r.y - lo].x : &scr[(int32_t)r.y - lo].y, 1u);
        atot += __popc(__ballot_sync(FULL, fwd)); btot += __popc(__ballot_sync(FULL, on && !fwd));
      }
      __syncwarp();
      uint32_t ca = 0, cb = 0;                                              // counts in the rows below the current 32
#pragma unroll 1
      for (uint32_t r0 = 0; r0 < R; r0 += 32) {
        uint32_t r = r0 + lane; bool on = r < R; uint2 cnt = on ? __ldcg(scr + r) : make_uint2(0u, 0u); uint32_t ia = cnt.x, ib = cnt.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t ua = __shfl_up_sync(FULL, ia, o), ub = __shfl_up_sync(FULL, ib, o); if (lane >= o) { ia += ua; ib += ub; } }
        // buy: num = orders of the forward class at or below the row, the rest of the denominator = reverse class at or above it; sell: the mirror image
        uint32_t num = buy ? ca + ia : atot - (ca + ia - cnt.x), oth = buy ? btot - (cb + ib - cnt.y) : cb + ib, den = num + oth;
        int32_t p = lo + (int32_t)r;
        double pr = den == 0 ? 0.0 : (double)num / (double)den, es = pr * (double)(buy ? v - p : p - v);
        if (on && (es > bes || (es == bes && p < bp))) { bes = es; bp = p; }
        ca += __shfl_sync(FULL, ia, 31); cb += __shfl_sync(FULL, ib, 31);
      }
    } else {
    uint32_t ncand = 3 * N + 2;                                             // price span wider than the scratch rows: the candidate form (exact, O(N^2 / 32))
#pragma unroll 1
    for (uint32_t ci = lane; ci < ncand; ci += 32) {
      int32_t p;
      if (ci < 3 * N) { uint32_t j = ci / 3; p = (int32_t)ldcg4(hl + ((top - j) & (cap - 1))).y + (int32_t)(ci - 3 * j) - 1; } else p = ci == 3 * N ? lo : hi;
      if (p < lo || p > hi) continue;
      uint32_t num = 0, den = 0;
#pragma unroll 4
      for (uint32_t j = 0; j < N; j++) {
        uint4 r = ldcg4(hl + ((top - j) & (cap - 1))); int32_t q = (int32_t)r.y; bool isb = r.w & 1u, tx = (r.w & 2u) != 0, le = q <= p, ge = q >= p;
        bool in_num = buy ? (le && (tx || !isb)) : (ge && (tx || isb)), extra = buy ? (!tx && isb && ge) : (!tx && !isb && le);   // :115-150 columns 0-2 (buy) / 0, 1, 3 (sell) over column 3 / 2
        num += in_num; den += in_num || extra;
      }
      double pr = den == 0 ? 0.0 : (double)num / (double)den;               // :152-159 nan_to_num(0 / 0)
      double es = pr * (double)(buy ? v - p : p - v);                       // :162-165
      if (es > bes || (es == bes && p < bp)) { bes = es; bp = p; }
    }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      uint64_t b = dbl_bits(bes); uint32_t l = __shfl_xor_sync(FULL, (uint32_t)b, o), h = __shfl_xor_sync(FULL, (uint32_t)(b >> 32), o); int32_t op = __shfl_xor_sync(FULL, bp, o);
      double oes = bits_dbl((uint64_t)l | ((uint64_t)h << 32));
      if (oes > bes || (oes == bes && op < bp)) { bes = oes; bp = op; }
    }
    best_p = bp; return bes > 0.0;                                          // :173 only a positive expected surplus is worth an order
  }
  // rmsc03: momentum agent k's ring of doubled mid prices lives in the per-environment int4 table
  __device__ __forceinline__ int32_t mid_load(int k, int slot) const { return __ldcg(reinterpret_cast<const int32_t *>(lob) + k * MOM_MIDS + slot); }
  // sum of the n (<= 64) most recent of the L stored doubled mids: lanes take entries lane, lane+32; 16-bit halves keep the REDUX sums exact
  __device__ __forceinline__ int64_t mid_sum(int k, int L, int n) const {
    static_assert((MOM_MIDS & (MOM_MIDS - 1)) == 0, "ring size");
    const int32_t *m = reinterpret_cast<const int32_t *>(lob) + k * MOM_MIDS; int32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) { int i = lane + 32 * j; if (i < n) { int32_t v = __ldcg(m + ((L - 1 - i) & (MOM_MIDS - 1))); lo += v & 0xffff; hi += v >> 16; } }
    lo = __reduce_add_sync(FULL, lo); hi = __reduce_add_sync(FULL, hi);
    return ((int64_t)hi << 16) + lo;
  }
  __device__ __forceinline__ void mid_store(int k, int slot, int32_t v) { if (lane == 0) __stcg(reinterpret_cast<int32_t *>(lob) + k * MOM_MIDS + slot, v); __syncwarp(); }
  // np.std (ddof 0) of log(mid_i / p0) over the n stored LOBs: lanes take LOBs lane, lane+32, ... (ABIDESEnvMetrics.py:183-192)
  __device__ double lob_midvol(int n, int head, double p0, bool &bad) const {
    double v[4]; double sum = 0.0; bool b = false;
#pragma unroll
    for (int k = 0; k < 4; k++) { int i = lane + 32 * k; v[k] = 0.0;
      if (i < n) { int4 a = __ldcg(lob + ((head + i) % LOB_CAP) * 3); if (a.x <= 0 || a.w <= 0) b = true; v[k] = log_ni((((double)a.x + (double)a.w) / 2) / p0); sum += v[k]; } }
    bad = bad || __any_sync(FULL, b);
    double mean = warp_sum(sum) / n, var = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) if (lane + 32 * k < n) var += (v[k] - mean) * (v[k] - mean);
    return sqrt(warp_sum(var) / n);
  }

  // ---- trader records: 12 lanes x 128-bit, HBM <-> shared ----
  __device__ __forceinline__ ZiAgent *agent_stage(int id) {
    sync();
    if (lane < (int)(sizeof(ZiAgent) / 16)) reinterpret_cast<uint4 *>(staged)[lane] = ldcg4(reinterpret_cast<const uint4 *>(agents + id) + lane);
    xsync();                                                                 // 12 lanes staged 16 bytes each; every lane reads the whole record
    return staged;
  }
  // the same copy in two halves: issue the 128-bit loads as soon as the recipient is known, park them in shared memory once the event is unpacked
  uint4 pre_v;
  __device__ __forceinline__ void agent_load_issue(int id) { if (lane < (int)(sizeof(ZiAgent) / 16)) pre_v = ldcg4(reinterpret_cast<const uint4 *>(agents + id) + lane); }
  __device__ __forceinline__ ZiAgent *agent_stage_issued(int) {
    sync();
    if (lane < (int)(sizeof(ZiAgent) / 16)) reinterpret_cast<uint4 *>(staged)[lane] = pre_v;
    xsync();
    return staged;
  }
  __device__ __forceinline__ void agent_prefetch(int id) const {
    const char *p = reinterpret_cast<const char *>(agents + id);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
  }
  __device__ __forceinline__ void agent_commit(int id) {
    __syncwarp();
    if (lane < (int)(sizeof(ZiAgent) / 16)) __stcg(reinterpret_cast<uint4 *>(agents + id) + lane, reinterpret_cast<const uint4 *>(staged)[lane]);
  }
  __device__ __forceinline__ double agent_lat_from(int id) const { return __ldcg(&agents[id].lat_from); }
};
typedef WarpCtxT<0> WarpCtx;
typedef WarpCtxT<1> WarpCtxSmallQ;
typedef WarpCtxT<2> WarpCtxHybridQ;
typedef WarpCtxT<3> WarpCtxNearQ;

}  // namespace abx
