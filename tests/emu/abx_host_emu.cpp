// abx_host_emu.cpp -- HOST EMULATION HARNESS (test tool, CPU CI only).
//
// Compiles the product's warp-uniform simulation logic (marl_optimal_execution_b200/csrc/abx_core.cuh) as plain
// C++ with a serial context in place of the warp-cooperative one, so that the scalar logic (event rules,
// exchange protocol, matching, ZI agent, sparse OU oracle, latency models, RNG tape plumbing, HBM record
// layouts) is checked against the oracle on the CPU before any GPU minute is spent.  It is NOT part of
// libabides_b200.so, is not importable from the package and is never a fallback: the product fails loudly
// without the CUDA library.  It exposes the same C ABI names as include/abides_b200.h for the subset it
// implements so tests can drive both through one Python wrapper class.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <unordered_map>
#include "../../marl_optimal_execution_b200/csrc/abx_host_common.h"

using namespace abx;

struct HostCtx {
  const SimParams &P; int env;
  uint4 *qkey, *qpay0, *qpay1, *qcache; ZiAgent *agents; int32_t *lvp, *lvq; uint32_t *lvht; uint4 *nodes; abx_trace_rec *tr;
  static constexpr int OUTN = OUT_CAP;
  uint32_t outbox_[OUT_CAP * OUT_WORDS]; ZiAgent staged; int cur_group, cur_slot; uint4 *idt; int4 *lob; uint4 *hl;
  HostCtx(const SimParams &P_, int e) : P(P_), env(e) {
    size_t q = (size_t)e * P.c.queue_cap; qkey = P.qkey + q; qpay0 = P.qpay0 + q; qpay1 = P.qpay1 + q; qcache = P.qcache + (size_t)e * P.n_qgroups;
    agents = P.agents + (size_t)e * P.c.n_agents; size_t l = (size_t)e * 2 * P.c.level_cap; lvp = P.lv_price + l; lvq = P.lv_qty + l; lvht = P.lv_ht + l;
    nodes = P.nodes + (size_t)e * P.c.order_cap; tr = P.trace ? P.trace + (size_t)e * P.c.trace_cap : nullptr; cur_group = cur_slot = -1;
    idt = P.idtab ? P.idtab + (size_t)e * P.n_ids : nullptr; lob = P.lobs ? P.lobs + (size_t)e * lob_stride_of(P.c) : nullptr;
    hl = P.hlog ? P.hlog + (size_t)e * hist_stride_of(P.c) : nullptr;
  }
  bool onchip_writer() { return true; }
  void sync() {}
  uint32_t *outbox() { return outbox_; }
  void trace(const abx_trace_rec &r, uint32_t i) { tr[i] = r; }
  // ---- queue ----
  void q_clear() { for (int g = 0; g < P.n_qgroups; g++) { qcache[g].x = qcache[g].y = 0xffffffffu; qcache[g].z = 0xffffffffu; qcache[g].w = 0; } }
  void group_recompute(int g) {
    uint64_t bh = KEY_EMPTY; uint32_t bu = 0xffffffffu; uint32_t mask = qcache[g].w;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) { uint4 k = qkey[g * 32 + i]; uint64_t h = (uint64_t)k.x | ((uint64_t)k.y << 32); if (key_less(h, k.z, bh, bu)) { bh = h; bu = k.z; } }
    qcache[g].x = (uint32_t)bh; qcache[g].y = (uint32_t)(bh >> 32); qcache[g].z = bu;
  }
  bool q_min(uint64_t &hi, uint32_t &uniq, int &grp) {
    uint64_t bh = KEY_EMPTY; uint32_t bu = 0xffffffffu; grp = -1;
    for (int g = 0; g < P.n_qgroups; g++) { if (!qcache[g].w) continue; uint64_t h = (uint64_t)qcache[g].x | ((uint64_t)qcache[g].y << 32); if (grp < 0 || key_less(h, qcache[g].z, bh, bu)) { bh = h; bu = qcache[g].z; grp = g; } }
    hi = bh; uniq = bu; return grp >= 0;
  }
  void q_fetch(int g, Event &e) {
    uint64_t bh = KEY_EMPTY; uint32_t bu = 0xffffffffu; int best = -1; uint32_t mask = qcache[g].w;
    for (int i = 0; i < 32; i++) if (mask >> i & 1) { uint4 k = qkey[g * 32 + i]; uint64_t h = (uint64_t)k.x | ((uint64_t)k.y << 32); if (best < 0 || key_less(h, k.z, bh, bu)) { bh = h; bu = k.z; best = i; } }
    cur_group = g; cur_slot = best; event_unpack(qkey[g * 32 + best], qpay0[g * 32 + best], qpay1[g * 32 + best], e);
  }
  void q_remove() { qcache[cur_group].w &= ~(1u << cur_slot); group_recompute(cur_group); }
  void q_requeue(int64_t t) {
    uint4 &k = qkey[cur_group * 32 + cur_slot]; uint64_t h = (uint64_t)k.x | ((uint64_t)k.y << 32);
    h = key_pack(t < KEY_T_MAX ? t : KEY_T_MAX, key_recipient(h), key_type(h)); k.x = (uint32_t)h; k.y = (uint32_t)(h >> 32); group_recompute(cur_group);
  }
  void q_settle(bool requeue, int64_t t) { if (requeue) q_requeue(t); else q_remove(); }
  bool q_push(const Event &e, int64_t) {
    for (int g = 0; g < P.n_qgroups; g++) if (qcache[g].w != 0xffffffffu) {
      int i = __builtin_ctz(~qcache[g].w); event_pack(e, qkey[g * 32 + i], qpay0[g * 32 + i], qpay1[g * 32 + i]); qcache[g].w |= 1u << i; group_recompute(g); return true; }
    return false;
  }
  // ---- ladders: ascending in "goodness", best level LAST.  side 0 bids (ascending price), side 1 asks (descending price) ----
  int32_t lv_price(int side, int i) { return lvp[side * P.c.level_cap + i]; }
  int32_t lv_qty(int side, int i) { return lvq[side * P.c.level_cap + i]; }
  uint32_t lv_head(int side, int i) { return lvht[side * P.c.level_cap + i] & 0xffffu; }
  uint32_t lv_tail(int side, int i) { return lvht[side * P.c.level_cap + i] >> 16; }
  void lv_set(int side, int i, int32_t qty, uint32_t head, uint32_t tail) { lvq[side * P.c.level_cap + i] = qty; lvht[side * P.c.level_cap + i] = head | (tail << 16); }
  void lv_find(int side, int32_t price, int n, int &pos, bool &found) {
    int best = -1;
    for (int i = 0; i < n; i++) { int32_t p = lv_price(side, i); if (side == 0 ? p <= price : p >= price) best = i; }
    if (n > 0 && (side == 0 ? lv_price(side, 0) > price : lv_price(side, 0) < price)) best = -1;
    found = best >= 0 && lv_price(side, best) == price; pos = found ? best : best + 1;
  }
  int lv_find_eq(int side, int32_t price, int limit, int &cnt) { int best = -1; cnt = 0; for (int i = 0; i < limit; i++) if (lv_price(side, i) == price) { best = i; cnt++; } return best; }
  void lv_setp(int side, int i, int32_t price) { lvp[side * P.c.level_cap + i] = price; }
  void lv_insert(int side, int pos, int n, int32_t price, int32_t qty, uint32_t head, uint32_t tail) {
    int b = side * P.c.level_cap;
    for (int i = n; i > pos; i--) { lvp[b + i] = lvp[b + i - 1]; lvq[b + i] = lvq[b + i - 1]; lvht[b + i] = lvht[b + i - 1]; }
    lvp[b + pos] = price; lvq[b + pos] = qty; lvht[b + pos] = head | (tail << 16);
  }
  void lv_remove(int side, int pos, int n) { int b = side * P.c.level_cap; for (int i = pos; i + 1 < n; i++) { lvp[b + i] = lvp[b + i + 1]; lvq[b + i] = lvq[b + i + 1]; lvht[b + i] = lvht[b + i + 1]; } }
  // ---- order nodes ----
  template <bool PRICED> NodeRec node_load(uint32_t i) { return node_unpack<PRICED>(nodes[i]); }
  template <bool PRICED> void node_store(uint32_t i, const NodeRec &r) { nodes[i] = node_pack<PRICED>(r); }
  // ---- ABIDESEnv shape ----
  EnvX *envx() { return P.envx + env; }
  uint2 ib_load(int i) { return P.idbook[(size_t)env * P.n_ids + i]; }
  void ib_store(int i, uint2 v) { P.idbook[(size_t)env * P.n_ids + i] = v; }
  uint4 id_load(int i) { return idt[i]; }
  void id_store(int i, uint4 v) { idt[i] = v; }
  int tab_find(int base, int n, uint32_t oid) { for (int i = 0; i < n; i++) if (idt[base + i].x == oid) return i; return -1; }
  void tab_remove(int base, int n, int f) { for (int i = f; i + 1 < n; i++) idt[base + i] = idt[base + i + 1]; }
  int4 row_load(int r) { return P.st_rows[r]; }
  uint32_t episode = 0; void set_episode(uint32_t ep) { episode = ep; }
  int4 day_rec() { return P.day_tab[P.n_days > 1 ? (int)(((uint32_t)env + episode) % (uint32_t)P.n_days) : 0]; }
  void clear_tables() {
    if (idt) memset(idt, 0, sizeof(uint4) * P.n_ids);
    if (P.idbook) memset(P.idbook + (size_t)env * P.n_ids, 0, sizeof(uint2) * P.n_ids);
    if (lob) memset(lob, 0, sizeof(int4) * LOB_CAP * 3);
  }
  int n_ts() { return day_rec().y; }
  int64_t ts_load(int k) { return P.st_ts[day_rec().x + k]; }
  int first_load(int k) { return P.st_first[day_rec().z + k]; }
  int day_row0() { return first_load(0); }
  int4 day_rec2() { return P.day_tab2[P.n_days > 1 ? (int)(((uint32_t)env + episode) % (uint32_t)P.n_days) : 0]; }
  int xid_first_row(int4 d2, int32_t id) { const int32_t *b = P.st_xid + d2.x; const int32_t *e = b + d2.y; const int32_t *it = std::lower_bound(b, e, id); return (it != e && *it == id) ? P.st_xfirst[d2.x + (it - b)] : -1; }
  void lob_store(int slot, const int32_t w[12]) { for (int k = 0; k < 3; k++) { int4 v; v.x = w[4 * k]; v.y = w[4 * k + 1]; v.z = w[4 * k + 2]; v.w = w[4 * k + 3]; lob[slot * 3 + k] = v; } }
  void lob_load(int slot, int32_t w[12]) { for (int k = 0; k < 3; k++) { int4 v = lob[slot * 3 + k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; } }
  void id_prefetch(int) {} void ib_prefetch(int) {}
  void snap_store(int k, int side, int n_total, int n_copy) { int2 *dst = P.snap + ((size_t)env * P.n_snap * 2 + (size_t)(k * 2 + side)) * P.snap_depth;
    for (int i = 0; i < n_copy; i++) { dst[i].x = lv_price(side, n_total - 1 - i); dst[i].y = lv_qty(side, n_total - 1 - i); } }
  int2 snap_load(int k, int side, int i) { return P.snap[((size_t)env * P.n_snap * 2 + (size_t)(k * 2 + side)) * P.snap_depth + i]; }
  // population 3: order-history log and the HBL belief argmax, the candidate-price form of abx_warp.cuh hbl_best one candidate at a time
  uint4 hist_load(int slot) { return hl[slot]; }
  void hist_store(int slot, uint4 v) { hl[slot] = v; }
  bool hbl_best(uint32_t hist_n, uint32_t e_lo, uint32_t e_hi, bool buy, int32_t v, int32_t &best_p, uint32_t &err) {
    uint32_t cap = (uint32_t)P.c.hist_log_cap, avail = hist_n < cap ? hist_n : cap, k_first = avail, k_end = avail; bool stop = false;
    for (uint32_t k = 0; k < avail && !stop; k++) {
      uint32_t ep = hl[(hist_n - 1 - k) & (cap - 1)].z;
      if (ep >= e_lo && ep <= e_hi && k_first == avail) k_first = k;
      if (ep < e_lo) { k_end = k; stop = true; }
    }
    if (!stop && hist_n > cap) err |= ABX_F_HISTORY_OVERFLOW;
    if (k_first >= k_end) { err |= ABX_F_REF_EXCEPTION; return false; }
    uint32_t N = k_end - k_first, top = hist_n - 1 - k_first;
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (uint32_t j = 0; j < N; j++) { int32_t q = (int32_t)hl[(top - j) & (cap - 1)].y; lo = std::min(lo, q); hi = std::max(hi, q); }
    double bes = -1.0e300; int32_t bp = INT32_MAX;
    uint32_t rows_cap = P.c.hbl_table_rows > 0 ? (uint32_t)P.c.hbl_table_rows : cap / 4; uint2 *scr = reinterpret_cast<uint2 *>(hl + cap);
    if ((int64_t)hi - (int64_t)lo < (int64_t)rows_cap) {                                // histogram + prefix-scan form (the reference's own table), one row at a time
      uint32_t R = (uint32_t)(hi - lo) + 1u, atot = 0, btot = 0, ca = 0, cb = 0;
      for (uint32_t r = 0; r < R; r++) { scr[r].x = 0; scr[r].y = 0; }
      for (uint32_t j = 0; j < N; j++) { uint4 r = hl[(top - j) & (cap - 1)]; bool isb = r.w & 1u, tx = (r.w & 2u) != 0, fwd = tx || (buy ? !isb : isb);
        if (fwd) { scr[(int32_t)r.y - lo].x++; atot++; } else { scr[(int32_t)r.y - lo].y++; btot++; } }
      for (uint32_t r = 0; r < R; r++) {
        uint32_t ia = ca + scr[r].x, ib = cb + scr[r].y, num = buy ? ia : atot - ca, oth = buy ? btot - cb : ib, den = num + oth; int32_t p = lo + (int32_t)r;
        double pr = den == 0 ? 0.0 : (double)num / (double)den, es = pr * (double)(buy ? v - p : p - v);
        if (es > bes || (es == bes && p < bp)) { bes = es; bp = p; }
        ca = ia; cb = ib;
      }
      best_p = bp; return bes > 0.0;
    }
    for (uint32_t ci = 0; ci < 3 * N + 2; ci++) {
      int32_t p = ci < 3 * N ? (int32_t)hl[(top - ci / 3) & (cap - 1)].y + (int32_t)(ci % 3) - 1 : (ci == 3 * N ? lo : hi);
      if (p < lo || p > hi) continue;
      uint32_t num = 0, den = 0;
      for (uint32_t j = 0; j < N; j++) {
        uint4 r = hl[(top - j) & (cap - 1)]; int32_t q = (int32_t)r.y; bool isb = r.w & 1u, tx = (r.w & 2u) != 0, le = q <= p, ge = q >= p;
        bool in_num = buy ? (le && (tx || !isb)) : (ge && (tx || isb)), extra = buy ? (!tx && isb && ge) : (!tx && !isb && le);
        num += in_num; den += in_num || extra;
      }
      double pr = den == 0 ? 0.0 : (double)num / (double)den, es = pr * (double)(buy ? v - p : p - v);
      if (es > bes || (es == bes && p < bp)) { bes = es; bp = p; }
    }
    best_p = bp; return bes > 0.0;
  }
  uint4 ord_load(int i) { return id_load(i); } void ord_store(int i, uint4 v) { id_store(i, v); }
  int64_t mid_sum(int k, int L, int n) { int64_t t = 0; for (int i = 0; i < n; i++) t += mid_load(k, (L - 1 - i) % MOM_MIDS); return t; }
  int32_t mid_load(int k, int slot) { return reinterpret_cast<int32_t *>(lob)[k * MOM_MIDS + slot]; }
  void mid_store(int k, int slot, int32_t v) { reinterpret_cast<int32_t *>(lob)[k * MOM_MIDS + slot] = v; }
  double lob_midvol(int n, int head, double p0, bool &bad) {
    double v[LOB_CAP], mean = 0, var = 0;
    for (int i = 0; i < n; i++) { int4 a = lob[((head + i) % LOB_CAP) * 3]; if (a.x <= 0 || a.w <= 0) bad = true; v[i] = log((((double)a.x + (double)a.w) / 2) / p0); mean += v[i]; }
    mean /= n; for (int i = 0; i < n; i++) var += (v[i] - mean) * (v[i] - mean);
    return sqrt(var / n);
  }
  // ---- agents ----
  ZiAgent *agent_stage(int id) { staged = agents[id]; return &staged; }
  void agent_load_issue(int) {} ZiAgent *agent_stage_issued(int id) { return agent_stage(id); } void agent_prefetch(int) {}
  void agent_commit(int id) { agents[id] = staged; }
  double agent_lat_from(int id) { return agents[id].lat_from; }
};

struct abx_sim {
  SimParams P; int n_envs; bool reset_done;
  std::vector<uint4> qkey, qpay0, qpay1, qcache, nodes; std::vector<ZiAgent> agents; std::vector<int32_t> lvp, lvq; std::vector<uint32_t> lvht;
  std::vector<EnvState> env; std::vector<abx_trace_rec> trace; std::vector<uint4> draw_log, evt, hlog; std::vector<uint64_t> tbits; std::vector<uint8_t> tkinds; std::vector<int64_t> toff;
  bool is_env; EnvStreamHost st; EnvDaysHost dh; bool has_days = false; std::vector<EnvX> envx; std::vector<uint4> idtab; std::vector<int4> lobs; std::vector<uint2> idbook; std::vector<int2> snap; int auto_reset = 0; std::vector<uint64_t> dq_seeds; std::vector<int32_t> dq_msizes, sched;
};

extern "C" {
const char *abx_strerror(int32_t st) { return status_string(st); }
const char *abx_last_cuda_error(void) { return "host emulation harness: no CUDA"; }
int32_t abx_device_count(void) { return 0; }
int32_t abx_config_sparse_zi(int32_t variant, abx_sim_config *cfg) { return config_sparse_zi(variant, cfg); }
int32_t abx_config_rmsc03(abx_sim_config *cfg) { return config_rmsc03(cfg); }
int32_t abx_config_rmsc03_pov(abx_sim_config *cfg) { return config_rmsc03_pov(cfg); }
int32_t abx_config_rmsc01(abx_sim_config *cfg) { return config_rmsc01(cfg); }
int32_t abx_config_rmsc02(abx_sim_config *cfg) { return config_rmsc02(cfg); }
int32_t abx_sim_pov_exec(abx_sim *h, int32_t env, int64_t *out, void *stream) {
  (void)stream; if (!h || !out || env < 0 || env >= h->n_envs || h->P.c.population != 1 || !h->P.c.n_pov_exec) return ABX_ERR_ARG;
  const ZiAgent &z = h->agents[(size_t)env * h->P.c.n_agents + h->P.c.n_agents - 1]; const ExecAux *ex = reinterpret_cast<const ExecAux *>(z.oid);
  out[0] = ex->rem_qty; out[1] = ex->n_executed; out[2] = z.n_orders; return ABX_OK;
}
typedef Sim<HostCtx, -1, ABX_LAT_ZERO, true, SHAPE_R3> R3SimHost;
typedef Sim<HostCtx, -1, -1, true, SHAPE_P3> P3SimHost;   // latency model decided at run time: zero (rmsc01) or matrix + noise (rmsc02)

int32_t abx_sim_create(const abx_sim_config *cfg, int32_t n_envs, int32_t device, abx_sim **out) {
  (void)device;
  if (!out || n_envs < 1 || config_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  abx_sim *h = new abx_sim(); memset(&h->P, 0, sizeof(h->P)); h->is_env = false; h->P.c = *cfg; h->P.n_envs = n_envs; h->n_envs = n_envs; h->reset_done = false; derive_params(h->P);
  size_t E = n_envs; const abx_sim_config &c = *cfg;
  h->qkey.resize(E * c.queue_cap); h->qpay0.resize(E * c.queue_cap); h->qpay1.resize(E * c.queue_cap); h->qcache.resize(E * h->P.n_qgroups);
  h->agents.resize(E * c.n_agents); h->lvp.resize(E * 2 * c.level_cap); h->lvq.resize(E * 2 * c.level_cap); h->lvht.resize(E * 2 * c.level_cap);
  h->nodes.resize(E * c.order_cap); h->env.resize(E); h->trace.resize(E * (size_t)c.trace_cap);
  h->P.qkey = h->qkey.data(); h->P.qpay0 = h->qpay0.data(); h->P.qpay1 = h->qpay1.data(); h->P.qcache = h->qcache.data(); h->P.agents = h->agents.data();
  h->P.lv_price = h->lvp.data(); h->P.lv_qty = h->lvq.data(); h->P.lv_ht = h->lvht.data(); h->P.nodes = h->nodes.data(); h->P.env = h->env.data();
  h->P.trace = c.trace_cap ? h->trace.data() : nullptr;
  h->draw_log.resize(E * (size_t)c.draw_log_cap); h->P.draw_log = c.draw_log_cap ? h->draw_log.data() : nullptr;
  h->evt.resize(E * (size_t)c.event_ring_cap); h->P.evt = c.event_ring_cap ? h->evt.data() : nullptr;
  if (c.population == 1) { h->P.dq_order_base = MM_ORDER_CAP + h->P.tv_ring; h->P.n_ids = h->P.dq_order_base + (c.n_pov_exec ? EXEC_ORDER_CAP : 0); h->idtab.resize(E * h->P.n_ids); h->lobs.resize(E * LOB_CAP * 3); h->P.idtab = h->idtab.data(); h->P.lobs = h->lobs.data();
    if (c.n_pov_exec) { h->P.n_snap = 1; h->P.snap_depth = c.level_cap; h->snap.resize(E * 2 * (size_t)c.level_cap); h->P.snap = h->snap.data(); } }
  if (c.population == 3) { h->P.n_ids = MM_ORDER_CAP + SUB_CAP; h->P.n_snap = SUB_CAP; h->P.snap_depth = SUB_LEVELS; h->snap.resize(E * (size_t)SUB_CAP * 2 * SUB_LEVELS); h->P.snap = h->snap.data(); h->idtab.resize(E * h->P.n_ids); h->lobs.resize(E * lob_stride_of(c)); h->hlog.resize(E * hist_stride_of(c)); h->P.idtab = h->idtab.data(); h->P.lobs = h->lobs.data(); h->P.hlog = h->hlog.data(); }
  *out = h; return ABX_OK;
}
int32_t abx_sim_destroy(abx_sim *h) { delete h; return ABX_OK; }
int64_t abx_sim_device_bytes(const abx_sim *h) { (void)h; return 0; }

static int32_t do_reset(abx_sim *h, const uint64_t *seeds) {
  const abx_sim_config &c = h->P.c;
  for (int e = 0; e < h->n_envs; e++) {
    uint64_t seed = seeds ? seeds[e] : 0; uint32_t err = 0;
    for (int id = 1; id < c.n_agents; id++) init_agent_record(h->P, e, id, seed, &h->agents[(size_t)e * c.n_agents + id], &err);
    EnvState s; init_env_state(h->P, seed, s); s.flags |= err;
    HostCtx ctx(h->P, e); ctx.q_clear();
    Sim<HostCtx> sim(ctx, h->P, s, e); sim.reset_env(); h->env[e] = sim.s;
  }
  h->reset_done = true; return ABX_OK;
}
int32_t abx_sim_reset_philox(abx_sim *h, const uint64_t *seeds, void *stream) {
  (void)stream; if (!h || !seeds || h->P.c.rng_mode != ABX_RNG_PHILOX) return ABX_ERR_ARG; return do_reset(h, seeds);
}
static int32_t reset_tape_impl(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to, const double *lat_from);
int32_t abx_sim_reset_tape(abx_sim *h, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to, const double *lat_from, void *stream) {
  (void)stream; return reset_tape_impl(h, 0, bits, kinds, off, lat_to, lat_from); }
int32_t abx_sim_reset_tape_shared(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to, const double *lat_from, void *stream) {
  (void)stream; if (!h || n_tapes < 1 || n_tapes > h->n_envs || !lat_to || !lat_from) return ABX_ERR_ARG;
  size_t na = (size_t)h->P.c.n_agents, E = (size_t)h->n_envs; std::vector<double> lt(E * na), lf(E * na);
  for (size_t e = 0; e < E; e++) { memcpy(&lt[e * na], lat_to + (e % n_tapes) * na, sizeof(double) * na); memcpy(&lf[e * na], lat_from + (e % n_tapes) * na, sizeof(double) * na); }
  return reset_tape_impl(h, n_tapes, bits, kinds, off, lt.data(), lf.data());
}
static int32_t reset_tape_impl(abx_sim *h, int32_t n_tapes, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const double *lat_to, const double *lat_from) {
  if (!h || !bits || !kinds || !off || !lat_to || !lat_from || h->P.c.rng_mode != ABX_RNG_TAPE) return ABX_ERR_ARG;
  h->P.n_tapes = n_tapes;
  size_t nS = (size_t)(n_tapes > 0 ? n_tapes : h->n_envs) * h->P.n_streams; int64_t total = off[nS];
  h->tbits.assign(bits, bits + total); h->tkinds.assign(kinds, kinds + total); h->toff.assign(off, off + nS + 1);
  h->P.tape_bits = h->tbits.data(); h->P.tape_kinds = h->tkinds.data(); h->P.tape_off = h->toff.data();
  for (int e = 0; e < h->n_envs; e++) for (int id = 0; id < h->P.c.n_agents; id++) { ZiAgent &z = h->agents[(size_t)e * h->P.c.n_agents + id]; z.lat_to = lat_to[(size_t)e * h->P.c.n_agents + id]; z.lat_from = lat_from[(size_t)e * h->P.c.n_agents + id]; }
  return do_reset(h, nullptr);
}
int32_t abx_sim_run(abx_sim *h, int64_t until_ns, void *stream) {
  (void)stream; if (!h) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  for (int e = 0; e < h->n_envs; e++) { HostCtx ctx(h->P, e);
    if (h->P.c.population == 1) { R3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_run(until_ns); h->env[e] = sim.s; }
    else if (h->P.c.population == 3) { P3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_run(until_ns); h->env[e] = sim.s; }
    else { Sim<HostCtx> sim(ctx, h->P, h->env[e], e); sim.run(until_ns); h->env[e] = sim.s; } }
  return ABX_OK;
}
int32_t abx_sim_run_each(abx_sim *h, const int64_t *until, void *stream) {
  (void)stream; if (!h || !until) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  for (int e = 0; e < h->n_envs; e++) { HostCtx ctx(h->P, e);
    if (h->P.c.population == 1) { R3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_run(until[e]); h->env[e] = sim.s; }
    else if (h->P.c.population == 3) { P3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_run(until[e]); h->env[e] = sim.s; }
    else { Sim<HostCtx> sim(ctx, h->P, h->env[e], e); sim.run(until[e]); h->env[e] = sim.s; } }
  return ABX_OK;
}
int32_t abx_sim_finalize(abx_sim *h, void *stream) {
  (void)stream; if (!h) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  for (int e = 0; e < h->n_envs; e++) { HostCtx ctx(h->P, e);
    if (h->P.c.population == 1) { R3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_finalize(); h->env[e] = sim.s; }
    else if (h->P.c.population == 3) { P3SimHost sim(ctx, h->P, h->env[e], e); sim.r3_finalize(); h->env[e] = sim.s; }
    else { Sim<HostCtx> sim(ctx, h->P, h->env[e], e); sim.finalize(); h->env[e] = sim.s; } }
  return ABX_OK;
}
int32_t abx_sim_stats(abx_sim *h, abx_env_stats *out, void *stream) {
  (void)stream; if (!h || !out) return ABX_ERR_ARG;
  for (int e = 0; e < h->n_envs; e++) {
    const EnvState &s = h->env[e]; HostCtx ctx(h->P, e); int nb = s.n_bid_lv, na = s.n_ask_lv;
    fill_stats(s, nb ? ctx.lv_price(0, nb - 1) : 0, nb ? ctx.lv_qty(0, nb - 1) : 0, na ? ctx.lv_price(1, na - 1) : 0, na ? ctx.lv_qty(1, na - 1) : 0, &out[e]);
  }
  return ABX_OK;
}
int32_t abx_sim_stats_device(abx_sim *h, abx_env_stats *out, void *stream) { return abx_sim_stats(h, out, stream); }
int32_t abx_sim_holdings(abx_sim *h, int32_t env, int64_t *out, void *stream) {
  (void)stream; if (!h || !out || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  for (int id = 1; id < h->P.c.n_agents; id++) { const ZiAgent &z = h->agents[(size_t)env * h->P.c.n_agents + id]; int64_t *r = out + 5 * (id - 1);
    r[0] = id; r[1] = z.shares; r[2] = z.cash; r[3] = z.cash + (int64_t)z.shares * ((z.flags & AF_HAS_LAST) ? z.last_trade : 0); r[4] = z.surplus; }
  return ABX_OK;
}
int32_t abx_sim_book_snapshot(abx_sim *h, int32_t env, int32_t is_bid, int32_t depth, int32_t *out, int32_t *n_levels, void *stream) {
  (void)stream; if (!h || !out || !n_levels || env < 0 || env >= h->n_envs || depth < 0) return ABX_ERR_ARG;
  HostCtx ctx(h->P, env); int side = is_bid ? 0 : 1, n = side ? h->env[env].n_ask_lv : h->env[env].n_bid_lv, m = depth < n ? depth : n;
  for (int k = 0; k < m; k++) { out[2 * k] = ctx.lv_price(side, n - 1 - k); out[2 * k + 1] = ctx.lv_qty(side, n - 1 - k); }
  *n_levels = m; return ABX_OK;
}
int32_t abx_sim_draw_log(abx_sim *h, int32_t env, abx_draw_rec *out, int32_t max_recs, int32_t *n_recs, void *stream) {
  (void)stream; if (!h || !out || !n_recs || env < 0 || env >= h->n_envs || h->is_env) return ABX_ERR_ARG;
  int n = (int)h->env[env].draw_n; if (n > max_recs) n = max_recs;
  if (n) memcpy(out, h->P.draw_log + (size_t)env * h->P.c.draw_log_cap, sizeof(uint4) * n);
  *n_recs = n; return ABX_OK;
}
int32_t abx_sim_events_device(abx_sim *h, abx_event_rec *out, uint32_t *counts, void *stream) {
  (void)stream; if (!h || h->is_env || !out || !counts || h->P.c.event_ring_cap <= 0) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  memcpy(out, h->evt.data(), sizeof(uint4) * h->evt.size()); for (int e = 0; e < h->n_envs; e++) counts[e] = h->env[e].evt_n; return ABX_OK;
}
int32_t abx_sim_agent_init(abx_sim *h, int32_t env, int32_t *theta, double *lat_to, double *lat_from, int32_t *sizes, int64_t *wakes, void *stream) {
  (void)stream; if (!h || h->is_env || env < 0 || env >= h->n_envs) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  agent_init_rows(h->P, &h->agents[(size_t)env * h->P.c.n_agents], theta, lat_to, lat_from, sizes, wakes); return ABX_OK;
}
int32_t abx_sim_trace(abx_sim *h, int32_t env, abx_trace_rec *out, int32_t max_recs, int32_t *n_recs, void *stream) {
  (void)stream; if (!h || !out || !n_recs || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  int n = (int)h->env[env].trace_n; if (n > max_recs) n = max_recs;
  if (n) memcpy(out, h->P.trace + (size_t)env * h->P.c.trace_cap, sizeof(abx_trace_rec) * n);
  if (h->is_env) { const std::vector<int64_t> &ido = h->has_days ? h->dh.days[(int)(((uint32_t)env + h->env[env].episode) % (uint32_t)h->dh.days.size())].id_orig : h->st.id_orig;
    for (int i = 0; i < n; i++) if (out[i].tag == 1 && (uint32_t)out[i].v[1] >= REPLAY_ID_BASE) out[i].v[1] = (int32_t)ido[(uint32_t)out[i].v[1] - REPLAY_ID_BASE]; }
  *n_recs = n; return ABX_OK;
}
int64_t abx_sim_launch_count(const abx_sim *h) { (void)h; return 0; }

typedef Sim<HostCtx, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_ENV> EnvSimHost;
int32_t abx_env_config_default(abx_env_config *cfg) { return env_config_default(cfg); }
int32_t abx_env_create_days(const abx_env_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out) {
  (void)device; if (!out || n_envs < 1 || env_shape_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  abx_sim *h = new abx_sim(); memset(&h->P, 0, sizeof(h->P)); h->is_env = true; h->has_days = true; h->n_envs = n_envs; h->reset_done = false;
  if (env_build_days(stream5, row_offsets, n_days, 4LL * cfg->n_horizon + 16, h->dh) != ABX_OK) { delete h; return ABX_ERR_ARG; }
  env_fill_params(*cfg, h->P); h->P.n_envs = n_envs; const abx_sim_config &c = h->P.c; size_t E = n_envs;
  h->P.n_ts = (int)h->dh.ts.size(); h->P.n_rows = (int)h->dh.rows.size(); h->P.n_ids = h->dh.max_ids; h->P.n_days = n_days;
  h->qkey.resize(E * c.queue_cap); h->qpay0.resize(E * c.queue_cap); h->qpay1.resize(E * c.queue_cap); h->qcache.resize(E * h->P.n_qgroups);
  h->agents.resize(4); h->lvp.resize(E * 2 * c.level_cap); h->lvq.resize(E * 2 * c.level_cap); h->lvht.resize(E * 2 * c.level_cap);
  h->nodes.resize(E * c.order_cap); h->env.resize(E); h->trace.resize(E * (size_t)c.trace_cap);
  h->envx.resize(E); h->idtab.resize(E * h->P.n_ids); h->lobs.resize(E * LOB_CAP * 3); h->idbook.resize(E * h->P.n_ids);
  h->P.qkey = h->qkey.data(); h->P.qpay0 = h->qpay0.data(); h->P.qpay1 = h->qpay1.data(); h->P.qcache = h->qcache.data(); h->P.agents = h->agents.data();
  h->P.lv_price = h->lvp.data(); h->P.lv_qty = h->lvq.data(); h->P.lv_ht = h->lvht.data(); h->P.nodes = h->nodes.data(); h->P.env = h->env.data();
  h->P.trace = c.trace_cap ? h->trace.data() : nullptr; h->P.envx = h->envx.data(); h->P.idtab = h->idtab.data(); h->P.lobs = h->lobs.data(); h->P.idbook = h->idbook.data();
  h->P.st_ts = h->dh.ts.data(); h->P.st_first = h->dh.first.data(); h->P.st_rows = h->dh.rows.data(); h->P.day_tab = h->dh.day_tab.data(); h->P.day_tab2 = h->dh.day_tab2.data(); h->P.st_xid = h->dh.xid.data(); h->P.st_xfirst = h->dh.xfirst.data();
  *out = h; return ABX_OK;
}
int32_t abx_env_create(const abx_env_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out) {
  int64_t off[2] = {0, n_rows}; return abx_env_create_days(cfg, stream5, off, 1, n_envs, device, out);
}
enum { RESET_ALL = 0, RESET_MASK = 1, RESET_DONE = 2 };
static bool reset_selected(abx_sim *h, int e, const uint8_t *mask, int mode, int advance_day, uint32_t &episode) {
  episode = 0; if (mode == RESET_ALL) return true;
  if (mode == RESET_MASK && !mask[e]) return false;
  if (mode == RESET_DONE && !(h->env[e].flags & ABX_F_DONE)) return false;
  episode = h->env[e].episode + (advance_day ? 1u : 0u); return true;
}
static void emu_env_reset(abx_sim *h, const uint8_t *mask, int mode, int advance_day) {
  for (int e = 0; e < h->n_envs; e++) {
    uint32_t episode; if (!reset_selected(h, e, mask, mode, advance_day, episode)) continue;
    HostCtx ctx(h->P, e); ctx.set_episode(episode); ctx.clear_tables();
    EnvState s; init_env_state(h->P, 0, s); s.last_trade = -1; s.episode = episode; init_envx(h->P, h->envx[e]);     // no oracle: last_trade None (ExchangeAgent.py:97-102)
    ctx.q_clear(); EnvSimHost sim(ctx, h->P, s, e); sim.env_reset(); h->env[e] = sim.s;
  }
}
static void emu_dq_reset(abx_sim *h, const uint8_t *mask, int mode, int advance_day);
int32_t abx_env_reset(abx_sim *h, void *stream) {
  (void)stream; if (!h || !h->is_env || h->P.c.population == 2) return ABX_ERR_ARG;
  emu_env_reset(h, nullptr, RESET_ALL, 0);
  h->reset_done = true; return ABX_OK;
}
int32_t abx_env_reset_mask(abx_sim *h, const uint8_t *mask, int32_t advance_day, void *stream) {
  (void)stream; if (!h || !h->is_env || !mask) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  if (h->P.c.population == 2) emu_dq_reset(h, mask, RESET_MASK, advance_day ? 1 : 0); else emu_env_reset(h, mask, RESET_MASK, advance_day ? 1 : 0);
  return ABX_OK;
}
int32_t abx_env_set_auto_reset(abx_sim *h, int32_t mode) { if (!h || !h->is_env || mode < 0 || mode > 2) return ABX_ERR_ARG; h->auto_reset = mode; return ABX_OK; }
int32_t abx_env_step_host(abx_sim *h, const double *actions, double *obs, double *reward, uint8_t *done, void *stream) {
  (void)stream; if (!h || !h->is_env || !actions || !obs || !done) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  for (int e = 0; e < h->n_envs; e++) {
    EnvX &x = h->envx[e];
    if (!(h->env[e].flags & ABX_F_DONE)) { HostCtx ctx(h->P, e); ctx.set_episode(h->env[e].episode); EnvSimHost sim(ctx, h->P, h->env[e], e); sim.env_step(actions[3 * e], actions[3 * e + 1], actions[3 * e + 2]); h->env[e] = sim.s; }
    else x.obs_len = 0;
    for (int i = 0; i < 9; i++) obs[9 * e + i] = x.obs_len ? x.obs[i] : 0.0;
    if (reward) reward[e] = 0.0; done[e] = (h->env[e].flags & ABX_F_DONE) ? 1 : 0;
  }
  if (h->auto_reset) emu_env_reset(h, nullptr, RESET_DONE, h->auto_reset == 2);
  return ABX_OK;
}
int32_t abx_env_step(abx_sim *h, const double *a, double *o, double *r, uint8_t *d, void *s) { return abx_env_step_host(h, a, o, r, d, s); }

// ---- DDQN execution shape ----
typedef Sim<HostCtx, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_DQ> DqSimHost;
int32_t abx_dq_config_default(abx_dq_config *cfg) { return dq_config_default(cfg); }
int32_t abx_dq_create_days(const abx_dq_config *cfg, const int64_t *stream5, const int64_t *row_offsets, int32_t n_days, int32_t n_envs, int32_t device, abx_sim **out) {
  (void)device; if (!out || n_envs < 1 || dq_config_validate(cfg) != ABX_OK) return ABX_ERR_ARG;
  abx_sim *h = new abx_sim(); memset(&h->P, 0, sizeof(h->P)); h->is_env = true; h->has_days = true; h->n_envs = n_envs; h->reset_done = false;
  if (env_build_days(stream5, row_offsets, n_days, dq_max_generated_ids(*cfg), h->dh) != ABX_OK) { delete h; return ABX_ERR_ARG; }
  dq_fill_params(*cfg, h->P); h->P.n_envs = n_envs; const abx_sim_config &c = h->P.c; size_t E = n_envs;
  h->P.n_ts = (int)h->dh.ts.size(); h->P.n_rows = (int)h->dh.rows.size(); h->P.dq_order_base = h->dh.max_ids; h->P.dq_id_limit = h->dh.min_id; h->P.n_days = n_days;
  h->P.n_ids = h->P.dq_order_base + (cfg->n_twap + (cfg->has_ddqn ? 1 : 0)) * EXEC_ORDER_CAP;
  h->qkey.resize(E * c.queue_cap); h->qpay0.resize(E * c.queue_cap); h->qpay1.resize(E * c.queue_cap); h->qcache.resize(E * h->P.n_qgroups);
  h->agents.resize(E * c.n_agents); h->lvp.resize(E * 2 * c.level_cap); h->lvq.resize(E * 2 * c.level_cap); h->lvht.resize(E * 2 * c.level_cap);
  h->nodes.resize(E * c.order_cap); h->env.resize(E); h->trace.resize(E * (size_t)c.trace_cap);
  h->envx.resize(E); h->idtab.resize(E * h->P.n_ids); h->lobs.resize(E * LOB_CAP * 3); h->idbook.resize(E * h->P.n_ids);
  h->P.qkey = h->qkey.data(); h->P.qpay0 = h->qpay0.data(); h->P.qpay1 = h->qpay1.data(); h->P.qcache = h->qcache.data(); h->P.agents = h->agents.data();
  h->P.lv_price = h->lvp.data(); h->P.lv_qty = h->lvq.data(); h->P.lv_ht = h->lvht.data(); h->P.nodes = h->nodes.data(); h->P.env = h->env.data();
  h->P.trace = c.trace_cap ? h->trace.data() : nullptr; h->P.envx = h->envx.data(); h->P.idtab = h->idtab.data(); h->P.lobs = h->lobs.data(); h->P.idbook = h->idbook.data();
  h->P.st_ts = h->dh.ts.data(); h->P.st_first = h->dh.first.data(); h->P.st_rows = h->dh.rows.data(); h->P.day_tab = h->dh.day_tab.data(); h->P.day_tab2 = h->dh.day_tab2.data(); h->P.st_xid = h->dh.xid.data(); h->P.st_xfirst = h->dh.xfirst.data();
  { int n_exec = cfg->n_twap + (cfg->has_ddqn ? 1 : 0); h->P.n_snap = n_exec > 0 ? n_exec : 1; h->P.snap_depth = DQ_DEPTH; h->snap.resize(E * (size_t)h->P.n_snap * 2 * DQ_DEPTH); h->P.snap = h->snap.data(); }
  *out = h; return ABX_OK;
}
int32_t abx_dq_create(const abx_dq_config *cfg, const int64_t *stream5, int64_t n_rows, int32_t n_envs, int32_t device, abx_sim **out) {
  int64_t off[2] = {0, n_rows}; return abx_dq_create_days(cfg, stream5, off, 1, n_envs, device, out);
}
static void emu_dq_reset(abx_sim *h, const uint8_t *mask, int mode, int advance_day) {
  for (int e = 0; e < h->n_envs; e++) {
    uint32_t episode; if (!reset_selected(h, e, mask, mode, advance_day, episode)) continue;
    HostCtx ctx(h->P, e); ctx.set_episode(episode); ctx.clear_tables();
    uint64_t seed = (h->dq_seeds.empty() ? 0 : h->dq_seeds[e]) + episode;
    EnvState s; init_env_state(h->P, seed, s); s.last_trade = -1; s.episode = episode; init_envx(h->P, h->envx[e]);
    for (int id = 2; id < h->P.c.n_agents; id++)
      init_agent_record_dq(h->P, e, id, seed, (!h->dq_msizes.empty() && id < 2 + h->P.dq_n_mom) ? h->dq_msizes[(size_t)e * h->P.dq_n_mom + id - 2] : -1, &h->agents[(size_t)e * h->P.c.n_agents + id]);
    ctx.q_clear(); DqSimHost sim(ctx, h->P, s, e); sim.env_reset(); h->env[e] = sim.s;
  }
}
int32_t abx_dq_set_schedule(abx_sim *h, int32_t k, const int32_t *qty, int32_t n) {
  if (!h || !h->is_env || h->P.c.population != 2 || k < 0 || k >= h->P.dq_n_twap || !qty || n < 1) return ABX_ERR_ARG;
  if (h->sched.empty()) h->sched.assign((size_t)h->P.dq_n_twap * h->P.n_h, -1);
  for (int i = 0; i < h->P.n_h; i++) h->sched[(size_t)k * h->P.n_h + i] = i < n ? qty[i] : -1;
  h->P.sched = h->sched.data(); return ABX_OK;
}
int32_t abx_dq_reset(abx_sim *h, const uint64_t *seeds, const int32_t *mom_sizes, void *stream) {
  (void)stream; if (!h || !h->is_env || h->P.c.population != 2) return ABX_ERR_ARG;
  h->dq_seeds.clear(); h->dq_msizes.clear();
  if (seeds) h->dq_seeds.assign(seeds, seeds + h->n_envs);
  if (mom_sizes && h->P.dq_n_mom > 0) h->dq_msizes.assign(mom_sizes, mom_sizes + (size_t)h->n_envs * h->P.dq_n_mom);
  emu_dq_reset(h, nullptr, RESET_ALL, 0);
  h->reset_done = true; return ABX_OK;
}
int32_t abx_dq_step_host(abx_sim *h, const int32_t *actions, double *obs, double *trans, double *reward, uint8_t *done, void *stream) {
  (void)stream; if (!h || !h->is_env || h->P.c.population != 2 || !obs || !trans || !done) return ABX_ERR_ARG; if (!h->reset_done) return ABX_ERR_STATE;
  for (int e = 0; e < h->n_envs; e++) {
    EnvX &x = h->envx[e]; bool was_done = (h->env[e].flags & ABX_F_DONE) != 0, paused = false;
    if (!was_done) { HostCtx ctx(h->P, e); ctx.set_episode(h->env[e].episode); DqSimHost sim(ctx, h->P, h->env[e], e); paused = sim.dq_step(actions ? actions[e] : 0); h->env[e] = sim.s; }
    for (int i = 0; i < 8; i++) obs[8 * e + i] = paused ? x.obs[i] : 0.0;
    for (int i = 0; i < 6; i++) trans[6 * e + i] = NAN;
    double rw = 0.0;
    if (!was_done && h->P.dq_has_ddqn) {
      const ZiAgent &z = h->agents[(size_t)e * h->P.c.n_agents + h->P.c.n_agents - 1]; const ExecAux *ex = reinterpret_cast<const ExecAux *>(z.oid);
      if (ex->exflags & EXF_E_VALID) { double *t = trans + 6 * e; t[0] = ex->e_s[0]; t[1] = ex->e_s[1]; t[2] = ex->e_a; t[3] = ex->e_sp[0]; t[4] = ex->e_sp[1]; t[5] = (ex->exflags & EXF_E_R) ? ex->e_r : NAN; }
      rw = ex->step_reward;
    }
    if (reward) reward[e] = rw; done[e] = (h->env[e].flags & ABX_F_DONE) ? 1 : 0;
  }
  if (h->auto_reset) emu_dq_reset(h, nullptr, RESET_DONE, h->auto_reset == 2);
  return ABX_OK;
}
int32_t abx_dq_step(abx_sim *h, const int32_t *a, double *o, double *t, double *r, uint8_t *d, void *s) { return abx_dq_step_host(h, a, o, t, r, d, s); }
int32_t abx_dq_holdings(abx_sim *h, int32_t env, int64_t *out, double *exec_out, void *stream) {
  (void)stream; if (!h || !h->is_env || h->P.c.population != 2 || !out || env < 0 || env >= h->n_envs) return ABX_ERR_ARG;
  dq_holdings_rows(h->P, &h->agents[(size_t)env * h->P.c.n_agents], h->envx[env], out, exec_out); return ABX_OK;
}

// ---- Book surface ----
typedef Sim<HostCtx, ABX_RNG_PHILOX, ABX_LAT_ZERO, true, SHAPE_BOOK> BookSimHost;
static std::unordered_map<abx_sim *, std::unordered_map<int64_t, int32_t>> g_book_ids;
int32_t abx_book_create(int32_t stream_history, int32_t level_cap, int32_t order_cap, int32_t trace_cap, int32_t n_envs, int32_t device, abx_sim **out) {
  (void)device; abx_env_config ec; env_config_default(&ec);
  ec.order_level = 0; ec.stream_history = stream_history; ec.queue_cap = 32; ec.level_cap = level_cap; ec.order_cap = order_cap; ec.trace_cap = trace_cap; ec.hash_pops = 0;
  if (!out || n_envs < 1 || env_config_validate(&ec) != ABX_OK) return ABX_ERR_ARG;
  abx_sim *h = new abx_sim(); memset(&h->P, 0, sizeof(h->P)); h->is_env = true; h->n_envs = n_envs; h->reset_done = false;
  env_fill_params(ec, h->P); h->P.n_envs = n_envs; h->P.c.n_agents = 2; const abx_sim_config &c = h->P.c; size_t E = n_envs;
  h->qkey.resize(E * c.queue_cap); h->qpay0.resize(E * c.queue_cap); h->qpay1.resize(E * c.queue_cap); h->qcache.resize(E * h->P.n_qgroups);
  h->agents.resize(4); h->lvp.resize(E * 2 * c.level_cap); h->lvq.resize(E * 2 * c.level_cap); h->lvht.resize(E * 2 * c.level_cap);
  h->nodes.resize(E * c.order_cap); h->env.resize(E); h->trace.resize(E * (size_t)c.trace_cap); h->envx.resize(E);
  h->P.qkey = h->qkey.data(); h->P.qpay0 = h->qpay0.data(); h->P.qpay1 = h->qpay1.data(); h->P.qcache = h->qcache.data(); h->P.agents = h->agents.data();
  h->P.lv_price = h->lvp.data(); h->P.lv_qty = h->lvq.data(); h->P.lv_ht = h->lvht.data(); h->P.nodes = h->nodes.data(); h->P.env = h->env.data();
  h->P.trace = c.trace_cap ? h->trace.data() : nullptr; h->P.envx = h->envx.data();
  g_book_ids.erase(h);                                                 // a recycled handle address must not inherit another book's id map
  *out = h; return ABX_OK;
}
int32_t abx_book_replay(abx_sim *h, const int64_t *ops9, int64_t n_ops, void *stream) {
  (void)stream; if (!h || !h->is_env || !ops9 || n_ops < 1) return ABX_ERR_ARG;
  std::vector<int64_t> dev_ops; int rc = book_ops_to_device(ops9, n_ops, g_book_ids[h], h->st.id_orig, dev_ops); if (rc != ABX_OK) return rc;
  int n_ids = (int)h->st.id_orig.size(); bool fresh = !h->reset_done;
  if (n_ids > h->P.n_ids) {
    int new_n = n_ids + n_ids / 2 + 64; std::vector<uint4> nw((size_t)h->n_envs * new_n); memset(nw.data(), 0, nw.size() * sizeof(uint4));
    if (!fresh) for (int e = 0; e < h->n_envs; e++) memcpy(&nw[(size_t)e * new_n], &h->idtab[(size_t)e * h->P.n_ids], sizeof(uint4) * h->P.n_ids);
    std::vector<uint2> nb((size_t)h->n_envs * new_n); memset(nb.data(), 0, nb.size() * sizeof(uint2));
    if (!fresh) for (int e = 0; e < h->n_envs; e++) memcpy(&nb[(size_t)e * new_n], &h->idbook[(size_t)e * h->P.n_ids], sizeof(uint2) * h->P.n_ids);
    h->idbook.swap(nb); h->P.idbook = h->idbook.data();
    h->idtab.swap(nw); h->P.idtab = h->idtab.data(); h->P.n_ids = new_n;
  }
  for (int e = 0; e < h->n_envs; e++) {
    HostCtx ctx(h->P, e); EnvState s;
    if (fresh) { init_env_state(h->P, 0, s); s.last_trade = -1; ctx.q_clear(); } else { s = h->env[e]; s.flags &= ~ABX_F_DONE; }
    BookSimHost sim(ctx, h->P, s, e); sim.book_replay(dev_ops.data(), n_ops); h->env[e] = sim.s;
  }
  h->reset_done = true; return ABX_OK;
}
}
