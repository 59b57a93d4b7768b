/*
 * abides_oracle.c -- CPU restatement (plain C) of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 * See abides_oracle.h for the rules about who may call this and the parity status (PINNED).
 *
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): the reference is CPython + libm, every
 * fp64 operation is individually rounded.  All file:line citations are relative to /root/reference.
 */
#include "abides_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ====================================================================================================
 * growable int64 row buffers
 * ==================================================================================================== */
typedef struct { int64_t *v; int64_t n, cap; } i64buf;
static void ib_push(i64buf *b, const int64_t *row, int w) {
  if (b->n + w > b->cap) { b->cap = b->cap ? b->cap * 2 : 4096; while (b->cap < b->n + w) b->cap *= 2; b->v = (int64_t *)realloc(b->v, sizeof(int64_t) * b->cap); }
  memcpy(b->v + b->n, row, sizeof(int64_t) * w); b->n += w;
}
static const uint64_t FNV_OFF = 0xCBF29CE484222325ULL, FNV_PRIME = 0x100000001B3ULL;
static inline uint64_t fnv_mix(uint64_t h, int64_t sv) { uint64_t v = (uint64_t)sv; for (int i = 0; i < 8; i++) { h = (h ^ (v & 0xFF)) * FNV_PRIME; v >>= 8; } return h; }

/* ====================================================================================================
 * numpy legacy RandomState: MT19937 + legacy distributions (numpy/random/src/legacy, mtrand.pyx).
 * The reference draws everything through np.random.RandomState.
 * ==================================================================================================== */
struct abo_rng {
  uint32_t mt[624]; int pos; int has_gauss; double gauss; uint32_t seed;
  int record; uint8_t *tk; uint64_t *tv; int64_t tn, tcap;
  /* external tape (abo_sim_set_external): the STANDARD variates come from a supplied list instead of MT19937 -- this is how a run of
   * the CUDA simulator under its own counter-based generator is re-run here draw for draw */
  int replay; const uint8_t *rk; const uint64_t *rv; int64_t rn, rpos; int rerr;   /* rerr: 1 underrun, 2 kind mismatch */
};
static void rng_seed(abo_rng *r, uint32_t seed) { /* mt19937_seed */
  r->seed = seed;
  for (int i = 0; i < 624; i++) { r->mt[i] = seed; seed = 1812433253U * (seed ^ (seed >> 30)) + (uint32_t)i + 1U; }
  r->pos = 624; r->has_gauss = 0; r->gauss = 0.0;
}
abo_rng *abo_rng_new(uint32_t seed) { abo_rng *r = (abo_rng *)calloc(1, sizeof(abo_rng)); rng_seed(r, seed); return r; }
void abo_rng_free(abo_rng *r) { if (r) { free(r->tk); free(r->tv); free(r); } }
static void rng_rec(abo_rng *r, uint8_t k, uint64_t bits) {
  if (!r->record) return;
  if (r->tn == r->tcap) { r->tcap = r->tcap ? r->tcap * 2 : 64; r->tk = (uint8_t *)realloc(r->tk, r->tcap); r->tv = (uint64_t *)realloc(r->tv, 8 * r->tcap); }
  r->tk[r->tn] = k; r->tv[r->tn] = bits; r->tn++;
}
static inline uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static void mt_gen(abo_rng *r) {
  uint32_t *mt = r->mt; int i; uint32_t y;
  for (i = 0; i < 624 - 397; i++) { y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU); mt[i] = mt[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfU); }
  for (; i < 623; i++) { y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU); mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfU); }
  y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU); mt[623] = mt[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfU);
  r->pos = 0;
}
static inline uint32_t rng_u32_raw(abo_rng *r) {
  if (r->pos == 624) mt_gen(r);
  uint32_t y = r->mt[r->pos++];
  y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680U; y ^= (y << 15) & 0xefc60000U; y ^= (y >> 18);
  return y;
}
static inline double rng_double_raw(abo_rng *r) { /* mt19937_next_double */
  int32_t a = (int32_t)(rng_u32_raw(r) >> 5), b = (int32_t)(rng_u32_raw(r) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double rng_gauss_raw(abo_rng *r) { /* legacy_gauss */
  if (r->has_gauss) { double t = r->gauss; r->has_gauss = 0; r->gauss = 0.0; return t; }
  double f, x1, x2, r2;
  do { x1 = 2.0 * rng_double_raw(r) - 1.0; x2 = 2.0 * rng_double_raw(r) - 1.0; r2 = x1 * x1 + x2 * x2; } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(-2.0 * log(r2) / r2);
  r->gauss = f * x1; r->has_gauss = 1;
  return f * x2;
}
static inline double bits_d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t rng_replay_next(abo_rng *r, uint8_t kind) {
  if (r->rpos >= r->rn) { r->rerr |= 1; return 0; }
  if (r->rk[r->rpos] != kind) r->rerr |= 2;
  return r->rv[r->rpos++];
}
uint32_t abo_rng_u32(abo_rng *r) { return rng_u32_raw(r); }
double abo_rng_double(abo_rng *r) { if (r->replay) return bits_d(rng_replay_next(r, 'u')); double u = rng_double_raw(r); rng_rec(r, 'u', dbits(u)); return u; }
double abo_rng_gauss(abo_rng *r) { if (r->replay) return bits_d(rng_replay_next(r, 'n')); double z = rng_gauss_raw(r); rng_rec(r, 'n', dbits(z)); return z; }
double abo_rng_std_exponential(abo_rng *r) { if (r->replay) return bits_d(rng_replay_next(r, 'e')); double e = -log(1.0 - rng_double_raw(r)); rng_rec(r, 'e', dbits(e)); return e; }
int64_t abo_rng_randint(abo_rng *r, int64_t low, int64_t high) { /* _rand_int64 + random_bounded_uint64_fill, use_masked */
  if (r->replay) return low + (int64_t)rng_replay_next(r, 'i');
  uint64_t rng = (uint64_t)(high - 1 - low), v;
  if (rng == 0) v = 0;
  else if (rng <= 0xFFFFFFFFULL) {
    if (rng == 0xFFFFFFFFULL) v = rng_u32_raw(r);
    else { uint64_t mask = rng; mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
           do { v = rng_u32_raw(r) & mask; } while (v > rng); }
  } else { uint64_t mask = rng; mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
           do { uint64_t hi = rng_u32_raw(r); uint64_t lo = rng_u32_raw(r); v = ((hi << 32) | lo) & mask; } while (v > rng); }
  rng_rec(r, 'i', v);
  return low + (int64_t)v;
}
/* scaled forms, written as the identities of SURVEY App. C */
static inline double rng_normal(abo_rng *r, double loc, double scale) { return loc + scale * abo_rng_gauss(r); }
static inline double rng_exponential(abo_rng *r, double scale) { return abo_rng_std_exponential(r) * scale; }
static inline double rng_uniform(abo_rng *r, double low, double high) { return low + (high - low) * abo_rng_double(r); }

/* Python int(round(x)) for a float x: round-half-even (nearbyint under the default rounding mode). */
static inline int64_t py_round(double x) { return (int64_t)nearbyint(x); }

/* ====================================================================================================
 * Order book -- util/OrderBook.py
 * ==================================================================================================== */
typedef struct { int64_t agent_id, order_id, quantity, limit_price, fill_price; int is_buy; } order_t; /* util/order/LimitOrder.py:14-20 */
typedef struct { order_t *o; int n, cap; } level_t;          /* one price level: FIFO list, oldest at 0 */
typedef struct { level_t *lv; int n, cap; } side_t;          /* OrderBook.bids / .asks, best at 0 (OrderBook.py:24-25) */
typedef struct { int64_t t, q; } tx_t;
typedef struct { int64_t order_id, limit_price; int is_buy; tx_t *tx; int ntx, captx; } hrec_t; /* history record (OrderBook.py:52-60), only what is read back */
typedef struct { hrec_t *r; int n, cap; int64_t serial; } hbucket_t;   /* serial: which history.insert(0, {}) created it (the exchange hands out REFERENCES to these dicts, ExchangeAgent.py:276) */

typedef void (*send_fn)(void *owner, int64_t recipient, int kind, const order_t *o);

struct abo_book {
  side_t bids, asks; int64_t last_trade; int has_last_trade;
  hbucket_t *hist; int nhist; int stream_history;
  hbucket_t *arch; int n_arch, cap_arch; int keep_dropped; int64_t next_serial;   /* buckets that fell off history[:stream_history+1] but may still be referenced by an agent's stream_history */
  int64_t now; void *owner; send_fn send;
  int64_t last_update; int has_update;                     /* OrderBook.last_update_ts (:36,169,338,372): what publishOrderBookData compares with */
  i64buf notes; /* standalone use */
  int64_t n_fills;
};

static void level_push(level_t *l, const order_t *o) { if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 4; l->o = (order_t *)realloc(l->o, sizeof(order_t) * l->cap); } l->o[l->n++] = *o; }
static void level_pop_at(level_t *l, int i) { memmove(l->o + i, l->o + i + 1, sizeof(order_t) * (l->n - i - 1)); l->n--; }
static void side_insert(side_t *s, int i, const order_t *o) {
  if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->lv = (level_t *)realloc(s->lv, sizeof(level_t) * s->cap); }
  memmove(s->lv + i + 1, s->lv + i, sizeof(level_t) * (s->n - i)); s->n++;
  s->lv[i].o = NULL; s->lv[i].n = s->lv[i].cap = 0; level_push(&s->lv[i], o);
}
static void side_delete(side_t *s, int i) { free(s->lv[i].o); memmove(s->lv + i, s->lv + i + 1, sizeof(level_t) * (s->n - i - 1)); s->n--; }

static hrec_t *hist_find(hbucket_t *b, int64_t id) { for (int i = 0; i < b->n; i++) if (b->r[i].order_id == id) return &b->r[i]; return NULL; }
static void hist_add_tx(hrec_t *r, int64_t t, int64_t q) { if (r->ntx == r->captx) { r->captx = r->captx ? r->captx * 2 : 2; r->tx = (tx_t *)realloc(r->tx, sizeof(tx_t) * r->captx); } r->tx[r->ntx].t = t; r->tx[r->ntx].q = q; r->ntx++; }
static void hist_bucket_free(hbucket_t *b) { for (int i = 0; i < b->n; i++) free(b->r[i].tx); free(b->r); b->r = NULL; b->n = b->cap = 0; }

static void book_note_send(void *owner, int64_t recipient, int kind, const order_t *o) { /* standalone book: record rows */
  abo_book *b = (abo_book *)owner;
  int64_t row[13] = { b->now, recipient, kind, o->order_id, o->is_buy, o->quantity, o->limit_price, o->fill_price, 0, 0, 0, 0, 0 };
  ib_push(&b->notes, row, 13);
}
static void book_init(abo_book *b, int stream_history, void *owner, send_fn send) { /* OrderBook.__init__ :21-36 */
  memset(b, 0, sizeof(*b)); b->stream_history = stream_history; b->owner = owner; b->send = send;
  b->hist = (hbucket_t *)calloc(stream_history + 2, sizeof(hbucket_t)); b->nhist = 1;
}
static void book_destroy(abo_book *b) {
  for (int i = 0; i < b->bids.n; i++) free(b->bids.lv[i].o);
  for (int i = 0; i < b->asks.n; i++) free(b->asks.lv[i].o);
  free(b->bids.lv); free(b->asks.lv);
  for (int i = 0; i < b->nhist; i++) hist_bucket_free(&b->hist[i]);
  for (int i = 0; i < b->n_arch; i++) hist_bucket_free(&b->arch[i]);
  free(b->arch);
  free(b->hist); free(b->notes.v);
}
/* isMatch :242-254 (same-side case cannot occur: caller picks the opposite side) */
static inline int is_match(const order_t *order, const order_t *o) { return order->is_buy ? order->limit_price >= o->limit_price : order->limit_price <= o->limit_price; }
/* isBetterPrice :440-453 */
static inline int is_better(const order_t *order, const order_t *o) { return order->is_buy ? order->limit_price > o->limit_price : order->limit_price < o->limit_price; }

/* executeOrder :172-240.  Returns 1 and fills *matched when the head of the best opposite level matches. */
static int book_execute(abo_book *b, order_t *order, order_t *matched) {
  side_t *book = order->is_buy ? &b->asks : &b->bids;
  if (book->n == 0) return 0;
  if (!is_match(order, &book->lv[0].o[0])) return 0;
  if (order->quantity >= book->lv[0].o[0].quantity) {       /* :204-210 consumed the whole resting order */
    *matched = book->lv[0].o[0]; level_pop_at(&book->lv[0], 0);
    if (book->lv[0].n == 0) side_delete(book, 0);
  } else {                                                  /* :212-217 partial */
    *matched = book->lv[0].o[0]; matched->quantity = order->quantity;
    book->lv[0].o[0].quantity -= matched->quantity;
  }
  matched->fill_price = matched->limit_price;               /* :221 */
  hrec_t *r = hist_find(&b->hist[0], order->order_id);      /* :227 incoming order's PRE-fill remaining qty */
  if (r) hist_add_tx(r, b->now, order->quantity);
  for (int i = 0; i < b->nhist; i++) { hrec_t *m = hist_find(&b->hist[i], matched->order_id); if (m) hist_add_tx(m, b->now, matched->quantity); } /* :230-237 */
  return 1;
}
/* enterOrder :256-282 */
static void book_enter(abo_book *b, const order_t *order) {
  side_t *book = order->is_buy ? &b->bids : &b->asks;
  if (book->n == 0) { side_insert(book, 0, order); return; }
  const order_t *last = &book->lv[book->n - 1].o[0];
  if (!is_better(order, last) && order->limit_price != last->limit_price) { side_insert(book, book->n, order); return; }
  for (int i = 0; i < book->n; i++) {
    if (is_better(order, &book->lv[i].o[0])) { side_insert(book, i, order); return; }
    else if (order->limit_price == book->lv[i].o[0].limit_price) { level_push(&book->lv[i], order); return; }
  }
}
/* handleLimitOrder :38-170 (symbol check is the caller's; book_log/prettyPrint have no state effect) */
static void book_handle_limit(abo_book *b, order_t order) {
  if (order.quantity <= 0) return;                                                   /* :47-49 */
  hbucket_t *h0 = &b->hist[0]; hrec_t *r = hist_find(h0, order.order_id);            /* :52-60 */
  if (r) { r->ntx = 0; r->limit_price = order.limit_price; r->is_buy = order.is_buy; }
  else { if (h0->n == h0->cap) { h0->cap = h0->cap ? h0->cap * 2 : 16; h0->r = (hrec_t *)realloc(h0->r, sizeof(hrec_t) * h0->cap); }
         h0->r[h0->n].order_id = order.order_id; h0->r[h0->n].limit_price = order.limit_price; h0->r[h0->n].is_buy = order.is_buy;
         h0->r[h0->n].tx = NULL; h0->r[h0->n].ntx = h0->r[h0->n].captx = 0; h0->n++; }
  int matching = 1; int64_t trade_qty = 0, trade_price = 0; int executed = 0;
  while (matching) {                                                                 /* :68-110 */
    order_t matched;
    if (book_execute(b, &order, &matched)) {
      order_t filled = order; filled.quantity = matched.quantity; filled.fill_price = matched.fill_price; /* :73-75 */
      order.quantity -= filled.quantity;                                             /* :77 */
      b->send(b->owner, order.agent_id, ABO_ORDER_EXECUTED, &filled);                /* :88 */
      b->send(b->owner, matched.agent_id, ABO_ORDER_EXECUTED, &matched);             /* :89-91 */
      trade_qty += filled.quantity; trade_price += filled.fill_price * filled.quantity; executed = 1; b->n_fills++; /* :94,131-137 */
      if (order.quantity <= 0) matching = 0;
    } else {
      book_enter(b, &order);                                                         /* :101 */
      b->send(b->owner, order.agent_id, ABO_ORDER_ACCEPTED, &order);                 /* :108 */
      matching = 0;
    }
  }
  if (executed) {                                                                    /* :131-149 */
    b->last_trade = py_round((double)trade_price / (double)trade_qty); b->has_last_trade = 1; /* int(round(a / b)): true division */
    /* history.insert(0, {}) then truncate to stream_history+1 */
    if (b->nhist == b->stream_history + 1) {
      if (b->keep_dropped) { if (b->n_arch == b->cap_arch) { b->cap_arch = b->cap_arch ? 2 * b->cap_arch : 64; b->arch = (hbucket_t *)realloc(b->arch, sizeof(hbucket_t) * b->cap_arch); } b->arch[b->n_arch++] = b->hist[b->nhist - 1]; }
      else hist_bucket_free(&b->hist[b->nhist - 1]);
      b->nhist--; }
    memmove(b->hist + 1, b->hist, sizeof(hbucket_t) * b->nhist); memset(&b->hist[0], 0, sizeof(hbucket_t)); b->hist[0].serial = ++b->next_serial; b->nhist++;
  }
  b->last_update = b->now; b->has_update = 1;                                        /* :169 */
}
/* cancelOrder :284-339 */
static void book_cancel(abo_book *b, const order_t *order) {
  side_t *book = order->is_buy ? &b->bids : &b->asks;
  if (book->n == 0) return;
  for (int i = 0; i < book->n; i++) {
    if (order->limit_price == book->lv[i].o[0].limit_price) {                        /* :306 level price is slot 0's */
      for (int ci = 0; ci < book->lv[i].n; ci++) {
        if (order->order_id == book->lv[i].o[ci].order_id) {
          order_t cancelled = book->lv[i].o[ci]; level_pop_at(&book->lv[i], ci);     /* :311 */
          /* :314-321 history cancellations: never read back on this path */
          if (book->lv[i].n == 0) side_delete(book, i);                              /* :324-325 */
          b->send(b->owner, order->agent_id, ABO_ORDER_CANCELLED, &cancelled);       /* :334-336 recipient = REQUEST's agent_id */
          b->last_update = b->now; b->has_update = 1;                                /* :338 only when the order was found */
          return;
        }
      }
    }
  }
}
/* modifyOrder :341-372 -- including the slot-0 overwrite and the live iteration */
static void book_modify(abo_book *b, const order_t *order, const order_t *new_order) {
  if (order->order_id != new_order->order_id) return;                                /* :343 */
  side_t *book = order->is_buy ? &b->bids : &b->asks;
  if (book->n == 0) return;
  for (int i = 0; i < book->n; i++) {
    if (order->limit_price == book->lv[i].o[0].limit_price) {                        /* :349 evaluated once per level */
      for (int mi = 0; mi < book->lv[i].n; mi++) {
        if (order->order_id == book->lv[i].o[mi].order_id) {                         /* :351 sees the overwritten slot 0 on later mi */
          book->lv[i].o[0] = *new_order;                                             /* :352 */
          for (int idx = 0; idx < b->nhist; idx++) {                                 /* :353-367 one ORDER_MODIFIED per bucket holding the id */
            if (!hist_find(&b->hist[idx], new_order->order_id)) continue;
            b->send(b->owner, order->agent_id, ABO_ORDER_MODIFIED, new_order);
          }
        }
      }
    }
  }
  b->last_update = b->now; b->has_update = 1;                                        /* :372 */
}
static const hbucket_t *book_bucket(const abo_book *b, int64_t serial) {
  for (int i = 0; i < b->nhist; i++) if (b->hist[i].serial == serial) return &b->hist[i];
  for (int i = b->n_arch - 1; i >= 0; i--) if (b->arch[i].serial == serial) return &b->arch[i];
  return NULL;
}
/* getInsideBids / getInsideAsks :377-398 */
static int book_inside(const abo_book *b, int is_bid, int depth, int64_t *out) {
  const side_t *s = is_bid ? &b->bids : &b->asks; int n = depth < s->n ? depth : s->n;
  for (int i = 0; i < n; i++) { int64_t q = 0; for (int k = 0; k < s->lv[i].n; k++) q += s->lv[i].o[k].quantity; out[2 * i] = s->lv[i].o[0].limit_price; out[2 * i + 1] = q; }
  return n;
}
/* get_transacted_volume :400-436: distinct (time, qty) tuples over the surviving history buckets, time >= now - lookback */
static int64_t book_transacted_volume(const abo_book *b, int64_t lookback) {
  tx_t *all = NULL; int n = 0, cap = 0;
  for (int i = 0; i < b->nhist; i++) for (int k = 0; k < b->hist[i].n; k++) for (int j = 0; j < b->hist[i].r[k].ntx; j++) {
    tx_t t = b->hist[i].r[k].tx[j]; int dup = 0;
    for (int m = 0; m < n; m++) if (all[m].t == t.t && all[m].q == t.q) { dup = 1; break; }
    if (dup) continue;
    if (n == cap) { cap = cap ? cap * 2 : 64; all = (tx_t *)realloc(all, sizeof(tx_t) * cap); }
    all[n++] = t;
  }
  int64_t start = b->now - lookback, sum = 0;
  for (int m = 0; m < n; m++) if (all[m].t >= start) sum += all[m].q;
  free(all); return sum;
}

/* ---- standalone book API ---- */
abo_book *abo_book_new(int stream_history) { abo_book *b = (abo_book *)malloc(sizeof(abo_book)); book_init(b, stream_history, b, book_note_send); return b; }
void abo_book_free(abo_book *b) { if (b) { book_destroy(b); free(b); } }
void abo_book_set_time(abo_book *b, int64_t t) { b->now = t; }
void abo_book_limit(abo_book *b, int64_t agent, int64_t id, int is_buy, int64_t price, int64_t qty) { order_t o = { agent, id, qty, price, 0, is_buy }; book_handle_limit(b, o); }
void abo_book_cancel(abo_book *b, int64_t agent, int64_t id, int is_buy, int64_t price) { order_t o = { agent, id, 0, price, 0, is_buy }; book_cancel(b, &o); }
void abo_book_modify(abo_book *b, int64_t agent, int64_t id, int is_buy, int64_t price, int64_t nid, int64_t nprice, int64_t nqty) {
  order_t o = { agent, id, 0, price, 0, is_buy }, n = { agent, nid, nqty, nprice, 0, is_buy }; book_modify(b, &o, &n);
}
int abo_book_inside(abo_book *b, int is_bid, int depth, int64_t *out) { return book_inside(b, is_bid, depth, out); }
int64_t abo_book_last_trade(abo_book *b) { return b->has_last_trade ? b->last_trade : -1; }
int64_t abo_book_transacted_volume(abo_book *b, int64_t lookback) { return book_transacted_volume(b, lookback); }
int abo_book_n_levels(abo_book *b, int is_bid) { return is_bid ? b->bids.n : b->asks.n; }
int abo_book_n_resting(abo_book *b) { int n = 0; for (int i = 0; i < b->bids.n; i++) n += b->bids.lv[i].n; for (int i = 0; i < b->asks.n; i++) n += b->asks.lv[i].n; return n; }
int64_t abo_book_n_notes(abo_book *b) { return b->notes.n / 13; }
const int64_t *abo_book_notes(abo_book *b) { return b->notes.v; }
void abo_book_clear_notes(abo_book *b) { b->notes.n = 0; }
int abo_book_level_orders(abo_book *b, int is_bid, int level, int64_t *out, int max_orders) {
  side_t *s = is_bid ? &b->bids : &b->asks; if (level >= s->n) return 0;
  int n = s->lv[level].n < max_orders ? s->lv[level].n : max_orders;
  for (int k = 0; k < n; k++) { out[3 * k] = s->lv[level].o[k].order_id; out[3 * k + 1] = s->lv[level].o[k].quantity; out[3 * k + 2] = s->lv[level].o[k].limit_price; }
  return n;
}

/* ====================================================================================================
 * Simulation: Kernel + ExchangeAgent + TradingAgent/ZeroIntelligenceAgent + SparseMeanRevertingOracle
 * ==================================================================================================== */
typedef struct {            /* one PriorityQueue entry: (deliverAt, (recipient, type, Message)) Kernel.py:425,462 */
  int64_t t; int32_t recipient; int32_t type; int64_t uniq;  /* Message.uniq (message/Message.py:33-34); -1 for msg=None */
  int32_t kind; int32_t sender;
  order_t order;                                             /* body["order"] */
  int64_t bid, bid_q, ask, ask_q, data; int mkt_closed;      /* QUERY_SPREAD reply (depth-1 lists; has_* by qty>0), body["data"] */
  int has_bid, has_ask;
  int64_t bid2, ask2; int n_bids, n_asks;                    /* depth > 1 replies: second level price, level counts (capped at 2) */
  order_t new_order;                                         /* MODIFY_ORDER body["new_order"] */
  int64_t tv, lookback;                                      /* QUERY_TRANSACTED_VOLUME: transacted_volume / lookback_period (ns); QUERY_ORDER_STREAM: lookback = length */
  int n_stream; int64_t stream_serial[16];                   /* QUERY_ORDER_STREAM reply: "orders" = history[1 : length + 1], as references to the book's dicts */
  int md_nb, md_na; int64_t md_bids[10], md_asks[10]; int has_data;   /* MARKET_DATA: `levels` (price, quantity) pairs a side, last_transaction (None before the first trade) */
  int sub_levels; double sub_freq;                           /* MARKET_DATA_SUBSCRIPTION_REQUEST body */
} event_t;

static inline int ev_less(const event_t *a, const event_t *b) { /* tuple order (t, recipient, type.value, msg.uniq) */
  if (a->t != b->t) return a->t < b->t;
  if (a->recipient != b->recipient) return a->recipient < b->recipient;
  if (a->type != b->type) return a->type < b->type;
  return a->uniq < b->uniq;
}
typedef struct { event_t *e; int n, cap; } heap_t;
static void heap_push(heap_t *h, const event_t *ev) { /* heapq.heappush */
  if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 4096; h->e = (event_t *)realloc(h->e, sizeof(event_t) * h->cap); }
  int i = h->n++; h->e[i] = *ev;
  while (i > 0) { int p = (i - 1) >> 1; if (!ev_less(&h->e[i], &h->e[p])) break; event_t t = h->e[i]; h->e[i] = h->e[p]; h->e[p] = t; i = p; }
}
static void heap_pop(heap_t *h, event_t *out) { /* heapq.heappop */
  *out = h->e[0]; h->n--; if (h->n == 0) return; h->e[0] = h->e[h->n];
  int i = 0; for (;;) { int l = 2 * i + 1, r = l + 1, m = i; if (l < h->n && ev_less(&h->e[l], &h->e[m])) m = l; if (r < h->n && ev_less(&h->e[r], &h->e[m])) m = r; if (m == i) break; event_t t = h->e[i]; h->e[i] = h->e[m]; h->e[m] = t; i = m; }
}

typedef struct { int64_t order_id, quantity, limit_price; int is_buy; } open_order_t;
enum { ST_AWAITING_WAKEUP = 0, ST_INACTIVE, ST_AWAITING_SPREAD, ST_AWAITING_TV_, ST_AWAITING_STREAM, ST_AWAITING_MARKET_DATA };

typedef struct {            /* TradingAgent (agent/TradingAgent.py:19-98) + ZeroIntelligenceAgent (:65-70) state */
  abo_rng *rs; int group;
  int64_t R_min, R_max; double eta;
  int64_t shares, cash, starting_cash;                  /* holdings[symbol], holdings["CASH"] */
  int has_open, has_close; int64_t mkt_open, mkt_close; /* mkt_open/mkt_close None until the replies arrive */
  int mkt_closed, first_wake, trading, state;
  int has_last_trade, has_daily_close; int64_t last_trade, daily_close;
  int has_known; int64_t bid, bid_q, ask, ask_q; int has_bid, has_ask;
  open_order_t *orders; int n_orders, cap_orders;       /* self.orders dict, insertion ordered */
  double r_t, sigma_t; int has_prev; int64_t prev_wake;
  int32_t theta[64]; int q_max;
  int64_t current_time;                                 /* Agent.currentTime */
  int64_t surplus;
  /* rmsc03 population (config/rmsc03.py): agent class and its extra state */
  int type;                                             /* AT_ZI, AT_NOISE, AT_VALUE, AT_MOMENTUM, AT_POVMM */
  int64_t wakeup_time, size;                            /* NoiseAgent.wakeup_time[0]; NoiseAgent/ValueAgent/MomentumAgent.size */
  double *mids; int n_mids, cap_mids; double avg20, avg50; int has20, has50;   /* MomentumAgent.mid_list, avg_20_list[-1], avg_50_list[-1] */
  int64_t order_size, last_mid, transacted_volume; int has_last_mid, aw_spread, aw_vol;   /* POVMarketMakerAgent */
  int64_t *kside[2], *fside[2]; int nk[2], nf[2], capk; int64_t px_rem, px_executed, px_n_executed;   /* POVExecutionAgent: known_bids/asks (price, qty pairs), lists in flight, rem_quantity */
  int L, n_stream; int64_t stream_serial[16];           /* HeuristicBeliefLearningAgent: L, stream_history[symbol] (references to the exchange's history dicts) */
  int subscribe, sub_requested; int kb_n, ka_n; int64_t kb[10], ka[10];   /* subscription mode: known_bids / known_asks as MARKET_DATA delivered them */
  int64_t mkm_min, mkm_max, last_spread;                /* MarketMakerAgent (agent/market_makers/MarketMakerAgent.py): min_size, max_size, last_spread (10, never updated) */
} zi_t;
enum { AT_ZI = 0, AT_NOISE, AT_VALUE, AT_MOMENTUM, AT_POVMM, AT_POVEXEC, AT_MKM, AT_HBL };

struct abo_sim {
  int variant; uint32_t seed; int trace; int n_agents;
  /* kernel */
  heap_t q; int64_t now, start_time, stop_time; int64_t *agent_time; int64_t *comp_delay; int64_t addl_delay;
  int64_t ttl; int64_t uniq; int64_t next_order_id; int started, loop_done;
  abo_rng *g, *kernel_rs, *lat_rs, *sym_rs, *exch_rs;
  int use_latency_model; double *latency; int n_noise; double jitter, jitter_clip, jitter_unit;
  /* exchange */
  abo_book book; int64_t mkt_open, mkt_close, pipeline_delay, exch_comp_delay;
  /* oracle (SparseMeanRevertingOracle) */
  double r_bar, kappa, fund_vol, megashock_lambda, megashock_mean, megashock_var;
  int64_t or_t, or_v; int64_t ms_t; double ms_v; double *gexp; int64_t n_gexp, cap_gexp;
  /* agents */
  zi_t *zi; double sigma_n, agent_kappa, sigma_s, lambda_a; int64_t order_size, starting_cash; double value_percent_aggr; int64_t value_depth_spread;
  double mm_pov; int64_t mm_min_size, mm_window, mm_ticks, mm_wake_ns, mom_wake_ns;   /* config/rmsc03.py:41-45,176-200 */
  struct { int agent, levels; double freq; int64_t last; } subs[128]; int n_subs;      /* ExchangeAgent.subscription_dict in insertion order */
  double mkm_sub_freq, mom_sub_freq;
  int64_t mkm_levels, mkm_wake_ns;                                                     /* MarketMakerAgent subscribe_num_levels (5), wake_up_freq ("1s") */
  int px_kind; int64_t px_limit;                        /* 0 POVExecutionAgent, 1 PassiveAgent (px_limit: its limit_price, 0 = None), 2 AggressiveAgent */
  int px_id, px_is_buy; double px_pov; int64_t px_quantity, px_start, px_end, px_freq, px_lookback;   /* POVExecutionAgent (agent/execution/baselines/pov_agent.py), 0 = none */
  /* traces */
  i64buf live_qty;                                      /* by order id: quantity of the AGENT's order object (a CANCEL_ORDER message carries a reference to it, TradingAgent.py:399-406: partial fills the agent books while the message is in flight show in what the exchange receives) */
  i64buf pops, ops, notes, snaps; uint64_t pop_hash, note_hash, snap_hash; uint64_t *ckpt; int64_t n_ckpt, cap_ckpt;
  int64_t c_limit, c_cancel, c_query, max_queue, max_bid_lv, max_ask_lv, max_resting;
};

/* ---------------- kernel services ---------------- */
static void k_put(abo_sim *s, const event_t *e) { heap_push(&s->q, e); if (s->q.n > s->max_queue) s->max_queue = s->q.n; }

/* Kernel.setWakeup :435-462 */
static void k_set_wakeup(abo_sim *s, int sender, int64_t t) {
  event_t e; memset(&e, 0, sizeof(e)); e.t = t; e.recipient = sender; e.type = ABO_T_WAKEUP; e.uniq = -1; k_put(s, &e);
}
/* Message() construction: message/Message.py:28-34 */
static inline int64_t new_uniq(abo_sim *s) { return s->uniq++; }

/* Kernel.sendMessage :347-433.  `e` carries the body and e->uniq from Message construction. */
static void k_send(abo_sim *s, int sender, int recipient, event_t *e, int64_t delay) {
  int64_t sent = s->now + (s->comp_delay[sender] + s->addl_delay + delay);                 /* :391-393 */
  int64_t deliver;
  if (s->use_latency_model) {                                                              /* :397-399, model/LatencyModel.py:109-140 */
    double min_latency = s->latency[(size_t)sender * s->n_agents + recipient];
    double x = rng_uniform(s->lat_rs, s->jitter_clip, 1.0);
    double latency = min_latency + ((s->jitter / pow(x, 3.0)) * (min_latency / s->jitter_unit));
    deliver = sent + (int64_t)latency;                                                     /* pd.Timedelta(float) truncates */
  } else {                                                                                 /* :410-412 */
    double latency = s->latency[(size_t)sender * s->n_agents + recipient];
    int64_t noise = abo_rng_randint(s->kernel_rs, 0, s->n_noise);                          /* choice(len, 1, <list as replace>) == randint(0,len) */
    deliver = sent + (int64_t)(latency + (double)noise);
  }
  e->t = deliver; e->recipient = recipient; e->type = ABO_T_MESSAGE; e->sender = sender;
  k_put(s, e);                                                                             /* :425 */
}

/* ---------------- SparseMeanRevertingOracle ---------------- */
static inline int64_t ns_from_float_string(double x) { return (int64_t)x; } /* pd.Timedelta("{}ns".format(float)) truncates */
static double g_exponential(abo_sim *s, double scale) { /* np.random.exponential on the GLOBAL stream :69,168 */
  double e = abo_rng_std_exponential(s->g);
  if (s->n_gexp == s->cap_gexp) { s->cap_gexp = s->cap_gexp ? s->cap_gexp * 2 : 32; s->gexp = (double *)realloc(s->gexp, 8 * s->cap_gexp); }
  s->gexp[s->n_gexp++] = e;
  return e * scale;
}
static void oracle_new_megashock(abo_sim *s, int64_t from) { /* :67-73, :168-171 */
  s->ms_t = from + ns_from_float_string(g_exponential(s, 1.0 / s->megashock_lambda));
  double msv = rng_normal(s->sym_rs, s->megashock_mean, sqrt(s->megashock_var));
  s->ms_v = abo_rng_randint(s->sym_rs, 0, 2) == 0 ? msv : -msv;
}
/* compute_fundamental_at_timestamp :88-125 */
static int64_t oracle_compute(abo_sim *s, int64_t ts, double v_adj, int64_t pt, int64_t pv) {
  int64_t d = ts - pt; double mu = s->r_bar, gamma = s->kappa, theta = s->fund_vol;
  double v = rng_normal(s->sym_rs, mu + ((double)pv - mu) * exp(-gamma * (double)d),
                        ((theta * theta) / (2 * gamma)) * (1 - exp(-2 * gamma * (double)d)));   /* variance formula passed as scale */
  v += v_adj; if (!(v > 0)) v = 0;                                                         /* max(0, v) */
  int64_t iv = py_round(v); s->or_t = ts; s->or_v = iv; return iv;
}
/* advance_fundamental_value_series :131-181 */
static int64_t oracle_advance(abo_sim *s, int64_t t) {
  int64_t pt = s->or_t, pv = s->or_v;
  if (t <= pt) return pv;
  while (s->ms_t < t) { int64_t v = oracle_compute(s, s->ms_t, s->ms_v, pt, pv); pt = s->ms_t; pv = v; oracle_new_megashock(s, pt); }
  return oracle_compute(s, t, 0.0, pt, pv);
}
/* observePrice :210-227 */
static int64_t oracle_observe(abo_sim *s, int64_t t, double sigma_n, abo_rng *rs) {
  int64_t r_t = (t >= s->mkt_close) ? oracle_advance(s, s->mkt_close - 1) : oracle_advance(s, t);
  if (sigma_n == 0) return r_t;
  return py_round(rng_normal(rs, (double)r_t, sqrt(sigma_n)));
}

/* ---------------- exchange ---------------- */
static void trace_note(abo_sim *s, int recipient, const event_t *e) {
  int64_t row[13] = { s->now, recipient, e->kind, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  if (e->kind == ABO_ORDER_ACCEPTED || e->kind == ABO_ORDER_EXECUTED || e->kind == ABO_ORDER_CANCELLED || e->kind == ABO_ORDER_MODIFIED) {
    row[3] = e->order.order_id; row[4] = e->order.is_buy; row[5] = e->order.quantity; row[6] = e->order.limit_price; row[7] = e->order.fill_price;
  }
  if (e->kind == ABO_QUERY_SPREAD) { row[7] = e->data; if (e->has_bid) { row[8] = e->bid; row[9] = e->bid_q; } if (e->has_ask) { row[10] = e->ask; row[11] = e->ask_q; } row[12] = e->mkt_closed; }
  if (e->kind == ABO_MARKET_DATA) {                          /* the recorder's row: level counts, position-weighted price / size sums, top of book, last trade */
    row[3] = e->md_nb; row[4] = e->md_na; row[7] = e->has_data ? e->data : -1;
    for (int i = 0; i < e->md_nb; i++) row[5] += (i + 1) * e->md_bids[2 * i];
    for (int i = 0; i < e->md_na; i++) row[6] += (i + 1) * e->md_asks[2 * i];
    if (e->md_nb) { row[8] = e->md_bids[0]; row[9] = e->md_bids[1]; }
    if (e->md_na) { row[10] = e->md_asks[0]; row[11] = e->md_asks[1]; }
    for (int i = 0; i < e->md_nb && i < e->md_na; i++) row[12] += (i + 1) * (e->md_bids[2 * i + 1] + 3 * e->md_asks[2 * i + 1]);
  }
  for (int i = 0; i < 13; i++) s->note_hash = fnv_mix(s->note_hash, row[i]);
  if (s->trace & ABO_TRACE_NOTES) ib_push(&s->notes, row, 13);
}
/* ExchangeAgent.sendMessage :471-485 */
static void exch_send(abo_sim *s, int recipient, event_t *e) {
  e->uniq = new_uniq(s);
  trace_note(s, recipient, e);
  int64_t delay = (e->kind == ABO_ORDER_ACCEPTED || e->kind == ABO_ORDER_CANCELLED || e->kind == ABO_ORDER_EXECUTED) ? s->pipeline_delay : 0;
  k_send(s, 0, recipient, e, delay);
}
static void exch_book_send(void *owner, int64_t recipient, int kind, const order_t *o) { /* book -> owner.sendMessage */
  abo_sim *s = (abo_sim *)owner; event_t e; memset(&e, 0, sizeof(e)); e.kind = kind; e.order = *o; exch_send(s, (int)recipient, &e);
}
static void trace_snap(abo_sim *s) {
  int64_t row[16]; memset(row, 0, sizeof(row));
  int nb = s->book.bids.n, na = s->book.asks.n; int64_t rest = abo_book_n_resting(&s->book);
  row[0] = nb; row[1] = na; row[2] = rest;
  book_inside(&s->book, 1, 3, row + 3); book_inside(&s->book, 0, 3, row + 9);
  row[15] = s->book.has_last_trade ? s->book.last_trade : -1;
  if (nb > s->max_bid_lv) s->max_bid_lv = nb;
  if (na > s->max_ask_lv) s->max_ask_lv = na;
  if (rest > s->max_resting) s->max_resting = rest;
  for (int i = 0; i < 16; i++) s->snap_hash = fnv_mix(s->snap_hash, row[i]);
  if (s->trace & ABO_TRACE_SNAPS) ib_push(&s->snaps, row, 16);
}
static void trace_op(abo_sim *s, int op, const order_t *o, int64_t np, int64_t nq) {
  if (!(s->trace & ABO_TRACE_OPS)) return;
  int64_t row[9] = { s->now, op, o->agent_id, o->order_id, o->is_buy, o->limit_price, o->quantity, np, nq }; ib_push(&s->ops, row, 9);
}
/* ExchangeAgent.publishOrderBookData :359-387 */
static void exch_publish(abo_sim *s) {
  for (int k = 0; k < s->n_subs; k++) {
    int64_t lu = s->book.last_update, last = s->subs[k].last;
    if (!(s->subs[k].freq == 0 || (lu > last && (double)(lu - last) >= s->subs[k].freq))) continue;
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_MARKET_DATA;
    int lv = s->subs[k].levels > 5 ? 5 : s->subs[k].levels;
    e.md_nb = book_inside(&s->book, 1, lv, e.md_bids); e.md_na = book_inside(&s->book, 0, lv, e.md_asks);
    e.data = s->book.last_trade; e.has_data = s->book.has_last_trade;
    exch_send(s, s->subs[k].agent, &e);
    s->subs[k].last = lu;
  }
}
/* ExchangeAgent.receiveMessage :129-340 */
static void exch_receive(abo_sim *s, const event_t *m) {
  s->comp_delay[0] = s->exch_comp_delay;                                                    /* :139 */
  int t_closed = s->now > s->mkt_close;
  if (t_closed) {                                                                           /* :142-160 */
    int is_order = m->kind == ABO_LIMIT_ORDER || m->kind == ABO_CANCEL_ORDER || m->kind == ABO_MODIFY_ORDER;
    int is_query = m->kind == ABO_QUERY_SPREAD || m->kind == ABO_QUERY_LAST_TRADE || m->kind == ABO_QUERY_TRANSACTED_VOLUME || m->kind == ABO_QUERY_ORDER_STREAM;
    if (is_order || !is_query) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_MKT_CLOSED; exch_send(s, m->sender, &e); return; }
  }
  if (m->kind == ABO_MARKET_DATA_SUBSCRIPTION_REQUEST) {                                   /* updateSubscriptionDict :342-357: dict[agent] = {symbol: [levels, freq, now]} (a repeat keeps its place) */
    int k = 0; while (k < s->n_subs && s->subs[k].agent != m->sender) k++;
    if (k == s->n_subs) { if (s->n_subs == 128) { fprintf(stderr, "abides_oracle: more than 128 subscribers\n"); return; } s->n_subs++; }
    s->subs[k].agent = m->sender; s->subs[k].levels = m->sub_levels; s->subs[k].freq = m->sub_freq; s->subs[k].last = s->now;
    return;
  }
  event_t e; memset(&e, 0, sizeof(e));
  switch (m->kind) {
    case ABO_QUERY_TRANSACTED_VOLUME: s->book.now = s->now; e.kind = ABO_QUERY_TRANSACTED_VOLUME; e.tv = book_transacted_volume(&s->book, m->lookback); e.mkt_closed = t_closed; exch_send(s, m->sender, &e); break; /* :280-303 */
    case ABO_QUERY_ORDER_STREAM: {                                                          /* :251-279: orders = history[1 : length + 1] */
      e.kind = ABO_QUERY_ORDER_STREAM; e.mkt_closed = t_closed; e.lookback = m->lookback;
      for (int k = 1; k <= (int)m->lookback && k < s->book.nhist && e.n_stream < 16; k++) e.stream_serial[e.n_stream++] = s->book.hist[k].serial;
      exch_send(s, m->sender, &e); break; }
    case ABO_WHEN_MKT_OPEN: s->comp_delay[0] = 0; e.kind = ABO_WHEN_MKT_OPEN; e.data = s->mkt_open; exch_send(s, m->sender, &e); break;     /* :175-183 */
    case ABO_WHEN_MKT_CLOSE: s->comp_delay[0] = 0; e.kind = ABO_WHEN_MKT_CLOSE; e.data = s->mkt_close; exch_send(s, m->sender, &e); break;  /* :184-192 */
    case ABO_QUERY_SPREAD: {                                                                /* :215-245, depth 1 on this path */
      int64_t pq[2]; s->c_query++;
      e.kind = ABO_QUERY_SPREAD;
      if (book_inside(&s->book, 1, 1, pq)) { e.has_bid = 1; e.bid = pq[0]; e.bid_q = pq[1]; }
      if (book_inside(&s->book, 0, 1, pq)) { e.has_ask = 1; e.ask = pq[0]; e.ask_q = pq[1]; }
      if (s->px_id && m->sender == s->px_id) {                                                /* depth = sys.maxsize: the whole book rides in the reply */
        zi_t *a = &s->zi[s->px_id]; int need = s->book.bids.n > s->book.asks.n ? s->book.bids.n : s->book.asks.n;
        if (need > a->capk) { a->capk = need * 2 + 64; for (int k = 0; k < 2; k++) { a->kside[k] = (int64_t *)realloc(a->kside[k], 16 * a->capk); a->fside[k] = (int64_t *)realloc(a->fside[k], 16 * a->capk); } }
        int depth = s->px_kind == 2 ? 100 : a->capk;                                          /* AggressiveAgent: getCurrentSpread(depth=100) */
        a->nf[0] = book_inside(&s->book, 1, depth, a->fside[0]); a->nf[1] = book_inside(&s->book, 0, depth, a->fside[1]); }
      e.data = s->book.last_trade; e.mkt_closed = t_closed; exch_send(s, m->sender, &e); break; }
    case ABO_LIMIT_ORDER: s->c_limit++; trace_op(s, 0, &m->order, 0, 0); s->book.now = s->now; book_handle_limit(&s->book, m->order); trace_snap(s); exch_publish(s); break; /* :304-312 */
    case ABO_CANCEL_ORDER: s->c_cancel++; { order_t seen = m->order; if (seen.order_id < s->live_qty.n && s->live_qty.v[seen.order_id] > 0) seen.quantity = s->live_qty.v[seen.order_id]; trace_op(s, 1, &seen, 0, 0); } s->book.now = s->now; book_cancel(&s->book, &m->order); trace_snap(s); exch_publish(s); break;   /* :313-325 */
    default: break;
  }
}

/* ---------------- trading agent / ZI ---------------- */
static void ta_send(abo_sim *s, int id, event_t *e) { e->uniq = new_uniq(s); k_send(s, id, 0, e, 0); } /* Agent.sendMessage :148-149 */

/* TradingAgent.getCurrentSpread :277-282 -- builds a second, never-sent Message (uniq += 2) */
static void ta_get_spread(abo_sim *s, int id) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_SPREAD; ta_send(s, id, &e); (void)new_uniq(s); }

/* ZeroIntelligenceAgent.wakeup :125-187 (+ TradingAgent.wakeup :142-158) */
static void zi_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id]; a->current_time = s->now; a->first_wake = 0;
  if (!a->has_open) {                                                                       /* TradingAgent.py:149-153 */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_OPEN; ta_send(s, id, &e);
    memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_CLOSE; ta_send(s, id, &e);
  }
  a->state = ST_INACTIVE;
  if (!a->has_open || !a->has_close) return;                                                /* :130-131 */
  a->trading = 1;
  if (a->mkt_closed && a->has_daily_close) return;                                          /* :145-147 */
  double delta_time = rng_exponential(a->rs, 1.0 / s->lambda_a);                            /* :157 */
  k_set_wakeup(s, id, s->now + py_round(delta_time));                                       /* :158 */
  if (a->mkt_closed && !a->has_daily_close) { ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD; return; } /* :162-166 */
  for (int i = 0; i < a->n_orders; i++) {                                                   /* cancelOrders :336-344 -> TradingAgent.cancelOrder :399-406 */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_CANCEL_ORDER;
    e.order.agent_id = id; e.order.order_id = a->orders[i].order_id; e.order.quantity = a->orders[i].quantity; e.order.limit_price = a->orders[i].limit_price; e.order.is_buy = a->orders[i].is_buy;
    ta_send(s, id, &e);
  }
  if (a->type == AT_HBL) {                                                                  /* :183-187 a subclass is left "ACTIVE"; HeuristicBeliefLearningAgent.wakeup :61-74 */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_ORDER_STREAM; e.lookback = a->L; ta_send(s, id, &e); a->state = ST_AWAITING_STREAM; return; }
  ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD;                                      /* :183-185 */
}
/* ZeroIntelligenceAgent.updateEstimates :189-275 */
static int zi_update_estimates(abo_sim *s, int id, int64_t *v_out, int *buy_out) {
  zi_t *a = &s->zi[id];
  int64_t obs_t = oracle_observe(s, a->current_time, s->sigma_n, a->rs);                    /* :190-192 */
  int64_t q = (int64_t)((double)a->shares / 100);                                           /* :203 int(x / 100) */
  int buy;
  if (q >= a->q_max) buy = 0; else if (q <= -a->q_max) buy = 1; else buy = (int)abo_rng_randint(a->rs, 0, 2); /* :205-213 */
  if (!a->has_prev) { a->prev_wake = a->mkt_open; a->has_prev = 1; }                         /* :217-218 */
  double kappa = s->agent_kappa, r_bar = s->r_bar, sigma_n = s->sigma_n;
  double delta = (double)(a->current_time - a->prev_wake);                                  /* :221 */
  double pw = pow(1 - kappa, delta);
  double r_tprime = (1 - pw) * r_bar;                                                       /* :229 */
  r_tprime += pw * a->r_t;                                                                  /* :230 */
  double pw2 = pow(1 - kappa, 2 * delta);
  double sigma_tprime = pw2 * a->sigma_t;                                                   /* :233 */
  sigma_tprime += ((1 - pw2) / (1 - pow(1 - kappa, 2.0))) * s->sigma_s;                     /* :234 */
  a->r_t = (sigma_n / (sigma_n + sigma_tprime)) * r_tprime;                                 /* :239 */
  a->r_t += (sigma_tprime / (sigma_n + sigma_tprime)) * (double)obs_t;                      /* :240 */
  a->sigma_t = (sigma_n * a->sigma_t) / (sigma_n + a->sigma_t);                             /* :242 */
  double d2 = (double)(a->mkt_close - a->current_time); if (!(d2 > 0)) d2 = 0;              /* :251 max(0, .) */
  double pw3 = pow(1 - kappa, d2);
  double r_T = (1 - pw3) * r_bar;                                                           /* :255 */
  r_T += pw3 * a->r_t;                                                                      /* :256 */
  int64_t r_Ti = py_round(r_T);                                                             /* :259 */
  a->prev_wake = a->current_time;                                                           /* :262 */
  q += a->q_max - 1;                                                                        /* :267 */
  int64_t idx = buy ? q + 1 : q; int n = 2 * a->q_max;                                      /* :268 Python list indexing */
  if (idx < 0) idx += n;
  if (idx < 0 || idx >= n) { fprintf(stderr, "abides_oracle: theta index out of range (reference would raise IndexError)\n"); return -1; }
  *v_out = r_Ti + a->theta[idx]; *buy_out = buy;                                            /* :270 */
  return 0;
}
/* TradingAgent.placeLimitOrder :309-349 */
static void ta_place_limit(abo_sim *s, int id, int64_t qty, int is_buy, int64_t price) {
  zi_t *a = &s->zi[id];
  int64_t oid = s->next_order_id++;                                                         /* util/order/Order.py:27,35-42 dense global ids */
  if (qty <= 0) return;
  if (a->n_orders == a->cap_orders) { a->cap_orders = a->cap_orders ? a->cap_orders * 2 : 4; a->orders = (open_order_t *)realloc(a->orders, sizeof(open_order_t) * a->cap_orders); }
  open_order_t *oo = &a->orders[a->n_orders++]; oo->order_id = oid; oo->quantity = qty; oo->limit_price = price; oo->is_buy = is_buy; /* :342 */
  if (s->trace & ABO_TRACE_OPS) { while (s->live_qty.n <= oid) { int64_t z = 0; ib_push(&s->live_qty, &z, 1); } s->live_qty.v[oid] = qty; }
  event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_LIMIT_ORDER;
  e.order.agent_id = id; e.order.order_id = oid; e.order.quantity = qty; e.order.limit_price = price; e.order.is_buy = is_buy;
  ta_send(s, id, &e);                                                                       /* :343 */
}
/* ZeroIntelligenceAgent.placeOrder :277-309 */
static void zi_place_order(abo_sim *s, int id) {
  zi_t *a = &s->zi[id]; int64_t v; int buy;
  if (zi_update_estimates(s, id, &v, &buy)) return;
  int64_t R = abo_rng_randint(a->rs, a->R_min, a->R_max + 1);                               /* :284 */
  int64_t p = buy ? v - R : v + R;                                                          /* :287 */
  int64_t ask_vol = a->has_ask ? a->ask_q : 0, bid_vol = a->has_bid ? a->bid_q : 0;         /* getKnownBidAsk :564-574 */
  if (buy && ask_vol > 0) { int64_t R_ask = v - a->ask; if ((double)R_ask >= a->eta * (double)R) p = a->ask; }   /* :291-297 */
  else if (!buy && bid_vol > 0) { int64_t R_bid = a->bid - v; if ((double)R_bid >= a->eta * (double)R) p = a->bid; } /* :298-305 */
  ta_place_limit(s, id, s->order_size, buy, p);                                             /* :308-309 (size 100) */
}
/* HeuristicBeliefLearningAgent.placeOrder (agent/HeuristicBeliefLearningAgent.py:76-174): belief of a successful transaction per limit price from the
 * orders that led up to the last L trades, limit price = argmax of probability x surplus; falls back to the ZI order without enough history. */
static void hbl_place_order(abo_sim *s, int id) {
  zi_t *a = &s->zi[id];
  if (a->n_stream < a->L) { zi_place_order(s, id); return; }                                /* :83-87 len(stream_history) < L */
  int64_t v; int buy;
  if (zi_update_estimates(s, id, &v, &buy)) return;                                         /* :94 */
  int64_t low_p = INT64_MAX, high_p = 0;                                                    /* :98-108 (sys.maxsize, 0) */
  for (int k = 0; k < a->n_stream; k++) { const hbucket_t *h = book_bucket(&s->book, a->stream_serial[k]); if (!h) { fprintf(stderr, "abides_oracle: history bucket gone\n"); continue; }
    for (int i = 0; i < h->n; i++) { int64_t p = h->r[i].limit_price; if (p < low_p) low_p = p; if (p > high_p) high_p = p; } }
  if (low_p > high_p) return;                                                               /* empty history dicts: np.zeros((negative, 8)) raises in the reference */
  int64_t R = high_p - low_p + 1;
  double *nd = (double *)calloc((size_t)R * 8, sizeof(double));                              /* :112 columns sa, sb, ua, ub, num, denom, Pr, Es */
  for (int k = 0; k < a->n_stream; k++) { const hbucket_t *h = book_bucket(&s->book, a->stream_serial[k]); if (!h) continue;
    for (int i = 0; i < h->n; i++) { const hrec_t *o = &h->r[i]; int64_t p = o->limit_price;           /* :115-139: "transactions" non-empty == successful */
      if (o->is_buy) { if (o->ntx) nd[(p - low_p) * 8 + 1] += 1; else nd[(p - low_p) * 8 + 3] += 1; }
      else { if (o->ntx) nd[(p - low_p) * 8 + 0] += 1; else nd[(p - low_p) * 8 + 2] += 1; } } }
  if (buy) {                                                                                /* :143-150 */
    for (int64_t r = 1; r < R; r++) for (int c = 0; c < 3; c++) nd[r * 8 + c] += nd[(r - 1) * 8 + c];
    for (int64_t r = R - 2; r >= 0; r--) nd[r * 8 + 3] += nd[(r + 1) * 8 + 3];
    for (int64_t r = 0; r < R; r++) nd[r * 8 + 4] = (nd[r * 8 + 0] + nd[r * 8 + 1]) + nd[r * 8 + 2];
  } else {
    for (int64_t r = R - 2; r >= 0; r--) { nd[r * 8 + 0] += nd[(r + 1) * 8 + 0]; nd[r * 8 + 1] += nd[(r + 1) * 8 + 1]; nd[r * 8 + 3] += nd[(r + 1) * 8 + 3]; }
    for (int64_t r = 1; r < R; r++) nd[r * 8 + 2] += nd[(r - 1) * 8 + 2];
    for (int64_t r = 0; r < R; r++) nd[r * 8 + 4] = (nd[r * 8 + 0] + nd[r * 8 + 1]) + nd[r * 8 + 3];
  }
  int64_t best = 0; double best_es = 0.0;
  for (int64_t r = 0; r < R; r++) {
    double den = ((nd[r * 8 + 0] + nd[r * 8 + 1]) + nd[r * 8 + 2]) + nd[r * 8 + 3];         /* :152 np.sum(nd[:, 0:4], axis=1) */
    double pr = den == 0.0 ? 0.0 : nd[r * 8 + 4] / den;                                     /* :158-159 nan_to_num(0 / 0) */
    double es = buy ? pr * (double)(v - (low_p + r)) : pr * (double)((low_p + r) - v);      /* :162-165 */
    if (r == 0 || es > best_es) { best = r; best_es = es; }                                 /* :168 np.argmax: first maximum */
  }
  free(nd);
  if (best_es > 0) ta_place_limit(s, id, s->order_size, buy, low_p + best);                 /* :173-185 int(round(best_p)) */
}
/* MarketMakerAgent (agent/market_makers/MarketMakerAgent.py), polling mode (subscribe False): wakeup :66-77, receiveMessage :79-108 */
static int64_t mkm_draw_size(zi_t *a) { double h = (double)abo_rng_randint(a->rs, a->mkm_min, a->mkm_max) / 2; return (int64_t)nearbyint(h); }   /* round(randint(min, max) / 2): half to even */
static void ta_cancel_all(abo_sim *s, int id);
static int ta_wakeup_common(abo_sim *s, int id);
/* TradingAgent.requestDataSubscription :160-172 */
static void ta_request_subscription(abo_sim *s, int id, int levels, double freq) {
  event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_MARKET_DATA_SUBSCRIPTION_REQUEST; e.sub_levels = levels; e.sub_freq = freq; ta_send(s, id, &e);
}
static void mkm_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id];
  int can_trade = ta_wakeup_common(s, id);
  if (a->subscribe) {                                                                        /* :69-72: the subscription is requested at the very first wake-up; nothing else ever happens on a wake-up */
    if (!a->sub_requested) { ta_request_subscription(s, id, (int)s->mkm_levels, s->mkm_sub_freq); a->sub_requested = 1; a->state = ST_AWAITING_MARKET_DATA; }
    return;
  }
  if (!can_trade) return;
  ta_cancel_all(s, id); ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD;                 /* getCurrentSpread(depth=subscribe_num_levels): the agent reads level 0 only */
}
static void ta_place_limit(abo_sim *s, int id, int64_t qty, int is_buy, int64_t price);
static void mkm_receive_tail(abo_sim *s, int id, const event_t *m) {
  zi_t *a = &s->zi[id];
  if (a->subscribe) {                                                                        /* :108-139 subscription mode */
    static const double split[5][5] = { { 1, 0, 0, 0, 0 }, { 0.5, 0.5, 0, 0, 0 }, { 0.34, 0.33, 0.33, 0, 0 }, { 0.25, 0.25, 0.25, 0.25, 0 }, { 0.20, 0.20, 0.20, 0.20, 0.20 } };   /* DEFAULT_LEVELS_QUOTE_DICT :6-12 */
    if (!(a->state == ST_AWAITING_MARKET_DATA && m->kind == ABO_MARKET_DATA)) return;
    ta_cancel_all(s, id);
    int num_levels = (int)abo_rng_randint(a->rs, 1, 5);                                      /* randint(1, len(levels_quote_dict)) -> 1..4 */
    if (a->kb_n && a->ka_n) {                                                                /* placeOrders :120-139 */
      a->size = mkm_draw_size(a);
      int64_t bp[5], bv[5], ap[5], av[5]; int nbq = 0, naq = 0;                              /* buy_quotes / sell_quotes: dicts keyed by price (a repeated key keeps its place, takes the new volume) */
      for (int i = 0; i < num_levels; i++) {
        int64_t vol = py_round(split[num_levels - 1][i] * (double)a->size);
        int64_t pb = i < a->kb_n ? a->kb[2 * i] : a->kb[2 * (a->kb_n - 1)] - 1, pa = i < a->ka_n ? a->ka[2 * i] : a->ka[2 * (a->ka_n - 1)] + 1;
        int k = 0; while (k < nbq && bp[k] != pb) k++; if (k == nbq) nbq++; bp[k] = pb; bv[k] = vol;
        k = 0; while (k < naq && ap[k] != pa) k++; if (k == naq) naq++; ap[k] = pa; av[k] = vol;
      }
      for (int k = 0; k < nbq; k++) ta_place_limit(s, id, bv[k], 1, bp[k]);
      for (int k = 0; k < naq; k++) ta_place_limit(s, id, av[k], 0, ap[k]);
    }
    return;
  }
  if (!(a->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD)) return;
  ta_cancel_all(s, id);
  int64_t mid = a->last_trade, spread;
  if (a->has_bid && a->bid != 0 && a->has_ask && a->ask != 0) { mid = (int64_t)((double)(a->ask + a->bid) / 2); spread = (int64_t)((double)llabs(a->ask - a->bid) / 2); }
  else spread = a->last_spread;
  for (int i = 0; i < 2 * (int)s->mkm_levels; i++) {
    a->size = mkm_draw_size(a);
    ta_place_limit(s, id, a->size, 1, mid - spread - i);
    ta_place_limit(s, id, a->size, 0, mid + spread + i);
  }
  k_set_wakeup(s, id, s->now + s->mkm_wake_ns);
  a->state = ST_AWAITING_WAKEUP;
}
static void orders_remove(zi_t *a, int i) { memmove(a->orders + i, a->orders + i + 1, sizeof(open_order_t) * (a->n_orders - i - 1)); a->n_orders--; }
static void povmm_receive_tail(abo_sim *s, int id, const event_t *m); static void momentum_place_orders(abo_sim *s, int id); static void povexec_receive_tail(abo_sim *s, int id, const event_t *m);
static void noise_place_order(abo_sim *s, int id); static void value_place_order(abo_sim *s, int id);
/* TradingAgent.receiveMessage :181-268 + ZeroIntelligenceAgent.receiveMessage :311-334 */
static void zi_receive(abo_sim *s, int id, const event_t *m) {
  zi_t *a = &s->zi[id]; a->current_time = s->now;
  int had = a->has_open && a->has_close;
  switch (m->kind) {
    case ABO_WHEN_MKT_OPEN: a->mkt_open = m->data; a->has_open = 1; break;
    case ABO_WHEN_MKT_CLOSE: a->mkt_close = m->data; a->has_close = 1; break;
    case ABO_ORDER_EXECUTED: {                                                              /* orderExecuted :422-462 */
      int64_t qty = m->order.is_buy ? m->order.quantity : -m->order.quantity;
      a->shares += qty; a->cash -= qty * m->order.fill_price;
      for (int i = 0; i < a->n_orders; i++) if (a->orders[i].order_id == m->order.order_id) {
        if (m->order.quantity >= a->orders[i].quantity) orders_remove(a, i); else { a->orders[i].quantity -= m->order.quantity; if (m->order.order_id < s->live_qty.n) s->live_qty.v[m->order.order_id] = a->orders[i].quantity; } break; }
      if (a->type == AT_POVEXEC) { a->px_executed += m->order.quantity; a->px_n_executed++; a->px_rem = s->px_quantity - a->px_executed; }   /* handleOrderExecution :103-107 */
      break; }
    case ABO_ORDER_ACCEPTED: break;
    case ABO_ORDER_CANCELLED:                                                               /* orderCancelled :476-489 */
      for (int i = 0; i < a->n_orders; i++) if (a->orders[i].order_id == m->order.order_id) { orders_remove(a, i); break; }
      break;
    case ABO_MKT_CLOSED: a->mkt_closed = 1; break;                                          /* marketClosed :492-499 */
    case ABO_MARKET_DATA:                                                                   /* handleMarketData :539-546 */
      a->kb_n = m->md_nb; a->ka_n = m->md_na; memcpy(a->kb, m->md_bids, sizeof(a->kb)); memcpy(a->ka, m->md_asks, sizeof(a->ka));
      a->has_known = 1; a->has_bid = m->md_nb > 0; a->bid = m->md_bids[0]; a->bid_q = m->md_bids[1]; a->has_ask = m->md_na > 0; a->ask = m->md_asks[0]; a->ask_q = m->md_asks[1];
      a->last_trade = m->data; a->has_last_trade = m->has_data;
      break;
    case ABO_QUERY_TRANSACTED_VOLUME: if (m->mkt_closed) a->mkt_closed = 1; a->transacted_volume = m->tv; break;   /* :248-251,556-558 */
    case ABO_QUERY_ORDER_STREAM: if (m->mkt_closed) a->mkt_closed = 1; a->n_stream = m->n_stream; memcpy(a->stream_serial, m->stream_serial, sizeof(a->stream_serial)); break;   /* :240-246,549-554 */
    case ABO_QUERY_SPREAD:                                                                  /* :232-238, querySpread :514-537, queryLastTrade :502-511 */
      if (m->mkt_closed) a->mkt_closed = 1;
      a->last_trade = m->data; a->has_last_trade = 1;
      if (a->mkt_closed) { a->daily_close = a->last_trade; a->has_daily_close = 1; }
      a->has_known = 1; a->has_bid = m->has_bid; a->bid = m->bid; a->bid_q = m->bid_q; a->has_ask = m->has_ask; a->ask = m->ask; a->ask_q = m->ask_q;
      if (a->type == AT_POVEXEC) for (int k = 0; k < 2; k++) { a->nk[k] = a->nf[k]; memcpy(a->kside[k], a->fside[k], 16 * (size_t)a->nf[k]); }
      break;
    default: break;
  }
  if (a->has_open && a->has_close && !had) {                                                /* :258-268 */
    int64_t off = a->type == AT_MOMENTUM ? s->mom_wake_ns : a->type == AT_POVMM ? s->mm_wake_ns : a->type == AT_MKM ? s->mkm_wake_ns : a->type == AT_POVEXEC ? (s->px_kind ? s->px_start - a->mkt_open : s->px_freq)   /* Passive / AggressiveAgent.getWakeFrequency: timestamp - mkt_open */
                : abo_rng_randint(a->rs, 0, 100);                                           /* ZI/Noise/Value.getWakeFrequency: randint(0, 100) ns */
    k_set_wakeup(s, id, a->mkt_open + off);
  }
  if (a->type == AT_POVMM) { povmm_receive_tail(s, id, m); return; }
  if (a->type == AT_MKM) { mkm_receive_tail(s, id, m); return; }
  if (a->type == AT_POVEXEC) { povexec_receive_tail(s, id, m); return; }
  if (a->type == AT_MOMENTUM) {                                                             /* MomentumAgent.receiveMessage :65-76 */
    if (a->subscribe) { if (a->state == ST_AWAITING_MARKET_DATA && m->kind == ABO_MARKET_DATA && a->kb_n && a->ka_n) momentum_place_orders(s, id); return; }   /* :71-75 placeOrders(bids[0][0], asks[0][0]) */
    if (a->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) { momentum_place_orders(s, id); k_set_wakeup(s, id, s->now + s->mom_wake_ns); a->state = ST_AWAITING_WAKEUP; }
    return;
  }
  if (a->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) {                      /* ZI :319-334, NoiseAgent :123-129, ValueAgent :245-251 */
    if (a->mkt_closed) return;
    if (a->type == AT_NOISE) noise_place_order(s, id); else if (a->type == AT_VALUE) value_place_order(s, id); else if (a->type == AT_HBL) hbl_place_order(s, id); else zi_place_order(s, id);
    a->state = ST_AWAITING_WAKEUP;
  }
  if (a->type == AT_HBL && a->state == ST_AWAITING_STREAM && m->kind == ABO_QUERY_ORDER_STREAM) {   /* HeuristicBeliefLearningAgent.receiveMessage :176-195 */
    if (a->mkt_closed) return;
    ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD;
  }
}


/* ====================================================================================================
 * rmsc03 population: NoiseAgent, ValueAgent, MomentumAgent, POVMarketMakerAgent (config/rmsc03.py)
 * ==================================================================================================== */
static void ta_cancel_all(abo_sim *s, int id) {                       /* cancelOrders / cancelAllOrders: one CANCEL_ORDER per open order */
  zi_t *a = &s->zi[id];
  for (int i = 0; i < a->n_orders; i++) {
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_CANCEL_ORDER;
    e.order.agent_id = id; e.order.order_id = a->orders[i].order_id; e.order.quantity = a->orders[i].quantity; e.order.limit_price = a->orders[i].limit_price; e.order.is_buy = a->orders[i].is_buy;
    ta_send(s, id, &e);
  }
}
/* TradingAgent.wakeup :142-158 -> can_trade */
static int ta_wakeup_common(abo_sim *s, int id) {
  zi_t *a = &s->zi[id]; a->current_time = s->now; a->first_wake = 0;
  if (!a->has_open) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_OPEN; ta_send(s, id, &e); memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_CLOSE; ta_send(s, id, &e); }
  return a->has_open && a->has_close && !a->mkt_closed;
}
/* NoiseAgent.wakeup (agent/NoiseAgent.py:82-112) */
static void noise_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id]; ta_wakeup_common(s, id); a->state = ST_INACTIVE;
  if (!a->has_open || !a->has_close) return;
  a->trading = 1;
  if (a->mkt_closed && a->has_daily_close) return;
  if (a->wakeup_time > s->now) k_set_wakeup(s, id, a->wakeup_time);                       /* :98-99 */
  ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD;                                      /* :101-108 (both branches query the spread) */
}
static void noise_place_order(abo_sim *s, int id) {                                         /* :114-121 */
  zi_t *a = &s->zi[id];
  int buy = (int)abo_rng_randint(s->g, 0, 2);                                               /* np.random.randint(0, 1 + 1): GLOBAL stream */
  if (buy && a->has_ask && a->ask != 0) ta_place_limit(s, id, a->size, 1, a->ask);
  else if (!buy && a->has_bid && a->bid != 0) ta_place_limit(s, id, a->size, 0, a->bid);
}
/* ValueAgent.wakeup (agent/ValueAgent.py:100-138): same shape as the ZI wakeup */
static void value_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id]; ta_wakeup_common(s, id); a->state = ST_INACTIVE;
  if (!a->has_open || !a->has_close) return;
  a->trading = 1;
  if (a->mkt_closed && a->has_daily_close) return;
  double delta_time = rng_exponential(a->rs, 1.0 / s->lambda_a);
  k_set_wakeup(s, id, s->now + py_round(delta_time));
  if (a->mkt_closed && !a->has_daily_close) { ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD; return; }
  ta_cancel_all(s, id);
  ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD;
}
/* ValueAgent.updateEstimates :140-205 + placeOrder :207-243 */
static void value_place_order(abo_sim *s, int id) {
  zi_t *a = &s->zi[id];
  int64_t obs_t = oracle_observe(s, a->current_time, s->sigma_n, a->rs);
  if (!a->has_prev) { a->prev_wake = a->mkt_open; a->has_prev = 1; }
  double kappa = s->agent_kappa, r_bar = s->r_bar, sigma_n = s->sigma_n;
  double delta = (double)(a->current_time - a->prev_wake);
  double pw = pow(1 - kappa, delta);
  double r_tprime = (1 - pw) * r_bar; r_tprime += pw * a->r_t;
  double pw2 = pow(1 - kappa, 2 * delta);
  double sigma_tprime = pw2 * a->sigma_t; sigma_tprime += ((1 - pw2) / (1 - pow(1 - kappa, 2.0))) * s->sigma_s;
  a->r_t = (sigma_n / (sigma_n + sigma_tprime)) * r_tprime; a->r_t += (sigma_tprime / (sigma_n + sigma_tprime)) * (double)obs_t;
  a->sigma_t = (sigma_n * a->sigma_t) / (sigma_n + a->sigma_t);
  double d2 = (double)(a->mkt_close - a->current_time); if (!(d2 > 0)) d2 = 0;
  double pw3 = pow(1 - kappa, d2);
  double r_T = (1 - pw3) * r_bar; r_T += pw3 * a->r_t;
  int64_t r_Ti = py_round(r_T); a->prev_wake = a->current_time;
  int buy; int64_t p;
  if (a->has_bid && a->bid != 0 && a->has_ask && a->ask != 0) {                              /* if bid and ask */
    int64_t mid = (int64_t)((double)(a->ask + a->bid) / 2); int64_t spread = llabs(a->ask - a->bid); int64_t adjust;
    if (abo_rng_double(s->g) < s->value_percent_aggr) adjust = 0;                            /* np.random.rand() < percent_aggr */
    else adjust = abo_rng_randint(s->g, 0, s->value_depth_spread * spread);                                      /* np.random.randint(0, depth_spread * spread) */
    if (r_Ti < mid) { buy = 0; p = a->bid + adjust; } else { buy = 1; p = a->ask - adjust; }
  } else { buy = (int)abo_rng_randint(s->g, 0, 2); p = r_Ti; }
  ta_place_limit(s, id, a->size, buy, p);
}
/* MomentumAgent.wakeup (agent/examples/MomentumAgent.py:53-63) */
static void momentum_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id];
  int can_trade = ta_wakeup_common(s, id);
  if (a->subscribe) { if (!a->sub_requested) { ta_request_subscription(s, id, 1, s->mom_sub_freq); a->sub_requested = 1; a->state = ST_AWAITING_MARKET_DATA; } return; }   /* :56-59 levels=1, freq=10e9 */
  if (can_trade) { ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD; }
}
static double np_round2(double x) { return nearbyint(x * 100.0) / 100.0; }                   /* numpy float64.round(2) */
static void momentum_place_orders(abo_sim *s, int id) {                                     /* placeOrders :78-93, ma :95-99 */
  zi_t *a = &s->zi[id];
  if (!(a->has_bid && a->bid != 0 && a->has_ask && a->ask != 0)) return;
  if (a->n_mids == a->cap_mids) { a->cap_mids = a->cap_mids ? 2 * a->cap_mids : 64; a->mids = (double *)realloc(a->mids, 8 * a->cap_mids); }
  a->mids[a->n_mids++] = (double)(a->bid + a->ask) / 2;
  int L = a->n_mids;
  for (int w = 0; w < 2; w++) { int n = w ? 50 : 20;
    if (L > n) { double c1 = 0, c0 = 0; for (int i = 0; i < L; i++) { c1 += a->mids[i]; if (i == L - 1 - n) c0 = c1; }      /* np.cumsum; ret[n:] - ret[:-n] */
      double v = np_round2((c1 - c0) / n); if (w) { a->avg50 = v; a->has50 = 1; } else { a->avg20 = v; a->has20 = 1; } } }
  if (a->has20 && a->has50) { if (a->avg20 >= a->avg50) ta_place_limit(s, id, a->size, 1, a->ask); else ta_place_limit(s, id, a->size, 0, a->bid); }
}
/* POVMarketMakerAgent.wakeup (agent/market_makers/POVMarketMakerAgent.py:83-100), with the getTransactedVolume alias */
static void povmm_wakeup(abo_sim *s, int id) {
  if (!ta_wakeup_common(s, id)) return;
  ta_get_spread(s, id);                                                                      /* depth = subscribe_num_levels = 1 */
  event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_TRANSACTED_VOLUME; e.lookback = s->mm_wake_ns; ta_send(s, id, &e);
}
static void povmm_receive_tail(abo_sim *s, int id, const event_t *m) {                      /* receiveMessage :102-150 */
  zi_t *a = &s->zi[id];
  if (m->kind == ABO_QUERY_TRANSACTED_VOLUME && a->aw_vol) {                                 /* updateOrderSize :152-156 */
    int64_t qty = py_round(s->mm_pov * (double)a->transacted_volume); a->order_size = qty >= s->mm_min_size ? qty : s->mm_min_size; a->aw_vol = 0; }
  if (m->kind == ABO_QUERY_SPREAD && a->aw_spread) {
    if (a->has_bid && a->bid != 0 && a->has_ask && a->ask != 0) { a->last_mid = (int64_t)((double)(a->ask + a->bid) / 2); a->has_last_mid = 1; a->aw_spread = 0; } }
  if (!a->aw_spread && !a->aw_vol) {
    ta_cancel_all(s, id);
    int64_t mid = a->last_mid, highest_bid = mid - 1, lowest_ask = mid + s->mm_window;      /* computeOrdersToPlace :158-177, anchor bottom */
    int64_t lowest_bid = highest_bid - s->mm_ticks, highest_ask = lowest_ask + s->mm_ticks;
    for (int64_t p = lowest_bid; p <= highest_bid; p++) ta_place_limit(s, id, a->order_size, 1, p);
    for (int64_t p = lowest_ask; p <= highest_ask; p++) ta_place_limit(s, id, a->order_size, 0, p);
    a->aw_spread = a->aw_vol = 1;
    k_set_wakeup(s, id, s->now + s->mm_wake_ns);
  }
}
/* POVExecutionAgent (agent/execution/baselines/pov_agent.py): wakeup :55-64, receiveMessage :69-99, placeMarketOrder TradingAgent.py:351-397 */
enum { ST_AWAITING_TV = 3 };
static void px_market_order(abo_sim *s, int id, int64_t quantity) {                          /* TradingAgent.placeMarketOrder :351-397: one limit order per level of the cached opposite side */
  zi_t *a = &s->zi[id];
  if (quantity <= 0) return;
  int k = s->px_is_buy ? 1 : 0; const int64_t *side = a->kside[k]; int n = a->nk[k];
  int nq = 0; int64_t *qp = (int64_t *)malloc(16 * (size_t)(n + 1));
  for (int i = 0; i < n; i++) { if (quantity <= side[2 * i + 1]) { qp[2 * nq] = side[2 * i]; qp[2 * nq + 1] = quantity; nq++; break; } qp[2 * nq] = side[2 * i]; qp[2 * nq + 1] = side[2 * i + 1]; nq++; quantity -= side[2 * i + 1]; }
  for (int i = 0; i < nq; i++) ta_place_limit(s, id, qp[2 * i + 1], s->px_is_buy, qp[2 * i]);
  free(qp);
}
static void povexec_wakeup(abo_sim *s, int id) {
  zi_t *a = &s->zi[id];
  if (!ta_wakeup_common(s, id)) return;
  if (s->px_kind) {                                                                          /* PassiveAgent.wakeup :40-58 / AggressiveAgent.wakeup :26-32 */
    if (s->now != s->px_start) return;
    if (s->px_kind == 1 && s->px_limit) { ta_place_limit(s, id, s->px_quantity, s->px_is_buy, s->px_limit); return; }
    ta_get_spread(s, id); a->state = ST_AWAITING_SPREAD; return;
  }
  if (a->px_rem > 0 && s->now < s->px_end) {
    k_set_wakeup(s, id, s->now + s->px_freq); ta_cancel_all(s, id); ta_get_spread(s, id);                 /* depth = sys.maxsize */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_TRANSACTED_VOLUME; e.lookback = s->px_lookback; ta_send(s, id, &e);
    a->state = ST_AWAITING_TV;
  }
}
static void povexec_receive_tail(abo_sim *s, int id, const event_t *m) {
  zi_t *a = &s->zi[id];
  if (s->px_kind) {                                                                          /* PassiveAgent.receiveMessage :60-70 / AggressiveAgent.receiveMessage :34-37 (the state is never left) */
    if (!(a->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD)) return;
    if (s->px_kind == 1) { int has = s->px_is_buy ? a->has_bid : a->has_ask; if (!has) { fprintf(stderr, "abides_oracle: PassiveAgent met an empty book side (limit price None)\n"); return; }
      ta_place_limit(s, id, s->px_quantity, s->px_is_buy, s->px_is_buy ? a->bid : a->ask); }
    else px_market_order(s, id, s->px_quantity);
    return;
  }
  if (s->now > s->px_end) return;
  if (a->px_rem > 0 && a->state == ST_AWAITING_TV && m->kind == ABO_QUERY_TRANSACTED_VOLUME && s->now > s->px_start) {
    int64_t quantity = py_round(s->px_pov * (double)a->transacted_volume);
    ta_cancel_all(s, id);
    if (quantity > 0) {                                                                      /* walk the cached opposite side, one limit order per level */
      int k = s->px_is_buy ? 1 : 0; const int64_t *side = a->kside[k]; int n = a->nk[k];
      int nq = 0; int64_t *qp = (int64_t *)malloc(16 * (size_t)(n + 1));
      for (int i = 0; i < n; i++) { if (quantity <= side[2 * i + 1]) { qp[2 * nq] = side[2 * i]; qp[2 * nq + 1] = quantity; nq++; break; } qp[2 * nq] = side[2 * i]; qp[2 * nq + 1] = side[2 * i + 1]; nq++; quantity -= side[2 * i + 1]; }
      for (int i = 0; i < nq; i++) ta_place_limit(s, id, qp[2 * i + 1], s->px_is_buy, qp[2 * i]);
      free(qp);
    }
  }
}
static void agent_wakeup(abo_sim *s, int id) {
  switch (s->zi[id].type) { case AT_NOISE: noise_wakeup(s, id); break; case AT_VALUE: value_wakeup(s, id); break; case AT_MOMENTUM: momentum_wakeup(s, id); break;
    case AT_POVMM: povmm_wakeup(s, id); break; case AT_POVEXEC: povexec_wakeup(s, id); break; case AT_MKM: mkm_wakeup(s, id); break; default: zi_wakeup(s, id); }
}

/* ---------------- config: config/sparse_zi_100.py / config/sparse_zi_1000.py ---------------- */
static abo_rng *new_stream(abo_sim *s) { /* RandomState(seed=np.random.randint(0, 2**32, dtype=uint64)) */
  abo_rng *r = abo_rng_new(rng_u32_raw(s->g)); r->record = (s->trace & ABO_TRACE_TAPES) != 0; return r;
}
#define NS_PER_S 1000000000LL

/* The numbers of the reference's config scripts as an abx_sim_config (the parameter surface the CUDA product takes as well: the same struct
 * drives both, so a parity test can change any field on both sides at once).  variant 100 / 1000: config/sparse_zi_100.py / sparse_zi_1000.py;
 * 3: config/rmsc03.py; 4: rmsc03 + one POVExecutionAgent in the style of config/execution_iabs_plots.py:200-226.  Capacities / rng_mode are the
 * product's business and stay 0 here. */
int abo_default_config(int variant, abx_sim_config *c) {
  if (!c) return -1;
  memset(c, 0, sizeof(*c)); c->version = ABX_VERSION;
  if (variant == 100 || variant == 1000) {
    static const int n1000[7] = { 143, 143, 143, 143, 143, 143, 142 }, n100[7] = { 15, 15, 14, 14, 14, 14, 14 };
    static const int rmin[7] = { 0, 0, 0, 0, 0, 250, 250 }, rmax[7] = { 250, 500, 1000, 1000, 2000, 500, 500 };     /* sparse_zi_1000.py:196-204 */
    static const double eta[7] = { 1, 1, 0.8, 1, 0.8, 0.8, 1 };
    c->n_groups = 7; c->q_max = 10; c->n_agents = 1;
    for (int g = 0; g < 7; g++) { c->groups[g].count = variant == 1000 ? n1000[g] : n100[g]; c->groups[g].r_min = rmin[g]; c->groups[g].r_max = rmax[g]; c->groups[g].eta = eta[g]; c->n_agents += c->groups[g].count; }
    c->start_ns = 0; c->stop_ns = 17 * 3600 * NS_PER_S;                               /* :86-88 midnight .. 17:00 */
    c->mkt_open_ns = (9 * 3600 + 30 * 60) * NS_PER_S; c->mkt_close_ns = 16 * 3600 * NS_PER_S;
    c->default_computation_delay_ns = NS_PER_S; c->starting_cash = 10000000; c->order_size = 100; c->stream_history = 10;
    c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4; c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;   /* :130-142 */
    c->sigma_n = 1000000.0; c->agent_kappa = 1.67e-15; c->sigma_s = 1e-4; c->sigma_pv = 5e6; c->lambda_a = 1e-12;                                  /* :232-250 */
    if (variant == 1000) { c->latency_model = ABX_LAT_MATRIX_NOISE; c->n_noise = 6; c->latency_mirrored = 1; c->latency_lo = 21000; c->latency_hi = 13000000; }   /* :264-286 */
    else { c->latency_model = ABX_LAT_CUBIC; c->n_noise = 1; c->latency_lo = 21000; c->latency_hi = 100000; c->jitter = 0.3; c->jitter_clip = 0.05; c->jitter_unit = 5.0; }   /* sparse_zi_100.py:305-318 */
    return 0;
  }
  if (variant == 3 || variant == 4) {
    c->population = 1; c->n_noise_agents = 50; c->n_value_agents = 10; c->n_mm_agents = 1; c->n_momentum_agents = 2; c->n_agents = 64; c->q_max = 10;
    c->mkt_open_ns = (9 * 3600 + 30 * 60) * NS_PER_S; c->mkt_close_ns = (9 * 3600 + 45 * 60) * NS_PER_S;   /* rmsc03.py:69-70 */
    c->start_ns = c->mkt_open_ns; c->stop_ns = c->mkt_close_ns + 60 * NS_PER_S;                              /* :205-207 */
    c->starting_cash = 10000000; c->stream_history = 10;
    c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4; c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;
    c->sigma_n = 1e5 / 10; c->agent_kappa = 1.67e-15; c->sigma_s = 100000; c->lambda_a = 7e-11;            /* :77-80; sigma_s is ValueAgent's default */
    c->latency_model = ABX_LAT_ZERO; c->n_noise = 1;
    c->size_lo = 20; c->size_hi = 50; c->value_depth_spread = 2; c->value_percent_aggr = 0.1;              /* NoiseAgent.py:34, ValueAgent.py:53-56 */
    c->noise_wake_lo_ns = 9 * 3600 * NS_PER_S; c->noise_wake_hi_ns = 16 * 3600 * NS_PER_S;                  /* :115-116 */
    c->mom_min_size = 1; c->mom_max_size = 10; c->mom_wake_ns = 20 * NS_PER_S;                              /* :190-192 */
    c->mm_pov = 0.05; c->mm_min_order_size = 20; c->mm_window_size = 5; c->mm_num_ticks = 20; c->mm_wake_ns = NS_PER_S;   /* :41-45 */
    if (variant == 4) { c->n_pov_exec = 1; c->n_agents += 1; c->pov_exec_is_buy = 1; c->pov_exec_pov = 0.5; c->pov_exec_quantity = 120000;
      c->pov_exec_start_ns = (9 * 3600 + 32 * 60) * NS_PER_S; c->pov_exec_end_ns = (9 * 3600 + 43 * 60) * NS_PER_S; c->pov_exec_freq_ns = 30 * NS_PER_S; c->pov_exec_lookback_ns = 30 * NS_PER_S; }
    return 0;
  }
  if (variant == 2) {                                                                       /* config/rmsc02.py: rmsc01 + subscription mode + the sparse_zi_1000 latency, midnight .. 17:00 */
    if (abo_default_config(1, c)) return -1;
    c->mkm_subscribe = 1; c->mom_subscribe = 1; c->mkm_sub_freq_ns = 10 * NS_PER_S; c->mom_sub_freq_ns = 10 * NS_PER_S;   /* :108,205; MarketMakerAgent subscribe_freq=10e9, MomentumAgent.py:58 */
    c->start_ns = 0; c->stop_ns = 17 * 3600 * NS_PER_S;                                      /* :264-265 */
    c->latency_model = ABX_LAT_MATRIX_NOISE; c->n_noise = 6; c->latency_mirrored = 0; c->latency_lo = 21000; c->latency_hi = 13000000;   /* :268-269 */
    return 0;
  }
  if (variant == 1) {                                                                       /* config/rmsc01.py */
    c->population = 3; c->n_mm_agents = 1; c->n_groups = 2; c->q_max = 10; c->n_momentum_agents = 24;
    c->groups[0].count = 50; c->groups[0].r_min = 0; c->groups[0].r_max = 100; c->groups[0].eta = 1;          /* ZI :134-157 */
    c->groups[1].count = 25; c->groups[1].r_min = 0; c->groups[1].r_max = 100; c->groups[1].eta = 1; c->hbl_L = 2;   /* HBL :163-187 */
    c->n_agents = 1 + 1 + 50 + 25 + 24;
    c->mkt_open_ns = (9 * 3600 + 30 * 60) * NS_PER_S; c->mkt_close_ns = 16 * 3600 * NS_PER_S; c->start_ns = c->mkt_open_ns; c->stop_ns = (16 * 3600 + 60) * NS_PER_S;   /* :72-73,246-247 */
    c->starting_cash = 10000000; c->order_size = 100; c->stream_history = 10;
    c->r_bar = 1e5; c->kappa = 1.67e-12; c->fund_vol = 1e-4; c->megashock_lambda_a = 2.77778e-13; c->megashock_mean = 1e3; c->megashock_var = 5e4;
    c->sigma_n = 10000; c->agent_kappa = 1.67e-15; c->sigma_s = 1e-4; c->sigma_pv = 5e4; c->lambda_a = 1e-12;        /* :141-153 */
    c->latency_model = ABX_LAT_ZERO; c->n_noise = 1;
    c->mom_min_size = 1; c->mom_max_size = 10; c->mom_wake_ns = 60 * NS_PER_S;                                     /* MomentumAgent default wake_up_freq "60s" */
    c->mkm_min_size = 500; c->mkm_max_size = 1000; c->mkm_num_levels = 5; c->mkm_wake_ns = NS_PER_S;             /* MarketMakerAgent :100-110 + defaults */
    return 0;
  }
  return -1;
}

/* config/sparse_zi_100.py / config/sparse_zi_1000.py (population 0) and config/rmsc03.py (population 1) with the parameters of `c`: the same
 * np.random.seed(seed) cascade in the scripts' source order, whatever the agent counts are. */
static abo_sim *new_population0(const abx_sim_config *c, uint32_t seed, int trace) {
  abo_sim *s = (abo_sim *)calloc(1, sizeof(abo_sim));
  int cubic = c->latency_model == ABX_LAT_CUBIC;
  s->variant = cubic ? 100 : 1000; s->seed = seed; s->trace = trace;
  s->pop_hash = s->note_hash = s->snap_hash = FNV_OFF;
  int n = c->n_agents; s->n_agents = n;
  s->g = abo_rng_new(seed);                                              /* np.random.seed(seed)  sparse_zi_1000.py:72 */
  s->start_time = c->start_ns; s->stop_time = c->stop_ns;                /* :86-88 */
  /* symbols["JPM"] :130-142 */
  s->r_bar = c->r_bar; s->kappa = c->kappa; s->agent_kappa = c->agent_kappa; s->sigma_s = c->sigma_s; s->fund_vol = c->fund_vol;
  s->megashock_lambda = c->megashock_lambda_a; s->megashock_mean = c->megashock_mean; s->megashock_var = c->megashock_var;
  s->order_size = c->order_size; s->starting_cash = c->starting_cash;
  s->sym_rs = new_stream(s);                                             /* :140 */
  s->kernel_rs = new_stream(s);                                          /* :146-148 */
  if (cubic) s->lat_rs = new_stream(s);                                  /* sparse_zi_100.py:152 */
  s->mkt_open = c->mkt_open_ns; s->mkt_close = c->mkt_close_ns;
  /* SparseMeanRevertingOracle.__init__ :36-79 */
  s->or_t = s->mkt_open; s->or_v = (int64_t)s->r_bar;
  oracle_new_megashock(s, s->mkt_open);
  /* ExchangeAgent :174-194 */
  s->exch_rs = new_stream(s); s->pipeline_delay = c->exchange_pipeline_delay_ns; s->exch_comp_delay = c->exchange_computation_delay_ns;
  book_init(&s->book, c->stream_history, s, exch_book_send);
  /* ZI agents :211-251 */
  s->sigma_n = c->sigma_n; s->lambda_a = c->lambda_a;
  s->zi = (zi_t *)calloc(n, sizeof(zi_t));
  int id = 1;
  for (int g = 0; g < c->n_groups; g++) for (int k = 0; k < c->groups[g].count; k++, id++) {
    zi_t *a = &s->zi[id]; a->rs = new_stream(s); a->group = g; a->R_min = c->groups[g].r_min; a->R_max = c->groups[g].r_max; a->eta = c->groups[g].eta;
    a->starting_cash = a->cash = c->starting_cash; a->first_wake = 1; a->state = ST_AWAITING_WAKEUP; a->r_t = s->r_bar; a->sigma_t = 0; a->q_max = c->q_max;
    /* theta: sorted(np.round(normal(0, sqrt(sigma_pv), size=2*q_max)), reverse=True) -> int   ZeroIntelligenceAgent.py:65-70 */
    double th[64]; int m = 2 * a->q_max;
    for (int i = 0; i < m; i++) th[i] = nearbyint(rng_normal(a->rs, 0.0, sqrt(c->sigma_pv)));
    for (int i = 1; i < m; i++) { double x = th[i]; int j = i - 1; while (j >= 0 && th[j] < x) { th[j + 1] = th[j]; j--; } th[j + 1] = x; }
    for (int i = 0; i < m; i++) a->theta[i] = (int32_t)th[i];
  }
  /* latency :264-286 / sparse_zi_100.py:305-318 : N x N uniform draws from the global stream, row-major */
  s->latency = (double *)malloc(sizeof(double) * (size_t)n * n);
  for (size_t i = 0; i < (size_t)n * n; i++) s->latency[i] = c->latency_lo + (c->latency_hi - c->latency_lo) * rng_double_raw(s->g);
  if (!cubic) {
    /* mirror the upper triangle, diagonal 20000; the ZI<->ZI 24h override never fires */
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { if (i > j) s->latency[(size_t)i * n + j] = s->latency[(size_t)j * n + i]; else if (i == j) s->latency[(size_t)i * n + j] = 20000; }
    s->n_noise = c->n_noise; s->use_latency_model = 0;
  } else { s->use_latency_model = 1; s->jitter = c->jitter; s->jitter_clip = c->jitter_clip; s->jitter_unit = c->jitter_unit; }
  /* Kernel.runner :97,105 */
  s->agent_time = (int64_t *)calloc(n, sizeof(int64_t)); s->comp_delay = (int64_t *)malloc(sizeof(int64_t) * n);
  for (int i = 0; i < n; i++) { s->agent_time[i] = s->start_time; s->comp_delay[i] = c->default_computation_delay_ns; }      /* defaultComputationDelay */
  return s;
}
abo_sim *abo_sim_new_sparse_zi(int variant, uint32_t seed, int trace) {
  abx_sim_config c; if ((variant != 100 && variant != 1000) || abo_default_config(variant, &c)) return NULL;
  return new_population0(&c, seed, trace);
}

/* config/rmsc03.py:49-232 (ticker/date only name things): 1 exchange, 50 noise, 10 value, 1 POV market maker, 2 momentum */
static double u_quadratic_inverse_cdf(double y) {                      /* util/util.py:35-58, a = 0, b = 1 */
  double alpha = 12.0 / pow(1.0 - 0.0, 3.0), beta = (1.0 + 0.0) / 2;
  double n = (3 / alpha) * y - pow(beta - 0.0, 3.0);
  double c = n < 0 ? -pow(-n, 1.0 / 3.0) : pow(n, 1.0 / 3.0);
  return c + beta;
}
/* n_pov_exec: one POVExecutionAgent is appended to the population with a RandomState outside the config's seed cascade (it never draws) */
static abo_sim *new_population1(const abx_sim_config *c, uint32_t seed, int trace) {
  abo_sim *s = (abo_sim *)calloc(1, sizeof(abo_sim));
  s->variant = 3; s->seed = seed; s->trace = trace; s->pop_hash = s->note_hash = s->snap_hash = FNV_OFF;
  int n_noise = c->n_noise_agents, n_value = c->n_value_agents, n_mm = c->n_mm_agents, n_mom = c->n_momentum_agents;
  int n0 = 1 + n_noise + n_value + n_mm + n_mom, n = n0 + (c->n_pov_exec ? 1 : 0); s->n_agents = n;
  if (c->n_pov_exec) { s->px_kind = c->exec_kind; s->px_limit = c->exec_limit_price; s->px_id = n0; s->px_pov = c->pov_exec_pov; s->px_quantity = c->pov_exec_quantity; s->px_is_buy = c->pov_exec_is_buy; s->px_start = c->pov_exec_start_ns; s->px_end = c->pov_exec_end_ns;
    s->px_freq = c->pov_exec_freq_ns; s->px_lookback = c->pov_exec_lookback_ns; }
  s->g = abo_rng_new(seed);                                              /* np.random.seed(seed) :58 */
  s->mkt_open = c->mkt_open_ns; s->mkt_close = c->mkt_close_ns;          /* :69-70 */
  s->start_time = c->start_ns; s->stop_time = c->stop_ns;                /* :205-207 */
  s->r_bar = c->r_bar; s->kappa = c->kappa; s->fund_vol = c->fund_vol; s->megashock_lambda = c->megashock_lambda_a; s->megashock_mean = c->megashock_mean; s->megashock_var = c->megashock_var;
  s->sigma_n = c->sigma_n; s->agent_kappa = c->agent_kappa; s->sigma_s = c->sigma_s; s->lambda_a = c->lambda_a;       /* :77-80 */
  s->mm_pov = c->mm_pov; s->mm_min_size = c->mm_min_order_size; s->mm_window = c->mm_window_size; s->mm_ticks = c->mm_num_ticks; s->mm_wake_ns = c->mm_wake_ns; s->mom_wake_ns = c->mom_wake_ns;
  s->value_percent_aggr = c->value_percent_aggr; s->value_depth_spread = c->value_depth_spread; s->starting_cash = c->starting_cash;
  s->sym_rs = new_stream(s);                                             /* :88 */
  s->or_t = s->mkt_open; s->or_v = (int64_t)s->r_bar; oracle_new_megashock(s, s->mkt_open);          /* :91 oracle __init__ */
  s->exch_rs = new_stream(s); s->pipeline_delay = c->exchange_pipeline_delay_ns; s->exch_comp_delay = c->exchange_computation_delay_ns;   /* :108 */
  book_init(&s->book, c->stream_history, s, exch_book_send);
  s->zi = (zi_t *)calloc(n, sizeof(zi_t));
  int64_t n_open = c->noise_wake_lo_ns, n_close = c->noise_wake_hi_ns;                                /* noise_mkt_open/close :115-116 */
  for (int id = 1; id < n; id++) {
    zi_t *a = &s->zi[id]; a->starting_cash = a->cash = c->starting_cash; a->first_wake = 1; a->state = ST_AWAITING_WAKEUP; a->q_max = c->q_max;
    if (id <= n_noise) {                                                 /* NoiseAgent :117-131: wakeup_time (global rand), seed, then __init__ size (global randint) */
      a->type = AT_NOISE; double mult = u_quadratic_inverse_cdf(rng_double_raw(s->g));
      /* float * Timedelta truncates to the Timedelta's unit: microseconds for string-parsed times under pandas >= 3 (the
         recording environment); the reference's pinned pandas 0.24 would truncate to nanoseconds (differs by < 1 us) */
      a->wakeup_time = n_open + (int64_t)(mult * (double)((n_close - n_open) / 1000)) * 1000;
      a->rs = new_stream(s); a->size = abo_rng_randint(s->g, c->size_lo, c->size_hi);
    } else if (id <= n_noise + n_value) { a->type = AT_VALUE; a->rs = new_stream(s); a->size = abo_rng_randint(s->g, c->size_lo, c->size_hi); a->r_t = s->r_bar; a->sigma_t = 0; }   /* :137-155 */
    else if (id <= n_noise + n_value + n_mm) { a->type = AT_POVMM; a->rs = new_stream(s); a->order_size = s->mm_min_size; a->aw_spread = a->aw_vol = 1; }                  /* :160-178 */
    else if (c->n_pov_exec && id == n0) { a->type = AT_POVEXEC; a->rs = abo_rng_new(0); a->px_rem = c->pov_exec_quantity; }
    else { a->type = AT_MOMENTUM; a->rs = new_stream(s); a->size = abo_rng_randint(a->rs, c->mom_min_size, c->mom_max_size); }                                            /* :183-200, MomentumAgent.py:42 */
  }
  s->kernel_rs = new_stream(s);                                          /* :201-204 */
  s->latency = (double *)calloc((size_t)n * n, sizeof(double)); s->n_noise = 1; s->use_latency_model = 0;   /* np.zeros, noise [0.0] :209-210 */
  s->agent_time = (int64_t *)calloc(n, sizeof(int64_t)); s->comp_delay = (int64_t *)calloc(n, sizeof(int64_t));
  for (int i = 0; i < n; i++) { s->agent_time[i] = s->start_time; s->comp_delay[i] = c->default_computation_delay_ns; }           /* defaultComputationDelay 0 :208 */
  return s;
}
abo_sim *abo_sim_new_rmsc03_pov(uint32_t seed, int trace, double pov, int64_t quantity, int is_buy, int64_t start_ns, int64_t end_ns, int64_t freq_ns, int64_t lookback_ns) {
  abx_sim_config c; abo_default_config(pov > 0 ? 4 : 3, &c);
  if (pov > 0) { c.pov_exec_pov = pov; c.pov_exec_quantity = quantity; c.pov_exec_is_buy = is_buy; c.pov_exec_start_ns = start_ns; c.pov_exec_end_ns = end_ns; c.pov_exec_freq_ns = freq_ns; c.pov_exec_lookback_ns = lookback_ns; }
  return new_population1(&c, seed, trace);
}
abo_sim *abo_sim_new_rmsc03(uint32_t seed, int trace) { return abo_sim_new_rmsc03_pov(seed, trace, 0.0, 0, 1, 0, 0, 0, 0); }
/* config/rmsc01.py (population 3): ExchangeAgent, MarketMakerAgent(s), ZeroIntelligenceAgents (groups[0]), HeuristicBeliefLearningAgents (groups[1]),
 * MomentumAgents -- seeds in the script's source order: exchange :89, market makers :109, symbol :129 (+ oracle __init__), ZI :154, HBL :184, momentum :205, kernel :243 */
static abo_sim *new_population3(const abx_sim_config *c, uint32_t seed, int trace) {
  abo_sim *s = (abo_sim *)calloc(1, sizeof(abo_sim));
  s->variant = 1; s->seed = seed; s->trace = trace; s->pop_hash = s->note_hash = s->snap_hash = FNV_OFF;
  int n_mm = c->n_mm_agents, n_zi = c->groups[0].count, n_hbl = c->groups[1].count, n_mom = c->n_momentum_agents, n = 1 + n_mm + n_zi + n_hbl + n_mom; s->n_agents = n;
  s->g = abo_rng_new(seed);
  s->mkt_open = c->mkt_open_ns; s->mkt_close = c->mkt_close_ns; s->start_time = c->start_ns; s->stop_time = c->stop_ns;
  s->r_bar = c->r_bar; s->kappa = c->kappa; s->fund_vol = c->fund_vol; s->megashock_lambda = c->megashock_lambda_a; s->megashock_mean = c->megashock_mean; s->megashock_var = c->megashock_var;
  s->sigma_n = c->sigma_n; s->agent_kappa = c->agent_kappa; s->sigma_s = c->sigma_s; s->lambda_a = c->lambda_a; s->order_size = c->order_size; s->starting_cash = c->starting_cash;
  s->mom_wake_ns = c->mom_wake_ns; s->mkm_levels = c->mkm_num_levels; s->mkm_wake_ns = c->mkm_wake_ns;
  s->exch_rs = new_stream(s); s->pipeline_delay = c->exchange_pipeline_delay_ns; s->exch_comp_delay = c->exchange_computation_delay_ns;
  book_init(&s->book, c->stream_history, s, exch_book_send); s->book.keep_dropped = 1;
  s->zi = (zi_t *)calloc(n, sizeof(zi_t));
  for (int id = 1; id <= n_mm; id++) { zi_t *a = &s->zi[id]; a->type = AT_MKM; a->rs = new_stream(s); a->mkm_min = c->mkm_min_size; a->mkm_max = c->mkm_max_size; a->last_spread = 10; a->size = mkm_draw_size(a); }   /* __init__ :55 */
  s->sym_rs = new_stream(s);
  s->or_t = s->mkt_open; s->or_v = (int64_t)s->r_bar; oracle_new_megashock(s, s->mkt_open);
  for (int id = 1 + n_mm; id < 1 + n_mm + n_zi + n_hbl; id++) {
    int g = id < 1 + n_mm + n_zi ? 0 : 1; zi_t *a = &s->zi[id]; a->type = g ? AT_HBL : AT_ZI; a->L = c->hbl_L; a->rs = new_stream(s); a->group = g;
    a->R_min = c->groups[g].r_min; a->R_max = c->groups[g].r_max; a->eta = c->groups[g].eta; a->r_t = s->r_bar; a->sigma_t = 0; a->q_max = c->q_max;
    double th[64]; int m = 2 * a->q_max;
    for (int i = 0; i < m; i++) th[i] = nearbyint(rng_normal(a->rs, 0.0, sqrt(c->sigma_pv)));
    for (int i = 1; i < m; i++) { double x = th[i]; int j = i - 1; while (j >= 0 && th[j] < x) { th[j + 1] = th[j]; j--; } th[j + 1] = x; }
    for (int i = 0; i < m; i++) a->theta[i] = (int32_t)th[i];
  }
  for (int id = 1 + n_mm + n_zi + n_hbl; id < n; id++) { zi_t *a = &s->zi[id]; a->type = AT_MOMENTUM; a->rs = new_stream(s); a->size = abo_rng_randint(a->rs, c->mom_min_size, c->mom_max_size); }
  for (int id = 1; id < n; id++) { zi_t *a = &s->zi[id]; a->starting_cash = a->cash = c->starting_cash; a->first_wake = 1; a->state = ST_AWAITING_WAKEUP; if (!a->q_max) a->q_max = c->q_max; }
  s->kernel_rs = new_stream(s);
  s->latency = (double *)calloc((size_t)n * n, sizeof(double)); s->n_noise = 1; s->use_latency_model = 0;
  if (c->latency_model == ABX_LAT_MATRIX_NOISE) {                                      /* config/rmsc02.py:268-269: np.random.uniform(low, high, size=(n, n)) row major, noise list of n_noise entries */
    for (size_t k = 0; k < (size_t)n * n; k++) s->latency[k] = rng_uniform(s->g, c->latency_lo, c->latency_hi);
    s->n_noise = c->n_noise;
  }
  for (int id = 1; id < n; id++) { zi_t *a = &s->zi[id]; a->subscribe = a->type == AT_MKM ? c->mkm_subscribe : a->type == AT_MOMENTUM ? c->mom_subscribe : 0; }
  s->mkm_sub_freq = (double)c->mkm_sub_freq_ns; s->mom_sub_freq = (double)c->mom_sub_freq_ns;
  s->agent_time = (int64_t *)calloc(n, sizeof(int64_t)); s->comp_delay = (int64_t *)calloc(n, sizeof(int64_t));
  for (int i = 0; i < n; i++) { s->agent_time[i] = s->start_time; s->comp_delay[i] = c->default_computation_delay_ns; }
  return s;
}
abo_sim *abo_sim_new_config(const abx_sim_config *c, uint32_t seed, int trace) {
  if (!c || c->n_agents < 2 || c->q_max < 1 || c->q_max > 32) return NULL;
  if (c->population == 0) { int n = 1; for (int g = 0; g < c->n_groups; g++) n += c->groups[g].count; if (n != c->n_agents) return NULL; return new_population0(c, seed, trace); }
  if (c->population == 3) { if (c->n_groups != 2 || 1 + c->n_mm_agents + c->groups[0].count + c->groups[1].count + c->n_momentum_agents != c->n_agents || c->hbl_L < 1 || c->hbl_L > 16) return NULL; return new_population3(c, seed, trace); }
  if (c->population == 1) { if (1 + c->n_noise_agents + c->n_value_agents + c->n_mm_agents + c->n_momentum_agents + (c->n_pov_exec ? 1 : 0) != c->n_agents || c->n_mm_agents > 1) return NULL; return new_population1(c, seed, trace); }
  return NULL;
}

/* External tapes: re-run a simulation whose random variates were drawn elsewhere (the CUDA simulator under its Philox streams).  Streams in the
 * product's order (include/abides_b200.h, abx_sim_reset_tape): 0 symbol, 1 kernel, 2 latency model, 3 global, 3 + a agent a; off has n_agents + 4
 * entries.  The start-of-run state those generators produced comes with them: theta [n_agents][20] (population 0), latency[a][0] / latency[0][a],
 * Noise/Value/Momentum sizes and NoiseAgent wake-up times (population 1).  Call before abo_sim_start / run; the arrays must outlive the run. */
int abo_sim_set_external(abo_sim *s, const uint64_t *bits, const uint8_t *kinds, const int64_t *off, const int32_t *theta, const double *lat_to,
                         const double *lat_from, const int32_t *sizes, const int64_t *wakes) {
  if (!s || s->started || !bits || !kinds || !off) return -1;
  int n = s->n_agents;
  for (int st = 0; st < n + 3; st++) {
    abo_rng *r = st == 0 ? s->sym_rs : st == 1 ? s->kernel_rs : st == 2 ? s->lat_rs : st == 3 ? s->g : s->zi[st - 3].rs;
    if (!r) continue;
    if (st == 1 && !s->use_latency_model && s->n_noise == 1) continue;   /* noise [x]: randint(0, 1) is 0 whatever the stream (the product draws nothing) */
    r->replay = 1; r->rk = kinds + off[st]; r->rv = bits + off[st]; r->rn = off[st + 1] - off[st]; r->rpos = 0; r->rerr = 0; r->has_gauss = 0;
  }
  for (int a = 1; a < n; a++) {
    zi_t *z = &s->zi[a];
    if (theta && (z->type == AT_ZI || z->type == AT_HBL)) for (int i = 0; i < 2 * z->q_max && i < 20; i++) z->theta[i] = theta[(size_t)a * 20 + i];
    if (lat_to && lat_from) { s->latency[(size_t)a * n] = lat_to[a]; s->latency[a] = lat_from[a]; }
    if (sizes && (z->type == AT_NOISE || z->type == AT_VALUE || z->type == AT_MOMENTUM)) z->size = sizes[a];
    if (wakes && z->type == AT_NOISE) z->wakeup_time = wakes[a];
  }
  s->or_t = s->mkt_open; s->or_v = (int64_t)s->r_bar; s->n_gexp = 0;
  oracle_new_megashock(s, s->mkt_open);                                  /* the oracle's __init__ draw, now from the supplied streams */
  return 0;
}
int abo_sim_rng_error(abo_sim *s) {
  int e = 0; abo_rng *rs[5] = { s->sym_rs, s->kernel_rs, s->lat_rs, s->g, s->exch_rs };
  for (int i = 0; i < 5; i++) if (rs[i]) e |= rs[i]->rerr;
  for (int a = 1; a < s->n_agents; a++) if (s->zi[a].rs) e |= s->zi[a].rs->rerr;
  return e;
}
/* draws left unread on the external tapes (0 when the run consumed exactly what the other side drew) */
int64_t abo_sim_external_unread(abo_sim *s) {
  int64_t u = 0; abo_rng *rs[4] = { s->sym_rs, s->kernel_rs, s->lat_rs, s->g };
  for (int i = 0; i < 4; i++) if (rs[i] && rs[i]->replay) u += rs[i]->rn - rs[i]->rpos;
  for (int a = 1; a < s->n_agents; a++) if (s->zi[a].rs && s->zi[a].rs->replay) u += s->zi[a].rs->rn - s->zi[a].rs->rpos;
  return u;
}
/* POVExecutionAgent: rem_quantity, executed orders, open orders */
void abo_sim_pov_exec(abo_sim *s, int64_t *out3) { out3[0] = out3[1] = out3[2] = 0; if (s->px_id) { zi_t *a = &s->zi[s->px_id]; out3[0] = a->px_rem; out3[1] = a->px_n_executed; out3[2] = a->n_orders; } }
void abo_sim_free(abo_sim *s) {
  if (!s) return;
  for (int i = 1; i < s->n_agents; i++) { abo_rng_free(s->zi[i].rs); free(s->zi[i].orders); free(s->zi[i].mids); for (int k = 0; k < 2; k++) { free(s->zi[i].kside[k]); free(s->zi[i].fside[k]); } }
  free(s->zi); abo_rng_free(s->g); abo_rng_free(s->kernel_rs); abo_rng_free(s->lat_rs); abo_rng_free(s->sym_rs); abo_rng_free(s->exch_rs);
  book_destroy(&s->book); free(s->latency); free(s->agent_time); free(s->comp_delay); free(s->q.e); free(s->gexp); free(s->live_qty.v);
  free(s->pops.v); free(s->ops.v); free(s->notes.v); free(s->snaps.v); free(s->ckpt); free(s);
}

/* Kernel.runner :154-175 */
void abo_sim_start(abo_sim *s) {
  if (s->started) return; s->started = 1;
  s->g->record = (s->trace & ABO_TRACE_TAPES) != 0;                      /* runtime draws on the GLOBAL np.random stream */
  s->book.last_trade = (int64_t)s->r_bar; s->book.has_last_trade = 1;   /* ExchangeAgent.kernelInitializing :91-102, getDailyOpenPrice */
  s->now = s->start_time;
  for (int i = 0; i < s->n_agents; i++) k_set_wakeup(s, i, s->start_time);   /* Agent.kernelStarting :78 */
}
static void ckpt_push(abo_sim *s) { if (s->n_ckpt == s->cap_ckpt) { s->cap_ckpt = s->cap_ckpt ? s->cap_ckpt * 2 : 256; s->ckpt = (uint64_t *)realloc(s->ckpt, 8 * s->cap_ckpt); } s->ckpt[s->n_ckpt++] = s->pop_hash; }

/* Kernel.runner hot loop :190-292 */
int64_t abo_sim_run_until(abo_sim *s, int64_t until, int *done) {
  int64_t n0 = s->ttl;
  abo_sim_start(s);
  while (!s->loop_done) {
    if (s->q.n == 0 || !(s->now <= s->stop_time)) { s->loop_done = 1; break; }              /* :190 condition tested BEFORE the pop */
    if (s->q.e[0].t > until) break;
    event_t ev; heap_pop(&s->q, &ev); s->now = ev.t;                                         /* :192 */
    s->ttl++;                                                                                /* :211 */
    s->pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s->pop_hash, ev.t), ev.recipient), ev.type), ev.uniq);
    if (s->ttl % 1000 == 0) ckpt_push(s);
    if (s->trace & ABO_TRACE_POPS) { int64_t row[5] = { ev.t, ev.recipient, ev.type, ev.uniq, ev.kind }; ib_push(&s->pops, row, 5); }
    s->addl_delay = 0;                                                                       /* :214 */
    int a = ev.recipient;
    if (s->agent_time[a] > s->now) { ev.t = s->agent_time[a]; k_put(s, &ev); continue; }     /* :224-230 / :258-264 requeue, same uniq */
    s->agent_time[a] = s->now;                                                               /* :234 / :268 */
    if (ev.type == ABO_T_WAKEUP) { if (a != 0) agent_wakeup(s, a); /* exchange: Agent.wakeup no-op */ }
    else { if (a == 0) exch_receive(s, &ev); else zi_receive(s, a, &ev); }
    s->agent_time[a] += s->comp_delay[a] + s->addl_delay;                                    /* :240-242 / :274-276 */
  }
  if (done) *done = s->loop_done;
  return s->ttl - n0;
}
/* Kernel.runner :310-311 -> TradingAgent.kernelStopping :110-140, ZeroIntelligenceAgent.kernelStopping :80-123 */
void abo_sim_stop(abo_sim *s) {
  for (int id = 1; id < s->n_agents; id++) {
    zi_t *a = &s->zi[id];
    if (a->type != AT_ZI && a->type != AT_HBL) { if (a->type == AT_VALUE) oracle_observe(s, a->current_time, 0, a->rs); a->surplus = 0; continue; }   /* ValueAgent.kernelStopping :57-61 advances the fundamental */
    double hr = nearbyint((double)a->shares / 100.0) * 100.0; /* round(int, -2): half-even on hundreds */
    int64_t H = (int64_t)(hr / 100);
    int64_t rT = oracle_observe(s, a->current_time, 0, a->rs);
    int64_t surplus = 0;
    if (H > 0) for (int64_t x = 1; x <= H; x++) surplus += a->theta[x + a->q_max - 1];
    else if (H < 0) { for (int64_t x = H + 1; x <= 0; x++) surplus += a->theta[x + a->q_max - 1]; surplus = -surplus; }
    surplus += rT * H; surplus += a->cash - a->starting_cash; a->surplus = surplus;
  }
}
int64_t abo_sim_run(abo_sim *s) {
  int done = 0; abo_sim_run_until(s, INT64_MAX, &done); ckpt_push(s); abo_sim_stop(s); return s->ttl;
}

/* ---------------- accessors ---------------- */
int abo_sim_n_agents(abo_sim *s) { return s->n_agents; }
int64_t abo_sim_n_pops(abo_sim *s) { return s->ttl; }
void abo_sim_holdings(abo_sim *s, int64_t *out) {
  for (int id = 1; id < s->n_agents; id++) { zi_t *a = &s->zi[id]; int64_t *r = out + 5 * (id - 1);
    r[0] = id; r[1] = a->shares; r[2] = a->cash; r[3] = a->cash + a->shares * (a->has_last_trade ? a->last_trade : 0); r[4] = a->surplus; } /* markToMarket :609-633 */
}
uint64_t abo_sim_pop_hash(abo_sim *s) { return s->pop_hash; }
int64_t abo_sim_n_hash_ckpt(abo_sim *s) { return s->n_ckpt; }
const uint64_t *abo_sim_hash_ckpt(abo_sim *s) { return s->ckpt; }
uint64_t abo_sim_note_hash(abo_sim *s) { return s->note_hash; }
uint64_t abo_sim_snap_hash(abo_sim *s) { return s->snap_hash; }
int64_t abo_sim_trace(abo_sim *s, int which, const int64_t **rows) {
  i64buf *b = which == 0 ? &s->pops : which == 1 ? &s->ops : which == 2 ? &s->notes : &s->snaps; int w = which == 0 ? 5 : which == 1 ? 9 : which == 2 ? 13 : 16;
  *rows = b->v; return b->n / w;
}
static abo_rng *stream_at(abo_sim *s, int i) {
  if (s->variant == 3 || s->variant == 1) { if (i == 0) return s->sym_rs; if (i == 1) return s->exch_rs; if (i == s->n_agents + 1) return s->kernel_rs; return s->zi[i - 1].rs; }
  int base = s->variant == 100 ? 4 : 3;
  if (i == 0) return s->sym_rs; if (i == 1) return s->kernel_rs;
  if (s->variant == 100) { if (i == 2) return s->lat_rs; if (i == 3) return s->exch_rs; } else if (i == 2) return s->exch_rs;
  return s->zi[i - base + 1].rs;
}
int abo_sim_n_streams(abo_sim *s) { return s->n_agents - 1 + (s->variant == 100 ? 4 : 3); }
int64_t abo_sim_global_tape(abo_sim *s, const uint8_t **k, const uint64_t **b) { *k = s->g->tk; *b = s->g->tv; return s->g->tn; }
void abo_sim_agent_info(abo_sim *s, int id, int64_t *out4) { out4[0] = s->zi[id].type; out4[1] = s->zi[id].size; out4[2] = s->zi[id].wakeup_time; out4[3] = s->zi[id].order_size; }
int64_t abo_sim_tape(abo_sim *s, int i, const uint8_t **k, const uint64_t **b) { abo_rng *r = stream_at(s, i); *k = r->tk; *b = r->tv; return r->tn; }
uint32_t abo_sim_stream_seed(abo_sim *s, int i) { return stream_at(s, i)->seed; }
int64_t abo_sim_global_exp_tape(abo_sim *s, const double **v) { *v = s->gexp; return s->n_gexp; }
void abo_sim_theta(abo_sim *s, int agent, int32_t *out) { memcpy(out, s->zi[agent].theta, sizeof(int32_t) * 2 * s->zi[agent].q_max); }
void abo_sim_latency_vectors(abo_sim *s, double *to_x, double *from_x) { int n = s->n_agents; for (int j = 0; j < n; j++) { to_x[j] = s->latency[(size_t)j * n]; from_x[j] = s->latency[j]; } }
void abo_sim_zi_params(abo_sim *s, int agent, double *out) { zi_t *a = &s->zi[agent]; out[0] = (double)a->R_min; out[1] = (double)a->R_max; out[2] = a->eta; out[3] = a->group; }
int64_t abo_sim_counter(abo_sim *s, int w) {
  switch (w) { case 0: return s->c_limit; case 1: return s->c_cancel; case 2: return s->book.n_fills; case 3: return s->c_query; case 4: return s->max_queue;
    case 5: return s->max_bid_lv; case 6: return s->max_ask_lv; case 7: return s->max_resting; case 8: return s->next_order_id; case 9: return s->uniq; default: return -1; }
}
void abo_sim_book_l1(abo_sim *s, int64_t *o) { int64_t pq[2]; o[0] = o[1] = o[2] = o[3] = 0; if (book_inside(&s->book, 1, 1, pq)) { o[0] = pq[0]; o[1] = pq[1]; } if (book_inside(&s->book, 0, 1, pq)) { o[2] = pq[0]; o[3] = pq[1]; } o[4] = s->book.has_last_trade ? s->book.last_trade : -1; }
int64_t abo_sim_fundamental(abo_sim *s) { return s->or_v; }

/* ====================================================================================================
 * ABIDESEnv: Exchange + MarketReplayAgent + DummyRLExecutionAgent under GymKernel
 *   ABIDESEnv.py:8-103, agent_config.py:42-154, GymKernel.py:25-306,364-389,
 *   agent/examples/MarketReplayAgent.py:14-100, agent/execution/rl/dummy_rl_execution_agent.py,
 *   agent/execution/baselines/execution_agent.py:9-130, ABIDESEnvMetrics.py:9-258
 * Zero latency, zero computation delay, noise [1.0] (no MT draw: randint(0,1)); no oracle (last_trade None until
 * the first trade).  Fully deterministic given the order stream and the actions.
 * ==================================================================================================== */
typedef struct { int64_t t, id, price, size; int is_buy; } srow_t;
typedef struct { int64_t key; int64_t qty, price; int is_buy; int used; } omap_ent;   /* TradingAgent.orders of the replay agent */
typedef struct { omap_ent *e; int cap, n; } omap_t;
static void omap_init(omap_t *m, int cap) { m->cap = 1; while (m->cap < cap * 2) m->cap <<= 1; m->e = (omap_ent *)calloc(m->cap, sizeof(omap_ent)); m->n = 0; }
static omap_ent *omap_find(omap_t *m, int64_t k) { uint64_t h = (uint64_t)k * 0x9E3779B97F4A7C15ULL; int i = (int)(h >> 40) & (m->cap - 1); while (m->e[i].used) { if (m->e[i].used == 1 && m->e[i].key == k) return &m->e[i]; i = (i + 1) & (m->cap - 1); } return NULL; }
static void omap_put(omap_t *m, int64_t k, int64_t qty, int64_t price, int is_buy) { uint64_t h = (uint64_t)k * 0x9E3779B97F4A7C15ULL; int i = (int)(h >> 40) & (m->cap - 1); while (m->e[i].used == 1) i = (i + 1) & (m->cap - 1); m->e[i].key = k; m->e[i].qty = qty; m->e[i].price = price; m->e[i].is_buy = is_buy; m->e[i].used = 1; m->n++; }
static void omap_del(omap_t *m, omap_ent *e) { e->used = 2; m->n--; (void)m; }   /* tombstone */

typedef struct { int64_t bid1, bidq1, bid2, ask1, askq1, ask2, data; int n_bids, n_asks; } lob_t;   /* what the agent reads of a stored QUERY_SPREAD body */

typedef struct {
  /* TradingAgent common */
  int has_open, has_close, mkt_closed; int64_t mkt_open, mkt_close; int64_t shares, cash; int has_last_trade; int64_t last_trade;
} tagent_t;

typedef struct { int64_t p, q; } pq_t;                                                /* one (price, quantity) level of a QUERY_SPREAD reply */
struct dq_state;
struct abo_env {
  heap_t q; int64_t now, stop_time; int64_t agent_time[16]; int64_t comp_delay[16]; int64_t ttl, uniq; int done;
  struct dq_state *dq;                      /* DDQN execution config (execution_marketreplay_ddqn.py): momentum + TWAP + DDQN agents */
  int64_t next_order_id; omap_t used_ids;   /* util/order/Order.py:8-9,35-42 global id allocator (ids seen so far) */
  abo_book book; int64_t mkt_open, mkt_close;
  /* replay agent (id 1) */
  tagent_t ra; srow_t *rows; int64_t n_rows; int64_t *ts; int64_t *ts_first; int64_t n_ts; int64_t wt_cursor; omap_t ra_orders;
  /* RL agent (id 2) */
  tagent_t rl; int rl_state; int rl_trade; double quantity, rem_quantity, executed_sum; int64_t n_executed;
  open_order_t *rl_orders; double *rl_oqty; int rl_n_orders, rl_cap_orders;
  int64_t *horizon; int n_h; int order_level; double steep;
  lob_t lobs[100]; int n_lobs, lob_head; int64_t p0; int metrics_init; int rem_time;
  /* step outputs */
  double obs[9]; int obs_len; int end_step;
  /* traces */
  int trace; i64buf pops, ops, notes, snaps; uint64_t pop_hash, note_hash, snap_hash; uint64_t *ckpt; int64_t n_ckpt, cap_ckpt;
  int64_t max_bid_lv, max_ask_lv, max_resting, max_queue;
};
typedef struct abo_env abo_env;

static void env_put(abo_env *s, const event_t *e) { heap_push(&s->q, e); if (s->q.n > s->max_queue) s->max_queue = s->q.n; }
static void env_set_wakeup(abo_env *s, int sender, int64_t t) { event_t e; memset(&e, 0, sizeof(e)); e.t = t; e.recipient = sender; e.type = ABO_T_WAKEUP; e.uniq = -1; env_put(s, &e); }
static void env_set_cancel(abo_env *s, int sender, int64_t t) { event_t e; memset(&e, 0, sizeof(e)); e.t = t; e.recipient = sender; e.type = ABO_T_CANCEL_ORDER; e.uniq = -1; env_put(s, &e); } /* GymKernel.setCancelOrder :364-389 */
/* Kernel.sendMessage :347-433 with zero latency; noise = choice(1, 1, [1.0])[0] == 0 without an MT draw */
static void env_send(abo_env *s, int sender, int recipient, event_t *e, int64_t delay) {
  e->uniq = s->uniq++; e->t = s->now + s->comp_delay[sender] + delay; e->recipient = recipient; e->type = ABO_T_MESSAGE; e->sender = sender; env_put(s, e);
}
static void env_trace_note(abo_env *s, int recipient, const event_t *e) {
  int64_t row[13] = { s->now, recipient, e->kind, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  if (e->kind == ABO_ORDER_ACCEPTED || e->kind == ABO_ORDER_EXECUTED || e->kind == ABO_ORDER_CANCELLED || e->kind == ABO_ORDER_MODIFIED) {
    row[3] = e->order.order_id; row[4] = e->order.is_buy; row[5] = e->order.quantity; row[6] = e->order.limit_price; row[7] = e->order.fill_price; }
  if (e->kind == ABO_QUERY_SPREAD) { row[7] = e->data; if (e->has_bid) { row[8] = e->bid; row[9] = e->bid_q; } if (e->has_ask) { row[10] = e->ask; row[11] = e->ask_q; } row[12] = e->mkt_closed; }
  for (int i = 0; i < 13; i++) s->note_hash = fnv_mix(s->note_hash, row[i]);
  if (s->trace & ABO_TRACE_NOTES) ib_push(&s->notes, row, 13);
}
static void env_exch_send(abo_env *s, int recipient, event_t *e) { /* ExchangeAgent.sendMessage :471-485 (pipeline delay 0) */
  /* uniq is taken at Message construction, which precedes the trace hook in the recorder as well */
  int64_t u = s->uniq; env_trace_note(s, recipient, e); (void)u; env_send(s, 0, recipient, e, 0);
}
static void env_book_send(void *owner, int64_t recipient, int kind, const order_t *o) { abo_env *s = (abo_env *)owner; event_t e; memset(&e, 0, sizeof(e)); e.kind = kind; e.order = *o; env_exch_send(s, (int)recipient, &e); }
static void env_trace_snap(abo_env *s) {
  int64_t row[16]; memset(row, 0, sizeof(row));
  int nb = s->book.bids.n, na = s->book.asks.n; int64_t rest = abo_book_n_resting(&s->book);
  row[0] = nb; row[1] = na; row[2] = rest; book_inside(&s->book, 1, 3, row + 3); book_inside(&s->book, 0, 3, row + 9);
  row[15] = s->book.has_last_trade ? s->book.last_trade : -1;
  if (nb > s->max_bid_lv) s->max_bid_lv = nb; if (na > s->max_ask_lv) s->max_ask_lv = na; if (rest > s->max_resting) s->max_resting = rest;
  for (int i = 0; i < 16; i++) s->snap_hash = fnv_mix(s->snap_hash, row[i]);
  if (s->trace & ABO_TRACE_SNAPS) ib_push(&s->snaps, row, 16);
}
static void env_trace_op(abo_env *s, int op, const order_t *o, int64_t np, int64_t nq) {
  if (!(s->trace & ABO_TRACE_OPS)) return;
  int64_t row[9] = { s->now, op, o->agent_id, o->order_id, o->is_buy, o->limit_price, o->quantity, np, nq }; ib_push(&s->ops, row, 9);
}
static void dq_snapshot(abo_env *s, int recipient);
/* ExchangeAgent.receiveMessage :129-340 (oracle None, depth-k QUERY_SPREAD, MODIFY_ORDER) */
static void env_exch_receive(abo_env *s, const event_t *m) {
  s->comp_delay[0] = 0;
  int t_closed = s->now > s->mkt_close;
  if (t_closed) {
    int is_order = m->kind == ABO_LIMIT_ORDER || m->kind == ABO_CANCEL_ORDER || m->kind == ABO_MODIFY_ORDER;
    int is_query = m->kind == ABO_QUERY_SPREAD;
    if (is_order || !is_query) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_MKT_CLOSED; env_exch_send(s, m->sender, &e); return; }
  }
  event_t e; memset(&e, 0, sizeof(e)); s->book.now = s->now;
  switch (m->kind) {
    case ABO_WHEN_MKT_OPEN: e.kind = ABO_WHEN_MKT_OPEN; e.data = s->mkt_open; env_exch_send(s, m->sender, &e); break;
    case ABO_WHEN_MKT_CLOSE: e.kind = ABO_WHEN_MKT_CLOSE; e.data = s->mkt_close; env_exch_send(s, m->sender, &e); break;
    case ABO_QUERY_SPREAD: {                                                                  /* :215-245 */
      int64_t pq[4]; e.kind = ABO_QUERY_SPREAD;
      e.n_bids = book_inside(&s->book, 1, 2, pq); if (e.n_bids > 0) { e.has_bid = 1; e.bid = pq[0]; e.bid_q = pq[1]; } if (e.n_bids > 1) e.bid2 = pq[2];
      e.n_asks = book_inside(&s->book, 0, 2, pq); if (e.n_asks > 0) { e.has_ask = 1; e.ask = pq[0]; e.ask_q = pq[1]; } if (e.n_asks > 1) e.ask2 = pq[2];
      e.data = s->book.has_last_trade ? s->book.last_trade : -1; e.mkt_closed = t_closed;
      if (s->dq) dq_snapshot(s, m->sender);                                                   /* depth-500 lists ride in the reply body */
      env_exch_send(s, m->sender, &e); break; }
    case ABO_LIMIT_ORDER: env_trace_op(s, 0, &m->order, 0, 0); book_handle_limit(&s->book, m->order); env_trace_snap(s); break;
    case ABO_CANCEL_ORDER: env_trace_op(s, 1, &m->order, 0, 0); book_cancel(&s->book, &m->order); env_trace_snap(s); break;
    case ABO_MODIFY_ORDER: env_trace_op(s, 2, &m->order, m->new_order.limit_price, m->new_order.quantity); book_modify(&s->book, &m->order, &m->new_order); env_trace_snap(s); break; /* :326-340 */
    default: break;
  }
}
/* Order.__init__ / generateOrderId (util/order/Order.py:27,35-42) */
static int64_t env_new_order_id(abo_env *s, int64_t given) {
  int64_t id;
  if (given) id = given;
  else { while (omap_find(&s->used_ids, s->next_order_id)) s->next_order_id++; id = s->next_order_id; }
  if (!omap_find(&s->used_ids, id)) { if (s->used_ids.n * 2 >= s->used_ids.cap) { omap_t old = s->used_ids; omap_init(&s->used_ids, old.cap); for (int i = 0; i < old.cap; i++) if (old.e[i].used == 1) omap_put(&s->used_ids, old.e[i].key, 0, 0, 0); free(old.e); } omap_put(&s->used_ids, id, 0, 0, 0); }
  return id;
}
/* TradingAgent.wakeup :142-158: returns can_trade */
static int env_ta_wakeup(abo_env *s, int id, tagent_t *a) {
  if (!a->has_open) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_OPEN; env_send(s, id, 0, &e, 0); memset(&e, 0, sizeof(e)); e.kind = ABO_WHEN_MKT_CLOSE; env_send(s, id, 0, &e, 0); }
  return a->has_open && a->has_close && !a->mkt_closed;
}
/* TradingAgent.receiveMessage :181-268 common part; returns 1 when the market hours just became known */
static int env_ta_receive(abo_env *s, tagent_t *a, const event_t *m) {
  (void)s; int had = a->has_open && a->has_close;
  switch (m->kind) {
    case ABO_WHEN_MKT_OPEN: a->mkt_open = m->data; a->has_open = 1; break;
    case ABO_WHEN_MKT_CLOSE: a->mkt_close = m->data; a->has_close = 1; break;
    case ABO_ORDER_EXECUTED: { int64_t qty = m->order.is_buy ? m->order.quantity : -m->order.quantity; a->shares += qty; a->cash -= qty * m->order.fill_price; break; }
    case ABO_MKT_CLOSED: a->mkt_closed = 1; break;
    case ABO_QUERY_SPREAD: if (m->mkt_closed) a->mkt_closed = 1; a->last_trade = m->data; a->has_last_trade = 1; break;
    default: break;
  }
  return a->has_open && a->has_close && !had;
}
/* MarketReplayAgent.placeOrder :69-96 for one row */
static void replay_place(abo_env *s, const srow_t *r) {
  omap_ent *ex = omap_find(&s->ra_orders, r->id);
  if (!ex && r->size > 0) {                                                          /* placeLimitOrder(order_id=ORDER_ID) TradingAgent.py:309-349 */
    int64_t oid = env_new_order_id(s, r->id);
    omap_put(&s->ra_orders, oid, r->size, r->price, r->is_buy);
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_LIMIT_ORDER; e.order.agent_id = 1; e.order.order_id = oid; e.order.quantity = r->size; e.order.limit_price = r->price; e.order.is_buy = r->is_buy;
    env_send(s, 1, 0, &e, 0);
  } else if (ex && r->size == 0) {                                                   /* cancelOrder :399-406 */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_CANCEL_ORDER; e.order.agent_id = 1; e.order.order_id = ex->key; e.order.quantity = ex->qty; e.order.limit_price = ex->price; e.order.is_buy = ex->is_buy;
    env_send(s, 1, 0, &e, 0);
  } else if (ex) {                                                                   /* modifyOrder :408-418 */
    int64_t nid = env_new_order_id(s, r->id);                                        /* LimitOrder(..., order_id=order_id): id 0 is "unset" -> a fresh generated id (Order.py:27) */
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_MODIFY_ORDER;
    e.order.agent_id = 1; e.order.order_id = ex->key; e.order.quantity = ex->qty; e.order.limit_price = ex->price; e.order.is_buy = ex->is_buy;
    e.new_order.agent_id = 1; e.new_order.order_id = nid; e.new_order.quantity = r->size; e.new_order.limit_price = r->price; e.new_order.is_buy = r->is_buy;
    env_send(s, 1, 0, &e, 0);
  }
}
/* MarketReplayAgent.wakeup :50-60 */
static void replay_wakeup(abo_env *s) {
  tagent_t *a = &s->ra; env_ta_wakeup(s, 1, a);
  if (!a->has_open || !a->has_close) return;
  if (s->wt_cursor >= s->n_ts) return;                                               /* wakeup_times[0] -> IndexError: nothing placed */
  env_set_wakeup(s, 1, s->ts[s->wt_cursor]); s->wt_cursor++;                         /* setWakeup(wakeup_times[0]); pop(0) */
  /* orders_dict[currentTime]: rows of the timestamp equal to now */
  int64_t lo = 0, hi = s->n_ts - 1, k = -1;
  while (lo <= hi) { int64_t mid = (lo + hi) / 2; if (s->ts[mid] == s->now) { k = mid; break; } if (s->ts[mid] < s->now) lo = mid + 1; else hi = mid - 1; }
  if (k < 0) return;                                                                 /* KeyError would propagate in the reference; cannot happen */
  int64_t r0 = s->ts_first[k], r1 = (k + 1 < s->n_ts) ? s->ts_first[k + 1] : s->n_rows;
  for (int64_t r = r0; r < r1; r++) replay_place(s, &s->rows[r]);
}
static void replay_receive(abo_env *s, const event_t *m) {
  tagent_t *a = &s->ra;
  int newly = env_ta_receive(s, a, m);
  if (m->kind == ABO_ORDER_EXECUTED) {                                               /* orderExecuted :422-462 */
    omap_ent *o = omap_find(&s->ra_orders, m->order.order_id);
    if (o) { if (m->order.quantity >= o->qty) omap_del(&s->ra_orders, o); else o->qty -= m->order.quantity; }
    a->last_trade = m->order.fill_price; a->has_last_trade = 1;                      /* MarketReplayAgent.receiveMessage :62-67 */
  } else if (m->kind == ABO_ORDER_CANCELLED) { omap_ent *o = omap_find(&s->ra_orders, m->order.order_id); if (o) omap_del(&s->ra_orders, o); }
  if (newly) env_set_wakeup(s, 1, a->mkt_open + (s->ts[0] - a->mkt_open));           /* getWakeFrequency :98-100 */
}
/* ---- DummyRLExecutionAgent ---- */
static void rl_orders_remove(abo_env *s, int i) { memmove(s->rl_orders + i, s->rl_orders + i + 1, sizeof(open_order_t) * (s->rl_n_orders - i - 1)); memmove(s->rl_oqty + i, s->rl_oqty + i + 1, sizeof(double) * (s->rl_n_orders - i - 1)); s->rl_n_orders--; }
static void rl_get_spread(abo_env *s) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_SPREAD; env_send(s, 2, 0, &e, 0); s->uniq++; }  /* getCurrentSpread(depth=500) + msg_copy */
/* wakeup :184-218 */
static void rl_wakeup(abo_env *s) {
  if (!env_ta_wakeup(s, 2, &s->rl)) return;
  if (s->rl_trade) { int k = -1; for (int i = 0; i < s->n_h; i++) if (s->horizon[i] > s->now) { k = i; break; }
    if (k >= 0) env_set_cancel(s, 2, s->horizon[k]); else s->rl_trade = 0; }                         /* requestedTime - Timedelta(0.5) == requestedTime */
  if (s->rl_trade) { int k = -1; for (int i = 0; i < s->n_h - 1; i++) if (s->horizon[i] > s->now) { k = i; break; }
    if (k >= 0) env_set_wakeup(s, 2, s->horizon[k]); else s->rl_trade = 0; }
  rl_get_spread(s); s->rl_state = ST_AWAITING_SPREAD;
}
/* receiveMessage :230-245 (+ ExecutionAgent.handleOrderExecution :88-99, ABIDESEnvMetrics.addLOB :62-84) */
static void rl_receive(abo_env *s, const event_t *m) {
  tagent_t *a = &s->rl;
  int newly = env_ta_receive(s, a, m);
  if (m->kind == ABO_ORDER_EXECUTED) {
    for (int i = 0; i < s->rl_n_orders; i++) if (s->rl_orders[i].order_id == m->order.order_id) {
      if ((double)m->order.quantity >= s->rl_oqty[i]) rl_orders_remove(s, i); else s->rl_oqty[i] -= (double)m->order.quantity; break; }
    s->executed_sum += (double)m->order.quantity; s->n_executed++; s->rem_quantity = s->quantity - s->executed_sum;
  } else if (m->kind == ABO_ORDER_CANCELLED) {
    for (int i = 0; i < s->rl_n_orders; i++) if (s->rl_orders[i].order_id == m->order.order_id) { rl_orders_remove(s, i); break; }
  }
  if (newly) env_set_wakeup(s, 2, a->mkt_open + (s->horizon[0] - a->mkt_open));               /* ExecutionAgent.getWakeFrequency :129-130 */
  if (s->rem_quantity > 0 && s->rl_state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) {
    s->rl_state = ST_AWAITING_WAKEUP;
    if (!s->metrics_init) { s->p0 = m->data; s->metrics_init = 1; }
    s->lob_head = (s->lob_head + 99) % 100;                                                   /* deque.appendleft, maxlen 100 */
    lob_t *l = &s->lobs[s->lob_head]; l->bid1 = m->bid; l->bidq1 = m->bid_q; l->bid2 = m->bid2; l->ask1 = m->ask; l->askq1 = m->ask_q; l->ask2 = m->ask2; l->data = m->data; l->n_bids = m->n_bids; l->n_asks = m->n_asks;
    if (s->n_lobs < 100) s->n_lobs++;
  }
}
static const lob_t *rl_lob(const abo_env *s, int idx) { if (idx < 0) idx += s->n_lobs; return &s->lobs[(s->lob_head + idx) % 100]; }   /* deque[idx], negative from the right */
/* numpy's pairwise summation for n <= 128 (numpy/core/src/umath/loops_utils.h.src pairwise_sum): np.mean / np.std use it */
static double np_pairwise_sum(const double *a, int n) {
  if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r += a[i]; return r; }
  double r[8]; for (int k = 0; k < 8; k++) r[k] = a[k];
  int i; for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; k++) r[k] += a[i + k];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; i++) res += a[i];
  return res;
}
/* get_observation :294-315; returns 0 on the ValueError paths of ABIDESEnvMetrics (empty side / no LOB) */
static int rl_observe(abo_env *s) {
  /* get_remaining_time :282-292 */
  int64_t curr = s->now - (s->now % (30 * NS_PER_S)); int rem = s->n_h;
  for (int i = 0; i < s->n_h; i++) if (s->horizon[i] == curr) { rem = s->n_h - 1 - i; break; }
  s->rem_time = rem; s->obs_len = 0;
  if (s->n_lobs == 0) return 0;
  const lob_t *l = rl_lob(s, 0);
  if (l->n_bids == 0 || l->n_asks == 0) return 0;
  double p0 = (double)s->p0, pt = (double)l->data;
  double mid = ((double)l->bid1 + (double)l->ask1) / 2;
  double *o = s->obs;
  o[0] = rem; o[1] = s->rem_quantity;
  o[2] = log(pt / p0);                                                                        /* getLogReturn */
  o[3] = (double)(l->ask1 - l->bid1);                                                         /* getBidAskSpread */
  o[4] = ((double)l->bidq1 - (double)l->askq1) / ((double)l->bidq1 + (double)l->askq1);       /* getVolImbalance level 1 */
  o[5] = tanh((double)l->ask1 / (double)l->askq1 - (double)l->bid1 / (double)l->bidq1);       /* getSmartPrice */
  { double v[100], mean = 0, var = 0; int n = s->n_lobs;                                      /* getMidPriceVolatility: np.std(ddof=0) */
    for (int i = 0; i < n; i++) { const lob_t *li = rl_lob(s, i); if (li->n_bids == 0 || li->n_asks == 0) return 0; v[i] = log((((double)li->bid1 + (double)li->ask1) / 2) / p0); mean += v[i]; }
    mean = np_pairwise_sum(v, n) / n; for (int i = 0; i < n; i++) v[i] = (v[i] - mean) * (v[i] - mean); var = np_pairwise_sum(v, n); o[6] = sqrt(var / n); }
  int d;                                                                                      /* getTradeDirection :197-216 */
  if (pt > mid) d = 1; else if (pt < mid) d = -1;
  else { const lob_t *ll = rl_lob(s, -1); if (ll->n_bids == 0 || ll->n_asks == 0) return 0; double lm = ((double)ll->bid1 + (double)ll->ask1) / 2; d = mid > lm ? 1 : -1; }
  o[7] = d; o[8] = 2 * d * (pt - mid) / mid;                                                  /* getEffectiveSpread */
  s->obs_len = 9; return 1;
}
/* place_orders :159-181 + process_action :138-157 */
static void rl_place_orders(abo_env *s, const double *action) {
  int n = s->order_level; double q0 = s->quantity, q = s->quantity /* metrics.rem_quantity is never updated (typo, App. A-10) */;
  double sum = 0; for (int i = 0; i < n; i++) sum = sum + action[1 + i];
  double o_hat[8], o[8];
  for (int i = 0; i < n; i++) o_hat[i] = sum == 0.0 ? 1.0 / n : action[1 + i] / sum;
  double q_hat = q / q0;
  double o_total = nearbyint(q0 * q_hat * pow(action[0], pow(q_hat, s->steep)));
  double part = 0; for (int i = 0; i < n; i++) o[i] = nearbyint(o_total * o_hat[i]);
  for (int i = 0; i < n - 1; i++) part = part + o[i];
  o[n - 1] = o_total - part;
  for (int l = 0; l < n; l++) {
    if (s->n_lobs == 0) continue;                                                             /* exception swallowed */
    const lob_t *lb = rl_lob(s, 0);
    if (lb->n_bids == 0 || lb->n_asks == 0) continue;                                         /* ValueError swallowed */
    if (l >= lb->n_bids || l >= lb->n_asks) continue;                                         /* IndexError swallowed (needs both sides at this level) */
    int64_t price = l == 0 ? lb->bid1 : lb->bid2;                                             /* BUY: bid price of level l+1 */
    int64_t oid = env_new_order_id(s, 0);                                                     /* LimitOrder() is constructed before the qty test */
    if (!(o[l] > 0)) continue;
    if (s->rl_n_orders == s->rl_cap_orders) { s->rl_cap_orders = s->rl_cap_orders ? 2 * s->rl_cap_orders : 8; s->rl_orders = (open_order_t *)realloc(s->rl_orders, sizeof(open_order_t) * s->rl_cap_orders); s->rl_oqty = (double *)realloc(s->rl_oqty, sizeof(double) * s->rl_cap_orders); }
    open_order_t *oo = &s->rl_orders[s->rl_n_orders]; oo->order_id = oid; oo->limit_price = price; oo->is_buy = 1; oo->quantity = (int64_t)o[l]; s->rl_oqty[s->rl_n_orders] = o[l]; s->rl_n_orders++;
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_LIMIT_ORDER; e.order.agent_id = 2; e.order.order_id = oid; e.order.quantity = (int64_t)o[l]; e.order.limit_price = price; e.order.is_buy = 1;
    env_send(s, 2, 0, &e, 0);
  }
}
/* cancelAllOrders :247-255 */
static void rl_cancel_all(abo_env *s) {
  for (int i = 0; i < s->rl_n_orders; i++) {
    event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_CANCEL_ORDER; e.order.agent_id = 2; e.order.order_id = s->rl_orders[i].order_id; e.order.quantity = (int64_t)s->rl_oqty[i]; e.order.limit_price = s->rl_orders[i].limit_price; e.order.is_buy = 1;
    env_send(s, 2, 0, &e, 0);
  }
}

abo_env *abo_env_new2(const int64_t *stream5, int64_t n_rows, double quantity, int order_level, int trace, int64_t stop_ns) {
  abo_env *s = (abo_env *)calloc(1, sizeof(abo_env));
  s->trace = trace; s->pop_hash = s->note_hash = s->snap_hash = FNV_OFF;
  s->mkt_open = (9 * 3600 + 30 * 60) * NS_PER_S; s->mkt_close = 16 * 3600 * NS_PER_S; s->stop_time = stop_ns;   /* ABIDESEnv.py:87-88 16:10; config/marketreplay.py:135 16:01 */
  book_init(&s->book, 10, s, env_book_send);
  s->rows = (srow_t *)malloc(sizeof(srow_t) * (n_rows > 0 ? n_rows : 1)); s->n_rows = n_rows;
  s->ts = (int64_t *)malloc(8 * (n_rows + 1)); s->ts_first = (int64_t *)malloc(8 * (n_rows + 1));
  for (int64_t i = 0; i < n_rows; i++) { srow_t *r = &s->rows[i]; r->t = stream5[5 * i]; r->id = stream5[5 * i + 1]; r->price = stream5[5 * i + 2]; r->size = stream5[5 * i + 3]; r->is_buy = (int)stream5[5 * i + 4];
    if (i == 0 || r->t != s->rows[i - 1].t) { s->ts[s->n_ts] = r->t; s->ts_first[s->n_ts] = i; s->n_ts++; } }
  omap_init(&s->ra_orders, (int)n_rows + 16); omap_init(&s->used_ids, 1 << 12);
  s->quantity = s->rem_quantity = quantity; s->order_level = order_level; s->steep = 0.5; s->rl_trade = 1; s->rl_state = ST_AWAITING_WAKEUP;
  s->n_h = 761; s->horizon = (int64_t *)malloc(8 * s->n_h);                                   /* date_range(09:40, 16:00, freq 30s)  agent_config.py:134-136 */
  for (int i = 0; i < s->n_h; i++) s->horizon[i] = (9 * 3600 + 40 * 60) * NS_PER_S + (int64_t)i * 30 * NS_PER_S;
  s->wt_cursor = 1;                                                                           /* wakeup_times = [*orders_dict]; first_wakeup stays in the list */
  s->wt_cursor = 0;
  for (int i = 0; i < (order_level > 0 ? 3 : 2); i++) env_set_wakeup(s, i, 0);               /* initRunner :139-146 kernelStarting; order_level 0: no RL agent (config/marketreplay.py) */
  return s;
}
abo_env *abo_env_new(const int64_t *stream5, int64_t n_rows, double quantity, int order_level, int trace) { return abo_env_new2(stream5, n_rows, quantity, order_level, trace, (16 * 3600 + 600) * NS_PER_S); }
static void dq_free(abo_env *s);
void abo_env_free(abo_env *s) {
  if (!s) return; dq_free(s); book_destroy(&s->book); free(s->rows); free(s->ts); free(s->ts_first); free(s->ra_orders.e); free(s->used_ids.e); free(s->rl_orders); free(s->rl_oqty);
  free(s->horizon); free(s->q.e); free(s->pops.v); free(s->ops.v); free(s->notes.v); free(s->snaps.v); free(s->ckpt); free(s);
}
/* GymKernel.stepRunner :158-306.  Returns obs length (0 or 9); *done as ABIDESEnv.step computes it (ABIDESEnv.py:42-46). */
int abo_env_step(abo_env *s, const double *action, double *obs_out, int *done) {
  if (s->order_level > 0) rl_place_orders(s, action);
  s->end_step = 0;
  while (!s->end_step && s->q.n > 0 && s->now <= s->stop_time) {
    event_t ev; heap_pop(&s->q, &ev); s->now = ev.t; s->ttl++;
    s->pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s->pop_hash, ev.t), ev.recipient), ev.type), ev.uniq);
    if (s->ttl % 1000 == 0) { if (s->n_ckpt == s->cap_ckpt) { s->cap_ckpt = s->cap_ckpt ? s->cap_ckpt * 2 : 256; s->ckpt = (uint64_t *)realloc(s->ckpt, 8 * s->cap_ckpt); } s->ckpt[s->n_ckpt++] = s->pop_hash; }
    if (s->trace & ABO_TRACE_POPS) { int64_t row[5] = { ev.t, ev.recipient, ev.type, ev.uniq, ev.kind }; ib_push(&s->pops, row, 5); }
    int a = ev.recipient;
    if (ev.type == ABO_T_CANCEL_ORDER) { rl_cancel_all(s); continue; }                        /* :241-245 (get_reward returns None) */
    if (s->agent_time[a] > s->now) { ev.t = s->agent_time[a]; env_put(s, &ev); continue; }
    s->agent_time[a] = s->now;
    if (ev.type == ABO_T_WAKEUP) { if (a == 1) replay_wakeup(s); else if (a == 2) rl_wakeup(s); }
    else {
      if (a == 0) env_exch_receive(s, &ev); else if (a == 1) replay_receive(s, &ev); else rl_receive(s, &ev);
      if (a == 2 && ev.kind == ABO_QUERY_SPREAD) { rl_observe(s); s->end_step = 1; }          /* :286-289 */
    }
    s->agent_time[a] += s->comp_delay[a];
  }
  *done = !(s->q.n > 0 && s->now <= s->stop_time);
  for (int i = 0; i < s->obs_len; i++) obs_out[i] = s->obs[i];
  return s->obs_len;
}
int64_t abo_env_n_pops(abo_env *s) { return s->ttl; }
uint64_t abo_env_pop_hash(abo_env *s) { return s->pop_hash; }
uint64_t abo_env_note_hash(abo_env *s) { return s->note_hash; }
uint64_t abo_env_snap_hash(abo_env *s) { return s->snap_hash; }
int64_t abo_env_n_hash_ckpt(abo_env *s) { return s->n_ckpt; }
const uint64_t *abo_env_hash_ckpt(abo_env *s) { return s->ckpt; }
int64_t abo_env_trace(abo_env *s, int which, const int64_t **rows) { i64buf *b = which == 0 ? &s->pops : which == 1 ? &s->ops : which == 2 ? &s->notes : &s->snaps; int w = which == 0 ? 5 : which == 1 ? 9 : which == 2 ? 13 : 16; *rows = b->v; return b->n / w; }
void abo_env_final(abo_env *s, double *out8) { out8[0] = s->rem_quantity; out8[1] = (double)s->rl.shares; out8[2] = (double)s->rl.cash; out8[3] = (double)s->n_executed; out8[4] = (double)s->ra.shares; out8[5] = (double)s->ra.cash; out8[6] = (double)s->ra_orders.n; out8[7] = (double)s->now; }
int64_t abo_env_counter(abo_env *s, int w) { switch (w) { case 0: return s->max_queue; case 1: return s->max_bid_lv; case 2: return s->max_ask_lv; case 3: return s->max_resting; case 4: return s->uniq; case 5: return s->next_order_id; default: return -1; } }

/* ====================================================================================================
 * DDQN execution config: config/execution/marketreplay/execution_marketreplay_ddqn.py (-a rl)
 *   Exchange (0) + MarketReplayAgent (1) + n_mom MomentumAgents (2..) + TWAPExecutionAgent + DDQLearningExecutionAgent
 *   under Kernel.runner, zero latency / computation delay, stop = horizon end + 10 min (:317-318).
 *   agent/examples/MomentumAgent.py:53-99, agent/execution/baselines/execution_agent.py:66-130, twap_agent.py:51-63,
 *   agent/execution/qlearning/ddqlearning_execution_agent.py:20-37,141-185,228-447,507-611, agent/execution/util.py:6-42,
 *   agent/TradingAgent.py:351-397 (placeMarketOrder), :564-574 (getKnownBidAsk).
 * The Q-network is NOT part of this restatement: the action of every decision tick is an input (the agent's
 * choose_action is np.argmax of the network output, :362-364).  The event loop pauses where the reference calls
 * choose_action (inside place_order :245) and resumes with the supplied action.
 * ==================================================================================================== */
typedef struct { tagent_t ta; int state; int64_t size; int has_bid, has_ask; int64_t bid, ask; double *mids; int n_mids, cap_mids; double avg20, avg50; int has20, has50; } mom_t;
typedef struct {
  tagent_t ta; int id, is_ddqn, state, trade; int64_t quantity, rem_qty, executed_sum, n_executed;
  open_order_t *orders; int n_orders, cap_orders;                                             /* self.orders, insertion ordered */
  pq_t *kb, *ka; int nkb, nka, cap_k;                                                        /* known_bids / known_asks */
  pq_t *fb, *fa; int nfb, nfa;                                                               /* lists of the in-flight QUERY_SPREAD reply */
  double arrival; int has_arrival; int64_t child_qty; int64_t *sched; int n_sched;   /* sched: per-bin child quantities (VWAPExecutionAgent.generate_schedule, vwap_agent.py:48-62); NULL = one quantity (TWAP) */
  int t, rem_time, pending, cur_s[2], sp[2]; double obs6[6];                                  /* DDQN: self.t, remaining_time, decision pending, self.s */
  double *pp; int n_pp, cap_pp; double *exp; int n_exp, cap_exp; double *rew; int n_rew, cap_rew; double *ahist; int n_ahist, cap_ahist;
  double step_reward;                                                                         /* sum of step rewards since the last decision */
} exec_t;
struct dq_state { int n_mom, n_twap, has_ddqn, n_agents; mom_t mom[8]; exec_t ex[3]; int n_ex; int64_t h0, h_step; int n_h; int64_t mom_wake_ns; int is_buy; int error; };
typedef struct dq_state dq_state;
enum { DQ_DEPTH = 500, DQ_ERR_EMPTY_SIDE = 1, DQ_ERR_LEVELS = 2 };

static exec_t *dq_exec(abo_env *s, int id) { dq_state *d = s->dq; int k = id - (2 + d->n_mom); return (k >= 0 && k < d->n_ex) ? &d->ex[k] : NULL; }
static void dq_snapshot(abo_env *s, int recipient) {                                          /* ExchangeAgent.py:231-245 with depth = 500 */
  exec_t *e = dq_exec(s, recipient); if (!e) return;
  int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * 2 * DQ_DEPTH);
  e->nfb = book_inside(&s->book, 1, DQ_DEPTH, tmp); for (int i = 0; i < e->nfb; i++) { e->fb[i].p = tmp[2 * i]; e->fb[i].q = tmp[2 * i + 1]; }
  e->nfa = book_inside(&s->book, 0, DQ_DEPTH, tmp); for (int i = 0; i < e->nfa; i++) { e->fa[i].p = tmp[2 * i]; e->fa[i].q = tmp[2 * i + 1]; }
  free(tmp);
}
static void dq_get_spread(abo_env *s, int id) { event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_QUERY_SPREAD; env_send(s, id, 0, &e, 0); s->uniq++; }   /* getCurrentSpread + msg_copy (TradingAgent.py:277-282) */
/* TradingAgent.placeLimitOrder :309-349 */
static void dq_place_limit(abo_env *s, int id, open_order_t **orders, int *n, int *cap, int64_t qty, int is_buy, int64_t price) {
  int64_t oid = env_new_order_id(s, 0);
  if (qty <= 0) return;
  if (orders) { if (*n == *cap) { *cap = *cap ? 2 * *cap : 16; *orders = (open_order_t *)realloc(*orders, sizeof(open_order_t) * *cap); }
    open_order_t *o = &(*orders)[(*n)++]; o->order_id = oid; o->quantity = qty; o->limit_price = price; o->is_buy = is_buy; }
  event_t e; memset(&e, 0, sizeof(e)); e.kind = ABO_LIMIT_ORDER; e.order.agent_id = id; e.order.order_id = oid; e.order.quantity = qty; e.order.limit_price = price; e.order.is_buy = is_buy;
  env_send(s, id, 0, &e, 0);
}
/* ---- MomentumAgent ---- */
static void dq_mom_wakeup(abo_env *s, int id) { mom_t *a = &s->dq->mom[id - 2]; if (env_ta_wakeup(s, id, &a->ta)) { dq_get_spread(s, id); a->state = ST_AWAITING_SPREAD; } }
static void dq_mom_receive(abo_env *s, int id, const event_t *m) {
  mom_t *a = &s->dq->mom[id - 2];
  int newly = env_ta_receive(s, &a->ta, m);
  if (m->kind == ABO_QUERY_SPREAD) { a->has_bid = m->has_bid; a->has_ask = m->has_ask; a->bid = m->bid; a->ask = m->ask; }
  if (newly) env_set_wakeup(s, id, a->ta.mkt_open + s->dq->mom_wake_ns);                      /* getWakeFrequency :89-90 */
  if (a->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) {                        /* :65-71 */
    if (a->has_bid && a->bid != 0 && a->has_ask && a->ask != 0) {                             /* placeOrders :78-93 */
      if (a->n_mids == a->cap_mids) { a->cap_mids = a->cap_mids ? 2 * a->cap_mids : 64; a->mids = (double *)realloc(a->mids, 8 * a->cap_mids); }
      a->mids[a->n_mids++] = (double)(a->bid + a->ask) / 2;
      int L = a->n_mids;
      for (int w = 0; w < 2; w++) { int n = w ? 50 : 20;
        if (L > n) { double c1 = 0, c0 = 0; for (int i = 0; i < L; i++) { c1 += a->mids[i]; if (i == L - 1 - n) c0 = c1; }
          double v = np_round2((c1 - c0) / n); if (w) { a->avg50 = v; a->has50 = 1; } else { a->avg20 = v; a->has20 = 1; } } }
      if (a->has20 && a->has50) { if (a->avg20 >= a->avg50) dq_place_limit(s, id, NULL, NULL, NULL, a->size, 1, a->ask); else dq_place_limit(s, id, NULL, NULL, NULL, a->size, 0, a->bid); }
    }
    env_set_wakeup(s, id, s->now + s->dq->mom_wake_ns); a->state = ST_AWAITING_WAKEUP;
  }
}
/* ---- execution agents: shared parts ---- */
static void dq_exec_orders_remove(exec_t *e, int i) { memmove(e->orders + i, e->orders + i + 1, sizeof(open_order_t) * (e->n_orders - i - 1)); e->n_orders--; }
static void dq_exec_cancel_all(abo_env *s, exec_t *e) {                                       /* cancelOrders / cancel_orders */
  for (int i = 0; i < e->n_orders; i++) { event_t ev; memset(&ev, 0, sizeof(ev)); ev.kind = ABO_CANCEL_ORDER; ev.order.agent_id = e->id; ev.order.order_id = e->orders[i].order_id;
    ev.order.quantity = e->orders[i].quantity; ev.order.limit_price = e->orders[i].limit_price; ev.order.is_buy = e->orders[i].is_buy; env_send(s, e->id, 0, &ev, 0); }
}
static void dq_place_market(abo_env *s, exec_t *e, int64_t quantity) {                         /* TradingAgent.placeMarketOrder :351-397 */
  if (quantity <= 0) return;
  const pq_t *side = s->dq->is_buy ? e->ka : e->kb; int n = s->dq->is_buy ? e->nka : e->nkb;
  if (n == 0) { s->dq->error |= DQ_ERR_EMPTY_SIDE; return; }                                  /* the reference iterates None: TypeError */
  int nq = 0; int64_t *qp = (int64_t *)malloc(16 * (n + 1));
  for (int i = 0; i < n; i++) { if (quantity <= side[i].q) { qp[2 * nq] = side[i].p; qp[2 * nq + 1] = quantity; nq++; break; } qp[2 * nq] = side[i].p; qp[2 * nq + 1] = side[i].q; nq++; quantity -= side[i].q; }
  for (int i = 0; i < nq; i++) dq_place_limit(s, e->id, &e->orders, &e->n_orders, &e->cap_orders, qp[2 * i + 1], s->dq->is_buy, qp[2 * i]);
  free(qp);
}
static int dq_horizon_index(const dq_state *d, int64_t t) { if (t < d->h0 || (t - d->h0) % d->h_step) return -1; int64_t k = (t - d->h0) / d->h_step; return k < d->n_h ? (int)k : -1; }
static void dq_exec_wakeup(abo_env *s, exec_t *e) {
  dq_state *d = s->dq;
  if (!env_ta_wakeup(s, e->id, &e->ta)) return;
  int64_t k = s->now < d->h0 ? 0 : (s->now - d->h0) / d->h_step + 1;                          /* first horizon time > now */
  if (e->is_ddqn) {                                                                           /* ddqlearning_execution_agent.py:141-153 */
    if (e->trade) { if (k < d->n_h) env_set_wakeup(s, e->id, d->h0 + k * d->h_step); else e->trade = 0; }
    dq_get_spread(s, e->id); e->state = ST_AWAITING_SPREAD;
  } else if (e->trade) {                                                                      /* execution_agent.py:66-78 */
    if (k < d->n_h) env_set_wakeup(s, e->id, d->h0 + k * d->h_step);
    dq_get_spread(s, e->id); e->state = ST_AWAITING_SPREAD;
  }
}
/* np.digitize(x, np.linspace(0, 1, 201)[1:-1]): number of split points k * (1/200), k = 1..199, that are <= x (agent/execution/util.py:6-42) */
static int dq_digitize(double x) { int k = 0; while (k < 199 && (double)(k + 1) * (1.0 / 200) <= x) k++; return k; }
/* DDQLearningExecutionAgent.get_observation :280-336 */
static void dq_get_observation(abo_env *s, exec_t *e, double *obs, int *disc) {
  dq_state *d = s->dq;
  int64_t curr = s->now - (s->now % d->h_step); int hi = dq_horizon_index(d, curr);
  e->rem_time = hi >= 0 ? d->n_h - 1 - hi : d->n_h;
  if (e->nkb == 0 || e->nka == 0) { d->error |= DQ_ERR_EMPTY_SIDE; for (int i = 0; i < 6; i++) obs[i] = 0; disc[0] = disc[1] = 0; return; }
  obs[0] = 2 * ((double)e->rem_time / (double)d->n_h) - 1;
  obs[1] = 2 * ((double)e->rem_qty / (double)e->quantity) - 1;
  obs[2] = (double)(e->ka[0].p - e->kb[0].p);
  obs[3] = (double)(e->ka[0].q - e->kb[0].q) / (double)(e->ka[0].q + e->kb[0].q);
  double mid = (double)(e->kb[0].p + e->ka[0].p) / 2;
  if (e->n_pp == e->cap_pp) { e->cap_pp = e->cap_pp ? 2 * e->cap_pp : 1024; e->pp = (double *)realloc(e->pp, 8 * e->cap_pp); }
  e->pp[e->n_pp++] = mid;
  obs[4] = s->now == d->h0 ? 0.0 : log(mid / e->pp[e->n_pp - 2]);
  obs[5] = log(mid / e->pp[0]);
  disc[0] = dq_digitize(obs[0]); disc[1] = dq_digitize(obs[1]);                               /* the grid has TWO dimensions: zip() drops the other four features */
}
static double *dq_exp_row(exec_t *e, int t) { return e->exp + 6 * t; }
/* handle_order_execution :507-548 / handle_order_acceptance :550-576 (after TradingAgent.receiveMessage) */
static void dq_ddqn_order_event(abo_env *s, exec_t *e, const event_t *m) {
  dq_state *d = s->dq;
  int64_t curr = s->now - (s->now % d->h_step);
  if (dq_horizon_index(d, curr) < 0) return;
  double o6[6]; int sp[2]; dq_get_observation(s, e, o6, sp);
  double *row = dq_exp_row(e, e->t - 1); row[3] = sp[0]; row[4] = sp[1]; e->cur_s[0] = sp[0]; e->cur_s[1] = sp[1];
  if (m->kind == ABO_ORDER_EXECUTED) {                                                        /* compute_reward :411-447 (BUY / SELL) */
    double fp = (double)m->order.fill_price, ar = e->arrival;
    double r = (1 - (d->is_buy ? (fp - ar) : (ar - fp)) / ar) * (double)m->order.quantity / (double)e->quantity * 10000;
    if (e->n_rew == e->cap_rew) { e->cap_rew = e->cap_rew ? 2 * e->cap_rew : 1024; e->rew = (double *)realloc(e->rew, 8 * e->cap_rew); }
    e->rew[e->n_rew++] = r; e->step_reward += r; row[5] = r;
  } else row[5] = 0;
}
static void dq_exec_receive(abo_env *s, exec_t *e, const event_t *m) {
  dq_state *d = s->dq;
  int newly = env_ta_receive(s, &e->ta, m);
  if (m->kind == ABO_ORDER_EXECUTED) {                                                        /* TradingAgent.orderExecuted :445-452 */
    for (int i = 0; i < e->n_orders; i++) if (e->orders[i].order_id == m->order.order_id) { if (m->order.quantity >= e->orders[i].quantity) dq_exec_orders_remove(e, i); else e->orders[i].quantity -= m->order.quantity; break; }
  } else if (m->kind == ABO_ORDER_CANCELLED) { for (int i = 0; i < e->n_orders; i++) if (e->orders[i].order_id == m->order.order_id) { dq_exec_orders_remove(e, i); break; } }
  else if (m->kind == ABO_QUERY_SPREAD) { e->nkb = e->nfb; e->nka = e->nfa; memcpy(e->kb, e->fb, sizeof(pq_t) * e->nfb); memcpy(e->ka, e->fa, sizeof(pq_t) * e->nfa); }   /* querySpread :514-537 */
  if (newly) env_set_wakeup(s, e->id, e->ta.mkt_open + (d->h0 - e->ta.mkt_open));             /* getWakeFrequency: start_time - mkt_open */
  if (m->kind == ABO_ORDER_EXECUTED) { e->executed_sum += m->order.quantity; e->n_executed++; e->rem_qty = e->quantity - e->executed_sum; }
  if (!e->is_ddqn) {                                                                          /* ExecutionAgent.receiveMessage :80-86 */
    if (e->rem_qty > 0 && e->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) {
      dq_exec_cancel_all(s, e);
      int hi = dq_horizon_index(d, s->now);                                                   /* placeOrders :107-124 */
      if (hi == d->n_h - 2) dq_place_market(s, e, e->rem_qty);
      else if (hi >= 0 && hi < d->n_h - 2) {
        if (e->nkb == 0 || e->nka == 0) { d->error |= DQ_ERR_EMPTY_SIDE; return; }
        if (hi == 0) { e->arrival = (double)(e->kb[0].p + e->ka[0].p) / 2; e->has_arrival = 1; }
        dq_place_limit(s, e->id, &e->orders, &e->n_orders, &e->cap_orders, (e->sched && hi < e->n_sched) ? e->sched[hi] : e->child_qty, d->is_buy, d->is_buy ? e->ka[0].p : e->kb[0].p);   /* self.schedule[pd.Interval(now, now + 30 s)] */
      }
    }
    return;
  }
  if (m->kind == ABO_ORDER_ACCEPTED || m->kind == ABO_ORDER_EXECUTED) dq_ddqn_order_event(s, e, m);   /* :158-163 */
  else { int hi = dq_horizon_index(d, s->now);
    if (hi >= 0 && hi < d->n_h - 1 && e->rem_qty > 0 && e->state == ST_AWAITING_SPREAD && m->kind == ABO_QUERY_SPREAD) {   /* :164-172 */
      dq_exec_cancel_all(s, e);
      /* place_order :228-245 up to choose_action */
      if (e->nkb == 0 || e->nka == 0) { d->error |= DQ_ERR_EMPTY_SIDE; return; }
      if (hi == 0) { e->arrival = (double)(e->kb[0].p + e->ka[0].p) / 2; e->has_arrival = 1; double o6[6]; dq_get_observation(s, e, o6, e->cur_s); }
      dq_get_observation(s, e, e->obs6, e->sp);
      e->pending = 1;
    } }
}
/* place_order :245-278 from choose_action's return on, then receiveMessage :172 (self.t += 1) */
static void dq_ddqn_resume(abo_env *s, exec_t *e, int a) {
  dq_state *d = s->dq;
  static const double SCALE[6] = { 0.1, 0.5, 1.0, 1.5, 2.0, 2.5 };
  static const double ALLOC[4][4] = { { 0, 0, 0, 0 }, { 1, 0, 0, 0 }, { 0.5, 0.5, 0, 0 }, { 0.34, 0.33, 0.33, 0 } };
  if (a < 0) a = 0; if (a > 23) a = 23;
  int alloc = a / 6; int64_t qty = e->child_qty;                                              /* take_action :367-409 */
  if (e->rem_time == 1) { qty = e->rem_qty; alloc = 0; } else { qty = py_round(SCALE[a % 6] * (double)qty); if (qty < 0) qty = 0; }
  if (e->n_ahist == e->cap_ahist) { e->cap_ahist = e->cap_ahist ? 2 * e->cap_ahist : 1024; e->ahist = (double *)realloc(e->ahist, 8 * e->cap_ahist); }
  e->ahist[e->n_ahist++] = (double)qty / (double)e->quantity;
  if (alloc == 0) dq_place_market(s, e, qty);
  else for (int lv = 0; lv < 4; lv++) {
    int64_t size = py_round(ALLOC[alloc][lv] * (double)qty);
    int n = d->is_buy ? e->nkb : e->nka; if (lv >= n) { d->error |= DQ_ERR_LEVELS; break; }   /* IndexError in the reference */
    int64_t price = d->is_buy ? e->kb[lv].p : e->ka[lv].p;
    if (size != 0) dq_place_limit(s, e->id, &e->orders, &e->n_orders, &e->cap_orders, size, d->is_buy, price);
  }
  if (e->t >= e->cap_exp) { e->cap_exp = e->cap_exp ? 2 * e->cap_exp : 1024; e->exp = (double *)realloc(e->exp, 8 * 6 * e->cap_exp); }
  double *row = dq_exp_row(e, e->t); row[0] = e->cur_s[0]; row[1] = e->cur_s[1]; row[2] = a; row[3] = e->sp[0]; row[4] = e->sp[1]; row[5] = NAN; e->n_exp = e->t + 1;
  e->cur_s[0] = e->sp[0]; e->cur_s[1] = e->sp[1]; e->t++; e->pending = 0;
}

/* stream5 as abo_env_new.  sizes: MomentumAgent.size per agent (config: random_state.randint(1, 10), MomentumAgent.py:42). */
abo_env *abo_dq_new(const int64_t *stream5, int64_t n_rows, int n_mom, const int64_t *mom_sizes, int n_twap, int has_ddqn, int is_buy, int64_t quantity,
                    int64_t h0_ns, int64_t h_step_ns, int n_h, int64_t mom_wake_ns, int trace) {
  if (n_mom < 0 || n_mom > 8 || n_twap < 0 || n_twap > 2 || n_twap + (has_ddqn ? 1 : 0) > 3 || n_h < 3) return NULL;
  abo_env *s = abo_env_new2(stream5, n_rows, (double)quantity, 0, trace, h0_ns + (int64_t)(n_h - 1) * h_step_ns + 600 * NS_PER_S);
  s->q.n = 0;                                                                                 /* rebuild the start-up wakeups for the full agent list */
  dq_state *d = (dq_state *)calloc(1, sizeof(dq_state)); s->dq = d;
  d->n_mom = n_mom; d->n_twap = n_twap; d->has_ddqn = has_ddqn; d->n_ex = n_twap + (has_ddqn ? 1 : 0); d->n_agents = 2 + n_mom + d->n_ex;
  d->h0 = h0_ns; d->h_step = h_step_ns; d->n_h = n_h; d->mom_wake_ns = mom_wake_ns; d->is_buy = is_buy;
  for (int i = 0; i < n_mom; i++) d->mom[i].size = mom_sizes[i];
  for (int k = 0; k < d->n_ex; k++) { exec_t *e = &d->ex[k]; e->id = 2 + n_mom + k; e->is_ddqn = has_ddqn && k == d->n_ex - 1; e->trade = 1; e->state = ST_AWAITING_WAKEUP;
    e->quantity = e->rem_qty = quantity; e->child_qty = e->is_ddqn ? (int64_t)((double)quantity / (double)(n_h - 1)) : (int64_t)((double)quantity / (double)n_h);   /* generate_schedule :499 / twap_agent.py:55 */
    e->kb = (pq_t *)malloc(sizeof(pq_t) * DQ_DEPTH); e->ka = (pq_t *)malloc(sizeof(pq_t) * DQ_DEPTH); e->fb = (pq_t *)malloc(sizeof(pq_t) * DQ_DEPTH); e->fa = (pq_t *)malloc(sizeof(pq_t) * DQ_DEPTH); }
  for (int i = 0; i < d->n_agents; i++) env_set_wakeup(s, i, 0);
  return s;
}
static void dq_free(abo_env *s) {
  dq_state *d = s->dq; if (!d) return;
  for (int i = 0; i < d->n_mom; i++) free(d->mom[i].mids);
  for (int k = 0; k < d->n_ex; k++) { exec_t *e = &d->ex[k]; free(e->orders); free(e->kb); free(e->ka); free(e->fb); free(e->fa); free(e->pp); free(e->exp); free(e->rew); free(e->ahist); free(e->sched); }
  free(d); s->dq = NULL;
}
/* The baseline execution agent k gets a per-bin schedule: what VWAPExecutionAgent.generate_schedule builds from its volume profile (round(profile[bin] * quantity)). */
int abo_dq_set_schedule(abo_env *s, int k, const int64_t *qty, int n) {
  dq_state *d = s ? s->dq : NULL; if (!d || k < 0 || k >= d->n_twap || !qty || n < 1) return -1;
  exec_t *e = &d->ex[k]; free(e->sched); e->sched = (int64_t *)malloc(sizeof(int64_t) * n); memcpy(e->sched, qty, sizeof(int64_t) * n); e->n_sched = n; return 0;
}
/* One decision step.  `action` completes the pending decision (ignored when none is pending: first call).  Runs Kernel.runner's loop
 * (Kernel.py:190-292) until the DDQN agent reaches choose_action again or the loop ends.  out8: the 6 observation features + s' (2 digitised
 * features); trans6: the finalised experience row of the previous tick (s0, s1, a, s'0, s'1, r; r NaN == None); returns 1 when a decision is
 * pending, 0 when the simulation ended (*done = 1). */
int abo_dq_step(abo_env *s, int action, double *out8, double *trans6, double *reward, int *done) {
  dq_state *d = s->dq; exec_t *dd = d->has_ddqn ? &d->ex[d->n_ex - 1] : NULL;
  if (dd && dd->pending) { dq_ddqn_resume(s, dd, action); s->agent_time[dd->id] += s->comp_delay[dd->id]; }
  if (dd) dd->step_reward = 0;
  int paused = 0;
  while (!paused && s->q.n > 0 && s->now <= s->stop_time) {
    event_t ev; heap_pop(&s->q, &ev); s->now = ev.t; s->ttl++;
    s->pop_hash = fnv_mix(fnv_mix(fnv_mix(fnv_mix(s->pop_hash, ev.t), ev.recipient), ev.type), ev.uniq);
    if (s->ttl % 1000 == 0) { if (s->n_ckpt == s->cap_ckpt) { s->cap_ckpt = s->cap_ckpt ? s->cap_ckpt * 2 : 256; s->ckpt = (uint64_t *)realloc(s->ckpt, 8 * s->cap_ckpt); } s->ckpt[s->n_ckpt++] = s->pop_hash; }
    if (s->trace & ABO_TRACE_POPS) { int64_t row[5] = { ev.t, ev.recipient, ev.type, ev.uniq, ev.kind }; ib_push(&s->pops, row, 5); }
    int a = ev.recipient;
    if (s->agent_time[a] > s->now) { ev.t = s->agent_time[a]; env_put(s, &ev); continue; }
    s->agent_time[a] = s->now;
    exec_t *e = dq_exec(s, a);
    if (ev.type == ABO_T_WAKEUP) { if (a == 1) replay_wakeup(s); else if (e) dq_exec_wakeup(s, e); else if (a >= 2) dq_mom_wakeup(s, a); }
    else { if (a == 0) env_exch_receive(s, &ev); else if (a == 1) replay_receive(s, &ev); else if (e) dq_exec_receive(s, e, &ev); else dq_mom_receive(s, a, &ev); }
    if (dd && dd->pending) { paused = 1; break; }                                             /* agent_time is advanced when the handler resumes */
    s->agent_time[a] += s->comp_delay[a];
  }
  *done = !paused;
  for (int i = 0; i < 8; i++) out8[i] = 0; for (int i = 0; i < 6; i++) trans6[i] = NAN; *reward = dd ? dd->step_reward : 0;
  if (dd) {
    if (paused) { for (int i = 0; i < 6; i++) out8[i] = dd->obs6[i]; out8[6] = dd->sp[0]; out8[7] = dd->sp[1]; }
    int tprev = dd->t - 1; if (tprev >= 0) for (int i = 0; i < 6; i++) trans6[i] = dq_exp_row(dd, tprev)[i];
  }
  return paused;
}
int abo_dq_error(abo_env *s) { return s->dq ? s->dq->error : 0; }
/* which: 0 price_path, 1 experience rows (x6), 2 step_reward_hist, 3 action_hist */
int64_t abo_dq_series(abo_env *s, int which, const double **v) {
  dq_state *d = s->dq; exec_t *e = &d->ex[d->n_ex - 1];
  switch (which) { case 0: *v = e->pp; return e->n_pp; case 1: *v = e->exp; return e->n_exp; case 2: *v = e->rew; return e->n_rew; default: *v = e->ahist; return e->n_ahist; }
}
/* per trader (ids 1..n-1): rows (id, shares, cash, last_trade, open orders) */
void abo_dq_holdings(abo_env *s, int64_t *out5) {
  dq_state *d = s->dq;
  for (int id = 1; id < d->n_agents; id++) { const tagent_t *a; int64_t no = 0; exec_t *e = dq_exec(s, id);
    if (id == 1) { a = &s->ra; no = s->ra_orders.n; } else if (e) { a = &e->ta; no = e->n_orders; } else { a = &d->mom[id - 2].ta; no = -1; }
    int64_t *r = out5 + 5 * (id - 1); r[0] = id; r[1] = a->shares; r[2] = a->cash; r[3] = a->has_last_trade ? a->last_trade : 0; r[4] = no; }
}
/* execution agent k (0 .. n_exec-1): rem_qty, arrival price, executed orders, remaining_time, t */
void abo_dq_exec_final(abo_env *s, int k, double *out5) { exec_t *e = &s->dq->ex[k]; out5[0] = (double)e->rem_qty; out5[1] = e->arrival; out5[2] = (double)e->n_executed; out5[3] = e->rem_time; out5[4] = e->t; }
