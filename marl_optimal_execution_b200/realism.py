"""Minute bars and the reference's seven stylized-fact metrics for whole batches of environments.

The reference logs LAST_TRADE events at the exchange (util/OrderBook.py:131-141), resamples them to minute bars after the run
(realism/realism_utils.py:22-44: `close` = last trade price of the minute, forward filled; `volume` = traded shares of the minute) and
computes seven metrics on them (realism/metrics/*.py).  A batched simulation does not need an event log for that: stepping every
environment to each minute boundary (one abx_run_kernel launch per minute) and reading `last_trade` and the fill counter from the
per-environment counters gives exactly those two columns.  The metrics below are vectorised restatements (numpy, over [n_envs, n_minutes]
arrays) of the reference classes; tests/golden/realism_metrics.npz holds the reference's own outputs for them.
"""
import numpy as np

NS = 10 ** 9


def minute_bars(sim, n_minutes=390, order_size=None, stream=None):
    """Run `sim` (a reset BatchedSim) minute by minute from the market open; returns (close [n_envs, n_minutes] float64 cents,
    volume [n_envs, n_minutes] shares).  A trade stamped exactly on a boundary belongs to the next bar, like pandas' left-closed
    resample bins.  `volume` counts fills x order_size: exact for the sparse_zi configs, where every order has the same size
    (ZeroIntelligenceAgent.py:308); bars before an environment's first trade hold the opening price where the reference has NaN."""
    cfg = sim.cfg
    size = int(cfg.order_size if order_size is None else order_size)
    t0 = int(cfg.mkt_open_ns)
    close = np.zeros((sim.n_envs, n_minutes))
    volume = np.zeros((sim.n_envs, n_minutes))
    prev = None
    for k in range(n_minutes):
        sim.run(t0 + (k + 1) * 60 * NS - 1, stream=stream)
        st = sim.stats(stream=stream)
        fills = st["fills"].astype(np.int64)
        if prev is None:
            prev = np.zeros_like(fills)
        close[:, k] = st["last_trade"]
        volume[:, k] = (fills - prev) * size
        prev = fills
    return close, volume


def minutely_returns(close):
    """MinutelyReturns.compute (realism/metrics/minutely_returns.py:9-13): diff of log close."""
    return np.diff(np.log(np.asarray(close, dtype=np.float64)), axis=-1)


def _corr(a, b):
    """Pearson correlation along the last axis (pandas Series.corr / np.corrcoef)."""
    a = a - a.mean(axis=-1, keepdims=True)
    b = b - b.mean(axis=-1, keepdims=True)
    den = np.sqrt((a * a).sum(axis=-1) * (b * b).sum(axis=-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (a * b).sum(axis=-1) / den


def autocorrelation(close, lag=1, window=30):
    """Autocorrelation.compute (autocorrelation.py:15-18): lag-`lag` autocorrelation of the returns over every full centred window."""
    r = minutely_returns(close)
    window = min(int(window), r.shape[-1])                 # a run shorter than the window (rmsc03's 15 minutes) is one window
    w = np.lib.stride_tricks.sliding_window_view(r, window, axis=-1)
    return _corr(w[..., lag:], w[..., :-lag])


def _bin_last(n, minutes, first_minute_of_day):
    """Row indices that pandas' resample("{minutes}T").last() keeps: bins are aligned to midnight, so with bars starting at 09:30
    (minute 570 of the day) 4-minute bins start at 09:28 and the first one holds two bars."""
    j = np.arange(n)
    keep = (first_minute_of_day + j + 1) % minutes == 0
    keep[n - 1] = True
    return j[keep]


def kurtosis(close, intervals=4, first_minute_of_day=570):
    """Kurtosis.compute (kurtosis.py:12-18): excess kurtosis (scipy.stats.kurtosis defaults: Fisher, biased) of the returns at 1..4 minute scales."""
    close = np.asarray(close, dtype=np.float64)
    out = []
    for i in range(1, intervals + 1):
        r = minutely_returns(close[..., _bin_last(close.shape[-1], i, first_minute_of_day)])
        d = r - r.mean(axis=-1, keepdims=True)
        m2, m4 = (d ** 2).mean(axis=-1), (d ** 4).mean(axis=-1)
        with np.errstate(invalid="ignore", divide="ignore"):
            out.append(m4 / m2 ** 2 - 3.0)
    return np.stack(out, axis=-1)


def aggregation_normality(close, minutes=10, first_minute_of_day=570):
    """AggregationNormality.compute (aggregation_normality.py:9-11): returns of the 10-minute bars."""
    close = np.asarray(close, dtype=np.float64)
    return minutely_returns(close[..., _bin_last(close.shape[-1], minutes, first_minute_of_day)])


def volatility_clustering(close, lags=10, mode="abs"):
    """VolatilityClustering.compute (volatility_clustering.py:19-25): autocorrelation of |r| (or r^2) at lags 1..10."""
    r = minutely_returns(close)
    v = np.abs(r) if mode == "abs" else r ** 2
    return np.stack([_corr(v[..., lag:], v[..., :-lag]) for lag in range(1, lags + 1)], axis=-1)


def returns_volatility_correlation(close):
    """ReturnsVolatilityCorrelation.compute (returns_volatility_correlation.py:10-13)."""
    r = minutely_returns(close)
    return _corr(r, np.abs(r))


def volume_volatility_correlation(close, volume):
    """VolumeVolatilityCorrelation.compute (volume_volatility_correlation.py:10-13)."""
    r = minutely_returns(close)
    return _corr(np.asarray(volume, dtype=np.float64)[..., 1:], np.abs(r))


def all_metrics(close, volume):
    return {"returns": minutely_returns(close), "autocorr": autocorrelation(close), "kurtosis": kurtosis(close), "aggnorm": aggregation_normality(close),
            "volclust": volatility_clustering(close), "retvol": returns_volatility_correlation(close), "volvol": volume_volatility_correlation(close, volume)}


# ---------------------------------------------------------------------------------------------------------------------------------------
# Device-side event ring (cfg.event_ring_cap > 0): the exchange's BEST_BID / BEST_ASK / LAST_TRADE log lines (util/OrderBook.py:114-141) and
# the order arrivals, reduced ON THE GPU (torch) to minute bars, the seven return metrics and the order-flow stylized facts of
# realism/order_flow_stylized_facts.py:84-103,176-221 -- one pass over the ring, no per-minute launches, exact volumes for every population.
# ---------------------------------------------------------------------------------------------------------------------------------------
EV_ORDER, EV_BEST_BID, EV_BEST_ASK, EV_LAST_TRADE = 0, 1, 2, 3


def events(sim, device=None, stream=None):
    """The rings of a BatchedSim as torch tensors on the simulator's GPU: (t_ns int64 [n_envs, cap], kind int64, a int64, b int64, valid bool), in
    chronological order per environment (the ring is unrolled; entries older than the capacity are gone and `valid` is False for unused slots)."""
    import ctypes as C
    import torch
    from . import _lib
    cap = int(sim.cfg.event_ring_cap)
    if cap <= 0:
        raise ValueError("the simulator was created with event_ring_cap == 0")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    raw = torch.empty(sim.n_envs, cap, 4, dtype=torch.int32, device=dev)
    cnt = torch.empty(sim.n_envs, dtype=torch.int32, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if stream is None else stream
    _lib.check(sim._L, sim._L.abx_sim_events_device(sim._h, C.c_void_p(raw.data_ptr()), C.c_void_p(cnt.data_ptr()), sp), "abx_sim_events_device")
    return unroll_events(raw, cnt)


def unroll_events(raw, counts):
    """raw int32 [n_envs, cap, 4] ring slots + counts [n_envs] -> chronological tensors (works on CPU tensors too: the CPU suite feeds it the emulation's ring)."""
    import torch
    n, cap, _ = raw.shape
    cnt = counts.to(torch.int64) & 0xFFFFFFFF
    start = torch.where(cnt > cap, cnt % cap, torch.zeros_like(cnt))                  # oldest surviving slot
    idx = (start[:, None] + torch.arange(cap, device=raw.device)[None, :]) % cap
    r = torch.gather(raw.to(torch.int64) & 0xFFFFFFFF, 1, idx[:, :, None].expand(n, cap, 4))
    t = r[..., 0] | ((r[..., 1] & 0x0FFFFFFF) << 32)
    kind = r[..., 1] >> 28
    sgn = lambda v: torch.where(v >= 2 ** 31, v - 2 ** 32, v)                         # noqa: E731  int32 payloads
    valid = torch.arange(cap, device=raw.device)[None, :] < torch.clamp(cnt, max=cap)[:, None]
    return t, kind, sgn(r[..., 2]), sgn(r[..., 3]), valid


def bars_from_events(t, kind, a, b, valid, mkt_open_ns, n_minutes=390, open_price=None):
    """Minute bars like realism/realism_utils.py:22-44 from the LAST_TRADE events: close = last trade price of the minute, forward filled (the price
    before the first trade: `open_price`, where the reference has NaN); volume = shares traded in the minute (exact: the LAST_TRADE quantities).
    Returns (close, volume) float64 [n_envs, n_minutes] on the events' device."""
    import torch
    n, cap = t.shape
    m = ((t - int(mkt_open_ns)) // (60 * NS)).clamp(min=-1, max=n_minutes)           # left-closed bins: a trade on a boundary opens the next bar
    tr = valid & (kind == EV_LAST_TRADE) & (m >= 0) & (m < n_minutes)
    mi = torch.where(tr, m, torch.zeros_like(m))
    volume = torch.zeros(n, n_minutes, dtype=torch.float64, device=t.device).scatter_add_(1, mi, torch.where(tr, b, torch.zeros_like(b)).to(torch.float64))
    pos = torch.arange(cap, device=t.device)[None, :].expand(n, cap)
    last = torch.full((n, n_minutes), -1, dtype=torch.int64, device=t.device).scatter_reduce_(1, mi, torch.where(tr, pos, torch.full_like(pos, -1)), reduce="amax")
    has = last >= 0
    px = torch.gather(a, 1, last.clamp(min=0)).to(torch.float64)
    # forward fill: index of the latest minute with a trade at or before each minute
    mm = torch.arange(n_minutes, device=t.device)[None, :].expand(n, n_minutes)
    src = torch.cummax(torch.where(has, mm, torch.full_like(mm, -1)), dim=1).values
    close = torch.gather(px, 1, src.clamp(min=0))
    if open_price is not None:
        close = torch.where(src >= 0, close, torch.full_like(close, float(open_price)))
    return close, volume


def order_flow_facts(t, kind, valid, mkt_open_ns, mkt_close_ns, binwidth_s=1):
    """realism/order_flow_stylized_facts.py: interarrival times of the order stream (:84-103, seconds between consecutive order arrivals) and the number
    of arrivals per `binwidth_s`-second bin (:176-221), per environment.  Returns dict(interarrival_mean, interarrival_std, interarrival_log10_hist
    [n_envs, 12] over 1e-9..1e3 s, n_orders, bin_counts int64 [n_envs, n_bins], bin_count_mean, bin_count_var)."""
    import torch
    n, cap = t.shape
    od = valid & (kind == EV_ORDER)
    tf = t.to(torch.float64) / 1e9
    # consecutive arrivals: sort order rows to the front (stable: chronological order is kept)
    key = torch.where(od, torch.zeros_like(t), torch.ones_like(t))
    order = torch.argsort(key, dim=1, stable=True)
    ts = torch.gather(tf, 1, order)
    k = od.sum(dim=1)
    pair = torch.arange(cap - 1, device=t.device)[None, :] < (k - 1).clamp(min=0)[:, None]
    d = (ts[:, 1:] - ts[:, :-1]) * pair
    cntp = pair.sum(dim=1).clamp(min=1).to(torch.float64)
    mean = d.sum(dim=1) / cntp
    var = (((d - mean[:, None]) * pair) ** 2).sum(dim=1) / cntp
    lg = torch.log10(d.clamp(min=1e-9))
    hb = ((lg + 9.0).floor().clamp(min=0, max=11)).to(torch.int64)
    hist = torch.zeros(n, 12, dtype=torch.int64, device=t.device).scatter_add_(1, torch.where(pair, hb, torch.zeros_like(hb)), pair.to(torch.int64))
    n_bins = int((int(mkt_close_ns) - int(mkt_open_ns)) // (binwidth_s * NS))
    bi = ((t - int(mkt_open_ns)) // (binwidth_s * NS))
    inb = od & (bi >= 0) & (bi < n_bins)
    bins = torch.zeros(n, n_bins, dtype=torch.int64, device=t.device).scatter_add_(1, torch.where(inb, bi, torch.zeros_like(bi)), inb.to(torch.int64))
    bf = bins.to(torch.float64)
    return {"interarrival_mean": mean, "interarrival_std": var.sqrt(), "interarrival_log10_hist": hist, "n_orders": k, "bin_counts": bins,
            "bin_count_mean": bf.mean(dim=1), "bin_count_var": bf.var(dim=1, unbiased=False)}


def spread_facts(t, kind, a, valid):
    """Quoted spread from the BEST_BID / BEST_ASK lines: per environment the mean and the last value of (best ask - best bid) over the events at which both
    sides are known (each event updates one side; the other side keeps its last logged value)."""
    import torch
    n, cap = t.shape
    pos = torch.arange(cap, device=t.device)[None, :].expand(n, cap)
    def last_seen(k):
        m = valid & (kind == k)
        src = torch.cummax(torch.where(m, pos, torch.full_like(pos, -1)), dim=1).values
        return torch.gather(a, 1, src.clamp(min=0)), src >= 0
    bid, hb = last_seen(EV_BEST_BID)
    ask, ha = last_seen(EV_BEST_ASK)
    upd = valid & ((kind == EV_BEST_BID) | (kind == EV_BEST_ASK)) & hb & ha
    sp = (ask - bid).to(torch.float64) * upd
    cnt = upd.sum(dim=1).clamp(min=1).to(torch.float64)
    lastpos = torch.where(upd, pos, torch.full_like(pos, -1)).amax(dim=1)
    return {"spread_mean": sp.sum(dim=1) / cnt, "spread_last": torch.gather(ask - bid, 1, lastpos.clamp(min=0)[:, None])[:, 0], "n_quotes": upd.sum(dim=1)}


def stylized_facts_gpu(sim, n_minutes=390, binwidth_s=1):
    """One pass over the device event ring of a finished (or running) BatchedSim: minute bars, the reference's seven return metrics and the order-flow /
    spread statistics, all reduced on the GPU.  The seven metrics are the numpy restatements above applied to the [n_envs, 390] bars (tiny)."""
    t, kind, a, b, valid = events(sim)
    cfg = sim.cfg
    close, volume = bars_from_events(t, kind, a, b, valid, cfg.mkt_open_ns, n_minutes, open_price=cfg.r_bar)
    out = {"close": close, "volume": volume}
    out.update(order_flow_facts(t, kind, valid, cfg.mkt_open_ns, cfg.mkt_close_ns, binwidth_s))
    out.update(spread_facts(t, kind, a, valid))
    out["metrics"] = all_metrics(close.cpu().numpy(), volume.cpu().numpy())
    return out
