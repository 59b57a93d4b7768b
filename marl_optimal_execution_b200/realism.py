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
