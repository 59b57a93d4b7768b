#!/usr/bin/env python3
"""Golden vectors for the seven stylized-fact metrics of the reference (realism/metrics/*.py), build container only.

  python tools/record_reference_realism.py tests/golden/realism_metrics.npz

Imports the UNMODIFIED metric classes from /root/reference/realism (matplotlib is absent from this image and only used by their
`visualize` methods, so a stub module stands in), feeds them seeded synthetic minute bars (390 rows of close / volume, the frame
realism_utils.get_trades :22-44 builds from the exchange's LAST_TRADE log) and stores inputs and outputs.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")


def main():
    out = sys.argv[1]
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    r0 = pd.DataFrame.resample                                    # pandas >= 3 dropped the 'T' alias the reference uses ("{}T".format(i), "10T")

    def resample(self, rule, *a, **k):
        if isinstance(rule, str) and rule.endswith("T"):
            rule = rule[:-1] + "min"
        return r0(self, rule, *a, **k)

    pd.DataFrame.resample = resample
    sys.path.insert(0, os.path.join(REF, "realism"))
    from metrics.aggregation_normality import AggregationNormality
    from metrics.autocorrelation import Autocorrelation
    from metrics.kurtosis import Kurtosis
    from metrics.minutely_returns import MinutelyReturns
    from metrics.returns_volatility_correlation import ReturnsVolatilityCorrelation
    from metrics.volatility_clustering import VolatilityClustering
    from metrics.volume_volatility_correlation import VolumeVolatilityCorrelation

    rs = np.random.RandomState(20260101)
    n_series, n = 6, 390
    idx = pd.date_range("2019-06-28 09:30:00", periods=n, freq="1min")
    closes, vols, res = [], [], {k: [] for k in ("returns", "autocorr", "kurtosis", "aggnorm", "volclust", "retvol", "volvol")}
    for s in range(n_series):
        vol = 0.0004 * (1 + 0.8 * np.abs(np.sin(np.arange(n) / 17.0 + s)))               # heteroskedastic: volatility clustering is non-trivial
        r = rs.standard_t(4, size=n) * vol
        close = np.round(100000 * np.exp(np.cumsum(r)))                                 # integer cents like OrderBook.last_trade
        if s == 3:
            close[50:60] = close[49]                                                    # a stale stretch (ffill of minutes without trades)
        volume = np.round(100 * (5 + 40 * np.abs(r) / vol.mean() + rs.poisson(10, n)))
        df = pd.DataFrame({"open": close, "high": close, "low": close, "close": close, "volume": volume}, index=idx)
        closes.append(close); vols.append(volume)
        res["returns"].append(MinutelyReturns().compute(df))
        res["autocorr"].append(Autocorrelation().compute(df))
        res["kurtosis"].append(Kurtosis().compute(df)[0])
        res["aggnorm"].append(AggregationNormality().compute(df))
        res["volclust"].append(VolatilityClustering().compute(df)[0])
        res["retvol"].append(ReturnsVolatilityCorrelation().compute(df)[0])
        res["volvol"].append(VolumeVolatilityCorrelation().compute(df)[0])
    np.savez_compressed(out, close=np.array(closes), volume=np.array(vols), **{k: np.array(v, dtype=np.float64) for k, v in res.items()})
    print("recorded", {k: np.array(v).shape for k, v in res.items()}, "->", out)


if __name__ == "__main__":
    main()
