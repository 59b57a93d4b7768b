"""Import shims for the reference: pandas>=1 moved json_normalize, pandas 2 dropped Timedelta.delta (total nanoseconds)."""
import pandas
import pandas.io.json

if not hasattr(pandas.io.json, "json_normalize"):
    pandas.io.json.json_normalize = pandas.json_normalize

if not hasattr(pandas.Timedelta, "delta"):
    pandas.Timedelta.delta = property(lambda self: self.value)
