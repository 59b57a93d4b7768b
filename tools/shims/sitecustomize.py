"""Import shim for the reference: pandas>=1 moved json_normalize."""
import pandas
import pandas.io.json

if not hasattr(pandas.io.json, "json_normalize"):
    pandas.io.json.json_normalize = pandas.json_normalize
