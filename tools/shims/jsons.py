"""Import shim for the reference: `jsons` is absent from this image.

The reference only calls `jsons.dump(order, strip_privates=True)` to build log payloads
(agent/ExchangeAgent.py:165,482; agent/TradingAgent.py:346); the value never feeds back
into simulation state.
"""


def dump(obj, strip_privates=True, **kw):
    try:
        return dict(vars(obj))
    except TypeError:
        return obj
